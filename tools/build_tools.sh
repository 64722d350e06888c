#!/bin/bash
# Builds the on-device self-test binaries next to their sources (git-ignored; they travel to the GPU box with gpurun).
# They link the MEASUREMENT build of the library (libviditq_b200_dbg.so, -DVQ_DEBUG_EPI): the GEMM self-test times the
# bisection epilogues (mainloop only, + TMEM loads, ...) that the product library does not contain.
set -e
cd "$(dirname "$0")/.."
python vidit-q_b200/build.py --debug > /dev/null
for t in gemm_selftest attn_selftest; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/$t tools/$t.cu \
       -Lvidit-q_b200 -l:libviditq_b200_dbg.so -lcuda -Xlinker -rpath -Xlinker '$ORIGIN/../vidit-q_b200'
done
ls -la tools/gemm_selftest tools/attn_selftest
