#!/bin/bash
# Bring-up round for vq_attn_spatial: self-test in every bisection mode + timing, then the GPU parity suite.
mkdir -p gpurun_out
timeout 120 tools/attn_selftest --time > gpurun_out/attn_selftest.log 2>&1; echo "attn_selftest rc=$?" >> gpurun_out/attn_selftest.log
cat gpurun_out/attn_selftest.log
if [ -n "$VQ_ATTN_EXTRA" ]; then bash -c "$VQ_ATTN_EXTRA"; fi
