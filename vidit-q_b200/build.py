"""Build libviditq_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

nvcc cross-compiles for sm_100a on a machine without a GPU; the resulting .so sits next to this file so that it travels
with the repo snapshot to the GPU box (it is git-ignored, not gpurun-ignored).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libviditq_b200.so")
SOURCES = ["vq_gemm_w8a8.cu", "vq_quant.cu", "vq_attention.cu", "vq_attn_spatial.cu", "vq_sampler.cu", "vq_embed.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared",
]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: cannot build libviditq_b200.so")
    return cand


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "viditq_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: build a tuning variant next to the product library (e.g. -DVQ_EPI_WARPS=8 for A/B timing)."""
    if out is None and not force and not needs_build():
        return LIB
    target = LIB if out is None else os.path.join(HERE, out)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines] + \
        [os.path.join(CSRC, s) for s in SOURCES] + ["-o", target]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libviditq_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
