"""Build libviditq_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

nvcc cross-compiles for sm_100a on a machine without a GPU; the resulting .so sits next to this file so that it travels
with the repo snapshot to the GPU box (it is git-ignored, not gpurun-ignored).  Translation units are compiled to objects
in parallel (build/ is git-ignored) and only re-compiled when they or a header changed.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libviditq_b200.so")
# tuning / measurement build: the bisection epilogues of the GEMM (mainloop-only etc.) exist only here
DEBUG_LIB = "libviditq_b200_dbg.so"
DEBUG_DEFINES = ("VQ_DEBUG_EPI",)
SOURCES = ["vq_gemm_w8a8.cu", "vq_linear.cu", "vq_quant.cu", "vq_attention.cu", "vq_attn_spatial.cu", "vq_attn_i8.cu", "vq_sampler.cu",
           "vq_embed.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-cudart", "shared",
]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: cannot build libviditq_b200.so")
    return cand


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    return hs + [os.path.join(HERE, "..", "include", "viditq_b200.h")]


def needs_build(target=LIB):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + _headers()
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, defines=()):
    """out / defines: build a variant next to the product library (e.g. -DVQ_DEBUG_EPI for the measurement build)."""
    target = LIB if out is None else os.path.join(HERE, out)
    if not force and not needs_build(target):
        return target
    tag = "prod" if not defines else "_".join(d.replace("=", "-") for d in defines)
    objdir = os.path.join(HERE, "build", tag)
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    hdr_time = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(objdir, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_time):
            jobs.append((s, [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines] +
                         ["-c", src, "-o", obj]))

    def run(job):
        return job[0], subprocess.run(job[1], capture_output=True, text=True)
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        results = list(pool.map(run, jobs))
    for name, res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError(f"nvcc failed on {name}")
        if verbose:
            sys.stderr.write(f"==== {name}\n" + res.stderr)
    objs = [os.path.join(objdir, s[:-3] + ".o") for s in SOURCES]
    res = subprocess.run([nvcc, "-shared", "-cudart", "shared", "-o", target] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("linking libviditq_b200.so failed")
    return target


def build_debug(force=False):
    return build(force=force, out=DEBUG_LIB, defines=DEBUG_DEFINES)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--debug" in sys.argv:
        print(build_debug(force="--force" in sys.argv))
