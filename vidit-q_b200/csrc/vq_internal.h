// Internal declarations shared by the .cu translation units of libviditq_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/viditq_b200.h"

#include <cstdlib>
#include <utility>

namespace vq {
int num_sms();

// Programmatic dependent launch (opt-in, VQ_PDL=1): every kernel of this library starts with `griddepcontrol.wait`
// (nothing before it touches global memory) followed by `griddepcontrol.launch_dependents`, so the next kernel's CTA
// scheduling and barrier/TMEM set-up may overlap the tail of this one.  Measured on B200 inside the captured denoise step:
// no gain (46.1 ms with, 45.7 ms without — a replayed graph already runs its kernels back to back, CUPTI shows no gaps),
// hence off by default; without the launch attribute the two instructions are no-ops.
inline bool pdl_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("VQ_PDL");
    return e && e[0] == '1';
  }();
  return on;
}

#ifdef __CUDACC__
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// first statements of every kernel, executed by every thread (a CTA that skipped the wait could let the grid finish,
// and release ITS dependents, before the grid it depends on has finished)
__device__ __forceinline__ void grid_dep_sync() {
  grid_dep_wait();
  grid_dep_launch();
}
#endif

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

int make_u8_kmajor_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch,
                        uint32_t box_rows);
}  // namespace vq
