// Internal declarations shared by the .cu translation units of libviditq_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/viditq_b200.h"

#include <cstdlib>
#include <utility>

namespace vq {
constexpr int kMaxDevices = 64;
int current_device();   // ordinal of the calling thread's current device, clamped to [0, kMaxDevices)
int num_sms();          // SM count of the current device

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
// nn.GELU(approximate="tanh") on a pair: 0.5 x (1 + tanh(u)) == x / (1 + exp(-2u)), u = k0 x (1 + k1 x^2).
// MUFU.EX2 / MUFU.RCP (~1e-7 relative) are far below one fp16 ulp; everything else is packed fp32 (FFMA2 / FMUL2).
__device__ __forceinline__ float2 gelu_tanh_pair(float2 x) {
  // exp(-2u) = 2^(c u), c = -2 log2(e), folded into the cubic's coefficients: w = x (A + B x^2)
  const float A = -2.0f * 1.4426950408889634f * 0.7978845608028654f;
  const float B = -2.0f * 1.4426950408889634f * 0.7978845608028654f * 0.044715f;
  const float2 x2 = __fmul2_rn(x, x);
  const float2 a = __ffma2_rn(x2, make_float2(B, B), make_float2(A, A));
  const float2 w = __fmul2_rn(x, a);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(w.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(w.y));
  const float2 d = __fadd2_rn(make_float2(e0, e1), make_float2(1.0f, 1.0f));
  float r0, r1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d.y));
  return __fmul2_rn(x, make_float2(r0, r1));
}

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

// cross attention on the tcgen05 flash-attention kernel of vq_attn_spatial.cu (N a multiple of 256, prompts <= 128 rows)
int attn_cross_tc(const void* q, const void* kv, void* out, const int* kv_start, const int* kv_len, int B, int N, int H,
                  int head_dim, long long kv_rows, float scale, void* stream);
int make_u8_kmajor_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch,
                        uint32_t box_rows);
// QuantLinear with the quantise pass inside the GEMM kernel (vq_gemm_w8a8.cu, QPRO): see vq_linear_w8a8, policy mode 2
int gemm_w8a8_qpro(const void* x, const void* shift, const void* scale, int rows_per_mod, const void* smooth, int n_bits,
                   uint8_t* codes, void* delta, void* zp, int32_t* rowsum, uint32_t* sync, const uint8_t* w_codes,
                   const VqColParam* col, int M, int N, int K, int epi, const void* res, int ldr, const void* gate,
                   int rows_per_gate, void* out, int ldo, uint32_t* status, cudaStream_t st);
int make_u8_tmap_ex(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, uint32_t box_cols,
                    uint32_t box_rows, bool swizzle128);
// rows x cols fp16 matrix (row pitch ld elements), box 32 rows x 32 columns, SWIZZLE_64B: the epilogue staging sub-tile
int make_f16_out_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld);
}  // namespace vq
