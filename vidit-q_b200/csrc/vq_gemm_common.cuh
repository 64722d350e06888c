// Tile constants, kernel arguments and the dequantising epilogue pieces shared by the two tcgen05 INT8 kernels:
// vq_gemm_w8a8_kernel (vq_gemm_w8a8.cu: codes in, TMA-staged A and B) and vq_linear_fused_kernel (vq_linear.cu: fp16 in,
// producer warps quantise the activation panel into the shared-memory A operand).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "vq_ptx.cuh"
#include "vq_internal.h"

namespace vq {

constexpr int BM = 128;
constexpr int BN = 192;
constexpr int BK = 128;  // bytes == u8 elements per K block (one 128B swizzle row)
constexpr int UMMA_K = 32;
constexpr int STAGES = 4;        // single-CTA tiles: 4 x (16 KB A + 24 KB B)
constexpr int PAIR_STAGES = 6;   // CTA pairs stage half a B tile each: 6 x (16 KB A + 12 KB B)
constexpr int MAX_STAGES = 6;
constexpr int B_PAIR_STAGE_BYTES = (BN / 2) * BK;   // cta_group::2: each CTA of the pair stages half of the B tile
constexpr int A_STAGE_BYTES = BM * BK;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK;  // 24 KB
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 256;  // TMEM column stride between accumulator stages
constexpr int TMEM_COLS = 512;
constexpr int NUM_EPI_WARPS = 8;                      // warp%4 = TMEM lane quarter, (warp-4)/4 = column half
constexpr int GEMM_THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int EPI_COLS = BN / 2;                      // 96 output columns per epilogue warp per tile
constexpr int EPI_CHUNK = 32;                         // columns per TMEM load / staging sub-tile (64 B of fp16 per row)
constexpr int EPI_NCHUNK = EPI_COLS / EPI_CHUNK;      // 3
constexpr int EPI_BUF_BYTES = 32 * EPI_CHUNK * 2;     // one sub-tile: 32 rows x 64 B, SWIZZLE_64B
constexpr int EPI_STAGING_BYTES = NUM_EPI_WARPS * EPI_NCHUNK * EPI_BUF_BYTES;   // one 32 x 96 strip per warp
constexpr int COLBUF_BYTES = 2 * BN * 16;                // per-tile {c1, zw, dw, bias} records, double-buffered
constexpr int OPERAND_BYTES_SINGLE = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES);
constexpr int OPERAND_BYTES_PAIR = PAIR_STAGES * (A_STAGE_BYTES + B_PAIR_STAGE_BYTES);
constexpr int OPERAND_BYTES = OPERAND_BYTES_PAIR > OPERAND_BYTES_SINGLE ? OPERAND_BYTES_PAIR : OPERAND_BYTES_SINGLE;
constexpr int SMEM_BYTES = OPERAND_BYTES + EPI_STAGING_BYTES + COLBUF_BYTES + 512 + 1024;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");

struct GemmArgs {
  int M, N, K;
  const __half* a_delta;    // [M] per-token step size (fp16, as the reference's DynamicActQuantizer.delta)
  const __half* a_zp;       // [M] per-token zero point (integer valued fp16)
  const int32_t* a_rowsum;  // [M] sum_k xq[m,k]
  int a_period;             // delta/zp row index = m % a_period (token statistics pooled over the batch, Q1)
  const VqColParam* col;    // [N] {c1, zw, dw, bias}
  __half* out;              // [M, ldo]
  int ldo;
  int epi;                  // VQ_EPI_*
  const __half* res;        // [M, ldr] residual (VQ_EPI_GATE_RESIDUAL)
  int ldr;
  const __half* gate;       // [M / rows_per_gate, N]
  int rows_per_gate;
  uint64_t store_policy;    // L2 cache policy of the output stores (kEvictFirst unless VQ_STORE_POLICY=normal)
  int group_m;              // rasterisation: m-panels per group (tile_to_mn)
  // ---- quantise-producer variant (QPRO): the activation codes this GEMM consumes are produced INSIDE the kernel by two
  // quantiser warpgroups running ahead of the MMAs (vq_gemm_w8a8.cu); a_codes / a_delta / a_zp / a_rowsum are then scratch
  const __half* qx;         // [M, K] fp16 source rows (K = 1152)
  const __half* q_shift;    // LayerNorm + modulate variant: [M / q_rows_per_mod, K]
  const __half* q_scale;
  const __half* q_smooth;   // [K] or null
  int q_rows_per_mod;
  float q_qmax;
  uint8_t* q_codes;         // scratch [M, K]: what tmap_a describes
  uint32_t* q_sync;         // [0] next row of the work queue; [1 + m] rows of m-panel m (TILE_M rows) quantised so far
  uint32_t* q_status;
};

constexpr int QP_THREADS = 640;           // 4 control warps + 8 epilogue warps + 8 quantiser warps
constexpr int QP_ROWS = 2;                // rows a quantiser warp keeps in flight
constexpr int QP_GRAB = 8;                // rows it takes from the work queue per atomic (one flag update per grab)
constexpr int QP_REGS_CONTROL = 40;       // setmaxnreg budgets per warpgroup: 128 x 40 + 256 x 144 + 256 x 72 = 60416 <= 640 x 96
constexpr int QP_REGS_EPILOGUE = 144;
constexpr int QP_REGS_QUANT = 72;

// Tile index -> (m-panel, n-tile): groups of `gm` m-panels; inside a group the m-panel runs fastest, then the n-tile; the
// last group may be narrower.  gm = num_m gives plain m-fastest order.
__device__ __forceinline__ void tile_to_mn(int tile, int num_m, int num_n, int gm, int& m, int& n) {
  const int per_group = gm * num_n;
  const int g = tile / per_group;
  const int first_m = g * gm;
  const int rem = tile - g * per_group;
  const int width = min(gm, num_m - first_m);
  n = rem / width;
  m = first_m + (rem - n * width);
}

constexpr int VQ_EPI_DEBUG_MAINLOOP = 3;  // internal: discard accumulators (measures the TMA->MMA pipeline alone)
constexpr int VQ_EPI_DEBUG_LOADS = 5;     // internal: + TMEM loads (no math, no stores)
constexpr int VQ_EPI_DEBUG_MATH = 6;      // internal: + dequant math (no staging / stores)
constexpr int VQ_EPI_DEBUG_STORES = 7;    // internal: TMEM loads + staging + TMA stores, no dequant math

// Dequantise 32 consecutive output columns of one row (thread = row): int32 zero-point correction, one fp32 FMA with
// dx * dw and the bias, one rounding to fp16, optional GELU. Results stay in registers (16 packed half2).
template <int EPI>
__device__ __forceinline__ void dequant_chunk(const uint32_t (&v)[32], int32_t zx, int32_t rs, float dx,
                                              const int4* colp, uint32_t (&packed)[16]) {
  // colp: this chunk's 16 column-PAIR records in shared memory, two int4 per pair (warp-uniform -> broadcast LDS.128):
  //   [2j]   = {c1(n), c1(n+1), zw(n), zw(n+1)}      [2j+1] = {dw(n), dw(n+1), bias(n), bias(n+1)} (fp32 bits)
  const float2 dx2 = make_float2(dx, dx);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int4 ci = lds_v4(colp + 2 * j);
    const int4 cf = lds_v4(colp + 2 * j + 1);
    const int32_t t0 = static_cast<int32_t>(v[2 * j]) - zx * ci.x - rs * ci.z;
    const int32_t t1 = static_cast<int32_t>(v[2 * j + 1]) - zx * ci.y - rs * ci.w;
    const float2 s2 = __fmul2_rn(dx2, make_float2(__int_as_float(cf.x), __int_as_float(cf.y)));
    const float2 f2 = __ffma2_rn(make_float2(static_cast<float>(t0), static_cast<float>(t1)), s2,
                                 make_float2(__int_as_float(cf.z), __int_as_float(cf.w)));
    __half2 h2 = __floats2half2_rn(f2.x, f2.y);
    if (EPI == VQ_EPI_GELU_TANH) {
      const float2 g = gelu_tanh_pair(__half22float2(h2));
      h2 = __floats2half2_rn(g.x, g.y);
    }
    packed[j] = *reinterpret_cast<uint32_t*>(&h2);
  }
}

// Write one chunk (32 columns of this thread's row) into its staging sub-tile: row-major 64-byte rows, 16-byte piece
// index XOR ((row >> 1) & 3) == CU_TENSOR_MAP_SWIZZLE_64B (conflict-free). For the gated residual the sub-tile already
// holds the residual (TMA load, same swizzle): x_new = res + gate * y with the reference's two fp16 roundings.
template <int EPI>
__device__ __forceinline__ void stage_chunk(const GemmArgs& p, uint32_t (&packed)[16], int row, bool row_ok, int col0,
                                            uint8_t* sub, int lane) {
  const uint32_t sw = (static_cast<uint32_t>(lane) >> 1) & 3u;
  const uint32_t base = smem_u32(sub) + lane * (EPI_CHUNK * 2);
  if (EPI == VQ_EPI_GATE_RESIDUAL) {
    const __half* gate_row = p.gate + static_cast<size_t>((row_ok ? row : 0) / p.rows_per_gate) * p.N;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int n = col0 + g * 8;
      uint4 gv = make_uint4(0, 0, 0, 0);
      if (n < p.N) gv = __ldg(reinterpret_cast<const uint4*>(gate_row + n));
      const int4 rv = lds_v4_addr(base + ((g ^ sw) << 4));
      const __half2* g2 = reinterpret_cast<const __half2*>(&gv);
      const __half2* r2 = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __half2 y = *reinterpret_cast<__half2*>(&packed[g * 4 + e]);
        __half2 o = __hadd2_rn(r2[e], __hmul2_rn(g2[e], y));   // _rn: two roundings, never one fp16 FMA
        packed[g * 4 + e] = *reinterpret_cast<uint32_t*>(&o);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < 4; ++g)
    sts_v4_addr(base + ((g ^ sw) << 4), packed[g * 4 + 0], packed[g * 4 + 1], packed[g * 4 + 2], packed[g * 4 + 3]);
}

}  // namespace vq
