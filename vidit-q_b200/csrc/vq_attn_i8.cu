// Opt-in INT8 Q/K/V spatial self-attention on tcgen05 (`vq_attn_spatial_i8`): both matrix products of the attention run on
// the integer tensor cores (tcgen05.mma.kind::i8, s32 accumulators in TMEM).
//
// NO REFERENCE COUNTERPART: the reference never quantises Q / K / V or the probabilities (the hooks are commented out,
// qdiff/models/quant_block.py:617-623, :630-632; STDiT / PixArt call flash-attn on the fp16 linear outputs,
// t2v/opensora/models/layers/blocks.py:169-188).  The default path of this library therefore keeps fp16 attention
// (vq_attn_spatial.cu); this kernel is the opt-in variant BASELINE.json's north_star names, with its own tolerance
// (oracle/attn_i8_oracle.py is its restatement).
//
// Scheme (per sequence of S tokens, per head, head_dim 72):
//   Q8 = rint(Q * (1 / sq)),  sq = max|Q[token, head, :]| / 127              per token and head
//   K8 = rint((K - mean_tokens K) * (1 / sk)), sk per 64-key block and head  (softmax is invariant to the mean)
//   V8 = rint(V * (1 / sv)),  sv = max over the sequence of |V[:, head, dim]| / 127   per channel, stored TRANSPOSED
//   S  = Q8 K8^T                       exact s32;  x = S * (sq sk scale log2e) - m + log2(255),  m = the EXACT row maximum
//   P8 = rint(2^x) as u8 (<= 255)
//   O  = sum_k P8 V8                   exact s32 accumulation in TMEM;  out = O * sv / sum_k 2^x   (un-rounded row sum)
// Two passes in front of the attention kernel produce the operands from the fused q|k|v GEMM output [n_seq * S, 3 * H * 72]:
//   ia_stats_kernel   per (sequence, channel): mean of K, max |V|                      (reads 2/3 of q|k|v)
//   ia_quant_kernel   per (64-token block, head): codes into [rows, {q,k}, H, 80] (72 codes + 8 zero bytes: TMA strides
//                     must be multiples of 16 B) and V8^T [n_seq, H, 72, S] (K-major B operand of P V), scales beside.
// The attention kernel has the skeleton of vq_attn_spatial_kernel (persistent CTAs, 12 warps: TMA producer, MMA issuer,
// TMEM owner, two softmax warpgroups with one thread per query row; work item = (sequence, head, 256 queries); two score
// buffers per query tile) with these differences:
//   * TWO PASSES over the key tiles per item instead of an online rescale: a u8 probability leaves one octave of head-room
//     for a lagging maximum, and on real score distributions that triggered an in-TMEM rescale of the integer accumulators
//     on every second key tile (first version: 380 us).  Pass A streams K8 and takes the exact row maximum (integer
//     maximum per 64-key tile x the tile's fp32 scale: score MMAs are cheap at INT8 rate, the buffer goes straight back to
//     the MMA warp); pass B streams K8 again and V8^T once, with a fixed maximum: no rescale, no branch, full 8-bit P;
//   * operands are bytes: Q8 / K8 rows are a 64-byte SWIZZLE_64B tile (dims 0..63: two K = 32 steps) plus a 32-byte
//     SWIZZLE_32B tile (dims 64..95; bytes past the tensor map's 80-byte extent are TMA zero fill): K = 96 for the MMA;
//     V8^T tiles are [80 dims (72..79 zero fill) x 64 keys] SWIZZLE_64B, two K = 32 steps of one N = 80 instruction each;
//   * Q8 is the A operand from shared memory (12 KB per score tile), P8 the A operand from TMEM (16 columns written over S);
//   * scores become floats with the 1.5 * 2^23 magic add (integer add + exact fp32 subtract: no I2F on the MUFU pipe), the
//     probabilities become bytes with the same constant (fp32 add rounds to nearest even, the byte is the low mantissa byte);
//   * the scales an item needs are prefetched one item ahead and published through shared memory (no global load inside
//     the key-tile loops).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include <climits>

#include "vq_internal.h"
#include "vq_ptx.cuh"

namespace vq {

constexpr int IA_D = 72;
constexpr int IA_DV = 72;                  // rows of V8^T per (sequence, head)
constexpr int IA_DP = 80;                  // bytes per (token, head) in the Q8 / K8 buffer
constexpr int IA_BM = 128;                 // queries per tile
constexpr int IA_BN = 64;                  // keys per tile = K scale block
constexpr int IA_QT = 2;                   // query tiles per item
constexpr int IA_STAGES = 4;
constexpr int IA_Q_A = IA_BM * 64;         // dims 0..63, SWIZZLE_64B
constexpr int IA_Q_B = IA_BM * 32;         // dims 64..95, SWIZZLE_32B
constexpr int IA_QTILE = IA_Q_A + IA_Q_B;
constexpr int IA_K_A = IA_BN * 64;
constexpr int IA_K_B = IA_BN * 32;
constexpr int IA_KTILE = IA_K_A + IA_K_B;
constexpr int IA_VBYTES = 80 * IA_BN;      // V8^T tile: 80 dim rows x 64 key bytes
constexpr int IA_VTILE = 6144;             // ring pitch (keeps every tile 512-byte aligned)
constexpr int IA_OSTAGE = 128 * IA_D * 2;
constexpr int IA_SMEM_K = IA_QT * IA_QTILE;
constexpr int IA_SMEM_V = IA_SMEM_K + IA_STAGES * IA_KTILE;
constexpr int IA_SMEM_O = IA_SMEM_V + IA_STAGES * IA_VTILE;
constexpr int IA_SMEM_BAR = IA_SMEM_O + IA_QT * IA_OSTAGE;
constexpr int IA_SMEM_SCALES = IA_SMEM_BAR + 512;                   // 2 tiles x 2 parities x 136 floats
constexpr int IA_SCALES = 136;                                      // 64 key-block scales (S <= 4096) + 72 V scales
constexpr int IA_SMEM_BYTES = IA_SMEM_SCALES + IA_QT * 2 * IA_SCALES * 4 + 1024;
constexpr int IA_THREADS = 384;
constexpr uint32_t IA_TMEM_COLS = 512;
constexpr uint32_t IA_O_COL = 256;         // O accumulators; S / P8 buffer b of tile t at t * 128 + b * 64
constexpr float IA_MAGIC = 12582912.0f;    // 1.5 * 2^23
constexpr float IA_LOG2_255 = 7.994353436858858f;

struct AttnI8Args {
  int n_seq, S, H;
  float scale_log2e;
  const float* sq;   // [n_seq * S, H]
  const float* sk;   // [n_seq * S / 64, H]
  const float* sv;   // [n_seq, H * 72]
};

__device__ __forceinline__ float ia_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for a pair on the FMA pipe: x = n + f, |f| <= 0.5, degree-5 interpolant (relative error 2.4e-7, that of ex2.approx:
// the probabilities are rounded to integers next, so the polynomial must not move them across a rounding boundary more
// often than the MUFU would)
__device__ __forceinline__ float2 ia_exp2_poly(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 xr = __fadd2_rn(x, make_float2(IA_MAGIC, IA_MAGIC));
  const float2 xi = __fadd2_rn(xr, make_float2(-IA_MAGIC, -IA_MAGIC));
  const float2 f = __fadd2_rn(x, make_float2(-xi.x, -xi.y));
  float2 p = __ffma2_rn(make_float2(0.0013390863314270973f, 0.0013390863314270973f), f,
                        make_float2(0.009676031768321991f, 0.009676031768321991f));
  p = __ffma2_rn(p, f, make_float2(0.055503569543361664f, 0.055503569543361664f));
  p = __ffma2_rn(p, f, make_float2(0.2402210682630539f, 0.2402210682630539f));
  p = __ffma2_rn(p, f, make_float2(0.6931471824645996f, 0.6931471824645996f));
  p = __ffma2_rn(p, f, make_float2(1.0000001192092896f, 1.0000001192092896f));
  float2 r;
  r.x = __uint_as_float(__float_as_uint(p.x) + (__float_as_uint(xr.x) << 23));
  r.y = __uint_as_float(__float_as_uint(p.y) + (__float_as_uint(xr.y) << 23));
  return r;
}

// ----------------------------------------------------------------------------- operand preparation
// per (sequence, channel): mean of K over the sequence's tokens, max |V| / 127.  Block = 32 chunks of 8 channels (16-byte
// loads) x 16 token slices; a chunk lies entirely in the k part (sum) or the v part (maximum) of the row.
__global__ void __launch_bounds__(512) ia_stats_kernel(const __half* __restrict__ qkv, float* __restrict__ kmean,
                                                       float* __restrict__ sv, float* __restrict__ svi, int S, int C) {
  grid_dep_sync();
  __shared__ float s_acc[16][32][8];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int col0 = (blockIdx.x * 32 + tx) * 8;     // first of this thread's 8 channels inside the k|v part (2C channels)
  const int seq = blockIdx.y;
  const bool is_k = col0 < C;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (col0 < 2 * C) {
    const __half* src = qkv + static_cast<size_t>(seq) * S * 3 * C + C + col0;
    for (int tok = ty; tok < S; tok += 16) {
      const int4 raw = *reinterpret_cast<const int4*>(src + static_cast<size_t>(tok) * 3 * C);
      const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 v = __half22float2(h2[i]);
        if (is_k) {
          acc[2 * i] += v.x;
          acc[2 * i + 1] += v.y;
        } else {
          acc[2 * i] = fmaxf(acc[2 * i], fabsf(v.x));
          acc[2 * i + 1] = fmaxf(acc[2 * i + 1], fabsf(v.y));
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) s_acc[ty][tx][i] = acc[i];
  __syncthreads();
  if (ty < 8 && col0 < 2 * C) {
    // thread (tx, ty < 8) finishes channel col0 + ty
    float r = s_acc[0][tx][ty];
    for (int i = 1; i < 16; ++i) r = is_k ? r + s_acc[i][tx][ty] : fmaxf(r, s_acc[i][tx][ty]);
    const int col = col0 + ty;
    if (is_k) {
      kmean[static_cast<size_t>(seq) * C + col] = r * (1.0f / static_cast<float>(S));
    } else {
      float a = r / 127.0f;
      a = a > 0.f ? a : 1.0f;
      sv[static_cast<size_t>(seq) * C + col - C] = a;
      svi[static_cast<size_t>(seq) * C + col - C] = 1.0f / a;
    }
  }
}

// code = rint(v * (1 / s)): one IEEE reciprocal per scale, then ONE fused multiply-add per pair with the 1.5 * 2^23 magic
// constant — the exact product is rounded to the nearest integer (half to even) in a single step and the code is the low
// byte of the result's bit pattern (two's complement).  No clamp: |v| <= the maximum the scale was taken from, so
// |v / s| <= 127 (1 + 2^-22) and rounds to at most 127.  The oracle (oracle/attn_i8_oracle.py quantise_qkv) forms the same
// exact product in fp64 and rounds it once.
__device__ __forceinline__ uint32_t ia_code4(const float (&x)[20], int i, float2 inv2) {
  const float2 magic = make_float2(IA_MAGIC, IA_MAGIC);
  const float2 a = __ffma2_rn(make_float2(x[i], x[i + 1]), inv2, magic);
  const float2 b = __ffma2_rn(make_float2(x[i + 2], x[i + 3]), inv2, magic);
  return __byte_perm(__byte_perm(__float_as_uint(a.x), __float_as_uint(a.y), 0x0040),
                     __byte_perm(__float_as_uint(b.x), __float_as_uint(b.y), 0x0040), 0x5410);
}
__device__ __forceinline__ int8_t ia_code1(float v, float inv) {
  return static_cast<int8_t>(__float_as_uint(fmaf(v, inv, IA_MAGIC)) & 0xffu);
}
// 20 fp16 values of one (token, quarter row) from the staged tile (8-byte aligned), dims past `nd` read as zero
__device__ __forceinline__ void ia_load20(const __half* tile, int token, int d0, int nd, float (&x)[20]) {
  const __half* p = tile + token * IA_D + d0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    uint2 r = make_uint2(0u, 0u);
    if (4 * i < nd) r = *reinterpret_cast<const uint2*>(p + 4 * i);
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    x[4 * i] = lo.x; x[4 * i + 1] = lo.y; x[4 * i + 2] = hi.x; x[4 * i + 3] = hi.y;
  }
}

// per (64-token block, head): Q8 / K8 rows of 80 bytes, V8 transposed, the scales.  The three [64 x 72] fp16 tiles are
// staged in shared memory with 16-byte loads, the codes leave through shared memory with 16-byte stores.  256 threads =
// 64 tokens x 4 threads; thread `sub` owns dims [20 sub, 20 sub + 20) (sub 3: dims 60..71 and the 8 zero pad bytes).
__global__ void __launch_bounds__(256, 3) ia_quant_kernel(const __half* __restrict__ qkv, const float* __restrict__ kmean,
                                                          const float* __restrict__ svi, int8_t* __restrict__ qk8,
                                                          int8_t* __restrict__ vt8, float* __restrict__ sq,
                                                          float* __restrict__ sk, int S, int H) {
  grid_dep_sync();
  __shared__ __align__(16) __half in_st[3][64 * IA_D];
  __shared__ __align__(16) int8_t qk_st[2][64 * IA_DP];
  __shared__ __align__(16) int8_t v_st[IA_DV * 64];
  __shared__ float s_red[8];
  const int C = H * IA_D;
  const int tid = threadIdx.x, token = tid >> 2, sub = tid & 3;
  // head is the fastest grid dimension: the 16 CTAs of a token block run together and read whole q|k|v rows between them
  const int h = blockIdx.x;
  const size_t row0 = static_cast<size_t>(blockIdx.y) * 64;
  const size_t row = row0 + token;
  const int seq = static_cast<int>(row0 / S);
  const int d0 = 20 * sub;
  const int nd = sub < 3 ? 20 : 12;
  for (int c = tid; c < 3 * 64 * 9; c += 256) {
    const int which = c / (64 * 9), r = (c % (64 * 9)) / 9, part = c % 9;
    *reinterpret_cast<int4*>(&in_st[which][r * IA_D + part * 8]) =
        *reinterpret_cast<const int4*>(qkv + (row0 + r) * 3 * C + which * C + h * IA_D + part * 8);
  }
  __syncthreads();
  float x[20];
  // ---- Q: per (token, head) scale
  ia_load20(in_st[0], token, d0, nd, x);
  float am = 0.f;
#pragma unroll
  for (int i = 0; i < 20; ++i) am = fmaxf(am, fabsf(x[i]));
  am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 1));
  am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 2));
  {
    float s_q = am / 127.0f;
    s_q = s_q > 0.f ? s_q : 1.0f;
    const float inv = 1.0f / s_q;
    if (sub == 0) sq[row * H + h] = s_q;
    uint32_t* dq = reinterpret_cast<uint32_t*>(&qk_st[0][token * IA_DP + d0]);
#pragma unroll
    for (int w = 0; w < 5; ++w) dq[w] = ia_code4(x, 4 * w, make_float2(inv, inv));
  }
  // ---- K: minus the sequence mean, per (64-token block, head) scale
  ia_load20(in_st[1], token, d0, nd, x);
  const float* km = kmean + static_cast<size_t>(seq) * C + h * IA_D + d0;
  am = 0.f;
#pragma unroll
  for (int i = 0; i < 20; ++i) {
    x[i] = i < nd ? x[i] - __ldg(km + i) : 0.f;
    am = fmaxf(am, fabsf(x[i]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
  if ((tid & 31) == 0) s_red[tid >> 5] = am;
  __syncthreads();
  am = s_red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) am = fmaxf(am, s_red[i]);
  {
    float s_k = am / 127.0f;
    s_k = s_k > 0.f ? s_k : 1.0f;
    const float inv = 1.0f / s_k;
    if (tid == 0) sk[static_cast<size_t>(blockIdx.y) * H + h] = s_k;
    uint32_t* dk = reinterpret_cast<uint32_t*>(&qk_st[1][token * IA_DP + d0]);
#pragma unroll
    for (int w = 0; w < 5; ++w) dk[w] = ia_code4(x, 4 * w, make_float2(inv, inv));
  }
  // ---- V: per (sequence, channel) scale, transposed
  ia_load20(in_st[2], token, d0, nd, x);
  const float* svp = svi + static_cast<size_t>(seq) * C + h * IA_D + d0;
#pragma unroll
  for (int i = 0; i < 20; ++i)
    if (i < nd) v_st[(d0 + i) * 64 + token] = ia_code1(x[i], __ldg(svp + i));
  __syncthreads();
  for (int c = tid; c < 2 * 64 * 5; c += 256) {
    const int which = c / (64 * 5), r = (c % (64 * 5)) / 5, part = c % 5;
    *reinterpret_cast<int4*>(qk8 + ((row0 + r) * 2 * H + which * H + h) * IA_DP + part * 16) =
        *reinterpret_cast<const int4*>(&qk_st[which][r * IA_DP + part * 16]);
  }
  const int tok0 = static_cast<int>(row0 % S);
  for (int c = tid; c < IA_DV * 4; c += 256) {
    const int d = c >> 2, part = c & 3;
    *reinterpret_cast<int4*>(vt8 + ((static_cast<size_t>(seq) * H + h) * IA_DV + d) * S + tok0 + part * 16) =
        *reinterpret_cast<const int4*>(v_st + d * 64 + part * 16);
  }
}

// ----------------------------------------------------------------------------- attention
template <int EMU>
__global__ void __launch_bounds__(IA_THREADS, 1)
vq_attn_i8_kernel(const __grid_constant__ CUtensorMap tmap_qa, const __grid_constant__ CUtensorMap tmap_qb,
                  const __grid_constant__ CUtensorMap tmap_ka, const __grid_constant__ CUtensorMap tmap_kb,
                  const __grid_constant__ CUtensorMap tmap_vt, const __grid_constant__ CUtensorMap tmap_o,
                  const AttnI8Args a) {
  extern __shared__ uint8_t ia_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ia_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + IA_SMEM_K;
  uint8_t* smem_v = smem + IA_SMEM_V;
  uint8_t* smem_o = smem + IA_SMEM_O;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + IA_SMEM_BAR);
  uint64_t* q_full = bars;                     // Q8 tiles of the item landed
  uint64_t* q_empty = bars + 1;                // the last score MMA of the item has read them
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = k_full + IA_STAGES;
  uint64_t* v_full = k_empty + IA_STAGES;
  uint64_t* v_empty = v_full + IA_STAGES;
  uint64_t* s_full = v_empty + IA_STAGES;      // [tile][buffer] scores are in TMEM
  uint64_t* p_full = s_full + 2 * IA_QT;       // [tile][buffer] P8 is in TMEM (and O rescaled if needed)
  uint64_t* o_full = p_full + 2 * IA_QT;       // [tile] the item's last P V has retired: O is complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + IA_QT);
  // per-item scales of each softmax warpgroup, double-buffered by item parity: [tile][parity][64 key-block scales | 72 sv]
  float* s_scales = reinterpret_cast<float*>(smem + IA_SMEM_SCALES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nqp = a.S / (IA_QT * IA_BM);
  const int nkv = a.S / IA_BN;                 // even, >= 4
  const int num_items = a.n_seq * a.H * nqp;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qa);
    tma_prefetch_desc(&tmap_qb);
    tma_prefetch_desc(&tmap_ka);
    tma_prefetch_desc(&tmap_kb);
    tma_prefetch_desc(&tmap_vt);
    tma_prefetch_desc(&tmap_o);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < IA_STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int i = 0; i < 2 * IA_QT; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);   // one arrival per softmax warp of the tile
    }
    for (int i = 0; i < IA_QT; ++i) mbar_init(&o_full[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, IA_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  grid_dep_sync();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t kc = 0, vc = 0;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int qp = item % nqp;
        const int h = (item / nqp) % a.H;
        const int seq = item / (nqp * a.H);
        const int row_q0 = seq * a.S + qp * (IA_QT * IA_BM);
        const int row_kv0 = seq * a.S;
        mbar_wait(q_empty, (it & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, IA_QT * IA_QTILE);
        for (int t = 0; t < IA_QT; ++t) {
          tma_load_4d_hint(smem_q + t * IA_QTILE, &tmap_qa, q_full, 0, h, 0, row_q0 + t * IA_BM, kEvictFirst);
          tma_load_4d_hint(smem_q + t * IA_QTILE + IA_Q_A, &tmap_qb, q_full, 64, h, 0, row_q0 + t * IA_BM, kEvictFirst);
        }
        // K8 is streamed twice per item (pass A: row maxima, pass B: probabilities), V8^T once (pass B)
        for (int i = 0; i < 2 * nkv; ++i, ++kc) {
          const int j = i < nkv ? i : i - nkv;
          const int s = kc % IA_STAGES;
          const uint32_t ph = (kc / IA_STAGES) & 1;
          mbar_wait(&k_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&k_full[s], IA_KTILE);
          tma_load_4d_hint(smem_k + s * IA_KTILE, &tmap_ka, &k_full[s], 0, h, 1, row_kv0 + j * IA_BN, kEvictLast);
          tma_load_4d_hint(smem_k + s * IA_KTILE + IA_K_A, &tmap_kb, &k_full[s], 64, h, 1, row_kv0 + j * IA_BN, kEvictLast);
          if (i >= nkv) {
            const int sv = vc % IA_STAGES;
            mbar_wait(&v_empty[sv], ((vc / IA_STAGES) & 1) ^ 1);
            mbar_arrive_expect_tx(&v_full[sv], IA_VBYTES);
            tma_load_3d_hint(smem_v + sv * IA_VTILE, &tmap_vt, &v_full[sv], j * IA_BN, 0, seq * a.H + h, kEvictLast);
            ++vc;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    constexpr uint32_t idesc_s = make_idesc_i8(IA_BM, IA_BN, 1, 1);    // s8 x s8
    constexpr uint32_t idesc_pv = make_idesc_i8(IA_BM, 80, 0, 1);      // u8 P x s8 V
    // descriptor high words: stride byte offset (8-row group pitch) | version 1 | layout (4 = SWIZZLE_64B, 6 = SWIZZLE_32B)
    constexpr uint64_t hi_sw64 = static_cast<uint64_t>((512u >> 4) | (1u << 14) | (4u << 29)) << 32;
    constexpr uint64_t hi_sw32 = static_cast<uint64_t>((256u >> 4) | (1u << 14) | (6u << 29)) << 32;
    constexpr uint32_t lbo = 1u << 16;   // ignored for swizzled K-major tiles
    const uint32_t q_lo = ((smem_u32(smem_q) & 0x3FFFFu) >> 4) | lbo;
    const uint32_t k_lo = ((smem_u32(smem_k) & 0x3FFFFu) >> 4) | lbo;
    const uint32_t v_lo = ((smem_u32(smem_v) & 0x3FFFFu) >> 4) | lbo;
    // S[t][b] = Q8[t] K8^T: K = 96 bytes = two 32-byte steps inside the 64-byte swizzle rows + one step in the 32-byte tile
    auto issue_s = [&](int t, int b, int ks) {
      const uint32_t qa = q_lo + t * (IA_QTILE >> 4);
      const uint32_t ka = k_lo + ks * (IA_KTILE >> 4);
      const uint32_t d = tmem_base + t * 128 + b * IA_BN;
      tc_mma_i8(d, hi_sw64 | qa, hi_sw64 | ka, idesc_s, 0u);
      tc_mma_i8(d, hi_sw64 | (qa + 2), hi_sw64 | (ka + 2), idesc_s, 1u);
      tc_mma_i8(d, hi_sw32 | (qa + (IA_Q_A >> 4)), hi_sw32 | (ka + (IA_K_A >> 4)), idesc_s, 1u);
      tc_commit(&s_full[2 * t + b]);
    };
    // O[t] (+)= P8[t][b] V8: K = 64 keys = two 32-byte steps of the [80 dims x 64 keys] tile
    auto issue_pv = [&](int t, int b, int vs, uint32_t acc, bool last) {
      const uint32_t va = v_lo + vs * (IA_VTILE >> 4);
      const uint32_t p = tmem_base + t * 128 + b * IA_BN;
      const uint32_t o = tmem_base + IA_O_COL + t * 128;
      tc_mma_i8_ts(o, p, hi_sw64 | va, idesc_pv, acc);
      tc_mma_i8_ts(o, p + 8, hi_sw64 | (va + 2), idesc_pv, 1u);
      if (last) tc_commit(&o_full[t]);   // one completion per item: the epilogue is its only waiter
    };
    uint32_t kc = 0, vc = 0;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      mbar_wait(q_full, it & 1);
      for (int i = 0; i < 2; ++i) {
        const int ks = kc % IA_STAGES;
        mbar_wait(&k_full[ks], (kc / IA_STAGES) & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_s(0, i, ks);
          issue_s(1, i, ks);
          tc_commit(&k_empty[ks]);
        }
        __syncwarp();
        ++kc;
      }
      // 2 nkv score tiles per item: i < nkv are pass A (the softmax warps only take the row maximum: the buffer is handed
      // back through p_full without a P V), i >= nkv are pass B (P8 written over the scores, O += P8 V8)
      const int n2 = 2 * nkv;
      for (int i = 0; i < n2; ++i) {
        const uint32_t bph = (static_cast<uint32_t>(it) * nkv + (i >> 1)) & 1;   // phase of the per-buffer barriers
        const int b = i & 1;
        const bool pv = i >= nkv;
        const bool more = i + 2 < n2;
        const int vs = vc % IA_STAGES;
        const int ks = kc % IA_STAGES;
        if (pv) mbar_wait(&v_full[vs], (vc / IA_STAGES) & 1);
        if (more) mbar_wait(&k_full[ks], (kc / IA_STAGES) & 1);
#pragma unroll
        for (int t = 0; t < IA_QT; ++t) {
          mbar_wait(&p_full[2 * t + b], bph);
          tc_fence_after();
          if (elect_one()) {
            if (pv) issue_pv(t, b, vs, i > nkv ? 1u : 0u, i + 1 == n2);
            if (more) issue_s(t, b, ks);
          }
          __syncwarp();
        }
        if (elect_one()) {
          if (pv) tc_commit(&v_empty[vs]);
          if (more) tc_commit(&k_empty[ks]);
          if (i + 3 == n2) tc_commit(q_empty);   // the item's last score MMAs are in flight: Q8 may be refilled behind them
        }
        __syncwarp();
        if (pv) ++vc;
        if (more) ++kc;
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax / correction / epilogue =====================
    const int t = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int wg_thread = threadIdx.x - (4 + 4 * t) * 32;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off + t * 128;
    const uint32_t o_addr = tmem_base + lane_off + IA_O_COL + t * 128;
    uint8_t* ostage = smem_o + t * IA_OSTAGE;
    // The scales an item needs — sq of this thread's row, the (sequence, head)'s key-block scales and V scales — are
    // loaded one item ahead into registers and published through shared memory at the top of the item: a global load
    // inside the key-tile loop put its whole latency into every iteration (ncu: long_scoreboard on the loop branch).
    auto item_coords = [&](int item, int& qp, int& h, int& seq) {
      qp = item % nqp;
      h = (item / nqp) % a.H;
      seq = item / (nqp * a.H);
    };
    auto prefetch = [&](int item, float& r_sq, float& r_a, float& r_b) {
      int qp, h, seq;
      item_coords(item, qp, h, seq);
      const int grow = seq * a.S + qp * (IA_QT * IA_BM) + t * IA_BM + row;
      r_sq = __ldg(a.sq + static_cast<size_t>(grow) * a.H + h);
      // thread i < nkv of the warpgroup carries key-block scale i, thread d < 72 carries V scale d
      r_a = wg_thread < nkv ? __ldg(a.sk + (static_cast<size_t>(seq) * nkv + wg_thread) * a.H + h) : 0.f;
      r_b = wg_thread < IA_D ? __ldg(a.sv + (static_cast<size_t>(seq) * a.H + h) * IA_D + wg_thread) : 0.f;
    };
    float r_sq = 0.f, r_a = 0.f, r_b = 0.f;
    if (static_cast<int>(blockIdx.x) < num_items) prefetch(blockIdx.x, r_sq, r_a, r_b);
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      int qp, h, seq;
      item_coords(item, qp, h, seq);
      float* sc = s_scales + (t * 2 + (it & 1)) * IA_SCALES;
      if (wg_thread < nkv) sc[wg_thread] = r_a;
      if (wg_thread < IA_D) sc[64 + wg_thread] = r_b;
      const float c_row = r_sq * a.scale_log2e;
      named_bar_sync(1 + t, 128);   // also orders this item's writes behind every thread's reads of two items ago
      if (item + static_cast<int>(gridDim.x) < num_items) prefetch(item + gridDim.x, r_sq, r_a, r_b);
      const float* skp = sc;
      const uint32_t uses = static_cast<uint32_t>(it) * nkv;   // completions of every per-buffer barrier before this item
      // ---- pass A: the exact row maximum (log2 units) over all key tiles — integer maximum per tile, one multiply
      float m = -INFINITY;
      for (int i = 0; i < nkv; ++i) {
        const int b = i & 1;
        const uint32_t sa = s_addr + b * IA_BN;
        const float c_rt = c_row * skp[i];
        mbar_wait(&s_full[2 * t + b], (uses + (i >> 1)) & 1);
        tc_fence_after();
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(sa, v0);
        tmem_ld_32x32b_x32(sa + 32, v1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * t + b]);   // the buffer may take the scores of tile i + 2
        int mi0 = INT_MIN, mi1 = INT_MIN, mi2 = INT_MIN, mi3 = INT_MIN;
#pragma unroll
        for (int k = 0; k < 32; k += 4) {   // VIMNMX3: two scores per instruction
          mi0 = __vimax3_s32(mi0, static_cast<int>(v0[k]), static_cast<int>(v0[k + 1]));
          mi1 = __vimax3_s32(mi1, static_cast<int>(v0[k + 2]), static_cast<int>(v0[k + 3]));
          mi2 = __vimax3_s32(mi2, static_cast<int>(v1[k]), static_cast<int>(v1[k + 1]));
          mi3 = __vimax3_s32(mi3, static_cast<int>(v1[k + 2]), static_cast<int>(v1[k + 3]));
        }
        m = fmaxf(m, static_cast<float>(max(max(mi0, mi1), max(mi2, mi3))) * c_rt);
      }
      // ---- pass B: P8 = rint(255 * 2^(S c - m)) as bytes over the first 16 columns of S; row sum of the un-rounded values
      const float neg = IA_LOG2_255 - m;
      const float2 pmagic = make_float2(IA_MAGIC, IA_MAGIC);
      const float2 nmagic = make_float2(-IA_MAGIC, -IA_MAGIC);
      float l = 0.f;   // row sum of the UN-rounded 2^x: normalising by the rounded bytes' sum would drop the mass of the
                       // many small probabilities from the denominator only (measured: 4.2e-2 -> 6.9e-2 against fp attention)
      for (int j = 0; j < nkv; ++j) {
        const int b = j & 1;
        const uint32_t sa = s_addr + b * IA_BN;
        const float c_rt = c_row * skp[j];
        mbar_wait(&s_full[2 * t + b], (uses + ((nkv + j) >> 1)) & 1);
        tc_fence_after();
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(sa, v0);
        tmem_ld_32x32b_x32(sa + 32, v1);
        tmem_ld_wait();
        const float2 c2 = make_float2(c_rt, c_rt);
        const float2 neg2 = make_float2(neg, neg);
        float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
        uint32_t pk[16];
#pragma unroll
        for (int w = 0; w < 16; ++w) {
          const uint32_t* src = w < 8 ? v0 : v1;
          const int k4 = 4 * (w & 7);
          uint32_t by[4];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            // integer -> float: 0x4B400000 + s is the float 1.5 * 2^23 + s (|s| < 2^22), the subtraction is exact
            // (folding the offset into the FMA addend instead costs 1e-3 of accuracy: the addend is then ~1e4 with a 1e-3 ulp)
            const float2 sf = __fadd2_rn(make_float2(__uint_as_float(0x4B400000u + src[k4 + 2 * hh]),
                                                     __uint_as_float(0x4B400000u + src[k4 + 2 * hh + 1])), nmagic);
            const float2 x = __ffma2_rn(sf, c2, neg2);
            float2 e;
            if ((((2 * w + hh) * EMU) & 15) < EMU) {
              e = ia_exp2_poly(x);
            } else {
              e.x = ia_exp2(x.x);
              e.y = ia_exp2(x.y);
            }
            if (hh) lb = __fadd2_rn(lb, e);
            else la = __fadd2_rn(la, e);
            const float2 y = __fadd2_rn(e, pmagic);   // round to nearest even: the low mantissa byte is the integer
            by[2 * hh] = __float_as_uint(y.x);
            by[2 * hh + 1] = __float_as_uint(y.y);
          }
          pk[w] = __byte_perm(__byte_perm(by[0], by[1], 0x0040), __byte_perm(by[2], by[3], 0x0040), 0x5410);
        }
        tmem_st_32x32b_x16(sa, pk);
        l += (la.x + la.y) + (lb.x + lb.y);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * t + b]);
      }
      // ---- epilogue: O * sv / l -> fp16 -> dense staging tile -> one TMA store of [128 queries x 72 dims]
      {
        mbar_wait(&o_full[t], it & 1);
        tc_fence_after();
        uint32_t v0[32], v1[32], w[8];
        tmem_ld_32x32b_x32(o_addr, v0);
        tmem_ld_32x32b_x32(o_addr + 32, v1);
        tmem_ld_32x32b_x8(o_addr + 64, w);
        tmem_ld_wait();
        const float inv = __fdividef(1.0f, l);
        const float* svp = sc + 64;
        uint32_t pk[36];
#pragma unroll
        for (int i = 0; i < 36; ++i) {
          const int x0 = static_cast<int>(i < 16 ? v0[2 * i] : (i < 32 ? v1[2 * i - 32] : w[2 * i - 64]));
          const int x1 = static_cast<int>(i < 16 ? v0[2 * i + 1] : (i < 32 ? v1[2 * i - 31] : w[2 * i - 63]));
          const float2 s2 = *(reinterpret_cast<const float2*>(svp) + i);
          const __half2 hv = __floats2half2_rn(__int2float_rn(x0) * (s2.x * inv), __int2float_rn(x1) * (s2.y * inv));
          pk[i] = *reinterpret_cast<const uint32_t*>(&hv);
        }
        if (wg_thread == 0) tma_store_wait_read<0>();
        named_bar_sync(1 + t, 128);
        const uint32_t dst = smem_u32(ostage) + row * (IA_D * 2);
#pragma unroll
        for (int i = 0; i < 9; ++i) sts_v4_addr(dst + 16 * i, pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        fence_proxy_async_smem();
        named_bar_sync(1 + t, 128);
        if (wg_thread == 0) {
          tma_store_3d(&tmap_o, ostage, 0, h, seq * a.S + qp * (IA_QT * IA_BM) + t * IA_BM);
          tma_store_commit();
        }
      }
    }
    if (wg_thread == 0) tma_store_wait<0>();
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, IA_TMEM_COLS);
  }
}

// ----------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiledIA)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiledIA ia_encode_fn() {
  static PFN_encodeTiledIA fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiledIA>(ptr);
  return fn;
}

static int ia_tmap(CUtensorMap* out, CUtensorMapDataType dt, uint32_t rank, const void* base, const cuuint64_t* gdim,
                   const cuuint64_t* gstride, const cuuint32_t* box, CUtensorMapSwizzle sw) {
  PFN_encodeTiledIA enc = ia_encode_fn();
  if (!enc) return VQ_ERR_DRIVER;
  cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  CUresult r = enc(out, dt, rank, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? VQ_OK : VQ_ERR_TMAP;
}

// workspace layout (all pieces 256-byte aligned)
struct IaLayout {
  size_t qk8, vt8, sq, sk, sv, kmean, svi, total;
};
static IaLayout ia_layout(int n_seq, int S, int H) {
  const size_t rows = static_cast<size_t>(n_seq) * S, C = static_cast<size_t>(H) * IA_D;
  auto up = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
  IaLayout L;
  L.qk8 = 0;
  L.vt8 = up(L.qk8 + rows * 2 * H * IA_DP);
  L.sq = up(L.vt8 + static_cast<size_t>(n_seq) * H * IA_DV * S);
  L.sk = up(L.sq + rows * H * 4);
  L.sv = up(L.sk + rows / IA_BN * H * 4);
  L.kmean = up(L.sv + static_cast<size_t>(n_seq) * C * 4);
  L.svi = up(L.kmean + static_cast<size_t>(n_seq) * C * 4);
  L.total = up(L.svi + static_cast<size_t>(n_seq) * C * 4);
  return L;
}

static int ia_check(const void* a, const void* b, int n_seq, int S, int H, int head_dim) {
  if (!a || !b || n_seq <= 0 || S <= 0 || H <= 0) return VQ_ERR_ARG;
  if (head_dim != IA_D || (S % (IA_QT * IA_BM)) != 0 || S < 4 * IA_BN || S > 64 * IA_BN) return VQ_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a) & 255) || (reinterpret_cast<uintptr_t>(b) & 15)) return VQ_ERR_ARG;
  if (static_cast<uint64_t>(n_seq) * S * 3 * H * IA_D >= (1ull << 40) || static_cast<uint64_t>(n_seq) * S / 64 > 65535)
    return VQ_ERR_UNSUPPORTED;
  return VQ_OK;
}

static int ia_quantise(const void* qkv, void* ws, int n_seq, int S, int H, cudaStream_t st) {
  const IaLayout L = ia_layout(n_seq, S, H);
  uint8_t* w = static_cast<uint8_t*>(ws);
  const int C = H * IA_D;
  float* kmean = reinterpret_cast<float*>(w + L.kmean);
  float* sv = reinterpret_cast<float*>(w + L.sv);
  float* svi = reinterpret_cast<float*>(w + L.svi);
  launch_pdl(ia_stats_kernel, dim3((2 * C + 255) / 256, n_seq), dim3(32, 16), 0, st, static_cast<const __half*>(qkv), kmean,
             sv, svi, S, C);
  if (cudaGetLastError() != cudaSuccess) return VQ_ERR_LAUNCH;
  launch_pdl(ia_quant_kernel, dim3(H, static_cast<unsigned>(static_cast<size_t>(n_seq) * S / 64)), dim3(256), 0, st,
             static_cast<const __half*>(qkv), static_cast<const float*>(kmean), static_cast<const float*>(svi),
             reinterpret_cast<int8_t*>(w + L.qk8), reinterpret_cast<int8_t*>(w + L.vt8),
             reinterpret_cast<float*>(w + L.sq), reinterpret_cast<float*>(w + L.sk), S, H);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

template <int EMU>
static int ia_launch(const CUtensorMap& qa, const CUtensorMap& qb, const CUtensorMap& ka, const CUtensorMap& kb,
                     const CUtensorMap& vt, const CUtensorMap& to, const AttnI8Args& a, int grid, cudaStream_t st) {
  static bool attr_dev[kMaxDevices] = {};
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    if (cudaFuncSetAttribute(vq_attn_i8_kernel<EMU>, cudaFuncAttributeMaxDynamicSharedMemorySize, IA_SMEM_BYTES) !=
        cudaSuccess)
      return VQ_ERR_LAUNCH;
    attr = true;
  }
  launch_pdl(vq_attn_i8_kernel<EMU>, dim3(grid), dim3(IA_THREADS), IA_SMEM_BYTES, st, qa, qb, ka, kb, vt, to, a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

static int ia_attend(const void* ws, void* out, int n_seq, int S, int H, float scale, cudaStream_t st) {
  const IaLayout L = ia_layout(n_seq, S, H);
  const uint8_t* w = static_cast<const uint8_t*>(ws);
  const uint64_t rows = static_cast<uint64_t>(n_seq) * S;
  CUtensorMap qa, qb, ka, kb, vt, to;
  int rc;
  {
    // codes [rows, {q,k}, H, 80 bytes] as (byte 80, head H, which 2, token rows); dims past 80 are TMA zero fill
    cuuint64_t gdim[4] = {IA_DP, static_cast<cuuint64_t>(H), 2, rows};
    cuuint64_t gstr[3] = {IA_DP, static_cast<cuuint64_t>(H) * IA_DP, 2ull * H * IA_DP};
    cuuint32_t box_a_q[4] = {64, 1, 1, IA_BM}, box_b_q[4] = {32, 1, 1, IA_BM};
    cuuint32_t box_a_k[4] = {64, 1, 1, IA_BN}, box_b_k[4] = {32, 1, 1, IA_BN};
    if ((rc = ia_tmap(&qa, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, w + L.qk8, gdim, gstr, box_a_q, CU_TENSOR_MAP_SWIZZLE_64B)))
      return rc;
    if ((rc = ia_tmap(&qb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, w + L.qk8, gdim, gstr, box_b_q, CU_TENSOR_MAP_SWIZZLE_32B)))
      return rc;
    if ((rc = ia_tmap(&ka, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, w + L.qk8, gdim, gstr, box_a_k, CU_TENSOR_MAP_SWIZZLE_64B)))
      return rc;
    if ((rc = ia_tmap(&kb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, w + L.qk8, gdim, gstr, box_b_k, CU_TENSOR_MAP_SWIZZLE_32B)))
      return rc;
  }
  {
    // V8^T [n_seq * H, 72 dims, S tokens] as (token S, dim 72, sequence-head): box = 64 keys x 80 dims (72..79 zero fill)
    cuuint64_t gdim[3] = {static_cast<cuuint64_t>(S), IA_DV, static_cast<cuuint64_t>(n_seq) * H};
    cuuint64_t gstr[2] = {static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(S) * IA_DV};
    cuuint32_t box[3] = {IA_BN, 80, 1};
    if ((rc = ia_tmap(&vt, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, w + L.vt8, gdim, gstr, box, CU_TENSOR_MAP_SWIZZLE_64B)))
      return rc;
  }
  {
    cuuint64_t gdim[3] = {IA_D, static_cast<cuuint64_t>(H), rows};
    cuuint64_t gstr[2] = {IA_D * 2, static_cast<cuuint64_t>(H) * IA_D * 2};
    cuuint32_t box[3] = {IA_D, 1, IA_BM};
    PFN_encodeTiledIA enc = ia_encode_fn();
    if (!enc) return VQ_ERR_DRIVER;
    cuuint32_t estr[3] = {1u, 1u, 1u};
    if (enc(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return VQ_ERR_TMAP;
  }
  static const int emu = [] {
    const char* e = getenv("VQ_IA_EMU");
    return e ? atoi(e) : 4;
  }();
  AttnI8Args a{n_seq, S, H, scale * 1.4426950408889634f, reinterpret_cast<const float*>(w + L.sq),
               reinterpret_cast<const float*>(w + L.sk), reinterpret_cast<const float*>(w + L.sv)};
  const long long items = static_cast<long long>(n_seq) * H * (S / (IA_QT * IA_BM));
  const int grid = static_cast<int>(items < num_sms() ? items : num_sms());
  switch (emu) {
    case 0: return ia_launch<0>(qa, qb, ka, kb, vt, to, a, grid, st);
    case 8: return ia_launch<8>(qa, qb, ka, kb, vt, to, a, grid, st);
    default: return ia_launch<4>(qa, qb, ka, kb, vt, to, a, grid, st);
  }
}

}  // namespace vq

extern "C" int64_t vq_attn_i8_workspace_bytes(int n_seq, int S, int H, int head_dim) {
  if (n_seq <= 0 || S <= 0 || H <= 0 || head_dim != vq::IA_D || (S % (vq::IA_QT * vq::IA_BM)) != 0 || S > 64 * vq::IA_BN)
    return -1;
  return static_cast<int64_t>(vq::ia_layout(n_seq, S, H).total);
}

extern "C" int vq_attn_i8_quantise(const void* qkv, void* workspace, int n_seq, int S, int H, int head_dim, void* stream) {
  int rc = vq::ia_check(workspace, qkv, n_seq, S, H, head_dim);
  if (rc != VQ_OK) return rc;
  return vq::ia_quantise(qkv, workspace, n_seq, S, H, static_cast<cudaStream_t>(stream));
}

extern "C" int vq_attn_i8_attend(const void* workspace, void* out, int n_seq, int S, int H, int head_dim, float scale,
                                 void* stream) {
  int rc = vq::ia_check(workspace, out, n_seq, S, H, head_dim);
  if (rc != VQ_OK) return rc;
  return vq::ia_attend(workspace, out, n_seq, S, H, scale, static_cast<cudaStream_t>(stream));
}

extern "C" int vq_attn_spatial_i8(const void* qkv, void* out, void* workspace, int n_seq, int S, int H, int head_dim,
                                  float scale, void* stream) {
  int rc = vq::ia_check(workspace, qkv, n_seq, S, H, head_dim);
  if (rc != VQ_OK) return rc;
  if (!out || (reinterpret_cast<uintptr_t>(out) & 15)) return VQ_ERR_ARG;
  rc = vq::ia_quantise(qkv, workspace, n_seq, S, H, static_cast<cudaStream_t>(stream));
  if (rc != VQ_OK) return rc;
  return vq::ia_attend(workspace, out, n_seq, S, H, scale, static_cast<cudaStream_t>(stream));
}
