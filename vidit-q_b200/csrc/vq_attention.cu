// Small-sequence attention kernels (head_dim 72, fp16 in/out, fp32 softmax) for the two STDiT attentions whose key
// length is tiny and which are therefore HBM-bound:
//   * temporal self-attention (reference stdit.py:112-118 -> blocks.py:151-195 on "(B S) T C"): T <= 16 keys per
//     (batch, spatial position, head). Reads q/k/v straight out of the fused q|k|v GEMM output in the (T S) token
//     layout through strides — the reference's rearrange copies do not exist here.
//   * cross attention (blocks.py:292-310, xformers BlockDiagonalMask): every image token of sample b attends to that
//     sample's (mask-selected) prompt tokens, L <= 128.
// One warp owns 16 query rows: S = Q K^T and O = P V run on mma.sync.m16n8k16 (f16 x f16 -> f32) with ldmatrix
// fragments from padded shared-memory tiles (pitch 88 halves: conflict-free); softmax in registers; the C fragments of
// S are reused as the A fragments of P. Keys are processed in 64-key chunks with online softmax.
// (The long spatial attention, S = 1024, stays on the library flash kernel for now; a tcgen05 version is a later row.)
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "vq_internal.h"
#include "vq_quant_common.cuh"

namespace vq {

constexpr int HD = 72;        // head dim
constexpr int HDP = 80;       // padded to a multiple of the MMA K (16)
constexpr int PITCH = 88;     // smem row pitch in halves (176 B: 8 consecutive rows hit 8 distinct 16B bank groups)
constexpr int ND = HD / 8;    // 9 output n-tiles

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const __half* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const __half* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], const __half* p) {
  uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_exp2(float x) {   // MUFU.EX2; exp2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Asynchronously copy `rows` rows of 72 halves (global row stride gstride elements) into a [tile_rows, PITCH] smem
// tile; rows >= rows and the pad columns 72..79 are zero-filled (cp.async with src-size 0). `nthr` threads cooperate.
// All copies of a thread are in flight together; the caller waits with cp_async_wait_all().
__device__ __forceinline__ void load_tile(__half* dst, const __half* src, long long gstride, int rows, int tile_rows,
                                          int tid, int nthr) {
  const int chunks = tile_rows * 10;  // 10 x 16 B per row: 9 data + 1 pad
  for (int c = tid; c < chunks; c += nthr) {
    const int r = c / 10, cc = c % 10;
    const bool real = cc < 9 && r < rows;
    const __half* g = real ? src + r * gstride + cc * 8 : src;
    cp_async16(dst + r * PITCH + cc * 8, g, real ? 16 : 0);
  }
}

// One warp: 16 query rows in sQ (rows >= nq are zero), keys/values in sK/sV (rows >= lk zero), processed in chunks of
// CK*16 keys with online softmax (running max / sum, accumulator rescale). Result rows [0, nq) are written to global
// through sQ as staging.
template <int CK, bool STORE = true>
__device__ __forceinline__ void warp_attend(__half* sQ, const __half* sK, const __half* sV, int lk, float scale_log2e,
                                            __half* out, long long ostride, int nq, int lane) {
  const int g = lane >> 2, t = lane & 3;
  float o[ND][4];
#pragma unroll
  for (int n = 0; n < ND; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int ntiles = (lk + 15) >> 4;
  for (int kt0 = 0; kt0 < ntiles; kt0 += CK) {
    // ---- S = Q K^T for this chunk: 2*CK n-tiles of 8 keys, 5 k-steps of 16 dims
    float s[2 * CK][4];
#pragma unroll
    for (int n = 0; n < 2 * CK; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HDP / 16; ++ks) {
      uint32_t a[4];
      ldsm_x4(a, sQ + ((lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH + ks * 16 + 8 * (lane >> 4));
#pragma unroll
      for (int kt = 0; kt < CK; ++kt) {
        if (kt0 + kt < ntiles) {   // warp-uniform
          uint32_t b[4];  // keys (kt0+kt)*16 + [0,8) / [8,16), dims ks*16 + [0,8) / [8,16)
          ldsm_x4(b, sK + ((kt0 + kt) * 16 + (lane & 7) + 8 * (lane >> 4)) * PITCH + ks * 16 + 8 * ((lane >> 3) & 1));
          mma16816(s[2 * kt], a, b[0], b[1]);
          mma16816(s[2 * kt + 1], a, b[2], b[3]);
        }
      }
    }
    // ---- online softmax (rows g and g+8 of this quad), fp32
    float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 2 * CK; ++n) {
      const int col = kt0 * 16 + n * 8 + 2 * t;
      if (col >= lk) s[n][0] = s[n][2] = -INFINITY;
      if (col + 1 >= lk) s[n][1] = s[n][3] = -INFINITY;
      c0 = fmaxf(c0, fmaxf(s[n][0], s[n][1]));
      c1 = fmaxf(c1, fmaxf(s[n][2], s[n][3]));
    }
    c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 1));
    c0 = fmaxf(c0, __shfl_xor_sync(0xffffffffu, c0, 2));
    c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 1));
    c1 = fmaxf(c1, __shfl_xor_sync(0xffffffffu, c1, 2));
    const float n0 = fmaxf(m0, c0), n1 = fmaxf(m1, c1);   // finite: every chunk holds at least one real key
    const float r0 = fast_exp2((m0 - n0) * scale_log2e), r1 = fast_exp2((m1 - n1) * scale_log2e);
    m0 = n0;
    m1 = n1;
    l0 *= r0;
    l1 *= r1;
#pragma unroll
    for (int n = 0; n < ND; ++n) {
      o[n][0] *= r0; o[n][1] *= r0; o[n][2] *= r1; o[n][3] *= r1;
    }
    uint32_t p[CK][4];
#pragma unroll
    for (int n = 0; n < 2 * CK; ++n) {
      float e0 = fast_exp2((s[n][0] - m0) * scale_log2e), e1 = fast_exp2((s[n][1] - m0) * scale_log2e);
      float e2 = fast_exp2((s[n][2] - m1) * scale_log2e), e3 = fast_exp2((s[n][3] - m1) * scale_log2e);
      l0 += e0 + e1;
      l1 += e2 + e3;
      // C fragments of key tiles (2kt, 2kt+1) are the A fragment of P for k-step kt
      p[n >> 1][(n & 1) * 2 + 0] = pack_h2(e0, e1);
      p[n >> 1][(n & 1) * 2 + 1] = pack_h2(e2, e3);
    }
    // ---- O += P V : 9 n-tiles of 8 dims, CK k-steps of 16 keys
#pragma unroll
    for (int kt = 0; kt < CK; ++kt) {
      if (kt0 + kt < ntiles) {
        const __half* vrow = sV + ((kt0 + kt) * 16 + (lane & 7) + 8 * ((lane >> 3) & 1)) * PITCH;
#pragma unroll
        for (int n = 0; n < ND - 1; n += 2) {
          uint32_t b[4];  // V^T fragments: keys [0,8)/[8,16) of the tile, dims n*8 + [0,8) / (n+1)*8 + [0,8)
          ldsm_x4_t(b, vrow + n * 8 + 8 * (lane >> 4));
          mma16816(o[n], p[kt], b[0], b[1]);
          mma16816(o[n + 1], p[kt], b[2], b[3]);
        }
        uint32_t b2[2];
        ldsm_x2_t(b2, vrow + (ND - 1) * 8);
        mma16816(o[ND - 1], p[kt], b2[0], b2[1]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  // ---- normalise, stage through sQ, coalesced 16-byte stores
  const float i0 = __fdividef(1.0f, l0), i1 = __fdividef(1.0f, l1);
  __syncwarp();
#pragma unroll
  for (int n = 0; n < ND; ++n) {
    *reinterpret_cast<uint32_t*>(sQ + g * PITCH + n * 8 + 2 * t) = pack_h2(o[n][0] * i0, o[n][1] * i0);
    *reinterpret_cast<uint32_t*>(sQ + (g + 8) * PITCH + n * 8 + 2 * t) = pack_h2(o[n][2] * i1, o[n][3] * i1);
  }
  __syncwarp();
  if (!STORE) return;      // the caller consumes the normalised fp16 rows from sQ (fused quantiser)
  for (int c = lane; c < 16 * 9; c += 32) {
    const int r = c / 9, cc = c % 9;
    if (r < nq) *(reinterpret_cast<uint4*>(out + r * ostride) + cc) = *reinterpret_cast<const uint4*>(sQ + r * PITCH + cc * 8);
  }
}

// ------------------------------------------------------------------------------------------------- temporal
struct TemporalArgs {
  const __half* qkv;   // [B, T, S, 3, H, 72] == fused q|k|v GEMM output [B*T*S, 3*H*72]
  __half* out;         // [B, T, S, H*72]
  int B, T, S, H;
  float scale_log2e;
};

constexpr int TEMPORAL_WARPS = 8;

__global__ void __launch_bounds__(TEMPORAL_WARPS * 32) vq_attn_temporal_kernel(const TemporalArgs a) {
  grid_dep_sync();
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __half* sQ = reinterpret_cast<__half*>(smem_attn) + warp * 3 * 16 * PITCH;
  __half* sK = sQ + 16 * PITCH;
  __half* sV = sK + 16 * PITCH;
  const long long unit = static_cast<long long>(blockIdx.x) * TEMPORAL_WARPS + warp;  // (b, s, h), h fastest
  const long long total = static_cast<long long>(a.B) * a.S * a.H;
  if (unit >= total) return;
  const int h = static_cast<int>(unit % a.H);
  const long long bs = unit / a.H;
  const int sidx = static_cast<int>(bs % a.S);
  const int b = static_cast<int>(bs / a.S);
  const int C = a.H * HD;
  const long long tok0 = static_cast<long long>(b) * a.T * a.S + sidx;   // token (b, t=0, s)
  const long long rstride = static_cast<long long>(a.S) * 3 * C;         // next frame, same spatial position
  const __half* q = a.qkv + tok0 * 3 * C + h * HD;
  // 16 rows x 10 sixteen-byte chunks (9 data + 1 zero pad) = 5 per lane; one (row, chunk) pattern serves q, k and v
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int c = lane + 32 * i;
    const int r = c / 10, cc = c - r * 10;
    const bool real = cc < 9 && r < a.T;
    const long long goff = real ? r * rstride + cc * 8 : 0;
    const int soff = r * PITCH + cc * 8;
    const int nbytes = real ? 16 : 0;
    cp_async16(sQ + soff, q + goff, nbytes);
    cp_async16(sK + soff, q + C + goff, nbytes);
    cp_async16(sV + soff, q + 2 * C + goff, nbytes);
  }
  cp_async_wait_all();
  __syncwarp();
  warp_attend<1>(sQ, sK, sV, a.T, a.scale_log2e, a.out + tok0 * C + h * HD, static_cast<long long>(a.S) * C, a.T, lane);
}

// ------------------------------------------------------------------------------------------------- temporal + quantiser
// Temporal attention with the projection's DynamicActQuantizer (a1) fused behind it: one 16-warp block owns ONE (batch,
// spatial position) — all H = 16 heads x all T <= 16 frames — so the T token rows it produces are complete (1152 channels)
// inside the block: after the attention every warp takes one token row out of the 16 heads' shared-memory tiles, computes
// its min / max, and writes u8 codes + (delta, zp, code sum).  The fp16 attention output (75 MB per launch at 32768 tokens)
// and the separate quantise pass that re-read it never exist; codes are bit-identical to vq_attn_temporal -> vq_act_quant.
struct TemporalQArgs {
  const __half* qkv;
  int B, T, S;
  float scale_log2e;
  const __half* smooth;    // [1152] or null
  float qmax;
  uint8_t* codes;          // [B*T*S, 1152], (T S) token order
  __half* delta;           // [B*T*S]
  __half* zp;
  int32_t* rowsum;
  uint32_t* status;
};

constexpr int TQ_HEADS = 16;

__global__ void __launch_bounds__(TQ_HEADS * 32, 1) vq_attn_temporal_quant_kernel(const TemporalQArgs a) {
  grid_dep_sync();
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __half* tiles = reinterpret_cast<__half*>(smem_attn);
  __half* sQ = tiles + warp * 3 * 16 * PITCH;
  __half* sK = sQ + 16 * PITCH;
  __half* sV = sK + 16 * PITCH;
  const int sidx = static_cast<int>(blockIdx.x % a.S);
  const int b = static_cast<int>(blockIdx.x / a.S);
  const int h = warp;
  constexpr int C = TQ_HEADS * HD;
  const long long tok0 = static_cast<long long>(b) * a.T * a.S + sidx;   // token (b, t = 0, s)
  const long long rstride = static_cast<long long>(a.S) * 3 * C;
  const __half* q = a.qkv + tok0 * 3 * C + h * HD;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const int c = lane + 32 * i;
    const int r = c / 10, cc = c - r * 10;
    const bool real = cc < 9 && r < a.T;
    const long long goff = real ? r * rstride + cc * 8 : 0;
    const int soff = r * PITCH + cc * 8;
    const int nbytes = real ? 16 : 0;
    cp_async16(sQ + soff, q + goff, nbytes);
    cp_async16(sK + soff, q + C + goff, nbytes);
    cp_async16(sV + soff, q + 2 * C + goff, nbytes);
  }
  cp_async_wait_all();
  __syncwarp();
  warp_attend<1, false>(sQ, sK, sV, a.T, a.scale_log2e, nullptr, 0, a.T, lane);
  __syncthreads();      // every head's normalised rows are in its sQ tile
  // ---- warp t quantises token row (b, t, s): unit u = lane + 32 i (4 halves) belongs to head u / 18
  const int t = warp;
  if (t >= a.T) return;
  UnitRegs<9> regs;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int u = lane + 32 * i;
    const int head = u / 18, w = u - head * 18;
    regs.u[i] = *reinterpret_cast<const uint2*>(tiles + head * 3 * 16 * PITCH + t * PITCH + w * 4);
  }
  if (a.smooth) uapply_smooth<9>(regs, a.smooth, lane);
  __half2 mn2 = __float2half2_rn(0.f), mx2 = mn2;   // the range always contains zero
  urow_minmax<9>(regs, mn2, mx2);
  float mn, mx;
  warp_minmax(mn2, mx2, mn, mx);
  const RowStats st = make_stats(mn, mx, a.qmax);
  const QuantConsts qc = make_consts(st.delta, st.zp, a.qmax);
  const long long m = tok0 + static_cast<long long>(t) * a.S;
  int sum = uquant_store_row<9>(regs, a.codes + m * C, lane, qc);
  sum = warp_sum_i(sum);
  if (lane == 0) {
    a.delta[m] = __float2half_rn(st.delta);
    a.zp[m] = __float2half_rn(st.zp);
    a.rowsum[m] = sum;
    if (st.degenerate && a.status) atomicOr(a.status, static_cast<uint32_t>(VQ_STATUS_EPS_DEGENERATE));
  }
}

// ------------------------------------------------------------------------------------------------- cross
struct CrossArgs {
  const __half* q;     // [B*N, H*72]
  const __half* kv;    // [sum(len), 2, H, 72]
  __half* out;         // [B*N, H*72]
  const int* kv_start; // [B] first prompt row of each sample
  const int* kv_len;   // [B] prompt length (<= 128)
  int B, N, H;
  float scale_log2e;
};

constexpr int CROSS_WARPS = 8;
constexpr int CROSS_KT = 8;   // 128 keys

// One CTA = one (sample, head) and a contiguous range of 16-query groups: K/V (<= 128 keys) are loaded once per CTA, every
// warp then walks its own groups with a double-buffered Q tile (cp.async of group i+1 in flight under the MMAs of group
// i), so there is no CTA-wide barrier after the first one and the K/V load is amortised over ~7 groups per warp.
__global__ void __launch_bounds__(CROSS_WARPS * 32, 2) vq_attn_cross_kernel(const CrossArgs a) {
  grid_dep_sync();
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __half* sK = reinterpret_cast<__half*>(smem_attn);
  __half* sV = sK + CROSS_KT * 16 * PITCH;
  __half* sQ = sV + CROSS_KT * 16 * PITCH + warp * 2 * 16 * PITCH;   // two buffers per warp
  const int h = blockIdx.y, b = blockIdx.z;
  const int C = a.H * HD;
  const int lk = a.kv_len[b];
  const __half* kbase = a.kv + static_cast<long long>(a.kv_start[b]) * 2 * C + h * HD;
  load_tile(sK, kbase, 2LL * C, lk, CROSS_KT * 16, threadIdx.x, CROSS_WARPS * 32);
  load_tile(sV, kbase + C, 2LL * C, lk, CROSS_KT * 16, threadIdx.x, CROSS_WARPS * 32);
  cp_async_commit();
  const int groups = (a.N + 15) >> 4;
  const int per_cta = (groups + gridDim.x - 1) / gridDim.x;
  const int g_end = min(groups, (static_cast<int>(blockIdx.x) + 1) * per_cta);
  int grp = blockIdx.x * per_cta + warp;
  const __half* qbase = a.q + static_cast<long long>(b) * a.N * C + h * HD;
  __half* obase = a.out + static_cast<long long>(b) * a.N * C + h * HD;
  if (grp < g_end) load_tile(sQ, qbase + static_cast<long long>(grp) * 16 * C, C, min(16, a.N - grp * 16), 16, lane, 32);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();
  int buf = 0;
  for (; grp < g_end; grp += CROSS_WARPS) {
    const int nxt = grp + CROSS_WARPS;
    if (nxt < g_end)
      load_tile(sQ + (buf ^ 1) * 16 * PITCH, qbase + static_cast<long long>(nxt) * 16 * C, C, min(16, a.N - nxt * 16), 16,
                lane, 32);
    cp_async_commit();
    warp_attend<4>(sQ + buf * 16 * PITCH, sK, sV, lk, a.scale_log2e, obase + static_cast<long long>(grp) * 16 * C, C,
                   min(16, a.N - grp * 16), lane);
    cp_async_wait_all();
    __syncwarp();
    buf ^= 1;
  }
}

}  // namespace vq

extern "C" int vq_attn_temporal(const void* qkv, void* out, int B, int T, int S, int H, int head_dim, float scale,
                                void* stream) {
  using namespace vq;
  if (!qkv || !out || B <= 0 || S <= 0 || H <= 0) return VQ_ERR_ARG;
  if (head_dim != HD || T <= 0 || T > 16) return VQ_ERR_UNSUPPORTED;
  TemporalArgs a{static_cast<const __half*>(qkv), static_cast<__half*>(out), B, T, S, H, scale * 1.4426950408889634f};
  const long long units = static_cast<long long>(B) * S * H;
  const int smem = TEMPORAL_WARPS * 3 * 16 * PITCH * 2;
  static bool attr_dev[kMaxDevices] = {};   // per-device function attribute (one process may drive several GPUs)
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    if (cudaFuncSetAttribute(vq_attn_temporal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return VQ_ERR_LAUNCH;
    attr = true;
  }
  const unsigned grid = static_cast<unsigned>((units + TEMPORAL_WARPS - 1) / TEMPORAL_WARPS);
  launch_pdl(vq_attn_temporal_kernel, dim3(grid), dim3(TEMPORAL_WARPS * 32), smem, static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

extern "C" int vq_attn_temporal_quant(const void* qkv, int B, int T, int S, int H, int head_dim, float scale,
                                      const void* smooth, int n_bits, uint8_t* codes, void* delta, void* zp,
                                      int32_t* rowsum, uint32_t* status, void* stream) {
  using namespace vq;
  if (!qkv || !codes || !delta || !zp || !rowsum || B <= 0 || S <= 0) return VQ_ERR_ARG;
  if (head_dim != HD || H != TQ_HEADS || T <= 0 || T > 16 || n_bits < 2 || n_bits > 8) return VQ_ERR_UNSUPPORTED;
  TemporalQArgs a{static_cast<const __half*>(qkv), B, T, S, scale * 1.4426950408889634f, static_cast<const __half*>(smooth),
                  static_cast<float>((1 << n_bits) - 1), codes, static_cast<__half*>(delta), static_cast<__half*>(zp),
                  rowsum, status};
  const int smem = TQ_HEADS * 3 * 16 * PITCH * 2;
  static bool attr_dev[kMaxDevices] = {};
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    if (cudaFuncSetAttribute(vq_attn_temporal_quant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return VQ_ERR_LAUNCH;
    attr = true;
  }
  const long long blocks = static_cast<long long>(B) * S;
  if (blocks > 0x7fffffffLL) return VQ_ERR_ARG;
  launch_pdl(vq_attn_temporal_quant_kernel, dim3(static_cast<unsigned>(blocks)), dim3(TQ_HEADS * 32), smem,
             static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

extern "C" int vq_attn_cross(const void* q, const void* kv, void* out, const int32_t* kv_start, const int32_t* kv_len,
                             int B, int N, int H, int head_dim, int max_len, int64_t kv_rows, float scale, void* stream) {
  using namespace vq;
  if (!q || !kv || !out || !kv_start || !kv_len || B <= 0 || N <= 0 || H <= 0) return VQ_ERR_ARG;
  if (head_dim != HD || max_len <= 0 || max_len > CROSS_KT * 16) return VQ_ERR_UNSUPPORTED;
  // image-sized query sets: the tcgen05 flash-attention kernel (two 64-key tiles, ragged prompt masked); the mma.sync
  // kernel below keeps the small / odd shapes (and is the bring-up reference: VQ_CROSS_TC=0)
  static const bool allow_tc = [] {
    const char* e = getenv("VQ_CROSS_TC");
    return !(e && e[0] == '0');
  }();
  if (allow_tc && kv_rows > 0 && (N % 256) == 0)
    return attn_cross_tc(q, kv, out, kv_start, kv_len, B, N, H, head_dim, kv_rows, scale, stream);
  CrossArgs a{static_cast<const __half*>(q), static_cast<const __half*>(kv), static_cast<__half*>(out), kv_start,
              kv_len, B, N, H, scale * 1.4426950408889634f};
  const int smem = (2 * CROSS_KT * 16 + 2 * CROSS_WARPS * 16) * PITCH * 2;
  static bool attr_dev[kMaxDevices] = {};   // per-device function attribute (one process may drive several GPUs)
  bool& attr = attr_dev[current_device()];
  if (!attr) {
    if (cudaFuncSetAttribute(vq_attn_cross_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return VQ_ERR_LAUNCH;
    attr = true;
  }
  // two CTAs per SM, all resident at once; each covers >= one 16-query group per warp
  const int groups = (N + 15) / 16;
  int chunks = 2 * num_sms() / (H * B);
  chunks = chunks < 1 ? 1 : chunks;
  const int max_chunks = (groups + CROSS_WARPS - 1) / CROSS_WARPS;
  if (chunks > max_chunks) chunks = max_chunks;
  dim3 grid(chunks, H, B);
  launch_pdl(vq_attn_cross_kernel, dim3(grid), dim3(CROSS_WARPS * 32), smem, static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}
