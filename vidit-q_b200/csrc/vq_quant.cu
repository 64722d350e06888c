// HBM-bound quantiser kernels for sm_100a.
//
//  * vq_act_quant             — DynamicActQuantizer.forward (reference qdiff/quantizer/dynamic_quantizer.py:16-45) with the
//                               'token' min/max init of base_quantizer.py:177-228, emitting u8 codes + per-token
//                               delta / zero-point + per-row code sums instead of the fake-quantised fp16 tensor.
//  * vq_ln_modulate_act_quant — the same with LayerNorm(eps=1e-6, no affine) + t2i_modulate (blocks.py:51) fused in
//                               front (stdit.py:104,125: the producers of the q/k/v and fc1 inputs).
//  * vq_prep_weight           — WeightQuantizer.forward (base_quantizer.py:129-144) with static per-channel delta/zp,
//                               run once at load instead of every forward (quant_layer.py:185).
//
// All arithmetic reproduces the reference's fp16 tensor semantics op by op (each torch op on a half tensor =
// fp32 compute + one round-to-nearest-even to fp16), so codes are bit-exact:
//     delta = h( h(max - min) / (2^b - 1) )          zp = rint( h( (-min) / delta ) )
//     q     = clamp( rint( h(x / delta) ) + zp, 0, 2^b - 1 )
// One warp owns one token row (all G pooled batch entries of it); the row lives in registers between the
// statistics pass and the quantise pass when G == 1, so x is read from HBM exactly once.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "vq_internal.h"
#include "vq_quant_common.cuh"

namespace vq {

struct ActQuantArgs {
  const __half* x;
  int G, rows, K;
  long long group_stride, ld;
  const __half* smooth;  // [K] or null
  const __half* shift;   // [G,K] (LN mode)
  const __half* scale;   // [G,K] (LN mode)
  int rows_per_mod;      // LN mode: row (g, r) uses modulation vector (g * rows + r) / rows_per_mod
  __half* y_out;         // optional [G*rows, K] transformed input (LN mode)
  int head_S;            // > 0: x is head-major [G*rows / head_S, H, head_S, 72] (attention output), H = K / 72
  const __half* addv;    // optional [add_period, K]: row r is quantised as h(x + addv[(r / rows_per_add) % add_period])
  int rows_per_add, add_period;
  int reverse;           // consume rows last-first (L2 reuse of the producer's tail); results do not depend on it
  float qmax;
  uint8_t* codes;
  __half* delta;
  __half* zp;
  int32_t* rowsum;
  uint32_t* status;
};

// x <- h(gelu_tanh(x)): the activation between fc1 and fc2 (timm Mlp.act, nn.GELU(approximate="tanh") on the fp16 fc1
// output), applied while the row is in registers on its way to fc2's quantiser — the fc1 GEMM then runs its plain
// bias epilogue (the GELU epilogue made that GEMM epilogue-bound) and this HBM-bound pass absorbs the MUFU work.
template <int MAXC>
__device__ __forceinline__ void apply_gelu(RowRegs<MAXC>& r, int nchunk, int lane) {
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) {
      __half2* x = reinterpret_cast<__half2*>(&r.c[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 g = gelu_tanh_pair(__half22float2(x[e]));
        x[e] = __floats2half2_rn(g.x, g.y);
      }
    }
  }
}

__device__ __forceinline__ size_t mod_offset(const ActQuantArgs& a, int g, int r) {
  return static_cast<size_t>((static_cast<long long>(g) * a.rows + r) / a.rows_per_mod) * a.K;
}

template <int MAXC, bool LN, bool GELU = false>
__global__ void __launch_bounds__(256) vq_act_quant_kernel(const ActQuantArgs a) {
  grid_dep_sync();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= a.rows) return;
  const int nchunk = a.K >> 3;
  RowRegs<MAXC> regs;
  __half2 mn2 = __float2half2_rn(0.f), mx2 = mn2;  // the range always contains zero
  for (int g = 0; g < a.G; ++g) {
    load_row<MAXC>(regs, a.x + g * a.group_stride + r * a.ld, nchunk, lane);
    if (LN) {
      apply_ln_modulate<MAXC>(regs, a.shift + mod_offset(a, g, r), a.scale + mod_offset(a, g, r),
                              a.K, nchunk, lane);
      if (a.smooth) apply_smooth<MAXC>(regs, a.smooth, nchunk, lane);
      if (a.y_out) {
        __half* yrow = a.y_out + (static_cast<size_t>(g) * a.rows + r) * a.K;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
          int ci = lane + 32 * i;
          if (ci < nchunk) reinterpret_cast<uint4*>(yrow)[ci] = regs.c[i];
        }
      }
    } else {
      if (GELU) apply_gelu<MAXC>(regs, nchunk, lane);
      if (a.smooth) apply_smooth<MAXC>(regs, a.smooth, nchunk, lane);
    }
    row_minmax<MAXC>(regs, nchunk, lane, mn2, mx2);
  }
  float mn, mx;
  warp_minmax(mn2, mx2, mn, mx);
  const RowStats st = make_stats(mn, mx, a.qmax);
  const QuantConsts qc = make_consts(st.delta, st.zp, a.qmax);
  if (lane == 0) {
    a.delta[r] = __float2half_rn(st.delta);
    a.zp[r] = __float2half_rn(st.zp);
    if (st.degenerate && a.status) atomicOr(a.status, static_cast<uint32_t>(VQ_STATUS_EPS_DEGENERATE));
  }
  for (int g = 0; g < a.G; ++g) {
    if (a.G > 1) {  // G == 1: the transformed row is still in registers
      load_row<MAXC>(regs, a.x + g * a.group_stride + r * a.ld, nchunk, lane);
      if (LN) {
        apply_ln_modulate<MAXC>(regs, a.shift + mod_offset(a, g, r), a.scale + mod_offset(a, g, r),
                                a.K, nchunk, lane);
        if (a.smooth) apply_smooth<MAXC>(regs, a.smooth, nchunk, lane);
      } else {
        if (GELU) apply_gelu<MAXC>(regs, nchunk, lane);
        if (a.smooth) apply_smooth<MAXC>(regs, a.smooth, nchunk, lane);
      }
    }
    const size_t orow = static_cast<size_t>(g) * a.rows + r;
    int s = quant_store_row<MAXC>(regs, a.codes + orow * a.K, nchunk, lane, qc);
    s = warp_sum_i(s);
    if (lane == 0) a.rowsum[orow] = s;
  }
}

// one row of the unit mapping from either a token-major or a head-major tensor
template <int U, bool HEADS>
__device__ __forceinline__ void uload_any(UnitRegs<U>& regs, const ActQuantArgs& a, int g, int r, int lane) {
  if (HEADS) {   // token r of sample g = (image b = r / S, position s = r % S) of a [*, H, S, 72] tensor
    const int bb = r / a.head_S, ss = r - bb * a.head_S;
    uload_row_heads<U>(regs, a.x + g * a.group_stride + (static_cast<size_t>(bb) * (a.K / 72) * a.head_S + ss) * 72,
                       a.head_S, lane);
  } else {
    uload_row<U>(regs, a.x + g * a.group_stride + r * a.ld, lane);
  }
}

template <int U>
__device__ __forceinline__ void uapply_add(UnitRegs<U>& r, const __half* addrow, int lane) {
#pragma unroll
  for (int i = 0; i < U; ++i) {
    const uint2 av = __ldg(reinterpret_cast<const uint2*>(addrow) + lane + 32 * i);
    __half2* x = reinterpret_cast<__half2*>(&r.u[i]);
    const __half2* ad = reinterpret_cast<const __half2*>(&av);
    x[0] = __hadd2_rn(x[0], ad[0]);
    x[1] = __hadd2_rn(x[1], ad[1]);
  }
}

template <int U, bool LN>
__device__ __forceinline__ void utransform(UnitRegs<U>& regs, const ActQuantArgs& a, int g, int r, int lane) {
  if (!LN && a.addv)   // the temporal position embedding of block 0 (stdit.py:113-115: x + tpe, an fp16 add), fused
    uapply_add<U>(regs, a.addv + static_cast<size_t>((r / a.rows_per_add) % a.add_period) * a.K, lane);
  if (LN) {
    uapply_ln_modulate<U>(regs, a.shift + mod_offset(a, g, r), a.scale + mod_offset(a, g, r), a.K, lane);
    if (a.smooth) uapply_smooth<U>(regs, a.smooth, lane);
    if (a.y_out) {
      uint2* yrow = reinterpret_cast<uint2*>(a.y_out + (static_cast<size_t>(g) * a.rows + r) * a.K);
#pragma unroll
      for (int i = 0; i < U; ++i) yrow[lane + 32 * i] = regs.u[i];
    }
  } else if (a.smooth) {
    uapply_smooth<U>(regs, a.smooth, lane);
  }
}

// Persistent: every warp walks rows r, r + W, r + 2W, ...  PF = true (for G == 1) keeps the NEXT row's loads in flight
// while the current one is quantised; PF = false spends those registers on a fourth resident block per SM instead (the
// default, see launch_act_quant).
template <int U, bool LN, bool HEADS, bool PF = true>
__global__ void __launch_bounds__(256, PF ? 3 : 4) vq_act_quant_unit_kernel(const ActQuantArgs a) {
  grid_dep_sync();
  const int lane = threadIdx.x & 31;
  const int wstride = gridDim.x * (blockDim.x >> 5);
  int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= a.rows) return;
  // a.reverse: rows are consumed LAST FIRST — the producer kernel wrote the tensor front to back, so its tail is what
  // still sits in the 126 MB L2 when this pass starts; reading forwards would evict that tail before reaching it
  const int rlast = a.rows - 1;
  UnitRegs<U> regs, nxt;
  if (PF && a.G == 1) uload_any<U, HEADS>(regs, a, 0, a.reverse ? rlast - q : q, lane);
  for (; q < a.rows; q += wstride) {
    const int r = a.reverse ? rlast - q : q;
    const bool has_next = PF && a.G == 1 && q + wstride < a.rows;
    if (!PF && a.G == 1) uload_any<U, HEADS>(regs, a, 0, r, lane);
    if (has_next) uload_any<U, HEADS>(nxt, a, 0, a.reverse ? rlast - (q + wstride) : q + wstride, lane);   // prefetch
    __half2 mn2 = __float2half2_rn(0.f), mx2 = mn2;  // the range always contains zero
    if (a.G == 1) {
      utransform<U, LN>(regs, a, 0, r, lane);
      urow_minmax<U>(regs, mn2, mx2);
    } else {
      for (int g = 0; g < a.G; ++g) {
        uload_any<U, HEADS>(regs, a, g, r, lane);
        utransform<U, LN>(regs, a, g, r, lane);
        urow_minmax<U>(regs, mn2, mx2);
      }
    }
    float mn, mx;
    warp_minmax(mn2, mx2, mn, mx);
    const RowStats st = make_stats(mn, mx, a.qmax);
    const QuantConsts qc = make_consts(st.delta, st.zp, a.qmax);
    if (lane == 0) {
      a.delta[r] = __float2half_rn(st.delta);
      a.zp[r] = __float2half_rn(st.zp);
      if (st.degenerate && a.status) atomicOr(a.status, static_cast<uint32_t>(VQ_STATUS_EPS_DEGENERATE));
    }
    for (int g = 0; g < a.G; ++g) {
      if (a.G > 1) {  // G == 1: the transformed row is still in registers
        uload_any<U, HEADS>(regs, a, g, r, lane);
        UnitRegs<U>& rr = regs;
        if (!LN && a.addv)
          uapply_add<U>(rr, a.addv + static_cast<size_t>((r / a.rows_per_add) % a.add_period) * a.K, lane);
        if (LN) {
          uapply_ln_modulate<U>(rr, a.shift + mod_offset(a, g, r), a.scale + mod_offset(a, g, r), a.K, lane);
          if (a.smooth) uapply_smooth<U>(rr, a.smooth, lane);
        } else if (a.smooth) {
          uapply_smooth<U>(rr, a.smooth, lane);
        }
      }
      const size_t orow = static_cast<size_t>(g) * a.rows + r;
      int s = uquant_store_row<U>(regs, a.codes + orow * a.K, lane, qc);
      s = warp_sum_i(s);
      if (lane == 0) a.rowsum[orow] = s;
    }
    if (has_next) regs = nxt;
  }
}

// K = 1152 * WPR (2304, 4608: the Mlp hidden width), G == 1: one warp per 1152-column SEGMENT of a row, so a lane holds 18
// registers of row data instead of 72 and five 8-warp blocks fit an SM; the WPR warps of a row combine their min / max
// and their code sums through shared memory (two block barriers per row batch).  GELU = true applies
// nn.GELU(approximate="tanh") to the loaded fp16 values first (MUFU-heavy: the occupancy is what hides it).
template <int U>
__device__ __forceinline__ void uapply_gelu(UnitRegs<U>& r) {
#pragma unroll
  for (int i = 0; i < U; ++i) {
    __half2* x = reinterpret_cast<__half2*>(&r.u[i]);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float2 g = gelu_tanh_pair(__half22float2(x[e]));
      x[e] = __floats2half2_rn(g.x, g.y);
    }
  }
}

template <int WPR, bool GELU, int OCC = 4>
__global__ void __launch_bounds__(256, OCC) vq_act_quant_seg_kernel(const ActQuantArgs a) {
  grid_dep_sync();
  constexpr int RPB = 8 / WPR;   // rows per block
  __shared__ float s_mn[8], s_mx[8];
  __shared__ int s_sum[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rl = warp / WPR, seg = warp - rl * WPR;
  const int r0 = blockIdx.x * RPB + rl;
  const bool ok = r0 < a.rows;
  const int r = (ok && a.reverse) ? a.rows - 1 - r0 : r0;   // last rows first: the producer's tail is what is still in L2
  UnitRegs<9> regs;
  __half2 mn2 = __float2half2_rn(0.f), mx2 = mn2;  // the range always contains zero
  if (ok) {
    uload_row<9>(regs, a.x + static_cast<size_t>(r) * a.ld + seg * 1152, lane);
    if (GELU) uapply_gelu<9>(regs);
    if (a.smooth) uapply_smooth<9>(regs, a.smooth + seg * 1152, lane);
    urow_minmax<9>(regs, mn2, mx2);
  }
  float mn, mx;
  warp_minmax(mn2, mx2, mn, mx);
  if (lane == 0) {
    s_mn[warp] = mn;
    s_mx[warp] = mx;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < WPR; ++j) {
    mn = fminf(mn, s_mn[rl * WPR + j]);
    mx = fmaxf(mx, s_mx[rl * WPR + j]);
  }
  const RowStats st = make_stats(mn, mx, a.qmax);
  const QuantConsts qc = make_consts(st.delta, st.zp, a.qmax);
  if (ok) {
    if (seg == 0 && lane == 0) {
      a.delta[r] = __float2half_rn(st.delta);
      a.zp[r] = __float2half_rn(st.zp);
      if (st.degenerate && a.status) atomicOr(a.status, static_cast<uint32_t>(VQ_STATUS_EPS_DEGENERATE));
    }
    int sum = uquant_store_row<9>(regs, a.codes + static_cast<size_t>(r) * a.K + seg * 1152, lane, qc);
    sum = warp_sum_i(sum);
    if (lane == 0) s_sum[warp] = sum;
  }
  __syncthreads();
  if (ok && seg == 0 && lane == 0) {
    int tot = 0;
#pragma unroll
    for (int j = 0; j < WPR; ++j) tot += s_sum[rl * WPR + j];
    a.rowsum[r] = tot;
  }
}

template <bool LN, bool GELU = false>
static int launch_act_quant(const ActQuantArgs& a_in, cudaStream_t st) {
  // VQ_AQ_REVERSE=0: consume rows front to back (A/B knob; default: last rows first)
  static const int reverse = [] {
    const char* e = getenv("VQ_AQ_REVERSE");
    return e ? atoi(e) : 1;
  }();
  ActQuantArgs a = a_in;
  a.reverse = reverse;
  const int nchunk = a.K >> 3;
  const int maxc = (nchunk + 31) / 32;
  const int warps = 8;
  dim3 grid((a.rows + warps - 1) / warps), block(warps * 32);
  if (!LN && a.G == 1 && a.head_S == 0 && (a.K == 4608 || a.K == 2304 || (GELU && a.K == 1152))) {
    const int wpr = a.K / 1152;
    dim3 g((a.rows * wpr + 7) / 8);
    // resident blocks per SM of the K = 4608 kernel (registers capped at 62 / 48 / 40): 5 measured best (GELU variant
    // 118.8 / 112.7 / 114.8 us, plain 86.0 / 79.9 / 77.9 us for 4 / 5 / 6 at M = 32768, cold L2); knob VQ_SEG_OCC
    static const int occ = [] {
      const char* e = getenv("VQ_SEG_OCC");
      return e ? atoi(e) : 5;
    }();
    if (wpr == 4 && occ == 5) launch_pdl(vq_act_quant_seg_kernel<4, GELU, 5>, g, block, 0, st, a);
    else if (wpr == 4 && occ == 6) launch_pdl(vq_act_quant_seg_kernel<4, GELU, 6>, g, block, 0, st, a);
    else if (wpr == 4) launch_pdl(vq_act_quant_seg_kernel<4, GELU>, g, block, 0, st, a);
    else if (wpr == 2) launch_pdl(vq_act_quant_seg_kernel<2, GELU>, g, block, 0, st, a);
    else launch_pdl(vq_act_quant_seg_kernel<1, GELU>, g, block, 0, st, a);
    return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
  }
  if (GELU) {   // other K (or pooled batches): chunk-mapped kernel
    if (maxc <= 5) launch_pdl(vq_act_quant_kernel<5, false, true>, grid, block, 0, st, a);
    else if (maxc <= 9) launch_pdl(vq_act_quant_kernel<9, false, true>, grid, block, 0, st, a);
    else if (maxc <= 18) launch_pdl(vq_act_quant_kernel<18, false, true>, grid, block, 0, st, a);
    else if (maxc <= 36) launch_pdl(vq_act_quant_kernel<36, false, true>, grid, block, 0, st, a);
    else return VQ_ERR_UNSUPPORTED;
    return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
  }
  // K = 1152 is 4.5 sixteen-byte chunks per lane (divergent in the chunk mapping) but exactly 9 eight-byte units;
  // K = 4608 is 18 full chunk rounds, where the chunk mapping is already branch-free and lighter on registers.
  if (a.K == 9 * 128) {
    const int blocks_needed = (a.rows + warps - 1) / warps;
    // Register prefetch of the next row (3 blocks per SM, 80 registers) against no prefetch and 4 blocks per SM (64
    // registers): inside the replayed step the extra warps win — plain 20.6 -> 18.1 us, LN 39.4 -> 36.8 us per launch at
    // M = 32768 (profiles/r01_s22_*).  VQ_AQ_NOPF selects per kernel (bit 0: plain, bit 1: LN; default both).
    static const int nopf = [] {
      const char* e = getenv("VQ_AQ_NOPF");
      return e ? atoi(e) : 3;
    }();
    const bool pf = !((nopf >> (LN ? 1 : 0)) & 1) || a.head_S > 0;
    const int persistent = num_sms() * (pf ? 3 : 4);   // resident 8-warp blocks per SM, rows strided across them
    const int g2 = blocks_needed < persistent ? blocks_needed : persistent;
    if (a.head_S > 0) launch_pdl(vq_act_quant_unit_kernel<9, LN, true>, g2, block, 0, st, a);
    else if (pf) launch_pdl(vq_act_quant_unit_kernel<9, LN, false>, g2, block, 0, st, a);
    else launch_pdl(vq_act_quant_unit_kernel<9, LN, false, false>, g2, block, 0, st, a);
  }
  else if (maxc <= 5) launch_pdl(vq_act_quant_kernel<5, LN>, grid, block, 0, st, a);
  else if (maxc <= 9) launch_pdl(vq_act_quant_kernel<9, LN>, grid, block, 0, st, a);
  else if (maxc <= 18) launch_pdl(vq_act_quant_kernel<18, LN>, grid, block, 0, st, a);
  else if (maxc <= 36) launch_pdl(vq_act_quant_kernel<36, LN>, grid, block, 0, st, a);
  else return VQ_ERR_UNSUPPORTED;
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

// ----------------------------------------------------------------------------- weight prep
struct PrepArgs {
  const __half* w;
  const __half* delta;
  const __half* zp;
  const __half* smooth;
  const __half* bias;
  int N, K;
  float qmax;
  uint8_t* codes;
  VqColParam* col;
};

__global__ void __launch_bounds__(256) vq_prep_weight_kernel(const PrepArgs a) {
  grid_dep_sync();
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= a.N) return;
  const float delta = __half2float(a.delta[n]);
  const float zp = __half2float(a.zp[n]);
  const QuantConsts qc = make_consts(delta, zp, a.qmax);
  const __half* wrow = a.w + static_cast<size_t>(n) * a.K;
  uint8_t* crow = a.codes + static_cast<size_t>(n) * a.K;
  // static (checkpoint) step sizes do not bound |w / delta|: pre-clamp the weight so that |w / delta| <= 511 keeps the
  // fp16 rounding trick exact; beyond that the code saturates to 0 / qmax either way
  const __half2 wl = __float2half2_rn(fminf(511.0f * delta, 65504.0f));
  int sum = 0;
  for (int ci = lane; ci < (a.K >> 3); ci += 32) {
    uint4 wv = __ldg(reinterpret_cast<const uint4*>(wrow) + ci);
    __half2* x = reinterpret_cast<__half2*>(&wv);
    if (a.smooth) {  // quant_layer.py:178 weight * channel_wise_scale (fp16 product)
      uint4 sv = __ldg(reinterpret_cast<const uint4*>(a.smooth) + ci);
      const __half2* sm = reinterpret_cast<const __half2*>(&sv);
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] = __hmul2_rn(x[e], sm[e]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) x[e] = __hmin2(__hmax2(x[e], __hneg2(wl)), wl);
    reinterpret_cast<uint2*>(crow)[ci] = quant_chunk(wv, qc, sum);
  }
  sum = warp_sum_i(sum);
  if (lane == 0) {
    VqColParam c;
    const int zw = __float2int_rn(zp);
    c.c1 = sum - a.K * zw;
    c.zw = zw;
    c.dw = delta;
    c.bias = a.bias ? __half2float(a.bias[n]) : 0.0f;
    a.col[n] = c;
  }
}

// ----------------------------------------------------------------------------- static (calibrated) activation scales
// BaseQuantizer.forward with init_done (base_quantizer.py:112-144) on an ActQuantizer whose delta / zero_point come from
// the PTQ checkpoint instead of the live tensor: per-tensor (`per_group: False`, w8a8_naive.yaml — one scalar pair) or
// static per-token ([rows] pairs, period = rows).  One streaming pass, one warp per row.
struct StaticQuantArgs {
  const __half* x;
  int M, K;
  long long ld;
  const __half* delta;   // [period]
  const __half* zp;      // [period]
  int period;            // row m uses index m % period
  const __half* smooth;  // [K] or null
  float qmax;
  uint8_t* codes;
  int32_t* rowsum;
  const __half* shift;   // LN variant: [M / rows_per_mod, K] modulation vectors
  const __half* scale;
  int rows_per_mod;
};

// GELU = true: nn.GELU(approximate="tanh") on the loaded values first (fc2's input in the fused schedule: no row statistics
// are needed with calibrated scales, so the row streams through chunk by chunk)
template <bool GELU>
__global__ void __launch_bounds__(256) vq_act_quant_static_kernel(const StaticQuantArgs a) {
  grid_dep_sync();
  const int lane = threadIdx.x & 31;
  const int wstride = gridDim.x * (blockDim.x >> 5);
  for (int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < a.M; m += wstride) {
    const int si = a.period == 1 ? 0 : m % a.period;
    const float delta = __half2float(a.delta[si]);
    const float zp = __half2float(a.zp[si]);
    const QuantConsts qc = make_consts(delta, zp, a.qmax);
    // calibrated step sizes do not bound |x / delta|: clamp so that the fp16 rounding trick stays exact (|x/delta| <=
    // 511); beyond that the code saturates to 0 / qmax either way
    const __half2 xl = __float2half2_rn(fminf(511.0f * delta, 65504.0f));
    const __half* xrow = a.x + static_cast<size_t>(m) * a.ld;
    uint8_t* crow = a.codes + static_cast<size_t>(m) * a.K;
    int sum = 0;
    for (int ci = lane; ci < (a.K >> 3); ci += 32) {
      uint4 xv = __ldg(reinterpret_cast<const uint4*>(xrow) + ci);
      __half2* x = reinterpret_cast<__half2*>(&xv);
      if (GELU) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 gl = gelu_tanh_pair(__half22float2(x[e]));
          x[e] = __floats2half2_rn(gl.x, gl.y);
        }
      }
      if (a.smooth) {   // quant_layer.py:140 input / channel_wise_scale
        uint4 sv = __ldg(reinterpret_cast<const uint4*>(a.smooth) + ci);
        const __half2* sm = reinterpret_cast<const __half2*>(&sv);
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = div_pair(x[e], sm[e]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] = __hmin2(__hmax2(x[e], __hneg2(xl)), xl);
      reinterpret_cast<uint2*>(crow)[ci] = quant_chunk(xv, qc, sum);
    }
    sum = warp_sum_i(sum);
    if (lane == 0) a.rowsum[m] = sum;
  }
}

// LayerNorm + t2i_modulate in front of the static quantiser (K <= 8 * 32 * MAXC: the row is held in registers for the two
// LayerNorm passes; same apply_ln_modulate as the dynamic chunk-mapped kernel)
template <int MAXC>
__global__ void __launch_bounds__(256) vq_ln_act_quant_static_kernel(const StaticQuantArgs a) {
  grid_dep_sync();
  const int lane = threadIdx.x & 31;
  const int nchunk = a.K >> 3;
  const int wstride = gridDim.x * (blockDim.x >> 5);
  for (int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < a.M; m += wstride) {
    const int si = a.period == 1 ? 0 : m % a.period;
    const float delta = __half2float(a.delta[si]);
    const QuantConsts qc = make_consts(delta, __half2float(a.zp[si]), a.qmax);
    const __half2 xl = __float2half2_rn(fminf(511.0f * delta, 65504.0f));
    RowRegs<MAXC> regs;
    load_row<MAXC>(regs, a.x + static_cast<size_t>(m) * a.ld, nchunk, lane);
    const size_t mo = static_cast<size_t>(m / a.rows_per_mod) * a.K;
    apply_ln_modulate<MAXC>(regs, a.shift + mo, a.scale + mo, a.K, nchunk, lane);
    if (a.smooth) apply_smooth<MAXC>(regs, a.smooth, nchunk, lane);
#pragma unroll
    for (int i = 0; i < MAXC; ++i) {
      __half2* x = reinterpret_cast<__half2*>(&regs.c[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] = __hmin2(__hmax2(x[e], __hneg2(xl)), xl);
    }
    int sum = quant_store_row<MAXC>(regs, a.codes + static_cast<size_t>(m) * a.K, nchunk, lane, qc);
    sum = warp_sum_i(sum);
    if (lane == 0) a.rowsum[m] = sum;
  }
}

// ----------------------------------------------------------------------------- per-channel |x| maxima (smooth-quant)
// `input.abs().max(dim=-2)[0]` of quant_layer.py:116,119 (and the three STDiT subclasses): the live activation statistic
// of smooth-quant's "dynamic" channel scale and of the running-stat EMA (quirk Q17: PixArt keeps
// smooth_quant_running_stat=True at inference, t2i/scripts/quant_txt2img.py:300).  x fp16 [G, n, K]; out[g, k] =
// max_r |x[g, r, k]| as the fp16 bit pattern widened to u32 (non-negative fp16 values order like unsigned integers), so
// row chunks combine with atomicMax; the caller zero-fills `out`.  gelu != 0: statistics of h(gelu_tanh(x)) — the
// fused schedules hand fc2 the pre-activation.
struct ColMaxArgs {
  const __half* x;
  int G, n, K, rows_per_block, gelu;
  uint32_t* out;
};

__global__ void __launch_bounds__(256) vq_col_absmax_kernel(const ColMaxArgs a) {
  grid_dep_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;   // 16-byte column chunk
  if (c >= (a.K >> 3)) return;
  const int g = blockIdx.z;
  const int r0 = blockIdx.y * a.rows_per_block;
  const int r1 = min(r0 + a.rows_per_block, a.n);
  const __half2 zero = __float2half2_rn(0.f);
  __half2 m[4] = {zero, zero, zero, zero};
  const uint4* base = reinterpret_cast<const uint4*>(a.x + (static_cast<size_t>(g) * a.n) * a.K) + c;
  const size_t pitch = static_cast<size_t>(a.K >> 3);
#pragma unroll 4
  for (int r = r0; r < r1; ++r) {
    uint4 v = __ldg(base + r * pitch);
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __half2 x = h[e];
      if (a.gelu) {
        const float2 gl = gelu_tanh_pair(__half22float2(x));
        x = __floats2half2_rn(gl.x, gl.y);
      }
      m[e] = __hmax2(m[e], __habs2(x));
    }
  }
  uint32_t* o = a.out + static_cast<size_t>(g) * a.K + c * 8;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const uint32_t w = *reinterpret_cast<uint32_t*>(&m[e]);
    atomicMax(o + 2 * e, w & 0xffffu);
    atomicMax(o + 2 * e + 1, w >> 16);
  }
}

// ----------------------------------------------------------------------------- row exchange packing (frame sharding)
// The frame-sharded forward moves per-token QUANTISED activations between ranks: each row
// travels as K code bytes followed by a 16-byte tail {delta fp16, zp fp16, rowsum i32, pad}.  The rows are a 4-D array
// (d0, d1, d2, d3); the exchange layouts are permutations of those dimensions, given as destination strides (in rows).
//   pack   : codes [rows, K] + delta / zp / rowsum arrays  ->  rows-with-tails at dst = i0 s0 + i1 s1 + i2 s2 + i3 s3
//   unpack : rows-with-tails (source order)                ->  codes [rows, K] + arrays at the permuted position
// One warp per row, 4-byte accesses: a single pass at HBM speed instead of the five strided byte copies (cat, permute,
// contiguous, two slices) the exchange cost as torch ops.
struct RowPackArgs {
  const uint8_t* src;      // pack: codes [rows, K];  unpack: rows with tails [rows, K + 16]
  uint8_t* dst;            // pack: rows with tails;   unpack: codes [rows, K]
  __half* delta;           // arrays: read by pack, written by unpack
  __half* zp;
  int32_t* rowsum;
  int rows, K;
  int d1, d2, d3;          // source dims (d0 implied)
  long long s0, s1, s2, s3;
  int unpack;
};

__global__ void __launch_bounds__(256) vq_row_pack_kernel(const RowPackArgs a) {
  grid_dep_sync();
  const int lane = threadIdx.x & 31;
  const int wstride = gridDim.x * (blockDim.x >> 5);
  const int words = a.K >> 2;
  const long long pitch_t = a.K + 16;
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < a.rows; r += wstride) {
    const int i3 = r % a.d3, q3 = r / a.d3;
    const int i2 = q3 % a.d2, q2 = q3 / a.d2;
    const int i1 = q2 % a.d1, i0 = q2 / a.d1;
    const long long dr = i0 * a.s0 + i1 * a.s1 + i2 * a.s2 + i3 * a.s3;
    if (!a.unpack) {
      const uint32_t* s = reinterpret_cast<const uint32_t*>(a.src + static_cast<long long>(r) * a.K);
      uint32_t* d = reinterpret_cast<uint32_t*>(a.dst + dr * pitch_t);
      for (int w = lane; w < words; w += 32) d[w] = __ldg(s + w);
      if (lane == 0) {
        const uint32_t dz = static_cast<uint32_t>(__half_as_ushort(a.delta[r])) |
                            (static_cast<uint32_t>(__half_as_ushort(a.zp[r])) << 16);
        *reinterpret_cast<uint4*>(a.dst + dr * pitch_t + a.K) = make_uint4(dz, static_cast<uint32_t>(a.rowsum[r]), 0u, 0u);
      }
    } else {
      const uint32_t* s = reinterpret_cast<const uint32_t*>(a.src + static_cast<long long>(r) * pitch_t);
      uint32_t* d = reinterpret_cast<uint32_t*>(a.dst + dr * a.K);
      for (int w = lane; w < words; w += 32) d[w] = __ldg(s + w);
      if (lane == 0) {
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(a.src + static_cast<long long>(r) * pitch_t + a.K));
        a.delta[dr] = __ushort_as_half(static_cast<unsigned short>(t.x & 0xffffu));
        a.zp[dr] = __ushort_as_half(static_cast<unsigned short>(t.x >> 16));
        a.rowsum[dr] = static_cast<int32_t>(t.y);
      }
    }
  }
}

}  // namespace vq

extern "C" int vq_row_pack(const uint8_t* src, uint8_t* dst, void* delta, void* zp, int32_t* rowsum, int rows, int K,
                           int d1, int d2, int d3, int64_t s0, int64_t s1, int64_t s2, int64_t s3, int unpack,
                           void* stream) {
  using namespace vq;
  if (!src || !dst || !delta || !zp || !rowsum || rows <= 0 || K <= 0 || (K % 16) != 0) return VQ_ERR_ARG;
  if (d1 <= 0 || d2 <= 0 || d3 <= 0 || rows % (d1 * d2 * d3) != 0) return VQ_ERR_ARG;
  RowPackArgs a{src, dst, static_cast<__half*>(delta), static_cast<__half*>(zp), rowsum, rows, K, d1, d2, d3, s0, s1, s2, s3,
                unpack};
  long long blocks = (rows + 7) / 8;
  const long long cap = 16LL * num_sms();
  if (blocks > cap) blocks = cap;
  launch_pdl(vq_row_pack_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

extern "C" int vq_col_absmax(const void* x, int G, int n, int K, int gelu, uint32_t* out_bits, void* stream) {
  using namespace vq;
  if (!x || !out_bits || G <= 0 || n <= 0 || K <= 0 || (K % 8) != 0) return VQ_ERR_ARG;
  ColMaxArgs a{static_cast<const __half*>(x), G, n, K, 64, gelu, out_bits};
  dim3 grid(((K >> 3) + 255) / 256, (n + a.rows_per_block - 1) / a.rows_per_block, G);
  if (grid.y > 65535u || grid.z > 65535u) return VQ_ERR_UNSUPPORTED;
  launch_pdl(vq_col_absmax_kernel, grid, dim3(256), 0, static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

extern "C" int vq_act_quant_static(const void* x, int M, int K, int64_t ld, const void* delta, const void* zp,
                                   int period, const void* smooth, int n_bits, uint8_t* codes, int32_t* rowsum,
                                   void* stream) {
  using namespace vq;
  if (!x || !codes || !delta || !zp || !rowsum || M <= 0 || K <= 0 || period <= 0) return VQ_ERR_ARG;
  if ((K % 8) != 0 || (ld % 8) != 0 || n_bits < 2 || n_bits > 8) return VQ_ERR_ARG;
  StaticQuantArgs a{static_cast<const __half*>(x), M, K, ld, static_cast<const __half*>(delta),
                    static_cast<const __half*>(zp), period, static_cast<const __half*>(smooth),
                    static_cast<float>((1 << n_bits) - 1), codes, rowsum, nullptr, nullptr, 1};
  const int warps = 8;
  long long blocks = (M + warps - 1) / warps;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  launch_pdl(vq_act_quant_static_kernel<false>, dim3(static_cast<unsigned>(blocks)), dim3(warps * 32), 0,
             static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

extern "C" int vq_gelu_act_quant_static(const void* x, int M, int K, int64_t ld, const void* delta, const void* zp,
                                        int period, const void* smooth, int n_bits, uint8_t* codes, int32_t* rowsum,
                                        void* stream) {
  using namespace vq;
  if (!x || !codes || !delta || !zp || !rowsum || M <= 0 || K <= 0 || period <= 0) return VQ_ERR_ARG;
  if ((K % 8) != 0 || (ld % 8) != 0 || n_bits < 2 || n_bits > 8) return VQ_ERR_ARG;
  StaticQuantArgs a{static_cast<const __half*>(x), M, K, ld, static_cast<const __half*>(delta),
                    static_cast<const __half*>(zp), period, static_cast<const __half*>(smooth),
                    static_cast<float>((1 << n_bits) - 1), codes, rowsum, nullptr, nullptr, 1};
  long long blocks = (M + 7) / 8;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  launch_pdl(vq_act_quant_static_kernel<true>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0,
             static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

extern "C" int vq_ln_modulate_act_quant_static(const void* x, const void* shift, const void* scale, int M, int K,
                                               int rows_per_mod, const void* delta, const void* zp, int period,
                                               const void* smooth, int n_bits, uint8_t* codes, int32_t* rowsum,
                                               void* stream) {
  using namespace vq;
  if (!x || !shift || !scale || !codes || !delta || !zp || !rowsum || M <= 0 || K <= 0 || period <= 0 || rows_per_mod <= 0)
    return VQ_ERR_ARG;
  if ((K % 8) != 0 || n_bits < 2 || n_bits > 8 || (M % rows_per_mod) != 0) return VQ_ERR_ARG;
  if (K > 8 * 32 * 9) return VQ_ERR_UNSUPPORTED;   // the row lives in registers (hidden sizes up to 2304)
  StaticQuantArgs a{static_cast<const __half*>(x), M, K, K, static_cast<const __half*>(delta),
                    static_cast<const __half*>(zp), period, static_cast<const __half*>(smooth),
                    static_cast<float>((1 << n_bits) - 1), codes, rowsum, static_cast<const __half*>(shift),
                    static_cast<const __half*>(scale), rows_per_mod};
  long long blocks = (M + 7) / 8;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  const dim3 grid(static_cast<unsigned>(blocks)), block(256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (K <= 8 * 32 * 5) launch_pdl(vq_ln_act_quant_static_kernel<5>, grid, block, 0, st, a);
  else launch_pdl(vq_ln_act_quant_static_kernel<9>, grid, block, 0, st, a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

extern "C" int vq_act_quant(const void* x, int G, int rows, int K, int64_t group_stride, int64_t ld,
                            const void* smooth, int n_bits, uint8_t* codes, void* delta, void* zp, int32_t* rowsum,
                            uint32_t* status, void* stream) {
  using namespace vq;
  if (!x || !codes || !delta || !zp || !rowsum || G <= 0 || rows <= 0 || K <= 0) return VQ_ERR_ARG;
  if ((K % 8) != 0 || (ld % 8) != 0 || (group_stride % 8) != 0 || n_bits < 2 || n_bits > 8) return VQ_ERR_ARG;
  ActQuantArgs a{};
  a.x = static_cast<const __half*>(x);
  a.G = G; a.rows = rows; a.K = K;
  a.group_stride = group_stride; a.ld = ld;
  a.smooth = static_cast<const __half*>(smooth);
  a.qmax = static_cast<float>((1 << n_bits) - 1);
  a.codes = codes;
  a.delta = static_cast<__half*>(delta);
  a.zp = static_cast<__half*>(zp);
  a.rowsum = rowsum;
  a.status = status;
  return launch_act_quant<false>(a, static_cast<cudaStream_t>(stream));
}

extern "C" int vq_add_act_quant(const void* x, const void* addv, int rows_per_add, int add_period, int G, int rows, int K,
                                const void* smooth, int n_bits, uint8_t* codes, void* delta, void* zp, int32_t* rowsum,
                                uint32_t* status, void* stream) {
  using namespace vq;
  if (!x || !addv || !codes || !delta || !zp || !rowsum || G <= 0 || rows <= 0) return VQ_ERR_ARG;
  if (rows_per_add <= 0 || add_period <= 0 || n_bits < 2 || n_bits > 8) return VQ_ERR_ARG;
  if (K != 9 * 128) return VQ_ERR_UNSUPPORTED;   // the branch-free K = 1152 kernel only
  ActQuantArgs a{};
  a.x = static_cast<const __half*>(x);
  a.G = G; a.rows = rows; a.K = K;
  a.group_stride = static_cast<long long>(rows) * K; a.ld = K;
  a.smooth = static_cast<const __half*>(smooth);
  a.addv = static_cast<const __half*>(addv);
  a.rows_per_add = rows_per_add;
  a.add_period = add_period;
  a.qmax = static_cast<float>((1 << n_bits) - 1);
  a.codes = codes;
  a.delta = static_cast<__half*>(delta);
  a.zp = static_cast<__half*>(zp);
  a.rowsum = rowsum;
  a.status = status;
  return launch_act_quant<false>(a, static_cast<cudaStream_t>(stream));
}

extern "C" int vq_gelu_act_quant(const void* x, int G, int rows, int K, int64_t group_stride, int64_t ld,
                                 const void* smooth, int n_bits, uint8_t* codes, void* delta, void* zp,
                                 int32_t* rowsum, uint32_t* status, void* stream) {
  using namespace vq;
  if (!x || !codes || !delta || !zp || !rowsum || G <= 0 || rows <= 0 || K <= 0) return VQ_ERR_ARG;
  if ((K % 8) != 0 || (ld % 8) != 0 || (group_stride % 8) != 0 || n_bits < 2 || n_bits > 8) return VQ_ERR_ARG;
  ActQuantArgs a{};
  a.x = static_cast<const __half*>(x);
  a.G = G; a.rows = rows; a.K = K;
  a.group_stride = group_stride; a.ld = ld;
  a.smooth = static_cast<const __half*>(smooth);
  a.qmax = static_cast<float>((1 << n_bits) - 1);
  a.codes = codes;
  a.delta = static_cast<__half*>(delta);
  a.zp = static_cast<__half*>(zp);
  a.rowsum = rowsum;
  a.status = status;
  return launch_act_quant<false, true>(a, static_cast<cudaStream_t>(stream));
}

extern "C" int vq_act_quant_heads(const void* x, int G, int rows, int H, int S, int head_dim, int n_bits,
                                  uint8_t* codes, void* delta, void* zp, int32_t* rowsum, uint32_t* status,
                                  void* stream) {
  using namespace vq;
  if (!x || !codes || !delta || !zp || !rowsum || G <= 0 || rows <= 0 || S <= 0 || (rows % S) != 0) return VQ_ERR_ARG;
  if (head_dim != 72 || H * head_dim != 9 * 128 || n_bits < 2 || n_bits > 8) return VQ_ERR_UNSUPPORTED;
  ActQuantArgs a{};
  a.x = static_cast<const __half*>(x);
  a.G = G; a.rows = rows; a.K = H * head_dim;
  a.group_stride = static_cast<long long>(rows) * a.K; a.ld = a.K;
  a.head_S = S;
  a.qmax = static_cast<float>((1 << n_bits) - 1);
  a.codes = codes;
  a.delta = static_cast<__half*>(delta);
  a.zp = static_cast<__half*>(zp);
  a.rowsum = rowsum;
  a.status = status;
  return launch_act_quant<false>(a, static_cast<cudaStream_t>(stream));
}

extern "C" int vq_ln_modulate_act_quant(const void* x, const void* shift, const void* scale, const void* smooth, int G,
                                        int rows, int K, int rows_per_mod, int n_bits, void* y_out, uint8_t* codes,
                                        void* delta, void* zp, int32_t* rowsum, uint32_t* status, void* stream) {
  using namespace vq;
  if (!x || !shift || !scale || !codes || !delta || !zp || !rowsum || G <= 0 || rows <= 0 || K <= 0)
    return VQ_ERR_ARG;
  if (rows_per_mod <= 0 || rows_per_mod > rows || (rows % rows_per_mod) != 0) return VQ_ERR_ARG;
  if ((K % 8) != 0 || n_bits < 2 || n_bits > 8) return VQ_ERR_ARG;
  ActQuantArgs a{};
  a.x = static_cast<const __half*>(x);
  a.G = G; a.rows = rows; a.K = K;
  a.group_stride = static_cast<long long>(rows) * K; a.ld = K;
  a.shift = static_cast<const __half*>(shift);
  a.scale = static_cast<const __half*>(scale);
  a.smooth = static_cast<const __half*>(smooth);
  a.rows_per_mod = rows_per_mod;
  a.y_out = static_cast<__half*>(y_out);
  a.qmax = static_cast<float>((1 << n_bits) - 1);
  a.codes = codes;
  a.delta = static_cast<__half*>(delta);
  a.zp = static_cast<__half*>(zp);
  a.rowsum = rowsum;
  a.status = status;
  return launch_act_quant<true>(a, static_cast<cudaStream_t>(stream));
}

extern "C" int vq_prep_weight(const void* w, const void* delta, const void* zp, const void* smooth, const void* bias,
                              int N, int K, int n_bits, uint8_t* codes, VqColParam* col, void* stream) {
  using namespace vq;
  if (!w || !delta || !zp || !codes || !col || N <= 0 || K <= 0) return VQ_ERR_ARG;
  if ((K % 8) != 0 || n_bits < 2 || n_bits > 8) return VQ_ERR_ARG;
  PrepArgs a{};
  a.w = static_cast<const __half*>(w);
  a.delta = static_cast<const __half*>(delta);
  a.zp = static_cast<const __half*>(zp);
  a.smooth = static_cast<const __half*>(smooth);
  a.bias = static_cast<const __half*>(bias);
  a.N = N; a.K = K;
  a.qmax = static_cast<float>((1 << n_bits) - 1);
  a.codes = codes;
  a.col = col;
  const int warps = 8;
  launch_pdl(vq_prep_weight_kernel, dim3((N + warps - 1) / warps), dim3(warps * 32), 0, static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

extern "C" int vq_status_read(const uint32_t* status_dev, uint32_t* host_out, void* stream) {
  if (!status_dev || !host_out) return VQ_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (cudaMemcpyAsync(host_out, status_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, st) != cudaSuccess)
    return VQ_ERR_LAUNCH;
  return cudaStreamSynchronize(st) == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}

extern "C" int vq_version(void) { return 100; }
extern "C" int vq_num_sms(void) { return vq::num_sms(); }
