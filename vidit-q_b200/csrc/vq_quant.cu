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

namespace vq {

__device__ __forceinline__ float h_round(float v) { return __half2float(__float2half_rn(v)); }

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// q = clamp(rint(h(x / delta)) + zp, 0, qmax); x, delta are fp16 values held in fp32, rdelta ~= 1/delta.
// Division-free but exact: for fp16 operands (11-bit significands) the true quotient is either exactly an fp16
// rounding midpoint or at least 2^-23 (relative) away from one, so (a) the reference's fp32-then-fp16 double rounding
// equals one direct rounding, and (b) a quotient with < 2^-24 relative error — one Newton step on x * (1/delta), the
// residual being exact in one FMA — rounds to the same fp16. (Checked against x/delta on 2e7 random pairs and all
// golden vectors; see tests.)  NaN (delta == 0, flagged as degenerate) converts to 0.
__device__ __forceinline__ int quant_code(float x, float delta, float rdelta, int zp, int qmax) {
  float q0 = x * rdelta;
  float e = fmaf(-q0, delta, x);
  float q1 = fmaf(e, rdelta, q0);
  int q = __half2int_rn(__float2half_rn(q1)) + zp;
  return min(max(q, 0), qmax);
}

struct RowStats {
  float delta;  // fp16 value
  float zp;     // integer
  bool degenerate;
};

__device__ __forceinline__ RowStats make_stats(float mn, float mx, float qmax) {
  // base_quantizer.py:191-194 (range always contains 0), :219 (delta), :221 (eps test), :228 (zero point)
  mn = fminf(mn, 0.0f);
  mx = fmaxf(mx, 0.0f);
  RowStats s;
  float range = h_round(mx - mn);
  s.delta = h_round(__fdiv_rn(range, qmax));
  s.degenerate = s.delta < 1e-6f;
  s.zp = rintf(h_round(__fdiv_rn(-mn, s.delta)));
  return s;
}

// A register-resident slice of one row: lane l holds 16-byte chunks l, l+32, ... (8 halves each).
template <int MAXC>
struct RowRegs {
  uint4 c[MAXC];
};

template <int MAXC>
__device__ __forceinline__ void load_row(RowRegs<MAXC>& r, const __half* row, int nchunk, int lane) {
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) r.c[i] = __ldg(reinterpret_cast<const uint4*>(row) + ci);
  }
}

// x <- h(x / s[k])   (quant_layer.py:140 `input = input / channel_wise_scale`)
template <int MAXC>
__device__ __forceinline__ void apply_smooth(RowRegs<MAXC>& r, const __half* smooth, int nchunk, int lane) {
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) {
      uint4 sv = __ldg(reinterpret_cast<const uint4*>(smooth) + ci);
      __half* x = reinterpret_cast<__half*>(&r.c[i]);
      const __half* s = reinterpret_cast<const __half*>(&sv);
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = __float2half_rn(__fdiv_rn(__half2float(x[e]), __half2float(s[e])));
    }
  }
}

// x <- h( h( LN(x) * h(1 + scale) ) + shift ), LN in fp32 with one rounding to fp16 (nn.LayerNorm on a half tensor).
template <int MAXC>
__device__ __forceinline__ void apply_ln_modulate(RowRegs<MAXC>& r, const __half* shift, const __half* scale, int K,
                                                  int nchunk, int lane) {
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) {
      const __half* x = reinterpret_cast<const __half*>(&r.c[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += __half2float(x[e]);
    }
  }
  const float mean = warp_sum(sum) / static_cast<float>(K);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) {
      const __half* x = reinterpret_cast<const __half*>(&r.c[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float d = __half2float(x[e]) - mean;
        sq = fmaf(d, d, sq);
      }
    }
  }
  const float var = warp_sum(sq) / static_cast<float>(K);
  const float rstd = 1.0f / sqrtf(var + 1e-6f);
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) {
      uint4 shv = __ldg(reinterpret_cast<const uint4*>(shift) + ci);
      uint4 scv = __ldg(reinterpret_cast<const uint4*>(scale) + ci);
      __half* x = reinterpret_cast<__half*>(&r.c[i]);
      const __half* sh = reinterpret_cast<const __half*>(&shv);
      const __half* sc = reinterpret_cast<const __half*>(&scv);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float ln = h_round((__half2float(x[e]) - mean) * rstd);
        float one_plus = h_round(1.0f + __half2float(sc[e]));
        float prod = h_round(ln * one_plus);
        x[e] = __float2half_rn(prod + __half2float(sh[e]));
      }
    }
  }
}

template <int MAXC>
__device__ __forceinline__ void row_minmax(const RowRegs<MAXC>& r, int nchunk, int lane, float& mn, float& mx) {
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) {
      const __half* x = reinterpret_cast<const __half*>(&r.c[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = __half2float(x[e]);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
      }
    }
  }
}

template <int MAXC>
__device__ __forceinline__ int quant_store_row(const RowRegs<MAXC>& r, uint8_t* codes_row, int nchunk, int lane,
                                               float delta, float zpf, float qmaxf) {
  const float rdelta = __frcp_rn(delta);
  const int zp = __float2int_rn(zpf);
  const int qmax = __float2int_rn(qmaxf);
  int sum = 0;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) {
      const __half* x = reinterpret_cast<const __half*>(&r.c[i]);
      uint32_t w[2] = {0u, 0u};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        int q = quant_code(__half2float(x[e]), delta, rdelta, zp, qmax);
        sum += q;
        w[e >> 2] |= static_cast<uint32_t>(q) << (8 * (e & 3));
      }
      reinterpret_cast<uint2*>(codes_row)[ci] = make_uint2(w[0], w[1]);
    }
  }
  return sum;
}

struct ActQuantArgs {
  const __half* x;
  int G, rows, K;
  long long group_stride, ld;
  const __half* smooth;  // [K] or null
  const __half* shift;   // [G,K] (LN mode)
  const __half* scale;   // [G,K] (LN mode)
  __half* y_out;         // optional [G*rows, K] transformed input (LN mode)
  float qmax;
  uint8_t* codes;
  __half* delta;
  __half* zp;
  int32_t* rowsum;
  uint32_t* status;
};

template <int MAXC, bool LN>
__global__ void __launch_bounds__(256) vq_act_quant_kernel(const ActQuantArgs a) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= a.rows) return;
  const int nchunk = a.K >> 3;
  RowRegs<MAXC> regs;
  float mn = 0.f, mx = 0.f;  // the range always contains zero
  for (int g = 0; g < a.G; ++g) {
    load_row<MAXC>(regs, a.x + g * a.group_stride + r * a.ld, nchunk, lane);
    if (LN) {
      apply_ln_modulate<MAXC>(regs, a.shift + static_cast<size_t>(g) * a.K, a.scale + static_cast<size_t>(g) * a.K,
                              a.K, nchunk, lane);
      if (a.y_out) {
        __half* yrow = a.y_out + (static_cast<size_t>(g) * a.rows + r) * a.K;
#pragma unroll
        for (int i = 0; i < MAXC; ++i) {
          int ci = lane + 32 * i;
          if (ci < nchunk) reinterpret_cast<uint4*>(yrow)[ci] = regs.c[i];
        }
      }
    } else if (a.smooth) {
      apply_smooth<MAXC>(regs, a.smooth, nchunk, lane);
    }
    row_minmax<MAXC>(regs, nchunk, lane, mn, mx);
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  const RowStats st = make_stats(mn, mx, a.qmax);
  if (lane == 0) {
    a.delta[r] = __float2half_rn(st.delta);
    a.zp[r] = __float2half_rn(st.zp);
    if (st.degenerate && a.status) atomicOr(a.status, static_cast<uint32_t>(VQ_STATUS_EPS_DEGENERATE));
  }
  for (int g = 0; g < a.G; ++g) {
    if (a.G > 1) {  // G == 1: the transformed row is still in registers
      load_row<MAXC>(regs, a.x + g * a.group_stride + r * a.ld, nchunk, lane);
      if (LN) {
        apply_ln_modulate<MAXC>(regs, a.shift + static_cast<size_t>(g) * a.K, a.scale + static_cast<size_t>(g) * a.K,
                                a.K, nchunk, lane);
      } else if (a.smooth) {
        apply_smooth<MAXC>(regs, a.smooth, nchunk, lane);
      }
    }
    const size_t orow = static_cast<size_t>(g) * a.rows + r;
    int s = quant_store_row<MAXC>(regs, a.codes + orow * a.K, nchunk, lane, st.delta, st.zp, a.qmax);
    s = warp_sum_i(s);
    if (lane == 0) a.rowsum[orow] = s;
  }
}

template <bool LN>
static int launch_act_quant(const ActQuantArgs& a, cudaStream_t st) {
  const int nchunk = a.K >> 3;
  const int maxc = (nchunk + 31) / 32;
  const int warps = 8;
  dim3 grid((a.rows + warps - 1) / warps), block(warps * 32);
  if (maxc <= 5) vq_act_quant_kernel<5, LN><<<grid, block, 0, st>>>(a);
  else if (maxc <= 9) vq_act_quant_kernel<9, LN><<<grid, block, 0, st>>>(a);
  else if (maxc <= 18) vq_act_quant_kernel<18, LN><<<grid, block, 0, st>>>(a);
  else if (maxc <= 36) vq_act_quant_kernel<36, LN><<<grid, block, 0, st>>>(a);
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
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= a.N) return;
  const float delta = __half2float(a.delta[n]);
  const float zp = __half2float(a.zp[n]);
  const float rdelta = __frcp_rn(delta);
  const int zpi = __float2int_rn(zp);
  const int qmaxi = __float2int_rn(a.qmax);
  const __half* wrow = a.w + static_cast<size_t>(n) * a.K;
  uint8_t* crow = a.codes + static_cast<size_t>(n) * a.K;
  int sum = 0;
  for (int ci = lane; ci < (a.K >> 3); ci += 32) {
    uint4 wv = __ldg(reinterpret_cast<const uint4*>(wrow) + ci);
    const __half* x = reinterpret_cast<const __half*>(&wv);
    uint4 sv = make_uint4(0, 0, 0, 0);
    if (a.smooth) sv = __ldg(reinterpret_cast<const uint4*>(a.smooth) + ci);
    const __half* s = reinterpret_cast<const __half*>(&sv);
    uint32_t w[2] = {0u, 0u};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = __half2float(x[e]);
      if (a.smooth) v = h_round(v * __half2float(s[e]));  // quant_layer.py:178 weight * channel_wise_scale (fp16)
      int q = quant_code(v, delta, rdelta, zpi, qmaxi);
      sum += q;
      w[e >> 2] |= static_cast<uint32_t>(q) << (8 * (e & 3));
    }
    reinterpret_cast<uint2*>(crow)[ci] = make_uint2(w[0], w[1]);
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

}  // namespace vq

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

extern "C" int vq_ln_modulate_act_quant(const void* x, const void* shift, const void* scale, int G, int rows, int K,
                                        int n_bits, void* y_out, uint8_t* codes, void* delta, void* zp,
                                        int32_t* rowsum, uint32_t* status, void* stream) {
  using namespace vq;
  if (!x || !shift || !scale || !codes || !delta || !zp || !rowsum || G <= 0 || rows <= 0 || K <= 0)
    return VQ_ERR_ARG;
  if ((K % 8) != 0 || n_bits < 2 || n_bits > 8) return VQ_ERR_ARG;
  ActQuantArgs a{};
  a.x = static_cast<const __half*>(x);
  a.G = G; a.rows = rows; a.K = K;
  a.group_stride = static_cast<long long>(rows) * K; a.ld = K;
  a.shift = static_cast<const __half*>(shift);
  a.scale = static_cast<const __half*>(scale);
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
  vq_prep_weight_kernel<<<(N + warps - 1) / warps, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(a);
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
