// Device helpers of the exact activation / weight quantiser, shared by the stand-alone quantise kernels (vq_quant.cu) and
// the one-launch fused QuantLinear (vq_linear.cu), whose producer warps quantise activation panels straight into shared memory.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

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
// row minimum and maximum from per-lane packed partials: one half2 (min, -max) shuffle chain instead of two fp32 ones
__device__ __forceinline__ void warp_minmax(__half2 mn2, __half2 mx2, float& mn, float& mx) {
  __half2 v = __halves2half2(__hmin(__low2half(mn2), __high2half(mn2)), __hneg(__hmax(__low2half(mx2), __high2half(mx2))));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint32_t u = __shfl_xor_sync(0xffffffffu, *reinterpret_cast<uint32_t*>(&v), o);
    v = __hmin2(v, *reinterpret_cast<__half2*>(&u));
  }
  mn = __low2float(v);
  mx = -__high2float(v);
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

// ---------------------------------------------------------------------------------------------------------------
// Exact, division-free quantisation on packed pairs.
//   q = clamp(rint(h(x / delta)) + zp, 0, qmax)          (x, delta fp16 values)
// (1) For fp16 operands (11-bit significands) the true quotient is either exactly an fp16 rounding midpoint or at
//     least 2^-23 (relative) away from one, so the reference's fp32-then-fp16 double rounding equals one direct
//     rounding, and a quotient with < 2^-24 relative error — one Newton step on x * (1/delta), the residual being exact
//     in one FMA — rounds to the same fp16 (validated against x/delta on 2e7 random pairs and every golden vector).
// (2) rint + zero-point + clamp run in fp16 with the 1.5*2^10 trick: h + 1536 rounds to an integer (RNE, ulp = 1 on
//     [1024, 2048)), adding (zp - 512) gives 1024 + rint + zp exactly, clamping to [1024, 1024 + qmax] leaves the code
//     in the low byte of each fp16 lane.  No F2I, two elements per instruction (FFMA2 / HADD2 / HMNMX2).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rcp_approx(float x) {   // MUFU.RCP, <= 1 ulp
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// fp16( a / b ) held in fp32, for fp16-representable a, b (or b a small integer): Newton-corrected reciprocal multiply;
// exact by the midpoint-distance argument above.
__device__ __forceinline__ float h_div(float a, float b) {
  const float r = rcp_approx(b);
  const float q0 = a * r;
  const float e = fmaf(-q0, b, a);
  return h_round(fmaf(e, r, q0));
}

struct QuantConsts {
  float2 delta2, rdelta2;   // delta and ~1/delta broadcast to both lanes
  __half2 zpm;              // zp - 512
  __half2 hi;               // 1024 + qmax
};

__device__ __forceinline__ QuantConsts make_consts(float delta, float zp, float qmax) {
  QuantConsts c;
  const float r = rcp_approx(delta);
  c.delta2 = make_float2(delta, delta);
  c.rdelta2 = make_float2(r, r);
  c.zpm = __float2half2_rn(zp - 512.0f);
  c.hi = __float2half2_rn(1024.0f + qmax);
  return c;
}

// two fp16 inputs -> 32-bit word whose 16-bit lanes hold 0x6400 + code
__device__ __forceinline__ uint32_t quant_pair(__half2 x, const QuantConsts& c) {
  const float2 xf = __half22float2(x);
  const float2 q0 = __fmul2_rn(xf, c.rdelta2);
  const float2 e = __ffma2_rn(make_float2(-q0.x, -q0.y), c.delta2, xf);
  const float2 q1 = __ffma2_rn(e, c.rdelta2, q0);
  __half2 y = __floats2half2_rn(q1.x, q1.y);
  y = __hadd2_rn(y, __float2half2_rn(1536.0f));
  y = __hadd2_rn(y, c.zpm);
  y = __hmin2(__hmax2(y, __float2half2_rn(1024.0f)), c.hi);
  return *reinterpret_cast<uint32_t*>(&y);
}

// 8 halves (one 16-byte chunk) -> 8 codes (two words); accumulates their sum via dp4a
__device__ __forceinline__ uint2 quant_chunk(const uint4& v, const QuantConsts& c, int& sum) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
  const uint32_t a = quant_pair(h[0], c), b = quant_pair(h[1], c), d = quant_pair(h[2], c), f = quant_pair(h[3], c);
  uint2 out;
  out.x = __byte_perm(a, b, 0x6420);
  out.y = __byte_perm(d, f, 0x6420);
  sum = static_cast<int>(__dp4a(out.x, 0x01010101u, static_cast<unsigned>(sum)));
  sum = static_cast<int>(__dp4a(out.y, 0x01010101u, static_cast<unsigned>(sum)));
  return out;
}

// h(x / s) for fp16 x, s: the same exact Newton-corrected reciprocal
__device__ __forceinline__ __half2 div_pair(__half2 x, __half2 s) {
  const float2 xf = __half22float2(x), sf = __half22float2(s);
  const float2 r = make_float2(rcp_approx(sf.x), rcp_approx(sf.y));
  const float2 q0 = __fmul2_rn(xf, r);
  const float2 e = __ffma2_rn(make_float2(-q0.x, -q0.y), sf, xf);
  const float2 q1 = __ffma2_rn(e, r, q0);
  return __floats2half2_rn(q1.x, q1.y);
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
  s.delta = h_div(range, qmax);
  s.degenerate = s.delta < 1e-6f;
  s.zp = rintf(h_div(-mn, s.delta));
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
      __half2* x = reinterpret_cast<__half2*>(&r.c[i]);
      const __half2* s = reinterpret_cast<const __half2*>(&sv);
#pragma unroll
      for (int e = 0; e < 4; ++e) x[e] = div_pair(x[e], s[e]);
    }
  }
}

// x <- h( h( LN(x) * h(1 + scale) ) + shift ), LN in fp32 with one rounding to fp16 (nn.LayerNorm on a half tensor).
template <int MAXC>
__device__ __forceinline__ void apply_ln_modulate(RowRegs<MAXC>& r, const __half* shift, const __half* scale, int K,
                                                  int nchunk, int lane) {
  float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) {
      const __half2* x = reinterpret_cast<const __half2*>(&r.c[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) sum2 = __fadd2_rn(sum2, __half22float2(x[e]));
    }
  }
  const float mean = warp_sum(sum2.x + sum2.y) / static_cast<float>(K);
  const float2 nmean2 = make_float2(-mean, -mean);
  float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) {
      const __half2* x = reinterpret_cast<const __half2*>(&r.c[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 d = __fadd2_rn(__half22float2(x[e]), nmean2);
        sq2 = __ffma2_rn(d, d, sq2);
      }
    }
  }
  const float var = warp_sum(sq2.x + sq2.y) / static_cast<float>(K);
  const float rstd = rsqrtf(var + 1e-6f);
  const float2 rstd2 = make_float2(rstd, rstd);
  const __half2 one = __float2half2_rn(1.0f);
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) {
      uint4 shv = __ldg(reinterpret_cast<const uint4*>(shift) + ci);
      uint4 scv = __ldg(reinterpret_cast<const uint4*>(scale) + ci);
      __half2* x = reinterpret_cast<__half2*>(&r.c[i]);
      const __half2* sh = reinterpret_cast<const __half2*>(&shv);
      const __half2* sc = reinterpret_cast<const __half2*>(&scv);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 ln = __fmul2_rn(__fadd2_rn(__half22float2(x[e]), nmean2), rstd2);
        const __half2 lnh = __floats2half2_rn(ln.x, ln.y);
        // each op rounded to fp16 separately, like the reference's half tensors (never contracted to an FMA)
        x[e] = __hadd2_rn(__hmul2_rn(lnh, __hadd2_rn(one, sc[e])), sh[e]);
      }
    }
  }
}

template <int MAXC>
__device__ __forceinline__ void row_minmax(const RowRegs<MAXC>& r, int nchunk, int lane, __half2& mn2, __half2& mx2) {
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) {
      const __half2* x = reinterpret_cast<const __half2*>(&r.c[i]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        mn2 = __hmin2(mn2, x[e]);
        mx2 = __hmax2(mx2, x[e]);
      }
    }
  }
}

template <int MAXC>
__device__ __forceinline__ int quant_store_row(const RowRegs<MAXC>& r, uint8_t* codes_row, int nchunk, int lane,
                                               const QuantConsts& qc) {
  int sum = 0;
#pragma unroll
  for (int i = 0; i < MAXC; ++i) {
    int ci = lane + 32 * i;
    if (ci < nchunk) reinterpret_cast<uint2*>(codes_row)[ci] = quant_chunk(r.c[i], qc, sum);
  }
  return sum;
}

// ---------------------------------------------------------------------------------------------------------------
// Branch-free row mapping for K = U * 128: the row is U x 32 units of 8 bytes (4 halves); lane l owns units
// l, l+32, ... — every lane does identical, fully unrolled work (K = 1152 -> U = 9, K = 4608 -> U = 36).
// ---------------------------------------------------------------------------------------------------------------
template <int U>
struct UnitRegs {
  uint2 u[U];
};

template <int U>
__device__ __forceinline__ void uload_row(UnitRegs<U>& r, const __half* row, int lane) {
#pragma unroll
  for (int i = 0; i < U; ++i) r.u[i] = __ldg(reinterpret_cast<const uint2*>(row) + lane + 32 * i);
}

// Token row gathered from a head-major attention output [n, H, S, 72]: unit u (4 halves) of token (b, s) belongs to
// head u / 18 and sits at ((b * H + head) * S + s) * 72 + (u % 18) * 4. `tok0` points at (b, head 0, s, 0).
template <int U>
__device__ __forceinline__ void uload_row_heads(UnitRegs<U>& r, const __half* tok0, int S, int lane) {
#pragma unroll
  for (int i = 0; i < U; ++i) {
    const int u = lane + 32 * i;
    const int head = u / 18, w = u - head * 18;
    r.u[i] = __ldg(reinterpret_cast<const uint2*>(tok0 + static_cast<size_t>(head) * S * 72) + w);
  }
}

template <int U>
__device__ __forceinline__ void uapply_smooth(UnitRegs<U>& r, const __half* smooth, int lane) {
#pragma unroll
  for (int i = 0; i < U; ++i) {
    const uint2 sv = __ldg(reinterpret_cast<const uint2*>(smooth) + lane + 32 * i);
    __half2* x = reinterpret_cast<__half2*>(&r.u[i]);
    const __half2* sm = reinterpret_cast<const __half2*>(&sv);
    x[0] = div_pair(x[0], sm[0]);
    x[1] = div_pair(x[1], sm[1]);
  }
}

template <int U>
__device__ __forceinline__ void uapply_ln_modulate(UnitRegs<U>& r, const __half* shift, const __half* scale, int K,
                                                   int lane) {
  float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < U; ++i) {
    const __half2* x = reinterpret_cast<const __half2*>(&r.u[i]);
    sum2 = __fadd2_rn(sum2, __half22float2(x[0]));
    sum2 = __fadd2_rn(sum2, __half22float2(x[1]));
  }
  const float mean = warp_sum(sum2.x + sum2.y) / static_cast<float>(K);
  const float2 nmean2 = make_float2(-mean, -mean);
  float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < U; ++i) {
    const __half2* x = reinterpret_cast<const __half2*>(&r.u[i]);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float2 d = __fadd2_rn(__half22float2(x[e]), nmean2);
      sq2 = __ffma2_rn(d, d, sq2);
    }
  }
  const float var = warp_sum(sq2.x + sq2.y) / static_cast<float>(K);
  const float rstd = rsqrtf(var + 1e-6f);
  const float2 rstd2 = make_float2(rstd, rstd);
  const __half2 one = __float2half2_rn(1.0f);
#pragma unroll
  for (int i = 0; i < U; ++i) {
    const uint2 shv = __ldg(reinterpret_cast<const uint2*>(shift) + lane + 32 * i);
    const uint2 scv = __ldg(reinterpret_cast<const uint2*>(scale) + lane + 32 * i);
    __half2* x = reinterpret_cast<__half2*>(&r.u[i]);
    const __half2* sh = reinterpret_cast<const __half2*>(&shv);
    const __half2* sc = reinterpret_cast<const __half2*>(&scv);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float2 ln = __fmul2_rn(__fadd2_rn(__half22float2(x[e]), nmean2), rstd2);
      const __half2 lnh = __floats2half2_rn(ln.x, ln.y);
      x[e] = __hadd2_rn(__hmul2_rn(lnh, __hadd2_rn(one, sc[e])), sh[e]);   // three separate fp16 roundings
    }
  }
}

template <int U>
__device__ __forceinline__ void urow_minmax(const UnitRegs<U>& r, __half2& mn2, __half2& mx2) {
#pragma unroll
  for (int i = 0; i < U; ++i) {
    const __half2* x = reinterpret_cast<const __half2*>(&r.u[i]);
    mn2 = __hmin2(mn2, __hmin2(x[0], x[1]));
    mx2 = __hmax2(mx2, __hmax2(x[0], x[1]));
  }
}

template <int U>
__device__ __forceinline__ int uquant_store_row(const UnitRegs<U>& r, uint8_t* codes_row, int lane,
                                                const QuantConsts& qc) {
  uint32_t sum = 0;
#pragma unroll
  for (int i = 0; i < U; ++i) {
    const __half2* x = reinterpret_cast<const __half2*>(&r.u[i]);
    const uint32_t w = __byte_perm(quant_pair(x[0], qc), quant_pair(x[1], qc), 0x6420);
    sum = __dp4a(w, 0x01010101u, sum);
    reinterpret_cast<uint32_t*>(codes_row)[lane + 32 * i] = w;
  }
  return static_cast<int>(sum);
}

}  // namespace vq
