// Patch embedding of the latent, fused with the spatial position embedding:
//   STDiT   x = x_embedder(x) (PatchEmbed3D: Conv3d kernel = stride = (1, 2, 2), blocks.py:60-110) ; rearrange "B (T S) C" ;
//           x = x + pos_embed                                              (stdit.py:255-258)
//   PixArt  x = x_embedder(x) (Conv2d kernel = stride = 2) + pos_embed     (PixArtMS.py:150-160)   — the T = 1 case
// i.e. per token 16 multiply-adds per channel.  The reference pays a cuDNN implicit-GEMM convolution, two layout
// conversion kernels, the "B C T H W -> B (T H W) C" transpose and a broadcast add (each a pass over the 37.7 MB hidden
// tensor); here it is one pass that only writes the hidden tensor.  Arithmetic as in the reference's fp16 graph: the fp32
// latent is rounded to fp16 (x.to(dtype)), the convolution accumulates in fp32 and rounds (+ bias) to fp16, the position
// embedding is a separate fp16 add.  (Summation order inside the 16-term dot product may differ from cuDNN's: last-bit.)
// HBM-bound: writes M * C * 2 bytes.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "vq_internal.h"

namespace vq {

constexpr int PE_MAXK = 16;          // in_channels * patch_h * patch_w
constexpr int PE_TOKENS = 64;        // tokens per CTA
constexpr int PE_CPT = 4;            // channels per thread

struct PatchEmbedArgs {
  const float* latent;   // [B, Cin, T, Hh, Ww] fp32
  const __half* weight;  // [C, Cin * ph * pw] (patch depth 1)
  const __half* bias;    // [C]
  const __half* pos;     // [S, C] or null
  __half* out;           // [B, T * S, C]
  int B, Cin, T, Hh, Ww, ph, pw, C;
};

__global__ void __launch_bounds__(320) vq_patch_embed_kernel(const PatchEmbedArgs a) {
  grid_dep_sync();
  __shared__ __align__(16) float patch[PE_TOKENS][PE_MAXK];
  const int gw = a.Ww / a.pw, gh = a.Hh / a.ph;
  const int S = gh * gw;
  const int K = a.Cin * a.ph * a.pw;
  const long long M = static_cast<long long>(a.B) * a.T * S;
  const long long tok0 = static_cast<long long>(blockIdx.x) * PE_TOKENS;
  // gather the patches of this CTA's tokens (fp32 latent -> fp16 rounding, as x.to(dtype) does)
  for (int i = threadIdx.x; i < PE_TOKENS * K; i += blockDim.x) {
    const int tl = i / K, k = i - tl * K;
    const long long tok = tok0 + tl;
    float v = 0.f;
    if (tok < M) {
      const int s = static_cast<int>(tok % S);
      const long long bt = tok / S;
      const int t = static_cast<int>(bt % a.T);
      const int b = static_cast<int>(bt / a.T);
      const int ci = k / (a.ph * a.pw), r = k - ci * (a.ph * a.pw);
      const int dy = r / a.pw, dx = r - dy * a.pw;
      const int hy = s / gw, wx = s - hy * gw;
      const long long off = (((static_cast<long long>(b) * a.Cin + ci) * a.T + t) * a.Hh + hy * a.ph + dy) * a.Ww + wx * a.pw + dx;
      v = __half2float(__float2half_rn(__ldg(a.latent + off)));
    }
    patch[tl][k] = v;
  }
  const int c0 = threadIdx.x * PE_CPT;
  const bool active = c0 < a.C;
  float w[PE_CPT][PE_MAXK];
  float bias[PE_CPT];
  if (active) {
#pragma unroll
    for (int j = 0; j < PE_CPT; ++j) {
      bias[j] = a.bias ? __half2float(a.bias[c0 + j]) : 0.f;
      if (K == PE_MAXK) {   // the thread's 4 x 16 weights are 128 contiguous bytes: 16-byte loads
        const uint4* wp = reinterpret_cast<const uint4*>(a.weight + static_cast<size_t>(c0 + j) * PE_MAXK);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const uint4 wv = __ldg(wp + v);
          const __half2* h2 = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(h2[e]);
            w[j][v * 8 + 2 * e] = f.x;
            w[j][v * 8 + 2 * e + 1] = f.y;
          }
        }
      } else {
#pragma unroll
        for (int k = 0; k < PE_MAXK; ++k) w[j][k] = k < K ? __half2float(a.weight[static_cast<size_t>(c0 + j) * K + k]) : 0.f;
      }
    }
  }
  __syncthreads();
  if (!active) return;
  const int ntok = static_cast<int>(M - tok0 < PE_TOKENS ? M - tok0 : PE_TOKENS);
  int s_idx = static_cast<int>(tok0 % S);   // spatial position of the token, kept incrementally (no division in the loop)
  __half* orow = a.out + static_cast<size_t>(tok0) * a.C + c0;
  for (int tl = 0; tl < ntok; ++tl, orow += a.C) {
    float acc[PE_CPT];
#pragma unroll
    for (int j = 0; j < PE_CPT; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k4 = 0; k4 < PE_MAXK; k4 += 4) {
      const float4 p4 = *reinterpret_cast<const float4*>(&patch[tl][k4]);   // broadcast 16-byte shared load
      const float p[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int j = 0; j < PE_CPT; ++j) acc[j] = fmaf(p[e], w[j][k4 + e], acc[j]);
    }
    __half2 h01 = __floats2half2_rn(acc[0] + bias[0], acc[1] + bias[1]);
    __half2 h23 = __floats2half2_rn(acc[2] + bias[2], acc[3] + bias[3]);
    if (a.pos) {   // separate fp16 add, as the reference's `x + pos_embed` on half tensors
      const uint2 pv = __ldg(reinterpret_cast<const uint2*>(a.pos + static_cast<size_t>(s_idx) * a.C + c0));
      h01 = __hadd2_rn(h01, *reinterpret_cast<const __half2*>(&pv.x));
      h23 = __hadd2_rn(h23, *reinterpret_cast<const __half2*>(&pv.y));
    }
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&h01);
    o.y = *reinterpret_cast<const uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(orow) = o;
    if (++s_idx == S) s_idx = 0;
  }
}

}  // namespace vq

extern "C" int vq_patch_embed(const float* latent, const void* weight, const void* bias, const void* pos, int B, int Cin,
                              int T, int Hh, int Ww, int ph, int pw, int C, void* out, void* stream) {
  using namespace vq;
  if (!latent || !weight || !out || B <= 0 || Cin <= 0 || T <= 0 || Hh <= 0 || Ww <= 0 || ph <= 0 || pw <= 0 || C <= 0)
    return VQ_ERR_ARG;
  if (Cin * ph * pw > PE_MAXK || (Hh % ph) != 0 || (Ww % pw) != 0 || (C % PE_CPT) != 0 || C > 320 * PE_CPT)
    return VQ_ERR_UNSUPPORTED;
  PatchEmbedArgs a{latent, static_cast<const __half*>(weight), static_cast<const __half*>(bias),
                   static_cast<const __half*>(pos), static_cast<__half*>(out), B, Cin, T, Hh, Ww, ph, pw, C};
  const long long M = static_cast<long long>(B) * T * (Hh / ph) * (Ww / pw);
  const unsigned grid = static_cast<unsigned>((M + PE_TOKENS - 1) / PE_TOKENS);
  launch_pdl(vq_patch_embed_kernel, dim3(grid), dim3(320), 0, static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}
