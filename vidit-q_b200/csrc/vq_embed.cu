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
constexpr int PE_TOKENS = 16;        // tokens per chunk; persistent CTAs walk chunks blockIdx.x, + gridDim.x, ...
constexpr int PE_CPT = 4;            // channels per thread

struct PatchEmbedArgs {
  const float* latent;   // [B, Cin, T, Hh, Ww] fp32
  const __half* weight;  // [C, Cin * ph * pw] (patch depth 1)
  const __half* bias;    // [C]
  const __half* pos;     // [S, C] or null
  __half* out;           // [B, T * S, C]
  int B, Cin, T, Hh, Ww, ph, pw, C;
};

// Round 2: persistent CTAs (two per SM) that load their 4 x 16 weights once and walk 16-token chunks (2048 chunks over 296
// CTAs: 99 % balanced; round 1 launched 512 CTAs of 64 tokens = 3.46 waves), and the 64 multiply-adds per token and thread
// issued as 32 packed FFMA2 (two channels per instruction; same sequential fp32 sums, bit-identical): the kernel was
// instruction-bound at 64 us for a 75 MB write.
__global__ void __launch_bounds__(320, 2) vq_patch_embed_kernel(const PatchEmbedArgs a) {
  grid_dep_sync();
  __shared__ __align__(16) float patch[2][PE_TOKENS][PE_MAXK];
  const int gw = a.Ww / a.pw, gh = a.Hh / a.ph;
  const int S = gh * gw;
  const int K = a.Cin * a.ph * a.pw;
  const long long M = static_cast<long long>(a.B) * a.T * S;
  const long long n_chunks = (M + PE_TOKENS - 1) / PE_TOKENS;
  const int c0 = threadIdx.x * PE_CPT;
  const bool active = c0 < a.C;
  // weights of this thread's 4 channels as channel PAIRS: w2[p][k] = {w[2p][k], w[2p+1][k]}
  float2 w2[PE_CPT / 2][PE_MAXK];
  float2 bias2[PE_CPT / 2];
  if (active) {
#pragma unroll
    for (int j = 0; j < PE_CPT; ++j) {
      const float bj = a.bias ? __half2float(a.bias[c0 + j]) : 0.f;
      if (j & 1) bias2[j >> 1].y = bj;
      else bias2[j >> 1].x = bj;
#pragma unroll
      for (int k = 0; k < PE_MAXK; ++k) {
        const float wv = k < K ? __half2float(__ldg(a.weight + static_cast<size_t>(c0 + j) * K + k)) : 0.f;
        if (j & 1) w2[j >> 1][k].y = wv;
        else w2[j >> 1][k].x = wv;
      }
    }
  }
  int buf = 0;
  for (long long chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x, buf ^= 1) {
    const long long tok0 = chunk * PE_TOKENS;
    // gather the patches of this chunk's tokens (fp32 latent -> fp16 rounding, as x.to(dtype) does); double-buffered so one
    // barrier per chunk suffices
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
      patch[buf][tl][k] = v;
    }
    __syncthreads();
    if (!active) continue;
    const int ntok = static_cast<int>(M - tok0 < PE_TOKENS ? M - tok0 : PE_TOKENS);
    int s_idx = static_cast<int>(tok0 % S);   // spatial position of the token, kept incrementally (no division in the loop)
    __half* orow = a.out + static_cast<size_t>(tok0) * a.C + c0;
    for (int tl = 0; tl < ntok; ++tl, orow += a.C) {
      float2 acc01 = make_float2(0.f, 0.f), acc23 = make_float2(0.f, 0.f);
#pragma unroll
      for (int k4 = 0; k4 < PE_MAXK; k4 += 4) {
        const float4 p4 = *reinterpret_cast<const float4*>(&patch[buf][tl][k4]);   // broadcast 16-byte shared load
        const float p[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 pp = make_float2(p[e], p[e]);
          acc01 = __ffma2_rn(pp, w2[0][k4 + e], acc01);
          acc23 = __ffma2_rn(pp, w2[1][k4 + e], acc23);
        }
      }
      __half2 h01 = __floats2half2_rn(acc01.x + bias2[0].x, acc01.y + bias2[0].y);
      __half2 h23 = __floats2half2_rn(acc23.x + bias2[1].x, acc23.y + bias2[1].y);
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
  const long long chunks = (M + PE_TOKENS - 1) / PE_TOKENS;
  const long long persistent = 2LL * num_sms();
  const unsigned grid = static_cast<unsigned>(chunks < persistent ? chunks : persistent);
  launch_pdl(vq_patch_embed_kernel, dim3(grid), dim3(320), 0, static_cast<cudaStream_t>(stream), a);
  return cudaGetLastError() == cudaSuccess ? VQ_OK : VQ_ERR_LAUNCH;
}
