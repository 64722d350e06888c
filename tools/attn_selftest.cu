// On-device bring-up test + timing of vq_attn_spatial (links libviditq_b200.so). The check is a naive fp32 attention (one
// thread per query row and head). Not part of the product.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/attn_selftest tools/attn_selftest.cu \
//          -Lvidit-q_b200 -lviditq_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../vidit-q_b200'
//   usage: attn_selftest [--time] [--modes 3,2,0] [--lbo N]   (2 = P := identity, 3 = also check the raw scores)
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

extern "C" int vq_attn_spatial(const void* qkv, void* out, int n_seq, int S, int H, int head_dim, float scale, void* stream);
extern "C" int vq_attn_spatial_debug(const void* qkv, void* out, int n_seq, int S, int H, int head_dim, float scale,
                                     int debug, float* dbg, unsigned v_lbo, void* stream);

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                   \
    }                                                                            \
  } while (0)

constexpr int D = 72;

// mode 0: softmax(scale q k^T) v; 1: mean of v over the sequence; 2: v of key (row % 128)
__global__ void ref_attn(const __half* qkv, float* out, int n_seq, int S, int H, float scale, int mode) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long total = (long long)n_seq * S * H;
  if (idx >= total) return;
  const int h = idx % H;
  const long long tok = idx / H;
  const int seq = tok / S;
  const int C = H * D;
  const __half* q = qkv + tok * 3 * C + h * D;
  const __half* kbase = qkv + (long long)seq * S * 3 * C + C + h * D;
  const __half* vbase = kbase + C;
  float o[D];
  for (int d = 0; d < D; ++d) o[d] = 0.f;
  if (mode == 2) {
    const int key = (int)(tok % S) % 128;
    for (int d = 0; d < D; ++d) out[idx * D + d] = __half2float(vbase[(long long)key * 3 * C + d]);
    return;
  }
  float qf[D];
  for (int d = 0; d < D; ++d) qf[d] = __half2float(q[d]);
  float mx = -INFINITY;
  if (mode == 0) {
    for (int k = 0; k < S; ++k) {
      float s = 0.f;
      for (int d = 0; d < D; ++d) s += qf[d] * __half2float(kbase[(long long)k * 3 * C + d]);
      mx = fmaxf(mx, s * scale);
    }
  }
  float l = 0.f;
  for (int k = 0; k < S; ++k) {
    float p = 1.0f;
    if (mode == 0) {
      float s = 0.f;
      for (int d = 0; d < D; ++d) s += qf[d] * __half2float(kbase[(long long)k * 3 * C + d]);
      p = expf(s * scale - mx);
    }
    l += p;
    for (int d = 0; d < D; ++d) o[d] += p * __half2float(vbase[(long long)k * 3 * C + d]);
  }
  for (int d = 0; d < D; ++d) out[idx * D + d] = o[d] / l;
}

// raw scores of the first key tile: ref[(item*2+t)*128+row][key]
__global__ void ref_scores(const __half* qkv, float* out, int n_seq, int S, int H) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // (item, t, row, key)
  const int nqp = S / 256;
  const long long total = (long long)n_seq * H * nqp * 2 * 128 * 128;
  if (idx >= total) return;
  const int key = idx % 128;
  const int row = (idx / 128) % 128;
  const int t = (idx / (128 * 128)) % 2;
  const long long item = idx / (2 * 128 * 128);
  const int qp = item % nqp;
  const int h = (item / nqp) % H;
  const int seq = item / (nqp * H);
  const int C = H * D;
  const __half* q = qkv + ((long long)seq * S + qp * 256 + t * 128 + row) * 3 * C + h * D;
  const __half* k = qkv + ((long long)seq * S + key) * 3 * C + C + h * D;
  float s = 0.f;
  for (int d = 0; d < D; ++d) s += __half2float(q[d]) * __half2float(k[d]);
  out[idx] = s;
}

static float frand(uint32_t& s) {   // uniform (-1, 1)
  s = s * 1664525u + 1013904223u;
  return ((s >> 8) * (1.0f / 8388608.0f)) - 1.0f;
}

int main(int argc, char** argv) {
  bool do_time = false;
  const char* modes = "3,2,0";
  unsigned lbo = 16;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--time")) do_time = true;
    if (!strcmp(argv[i], "--modes") && i + 1 < argc) modes = argv[++i];
    if (!strcmp(argv[i], "--lbo") && i + 1 < argc) lbo = atoi(argv[++i]);
  }
  const int H = 16, S = 1024, C = H * D;
  const float scale = 1.0f / sqrtf((float)D);
  const int n_seq = 2;
  const size_t rows = (size_t)n_seq * S;
  std::vector<__half> h_qkv(rows * 3 * C);
  uint32_t seed = 12345;
  for (size_t r = 0; r < rows; ++r)
    for (int c = 0; c < 3 * C; ++c) {
      float v = 1.7f * frand(seed);   // unit-ish variance
      // key rows late in the sequence are larger: the running maximum rises across key tiles (exercises the rescale)
      if (c >= C && c < 2 * C && (r % S) >= 600 && (r % 7) == 0) v *= 3.0f;
      h_qkv[r * 3 * C + c] = __float2half(v);
    }
  __half *d_qkv, *d_out;
  float *d_ref, *d_dbg, *d_sref;
  CK(cudaMalloc(&d_qkv, h_qkv.size() * 2));
  CK(cudaMalloc(&d_out, rows * C * 2));
  CK(cudaMalloc(&d_ref, rows * C * 4));
  const size_t n_scores = (size_t)n_seq * H * (S / 256) * 2 * 128 * 128;
  CK(cudaMalloc(&d_dbg, n_scores * 4));
  CK(cudaMalloc(&d_sref, n_scores * 4));
  CK(cudaMemcpy(d_qkv, h_qkv.data(), h_qkv.size() * 2, cudaMemcpyHostToDevice));
  std::vector<__half> h_out(rows * C);
  std::vector<float> h_ref(rows * C);
  int failures = 0;
  for (const char* p = modes; *p; ++p) {
    if (*p == ',') continue;
    const int mode = *p - '0';
    CK(cudaMemset(d_out, 0xFF, rows * C * 2));
    CK(cudaMemset(d_dbg, 0, n_scores * 4));
    int rc = mode == 0 ? vq_attn_spatial(d_qkv, d_out, n_seq, S, H, D, scale, nullptr)
                       : vq_attn_spatial_debug(d_qkv, d_out, n_seq, S, H, D, scale, mode, d_dbg, lbo, nullptr);
    if (rc != 0) {
      printf("mode %d: launch rc=%d\n", mode, rc);
      return 3;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("mode %d: kernel error %s\n", mode, cudaGetErrorString(e));
      return 4;
    }
    if (mode == 3) {
      const long long total = (long long)n_scores;
      ref_scores<<<(unsigned)((total + 255) / 256), 256>>>(d_qkv, d_sref, n_seq, S, H);
      CK(cudaDeviceSynchronize());
      std::vector<float> a(n_scores), b(n_scores);
      CK(cudaMemcpy(a.data(), d_dbg, n_scores * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(b.data(), d_sref, n_scores * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      long long bad = 0, first = -1;
      for (size_t i = 0; i < n_scores; ++i) {
        double err = fabs((double)a[i] - b[i]);
        if (err > maxerr) maxerr = err;
        if (err > 0.05) {
          ++bad;
          if (first < 0) first = (long long)i;
        }
      }
      printf("mode 3 (scores of key tile 0): max abs err %.4g, bad %lld / %zu", maxerr, bad, n_scores);
      if (first >= 0)
        printf("  first bad idx %lld (item %lld t %lld row %lld key %lld): got %g want %g", first, first / 32768,
               (first / 16384) % 2, (first / 128) % 128, first % 128, a[first], b[first]);
      printf("\n");
      if (bad) ++failures;
    }
    const int rmode = mode == 3 ? 0 : mode;
    const long long total = (long long)rows * H;
    ref_attn<<<(unsigned)((total + 127) / 128), 128>>>(d_qkv, d_ref, n_seq, S, H, scale, rmode);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h_out.data(), d_out, rows * C * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h_ref.data(), d_ref, rows * C * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    long long bad = 0, first = -1, nan = 0;
    for (size_t i = 0; i < rows * C; ++i) {
      const float g = __half2float(h_out[i]);
      if (!(g == g)) {
        ++nan;
        ++bad;
        if (first < 0) first = (long long)i;
        continue;
      }
      const double err = fabs((double)g - h_ref[i]);
      if (err > maxerr) maxerr = err;
      if (fabs(h_ref[i]) > maxref) maxref = fabs(h_ref[i]);
      if (err > 4e-3 + 4e-3 * fabs(h_ref[i])) {
        ++bad;
        if (first < 0) first = (long long)i;
      }
    }
    printf("mode %d output: max abs err %.4g (max |ref| %.3g), bad %lld (nan %lld) / %zu", mode, maxerr, maxref, bad, nan,
           rows * C);
    if (first >= 0)
      printf("  first bad: token %lld head %lld dim %lld got %g want %g", first / C, (first % C) / D, first % D,
             __half2float(h_out[first]), h_ref[first]);
    printf("\n");
    if (bad) ++failures;
  }
  if (do_time) {
    for (int ns : {16, 32}) {
      const size_t r2 = (size_t)ns * S;
      __half *q2, *o2;
      CK(cudaMalloc(&q2, r2 * 3 * C * 2));
      CK(cudaMalloc(&o2, r2 * C * 2));
      for (size_t off = 0; off < r2; off += rows)
        CK(cudaMemcpy(q2 + off * 3 * C, d_qkv, rows * 3 * C * 2, cudaMemcpyDeviceToDevice));
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      for (int i = 0; i < 3; ++i) vq_attn_spatial(q2, o2, ns, S, H, D, scale, nullptr);
      CK(cudaDeviceSynchronize());
      const int reps = 20;
      CK(cudaEventRecord(e0));
      for (int i = 0; i < reps; ++i) vq_attn_spatial(q2, o2, ns, S, H, D, scale, nullptr);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double us = ms * 1000.0 / reps;
      const double flop = 4.0 * ns * H * (double)S * S * D;
      printf("time n_seq=%d: %.1f us  (%.1f TFLOP/s at d=72)\n", ns, us, flop / us * 1e-6);
      cudaFree(q2);
      cudaFree(o2);
    }
  }
  printf(failures ? "SELFTEST FAILED (%d)\n" : "SELFTEST PASSED\n", failures);
  return failures ? 1 : 0;
}
