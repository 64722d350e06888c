// On-device self-test + timing of vq_gemm_w8a8 (links libviditq_b200.so). The check is a naive one-thread-per-output
// integer GEMM with the same epilogue arithmetic, so outputs must be bit-identical. Not part of the product.
//   usage: gemm_selftest [--time]
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../include/viditq_b200.h"

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e_ = (x);                                                        \
    if (e_ != cudaSuccess) {                                                     \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                   \
    }                                                                            \
  } while (0)

__device__ float ref_gelu(float x) {
  // same operation order as the kernel's packed version (gelu_tanh_pair)
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  float x2 = __fmul_rn(x, x);
  float a = __fmaf_rn(x2, k0k1, k0);
  float u = __fmul_rn(x, a);
  float w = __fmul_rn(u, -2.0f * 1.4426950408889634f);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(w));
  float d = __fadd_rn(e, 1.0f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return __fmul_rn(x, r);
}

__global__ void ref_gemm(const uint8_t* a, const __half* ad, const __half* az, const int32_t* ars, int period,
                         const uint8_t* w, const VqColParam* col, int M, int N, int K, int epi, const __half* res,
                         const __half* gate, int rpg, __half* out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  int acc = 0;
  for (int k = 0; k < K; ++k) acc += (int)a[(size_t)m * K + k] * (int)w[(size_t)n * K + k];
  int srow = m % period;
  float dx = __half2float(ad[srow]);
  int zx = __float2int_rn(__half2float(az[srow]));
  int rs = ars[m];
  VqColParam c = col[n];
  int t = acc - zx * c.c1 - rs * c.zw;
  float f = __fmaf_rn((float)t, __fmul_rn(dx, c.dw), c.bias);
  __half y = __float2half_rn(f);
  if (epi == VQ_EPI_GELU_TANH) y = __float2half_rn(ref_gelu(__half2float(y)));
  if (epi == VQ_EPI_GATE_RESIDUAL) {
    __half g = gate[(size_t)(m / rpg) * N + n];
    y = __hadd_rn(res[(size_t)m * N + n], __hmul_rn(g, y));
  }
  out[(size_t)m * N + n] = y;
}

static uint32_t rng_state = 12345u;
static uint32_t rnd() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 8;
}

static int run_case(int M, int N, int K, int epi, int period, bool timeit, bool inplace = false) {
  std::vector<uint8_t> ha((size_t)M * K), hw((size_t)N * K);
  for (auto& v : ha) v = rnd() & 255;
  for (auto& v : hw) v = rnd() & 255;
  std::vector<__half> had(period), haz(period), hres((size_t)M * N), hgate;
  std::vector<int32_t> hars(M);
  for (int i = 0; i < period; ++i) {
    had[i] = __float2half_rn(0.01f + (rnd() % 1000) * 1e-5f);
    haz[i] = __float2half_rn((float)(rnd() % 256));
  }
  for (int m = 0; m < M; ++m) {
    int s = 0;
    for (int k = 0; k < K; ++k) s += ha[(size_t)m * K + k];
    hars[m] = s;
  }
  std::vector<VqColParam> hcol(N);
  for (int n = 0; n < N; ++n) {
    int zw = rnd() % 256, s = 0;
    for (int k = 0; k < K; ++k) s += hw[(size_t)n * K + k];
    hcol[n].c1 = s - K * zw;
    hcol[n].zw = zw;
    hcol[n].dw = 1e-4f + (rnd() % 1000) * 1e-7f;
    hcol[n].bias = ((int)(rnd() % 2001) - 1000) * 1e-3f;
  }
  for (auto& v : hres) v = __float2half_rn(((int)(rnd() % 2001) - 1000) * 1e-3f);
  int rpg = M >= 4 ? (M + 1) / 2 : M;
  int ngate = (M + rpg - 1) / rpg;
  hgate.resize((size_t)ngate * N);
  for (auto& v : hgate) v = __float2half_rn(((int)(rnd() % 2001) - 1000) * 1e-3f);

  uint8_t *da, *dw;
  __half *dad, *daz, *dres, *dgate, *dout, *dref;
  int32_t* dars;
  VqColParam* dcol;
  CK(cudaMalloc(&da, ha.size()));
  CK(cudaMalloc(&dw, hw.size()));
  CK(cudaMalloc(&dad, period * 2));
  CK(cudaMalloc(&daz, period * 2));
  CK(cudaMalloc(&dars, M * 4));
  CK(cudaMalloc(&dcol, N * sizeof(VqColParam)));
  CK(cudaMalloc(&dres, (size_t)M * N * 2));
  CK(cudaMalloc(&dgate, hgate.size() * 2));
  CK(cudaMalloc(&dout, (size_t)M * N * 2));
  CK(cudaMalloc(&dref, (size_t)M * N * 2));
  CK(cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), hw.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dad, had.data(), period * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(daz, haz.data(), period * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dars, hars.data(), M * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dcol, hcol.data(), N * sizeof(VqColParam), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dres, hres.data(), (size_t)M * N * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dgate, hgate.data(), hgate.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xFF, (size_t)M * N * 2));

  if (inplace) CK(cudaMemcpy(dout, dres, (size_t)M * N * 2, cudaMemcpyDeviceToDevice));  // out aliases the residual
  int rc = vq_gemm_w8a8(da, dad, daz, dars, period, dw, dcol, M, N, K, epi, inplace ? dout : dres, N, dgate, rpg, dout,
                        N, 0);
  if (rc != VQ_OK) {
    printf("vq_gemm_w8a8 rc=%d\n", rc);
    return 1;
  }
  CK(cudaDeviceSynchronize());
  dim3 blk(128), grd((N + 127) / 128, M);
  ref_gemm<<<grd, blk>>>(da, dad, daz, dars, period, dw, dcol, M, N, K, epi, dres, dgate, rpg, dref);
  CK(cudaDeviceSynchronize());
  std::vector<uint16_t> ho((size_t)M * N), hr((size_t)M * N);
  CK(cudaMemcpy(ho.data(), dout, ho.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hr.data(), dref, hr.size() * 2, cudaMemcpyDeviceToHost));
  size_t bad = 0, first = (size_t)-1;
  for (size_t i = 0; i < ho.size(); ++i)
    if (ho[i] != hr[i]) {
      if (bad == 0) first = i;
      ++bad;
    }
  printf("case M=%d N=%d K=%d epi=%d%s period=%d : mismatches %zu / %zu", M, N, K, epi, inplace ? " (in-place)" : "",
         period, bad, ho.size());
  if (bad) {
    __half a, b;
    memcpy(&a, &ho[first], 2);
    memcpy(&b, &hr[first], 2);
    printf("  first at (m=%zu,n=%zu) got %f want %f", first / N, first % N, __half2float(a), __half2float(b));
  }
  printf("\n");

  if (timeit && !bad) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 5; ++i)
      vq_gemm_w8a8(da, dad, daz, dars, period, dw, dcol, M, N, K, epi, dres, N, dgate, rpg, dout, N, 0);
    CK(cudaDeviceSynchronize());
    const int iters = 50;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i)
      vq_gemm_w8a8(da, dad, daz, dars, period, dw, dcol, M, N, K, epi, dres, N, dgate, rpg, dout, N, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    double us = ms * 1e3 / iters;
    double tops = 2.0 * M * N * (double)K / (us * 1e-6) / 1e12;
    printf("   time %.2f us  -> %.1f TOPS (warm L2, back-to-back)\n", us, tops);
    // internal debug epilogues: 3 = mainloop only, 5 = + TMEM loads, 6 = + dequant math, 7 = loads + staging + TMA stores
    const int dbg[4] = {3, 5, 6, 7};
    const char* dname[4] = {"mainloop-only", "+tmem loads", "+loads+math", "loads+stores"};
    for (int di = 0; di < 4; ++di) {
      for (int i = 0; i < 3; ++i)
        vq_gemm_w8a8(da, dad, daz, dars, period, dw, dcol, M, N, K, dbg[di], dres, N, dgate, rpg, dout, N, 0);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      for (int i = 0; i < iters; ++i)
        vq_gemm_w8a8(da, dad, daz, dars, period, dw, dcol, M, N, K, dbg[di], dres, N, dgate, rpg, dout, N, 0);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      us = ms * 1e3 / iters;
      printf("   %-14s %.2f us -> %.1f TOPS\n", dname[di], us, 2.0 * M * N * (double)K / (us * 1e-6) / 1e12);
    }
  }
  cudaFree(da); cudaFree(dw); cudaFree(dad); cudaFree(daz); cudaFree(dars); cudaFree(dcol);
  cudaFree(dres); cudaFree(dgate); cudaFree(dout); cudaFree(dref);
  return bad ? 1 : 0;
}

int main(int argc, char** argv) {
  bool timeit = argc > 1 && !strcmp(argv[1], "--time");
  bool time32 = argc > 1 && !strcmp(argv[1], "--time32");   // the shapes a stacked cfg_split step launches (M = 32768)
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs=%d lib version %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount,
         vq_version());
  int fails = 0;
  if (argc > 5 && !strcmp(argv[1], "--case")) {  // single case, no timing loop: for ncu captures
    int M = atoi(argv[2]), N = atoi(argv[3]), K = atoi(argv[4]), epi = atoi(argv[5]);
    return run_case(M, N, K, epi, M, false);
  }
  if (time32) {
    fails += run_case(32768, 1152, 1152, VQ_EPI_GATE_RESIDUAL, 32768, true, true);
    fails += run_case(32768, 3456, 1152, VQ_EPI_BIAS, 32768, true);
    fails += run_case(32768, 4608, 1152, VQ_EPI_BIAS, 32768, true);
    fails += run_case(32768, 1152, 4608, VQ_EPI_GATE_RESIDUAL, 32768, true, true);
    printf(fails ? "SELFTEST FAILED (%d cases)\n" : "SELFTEST PASSED\n", fails);
    return fails ? 1 : 0;
  }
  fails += run_case(128, 192, 128, VQ_EPI_BIAS, 128, false);    // one tile, one K block
  fails += run_case(128, 192, 1152, VQ_EPI_BIAS, 128, false);   // full K pipeline (9 blocks > 5 stages)
  fails += run_case(256, 384, 1152, VQ_EPI_BIAS, 256, false);   // 4 tiles
  fails += run_case(120, 2304, 1152, VQ_EPI_BIAS, 120, false);  // kv_linear shape: ragged M
  fails += run_case(200, 32, 1152, VQ_EPI_BIAS, 100, false);    // PixArt final_layer: N tail, pooled period
  fails += run_case(2048, 1152, 1152, VQ_EPI_GATE_RESIDUAL, 1024, false);
  fails += run_case(2048, 1152, 1152, VQ_EPI_GATE_RESIDUAL, 1024, false, true);   // TMA reduce-add path
  fails += run_case(300, 200, 1152, VQ_EPI_GATE_RESIDUAL, 300, false, true);      // ragged M / N, in place
  fails += run_case(2048, 4608, 1152, VQ_EPI_GELU_TANH, 2048, false);
  fails += run_case(2048, 1152, 4608, VQ_EPI_BIAS, 2048, false);
  fails += run_case(16384, 1152, 1152, VQ_EPI_BIAS, 16384, timeit);  // > 148 tiles: persistent loop, TMEM double buffer
  if (timeit) {
    fails += run_case(16384, 4608, 1152, VQ_EPI_GELU_TANH, 16384, true);
    fails += run_case(16384, 1152, 4608, VQ_EPI_GATE_RESIDUAL, 16384, true);
    fails += run_case(16384, 1152, 1152, VQ_EPI_GATE_RESIDUAL, 16384, true, true);
    fails += run_case(16384, 3456, 1152, VQ_EPI_BIAS, 16384, true);
  }
  printf(fails ? "SELFTEST FAILED (%d cases)\n" : "SELFTEST PASSED\n", fails);
  return fails ? 1 : 0;
}
