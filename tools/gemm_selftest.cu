// gemm_selftest.cu — torch-free check of zb_gemm (all operand layouts, tails, epilogues, split-K) against a
// CPU fp32 reference over the same bf16-rounded inputs, plus a timing sweep.  Test infrastructure only.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 tools/gemm_selftest.cu \
//          -o tools/gemm_selftest -L zero_b200 -lzero_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../zero_b200'
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/zero_b200.h"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}
static float bf16r(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct Case {
  int m, n, k, a_mn, b_mn, flags, d_f32, split;
};

static int run_case(const Case& c, bool verbose) {
  const int M = c.m, N = c.n, K = c.k;
  // logical A(m,k), B(n,k)
  std::vector<float> A((size_t)M * K), B((size_t)N * K), bias(N), maskv((size_t)M * N), D0((size_t)M * N);
  for (auto& x : A) x = bf16r(frand());
  for (auto& x : B) x = bf16r(frand());
  for (auto& x : bias) x = frand();
  for (auto& x : maskv) x = bf16r(frand());
  for (auto& x : D0) x = (c.flags & ZB_EPI_ACCUM) ? frand() : 0.f;
  const int lda = c.a_mn ? ((M + 7) / 8 * 8) : ((K + 7) / 8 * 8);
  const int ldb = c.b_mn ? ((N + 7) / 8 * 8) : ((K + 7) / 8 * 8);
  const int ldd = (N + 7) / 8 * 8;
  std::vector<__nv_bfloat16> hA((size_t)(c.a_mn ? K : M) * lda), hB((size_t)(c.b_mn ? K : N) * ldb),
      hM((size_t)M * ldd);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      size_t idx = c.a_mn ? (size_t)k * lda + m : (size_t)m * lda + k;
      hA[idx] = __float2bfloat16(A[(size_t)m * K + k]);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      size_t idx = c.b_mn ? (size_t)k * ldb + n : (size_t)n * ldb + k;
      hB[idx] = __float2bfloat16(B[(size_t)n * K + k]);
    }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) hM[(size_t)m * ldd + n] = __float2bfloat16(maskv[(size_t)m * N + n]);

  void *dA, *dB, *dD, *dMask;
  float* dBias;
  const size_t dbytes = (size_t)M * ldd * (c.d_f32 ? 4 : 2);
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, dbytes));
  CK(cudaMalloc(&dMask, hM.size() * 2));
  CK(cudaMalloc(&dBias, N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dMask, hM.data(), hM.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBias, bias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, dbytes));
  if (c.flags & ZB_EPI_ACCUM) {
    std::vector<float> h((size_t)M * ldd, 0.f);
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) h[(size_t)m * ldd + n] = D0[(size_t)m * N + n];
    CK(cudaMemcpy(dD, h.data(), dbytes, cudaMemcpyHostToDevice));
  }

  zb_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.a = dA; g.b = dB; g.d = dD; g.m = M; g.n = N; g.k = K; g.lda = lda; g.ldb = ldb; g.ldd = ldd;
  g.a_layout = c.a_mn; g.b_layout = c.b_mn; g.d_dtype = c.d_f32 ? ZB_F32 : ZB_BF16; g.flags = c.flags;
  g.bias = dBias; g.mask = dMask; g.ldmask = ldd; g.alpha = 0.5f; g.split_k = c.split;
  int rc = zb_gemm(&g, 0);
  if (rc != 0) {
    printf("FAIL rc=%d (%s) m=%d n=%d k=%d\n", rc, zb_last_error_string(), M, N, K);
    return 1;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("FAIL kernel error %s m=%d n=%d k=%d a_mn=%d b_mn=%d\n", cudaGetErrorString(e), M, N, K, c.a_mn, c.b_mn);
    exit(3);
  }
  std::vector<float> out((size_t)M * N);
  if (c.d_f32) {
    std::vector<float> h((size_t)M * ldd);
    CK(cudaMemcpy(h.data(), dD, dbytes, cudaMemcpyDeviceToHost));
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) out[(size_t)m * N + n] = h[(size_t)m * ldd + n];
  } else {
    std::vector<__nv_bfloat16> h((size_t)M * ldd);
    CK(cudaMemcpy(h.data(), dD, dbytes, cudaMemcpyDeviceToHost));
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) out[(size_t)m * N + n] = __bfloat162float(h[(size_t)m * ldd + n]);
  }
  double max_err = 0, max_ref = 0;
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      const float* ar = &A[(size_t)m * K];
      const float* br = &B[(size_t)n * K];
      for (int k = 0; k < K; ++k) acc += (double)ar[k] * br[k];
      double v = acc * 0.5;
      if (c.flags & ZB_EPI_BIAS) v += bias[n];
      if (c.flags & ZB_EPI_RELU) v = v > 0 ? v : 0;
      if (c.flags & ZB_EPI_RELU_MASK) v = maskv[(size_t)m * N + n] > 0 ? v : 0;
      if (c.flags & ZB_EPI_ACCUM) v += D0[(size_t)m * N + n];
      double err = fabs(v - out[(size_t)m * N + n]);
      double tol = c.d_f32 ? 1e-3 + 1e-4 * fabs(v) : 2e-2 + 1e-2 * fabs(v);
      if (err > tol) {
        if (bad < 5 && verbose) printf("   mismatch at (%d,%d): got %f want %f\n", m, n, out[(size_t)m * N + n], v);
        ++bad;
      }
      if (err > max_err) max_err = err;
      if (fabs(v) > max_ref) max_ref = fabs(v);
    }
  printf("%s m=%5d n=%5d k=%5d A=%s B=%s flags=%2d f32=%d split=%d  max_err=%.3e (max|ref|=%.2f) bad=%d\n",
         bad ? "FAIL" : "ok  ", M, N, K, c.a_mn ? "MN" : "K ", c.b_mn ? "MN" : "K ", c.flags, c.d_f32, c.split, max_err,
         max_ref, bad);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dMask); cudaFree(dBias);
  return bad ? 1 : 0;
}

static void time_case(int M, int N, int K, int a_mn, int b_mn, int flags, int d_f32) {
  void *dA, *dB, *dD;
  const size_t an = (size_t)M * K, bn = (size_t)N * K;
  CK(cudaMalloc(&dA, an * 2));
  CK(cudaMalloc(&dB, bn * 2));
  CK(cudaMalloc(&dD, (size_t)M * N * (d_f32 ? 4 : 2)));
  std::vector<__nv_bfloat16> h(an > bn ? an : bn);
  for (auto& x : h) x = __float2bfloat16(frand());
  CK(cudaMemcpy(dA, h.data(), an * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, h.data(), bn * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, (size_t)M * N * (d_f32 ? 4 : 2)));
  zb_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.a = dA; g.b = dB; g.d = dD; g.m = M; g.n = N; g.k = K;
  g.lda = a_mn ? M : K; g.ldb = b_mn ? N : K; g.ldd = N;
  g.a_layout = a_mn; g.b_layout = b_mn; g.d_dtype = d_f32 ? ZB_F32 : ZB_BF16; g.flags = flags; g.alpha = 1.f;
  float* dBias;
  CK(cudaMalloc(&dBias, (size_t)N * 4));
  CK(cudaMemset(dBias, 0, (size_t)N * 4));
  g.bias = dBias;
  if (zb_gemm(&g, 0) != 0) {
    printf("time_case: %s\n", zb_last_error_string());
    return;
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 5; ++i) zb_gemm(&g, 0);
  CK(cudaDeviceSynchronize());
  const int iters = 20;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) zb_gemm(&g, 0);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= iters;
  printf("time m=%6d n=%6d k=%6d A=%s B=%s flags=%d f32=%d: %.3f ms  %.1f TFLOP/s\n", M, N, K, a_mn ? "MN" : "K ",
         b_mn ? "MN" : "K ", flags, d_f32, ms, 2.0 * M * N * K / ms / 1e9);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

static int only_a = -1, only_b = -1;
static int run_case_f(const Case& c, bool verbose) {
  if (only_a >= 0 && (c.a_mn != only_a || c.b_mn != only_b)) return 0;
  return run_case(c, verbose);
}
#define run_case run_case_f

int main(int argc, char** argv) {
  int fails = 0;
  printf("abi %d\n", zb_abi_version());
  if (argc > 8 && !strcmp(argv[1], "--one")) {  // --one m n k a_mn b_mn flags f32 : time (or profile) one shape
    time_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]));
    return 0;
  }
  if (argc > 3 && !strcmp(argv[1], "--layout")) {
    only_a = atoi(argv[2]);
    only_b = atoi(argv[3]);
  }
  // stage 1: the smallest single-tile problems, one per layout, so a descriptor bug is isolated early
  for (int a_mn = 0; a_mn < 2; ++a_mn)
    for (int b_mn = 0; b_mn < 2; ++b_mn) fails += run_case({128, 64, 64, a_mn, b_mn, 0, 1, 1}, true);
  for (int a_mn = 0; a_mn < 2; ++a_mn)
    for (int b_mn = 0; b_mn < 2; ++b_mn) fails += run_case({128, 256, 128, a_mn, b_mn, 0, 1, 1}, true);
  // stage 2: multi-tile, tails, tile widths
  const Case cases[] = {
      {256, 128, 256, 0, 1, 0, 0, 1},
      {384, 512, 512, 0, 1, ZB_EPI_BIAS, 0, 1},
      {384, 512, 512, 0, 0, ZB_EPI_BIAS | ZB_EPI_RELU, 0, 1},
      {200, 136, 72, 0, 1, ZB_EPI_BIAS, 0, 1},
      {200, 136, 72, 0, 0, 0, 1, 1},
      {200, 136, 72, 1, 1, 0, 1, 1},
      {1000, 1000, 520, 0, 0, 0, 0, 1},
      {4096, 1536, 512, 0, 1, ZB_EPI_BIAS, 0, 1},
      {4096, 2048, 512, 0, 1, ZB_EPI_BIAS | ZB_EPI_RELU, 0, 1},
      {4096, 512, 2048, 0, 0, ZB_EPI_RELU_MASK, 0, 1},
      {512, 1536, 4096, 1, 1, ZB_EPI_ACCUM, 1, 0},
      {512, 2048, 4096, 1, 1, ZB_EPI_ACCUM, 1, 4},
      {1000, 128, 1024, 1, 1, ZB_EPI_ACCUM, 1, 0},
      {1024, 4000, 128, 0, 0, 0, 1, 1},
      {1024, 128, 4000, 0, 1, 0, 0, 1},
      {4000, 128, 1024, 1, 1, ZB_EPI_ACCUM, 1, 0},
  };
  for (const auto& c : cases) fails += run_case(c, true);
  printf("%s: %d failing cases\n", fails ? "SELFTEST FAILED" : "SELFTEST PASSED", fails);
  if (argc > 1 && !strcmp(argv[1], "--time")) {
    time_case(4096, 1536, 512, 0, 1, ZB_EPI_BIAS, 0);
    time_case(4096, 512, 512, 0, 1, 0, 0);
    time_case(4096, 2048, 512, 0, 1, 0, 0);
    time_case(4096, 512, 2048, 0, 1, 0, 0);
    time_case(4096, 32000, 512, 0, 0, 0, 1);
    time_case(4096, 512, 32000, 0, 1, 0, 0);
    time_case(32000, 512, 4096, 1, 1, ZB_EPI_ACCUM, 1);
    time_case(512, 2048, 4096, 1, 1, ZB_EPI_ACCUM, 1);
    time_case(262144, 1536, 512, 0, 1, 0, 0);
    time_case(262144, 2048, 512, 0, 1, 0, 0);
    time_case(262144, 512, 2048, 0, 1, 0, 0);
    time_case(512, 2048, 262144, 1, 1, ZB_EPI_ACCUM, 1);
    time_case(8192, 8192, 8192, 0, 0, 0, 0);
    time_case(8192, 8192, 8192, 0, 1, 0, 0);
  }
  return fails ? 1 : 0;
}
