// Standalone bring-up / micro-benchmark harness for gemm_sm100.cuh (not part of the shipped library).
// Checks every operand-major combination against a naive CUDA-core GEMM and times the hot shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 harness_gemm.cu -o harness_gemm
#include "gemm_sm100.cuh"
#include <vector>
#include <cstdlib>
#include <cmath>
#include <cstring>

using namespace tfx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void naive_gemm(const __nv_bfloat16* A, long long lda, bool a_mn, const __nv_bfloat16* B, long long ldb, bool b_mn,
                           float* C, int M, int N, int K) {
  int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    float a = __bfloat162float(a_mn ? A[(long long)k * lda + m] : A[(long long)m * lda + k]);
    float b = __bfloat162float(b_mn ? B[(long long)k * ldb + n] : B[(long long)n * ldb + k]);
    acc += a * b;
  }
  C[(long long)m * N + n] = acc;
}

static __nv_bfloat16* rand_bf16(size_t n, unsigned seed, float scale = 1.f) {
  std::vector<__nv_bfloat16> h(n);
  srand(seed);
  for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16(scale * ((rand() % 2001) - 1000) / 1000.f);
  __nv_bfloat16* d; CK(cudaMalloc(&d, n * 2)); CK(cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice));
  return d;
}

template <int BN, bool A_MN, bool B_MN>
static int check(const char* name, int M, int N, int K, int k_splits, int sms) {
  long long lda = A_MN ? ((M + 7) / 8 * 8) : ((K + 7) / 8 * 8);
  long long ldb = B_MN ? ((N + 7) / 8 * 8) : ((K + 7) / 8 * 8);
  size_t na = (size_t)lda * (A_MN ? K : M), nb = (size_t)ldb * (B_MN ? K : N);
  __nv_bfloat16 *A = rand_bf16(na, 1), *B = rand_bf16(nb, 2);
  float *C, *Cref; CK(cudaMalloc(&C, (size_t)M * N * 4)); CK(cudaMalloc(&Cref, (size_t)M * N * 4));
  CK(cudaMemset(C, 0, (size_t)M * N * 4));
  naive_gemm<<<dim3((N + 127) / 128, M), 128>>>(A, lda, A_MN, B, ldb, B_MN, Cref, M, N, K);
  CK(cudaDeviceSynchronize());
  GemmParams p; memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.k_splits = k_splits; p.out_f32 = C; p.ld_f32 = N; p.alpha = 1.f; p.accumulate_f32 = k_splits > 1;
  GemmOperand oa{A, lda, A_MN}, ob{B, ldb, B_MN};
  int rc = launch_gemm_t<BN, A_MN, B_MN, EPI_STORE>(oa, ob, p, sms, 0);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc || e != cudaSuccess) { printf("[%s] LAUNCH FAILED rc=%d err=%s\n", name, rc, cudaGetErrorString(e)); return 1; }
  std::vector<float> h((size_t)M * N), r((size_t)M * N);
  CK(cudaMemcpy(h.data(), C, h.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r.data(), Cref, r.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0; size_t bad = 0, first_bad = (size_t)-1;
  for (size_t i = 0; i < h.size(); ++i) {
    double d = fabs((double)h[i] - r[i]);
    if (d > maxerr) maxerr = d;
    if (fabs(r[i]) > maxref) maxref = fabs(r[i]);
    if (d > 1e-2 + 1e-3 * fabs(r[i])) { if (!bad) first_bad = i; ++bad; }
  }
  printf("[%s] M=%d N=%d K=%d splits=%d  maxerr=%.4g (maxref %.3g)  bad=%zu/%zu %s\n", name, M, N, K, k_splits, maxerr, maxref, bad, h.size(), bad ? "FAIL" : "ok");
  if (bad) {
    size_t i = first_bad;
    printf("   first bad at (%zu,%zu): got %.5f want %.5f; C[0,0..3] = %.4f %.4f %.4f %.4f want %.4f %.4f %.4f %.4f\n", i / N, i % N, h[i], r[i],
           h[0], h[1], h[2], h[3], r[0], r[1], r[2], r[3]);
  }
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(Cref);
  return bad ? 1 : 0;
}

template <int BN, bool A_MN, bool B_MN>
static void bench(const char* name, int M, int N, int K, int k_splits, int sms) {
  long long lda = A_MN ? M : K, ldb = B_MN ? N : K;
  size_t na = (size_t)lda * (A_MN ? K : M), nb = (size_t)ldb * (B_MN ? K : N);
  __nv_bfloat16 *A, *B; CK(cudaMalloc(&A, na * 2)); CK(cudaMalloc(&B, nb * 2));
  CK(cudaMemset(A, 0x11, na * 2)); CK(cudaMemset(B, 0x11, nb * 2));
  float* C = nullptr; __nv_bfloat16* Cb = nullptr;
  GemmParams p; memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.k_splits = k_splits; p.alpha = 1.f;
  if (k_splits > 1) { CK(cudaMalloc(&C, (size_t)M * N * 4)); CK(cudaMemset(C, 0, (size_t)M * N * 4)); p.out_f32 = C; p.ld_f32 = N; p.accumulate_f32 = 1; }
  else { CK(cudaMalloc(&Cb, (size_t)M * N * 2)); p.out_bf16 = Cb; p.ld_bf16 = N; }
  GemmOperand oa{A, lda, A_MN}, ob{B, ldb, B_MN};
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch_gemm_t<BN, A_MN, B_MN, EPI_STORE>(oa, ob, p, sms, 0);
  CK(cudaDeviceSynchronize());
  const int iters = 10;
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) launch_gemm_t<BN, A_MN, B_MN, EPI_STORE>(oa, ob, p, sms, 0);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
  double tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12;
  printf("[bench %s] BN=%d M=%d N=%d K=%d splits=%d : %.3f ms  %.1f TFLOP/s\n", name, BN, M, N, K, k_splits, ms, tf);
  cudaFree(A); cudaFree(B); if (C) cudaFree(C); if (Cb) cudaFree(Cb);
}

int main(int argc, char** argv) {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s  sm_%d%d  SMs %d\n", prop.name, prop.major, prop.minor, sms);
  if (argc > 1 && !strcmp(argv[1], "--one")) {
    bench<128, false, false>("qkvg fwd", 65536, 1664, 512, 1, sms);
    return 0;
  }
  int fails = 0;
  // single tile, single k-block first: isolates descriptor errors
  fails += check<128, false, false>("NT 1tile 1kb", 128, 128, 64, 1, sms);
  fails += check<128, false, false>("NT 1tile 8kb", 128, 128, 512, 1, sms);
  fails += check<128, false, false>("NT ragged", 300, 384, 520, 1, sms);
  fails += check<128, false, false>("NT many", 4000, 1664, 512, 1, sms);
  fails += check<256, false, false>("NT BN256", 1000, 768, 512, 1, sms);
  fails += check<128, false, true>("NN 1tile", 128, 128, 64, 1, sms);
  fails += check<128, false, true>("NN dgrad", 1000, 512, 1664, 1, sms);
  fails += check<256, false, true>("NN dgrad BN256", 1000, 512, 2816, 1, sms);
  fails += check<128, true, true>("TN 1tile", 128, 128, 64, 1, sms);
  fails += check<128, true, true>("TN wgrad", 1664, 512, 4096, 1, sms);
  fails += check<128, true, true>("TN wgrad splitK", 1664, 512, 8200, 8, sms);
  fails += check<128, true, false>("TT", 256, 256, 256, 1, sms);
  printf("correctness: %d failing cases\n", fails);
  if (argc > 1 && !strcmp(argv[1], "--bench")) {
    bench<128, false, false>("qkvg fwd", 65536, 1664, 512, 1, sms);
    bench<256, false, false>("qkvg fwd", 65536, 1792, 512, 1, sms);
    bench<128, false, false>("ffn_in fwd", 65536, 2816, 512, 1, sms);
    bench<256, false, false>("ffn_in fwd", 65536, 2816, 512, 1, sms);
    bench<128, false, false>("ffn_out fwd", 65536, 512, 1408, 1, sms);
    bench<256, false, false>("ffn_out fwd", 65536, 512, 1408, 1, sms);
    bench<128, false, true>("ffn_in dgrad", 65536, 512, 2816, 1, sms);
    bench<256, false, true>("ffn_in dgrad", 65536, 512, 2816, 1, sms);
    bench<128, true, true>("ffn_in wgrad", 2816, 512, 65536, 8, sms);
    bench<256, true, true>("ffn_in wgrad", 2816, 512, 65536, 16, sms);
    bench<128, false, false>("big square", 8192, 8192, 8192, 1, sms);
    bench<256, false, false>("big square", 8192, 8192, 8192, 1, sms);
  }
  return fails ? 1 : 0;
}
