// C-ABI entry points of the tcgen05 GEMM family (see gemm_sm100.cuh).  Each replaces an ATen matmul call
// site of the reference (cited per function in include/tfx_b200.h).
#include "gemm_sm100.cuh"
#include "common.cuh"
#include "../../include/tfx_b200.h"
#include <string.h>

namespace tfx { int num_sms(); }
using namespace tfx;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

static int finish(int rc, const char* what) {
  if (rc == 0) return 0;
  if (rc <= -1000) set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d) - check 16-byte alignment of pointers and row pitches", what, -(rc + 1000));
  else if (rc == -1) set_error("%s: cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)", what);
  else if (rc == -2) set_error("%s: cudaFuncSetAttribute(max dynamic smem) failed: %s", what, cudaGetErrorString(cudaGetLastError()));
  else set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(cudaGetLastError()));
  return rc;
}

template <int EPI, int BN>
static int dispatch_major(const GemmOperand& A, const GemmOperand& B, const GemmParams& p, cudaStream_t st) {
  const int sms = num_sms();
  if (!A.mn_major && !B.mn_major) return launch_gemm_t<BN, false, false, EPI>(A, B, p, sms, st);
  if constexpr (EPI == EPI_STORE) {
    if (!A.mn_major && B.mn_major) return launch_gemm_t<BN, false, true, EPI>(A, B, p, sms, st);
    if (A.mn_major && B.mn_major) return launch_gemm_t<BN, true, true, EPI>(A, B, p, sms, st);
    return launch_gemm_t<BN, true, false, EPI>(A, B, p, sms, st);
  }
  return -3;
}

extern "C" {

int tfx_gemm_set_cluster_mode(int mode) {
  TFX_REQUIRE(mode >= 1 && mode <= 3, "gemm_set_cluster_mode: mode %d not in {1 (never pair), 2 (always pair), 3 (pair long-K launches, default)}", mode);
  gemm_cluster_mode_ref() = mode;
  return 0;
}

int tfx_gemm_store(const void* A, long long lda, int a_mn_major, const void* B, long long ldb, int b_mn_major, int M, int N, int K,
                   float* out_f32, long long ld_f32, void* out_bf16, long long ld_bf16, const float* bias, const long long* row_off,
                   float alpha, int accumulate, int k_splits, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  TFX_REQUIRE(K > 0, "gemm_store: K must be > 0");
  TFX_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "gemm_store: operand row pitches (%lld, %lld) must be multiples of 8 bf16", lda, ldb);
  TFX_REQUIRE(!(k_splits > 1) || (accumulate && out_f32 && !out_bf16 && !bias), "gemm_store: split-K requires fp32 accumulate output only");
  GemmParams p; memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.k_splits = k_splits < 1 ? 1 : k_splits;
  p.out_f32 = out_f32; p.ld_f32 = ld_f32; p.out_bf16 = (__nv_bfloat16*)out_bf16; p.ld_bf16 = ld_bf16; p.bias = bias; p.row_off = row_off;
  p.alpha = alpha; p.accumulate_f32 = accumulate;
  GemmOperand a{A, lda, a_mn_major != 0}, b{B, ldb, b_mn_major != 0};
  // wide N tiles when there is enough N to fill them; the 128-wide tile otherwise
  const bool wide = (N % 256 == 0) || N >= 1024;
  const int rc = wide ? dispatch_major<EPI_STORE, 256>(a, b, p, ST(stream)) : dispatch_major<EPI_STORE, 128>(a, b, p, ST(stream));
  return finish(rc, "gemm_store");
}

int tfx_gemm_qkvg(const void* u, long long ldu, const void* W, long long ldw, int M, int H, int D, void* q, void* k, void* v, float* gates, float* qk_inv,
                  const float* q_gamma, const float* k_gamma, const int* rope_pos, const float* rope_cs_t, int rope_len, const int* kv_rows, float* mix_pre, void* stream) {
  if (M <= 0) return 0;
  TFX_REQUIRE(!mix_pre || H <= 16, "gemm_qkvg: the value-residual mix columns share the 32-column gate slab: heads must be <= 16 (got %d)", H);
  TFX_REQUIRE(H >= 2 && H % 2 == 0 && H <= 32, "gemm_qkvg: heads must be even and in [2, 32] (got %d)", H);
  TFX_REQUIRE(ldu % 8 == 0 && ldw % 8 == 0, "gemm_qkvg: row pitches must be multiples of 8");
  GemmParams p; memset(&p, 0, sizeof(p));
  p.M = M; p.N = 3 * H * 64 + 128; p.K = D; p.k_splits = 1; p.H = H;
  p.q = (__nv_bfloat16*)q; p.k = (__nv_bfloat16*)k; p.v = (__nv_bfloat16*)v; p.gates = gates; p.qk_inv = qk_inv;
  p.q_gamma = q_gamma; p.k_gamma = k_gamma; p.rope_pos = rope_pos; p.rope_cs = (const float2*)rope_cs_t; p.rope_len = rope_len; p.kv_rows = kv_rows; p.mix_pre = mix_pre;
  GemmOperand a{u, ldu, false}, b{W, ldw, false};
  // 256-wide tiles (4 heads per tile, 16 epilogue warps) whenever the head count allows it; 2 heads per 128-wide tile otherwise
  if (H % 4 == 0) return finish(launch_gemm_t<256, false, false, EPI_QKVG>(a, b, p, num_sms(), ST(stream)), "gemm_qkvg");
  return finish(launch_gemm_t<128, false, false, EPI_QKVG>(a, b, p, num_sms(), ST(stream)), "gemm_qkvg");
}

int tfx_gemm_resid(const void* A, long long lda, const void* A2, long long lda2, int K1, const void* W, long long ldw, int M, int N, int K, const float* bias,
                   const float* x_res, float* x_out, void* x_out_bf16, void* y_bf16, const int* cond_row, const float* zgate, long long zgate_ld,
                   const float* layerscale, void* stream) {
  if (M <= 0) return 0;
  TFX_REQUIRE(N % 32 == 0, "gemm_resid: N (%d) must be a multiple of 32", N);
  TFX_REQUIRE(!A2 || K1 % 64 == 0, "gemm_resid: K1 (%d) must be a multiple of 64", K1);
  TFX_REQUIRE(x_out || x_out_bf16, "gemm_resid: at least one of x_out (fp32) / x_out_bf16 is required");
  GemmParams p; memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.k_splits = 1; p.K1 = A2 ? K1 : K; p.bias = bias;
  p.x_res = x_res; p.x_out = x_out; p.x_out_bf16 = (__nv_bfloat16*)x_out_bf16; p.y_bf16 = (__nv_bfloat16*)y_bf16;
  p.cond_row = cond_row; p.zgate = zgate; p.zgate_ld = zgate_ld; p.ls = layerscale;
  GemmOperand a{A, lda, false, A2, lda2}, b{W, ldw, false};
  return finish(launch_gemm_t<128, false, false, EPI_RESID>(a, b, p, num_sms(), ST(stream)), "gemm_resid");
}

int tfx_gemm_geglu(const void* u, long long ldu, const void* W1p, long long ldw, const float* b1p, int M, int Np, int K, void* vg, void* h, void* stream) {
  if (M <= 0) return 0;
  TFX_REQUIRE(Np % 128 == 0, "gemm_geglu: packed N (%d) must be a multiple of 128", Np);
  GemmParams p; memset(&p, 0, sizeof(p));
  p.M = M; p.N = Np; p.K = K; p.k_splits = 1; p.bias = b1p; p.vg = (__nv_bfloat16*)vg; p.h = (__nv_bfloat16*)h;
  GemmOperand a{u, ldu, false}, b{W1p, ldw, false};
  return finish(launch_gemm_t<256, false, false, EPI_GEGLU>(a, b, p, num_sms(), ST(stream)), "gemm_geglu");
}

}  // extern "C"
