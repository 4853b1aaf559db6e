// extern "C" operator-level entry points declared in include/vla_b200.h (parity-test surface).
#include "../../include/vla_b200.h"
#include "gemm.cuh"
#include "ops.cuh"

#include <string>

namespace {
thread_local std::string g_err;
int fail(int code, const char* msg) {
  g_err = msg ? msg : "unknown error";
  return code;
}
}  // namespace

extern "C" {

const char* vla_global_error(void) { return g_err.c_str(); }

long long vla_total_launch_count(void) { return vla::gemm_launch_count() + vla::ops_launch_count(); }

int vla_profile_gemm(int enable) {
  vla::gemm_profile_enable(enable != 0);
  return 0;
}

int vla_profile_gemm_read(double* total_ms, long long* launches) {
  return vla::gemm_profile_read(total_ms, launches);
}

int vla_op_gemm(const void* A, long long a_batch_stride, int lda, int rows, int batches, const void* W,
                int ldw, int N, int K, void* C, long long c_batch_stride, int ldc, const float* bias,
                const float* colscale, const void* resid, long long r_batch_stride, int ldr, int act,
                int force_bn, void* stream) {
  vla::GemmArgs g;
  g.A = static_cast<const __nv_bfloat16*>(A);
  g.a_batch_stride = a_batch_stride;
  g.lda = lda;
  g.rows = rows;
  g.batches = batches;
  g.W = static_cast<const __nv_bfloat16*>(W);
  g.ldw = ldw;
  g.N = N;
  g.K = K;
  g.C = static_cast<__nv_bfloat16*>(C);
  g.c_batch_stride = c_batch_stride;
  g.ldc = ldc;
  g.bias = bias;
  g.colscale = colscale;
  g.resid = static_cast<const __nv_bfloat16*>(resid);
  g.r_batch_stride = r_batch_stride;
  g.ldr = ldr;
  g.act = act;
  g.force_bn = force_bn;
  const char* err = nullptr;
  int rc = vla::gemm_launch(g, static_cast<cudaStream_t>(stream), &err);
  if (rc) return fail(rc, err);
  return 0;
}

int vla_op_layernorm(const void* x, int rows, int dim, int ldx, const float* w, const float* b, float eps,
                     void* y, int ldy, void* stream) {
  const char* err = nullptr;
  int rc = vla::layernorm_launch(static_cast<const __nv_bfloat16*>(x), rows, dim, ldx, w, b, eps,
                                 static_cast<__nv_bfloat16*>(y), ldy, static_cast<cudaStream_t>(stream), &err);
  return rc ? fail(rc, err) : 0;
}

int vla_op_fold_norm(void* W, int N, int K, int ldw, const float* norm_w, const float* norm_b, float* bias,
                     float* colsum, void* stream) {
  const char* err = nullptr;
  int rc = vla::fold_norm_launch(static_cast<__nv_bfloat16*>(W), N, K, ldw, norm_w, norm_b, bias, colsum,
                                 static_cast<cudaStream_t>(stream), &err);
  return rc ? fail(rc, err) : 0;
}

int vla_op_norm_gemm(const void* x, int rows, int ldx, const void* W, int ldw, int N, int K, void* C, int ldc,
                     const float* bias, const float* colsum, int rms, float eps, int act, float* stats,
                     void* stream) {
  const char* err = nullptr;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = vla::row_stats_launch(static_cast<const __nv_bfloat16*>(x), rows, K, ldx, rms, eps, stats, s, &err);
  if (rc) return fail(rc, err);
  vla::GemmArgs g;
  g.A = static_cast<const __nv_bfloat16*>(x);
  g.lda = ldx;
  g.rows = rows;
  g.W = static_cast<const __nv_bfloat16*>(W);
  g.ldw = ldw;
  g.N = N;
  g.K = K;
  g.C = static_cast<__nv_bfloat16*>(C);
  g.ldc = ldc;
  g.bias = bias;
  g.act = act;
  g.row_stats = stats;
  g.colsum = rms ? nullptr : colsum;
  rc = vla::gemm_launch(g, s, &err);
  return rc ? fail(rc, err) : 0;
}

int vla_op_block_tail(const void* a, int lda, int rows, const void* W1, int ldw1, int K1, void* x, int D,
                      const float* bias1, const float* colscale1, const void* W2, int ldw2, int N2, void* out, int ldo,
                      const float* bias2, const float* colsum2, int rms, float eps, int act, float* partials,
                      void* stream) {
  const char* err = nullptr;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  vla::GemmArgs g;  // x += colscale1 * (a @ W1^T + bias1), leaving the row statistics of the new x
  g.A = static_cast<const __nv_bfloat16*>(a); g.lda = lda; g.rows = rows;
  g.W = static_cast<const __nv_bfloat16*>(W1); g.ldw = ldw1; g.N = D; g.K = K1;
  g.C = static_cast<__nv_bfloat16*>(x); g.ldc = D; g.bias = bias1; g.colscale = colscale1;
  g.resid = static_cast<const __nv_bfloat16*>(x); g.ldr = D; g.stat_out = partials; g.resid_staged = 1;
  int rc = vla::gemm_launch(g, s, &err);
  if (rc) return fail(rc, err);
  g = vla::GemmArgs();  // out = act(Linear(Norm(x))) with the norm folded into W2 and the statistics from the partials
  g.A = static_cast<const __nv_bfloat16*>(x); g.lda = D; g.rows = rows;
  g.W = static_cast<const __nv_bfloat16*>(W2); g.ldw = ldw2; g.N = N2; g.K = D;
  g.C = static_cast<__nv_bfloat16*>(out); g.ldc = ldo; g.bias = bias2; g.act = act;
  g.stat_in = partials; g.stat_dim = D; g.stat_eps = eps; g.stat_rms = rms; g.colsum = rms ? nullptr : colsum2;
  rc = vla::gemm_launch(g, s, &err);
  return rc ? fail(rc, err) : 0;
}

int vla_op_center_crop_u8(const uint8_t* in, uint8_t* out, long long n_images, int H, int W, int out_size,
                          float crop_scale, void* stream) {
  const char* err = nullptr;
  int rc = vla::center_crop_u8_launch(in, out, n_images, H, W, out_size, crop_scale, static_cast<cudaStream_t>(stream), &err);
  return rc ? fail(rc, err) : 0;
}

int vla_op_rmsnorm(const void* x, int rows, int dim, int ldx, const float* w, float eps, void* y, int ldy,
                   void* stream) {
  const char* err = nullptr;
  int rc = vla::rmsnorm_launch(static_cast<const __nv_bfloat16*>(x), rows, dim, ldx, w, eps,
                               static_cast<__nv_bfloat16*>(y), ldy, static_cast<cudaStream_t>(stream), &err);
  return rc ? fail(rc, err) : 0;
}

int vla_op_attention(const void* qkv, int ld_qkv, int q_off, int k_off, int v_off, int B, int S, int n_heads,
                     int group, int hd, int causal, void* out, int ld_out, void* stream) {
  const char* err = nullptr;
  int rc = vla::attention_launch(static_cast<const __nv_bfloat16*>(qkv), ld_qkv, q_off, k_off, v_off, B, S,
                                 n_heads, group, hd, causal, static_cast<__nv_bfloat16*>(out), ld_out,
                                 static_cast<cudaStream_t>(stream), &err);
  return rc ? fail(rc, err) : 0;
}

int vla_op_gemm_rope(const void* A, int lda, int rows, const void* W, int ldw, int N, int K, void* C, int ldc,
                     const float* bias, const float* cos_t, const float* sin_t, int rope_cols, int S, void* stream) {
  vla::GemmArgs g;
  g.A = static_cast<const __nv_bfloat16*>(A);
  g.lda = lda;
  g.rows = rows;
  g.W = static_cast<const __nv_bfloat16*>(W);
  g.ldw = ldw;
  g.N = N;
  g.K = K;
  g.C = static_cast<__nv_bfloat16*>(C);
  g.ldc = ldc;
  g.bias = bias;
  g.rope_cols = rope_cols;
  g.rope_S = S;
  const char* err = nullptr;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!cos_t || !sin_t || S <= 0) return fail(-1, "gemm_rope: cos/sin tables required");
  uint32_t* cs = nullptr;  // the epilogue reads the table transposed and packed
  if (cudaMallocAsync(&cs, sizeof(uint32_t) * 32 * S, s) != cudaSuccess) return fail(-4, "gemm_rope: cudaMallocAsync failed");
  int rc = vla::rope_pack_launch(cos_t, sin_t, S, cs, s, &err);
  g.rope_cs = cs;
  if (!rc) rc = vla::gemm_launch(g, s, &err);
  cudaFreeAsync(cs, s);
  return rc ? fail(rc, err) : 0;
}

int vla_op_cross_attention(const void* q, int ld_q, int Sq, const void* k, const void* v, int ld_kv, int Skv, int B,
                           int n_heads, int group, int hd, int causal, void* out, int ld_out, void* stream) {
  const char* err = nullptr;
  int rc = vla::cross_attention_launch(static_cast<const __nv_bfloat16*>(q), ld_q, Sq, static_cast<const __nv_bfloat16*>(k),
                                       static_cast<const __nv_bfloat16*>(v), ld_kv, Skv, B, n_heads, group, hd, causal,
                                       static_cast<__nv_bfloat16*>(out), ld_out, static_cast<cudaStream_t>(stream), &err);
  return rc ? fail(rc, err) : 0;
}

int vla_set_attention_impl(int impl) {
  vla::attention_set_impl(impl);
  return 0;
}

int vla_op_rope(void* x, int ld, int off, int n_heads, int B, int S, float theta, void* stream) {
  const char* err = nullptr;
  int rc = vla::rope_launch(static_cast<__nv_bfloat16*>(x), ld, off, n_heads, B, S, theta,
                            static_cast<cudaStream_t>(stream), &err);
  return rc ? fail(rc, err) : 0;
}

}  // extern "C"
