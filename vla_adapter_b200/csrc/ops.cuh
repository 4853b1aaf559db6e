// Host-side launchers of the bandwidth-bound / small kernels (ops.cu, attention.cu, policy.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace vla {

long long ops_launch_count();
void ops_count_launch(int n = 1);

// nn.LayerNorm over the last dim; one warp per row, fp32 statistics, bf16 in/out.
int layernorm_launch(const __nv_bfloat16* x, int rows, int dim, int ldx, const float* w, const float* b,
                     float eps, __nv_bfloat16* y, int ldy, cudaStream_t s, const char** err);
// Same, for a 3-D row view (batch slabs), used to normalise "rows [r0, r0+rows) of every image".
int layernorm_launch_3d(const __nv_bfloat16* x, int rows, int batches, long long x_bs, int dim, int ldx,
                        const float* w, const float* b, float eps, __nv_bfloat16* y, long long y_bs, int ldy,
                        cudaStream_t s, const char** err);

// Qwen2RMSNorm (transformers Qwen2RMSNorm.forward): bf16(x * rsqrt(mean x^2 + eps)) * w.
// Norm folded into the next GEMM (gemm.cuh: GemmArgs::row_stats): per-row (rstd, -mean * rstd) of x, and the
// one-time fold of the norm's weight / bias into that GEMM's W / bias (+ the column sums its epilogue needs).
// partial_slots > 0: instead of (rstd, -mean * rstd) write the producer-GEMM format of gemm.cuh (STAT_SLOTS partial
// (sum x, sum x^2) pairs per row, everything in slot 0) - for the first layer of a stack, whose rows no GEMM produced.
int row_stats_launch(const __nv_bfloat16* x, int rows, int dim, int ldx, int rms, float eps, float* stats,
                     cudaStream_t s, const char** err, int partial_slots = 0);
int fold_norm_launch(__nv_bfloat16* W, int N, int K, int ldw, const float* g, const float* b, float* bias,
                     float* colsum, cudaStream_t s, const char** err);
int rmsnorm_launch(const __nv_bfloat16* x, int rows, int dim, int ldx, const float* w, float eps,
                   __nv_bfloat16* y, int ldy, cudaStream_t s, const char** err);

// cos/sin tables (fp32 values already rounded to bf16), [S][half] each; inv_freq_j = theta^(-2j/(2*half)).
// [S][32] fp32 tables -> the transposed packed table of the GEMM's RoPE epilogue (gemm.cuh: GemmArgs::rope_cs)
int rope_pack_launch(const float* cos_t, const float* sin_t, int S, uint32_t* cs, cudaStream_t s, const char** err);
int rope_table_launch(float* cos_t, float* sin_t, int S, int half, float theta, cudaStream_t s,
                      const char** err);
// HF rotate_half RoPE, in place, width-64 heads: out = bf16(bf16(x*cos) + bf16(rot(x)*sin)).
int rope_apply_launch(__nv_bfloat16* x, int ld, int off, int n_heads, int B, int S, const float* cos_t,
                      const float* sin_t, cudaStream_t s, const char** err);
int rope_launch(__nv_bfloat16* x, int ld, int off, int n_heads, int B, int S, float theta, cudaStream_t s,
                const char** err);

// Flash attention (attention.cu). hd in {64, 72}.
int attention_launch(const __nv_bfloat16* qkv, int ld_qkv, int q_off, int k_off, int v_off, int B, int S,
                     int n_heads, int group, int hd, int causal, __nv_bfloat16* out, int ld_out,
                     cudaStream_t s, const char** err);

// General form: q rows [b*Sq, (b+1)*Sq) of `q` (head h at column h*hd), k/v rows [b*Skv, (b+1)*Skv) of
// `k` / `v` (kv head h/group at column (h/group)*hd).  hd in {64, 72, 112}.  Used with Sq = chunk_len,
// Skv = chunk_len + 65 + NP, hd = 112 for the Bridge-Attention core of the policy head.
int cross_attention_launch(const __nv_bfloat16* q, int ld_q, int Sq, const __nv_bfloat16* k, const __nv_bfloat16* v,
                           int ld_kv, int Skv, int B, int n_heads, int group, int hd, int causal,
                           __nv_bfloat16* out, int ld_out, cudaStream_t s, const char** err);

// tcgen05/TMEM/TMA flash attention (attention_tc.cu) for hd in {64, 72}; returns 1 if the shape is not served.
int attention_tc_launch(const __nv_bfloat16* q, int ld_q, int Sq, const __nv_bfloat16* k, const __nv_bfloat16* v,
                        int ld_kv, int Skv, int B, int n_heads, int group, int hd, int causal, __nv_bfloat16* out,
                        int ld_out, int q_rows, cudaStream_t s, const char** err);  // q_rows: rows between samples of q/out
// 0 = auto (tcgen05 for hd 64/72 with Sq >= 32, mma.sync otherwise), 1 = mma.sync only, 2 = tcgen05 whenever possible
void attention_set_impl(int impl);

// Patch-embed im2col: pixel_values (B, 6*n_img, 224, 224) bf16 -> A[(b*n_img+i)*256 + p, 592]
// with k = c*196 + ky*14 + kx (Conv2d weight order), columns 588..591 zero.  tower 0 reads channels
// [6i, 6i+3) (DINOv2), tower 1 reads [6i+3, 6i+6) (SigLIP)  (modeling_prismatic.py:220-230).
int im2col_launch(const __nv_bfloat16* pix, int B, int n_img, int tower, __nv_bfloat16* out, cudaStream_t s,
                  const char** err);

// Same, from uint8 HWC frames (B, n_img, 224, 224, 3): value = lut[tower][c][u8] (ToTensor + Normalize + bf16 cast
// tabulated by the host; lut is [2][3][256] bf16).
int im2col_u8_launch(const uint8_t* img, int B, int n_img, int tower, const __nv_bfloat16* lut, __nv_bfloat16* out,
                     cudaStream_t s, const char** err);

// The reference's centre crop (openvla_utils.py:616-648): (n, H, W, 3) uint8 -> (n, out, out, 3) uint8, the centred box of
// area `crop_scale` resampled bilinearly with TensorFlow's crop_and_resize arithmetic (bit-exact against
// oracle/image_prep.py).
int center_crop_u8_launch(const uint8_t* in, uint8_t* out, long long n_images, int H, int W, int out_size,
                          float crop_scale, cudaStream_t s, const char** err);

// Writes the DINOv2 prefix rows (cls + 4 register tokens, no pos-embed) of every image slab.
int prefix_tokens_launch(__nv_bfloat16* x, int n_slabs, long long slab_stride, int dim,
                         const __nv_bfloat16* prefix /*[5, dim]*/, int n_prefix, cudaStream_t s,
                         const char** err);

// LLM input assembly (modeling_prismatic.py:418-454, 500-502): for sequence position s of sample b
//   s == 0           -> embed[ext_ids[b,0]]
//   1 <= s <= NP     -> left untouched (projector output is written there by the GEMM epilogue)
//   s > NP, j = s-NP -> aq_index[b,j] >= 0 ? action_queries[aq_index[b,j]] : embed[ext_ids[b,j]]
int assemble_launch(__nv_bfloat16* x, int B, int S, int NP, int Lext, int dim, const int64_t* ext_ids,
                    const int32_t* aq_index, const __nv_bfloat16* embed, int vocab,
                    const __nv_bfloat16* aq_table, int n_aq, int* err_flag, cudaStream_t s, const char** err);

// out[m, n] = act(sum_k x[m,k] W[n,k] + b[n]) for skinny problems (tiny M or N); one warp per output.
// x may be fp32 (x_is_f32) - it is rounded to bf16 first (reference casts proprio to bf16, AH:53).
// out_f32 != null additionally receives the bf16-rounded value as fp32.
int skinny_linear_launch(const void* x, int x_is_f32, int ldx, int M, int K, const __nv_bfloat16* W, int ldw,
                         int N, const float* bias, int act, __nv_bfloat16* out, int ldo, float* out_f32,
                         cudaStream_t s, const char** err);

// Pro-variant RoPE (action_heads.py:125-164, 381-386), in place: q rows (B*T, ld 896) at positions t, and the K
// half of the per-sample key/value buffer kv [B][T+65+NP][1792]: self rows at positions 0..T-1, the 65
// h_a ++ p rows at 0..64, the NP h_t rows at 0..NP-1.
// kv_only != 0: only the cond / vision key rows [T, T+65+NP) (q may be null).
int policy_rope_launch(__nv_bfloat16* q, __nv_bfloat16* kv, int B, int T, int NP, const float* cos_t,
                       const float* sin_t, cudaStream_t s, const char** err, int kv_only = 0);

// Final regression epilogue: out_norm[b,t,a] = fc2(LN(x)) ; out_unnorm = where(mask, 0.5*(a+1)*(hi-lo+1e-8)+lo, a)
int head_out_launch(const __nv_bfloat16* x, int rows, const float* ln_w, const float* ln_b,
                    const __nv_bfloat16* W /*[A, 896]*/, const float* bias, int A, const float* hi,
                    const float* lo, const uint8_t* mask, float* out_norm, float* out_unnorm, cudaStream_t s,
                    const char** err);

// Pro-variant RoPE table (action_heads.py:150-164): angle(t, j) = t * inv_freq[j mod 56], hd = 112.
int policy_rope_table_launch(float* cos_t, float* sin_t, int max_pos, cudaStream_t s, const char** err);

// dst[r, :] = src[:] for r in [0, rows)  (the input-independent MLPResNet prologue x0, AH:114-116)
int broadcast_row_launch(const __nv_bfloat16* src, int dim, int rows, __nv_bfloat16* dst, cudaStream_t s,
                         const char** err);

// Copies rows [r0, r0+rows) of every slab into a dense buffer (tap extraction).  With `len` (device, one prompt length
// per slab) the window of slab b is shifted by len[b] - len_ref: samples whose prompts have different lengths keep their
// ActionQuery rows at different offsets of the (right-padded) LLM sequence.
int gather_rows_launch(const __nv_bfloat16* src, long long src_bs, int ld, int r0, int rows, int batches,
                       int dim, __nv_bfloat16* dst, cudaStream_t s, const char** err, const int32_t* len = nullptr,
                       int len_ref = 0, int* err_flag = nullptr);

}  // namespace vla
