/*
 * vla_b200.h - C ABI of libvla_b200.so, the sm_100a engine behind VLA-Adapter's predict_action path.
 *
 * The reference has no FFI layer: its seam is a set of plain Python method calls
 * (SURVEY.md section 8b).  Each entry point below names the reference interface it replaces.
 * Signatures carry only plain pointers and sizes (no torch / C++ types).  All device work is
 * enqueued on the caller-supplied cudaStream_t (passed as void*; NULL = legacy default stream).
 *
 * Conventions
 *   - every function returns 0 on success, a negative vla_status otherwise; the message is
 *     available through vla_last_error(engine) (engine functions) or vla_global_error() (ops);
 *   - the engine is not thread-safe; one engine per device;
 *   - the caller owns all I/O buffers; the engine owns its (repacked) weights and workspace;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef VLA_B200_H
#define VLA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum vla_status {
  VLA_OK = 0,
  VLA_ERR_INVALID = -1,      /* bad shape / argument (reference: Python assert / ValueError) */
  VLA_ERR_DTYPE = -2,        /* unsupported dtype */
  VLA_ERR_MISSING = -3,      /* a required tensor was never loaded */
  VLA_ERR_CUDA = -4,         /* CUDA runtime / driver error */
  VLA_ERR_NOT_FINALIZED = -5 /* vla_predict before vla_finalize */
} vla_status;

typedef enum vla_dtype { VLA_BF16 = 0, VLA_F32 = 1, VLA_F16 = 2 } vla_dtype;

typedef enum vla_variant {
  VLA_HEAD_BASE = 0, /* MLPResNetBlock      (prismatic/models/action_heads.py:168) */
  VLA_HEAD_PRO = 1   /* MLPResNetBlock_Pro  (prismatic/models/action_heads.py:287) */
} vla_variant;

/*
 * Engine configuration.  The reference bakes most of these into import-time globals
 * (prismatic/vla/constants.py:88-91) and the HF config (pretrained_models/configs/config.json);
 * here they are explicit.  Widths of the three sub-networks are fixed by the architecture
 * (DINOv2-L/14, SigLIP-so400m/14, Qwen2.5-0.5B, 8x112 policy heads); depths and vocabulary are
 * configurable so that reduced-depth models can be used in parity tests.
 */
typedef struct vla_cfg {
  int32_t n_images;      /* images per sample (1..3); NUM_PATCHES = 256 * n_images (MP:953) */
  int32_t chunk_len;     /* NUM_ACTIONS_CHUNK (constants.py:29)  */
  int32_t action_dim;    /* ACTION_DIM                              */
  int32_t proprio_dim;   /* PROPRIO_DIM                             */
  int32_t variant;       /* vla_variant                             */
  int32_t dino_depth;    /* timm blocks in the DINOv2 tower (24); output taken after block depth-2 */
  int32_t siglip_depth;  /* timm blocks in the SigLIP tower (27)    */
  int32_t llm_layers;    /* Qwen2 decoder layers (24; the policy needs exactly 24 taps) */
  int32_t vocab_size;    /* rows of embed_tokens (151936)           */
  int32_t max_batch;     /* workspace is sized for this many samples per call */
  int32_t max_prompt_len;/* and this many prompt tokens (L)         */
  int32_t causal;        /* 1 = causal LLM attention (stock HF Qwen2), 0 = bidirectional */
} vla_cfg;

typedef struct vla_engine vla_engine;

/* Replaces: OpenVLAForActionPrediction.__init__ (modeling_prismatic.py:736) + get_action_head /
 * get_proprio_projector (experiments/robot/openvla_utils.py:482, 412) - allocates an empty engine. */
int vla_create(const vla_cfg* cfg, vla_engine** out);

/* Replaces: load_state_dict of the three reference modules (openvla_utils.py:230-250, 445, 530).
 * `name` is the reference state_dict key, prefixed by its module:
 *   "vla."  + OpenVLAForActionPrediction key   (e.g. vla.vision_backbone.featurizer.blocks.0.attn.qkv.weight)
 *   "head." + L1RegressionActionHead key       (e.g. head.model.mlp_resnet_blocks.3.q_proj.weight)
 *   "proprio." + ProprioProjector key          (e.g. proprio.fc1.weight)
 * `ptr` may be a device or host pointer; the engine copies and repacks, the caller keeps ownership.
 * Unknown names are rejected with VLA_ERR_INVALID, names the path does not read (e.g. lm_head,
 * film_gen, attn_pool) are accepted and ignored. */
int vla_load_tensor(vla_engine* e, const char* name, const void* ptr, int dtype, int ndim,
                    const int64_t* shape);

/* Replaces: OpenVLAForActionPrediction._unnormalize_actions statistics lookup
 * (modeling_prismatic.py:786-805, 998-1001).  hi/lo are q99/q01 (BOUNDS_Q99) or max/min (BOUNDS);
 * mask[i] != 0 means "un-normalise dimension i".  All arrays have action_dim entries (host). */
int vla_set_action_stats(vla_engine* e, const double* hi, const double* lo, const uint8_t* mask);

/* Checks that every tensor is present, precomputes input-independent constants (the MLPResNet
 * prologue x0 = ReLU(fc1(LN(0))), action_heads.py:114-116; RoPE tables; tanh(gating_factor)),
 * allocates the workspace for (max_batch, max_prompt_len). */
int vla_finalize(vla_engine* e);

/* Replaces: OpenVLAForActionPrediction.predict_action (modeling_prismatic.py:892-972), batched.
 *   pixel_values : device, bf16, (B, 6*n_images, 224, 224) contiguous   (MP:919)
 *   ext_ids      : device, int64, (B, L+65) = prompt ids, 64 placeholder ids, STOP   (MP:748-758)
 *   aq_index     : device, int32, (B, L+65): ActionQuery row to splice at that column or -1
 *                  (the all_actions_mask / masked_indices of MP:442-452, as indices)
 *   prompt_len   : device, int32, (B) prompt lengths L_b in [1, L], or NULL when every prompt has L tokens.  With
 *                  lengths, row b of ext_ids / aq_index is RIGHT-padded: [ids(L_b) | 64 placeholders | STOP | pad], pad
 *                  columns holding any valid id and aq_index -1.  The reference is bs=1 and never pads (MP:855); chat
 *                  prompts span 40-56 tokens (openvla_utils.py:783), so a served batch mixes lengths.  Causal attention
 *                  keeps every real row identical to the un-padded run of that sample.
 *   proprio      : device, fp32, (B, proprio_dim), already normalised (openvla_utils.py:671-701)
 *   out_norm     : device, fp32, (B, chunk_len, action_dim) normalised actions      (MP:871-872)
 *   out_unnorm   : device, fp32, (B, chunk_len, action_dim) un-normalised actions   (MP:799-803)
 *   out_last_ha  : device, bf16, (B, 64, 896) last-layer ActionQuery states or NULL (MP:855, 972)
 */
int vla_predict(vla_engine* e, const void* pixel_values, const int64_t* ext_ids,
                const int32_t* aq_index, const int32_t* prompt_len, const float* proprio, int B, int L,
                float* out_norm, float* out_unnorm, void* out_last_ha, void* stream);

/* Same call with HOST buffers (pinned or pageable): the engine stages inputs through its own
 * pinned buffers, runs vla_predict and copies the results back; synchronises the stream.
 * This is the end-to-end path bench.py reports as `e2e`. */
int vla_predict_host(vla_engine* e, const void* pixel_values, const int64_t* ext_ids,
                     const int32_t* aq_index, const int32_t* prompt_len, const float* proprio, int B, int L,
                     float* out_norm, float* out_unnorm, void* out_last_ha, void* stream);

/* Device-side image front-end: the same forward from uint8 frames.  `images` is
 * (B, n_images, 224, 224, 3) uint8, HWC, already resized / centre-cropped like PrismaticImageProcessor.apply_transform
 * does on the CPU (prismatic/extern/hf/processing_prismatic.py:128-145); the per-backbone ToTensor + Normalize and the
 * bf16 cast of get_vla_action (experiments/robot/openvla_utils.py:786-796) are applied on the device through a table
 * built with the processor's fp32 arithmetic, so results are bit-identical to feeding the CPU-normalised pixel_values
 * to vla_predict.  vla_predict_u8 takes device pointers, vla_predict_host_u8 host pointers (H2D / D2H inside). */
int vla_predict_u8(vla_engine* e, const uint8_t* images, const int64_t* ext_ids, const int32_t* aq_index,
                   const int32_t* prompt_len, const float* proprio, int B, int L, float* out_norm, float* out_unnorm,
                   void* out_last_ha, void* stream);
int vla_predict_host_u8(vla_engine* e, const uint8_t* images, const int64_t* ext_ids, const int32_t* aq_index,
                        const int32_t* prompt_len, const float* proprio, int B, int L, float* out_norm,
                        float* out_unnorm, void* out_last_ha, void* stream);
/* Normalisation statistics of the two backbones, mean[2][3] / std[2][3] (backbone 0 = DINOv2, 1 = SigLIP; the
 * defaults are those of pretrained_models/configs/preprocessor_config.json).  May be called before or after
 * vla_finalize. */
int vla_set_image_norm(vla_engine* e, const float* mean, const float* stdv);

/* Device-side centre crop for the uint8 entry points: with crop_scale > 0 (the reference uses 0.9,
 * experiments/robot/openvla_utils.py:616-648 `center_crop_image` -> :568-613 `crop_and_resize`) every 224 x 224 frame
 * handed to vla_predict_u8 / vla_predict_host_u8 is first replaced by the centred box of that relative AREA, resampled
 * bilinearly to 224 x 224 with TensorFlow's crop_and_resize arithmetic and converted back to uint8 like
 * tf.image.convert_image_dtype - bit-exact against the CPU restatement oracle/image_prep.py (TensorFlow itself is not
 * available offline: that restatement is unpinned).  0 switches it off.  After vla_finalize.  What stays on the CPU
 * is the first half of the reference's preparation: the JPEG round trip and the lanczos3 antialiased resize from the
 * camera resolution (:560-565), which depend on TensorFlow's JPEG codec.
 * vla_op_center_crop_u8 is the kernel alone: (n, H, W, 3) uint8 -> (n, out, out, 3) uint8, device pointers. */
int vla_set_center_crop(vla_engine* e, float crop_scale);
int vla_op_center_crop_u8(const uint8_t* in, uint8_t* out, long long n_images, int H, int W, int out_size,
                          float crop_scale, void* stream);

/* Per-subsystem timing of the forward, for bench.py: with vla_segment_timing(e, 1) the next calls run eagerly (no
 * graph replay) with CUDA events at the subsystem boundaries; vla_segment_times returns the device time [ms] of the
 * last call's (0) vision towers + projector, (1) LLM input assembly + Qwen2.5 prefill, (2) Bridge-Attention policy. */
int vla_segment_timing(vla_engine* e, int enable);
int vla_segment_times(vla_engine* e, float* ms3);

/* Debug taps for parity tests: copies an intermediate of the LAST vla_predict call to `dst`
 * (device or host).  Names: "patches" (B,NP,2176) tower output, "projected" (B,NP,896),
 * "llm_in" (B,S,896), "hidden.<i>" i in 1..24 (B,S,896) as returned by HF output_hidden_states,
 * "head_x.<i>" policy state after block i (B,T,896).  Returns the number of bytes through *bytes. */
int vla_get_tap(vla_engine* e, const char* name, void* dst, size_t capacity, size_t* bytes);

/* The device-pointer calls (vla_predict, vla_predict_u8) only ENQUEUE work, like the reference's CUDA calls behind
 * predict_action before its .cpu() at MP:872: what the kernels find out (a token id outside the vocabulary - the
 * reference's embedding lookup would raise a device assert - or a bad ActionQuery index) is reported by this call,
 * which waits for `stream`, returns VLA_ERR_INVALID / VLA_ERR_CUDA accordingly and clears the flag.  The offending
 * row of the LLM input is zero-filled, never left stale.  vla_predict_host* perform this check themselves. */
int vla_check_errors(vla_engine* e, void* stream);

/* Kernel launches issued by this engine's last vla_predict call (own kernels only). */
long long vla_last_launch_count(const vla_engine* e);

const char* vla_last_error(const vla_engine* e);
void vla_destroy(vla_engine* e);

/* ------------------------------------------------------------------------------------------
 * Operator-level entry points (used by the parity tests; each is one of the engine's kernels).
 * All pointers are device pointers, bf16 unless noted.
 * ------------------------------------------------------------------------------------------ */

/* C[b,r,n] = epi(sum_k A[b,r,k] W[n,k]) - nn.Linear; see csrc/gemm.cuh for the epilogue.
 * act: 0 none, 1 GELU(erf), 2 ReLU, 3 SwiGLU (interleaved gate/up rows). */
int vla_op_gemm(const void* A, long long a_batch_stride, int lda, int rows, int batches,
                const void* W, int ldw, int N, int K, void* C, long long c_batch_stride, int ldc,
                const float* bias, const float* colscale, const void* resid,
                long long r_batch_stride, int ldr, int act, int force_bn, void* stream);

/* y = LayerNorm(x) * w + b over the last dim (nn.LayerNorm, eps given); w, b fp32. */
int vla_op_layernorm(const void* x, int rows, int dim, int ldx, const float* w, const float* b,
                     float eps, void* y, int ldy, void* stream);

/* y = x * rsqrt(mean(x^2) + eps) * w (Qwen2RMSNorm); w fp32. */
int vla_op_rmsnorm(const void* x, int rows, int dim, int ldx, const float* w, float eps, void* y,
                   int ldy, void* stream);

/* A LayerNorm / RMSNorm in front of a Linear, folded into the GEMM (how the engine runs norm1 -> qkv, norm2 -> fc1,
 * input_layernorm -> q|k|v and post_attention_layernorm -> gate|up; csrc/gemm.cuh GemmArgs::row_stats).
 * vla_op_fold_norm, once per weight: W[n,k] <- bf16(W[n,k] * norm_w[k]) IN PLACE, bias[n] += sum_k W[n,k] norm_b[k]
 *   (norm_b NULL for RMSNorm; bias fp32, required when norm_b is given), colsum[n] = sum_k W'[n,k] (fp32, may be NULL
 *   for RMSNorm).
 * vla_op_norm_gemm: C = act(Linear(Norm(x))) from the RAW rows x [rows, K] and the folded W / bias / colsum;
 *   rms != 0 selects RMSNorm; stats is scratch for 2*rows floats. act as in vla_op_gemm (3 = SwiGLU, RMSNorm only). */
int vla_op_fold_norm(void* W, int N, int K, int ldw, const float* norm_w, const float* norm_b, float* bias,
                     float* colsum, void* stream);
int vla_op_norm_gemm(const void* x, int rows, int ldx, const void* W, int ldw, int N, int K, void* C, int ldc,
                     const float* bias, const float* colsum, int rms, float eps, int act, float* stats,
                     void* stream);

/* The tail of a transformer block the way the engine chains it, statistics kernel-free (csrc/gemm.cuh stat_out /
 * stat_in): (1) x[rows, D] += colscale1 * (a @ W1^T + bias1) IN PLACE - the residual box is TMA-loaded into the
 * epilogue's staging box and added in fp32 - while the epilogue leaves 12 partial (sum, sum of squares) pairs per row in
 * `partials` ([rows][12][2] fp32); (2) out = act(Linear(Norm(x))) with the norm folded into W2 / bias2 / colsum2
 * (vla_op_fold_norm) and mean / rstd finished from the partials in the second GEMM's epilogue.  rms != 0: RMSNorm. */
int vla_op_block_tail(const void* a, int lda, int rows, const void* W1, int ldw1, int K1, void* x, int D,
                      const float* bias1, const float* colscale1, const void* W2, int ldw2, int N2, void* out, int ldo,
                      const float* bias2, const float* colsum2, int rms, float eps, int act, float* partials,
                      void* stream);

/* Multi-head attention over a packed qkv buffer.
 *   q at qkv[row, q_off + h*hd], k at qkv[row, k_off + (h/group)*hd], v likewise; row = b*S + s.
 *   hd in {64, 72}; causal != 0 applies the lower-triangular mask; scale = hd^-0.5. */
int vla_op_attention(const void* qkv, int ld_qkv, int q_off, int k_off, int v_off, int B, int S,
                     int n_heads, int group, int hd, int causal, void* out, int ld_out,
                     void* stream);

/* C[m, :] = rope(A[m, :] @ W^T + bias): the Qwen2 q/k/v projection with HF rotate_half RoPE fused into the epilogue.
 * Output columns [0, rope_cols) are heads of width 64 rotated with position m % S; cos_t / sin_t are [S][32] fp32
 * tables holding bf16 values (angle = pos * theta^(-2j/64)); columns >= rope_cols (the v heads) pass through. */
int vla_op_gemm_rope(const void* A, int lda, int rows, const void* W, int ldw, int N, int K, void* C, int ldc,
                     const float* bias, const float* cos_t, const float* sin_t, int rope_cols, int S, void* stream);

/* General form: q rows [b*Sq, (b+1)*Sq) of `q` (head h at column h*hd), k / v rows [b*Skv, (b+1)*Skv) of `k` / `v`
 * (kv head h/group at column (h/group)*hd); hd in {64, 72, 112}.  With Sq = chunk_len, Skv = chunk_len + 65 + 256n,
 * hd = 112 this is the Bridge-Attention core of MLPResNetBlock (prismatic/models/action_heads.py:256-279). */
int vla_op_cross_attention(const void* q, int ld_q, int Sq, const void* k, const void* v, int ld_kv, int Skv, int B,
                           int n_heads, int group, int hd, int causal, void* out, int ld_out, void* stream);

/* Attention kernel selection for vla_op_attention and the engine: 0 = auto (tcgen05/TMEM/TMA kernel for head
 * dims 64 and 72, the small mma.sync kernel for the policy's 8-query hd-112 cross-attention), 1 = mma.sync
 * kernel everywhere (A/B comparison), 2 = tcgen05 kernel whenever the head dim allows.  Returns 0. */
int vla_set_attention_impl(int impl);

/* In-place HF rotate_half RoPE (theta) on `n_heads` heads of width 64 starting at column `off`
 * of each row; position = row % S. */
int vla_op_rope(void* x, int ld, int off, int n_heads, int B, int S, float theta, void* stream);

/* Per-launch CUDA-event timing of the tcgen05 GEMM kernel, for bench.py's roofline object: after
 * vla_profile_gemm(1) every GEMM launch is bracketed by events on its own stream;
 * vla_profile_gemm_read() synchronises them, returns the summed kernel time [ms] and the number of
 * launches since the last read, and clears the records. */
int vla_profile_gemm(int enable);
int vla_profile_gemm_read(double* total_ms, long long* launches);

const char* vla_global_error(void);
long long vla_total_launch_count(void);

/* Device watchdog.  Every mbarrier wait inside the tcgen05 kernels is bounded (default 10 s; VLA_WATCHDOG_MS or
 * vla_watchdog_set_timeout_ms override): on timeout the waiting warp records (kernel, role, barrier, parity, CTA,
 * thread, SM) in a host-mapped buffer and traps, so a protocol deadlock ends as VLA_ERR_CUDA with a message that names
 * the barrier instead of a hung process (the reference relies on torch.distributed's timeout for the same purpose,
 * vla-scripts/evaluate_calvin.py:877).  vla_watchdog_report copies the records so far (text, NUL-terminated, truncated
 * to `capacity`) and returns the full length (0 = none); vla_watchdog_selftest launches a kernel that deadlocks on
 * purpose and returns VLA_ERR_CUDA once the watchdog has fired (the CUDA context is unusable afterwards: tests run
 * it in a child process). */
int vla_watchdog_report(char* buf, size_t capacity);
int vla_watchdog_set_timeout_ms(int ms);
int vla_watchdog_selftest(void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VLA_B200_H */
