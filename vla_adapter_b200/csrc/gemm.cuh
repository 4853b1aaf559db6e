// Host-side interface of the tcgen05 GEMM (see gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace vla {

enum GemmAct : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_SWIGLU = 3 };

// Row statistics without a statistics kernel: a GEMM that writes rows of a residual stream leaves, per row, STAT_SLOTS
// partial (sum x, sum x^2) pairs (one per column tile half, unused slots zero); the GEMM that consumes those rows
// behind a folded LayerNorm / RMSNorm sums the slots in its epilogue.  [rows][STAT_SLOTS] float2.
constexpr int STAT_SLOTS = 12;

// C[b, r, n] = epi( sum_k A[b, r, k] * W[n, k] )      (nn.Linear convention: W is [N, K], K contiguous)
//
// A and C are 3-D *views*: `batches` slabs of `rows` rows each, slab b starting at
// A + b * a_batch_stride (elements).  This is how the engine reads "rows [r0, r0+rows) of every
// sample" of a hidden state (reference slices at modeling_prismatic.py:855,858) and how it writes
// patch rows behind the prefix tokens / into the LLM input buffer without gather copies.
//
// Epilogue (fp32): v = acc + bias[n]; v = act(v); v = colscale[n] * v; v += resid[b, r, n]; C = bf16(v)
// ACT_SWIGLU: W rows are interleaved in groups of 16 (g0..15, u0..15, g16.., u16..); the epilogue
// writes silu(g) * u to column n/2 (bias/colscale/resid are not applied in this mode).
struct GemmArgs {
  const __nv_bfloat16* A = nullptr;
  long long a_batch_stride = 0;  // elements
  int lda = 0;                   // elements, multiple of 8
  int rows = 0;                  // rows per batch slab
  int batches = 1;
  const __nv_bfloat16* W = nullptr;
  int ldw = 0;  // elements, multiple of 8
  int N = 0;    // multiple of 8
  int K = 0;    // logical K (TMA zero-fills the tail of the last 64-wide K block)
  __nv_bfloat16* C = nullptr;
  long long c_batch_stride = 0;
  int ldc = 0;
  const float* bias = nullptr;      // [N] fp32
  const float* colscale = nullptr;  // [N] fp32 (LayerScale)
  const __nv_bfloat16* resid = nullptr;
  long long r_batch_stride = 0;
  int ldr = 0;
  int act = ACT_NONE;
  int force_bn = 0;  // 0 = heuristic, else 64 / 128 / 192 / 224 / 256
  // HF rotate_half RoPE fused into the epilogue (Qwen2 q/k projection): output columns [0, rope_cols) are heads of
  // width 64 rotated with the cos/sin of position (row % rope_S).  rope_cs is the table TRANSPOSED and packed,
  // [32][rope_S] words of (bf16 cos | bf16 sin << 16) (rope_pack_launch): the 32 lanes of an epilogue warp are 32
  // consecutive rows = positions, so one load instruction reads one line.
  const uint32_t* rope_cs = nullptr;
  int rope_ld = 0;  // positions per table row (>= rope_S; 0 = rope_S)
  int rope_cols = 0;
  int rope_S = 0;
  // LayerNorm / RMSNorm of A folded into the GEMM: W already holds W * diag(norm weight) and `bias` holds
  // bias + W @ norm_bias; the epilogue turns the raw product into the product of the NORMALISED rows,
  //   acc <- rstd[r] * acc + (-mean[r] * rstd[r]) * colsum[n],
  // with (rstd, -mean * rstd) per row from row_stats_launch() and colsum[n] = sum_k W'[n, k] (nullptr for RMSNorm,
  // whose mean term is absent).  One row view only (batches == 1).
  const float* row_stats = nullptr;  // [rows][2] fp32
  const float* colsum = nullptr;     // [N] fp32
  // The same folded norm with the statistics taken from the PRODUCER's partial sums instead of row_stats:
  // stat_in = [rows][STAT_SLOTS] float2 written by an earlier gemm_launch with stat_out; the epilogue computes
  // mean = sum x / stat_dim, rstd = rsqrt(E[x^2] - mean^2 + stat_eps)  (stat_rms: rstd = rsqrt(E[x^2] + eps), no mean).
  const float* stat_in = nullptr;
  int stat_dim = 0;
  float stat_eps = 0.f;
  int stat_rms = 0;
  // Producer side: partial (sum, sum of squares) of every output row of THIS GEMM (the values before their bf16
  // rounding), one row view only; restricts the tile width so that 2 * column tiles <= STAT_SLOTS.
  float* stat_out = nullptr;
  // Residual handling: -1 auto, 0 legacy (in place: bf16 TMA reduce-add into C; out of place: per-thread row loads),
  // 1 staged (the residual box is TMA-loaded into the epilogue's staging box, added in fp32, stored by TMA).
  int resid_staged = -1;
};

// Returns 0 on success, negative on error (message in *err if non-null).
int gemm_launch(const GemmArgs& a, cudaStream_t stream, const char** err);

// dst view = src view: `rows` x `cols` bf16 per batch, 16-byte vectors; the source batch stride may be 0.
int copy_view_launch(const __nv_bfloat16* src, long long s_bs, int lds, __nv_bfloat16* dst, long long d_bs, int ldd,
                     int rows, int batches, int cols, cudaStream_t stream, const char** err);

// Number of GEMM kernel launches issued since process start (for bench.py's gpu_launches).
long long gemm_launch_count();

// Per-launch CUDA-event timing of the GEMM kernel (bench.py roofline leg).
void gemm_profile_enable(bool on);
bool gemm_profile_enabled();
int gemm_profile_read(double* total_ms, long long* launches);

}  // namespace vla
