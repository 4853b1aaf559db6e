// Host interface of the single-kernel Bridge-Attention policy for small batches (policy_fused.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stddef.h>

namespace vla {

struct PolicyBlockW {
  const __nv_bfloat16 *wq, *wkvs, *wo, *wffn;  // [896][896], [1792][896] (K rows then V rows), [896][896], [896][896]
  const float *bq, *bkvs, *bo, *lnw, *lnb, *bffn;
  __nv_bfloat16* kv;     // [B][NK][1792]: this block's key/value buffer (rows >= T already projected)
  __nv_bfloat16* x_out;  // [B*T][896]: the policy state after this block
};

constexpr int POLICY_FUSED_MAX_BLOCKS = 24;

// Passed by value as the kernel's (grid-constant) parameter: the per-block pointers are read from the constant bank at
// the point of use instead of living in registers across a block.  ~2.4 KB of the 4 KB parameter space.
struct PolicyFusedArgs {
  PolicyBlockW blocks[POLICY_FUSED_MAX_BLOCKS];
  int n_blocks;
  const __nv_bfloat16* x0;     // [B*T][896] state before block 0
  __nv_bfloat16* ao;           // [B*T][896] scratch: attention output
  __nv_bfloat16* y;            // [B*T][896] scratch: o-projection + residual
  int T, NK, pro;
  const float *rope_cos, *rope_sin;  // Pro variant: [pos][112] tables (policy_rope_table_launch)
  float scale_log2, ln_eps;
  int B;                 // samples = worker clusters; the CTAs beyond B * 8 are L2 prefetchers
  int* progress;         // device word, 0 between launches: the block the workers of sample 0 are in
  long long* prof;       // optional [n_blocks][8] clock64 stamps of CTA 0 (VLA_POLICY_PROF=1), else nullptr
};

size_t policy_fused_smem_bytes();
// One cluster of 8 CTAs per sample (+ 4 clusters of L2 prefetchers): every policy block in one launch.  T <= 16.
// prefetch_clusters < 0: the default (4); fewer when the launch shares the GPU with other work.
int policy_fused_launch(const PolicyFusedArgs& a, int B, cudaStream_t s, const char** err, int prefetch_clusters = -1);

}  // namespace vla
