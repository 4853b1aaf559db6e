// Bridge-Attention core of the policy head and the final regression epilogue.
//
// Reference: MLPResNetBlock.forward (prismatic/models/action_heads.py:218-283) and
// MLPResNetBlock_Pro.forward (:337-410).  Per (sample, head) the T chunk queries attend, in ONE softmax,
// over three key/value segments:
//     self   : the T rows of x                       (scale 1)
//     cond   : 64 ActionQuery rows h_a ++ 1 proprio row p   (scale 1)       [base "task", Pro "adapter"]
//     vision : NP raw rows h_t                       (scale tanh(gating_factor))  [base "adapter", Pro "task"]
// scores = [q k_self^T | q k_cond^T | g * q k_vis^T] / sqrt(112) -> softmax -> weighted sum of V.
// Pro additionally rotates q/k_self (positions 0..T-1), k_cond (0..64) and k_vis (0..NP-1) with the
// interleaved-pair / concat-frequency RoPE of action_heads.py:125-164.
// The K/V projections of all segments are produced by the tcgen05 GEMM; this kernel is the small
// irregular part (T <= 32 query rows), done on CUDA cores with fp32 accumulation.
#include "common.cuh"
#include "ops.cuh"

namespace vla {

namespace {

constexpr int PH = 8;       // heads
constexpr int PHD = 112;    // head dim
constexpr int PD = 896;
constexpr int PKV = 1792;   // K | V row
constexpr int P_THREADS = 256;
constexpr int P_MAXT = 32;

// Loads one 112-wide bf16 row (16-byte aligned) into fp32 registers.
VLA_DEVINL void load_row112(const __nv_bfloat16* p, float (&f)[PHD]) {
#pragma unroll
  for (int c = 0; c < PHD / 8; ++c) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + c);
    const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
    f[c * 8 + 0] = a0.x; f[c * 8 + 1] = a0.y; f[c * 8 + 2] = a1.x; f[c * 8 + 3] = a1.y;
    f[c * 8 + 4] = a2.x; f[c * 8 + 5] = a2.y; f[c * 8 + 6] = a3.x; f[c * 8 + 7] = a3.y;
  }
}

// apply_rope (action_heads.py:125-146) at position `pos`: pairs are (2i, 2i+1), the angle of lane j is
// pos * inv_freq[j mod 56]; result rounded to bf16 like the eager bf16 reference.
VLA_DEVINL void rope112(float (&f)[PHD], const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                        int pos) {
  const float* c = cos_t + pos * PHD;
  const float* s = sin_t + pos * PHD;
#pragma unroll
  for (int i = 0; i < PHD / 2; ++i) {
    const float x0 = f[2 * i], x1 = f[2 * i + 1];
    f[2 * i] = bf16_round(bf16_round(x0 * c[2 * i]) + bf16_round(-x1 * s[2 * i]));
    f[2 * i + 1] = bf16_round(bf16_round(x1 * c[2 * i + 1]) + bf16_round(x0 * s[2 * i + 1]));
  }
}

__global__ void __launch_bounds__(P_THREADS)
policy_attn_kernel(const PolicyAttnArgs a) {
  extern __shared__ float psm[];
  const int T = a.T, NP = a.NP;
  const int NK = T + 65 + NP;
  const int NKP = NK + 1;             // padded row stride of the score matrix
  float* sq = psm;                    // [T][112]
  float* sc = psm + T * PHD;          // [T][NKP]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x;

  // ---- queries (one thread per row; T <= 32)
  if (tid < T) {
    float f[PHD];
    load_row112(a.qkv_self + (static_cast<long long>(b) * T + tid) * (3 * PD) + h * PHD, f);
    if (a.pro) rope112(f, a.rope_cos, a.rope_sin, tid);
#pragma unroll
    for (int d = 0; d < PHD; ++d) sq[tid * PHD + d] = f[d];
  }
  __syncthreads();

  // ---- scores: one thread per key
  const float inv_sqrt = rsqrtf(static_cast<float>(PHD));
  for (int j = tid; j < NK; j += P_THREADS) {
    const __nv_bfloat16* kp;
    int pos;
    float scale = inv_sqrt;
    if (j < T) {
      kp = a.qkv_self + (static_cast<long long>(b) * T + j) * (3 * PD) + PD + h * PHD;
      pos = j;
    } else if (j < T + 64) {
      kp = a.kv_a + (static_cast<long long>(b) * 64 + (j - T)) * PKV + h * PHD;
      pos = j - T;
    } else if (j == T + 64) {
      kp = a.kv_p + static_cast<long long>(b) * a.ld_p + h * PHD;
      pos = 64;
    } else {
      kp = a.kv_t + (static_cast<long long>(b) * NP + (j - T - 65)) * PKV + h * PHD;
      pos = j - T - 65;
      scale *= a.gate;
    }
    float f[PHD];
    load_row112(kp, f);
    if (a.pro) rope112(f, a.rope_cos, a.rope_sin, pos);
    for (int t = 0; t < T; ++t) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < PHD; ++d) acc += sq[t * PHD + d] * f[d];
      sc[t * NKP + j] = acc * scale;
    }
  }
  __syncthreads();

  // ---- softmax over all NK keys, one warp per query row
  const int warp = tid >> 5, lane = tid & 31;
  for (int t = warp; t < T; t += P_THREADS / 32) {
    float* row = sc + t * NKP;
    float m = -INFINITY;
    for (int j = lane; j < NK; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < NK; j += 32) {
      const float e = __expf(row[j] - m);
      row[j] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int j = lane; j < NK; j += 32) row[j] *= inv;
  }
  __syncthreads();

  // ---- out[t][d] = sum_j P[t][j] V[j][d]; thread = (d, half of the query rows)
  const int d = tid % PHD;
  const int tg = tid / PHD;  // 0, 1 active; 2 idle (256 = 2*112 + 32)
  if (tg < 2) {
    const int t_begin = tg * ((T + 1) / 2);
    const int t_end = (tg == 0) ? (T + 1) / 2 : T;
    float acc[(P_MAXT + 1) / 2];
#pragma unroll
    for (int i = 0; i < (P_MAXT + 1) / 2; ++i) acc[i] = 0.f;
    for (int j = 0; j < NK; ++j) {
      const __nv_bfloat16* vp;
      if (j < T) vp = a.qkv_self + (static_cast<long long>(b) * T + j) * (3 * PD) + 2 * PD;
      else if (j < T + 64) vp = a.kv_a + (static_cast<long long>(b) * 64 + (j - T)) * PKV + PD;
      else if (j == T + 64) vp = a.kv_p + static_cast<long long>(b) * a.ld_p + PD;
      else vp = a.kv_t + (static_cast<long long>(b) * NP + (j - T - 65)) * PKV + PD;
      const float v = __bfloat162float(vp[h * PHD + d]);
#pragma unroll
      for (int i = 0; i < (P_MAXT + 1) / 2; ++i) {
        const int t = t_begin + i;
        if (t < t_end) acc[i] += sc[t * NKP + j] * v;
      }
    }
#pragma unroll
    for (int i = 0; i < (P_MAXT + 1) / 2; ++i) {
      const int t = t_begin + i;
      if (t < t_end) a.out[(static_cast<long long>(b) * T + t) * PD + h * PHD + d] = __float2bfloat16_rn(acc[i]);
    }
  }
}

__global__ void policy_rope_table_kernel(float* cos_t, float* sin_t, int max_pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max_pos * PHD) return;
  const int pos = i / PHD, j = i % PHD;
  // RotaryPositionEmbedding (action_heads.py:157-164): inv_freq = 1/10000^(2i/112), emb = cat(freqs, freqs)
  const float inv_freq = static_cast<float>(1.0 / pow(10000.0, (2.0 * (j % (PHD / 2))) / PHD));
  const float ang = static_cast<float>(pos) * inv_freq;
  cos_t[i] = bf16_round(static_cast<float>(cos(static_cast<double>(ang))));
  sin_t[i] = bf16_round(static_cast<float>(sin(static_cast<double>(ang))));
}

// One warp per row: LayerNorm(896) -> bf16 -> fc2 (A outputs) -> bf16 -> fp32 normalised action,
// plus the un-normalisation of modeling_prismatic.py:799-803 in fp32.
__global__ void __launch_bounds__(256)
head_out_kernel(const __nv_bfloat16* __restrict__ x, int rows, const float* __restrict__ ln_w,
                const float* __restrict__ ln_b, const __nv_bfloat16* __restrict__ W,
                const float* __restrict__ bias, int A, const float* __restrict__ hi,
                const float* __restrict__ lo, const uint8_t* __restrict__ mask, float* __restrict__ out_norm,
                float* __restrict__ out_unnorm) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const __nv_bfloat16* xr = x + static_cast<long long>(row) * PD;
  constexpr int NV = PD / 8;  // 112 vectors; lane handles v = lane, lane+32, lane+64, lane+96(<112)
  float v[4][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int vi = lane + i * 32;
    if (vi < NV) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + vi * 8);
      const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
      v[i][0] = a0.x; v[i][1] = a0.y; v[i][2] = a1.x; v[i][3] = a1.y;
      v[i][4] = a2.x; v[i][5] = a2.y; v[i][6] = a3.x; v[i][7] = a3.y;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
  }
  sum = warp_sum(sum);
  const float mean = sum / PD;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (lane + i * 32 < NV) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float dlt = v[i][j] - mean;
        var += dlt * dlt;
      }
    }
  }
  var = warp_sum(var);
  const float rstd = rsqrtf(var / PD + 1e-5f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int vi = lane + i * 32;
    if (vi < NV) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[i][j] = bf16_round((v[i][j] - mean) * rstd * ln_w[vi * 8 + j] + ln_b[vi * 8 + j]);
    }
  }
  for (int n = 0; n < A; ++n) {
    const __nv_bfloat16* wr = W + static_cast<long long>(n) * PD;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int vi = lane + i * 32;
      if (vi < NV) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(wr + vi * 8));
        const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
        acc += v[i][0] * a0.x + v[i][1] * a0.y + v[i][2] * a1.x + v[i][3] * a1.y + v[i][4] * a2.x +
               v[i][5] * a2.y + v[i][6] * a3.x + v[i][7] * a3.y;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const float an = bf16_round(acc + bias[n]);  // the head's output is bf16 (MP:871-872 .float() after)
      out_norm[static_cast<long long>(row) * A + n] = an;
      if (out_unnorm) {
        const float un = mask[n] ? 0.5f * (an + 1.f) * (hi[n] - lo[n] + 1e-8f) + lo[n] : an;
        out_unnorm[static_cast<long long>(row) * A + n] = un;
      }
    }
  }
}

}  // namespace

int policy_attention_launch(const PolicyAttnArgs& a, cudaStream_t s, const char** err) {
  if (a.T <= 0 || a.T > P_MAXT) {
    if (err) *err = "policy attention: chunk_len must be in [1, 32]";
    return -1;
  }
  const int NK = a.T + 65 + a.NP;
  const size_t smem = sizeof(float) * (static_cast<size_t>(a.T) * PHD + static_cast<size_t>(a.T) * (NK + 1));
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    if (cudaFuncSetAttribute(policy_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem)) != cudaSuccess) {
      if (err) *err = "policy attention: shared memory request too large";
      return -4;
    }
    smem_set = smem;
  }
  dim3 grid(PH, a.B);
  policy_attn_kernel<<<grid, P_THREADS, smem, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = cudaGetErrorString(e);
    return -4;
  }
  ops_count_launch();
  return 0;
}

int policy_rope_table_launch(float* cos_t, float* sin_t, int max_pos, cudaStream_t s, const char** err) {
  const int total = max_pos * PHD;
  policy_rope_table_kernel<<<(total + 255) / 256, 256, 0, s>>>(cos_t, sin_t, max_pos);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = cudaGetErrorString(e);
    return -4;
  }
  ops_count_launch();
  return 0;
}

int head_out_launch(const __nv_bfloat16* x, int rows, const float* ln_w, const float* ln_b,
                    const __nv_bfloat16* W, const float* bias, int A, const float* hi, const float* lo,
                    const uint8_t* mask, float* out_norm, float* out_unnorm, cudaStream_t s, const char** err) {
  head_out_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, rows, ln_w, ln_b, W, bias, A, hi, lo, mask, out_norm,
                                                out_unnorm);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = cudaGetErrorString(e);
    return -4;
  }
  ops_count_launch();
  return 0;
}

}  // namespace vla
