// Policy-head specific kernels: the Pro variant's RoPE and the final regression epilogue.
//
// The Bridge-Attention core itself (MLPResNetBlock.forward, prismatic/models/action_heads.py:218-283, and
// MLPResNetBlock_Pro.forward, :337-410) is laid out by the engine as ONE key/value buffer per sample,
//     rows [0, T)          self   : k/v of the T chunk rows x                      (scale 1)
//     rows [T, T+65)       cond   : k/v of 64 ActionQuery rows h_a ++ proprio row p (scale 1)
//     rows [T+65, T+65+NP) vision : k/v of the NP raw rows h_t, K pre-scaled by tanh(gating_factor)
// all produced by the tcgen05 GEMM writing through 3-D views, so that
//     softmax([q k_self^T | q k_cond^T | g q k_vis^T] / sqrt(112)) [v_self; v_cond; v_vis]
// is a single cross-attention call (attention.cu, hd = 112) on tensor cores.
#include "common.cuh"
#include "launch.cuh"
#include "ops.cuh"

namespace vla {

namespace {

constexpr int PHD = 112;    // head dim
constexpr int PD = 896;
constexpr int PKV = 1792;   // K | V row

// apply_rope (action_heads.py:125-146): pairs are (2i, 2i+1); the angle of lane j is pos * inv_freq[j mod 56]
// (concat-style table, action_heads.py:162-163); every op rounds to bf16 like the eager bf16 reference.
// One thread rotates 8 contiguous lanes (4 pairs) of one head of one row.
__global__ void __launch_bounds__(256)
policy_rope_kernel(__nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ kv, int B, int T, int NP,
                   const float* __restrict__ cos_t, const float* __restrict__ sin_t, int kv_only) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, their writes visible
  pdl_launch_dependents();
  const int NK = T + 65 + NP;
  // kv_only: just the cond / vision key rows [T, NK) (the fused small-batch policy kernel rotates q and the self keys itself)
  const int skip = kv_only ? 2 * T : 0;
  const int rows_per_b = T + NK - skip;  // q rows then kv rows
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * rows_per_b * (PD / 8);
  if (idx >= total) return;
  const int chunk = static_cast<int>(idx % (PD / 8));  // 8-lane chunk within the 896-wide row
  const long long t2 = idx / (PD / 8);
  const int r = static_cast<int>(t2 % rows_per_b) + skip;
  const int b = static_cast<int>(t2 / rows_per_b);
  __nv_bfloat16* p;
  int pos;
  if (r < T) {
    p = q + (static_cast<long long>(b) * T + r) * PD;
    pos = r;
  } else {
    const int j = r - T;
    p = kv + (static_cast<long long>(b) * NK + j) * PKV;
    pos = j < T ? j : (j < T + 65 ? j - T : j - T - 65);
  }
  p += chunk * 8;
  const int lane0 = (chunk * 8) % PHD;  // lane within the head (112 = 14 chunks, chunks never straddle heads)
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  const float* c = cos_t + pos * PHD + lane0;
  const float* s = sin_t + pos * PHD + lane0;
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 x = unpack_bf16(w[i]);
    const float y0 = bf16_round(bf16_round(x.x * c[2 * i]) + bf16_round(-x.y * s[2 * i]));
    const float y1 = bf16_round(bf16_round(x.y * c[2 * i + 1]) + bf16_round(x.x * s[2 * i + 1]));
    o[i] = pack_bf16(y0, y1);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void policy_rope_table_kernel(float* cos_t, float* sin_t, int max_pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= max_pos * PHD) return;
  const int pos = i / PHD, j = i % PHD;
  // RotaryPositionEmbedding (action_heads.py:157-164): inv_freq = 1/10000^(2i/112), emb = cat(freqs, freqs).
  // AS DEPLOYED the head is cast with .to(torch.bfloat16) (experiments/robot/openvla_utils.py:515), which also
  // casts the non-persistent inv_freq buffer; t = arange(dtype=inv_freq.dtype) (action_heads.py:161) and the
  // outer product are then bf16 too, so the angle is bf16(bf16(t) * bf16(inv_freq)).  Pinned by
  // tests/golden/libero_pro.npz.
  const float inv_freq = bf16_round(1.0f / powf(10000.0f, static_cast<float>(2 * (j % (PHD / 2))) / PHD));
  const float ang = bf16_round(bf16_round(static_cast<float>(pos)) * inv_freq);
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
  pdl_wait();  // programmatic dependent launch: predecessors complete, their writes visible
  pdl_launch_dependents();
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

inline int finish(const char** err) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = cudaGetErrorString(e);
    return -4;
  }
  ops_count_launch();
  return 0;
}

}  // namespace

int policy_rope_launch(__nv_bfloat16* q, __nv_bfloat16* kv, int B, int T, int NP, const float* cos_t,
                       const float* sin_t, cudaStream_t s, const char** err, int kv_only) {
  const long long total = static_cast<long long>(B) * ((kv_only ? 0 : 2 * T) + 65 + NP) * (PD / 8);
  launch_kernel(policy_rope_kernel, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, s, q, kv, B, T, NP, cos_t, sin_t,
                kv_only);
  return finish(err);
}

int policy_rope_table_launch(float* cos_t, float* sin_t, int max_pos, cudaStream_t s, const char** err) {
  const int total = max_pos * PHD;
  policy_rope_table_kernel<<<(total + 255) / 256, 256, 0, s>>>(cos_t, sin_t, max_pos);
  return finish(err);
}

int head_out_launch(const __nv_bfloat16* x, int rows, const float* ln_w, const float* ln_b,
                    const __nv_bfloat16* W, const float* bias, int A, const float* hi, const float* lo,
                    const uint8_t* mask, float* out_norm, float* out_unnorm, cudaStream_t s, const char** err) {
  launch_kernel(head_out_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, x, rows, ln_w, ln_b, W, bias, A, hi, lo, mask, out_norm,
                                                out_unnorm);
  return finish(err);
}

}  // namespace vla
