// Bandwidth-bound kernels of the predict_action path: norms, RoPE, im2col, token assembly,
// skinny linears.  All are vectorised (16-byte) and coalesced; reductions use warp shuffles.
#include "common.cuh"
#include "launch.cuh"
#include "ops.cuh"

#include <atomic>

namespace vla {

namespace {
std::atomic<long long> g_ops_launches{0};

inline int check_launch(const char** err) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = cudaGetErrorString(e);
    return -4;
  }
  g_ops_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

constexpr int MAX_VEC_PER_LANE = 5;  // dim <= 5*32*8 = 1280

// ------------------------------------------------------------------ LayerNorm / RMSNorm
// One warp per row.  The row is read once into registers (16-byte loads), statistics in fp32.
template <bool RMS>
__global__ void __launch_bounds__(256, 5)
norm_kernel(const __nv_bfloat16* __restrict__ x, int rows, int batches, long long x_bs, int dim, int ldx,
            const float* __restrict__ w, const float* __restrict__ b, float eps,
            __nv_bfloat16* __restrict__ y, long long y_bs, int ldy) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, their writes visible
  pdl_launch_dependents();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(rows) * batches;
  if (warp >= total) return;
  const int bi = warp / rows;
  const int r = warp - bi * rows;
  const __nv_bfloat16* xr = x + bi * x_bs + static_cast<long long>(r) * ldx;
  __nv_bfloat16* yr = y + bi * y_bs + static_cast<long long>(r) * ldy;
  const int nvec = dim >> 3;

  // The row stays packed (bf16) in registers - 4 registers per 16-byte vector instead of 8 floats - so that the
  // kernel fits 2048 threads per SM and keeps enough loads in flight to approach the HBM roofline.
  uint4 u[MAX_VEC_PER_LANE];
#pragma unroll
  for (int i = 0; i < MAX_VEC_PER_LANE; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) u[i] = *reinterpret_cast<const uint4*>(xr + vi * 8);
  }
  // One pass, one round of shuffles: RMSNorm needs sum x^2; LayerNorm takes mean and variance from the sums of
  // (x - p) and (x - p)^2 around a pivot p (the row's first element), which stays well conditioned when |mean| >> std.
  const float pivot = RMS ? 0.f : __shfl_sync(0xffffffffu, unpack_bf16(u[0].x).x, 0);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC_PER_LANE; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint32_t w4[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16(w4[j]);
        const float d0 = a.x - pivot, d1 = a.y - pivot;
        if (!RMS) s1 += d0 + d1;
        s2 += d0 * d0 + d1 * d1;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    if (!RMS) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  float mean = 0.f, rstd;
  if (RMS) {
    rstd = rsqrtf(s2 / dim + eps);
  } else {
    const float m1 = s1 / dim;
    mean = pivot + m1;
    rstd = rsqrtf(fmaxf(s2 / dim - m1 * m1, 0.f) + eps);
  }
#pragma unroll
  for (int i = 0; i < MAX_VEC_PER_LANE; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + vi * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + vi * 8 + 4));
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const uint32_t w4[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
      float v[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16(w4[j]);
        v[2 * j] = a.x;
        v[2 * j + 1] = a.y;
      }
      float o[8];
      if (RMS) {
        // HF: weight * (x_fp32 * rsqrt(var + eps)).to(bf16)
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = ww[j] * bf16_round(v[j] * rstd);
      } else {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + vi * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(b + vi * 8 + 4));
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * ww[j] + bb[j];
      }
      *reinterpret_cast<uint4*>(yr + vi * 8) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                                                          pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
    }
  }
}

// ------------------------------------------------------------------ row statistics of a norm folded into the next GEMM
// (rstd, -mean * rstd) per row; the GEMM that consumes the raw rows applies them in its epilogue (gemm.cuh: row_stats).
// Read-only: half the traffic of norm_kernel, and nothing of the row has to stay in registers - each warp takes two
// rows at a time to keep more loads in flight.
template <bool RMS>
__global__ void __launch_bounds__(256, 4)
row_stats_kernel(const __nv_bfloat16* __restrict__ x, int rows, int dim, int ldx, float eps, float2* __restrict__ stats,
                 int partial_slots) {
  pdl_wait();
  pdl_launch_dependents();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int r0 = warp * 2;
  if (r0 >= rows) return;
  const bool two = r0 + 1 < rows;
  const __nv_bfloat16* xa = x + static_cast<long long>(r0) * ldx;
  const __nv_bfloat16* xb = xa + (two ? ldx : 0);
  const int nvec = dim >> 3;
  uint4 ua[MAX_VEC_PER_LANE], ub[MAX_VEC_PER_LANE];
#pragma unroll
  for (int i = 0; i < MAX_VEC_PER_LANE; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      ua[i] = *reinterpret_cast<const uint4*>(xa + vi * 8);
      ub[i] = *reinterpret_cast<const uint4*>(xb + vi * 8);
    }
  }
  // same single-pass pivoted sums as norm_kernel
  const float pa = RMS ? 0.f : __shfl_sync(0xffffffffu, unpack_bf16(ua[0].x).x, 0);
  const float pb = RMS ? 0.f : __shfl_sync(0xffffffffu, unpack_bf16(ub[0].x).x, 0);
  float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC_PER_LANE; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint32_t wa[4] = {ua[i].x, ua[i].y, ua[i].z, ua[i].w};
      const uint32_t wb[4] = {ub[i].x, ub[i].y, ub[i].z, ub[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fa = unpack_bf16(wa[j]), fb = unpack_bf16(wb[j]);
        const float d0 = fa.x - pa, d1 = fa.y - pa, e0 = fb.x - pb, e1 = fb.y - pb;
        if (!RMS) {
          a1 += d0 + d1;
          b1 += e0 + e1;
        }
        a2 += d0 * d0 + d1 * d1;
        b2 += e0 * e0 + e1 * e1;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    if (!RMS) {
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      b1 += __shfl_xor_sync(0xffffffffu, b1, o);
    }
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    b2 += __shfl_xor_sync(0xffffffffu, b2, o);
  }
  if (lane < 2 && (lane == 0 || two)) {
    const float s1 = lane ? b1 : a1, s2 = lane ? b2 : a2, pv = lane ? pb : pa;
    if (partial_slots > 0) {
      // the producer-GEMM format (gemm.cuh: STAT_SLOTS partial (sum x, sum x^2) pairs per row): everything in slot 0
      float2* sp = stats + static_cast<long long>(r0 + lane) * partial_slots;
      sp[0] = make_float2(s1 + dim * pv, s2 + 2.f * pv * s1 + dim * pv * pv);
      for (int k = 1; k < partial_slots; ++k) sp[k] = make_float2(0.f, 0.f);
      return;
    }
    float2 o;
    if (RMS) {
      o = make_float2(rsqrtf(s2 / dim + eps), 0.f);
    } else {
      const float m1 = s1 / dim;
      const float rstd = rsqrtf(fmaxf(s2 / dim - m1 * m1, 0.f) + eps);
      o = make_float2(rstd, -(pv + m1) * rstd);
    }
    stats[r0 + lane] = o;
  }
}

// One-time weight fold (engine finalize): W[n, k] <- bf16(W[n, k] * g[k]); bias[n] += sum_k W[n, k] * b[k];
// colsum[n] = sum_k W'[n, k] (of the ROUNDED products - what the tensor cores will multiply).  One warp per row n.
__global__ void fold_norm_kernel(__nv_bfloat16* __restrict__ W, int N, int K, int ldw, const float* __restrict__ g,
                                 const float* __restrict__ b, float* __restrict__ bias, float* __restrict__ colsum) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  __nv_bfloat16* w = W + static_cast<long long>(n) * ldw;
  float su = 0.f, sv = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float wf = __bfloat162float(w[k]);
    if (b) sv += wf * b[k];
    const __nv_bfloat16 wr = __float2bfloat16_rn(wf * g[k]);
    w[k] = wr;
    su += __bfloat162float(wr);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    su += __shfl_xor_sync(0xffffffffu, su, o);
    sv += __shfl_xor_sync(0xffffffffu, sv, o);
  }
  if (lane == 0) {
    if (colsum) colsum[n] = su;
    if (bias && b) bias[n] += sv;
  }
}

// [S][32] fp32 cos / sin tables (bf16 values) -> [32][S] words (bf16 cos | bf16 sin << 16) for the GEMM's RoPE epilogue
__global__ void rope_pack_kernel(const float* __restrict__ cos_t, const float* __restrict__ sin_t, int S,
                                 uint32_t* __restrict__ cs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * 32) return;
  const int j = i / S, pos = i - j * S;
  const uint32_t c = __float_as_uint(bf16_round(cos_t[pos * 32 + j])) >> 16;
  const uint32_t sn = __float_as_uint(bf16_round(sin_t[pos * 32 + j])) & 0xffff0000u;
  cs[i] = c | sn;
}

// ------------------------------------------------------------------ RoPE (HF Qwen2, rotate_half)
__global__ void rope_table_kernel(float* cos_t, float* sin_t, int S, int half, float theta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * half) return;
  const int pos = i / half, j = i - pos * half;
  // transformers Qwen2RotaryEmbedding: inv_freq fp32, angle = pos * inv_freq in fp32, cos/sin -> bf16
  const float inv_freq = static_cast<float>(1.0 / pow(static_cast<double>(theta), (2.0 * j) / (2.0 * half)));
  const float ang = static_cast<float>(pos) * inv_freq;
  cos_t[i] = bf16_round(static_cast<float>(cos(static_cast<double>(ang))));
  sin_t[i] = bf16_round(static_cast<float>(sin(static_cast<double>(ang))));
}

// One thread handles 8 contiguous dims j..j+7 (j < 32) of one head and their partners j+32.
__global__ void __launch_bounds__(256)
rope_apply_kernel(__nv_bfloat16* __restrict__ x, int ld, int off, int n_heads, long long rows, int S,
                  const float* __restrict__ cos_t, const float* __restrict__ sin_t) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, their writes visible
  pdl_launch_dependents();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = rows * n_heads * 4;
  if (idx >= total) return;
  const int part = static_cast<int>(idx & 3);
  const long long t = idx >> 2;
  const int h = static_cast<int>(t % n_heads);
  const long long row = t / n_heads;
  const int pos = static_cast<int>(row % S);
  __nv_bfloat16* p = x + row * ld + off + h * 64 + part * 8;
  const uint4 lo = *reinterpret_cast<const uint4*>(p);
  const uint4 hi = *reinterpret_cast<const uint4*>(p + 32);
  const uint32_t lw[4] = {lo.x, lo.y, lo.z, lo.w}, hw[4] = {hi.x, hi.y, hi.z, hi.w};
  float a[8], bq[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = unpack_bf16(lw[i]), g = unpack_bf16(hw[i]);
    a[2 * i] = f.x; a[2 * i + 1] = f.y; bq[2 * i] = g.x; bq[2 * i + 1] = g.y;
  }
  const float* c = cos_t + pos * 32 + part * 8;
  const float* s = sin_t + pos * 32 + part * 8;
  float o1[8], o2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float cs = c[i], sn = s[i];
    // q_embed = (q * cos) + (rotate_half(q) * sin), each op rounded to bf16 like the eager reference
    o1[i] = bf16_round(bf16_round(a[i] * cs) + bf16_round(-bq[i] * sn));
    o2[i] = bf16_round(bf16_round(bq[i] * cs) + bf16_round(a[i] * sn));
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(o1[0], o1[1]), pack_bf16(o1[2], o1[3]),
                                            pack_bf16(o1[4], o1[5]), pack_bf16(o1[6], o1[7]));
  *reinterpret_cast<uint4*>(p + 32) = make_uint4(pack_bf16(o2[0], o2[1]), pack_bf16(o2[2], o2[3]),
                                                 pack_bf16(o2[4], o2[5]), pack_bf16(o2[6], o2[7]));
}

// ------------------------------------------------------------------ im2col straight from uint8 frames
// Device-side image front-end: frames already resized / centre-cropped to 224x224 arrive as uint8 HWC
// (PIL layout); the processor's ToTensor + Normalize of this tower (processing_prismatic.py:128-145, means / stds of
// preprocessor_config.json) and the cast to bf16 are a 3 x 256 lookup table built on the host with the processor's
// own fp32 arithmetic, so the result is bit-identical to the CPU path while the H2D copy shrinks 4x.
__global__ void __launch_bounds__(256)
im2col_u8_kernel(const uint8_t* __restrict__ img, int n_img, int tower, long long n_slabs,
                 const __nv_bfloat16* __restrict__ lut /*[2][3][256]*/, __nv_bfloat16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int KP = 592;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = n_slabs * 256 * 43;  // 42 (c,ky) segments + 1 zero-pad segment
  if (idx >= total) return;
  const int seg = static_cast<int>(idx % 43);
  const long long pr = idx / 43;  // slab*256 + patch
  const int patch = static_cast<int>(pr & 255);
  const long long slab = pr >> 8;
  __nv_bfloat16* orow = out + pr * KP;
  if (seg == 42) {
    *reinterpret_cast<uint2*>(orow + 588) = make_uint2(0u, 0u);
    return;
  }
  const int c = seg / 14, ky = seg - c * 14;
  const int py = patch >> 4, px = patch & 15;
  const uint8_t* src = img + ((slab * 224 + (py * 14 + ky)) * 224LL + px * 14) * 3 + c;  // HWC
  const __nv_bfloat16* t = lut + (tower * 3 + c) * 256;
  __nv_bfloat16* dst = orow + c * 196 + ky * 14;
#pragma unroll
  for (int i = 0; i < 14; ++i) dst[i] = t[__ldg(src + i * 3)];
}

// ------------------------------------------------------------------ centre crop (device-side image front-end)
// center_crop_image of the reference (experiments/robot/openvla_utils.py:616-648 -> crop_and_resize :568-613): the
// centred box of relative side sqrt(crop_scale), resampled bilinearly to out x out with TensorFlow's crop_and_resize
// arithmetic, uint8 -> [0,1] float -> uint8 (x 255.5, truncated).  Every product and sum is rounded separately
// (__fmul_rn / __fadd_rn: no FMA contraction), in the order of oracle/image_prep.py, so the two agree bit for bit.
__global__ void __launch_bounds__(256)
center_crop_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long n_images, int H, int W,
                      int osz, float y0, float x0, float hs, float ws) {
  pdl_wait();
  pdl_launch_dependents();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_images * osz * osz) return;
  const int ox = static_cast<int>(idx % osz);
  const int oy = static_cast<int>((idx / osz) % osz);
  const long long im = idx / (static_cast<long long>(osz) * osz);
  const float in_y = __fadd_rn(y0, __fmul_rn(static_cast<float>(oy), hs));
  const float in_x = __fadd_rn(x0, __fmul_rn(static_cast<float>(ox), ws));
  const float fy = floorf(in_y), fx = floorf(in_x);
  const int top = static_cast<int>(fy), bot = static_cast<int>(ceilf(in_y));
  const int left = static_cast<int>(fx), right = static_cast<int>(ceilf(in_x));
  const float yl = __fsub_rn(in_y, fy), xl = __fsub_rn(in_x, fx);
  const uint8_t* base = in + im * H * W * 3LL;
  const float k = static_cast<float>(1.0 / 255.0);
  uint8_t* o = out + idx * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float tl = __fmul_rn(static_cast<float>(base[(static_cast<long long>(top) * W + left) * 3 + c]), k);
    const float tr = __fmul_rn(static_cast<float>(base[(static_cast<long long>(top) * W + right) * 3 + c]), k);
    const float bl = __fmul_rn(static_cast<float>(base[(static_cast<long long>(bot) * W + left) * 3 + c]), k);
    const float br = __fmul_rn(static_cast<float>(base[(static_cast<long long>(bot) * W + right) * 3 + c]), k);
    const float t = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl));
    const float b = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
    float v = __fadd_rn(t, __fmul_rn(__fsub_rn(b, t), yl));
    v = fminf(fmaxf(v, 0.f), 1.f);
    int q = __float2int_rz(__fmul_rn(v, 255.5f));
    q = q < 0 ? 0 : (q > 255 ? 255 : q);
    o[c] = static_cast<uint8_t>(q);
  }
}

// ------------------------------------------------------------------ im2col for the 14x14/14 patch conv
// One thread per (patch row vector of 14 pixels): reads 14 contiguous bf16 (28 B) of the image, writes
// them at k = c*196 + ky*14 + [0,14).  Grid: (image slab, patch) x (c, ky).
__global__ void __launch_bounds__(256)
im2col_kernel(const __nv_bfloat16* __restrict__ pix, int n_img, int tower, long long n_slabs,
              __nv_bfloat16* __restrict__ out) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, their writes visible
  pdl_launch_dependents();
  constexpr int KP = 592;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = n_slabs * 256 * 43;  // 42 (c,ky) segments + 1 zero-pad segment
  if (idx >= total) return;
  const int seg = static_cast<int>(idx % 43);
  const long long pr = idx / 43;  // slab*256 + patch
  const int patch = static_cast<int>(pr & 255);
  const long long slab = pr >> 8;
  __nv_bfloat16* orow = out + pr * KP;
  if (seg == 42) {
    *reinterpret_cast<uint2*>(orow + 588) = make_uint2(0u, 0u);
    return;
  }
  const int c = seg / 14, ky = seg - c * 14;
  const int py = patch >> 4, px = patch & 15;
  const long long b = slab / n_img;
  const int img = static_cast<int>(slab - b * n_img);
  const int ch = img * 6 + tower * 3 + c;
  const __nv_bfloat16* src =
      pix + ((b * (6 * n_img) + ch) * 224 + (py * 14 + ky)) * 224LL + px * 14;  // 4-byte aligned
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d32 = reinterpret_cast<uint32_t*>(orow + c * 196 + ky * 14);  // 4-byte aligned
#pragma unroll
  for (int i = 0; i < 7; ++i) d32[i] = __ldg(s32 + i);
}

__global__ void prefix_tokens_kernel(__nv_bfloat16* __restrict__ x, int n_slabs, long long slab_stride,
                                     int dim, const __nv_bfloat16* __restrict__ prefix, int n_prefix) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, their writes visible
  pdl_launch_dependents();
  const int nvec = dim >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(n_slabs) * n_prefix * nvec;
  if (idx >= total) return;
  const int v = static_cast<int>(idx % nvec);
  const long long t = idx / nvec;
  const int r = static_cast<int>(t % n_prefix);
  const long long slab = t / n_prefix;
  const uint4 val = __ldg(reinterpret_cast<const uint4*>(prefix + static_cast<long long>(r) * dim + v * 8));
  *reinterpret_cast<uint4*>(x + slab * slab_stride + static_cast<long long>(r) * dim + v * 8) = val;
}

// ------------------------------------------------------------------ LLM input assembly
__global__ void __launch_bounds__(128)
assemble_kernel(__nv_bfloat16* __restrict__ x, int B, int S, int NP, int Lext, int dim,
                const int64_t* __restrict__ ext_ids, const int32_t* __restrict__ aq_index,
                const __nv_bfloat16* __restrict__ embed, int vocab, const __nv_bfloat16* __restrict__ aq_table,
                int n_aq, int* err_flag) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, their writes visible
  pdl_launch_dependents();
  // one block per (sample, text column j in [0, Lext))
  const int j = blockIdx.x % Lext;
  const int b = blockIdx.x / Lext;
  const int s = (j == 0) ? 0 : NP + j;
  const int aq = aq_index[static_cast<long long>(b) * Lext + j];
  // An out-of-range id / index raises the device flag (vla_check_errors, vla_predict_host) and the row is ZERO-filled:
  // the call's result is then deterministic garbage for that sample, never a stale row of an earlier call.
  const __nv_bfloat16* src = nullptr;
  if (aq >= 0) {
    if (aq >= n_aq) {
      if (threadIdx.x == 0) atomicExch(err_flag, 2);
    } else {
      src = aq_table + static_cast<long long>(aq) * dim;
    }
  } else {
    const int64_t id = ext_ids[static_cast<long long>(b) * Lext + j];
    if (id < 0 || id >= vocab) {
      if (threadIdx.x == 0) atomicExch(err_flag, 1);
    } else {
      src = embed + id * dim;
    }
  }
  __nv_bfloat16* dst = x + (static_cast<long long>(b) * S + s) * dim;
  for (int v = threadIdx.x; v < (dim >> 3); v += blockDim.x)
    reinterpret_cast<uint4*>(dst)[v] = src ? __ldg(reinterpret_cast<const uint4*>(src) + v) : make_uint4(0u, 0u, 0u, 0u);
}

// ------------------------------------------------------------------ skinny linear
__global__ void __launch_bounds__(256)
skinny_linear_kernel(const void* __restrict__ x, int x_is_f32, int ldx, int M, int K,
                     const __nv_bfloat16* __restrict__ W, int ldw, int N, const float* __restrict__ bias,
                     int act, __nv_bfloat16* __restrict__ out, int ldo, float* __restrict__ out_f32) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, their writes visible
  pdl_launch_dependents();
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= static_cast<long long>(M) * N) return;
  const int n = static_cast<int>(warp % N);
  const int m = static_cast<int>(warp / N);
  const __nv_bfloat16* wr = W + static_cast<long long>(n) * ldw;
  float acc = 0.f;
  if (x_is_f32) {
    const float* xr = static_cast<const float*>(x) + static_cast<long long>(m) * ldx;
    for (int k = lane; k < K; k += 32) acc += bf16_round(xr[k]) * __bfloat162float(wr[k]);
  } else {
    const __nv_bfloat16* xr = static_cast<const __nv_bfloat16*>(x) + static_cast<long long>(m) * ldx;
    if ((K & 7) == 0 && (ldx & 7) == 0 && (ldw & 7) == 0) {
      for (int v = lane; v < (K >> 3); v += 32) {
        const uint4 a = *reinterpret_cast<const uint4*>(xr + v * 8);
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(wr + v * 8));
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 fa = unpack_bf16(aw[i]), fw = unpack_bf16(ww[i]);
          acc += fa.x * fw.x + fa.y * fw.y;
        }
      }
    } else {
      for (int k = lane; k < K; k += 32) acc += __bfloat162float(xr[k]) * __bfloat162float(wr[k]);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (bias) acc += bias[n];
    if (act == 1) acc = gelu_erf(acc);
    else if (act == 2) acc = fmaxf(acc, 0.f);
    const __nv_bfloat16 o = __float2bfloat16_rn(acc);
    if (out) out[static_cast<long long>(m) * ldo + n] = o;
    if (out_f32) out_f32[static_cast<long long>(m) * N + n] = __bfloat162float(o);
  }
}

// dst[b, r, :] = src[b, r0 + (len ? len[b] - len_ref : 0) + r, :]: rows [r0, r0 + rows) of every slab, shifted per
// slab when the samples' prompts have different lengths (len_ref = the length r0 was computed for; a length outside
// [1, len_ref] raises error flag 3 and reads as len_ref).
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, long long src_bs, int ld, int r0,
                                   int rows, int batches, int dim, __nv_bfloat16* __restrict__ dst,
                                   const int32_t* __restrict__ len, int len_ref, int* err_flag) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, their writes visible
  pdl_launch_dependents();
  const int nvec = dim >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(batches) * rows * nvec;
  if (idx >= total) return;
  const int v = static_cast<int>(idx % nvec);
  const long long t = idx / nvec;
  const int r = static_cast<int>(t % rows);
  const long long b = t / rows;
  int shift = 0;
  if (len) {
    const int l = len[b];
    if (l < 1 || l > len_ref) {
      if (err_flag && v == 0 && r == 0) atomicExch(err_flag, 3);
    } else {
      shift = l - len_ref;
    }
  }
  reinterpret_cast<uint4*>(dst)[idx] =
      *reinterpret_cast<const uint4*>(src + b * src_bs + static_cast<long long>(r0 + shift + r) * ld + v * 8);
}

__global__ void broadcast_row_kernel(const __nv_bfloat16* __restrict__ src, int dim, int rows,
                                     __nv_bfloat16* __restrict__ dst) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, their writes visible
  pdl_launch_dependents();
  const int nvec = dim >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(rows) * nvec) return;
  reinterpret_cast<uint4*>(dst)[idx] = __ldg(reinterpret_cast<const uint4*>(src) + (idx % nvec));
}

}  // namespace

long long ops_launch_count() { return g_ops_launches.load(); }
void ops_count_launch(int n) { g_ops_launches.fetch_add(n, std::memory_order_relaxed); }

int layernorm_launch_3d(const __nv_bfloat16* x, int rows, int batches, long long x_bs, int dim, int ldx,
                        const float* w, const float* b, float eps, __nv_bfloat16* y, long long y_bs, int ldy,
                        cudaStream_t s, const char** err) {
  if ((dim & 7) || dim > MAX_VEC_PER_LANE * 256 || (ldx & 7) || (ldy & 7)) {
    if (err) *err = "layernorm: dim must be a multiple of 8 and <= 1280";
    return -1;
  }
  const long long total = static_cast<long long>(rows) * batches;
  const int blocks = static_cast<int>((total + 7) / 8);
  launch_kernel(norm_kernel<false>, dim3(blocks), dim3(256), 0, s, x, rows, batches, x_bs, dim, ldx, w, b, eps, y, y_bs, ldy);
  return check_launch(err);
}

int layernorm_launch(const __nv_bfloat16* x, int rows, int dim, int ldx, const float* w, const float* b,
                     float eps, __nv_bfloat16* y, int ldy, cudaStream_t s, const char** err) {
  return layernorm_launch_3d(x, rows, 1, 0, dim, ldx, w, b, eps, y, 0, ldy, s, err);
}

int rmsnorm_launch(const __nv_bfloat16* x, int rows, int dim, int ldx, const float* w, float eps,
                   __nv_bfloat16* y, int ldy, cudaStream_t s, const char** err) {
  if ((dim & 7) || dim > MAX_VEC_PER_LANE * 256 || (ldx & 7) || (ldy & 7)) {
    if (err) *err = "rmsnorm: dim must be a multiple of 8 and <= 1280";
    return -1;
  }
  const int blocks = (rows + 7) / 8;
  launch_kernel(norm_kernel<true>, dim3(blocks), dim3(256), 0, s, x, rows, 1, 0, dim, ldx, w, nullptr, eps, y, 0, ldy);
  return check_launch(err);
}

int row_stats_launch(const __nv_bfloat16* x, int rows, int dim, int ldx, int rms, float eps, float* stats,
                     cudaStream_t s, const char** err, int partial_slots) {
  if ((dim & 7) || dim > MAX_VEC_PER_LANE * 256 || (ldx & 7) || (reinterpret_cast<uintptr_t>(stats) & 7)) {
    if (err) *err = "row_stats: dim must be a multiple of 8 and <= 1280";
    return -1;
  }
  const int blocks = (rows + 15) / 16;  // 8 warps x 2 rows
  if (rms) launch_kernel(row_stats_kernel<true>, dim3(blocks), dim3(256), 0, s, x, rows, dim, ldx, eps, reinterpret_cast<float2*>(stats), partial_slots);
  else launch_kernel(row_stats_kernel<false>, dim3(blocks), dim3(256), 0, s, x, rows, dim, ldx, eps, reinterpret_cast<float2*>(stats), partial_slots);
  return check_launch(err);
}

int fold_norm_launch(__nv_bfloat16* W, int N, int K, int ldw, const float* g, const float* b, float* bias,
                     float* colsum, cudaStream_t s, const char** err) {
  if (!W || !g || (b && !bias)) {
    if (err) *err = "fold_norm: a norm bias needs a GEMM bias to fold into";
    return -1;
  }
  fold_norm_kernel<<<(N + 7) / 8, 256, 0, s>>>(W, N, K, ldw, g, b, bias, colsum);
  return check_launch(err);
}

int rope_pack_launch(const float* cos_t, const float* sin_t, int S, uint32_t* cs, cudaStream_t s, const char** err) {
  rope_pack_kernel<<<(S * 32 + 255) / 256, 256, 0, s>>>(cos_t, sin_t, S, cs);
  return check_launch(err);
}

int rope_table_launch(float* cos_t, float* sin_t, int S, int half, float theta, cudaStream_t s,
                      const char** err) {
  const int total = S * half;
  rope_table_kernel<<<(total + 255) / 256, 256, 0, s>>>(cos_t, sin_t, S, half, theta);
  return check_launch(err);
}

int rope_apply_launch(__nv_bfloat16* x, int ld, int off, int n_heads, int B, int S, const float* cos_t,
                      const float* sin_t, cudaStream_t s, const char** err) {
  if ((ld & 7) || (off & 7)) {
    if (err) *err = "rope: ld/off must be multiples of 8";
    return -1;
  }
  const long long rows = static_cast<long long>(B) * S;
  const long long total = rows * n_heads * 4;
  launch_kernel(rope_apply_kernel, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, s, x, ld, off, n_heads, rows, S, cos_t,
                                                                          sin_t);
  return check_launch(err);
}

int rope_launch(__nv_bfloat16* x, int ld, int off, int n_heads, int B, int S, float theta, cudaStream_t s,
                const char** err) {
  float* tab = nullptr;
  if (cudaMallocAsync(&tab, sizeof(float) * 2 * S * 32, s) != cudaSuccess) {
    if (err) *err = "rope: cudaMallocAsync failed";
    return -4;
  }
  int rc = rope_table_launch(tab, tab + S * 32, S, 32, theta, s, err);
  if (!rc) rc = rope_apply_launch(x, ld, off, n_heads, B, S, tab, tab + S * 32, s, err);
  cudaFreeAsync(tab, s);
  return rc;
}

int im2col_u8_launch(const uint8_t* img, int B, int n_img, int tower, const __nv_bfloat16* lut, __nv_bfloat16* out,
                     cudaStream_t s, const char** err) {
  const long long n_slabs = static_cast<long long>(B) * n_img;
  const long long total = n_slabs * 256 * 43;
  launch_kernel(im2col_u8_kernel, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, s, img, n_img, tower, n_slabs,
                lut, out);
  return check_launch(err);
}

int center_crop_u8_launch(const uint8_t* in, uint8_t* out, long long n_images, int H, int W, int out_size,
                          float crop_scale, cudaStream_t s, const char** err) {
  if (n_images <= 0 || H < 2 || W < 2 || out_size < 2 || !(crop_scale > 0.f) || crop_scale > 1.f) {
    if (err) *err = "center_crop: bad shape or crop_scale outside (0, 1]";
    return -1;
  }
  // the box of openvla_utils.py:588-603, every operation rounded to fp32 in the order of oracle/image_prep.py
  volatile float side = sqrtf(crop_scale);
  if (side > 1.f) side = 1.f;
  volatile float off = (1.f - side) / 2.f;
  volatile float y2 = off + side;
  volatile float dh = (y2 - off) * static_cast<float>(H - 1);
  volatile float hs = dh / static_cast<float>(out_size - 1);
  volatile float dw = (y2 - off) * static_cast<float>(W - 1);
  volatile float ws = dw / static_cast<float>(out_size - 1);
  volatile float y0 = off * static_cast<float>(H - 1);
  volatile float x0 = off * static_cast<float>(W - 1);
  const long long total = n_images * out_size * out_size;
  launch_kernel(center_crop_u8_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, in, out, n_images, H, W,
                out_size, static_cast<float>(y0), static_cast<float>(x0), static_cast<float>(hs), static_cast<float>(ws));
  return check_launch(err);
}

int im2col_launch(const __nv_bfloat16* pix, int B, int n_img, int tower, __nv_bfloat16* out, cudaStream_t s,
                  const char** err) {
  const long long n_slabs = static_cast<long long>(B) * n_img;
  const long long total = n_slabs * 256 * 43;
  launch_kernel(im2col_kernel, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, s, pix, n_img, tower, n_slabs, out);
  return check_launch(err);
}

int prefix_tokens_launch(__nv_bfloat16* x, int n_slabs, long long slab_stride, int dim,
                         const __nv_bfloat16* prefix, int n_prefix, cudaStream_t s, const char** err) {
  const long long total = static_cast<long long>(n_slabs) * n_prefix * (dim >> 3);
  launch_kernel(prefix_tokens_kernel, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, s, x, n_slabs, slab_stride, dim,
                                                                            prefix, n_prefix);
  return check_launch(err);
}

int assemble_launch(__nv_bfloat16* x, int B, int S, int NP, int Lext, int dim, const int64_t* ext_ids,
                    const int32_t* aq_index, const __nv_bfloat16* embed, int vocab,
                    const __nv_bfloat16* aq_table, int n_aq, int* err_flag, cudaStream_t s, const char** err) {
  launch_kernel(assemble_kernel, dim3(B * Lext), dim3(128), 0, s, x, B, S, NP, Lext, dim, ext_ids, aq_index, embed, vocab, aq_table,
                                           n_aq, err_flag);
  return check_launch(err);
}

int skinny_linear_launch(const void* x, int x_is_f32, int ldx, int M, int K, const __nv_bfloat16* W, int ldw,
                         int N, const float* bias, int act, __nv_bfloat16* out, int ldo, float* out_f32,
                         cudaStream_t s, const char** err) {
  const long long warps = static_cast<long long>(M) * N;
  launch_kernel(skinny_linear_kernel, dim3(static_cast<int>((warps + 7) / 8)), dim3(256), 0, s, x, x_is_f32, ldx, M, K, W, ldw, N,
                                                                        bias, act, out, ldo, out_f32);
  return check_launch(err);
}

int broadcast_row_launch(const __nv_bfloat16* src, int dim, int rows, __nv_bfloat16* dst, cudaStream_t s,
                         const char** err) {
  const long long total = static_cast<long long>(rows) * (dim >> 3);
  launch_kernel(broadcast_row_kernel, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, s, src, dim, rows, dst);
  return check_launch(err);
}

int gather_rows_launch(const __nv_bfloat16* src, long long src_bs, int ld, int r0, int rows, int batches,
                       int dim, __nv_bfloat16* dst, cudaStream_t s, const char** err, const int32_t* len, int len_ref,
                       int* err_flag) {
  const long long total = static_cast<long long>(batches) * rows * (dim >> 3);
  launch_kernel(gather_rows_kernel, dim3(static_cast<int>((total + 255) / 256)), dim3(256), 0, s, src, src_bs, ld, r0, rows, batches,
                                                                          dim, dst, len, len_ref, err_flag);
  return check_launch(err);
}

}  // namespace vla
