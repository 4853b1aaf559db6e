// Fused softmax(QK^T / sqrt(hd)) V for the three attention shapes of the path:
//   DINOv2   : 16 heads x 64, S = 261, bidirectional        (timm Attention, SDPA scale hd^-0.5)
//   SigLIP   : 16 heads x 72, S = 256, bidirectional        (head dim zero-padded to 80 in smem)
//   Qwen2.5  : 14 q heads / 2 kv heads x 64, S ~ 625, causal (transformers Qwen2Attention, repeat_kv)
// v1 implementation: FlashAttention-2 style tiling (64 q rows x 64 kv rows per step, online softmax in
// fp32, cp.async double-buffered K/V) on warp-level mma.sync.m16n8k16 bf16 tensor-core tiles.
// Attention is ~4 % of the path's FLOPs; the tcgen05 GEMM carries the other 96 %.
#include "common.cuh"
#include "launch.cuh"
#include "ops.cuh"

#include <cstdio>

namespace vla {

namespace {

constexpr int ATT_BM = 64;   // q rows per CTA (4 warps x 16)
constexpr int ATT_BN = 64;   // kv rows per step
constexpr int ATT_THREADS = 128;

VLA_DEVINL void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
VLA_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
VLA_DEVINL void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
VLA_DEVINL void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
VLA_DEVINL void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
VLA_DEVINL void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int HD, int HDP>
__global__ void __launch_bounds__(ATT_THREADS)
flash_attn_kernel(const __nv_bfloat16* __restrict__ qp, int ld_q, int Sq, const __nv_bfloat16* __restrict__ kp,
                  const __nv_bfloat16* __restrict__ vp, int ld, int S, int group, int causal, float scale_log2,
                  __nv_bfloat16* __restrict__ out, int ld_out) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, their writes visible
  pdl_launch_dependents();
  constexpr int LDS = HDP + 8;          // smem row stride (elements): odd multiple of 16 B
  constexpr int CH = HD / 8;            // 16-byte chunks per global row
  constexpr int KSTEPS = HDP / 16;      // k-steps of QK^T
  constexpr int ONB = HDP / 8;          // n-blocks of the output
  extern __shared__ __align__(16) uint8_t att_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(att_smem);
  __nv_bfloat16* sK = sQ + ATT_BM * LDS;            // [2][ATT_BN][LDS]
  __nv_bfloat16* sV = sK + 2 * ATT_BN * LDS;        // [2][ATT_BN][LDS]

  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / group;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // q rows: [b*Sq, (b+1)*Sq) of qp; k/v rows: [b*S, (b+1)*S) of kp/vp (self-attention passes Sq == S)
  const long long q_base = static_cast<long long>(b) * Sq;
  const long long row_base = static_cast<long long>(b) * S;
  const __nv_bfloat16* gq = qp + q_base * ld_q + h * HD;
  const __nv_bfloat16* gk = kp + row_base * ld + kvh * HD;
  const __nv_bfloat16* gv = vp + row_base * ld + kvh * HD;

  // zero the pad columns once (cp.async never writes them)
  if (HDP > HD) {
    for (int i = tid; i < ATT_BM * 5; i += ATT_THREADS) {
      const int r = i / 5, which = i % 5;
      __nv_bfloat16* base = which == 0 ? sQ : (which <= 2 ? sK + (which - 1) * ATT_BN * LDS
                                                          : sV + (which - 3) * ATT_BN * LDS);
      *reinterpret_cast<uint4*>(base + r * LDS + HD) = make_uint4(0, 0, 0, 0);
    }
  }

  const int q0 = qb * ATT_BM;
  for (int i = tid; i < ATT_BM * CH; i += ATT_THREADS) {
    const int r = i / CH, c = i % CH;
    const bool ok = (q0 + r) < Sq;
    cp_async16(smem_u32(sQ + r * LDS + c * 8), gq + static_cast<long long>(ok ? q0 + r : 0) * ld_q + c * 8, ok);
  }
  auto load_kv = [&](int it, int buf) {
    const int k0 = it * ATT_BN;
    for (int i = tid; i < ATT_BN * CH; i += ATT_THREADS) {
      const int r = i / CH, c = i % CH;
      const bool ok = (k0 + r) < S;
      const long long off = static_cast<long long>(ok ? k0 + r : 0) * ld + c * 8;
      cp_async16(smem_u32(sK + (buf * ATT_BN + r) * LDS + c * 8), gk + off, ok);
      cp_async16(smem_u32(sV + (buf * ATT_BN + r) * LDS + c * 8), gv + off, ok);
    }
  };
  int n_it = (S + ATT_BN - 1) / ATT_BN;
  if (causal && qb + 1 < n_it) n_it = qb + 1;
  load_kv(0, 0);
  cp_async_commit();

  float o[ONB][4];
#pragma unroll
  for (int i = 0; i < ONB; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t qf[KSTEPS][4];

  const int g = lane >> 2, t = lane & 3;
  const int qrow0 = q0 + warp * 16 + g;  // this thread's rows: qrow0 and qrow0 + 8

  for (int it = 0; it < n_it; ++it) {
    const int buf = it & 1;
    if (it + 1 < n_it) load_kv(it + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();

    if (it == 0) {
#pragma unroll
      for (int kk = 0; kk < KSTEPS; ++kk) {
        const uint32_t addr = smem_u32(sQ + (warp * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8);
        ldsm_x4(addr, qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
      }
    }

    // ---- S = Q K^T
    float sc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
    const __nv_bfloat16* bK = sK + buf * ATT_BN * LDS;
#pragma unroll
    for (int kk = 0; kk < KSTEPS; ++kk) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        const uint32_t addr =
            smem_u32(bK + (np * 16 + (lane & 7) + ((lane >> 4) << 3)) * LDS + kk * 16 + ((lane >> 3) & 1) * 8);
        ldsm_x4(addr, b0, b1, b2, b3);
        mma_bf16(sc[2 * np], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b0, b1);
        mma_bf16(sc[2 * np + 1], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b2, b3);
      }
    }

    // ---- mask + online softmax (rows g and g+8 of this warp's 16)
    const int k0 = it * ATT_BN;
    const bool need_mask = (k0 + ATT_BN > S) || (causal && it == qb);
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (need_mask) {
          const int col = k0 + nb * 8 + t * 2 + (e & 1);
          const int row = qrow0 + (e >> 1) * 8;
          if (col >= S || (causal && col > row)) sc[nb][e] = -INFINITY;
        }
        mx[e >> 1] = fmaxf(mx[e >> 1], sc[nb][e]);
      }
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f((m_run[r] - m_new) * scale_log2);
      m_run[r] = m_new;
      msc[r] = (m_new == -INFINITY) ? 0.f : m_new * scale_log2;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[8][2];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      const float p0 = exp2f(sc[nb][0] * scale_log2 - msc[0]);
      const float p1 = exp2f(sc[nb][1] * scale_log2 - msc[0]);
      const float p2 = exp2f(sc[nb][2] * scale_log2 - msc[1]);
      const float p3 = exp2f(sc[nb][3] * scale_log2 - msc[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nb][0] = pack_bf16(p0, p1);
      pf[nb][1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int i = 0; i < ONB; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }

    // ---- O += P V
    const __nv_bfloat16* bV = sV + buf * ATT_BN * LDS;
#pragma unroll
    for (int kk = 0; kk < ATT_BN / 16; ++kk) {
      const uint32_t a0 = pf[2 * kk][0], a1 = pf[2 * kk][1], a2 = pf[2 * kk + 1][0], a3 = pf[2 * kk + 1][1];
#pragma unroll
      for (int np = 0; np < ONB / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const uint32_t addr =
            smem_u32(bV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + np * 16 + (lane >> 4) * 8);
        ldsm_x4_t(addr, b0, b1, b2, b3);
        mma_bf16(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    __syncthreads();
  }

  // ---- finalize: O / l, bf16, store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  __nv_bfloat16* go = out + q_base * ld_out + h * HD;
#pragma unroll
  for (int nb = 0; nb < ONB; ++nb) {
    const int col = nb * 8 + t * 2;
    if (col < HD) {
      if (qrow0 < Sq)
        *reinterpret_cast<uint32_t*>(go + static_cast<long long>(qrow0) * ld_out + col) =
            pack_bf16(o[nb][0] * inv0, o[nb][1] * inv0);
      if (qrow0 + 8 < Sq)
        *reinterpret_cast<uint32_t*>(go + static_cast<long long>(qrow0 + 8) * ld_out + col) =
            pack_bf16(o[nb][2] * inv1, o[nb][3] * inv1);
    }
  }
}


// ------------------------------------------------------------------ few-query attention (the policy's Bridge-Attention)
// Sq <= 16 query rows against hundreds of keys (T = 8 queries x 585 keys x hd 112 in the LIBERO config): one CTA per
// (16-row query tile, head, sample); its four warps split the KEYS (each runs an online softmax over its quarter
// with cp.async double-buffered 32-key tiles), then merge their (max, sum, O) through shared memory.  The generic
// kernel above gives each warp its own 16 query rows, which leaves three of four warps idle when there are 8 queries.
constexpr int SK_BN = 32;  // keys per step and warp

template <int HD, int HDP>
__global__ void __launch_bounds__(ATT_THREADS)
splitkv_attn_kernel(const __nv_bfloat16* __restrict__ qp, int ld_q, int Sq, const __nv_bfloat16* __restrict__ kp,
                    const __nv_bfloat16* __restrict__ vp, int ld, int S, int group, float scale_log2,
                    __nv_bfloat16* __restrict__ out, int ld_out, int q_rows) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int LDS = HDP + 8;
  constexpr int CH = HD / 8;
  constexpr int KSTEPS = HDP / 16;
  constexpr int ONB = HDP / 8;
  extern __shared__ __align__(16) uint8_t att_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(att_smem);  // [16][LDS]
  __nv_bfloat16* sKV = sQ + 16 * LDS;                               // per warp: [2 stages][K | V][SK_BN][LDS]
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / group;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const __nv_bfloat16* gq = qp + static_cast<long long>(b) * q_rows * ld_q + h * HD;
  const __nv_bfloat16* gk = kp + static_cast<long long>(b) * S * ld + kvh * HD;
  const __nv_bfloat16* gv = vp + static_cast<long long>(b) * S * ld + kvh * HD;
  __nv_bfloat16* wK = sKV + warp * (4 * SK_BN * LDS);
  __nv_bfloat16* wV = wK + 2 * SK_BN * LDS;

  if (HDP > HD) {
    for (int i = tid; i < 16; i += ATT_THREADS) *reinterpret_cast<uint4*>(sQ + i * LDS + HD) = make_uint4(0, 0, 0, 0);
    for (int i = lane; i < 4 * SK_BN; i += 32) *reinterpret_cast<uint4*>(wK + i * LDS + HD) = make_uint4(0, 0, 0, 0);
  }
  const int q0 = qb * 16;
  for (int i = tid; i < 16 * CH; i += ATT_THREADS) {
    const int r = i / CH, c = i % CH;
    const bool ok = (q0 + r) < Sq;
    cp_async16(smem_u32(sQ + r * LDS + c * 8), gq + static_cast<long long>(ok ? q0 + r : 0) * ld_q + c * 8, ok);
  }
  // this warp's keys: a contiguous quarter, in steps of SK_BN
  const int per_warp = ((S + 4 * SK_BN - 1) / (4 * SK_BN)) * SK_BN;
  const int k_begin = warp * per_warp, k_end = min(S, k_begin + per_warp);
  const int n_it = k_end > k_begin ? (k_end - k_begin + SK_BN - 1) / SK_BN : 0;
  auto load_kv = [&](int it, int buf) {
    const int k0 = k_begin + it * SK_BN;
    for (int i = lane; i < SK_BN * CH; i += 32) {
      const int r = i / CH, c = i % CH;
      const bool ok = (k0 + r) < k_end;
      const long long off = static_cast<long long>(ok ? k0 + r : 0) * ld + c * 8;
      cp_async16(smem_u32(wK + (buf * SK_BN + r) * LDS + c * 8), gk + off, ok);
      cp_async16(smem_u32(wV + (buf * SK_BN + r) * LDS + c * 8), gv + off, ok);
    }
  };
  if (n_it) load_kv(0, 0);
  cp_async_commit();

  float o[ONB][4];
#pragma unroll
  for (int i = 0; i < ONB; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t qf[KSTEPS][4];
  const int g = lane >> 2, t = lane & 3;

  cp_async_wait<0>();
  __syncthreads();  // Q (loaded by all threads) is visible; every warp's first tile is its own business
#pragma unroll
  for (int kk = 0; kk < KSTEPS; ++kk) {
    const uint32_t addr = smem_u32(sQ + (lane & 15) * LDS + kk * 16 + (lane >> 4) * 8);
    ldsm_x4(addr, qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
  }
  for (int it = 0; it < n_it; ++it) {
    const int buf = it & 1;
    if (it + 1 < n_it) load_kv(it + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    float sc[SK_BN / 8][4];
#pragma unroll
    for (int i = 0; i < SK_BN / 8; ++i) sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
    const __nv_bfloat16* bK = wK + buf * SK_BN * LDS;
#pragma unroll
    for (int kk = 0; kk < KSTEPS; ++kk) {
#pragma unroll
      for (int np = 0; np < SK_BN / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        const uint32_t addr =
            smem_u32(bK + (np * 16 + (lane & 7) + ((lane >> 4) << 3)) * LDS + kk * 16 + ((lane >> 3) & 1) * 8);
        ldsm_x4(addr, b0, b1, b2, b3);
        mma_bf16(sc[2 * np], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b0, b1);
        mma_bf16(sc[2 * np + 1], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b2, b3);
      }
    }
    const int k0 = k_begin + it * SK_BN;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < SK_BN / 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = k0 + nb * 8 + t * 2 + (e & 1);
        if (col >= k_end) sc[nb][e] = -INFINITY;
        mx[e >> 1] = fmaxf(mx[e >> 1], sc[nb][e]);
      }
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f((m_run[r] - m_new) * scale_log2);
      m_run[r] = m_new;
      msc[r] = (m_new == -INFINITY) ? 0.f : m_new * scale_log2;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[SK_BN / 8][2];
#pragma unroll
    for (int nb = 0; nb < SK_BN / 8; ++nb) {
      const float p0 = exp2f(sc[nb][0] * scale_log2 - msc[0]);
      const float p1 = exp2f(sc[nb][1] * scale_log2 - msc[0]);
      const float p2 = exp2f(sc[nb][2] * scale_log2 - msc[1]);
      const float p3 = exp2f(sc[nb][3] * scale_log2 - msc[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[nb][0] = pack_bf16(p0, p1);
      pf[nb][1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int i = 0; i < ONB; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    const __nv_bfloat16* bV = wV + buf * SK_BN * LDS;
#pragma unroll
    for (int kk = 0; kk < SK_BN / 16; ++kk) {
      const uint32_t a0 = pf[2 * kk][0], a1 = pf[2 * kk][1], a2 = pf[2 * kk + 1][0], a3 = pf[2 * kk + 1][1];
#pragma unroll
      for (int np = 0; np < ONB / 2; ++np) {
        uint32_t b0, b1, b2, b3;
        const uint32_t addr =
            smem_u32(bV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + np * 16 + (lane >> 4) * 8);
        ldsm_x4_t(addr, b0, b1, b2, b3);
        mma_bf16(o[2 * np], a0, a1, a2, a3, b0, b1);
        mma_bf16(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    __syncwarp();
  }
  // ---- merge the four warps' partial results: per query row (max, sum) and the 16 x HD partial outputs
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  cp_async_wait<0>();
  __syncthreads();  // all warps are done with their K/V tiles: the tile memory becomes the merge buffer
  float* mrg_m = reinterpret_cast<float*>(sKV);            // [4][16]
  float* mrg_l = mrg_m + 64;                               // [4][16]
  float* mrg_o = mrg_l + 64;                               // [4][16][HDP]
  if (t == 0) {
    mrg_m[warp * 16 + g] = m_run[0];
    mrg_m[warp * 16 + g + 8] = m_run[1];
    mrg_l[warp * 16 + g] = l_run[0];
    mrg_l[warp * 16 + g + 8] = l_run[1];
  }
#pragma unroll
  for (int nb = 0; nb < ONB; ++nb) {
    const int col = nb * 8 + t * 2;
    *reinterpret_cast<float2*>(mrg_o + (warp * 16 + g) * HDP + col) = make_float2(o[nb][0], o[nb][1]);
    *reinterpret_cast<float2*>(mrg_o + (warp * 16 + g + 8) * HDP + col) = make_float2(o[nb][2], o[nb][3]);
  }
  __syncthreads();
  // every thread finalises a few (row, column pair) outputs
  for (int i = tid; i < 16 * (HD / 2); i += ATT_THREADS) {
    const int r = i / (HD / 2), c = (i % (HD / 2)) * 2;
    if (q0 + r >= Sq) continue;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) M = fmaxf(M, mrg_m[w * 16 + r]);
    float L = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float mw = mrg_m[w * 16 + r];
      const float f = (mw == -INFINITY) ? 0.f : exp2f((mw - M) * scale_log2);
      L += f * mrg_l[w * 16 + r];
      const float2 ov = *reinterpret_cast<const float2*>(mrg_o + (w * 16 + r) * HDP + c);
      a0 += f * ov.x;
      a1 += f * ov.y;
    }
    const float inv = 1.f / L;
    *reinterpret_cast<uint32_t*>(out + (static_cast<long long>(b) * q_rows + q0 + r) * ld_out + h * HD + c) =
        pack_bf16(a0 * inv, a1 * inv);
  }
}

template <int HD, int HDP>
int launch_splitkv(const __nv_bfloat16* q, int ld_q, int Sq, const __nv_bfloat16* k, const __nv_bfloat16* v, int ld,
                   int B, int S, int n_heads, int group, __nv_bfloat16* out, int ld_out, int q_rows, cudaStream_t s,
                   const char** err) {
  constexpr int LDS = HDP + 8;
  constexpr int SMEM_TILES = (16 + 4 * 4 * SK_BN) * LDS * 2;
  constexpr int SMEM_MERGE = 16 * LDS * 2 + (128 + 4 * 16 * HDP) * 4;
  constexpr int SMEM = SMEM_TILES > SMEM_MERGE ? SMEM_TILES : SMEM_MERGE;
  static PerDeviceFlag attr_flag;  // the shared-memory opt-in is per device
  bool& attr_set = attr_flag.here();
  if (!attr_set) {
    if (cudaFuncSetAttribute(splitkv_attn_kernel<HD, HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) !=
        cudaSuccess) {
      if (err) *err = "attention: cudaFuncSetAttribute failed";
      return -4;
    }
    attr_set = true;
  }
  const float scale_log2 = (1.0f / sqrtf(static_cast<float>(HD))) * 1.4426950408889634f;
  dim3 grid((Sq + 15) / 16, n_heads, B);
  launch_kernel(splitkv_attn_kernel<HD, HDP>, dim3(grid), dim3(ATT_THREADS), SMEM, s, q, ld_q, Sq, k, v, ld, S, group,
                scale_log2, out, ld_out, q_rows);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = cudaGetErrorString(e);
    return -4;
  }
  ops_count_launch();
  return 0;
}

template <int HD, int HDP>
int launch_attn(const __nv_bfloat16* q, int ld_q, int Sq, const __nv_bfloat16* k, const __nv_bfloat16* v, int ld,
                int B, int S, int n_heads, int group, int causal, __nv_bfloat16* out, int ld_out, cudaStream_t s,
                const char** err) {
  constexpr int LDS = HDP + 8;
  constexpr int SMEM = (ATT_BM + 4 * ATT_BN) * LDS * 2;
  static PerDeviceFlag attr_flag;  // the shared-memory opt-in is per device
  bool& attr_set = attr_flag.here();
  if (!attr_set) {
    if (cudaFuncSetAttribute(flash_attn_kernel<HD, HDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) !=
        cudaSuccess) {
      if (err) *err = "attention: cudaFuncSetAttribute failed";
      return -4;
    }
    attr_set = true;
  }
  const float scale_log2 = (1.0f / sqrtf(static_cast<float>(HD))) * 1.4426950408889634f;
  dim3 grid((Sq + ATT_BM - 1) / ATT_BM, n_heads, B);
  launch_kernel(flash_attn_kernel<HD, HDP>, dim3(grid), dim3(ATT_THREADS), SMEM, s, q, ld_q, Sq, k, v, ld, S, group, causal, scale_log2,
                                                            out, ld_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = cudaGetErrorString(e);
    return -4;
  }
  ops_count_launch();
  return 0;
}

}  // namespace

static int g_attn_impl = 0;  // 0 = auto, 1 = force the mma.sync kernel, 2 = force tcgen05 whenever the head dim allows
void attention_set_impl(int impl) { g_attn_impl = impl; }

int attention_launch(const __nv_bfloat16* qkv, int ld_qkv, int q_off, int k_off, int v_off, int B, int S,
                     int n_heads, int group, int hd, int causal, __nv_bfloat16* out, int ld_out,
                     cudaStream_t s, const char** err) {
  if ((ld_qkv & 7) || (q_off & 7) || (k_off & 7) || (v_off & 7) || (ld_out & 1) || group <= 0 ||
      (n_heads % group)) {
    if (err) *err = "attention: offsets/strides must be multiples of 8 elements";
    return -1;
  }
  return cross_attention_launch(qkv + q_off, ld_qkv, S, qkv + k_off, qkv + v_off, ld_qkv, S, B, n_heads, group, hd,
                                causal, out, ld_out, s, err);
}

int cross_attention_launch(const __nv_bfloat16* q, int ld_q, int Sq, const __nv_bfloat16* k, const __nv_bfloat16* v,
                           int ld_kv, int Skv, int B, int n_heads, int group, int hd, int causal,
                           __nv_bfloat16* out, int ld_out, cudaStream_t s, const char** err) {
  if ((ld_q & 7) || (ld_kv & 7) || (ld_out & 1) || group <= 0 || (n_heads % group) ||
      (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(k) & 15) ||
      (reinterpret_cast<uintptr_t>(v) & 15)) {
    if (err) *err = "attention: strides must be multiples of 8 elements and pointers 16-byte aligned";
    return -1;
  }
  if (causal && Sq != Skv) {
    if (err) *err = "attention: causal masking needs Sq == Skv";
    return -1;
  }
  // The large shapes (hd 64 / 72, at least one full-ish query tile) run on the tcgen05 kernel (attention_tc.cu).
  if (g_attn_impl != 1 && (hd == 64 || hd == 72) && (Sq >= 32 || g_attn_impl == 2)) {
    // A few query rows past the last full 128-row tile (DINOv2: 261 = 2 x 128 + 5) would occupy a whole tcgen05 work
    // item - a 128-row MMA tile and every K/V tile of the head - for a handful of rows.  Those rows go to the
    // key-splitting kernel instead, the full tiles to the tensor-core kernel; both see all the keys.
    const int rem = Sq % 128;
    static const bool no_split = getenv("VLA_FA_NO_ROW_SPLIT") != nullptr;
    if (!causal && hd == 64 && Sq > 128 && rem > 0 && rem <= 16 && Skv >= 128 && !no_split) {
      const int full = Sq - rem;
      int rc = attention_tc_launch(q, ld_q, full, k, v, ld_kv, Skv, B, n_heads, group, hd, 0, out, ld_out, Sq, s, err);
      if (rc < 0) return rc;
      if (rc == 0)
        return launch_splitkv<64, 64>(q + static_cast<long long>(full) * ld_q, ld_q, rem, k, v, ld_kv, B, Skv, n_heads,
                                      group, out + static_cast<long long>(full) * ld_out, ld_out, Sq, s, err);
    }
    const int rc = attention_tc_launch(q, ld_q, Sq, k, v, ld_kv, Skv, B, n_heads, group, hd, causal, out, ld_out, Sq, s, err);
    if (rc <= 0) return rc;
    // The shape is outside what the tensor-core kernel serves (work-item table, alignment).  The mma.sync kernel below
    // computes the same result ~2x slower; say so once instead of degrading silently - and refuse when the caller asked
    // for the tensor-core kernel explicitly (vla_set_attention_impl(2)).
    if (g_attn_impl == 2) {
      if (err) *err = "attention: shape not served by the tcgen05 kernel (vla_set_attention_impl(2) forbids the mma.sync path)";
      return -1;
    }
    static bool warned = false;
    if (!warned) {
      warned = true;
      fprintf(stderr, "libvla_b200: attention shape (Sq %d, Skv %d, heads %d, hd %d) runs on the slower mma.sync kernel\n",
              Sq, Skv, n_heads, hd);
    }
  }
  // few queries against many keys (the policy's Bridge-Attention): the warps split the keys instead of the queries
  if (!causal && Sq <= 32 && Skv >= 128 && g_attn_impl != 1 && !(ld_out & 1)) {
    if (hd == 112) return launch_splitkv<112, 112>(q, ld_q, Sq, k, v, ld_kv, B, Skv, n_heads, group, out, ld_out, Sq, s, err);
    if (hd == 64) return launch_splitkv<64, 64>(q, ld_q, Sq, k, v, ld_kv, B, Skv, n_heads, group, out, ld_out, Sq, s, err);
  }
  if (hd == 64) return launch_attn<64, 64>(q, ld_q, Sq, k, v, ld_kv, B, Skv, n_heads, group, causal, out, ld_out, s, err);
  if (hd == 72) return launch_attn<72, 80>(q, ld_q, Sq, k, v, ld_kv, B, Skv, n_heads, group, causal, out, ld_out, s, err);
  if (hd == 112) return launch_attn<112, 112>(q, ld_q, Sq, k, v, ld_kv, B, Skv, n_heads, group, causal, out, ld_out, s, err);
  if (err) *err = "attention: head dim must be 64, 72 or 112";
  return -1;
}

}  // namespace vla
