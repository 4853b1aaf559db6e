// The whole Bridge-Attention policy (24 blocks) as ONE kernel for small batches.
//
// Reference: MLPResNet.forward (prismatic/models/action_heads.py:111-121) looping MLPResNetBlock.forward (:218-283) or
// MLPResNetBlock_Pro.forward (:337-410).  At bs <= 8 the engine's per-block chain (q projection, self K|V projection,
// [RoPE,] attention, o projection + residual, LayerNorm, ffn + ReLU: 7-8 launches on 8 x 896 activations) is bound by
// launch latency, not by work: 24 blocks x 8 launches are ~1.2 ms of the 4.9 ms bs=1 forward.  Here one thread-block
// CLUSTER of 8 CTAs owns one sample and walks all 24 blocks; CTA h owns attention head h (8 heads x 112 = 896):
//   phase 1  q_h, k_self_h, v_self_h = x W^T + b for head h's 3 x 112 columns (mma.sync m16n8k16, A = the sample's
//            T <= 16 rows of x in shared memory, B = weight rows streamed from L2 with 256-bit loads), Pro: RoPE on
//            q_h / k_self_h; q_h stays in shared memory, k|v go to rows [0, T) of the block's key/value buffer
//   phase 2  softmax(q_h K_h^T / sqrt(112)) V_h over the NK = T + 65 + NP keys of the buffer (the cond / vision rows were
//            projected beforehand by the tcgen05 GEMM on the side stream): the 12 warps split the keys, online softmax
//            per warp over double-buffered cp.async tiles, merge through shared memory (the algorithm of
//            splitkv_attn_kernel in attention.cu)
//   -- cluster barrier --   (the 8 heads' outputs are exchanged through global memory / L2)
//   phase 3  y[:, 112h : 112h+112] = o Wo^T + bo + x          (fp32 add of the residual, one bf16 rounding)
//   -- cluster barrier --
//   phase 4  LayerNorm(y) over the full 896 columns (every CTA normalises the T rows itself: 14 KB of reads)
//   phase 5  x'[:, 112h : 112h+112] = ReLU(LN(y) Wffn^T + b)
//   -- cluster barrier --   next block
// Every bf16 rounding of the multi-kernel path is kept at the same place (projection outputs, RoPE products, P before
// P V, the attention output, y, LN output, x'), so the two paths agree to fp32 summation order.
// Beside the worker clusters (one per sample) the grid holds PF_PREFETCH_CL clusters of L2 prefetchers (below).
// Measured (bs=1, profiles/r02_policy_fused_phases.txt): ~30 us per block, 0.77 ms for the 24 blocks against ~1.0 ms for
// the per-block launches; the weight phases run at the one SM's L2 read rate (~36 B/clk).
#include "common.cuh"
#include "launch.cuh"
#include "ops.cuh"
#include "policy_fused.cuh"

namespace vla {

namespace {

constexpr int PF_THREADS = 384;
constexpr int PF_WARPS = 12;       // 12 warps x 168 registers (14 warps / deeper unrolls measured slower: the loads then
                                  // queue at the L1 tag stage - 8 lines per warp load - instead of waiting in registers)
constexpr int PF_CL = 8;          // CTAs per cluster = attention heads
constexpr int D = 896, HD = 112, PKV = 1792;
constexpr int LDA = 904;          // row stride of the A operand in shared memory (bf16): 452 words = 4 banks per row, so
                                  // the 16-byte fragment loads (32 bytes apart along t) of rows g and g+1 do not collide
constexpr int LDQ = 120;          // row stride of q_h / staged K, V tiles
constexpr int SKB = 16;           // keys per step and warp in the attention phase
constexpr int NT_MAX = 4;         // column tiles (8 wide) one warp carries at once in phase 1 (42 tiles / 12 warps)
constexpr int NT_OUT = (14 + PF_WARPS - 1) / PF_WARPS;  // ... in phases 3 and 5 (14 tiles)
constexpr int UNROLL_OUT = 7;
static_assert(PF_WARPS * NT_MAX >= 42, "phase 1 is a single pass");
static_assert(PF_WARPS * 16 * (HD + 2) * 4 <= PF_WARPS * 4 * SKB * LDQ * 2, "the merge buffer reuses the K/V tile memory");

VLA_DEVINL void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
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
VLA_DEVINL void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

VLA_DEVINL void ldg256(const void* p, uint4& lo, uint4& hi) {  // sm_100: 256-bit loads (LDG.E.256), 32-byte aligned
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}

// acc[i] (16 x 8, rows of A x weight rows n0[i] .. n0[i]+7) = A[16 x 896] W^T for `nt` column tiles of this warp at once.
// The weight words come straight from global memory (W is [N][896], K contiguous) with ONE 32-byte load per lane, tile
// and 64-wide k block - the four lanes of a row fetch a whole 128-byte line per instruction - and UNROLL blocks in
// flight: the phase is bound by the load path of the one SM that streams its share of the block's weights from L2
// (latency, and the L1 tag stage that takes one line per cycle), so bytes per instruction and bytes in flight are what
// counts.  (History: 4-byte loads, two k steps in flight: ~28 us per phase; 16-byte loads: 9 / 5 us.)  The k index
// inside a 64-block is permuted the same way for A and B (a sum over k does not care): lane t owns elements
// [16t, 16t+16) of the block; mma step s = 0..3 takes elements 16t+4s+{0,1} as its "k = 2t, 2t+1" pair and
// 16t+4s+{2,3} as its "k = 2t+8, 2t+9" pair, which makes both fragments contiguous: the A fragments of a block are
// two 16-byte shared-memory loads per row.
template <int NT, int UNROLL>
VLA_DEVINL void warp_gemm_16xK(const __nv_bfloat16* sA, const __nv_bfloat16* (&wrow)[NT], int nt, int lane,
                               float (&acc)[NT][4]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const __nv_bfloat16* a_lo = sA + g * LDA + 16 * t;        // row g
  const __nv_bfloat16* a_hi = sA + (g + 8) * LDA + 16 * t;  // row g + 8
  static_assert((D / 64) % UNROLL == 0, "k blocks must divide by the unroll factor");
  for (int kb = 0; kb < D; kb += 64 * UNROLL) {
    uint4 w0[UNROLL][NT], w1[UNROLL][NT];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int i = 0; i < NT; ++i)
        if (i < nt) ldg256(wrow[i] + kb + 64 * u + 16 * t, w0[u][i], w1[u][i]);  // wrow[i]: row n0 + g
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint4 lo0 = *reinterpret_cast<const uint4*>(a_lo + kb + 64 * u), lo1 = *reinterpret_cast<const uint4*>(a_lo + kb + 64 * u + 8);
      const uint4 hi0 = *reinterpret_cast<const uint4*>(a_hi + kb + 64 * u), hi1 = *reinterpret_cast<const uint4*>(a_hi + kb + 64 * u + 8);
#pragma unroll
      for (int i = 0; i < NT; ++i) {
        if (i < nt) {
          mma_bf16(acc[i], lo0.x, hi0.x, lo0.y, hi0.y, w0[u][i].x, w0[u][i].y);
          mma_bf16(acc[i], lo0.z, hi0.z, lo0.w, hi0.w, w0[u][i].z, w0[u][i].w);
          mma_bf16(acc[i], lo1.x, hi1.x, lo1.y, hi1.y, w1[u][i].x, w1[u][i].y);
          mma_bf16(acc[i], lo1.z, hi1.z, lo1.w, hi1.w, w1[u][i].z, w1[u][i].w);
        }
      }
    }
  }
}

// rows [0, T) x 896 of a global bf16 matrix -> the A operand in shared memory (rows T..15 stay zero)
VLA_DEVINL void load_rows_to_A(__nv_bfloat16* sA, const __nv_bfloat16* src, int T, int tid) {
  for (int i = tid; i < T * (D / 8); i += PF_THREADS) {
    const int r = i / (D / 8), c = i % (D / 8);
    // (__ldcg: the rows were written by the other CTAs of the cluster - never through this SM's L1)
    *reinterpret_cast<uint4*>(sA + r * LDA + c * 8) = __ldcg(reinterpret_cast<const uint4*>(src + static_cast<long long>(r) * D + c * 8));
  }
}

// ---- L2 prefetchers.  The policy's weights (8 MB per block, 193 MB in all) and the cond / vision K|V rows (2 MB per
// block and sample) do not survive in L2 from one call to the next (the towers and the LLM stream ~2.5 GB in between),
// and 8 worker SMs cannot pull them from HBM at more than ~1 TB/s, latency-bound.  The SMs the small batch leaves idle
// do that part: PF_PREFETCH_CL extra clusters run ahead of the workers by PF_AHEAD blocks (paced through one progress
// word the workers of sample 0 publish; L2 holds ~12 blocks, so running further ahead would evict what was fetched)
// and issue bulk L2 prefetches.  They are an optimisation only: nothing waits for them, and they give up after
// PF_PREFETCH_LIMIT clocks whatever the workers do.
constexpr int PF_PREFETCH_CL = 4;
constexpr int PF_AHEAD = 2;
constexpr long long PF_PREFETCH_LIMIT = 3000000;  // ~1.6 ms at 1.9 GHz
constexpr int PF_CHUNK = 4096;

VLA_DEVINL void prefetch_l2_bulk(const void* p, unsigned int bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __noinline__ void policy_prefetch_role(const PolicyFusedArgs& a, int pcta, int npcta) {
  const long long tid_g = static_cast<long long>(pcta) * PF_THREADS + threadIdx.x;
  const long long nthr = static_cast<long long>(npcta) * PF_THREADS;
  const long long t0 = clock64();
  const volatile int* progress = a.progress;
  for (int blk = 0; blk < a.n_blocks; ++blk) {
    const int need = blk - PF_AHEAD;
    if (need > 0) {
      int give_up = 0;
      if ((threadIdx.x & 31) == 0) {
        while (*progress < need) {
          if (clock64() - t0 > PF_PREFETCH_LIMIT) { give_up = 1; break; }
          __nanosleep(200);
        }
      }
      if (__shfl_sync(0xffffffffu, give_up, 0)) return;
    }
    const PolicyBlockW& w = a.blocks[blk];
    const void* base[5] = {w.wq, w.wkvs, w.wo, w.wffn, w.kv};
    const long long bytes[5] = {2LL * D * D, 4LL * D * D, 2LL * D * D, 2LL * D * D, 2LL * a.B * a.NK * PKV};
#pragma unroll
    for (int sgm = 0; sgm < 5; ++sgm)
      for (long long c = tid_g * PF_CHUNK; c < bytes[sgm]; c += nthr * PF_CHUNK) {
        const long long left = bytes[sgm] - c;
        prefetch_l2_bulk(static_cast<const uint8_t*>(base[sgm]) + c, static_cast<unsigned int>(left < PF_CHUNK ? left : PF_CHUNK));
      }
  }
}

__global__ void __cluster_dims__(PF_CL, 1, 1) __launch_bounds__(PF_THREADS, 1)
policy_fused_kernel(const __grid_constant__ PolicyFusedArgs a) {
  extern __shared__ __align__(16) uint8_t pf_smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(pf_smem);        // [16][LDQ]
  __nv_bfloat16* sA = sQ + 16 * LDQ;                                     // [16][LDA]  (rows >= T are never read back)
  __nv_bfloat16* sKV = sA;  // per warp [2 stages][K | V][SKB][LDQ], then the merge buffer: the attention phase does not
                            // need the A operand (phase 3 reloads it), so the two share the memory
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int h = static_cast<int>(cluster_ctarank());                     // this CTA's attention head
  const int b = static_cast<int>(blockIdx.x) / PF_CL;                    // this cluster's sample
  const int T = a.T, NK = a.NK;
  const long long row0 = static_cast<long long>(b) * T;                  // first row of the sample in [B*T][896] buffers

  pdl_wait();
  pdl_launch_dependents();
  if (static_cast<int>(blockIdx.x) >= a.B * PF_CL) {  // whole clusters: no cluster barrier is shared with the workers
    policy_prefetch_role(a, static_cast<int>(blockIdx.x) - a.B * PF_CL, static_cast<int>(gridDim.x) - a.B * PF_CL);
    return;
  }

  // rows T..15 of A and q hold whatever the memory held: an output row of an MMA depends on its own A row only, and
  // rows >= T of every result are dropped
  for (int i = tid; i < 16 * LDQ / 8; i += PF_THREADS) reinterpret_cast<uint4*>(sQ)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const __nv_bfloat16* x_in = a.x0 + row0 * D;
  load_rows_to_A(sA, x_in, T, tid);
  __syncthreads();

  for (int blk = 0; blk < a.n_blocks; ++blk) {
    const PolicyBlockW& w = a.blocks[blk];  // constant bank
    if (blockIdx.x == 0 && tid == 0) *reinterpret_cast<volatile int*>(a.progress) = blk;
    const bool prof = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
    if (prof) a.prof[blk * 8 + 0] = clock64();
    __nv_bfloat16* kv = w.kv + static_cast<long long>(b) * NK * PKV;

    // ------------------------------------------------------------ phase 1: q_h | k_self_h | v_self_h  (42 column tiles)
    for (int base = warp; base < 42; base += PF_WARPS * NT_MAX) {
      const __nv_bfloat16* wrow[NT_MAX];
      int tiles[NT_MAX], nt = 0;
#pragma unroll
      for (int i = 0; i < NT_MAX; ++i) {
        const int tile = base + i * PF_WARPS;
        tiles[i] = tile;
        wrow[i] = w.wq;
        if (tile < 42) {
          nt = i + 1;
          const int seg = tile / 14, col = h * HD + (tile % 14) * 8 + g;  // output column within the segment's 896
          wrow[i] = (seg == 0 ? w.wq + static_cast<long long>(col) * D
                              : w.wkvs + static_cast<long long>((seg == 1 ? 0 : D) + col) * D);
        }
      }
      float acc[NT_MAX][4];
      warp_gemm_16xK<NT_MAX, 2>(sA, wrow, nt, lane, acc);
#pragma unroll
      for (int i = 0; i < NT_MAX; ++i) {
        if (i >= nt) continue;
        const int seg = tiles[i] / 14, hc = (tiles[i] % 14) * 8 + 2 * t;  // column within the head (even)
        const int col = h * HD + hc;
        const float* bias = seg == 0 ? w.bq + col : w.bkvs + (seg == 1 ? 0 : D) + col;
        const float b0 = __ldg(bias), b1 = __ldg(bias + 1);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int r = g + half * 8;
          if (r >= T) continue;
          float v0 = bf16_round(acc[i][2 * half] + b0), v1 = bf16_round(acc[i][2 * half + 1] + b1);
          if (a.pro && seg < 2) {  // apply_rope (AH:125-146): pair (2i, 2i+1), table angle of lane j, bf16 products
            const float* c = a.rope_cos + r * HD + hc;
            const float* s = a.rope_sin + r * HD + hc;
            const float y0 = bf16_round(bf16_round(v0 * c[0]) + bf16_round(-v1 * s[0]));
            const float y1 = bf16_round(bf16_round(v1 * c[1]) + bf16_round(v0 * s[1]));
            v0 = y0;
            v1 = y1;
          }
          const uint32_t pk = pack_bf16(v0, v1);
          if (seg == 0) *reinterpret_cast<uint32_t*>(sQ + r * LDQ + hc) = pk;
          else *reinterpret_cast<uint32_t*>(kv + static_cast<long long>(r) * PKV + (seg == 1 ? 0 : D) + col) = pk;
        }
      }
    }
    if (prof) a.prof[blk * 8 + 1] = clock64();
    __syncthreads();  // q_h complete in shared memory; this CTA's self K|V rows are written (read back below by cp.async)

    // ------------------------------------------------------------ phase 2: attention of head h over NK keys
    {
      const __nv_bfloat16* gk = kv + h * HD;
      const __nv_bfloat16* gv = kv + D + h * HD;
      __nv_bfloat16* wK = sKV + warp * (4 * SKB * LDQ);   // two stages of one K and one V tile per warp
      __nv_bfloat16* wV = wK + 2 * SKB * LDQ;
      const int per_warp = ((NK + PF_WARPS * SKB - 1) / (PF_WARPS * SKB)) * SKB;
      const int k_begin = warp * per_warp, k_end = min(NK, k_begin + per_warp);
      const int n_it = k_end > k_begin ? (k_end - k_begin + SKB - 1) / SKB : 0;
      auto load_kv = [&](int it, int buf) {
        const int k0 = k_begin + it * SKB;
        for (int i = lane; i < SKB * (HD / 8); i += 32) {
          const int r = i / (HD / 8), c = i % (HD / 8);
          const bool ok = (k0 + r) < k_end;
          const long long off = static_cast<long long>(ok ? k0 + r : 0) * PKV + c * 8;
          cp_async16(smem_u32(wK + (buf * SKB + r) * LDQ + c * 8), gk + off, ok);
          cp_async16(smem_u32(wV + (buf * SKB + r) * LDQ + c * 8), gv + off, ok);
        }
      };
      if (n_it) load_kv(0, 0);
      cp_async_commit();
      float o[HD / 8][4];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
      float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
      const uint32_t q_addr = smem_u32(sQ + (lane & 15) * LDQ + (lane >> 4) * 8);  // q fragments are re-read per step:
                                                                                   // 128 registers per thread at 16 warps
      const float sl2 = a.scale_log2;
      for (int it = 0; it < n_it; ++it) {
        const int buf = it & 1;
        if (it + 1 < n_it) load_kv(it + 1, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        float sc[2][4];
        sc[0][0] = sc[0][1] = sc[0][2] = sc[0][3] = sc[1][0] = sc[1][1] = sc[1][2] = sc[1][3] = 0.f;
        const __nv_bfloat16* bK = wK + buf * SKB * LDQ;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          uint32_t b0, b1, b2, b3, q0, q1, q2, q3;
          ldsm_x4(q_addr + kk * 32, q0, q1, q2, q3);
          ldsm_x4(smem_u32(bK + ((lane & 7) + ((lane >> 4) << 3)) * LDQ + kk * 16 + ((lane >> 3) & 1) * 8), b0, b1, b2, b3);
          mma_bf16(sc[0], q0, q1, q2, q3, b0, b1);
          mma_bf16(sc[1], q0, q1, q2, q3, b2, b3);
        }
        const int k0 = k_begin + it * SKB;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = k0 + nb * 8 + t * 2 + (e & 1);
            if (col >= k_end) sc[nb][e] = -INFINITY;
            mx[e >> 1] = fmaxf(mx[e >> 1], sc[nb][e]);
          }
        float corr[2], msc[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
          mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
          const float m_new = fmaxf(m_run[r], mx[r]);
          corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f((m_run[r] - m_new) * sl2);
          m_run[r] = m_new;
          msc[r] = (m_new == -INFINITY) ? 0.f : m_new * sl2;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[2][2];
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
          const float p0 = exp2f(sc[nb][0] * sl2 - msc[0]), p1 = exp2f(sc[nb][1] * sl2 - msc[0]);
          const float p2 = exp2f(sc[nb][2] * sl2 - msc[1]), p3 = exp2f(sc[nb][3] * sl2 - msc[1]);
          rs[0] += p0 + p1;
          rs[1] += p2 + p3;
          pf[nb][0] = pack_bf16(p0, p1);
          pf[nb][1] = pack_bf16(p2, p3);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
          o[i][0] *= corr[0]; o[i][1] *= corr[0];
          o[i][2] *= corr[1]; o[i][3] *= corr[1];
        }
        const __nv_bfloat16* bV = wV + buf * SKB * LDQ;
#pragma unroll
        for (int np = 0; np < HD / 16; ++np) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(smem_u32(bV + ((lane & 7) + ((lane >> 3) & 1) * 8) * LDQ + np * 16 + (lane >> 4) * 8), b0, b1, b2, b3);
          mma_bf16(o[2 * np], pf[0][0], pf[0][1], pf[1][0], pf[1][1], b0, b1);
          mma_bf16(o[2 * np + 1], pf[0][0], pf[0][1], pf[1][0], pf[1][1], b2, b3);
        }
        __syncwarp();
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
      }
      cp_async_wait<0>();
      if (prof) a.prof[blk * 8 + 2] = clock64();
      __syncthreads();  // every warp is done with its K/V tiles: the tile memory becomes the merge buffer
      float* mrg_m = reinterpret_cast<float*>(sKV);        // [8][16]
      float* mrg_l = mrg_m + PF_WARPS * 16;                // [8][16]
      float* mrg_o = mrg_l + PF_WARPS * 16;                // [8][16][HD]
      if (t == 0) {
        mrg_m[warp * 16 + g] = m_run[0];
        mrg_m[warp * 16 + g + 8] = m_run[1];
        mrg_l[warp * 16 + g] = l_run[0];
        mrg_l[warp * 16 + g + 8] = l_run[1];
      }
#pragma unroll
      for (int nb = 0; nb < HD / 8; ++nb) {
        const int col = nb * 8 + t * 2;
        *reinterpret_cast<float2*>(mrg_o + (warp * 16 + g) * HD + col) = make_float2(o[nb][0], o[nb][1]);
        *reinterpret_cast<float2*>(mrg_o + (warp * 16 + g + 8) * HD + col) = make_float2(o[nb][2], o[nb][3]);
      }
      __syncthreads();
      for (int i = tid; i < T * (HD / 2); i += PF_THREADS) {
        const int r = i / (HD / 2), c = (i % (HD / 2)) * 2;
        float M = -INFINITY;
#pragma unroll
        for (int ww = 0; ww < PF_WARPS; ++ww) M = fmaxf(M, mrg_m[ww * 16 + r]);
        float L = 0.f, a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int ww = 0; ww < PF_WARPS; ++ww) {
          const float mw = mrg_m[ww * 16 + r];
          const float f = (mw == -INFINITY) ? 0.f : exp2f((mw - M) * sl2);
          L += f * mrg_l[ww * 16 + r];
          const float2 ov = *reinterpret_cast<const float2*>(mrg_o + (ww * 16 + r) * HD + c);
          a0 += f * ov.x;
          a1 += f * ov.y;
        }
        const float inv = 1.f / L;
        *reinterpret_cast<uint32_t*>(a.ao + (row0 + r) * D + h * HD + c) = pack_bf16(a0 * inv, a1 * inv);
      }
    }
    if (prof) a.prof[blk * 8 + 3] = clock64();
    cluster_sync_all();  // all 8 heads of the sample have written their slice of the attention output
    if (prof) a.prof[blk * 8 + 4] = clock64();

    // ------------------------------------------------------------ phase 3: y_h = o Wo^T + bo + x  (14 column tiles)
    load_rows_to_A(sA, a.ao + row0 * D, T, tid);
    __syncthreads();
    {
      const __nv_bfloat16* wrow[NT_OUT];
      int nt = 0;
#pragma unroll
      for (int i = 0; i < NT_OUT; ++i) {
        const int tile = warp + i * PF_WARPS;
        wrow[i] = w.wo;
        if (tile < 14) {
          nt = i + 1;
          wrow[i] = w.wo + static_cast<long long>(h * HD + tile * 8 + g) * D;
        }
      }
      float acc[NT_OUT][4];
      if (NT_OUT == 1 || nt == 2) {
        warp_gemm_16xK<NT_OUT, UNROLL_OUT>(sA, wrow, nt, lane, acc);
      } else {  // one tile: its own instantiation keeps the 14 loads of the two-tile warps out of this warp's registers
        const __nv_bfloat16* w1[1] = {wrow[0]};
        float acc1[1][4];
        warp_gemm_16xK<1, UNROLL_OUT>(sA, w1, nt, lane, acc1);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[0][e] = acc1[0][e];
      }
#pragma unroll
      for (int i = 0; i < NT_OUT; ++i) {
        if (i >= nt) continue;
        const int col = h * HD + (warp + i * PF_WARPS) * 8 + 2 * t;
        const float b0 = __ldg(w.bo + col), b1 = __ldg(w.bo + col + 1);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int r = g + half * 8;
          if (r >= T) continue;
          const float2 xr = unpack_bf16(__ldcg(reinterpret_cast<const unsigned int*>(x_in + static_cast<long long>(r) * D + col)));
          *reinterpret_cast<uint32_t*>(a.y + (row0 + r) * D + col) =
              pack_bf16(acc[i][2 * half] + b0 + xr.x, acc[i][2 * half + 1] + b1 + xr.y);
        }
      }
    }
    if (prof) a.prof[blk * 8 + 5] = clock64();
    cluster_sync_all();  // y complete

    // ------------------------------------------------------------ phase 4: LayerNorm(y) -> A operand (bf16); one warp per row
    for (int r = warp; r < T; r += PF_WARPS) {
      const __nv_bfloat16* yr = a.y + (row0 + r) * D;
      float v[4][8];
      // single pass around a pivot (the row's first element), like norm_kernel in ops.cu
      const float pivot = unpack_bf16(__ldcg(reinterpret_cast<const unsigned int*>(yr))).x;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int vi = lane + i * 32;
        if (vi < D / 8) {
          const uint4 u = __ldcg(reinterpret_cast<const uint4*>(yr + vi * 8));
          const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
          v[i][0] = a0.x; v[i][1] = a0.y; v[i][2] = a1.x; v[i][3] = a1.y;
          v[i][4] = a2.x; v[i][5] = a2.y; v[i][6] = a3.x; v[i][7] = a3.y;
#pragma unroll
          for (int j = 0; j < 8; j += 2) {  // pairwise, in the order of norm_kernel
            const float d0 = v[i][j] - pivot, d1 = v[i][j + 1] - pivot;
            s1 += d0 + d1;
            s2 += d0 * d0 + d1 * d1;
          }
        }
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      const float m1 = s1 / D, mean = pivot + m1;
      const float rstd = rsqrtf(fmaxf(s2 / D - m1 * m1, 0.f) + a.ln_eps);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int vi = lane + i * 32;
        if (vi < D / 8) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(w.lnw + vi * 8)), w1 = __ldg(reinterpret_cast<const float4*>(w.lnw + vi * 8 + 4));
          const float4 c0 = __ldg(reinterpret_cast<const float4*>(w.lnb + vi * 8)), c1 = __ldg(reinterpret_cast<const float4*>(w.lnb + vi * 8 + 4));
          const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          const float bb[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
          float o8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] = (v[i][j] - mean) * rstd * ww[j] + bb[j];
          *reinterpret_cast<uint4*>(sA + r * LDA + vi * 8) =
              make_uint4(pack_bf16(o8[0], o8[1]), pack_bf16(o8[2], o8[3]), pack_bf16(o8[4], o8[5]), pack_bf16(o8[6], o8[7]));
        }
      }
    }
    __syncthreads();
    if (prof) a.prof[blk * 8 + 6] = clock64();

    // ------------------------------------------------------------ phase 5: x'_h = ReLU(LN(y) Wffn^T + b)
    {
      const __nv_bfloat16* wrow[NT_OUT];
      int nt = 0;
#pragma unroll
      for (int i = 0; i < NT_OUT; ++i) {
        const int tile = warp + i * PF_WARPS;
        wrow[i] = w.wffn;
        if (tile < 14) {
          nt = i + 1;
          wrow[i] = w.wffn + static_cast<long long>(h * HD + tile * 8 + g) * D;
        }
      }
      float acc[NT_OUT][4];
      if (NT_OUT == 1 || nt == 2) {
        warp_gemm_16xK<NT_OUT, UNROLL_OUT>(sA, wrow, nt, lane, acc);
      } else {  // one tile: its own instantiation keeps the 14 loads of the two-tile warps out of this warp's registers
        const __nv_bfloat16* w1[1] = {wrow[0]};
        float acc1[1][4];
        warp_gemm_16xK<1, UNROLL_OUT>(sA, w1, nt, lane, acc1);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[0][e] = acc1[0][e];
      }
#pragma unroll
      for (int i = 0; i < NT_OUT; ++i) {
        if (i >= nt) continue;
        const int col = h * HD + (warp + i * PF_WARPS) * 8 + 2 * t;
        const float b0 = __ldg(w.bffn + col), b1 = __ldg(w.bffn + col + 1);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int r = g + half * 8;
          if (r >= T) continue;
          *reinterpret_cast<uint32_t*>(w.x_out + (row0 + r) * D + col) =
              pack_bf16(fmaxf(acc[i][2 * half] + b0, 0.f), fmaxf(acc[i][2 * half + 1] + b1, 0.f));
        }
      }
    }
    if (prof) a.prof[blk * 8 + 7] = clock64();
    cluster_sync_all();  // x' complete: it is the next block's x
    x_in = w.x_out + row0 * D;
    load_rows_to_A(sA, x_in, T, tid);
    __syncthreads();
  }
  if (blockIdx.x == 0 && tid == 0) *reinterpret_cast<volatile int*>(a.progress) = 0;  // the next launch starts from 0
}

}  // namespace

size_t policy_fused_smem_bytes() {
  constexpr size_t a_bytes = 16 * LDA * 2, kv_bytes = static_cast<size_t>(PF_WARPS) * 4 * SKB * LDQ * 2;
  return 16 * LDQ * 2 + (a_bytes > kv_bytes ? a_bytes : kv_bytes);
}

int policy_fused_launch(const PolicyFusedArgs& a, int B, cudaStream_t s, const char** err, int prefetch_clusters) {
  if (a.T < 1 || a.T > 16 || B < 1 || B != a.B || !a.progress || a.n_blocks < 1 || a.n_blocks > POLICY_FUSED_MAX_BLOCKS) {
    if (err) *err = "policy_fused: chunk length must be 1..16 and the batch positive";
    return -1;
  }
  const size_t smem = policy_fused_smem_bytes();
  static PerDeviceFlag attr_flag;
  bool& attr_set = attr_flag.here();
  if (!attr_set) {
    if (cudaFuncSetAttribute(policy_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
      if (err) *err = "policy_fused: cudaFuncSetAttribute failed";
      return -4;
    }
    attr_set = true;
  }
  const int pf_cl = prefetch_clusters < 0 ? PF_PREFETCH_CL : (prefetch_clusters > PF_PREFETCH_CL ? PF_PREFETCH_CL : prefetch_clusters);
  launch_kernel(policy_fused_kernel, dim3((B + pf_cl) * PF_CL), dim3(PF_THREADS), smem, s, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = cudaGetErrorString(e);
    return -4;
  }
  ops_count_launch();
  return 0;
}

}  // namespace vla
