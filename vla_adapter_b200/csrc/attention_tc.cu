// softmax(Q K^T / sqrt(hd)) V on tcgen05 tensor cores with TMEM accumulators, operands fed by TMA.
//
// Serves the three big attention shapes of the path (head dims 64 and 72; the policy's 8-query
// cross-attention stays on the small mma.sync kernel of attention.cu):
//   DINOv2  : 16 heads x 64, S = 261, bidirectional   (timm Attention restated in film_vit_wrapper.py:69)
//   SigLIP  : 16 heads x 72, S = 256, bidirectional
//   Qwen2.5 : 14 q / 2 kv heads x 64, S ~ 625, causal (modeling_prismatic.py:834-845 -> Qwen2Attention)
//
// Persistent kernel, one CTA (640 threads) per SM.  A work item is a PAIR of 128-row query tiles ("slots" A and B) that
// share one K/V stream: two query tiles of one head (ViT) or two query heads of one kv group (Qwen GQA), so every K/V
// tile is fetched once for 256 query rows.  Roles (setmaxnreg moves registers from warpgroup 0 to the softmax warps):
//   warp 0      : TMA producer - Q tiles of both slots (double-buffered), K/V tiles of 128 keys through an mbarrier
//                 ring that runs ahead across work items (the next item's operands land while this one computes)
//   warps 1, 2  : MMA issuers, one per slot - S = Q K^T (SS, 128 x keys x hd) into TMEM, then O += P V (TS: P is read
//                 from TMEM, V is the MN-major smem operand exactly as TMA delivers it); tcgen05.commit -> mbarriers
//   warps 4-11  : softmax + epilogue of slot A, warps 12-19 of slot B.  The two warps of one TMEM lane quarter own
//                 the same 32 query rows (one thread per row) and split the tile's 128 keys; row max and row sum are
//                 exchanged through shared memory.  tcgen05.ld the fp32 scores, mask, running max with lazy
//                 rescaling of O (only when the max grows by > 2^8), exp2 (packed FFMA2 / FADD2), P -> bf16 ->
//                 tcgen05.st; finally O / l -> bf16 -> global.
// TMEM columns, head dim 64: S_A S_B | P_A P_B | O_A O_B (P outside the score columns: the next tile's Q K^T is issued
// as soon as the scores have been read, and P V runs beside the next tile's softmax).  Head dim 72 (2 x 80 accumulator
// columns) keeps P aliased over S: S_A S_B | O_A O_B.
//
// Head dim 72 is handled as a 64-wide main block (128B-swizzled tiles) plus a 16-wide tail block
// (32B-swizzled tiles) whose columns 72..79 are zero-filled by TMA: the tensor maps are per-head 4-D views
// (d, head, row, sample) so that out-of-head columns count as out of bounds.
// An experimental configuration with 64-key tiles and two CTAs per SM is kept behind VLA_FA_BN64=1 (measured a wash).
#include "common.cuh"
#include "launch.cuh"
#include "ops.cuh"
#include "watchdog.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace vla {

namespace {

constexpr int FA_BM = 128;
constexpr uint32_t FA_TILE_BYTES = 128 * 128;  // Q tile: 128 rows x 64 bf16
constexpr uint32_t FA_TAIL_BYTES = 128 * 32;   // Q tail: 128 rows x 16 bf16
constexpr int FA_MAX_ITEMS = 40;               // work items per (sample, kv head): 3 images, L <= 127 -> 8 x 7 units = 28 items

struct FaDev {
  int Sq, Skv, group, kv_heads, causal;
  int q_rows;  // rows between two samples of Q / out (>= Sq: the caller may hand a row range of each sample)
  int bn;       // keys per tile (64: two CTAs per SM, 128: one)
  int n_bg;     // samples * kv heads
  int n_items;  // n_bg * items per (sample, kv head)
  float scale_log2;
  __nv_bfloat16* out;
  int ld_out;
  unsigned int* trace;  // timing experiments only (VLA_FA_TRACE): [0] = count, then (event, clock) pairs of CTA 0
  // timing experiments only (VLA_FA_DEBUG bit mask): 1 = softmax warps only move the barriers, 2 = no PV MMAs,
  // 4 = no QK MMAs, 8 = no exp-phase turn taking, 16 = no early QK issue, 32 = no MUFU work, 64 = no TMEM loads,
  // 128 = no P stores, 256 = no row-max exchange, 512 = no output stores, 1024 = no O loads.  Findings (Qwen shape,
  // 140 us): removing the exponentials, the TMEM traffic and the exchange changes nothing (<= 3 %); the output stores
  // cost ~10 %; bit 1 alone (no softmax code at all) gives 59 us - the kernel is bound by the latency of the per-tile
  // dependency chain through the softmax warps, not by any one unit.
  int debug;
  // per (sample, kv head): query head (relative to the group) and query tile of slot A / slot B, packed
  // hA | qA << 8 | hB << 16 | qB << 24; hB == 0xff: slot B idle
  uint32_t items[FA_MAX_ITEMS];
};

// The timing-experiment switches (FaDev::debug) and the event trace exist in the trace build only
// (-DVLA_FA_TRACE_BUILD); in the product build every switch is a compile-time zero and its branch disappears.
#ifdef VLA_FA_TRACE_BUILD
#define FA_DBG(p, mask) ((p).debug & (mask))
#else
#define FA_DBG(p, mask) 0
#endif
// Exp-phase turn taking between the two slots (A(j) -> B(j) -> A(j+1) ...: one slot owns the MUFU while the other's
// P -> PV -> next QK chain runs) is OFF: measured on B200 at bs=64 it is worth +4 % on the DINOv2 shape, +5 % on SigLIP
// and -1.5 % on Qwen - 9.30 ms against 9.24 ms of attention per step, nothing - and its original form was the source of
// an intermittent deadlock (see do_turn below).  -DVLA_FA_TURNS=1 brings the (repaired) protocol back.
#if defined(VLA_FA_TURNS) && VLA_FA_TURNS
constexpr bool FA_TURNS = true;
#else
constexpr bool FA_TURNS = false;
#endif

struct FaItem {
  int b, g;
  int h[2], qb[2], n[2];  // per slot: absolute query head, query tile, number of key tiles (0 = idle)
  int nmax;
};

// Debug event log of CTA 0 (VLA_FA_TRACE): four single-writer regions (producer, MMA, softmax A, softmax B) of 4000
// (event, clock) pairs each, written fire-and-forget so that tracing barely perturbs the timeline.
// Compiled in only with -DVLA_FA_TRACE_BUILD (the bookkeeping costs registers in the softmax warps).
VLA_DEVINL void fa_trace(const FaDev& p, int role, unsigned int& cnt, unsigned int ev) {
#ifdef VLA_FA_TRACE_BUILD
  if (p.trace && blockIdx.x == 0 && cnt < 4000) {
    unsigned int* dst = p.trace + 4 + role * 8000 + 2 * cnt;
    dst[0] = ev;
    dst[1] = static_cast<unsigned int>(clock64());
    ++cnt;
    p.trace[role] = cnt;
  }
#endif
}

VLA_DEVINL FaItem fa_decode(const FaDev& p, int item) {
  FaItem it;
  const int r = item / p.n_bg, bg = item - r * p.n_bg;
  it.b = bg / p.kv_heads;
  it.g = bg - it.b * p.kv_heads;
  const uint32_t w = p.items[r];
  const int n_all = (p.Skv + p.bn - 1) / p.bn;
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    const int hr = (w >> (16 * x)) & 0xff, q = (w >> (16 * x + 8)) & 0xff;
    it.h[x] = it.g * p.group + hr;
    it.qb[x] = q;
    const int n_c = ((q + 1) * FA_BM + p.bn - 1) / p.bn;  // causal: key tiles up to this query tile's last row
    it.n[x] = hr == 0xff ? 0 : ((p.causal && n_c < n_all) ? n_c : n_all);
  }
  it.nmax = it.n[0] > it.n[1] ? it.n[0] : it.n[1];
  return it;
}

VLA_DEVINL void tma_load_4d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// Generic shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1).
VLA_DEVINL uint64_t make_smem_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;  // LBO: unused by every layout in this kernel (single atom along MN / K)
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}
constexpr uint32_t LAYOUT_SW128 = 2, LAYOUT_SW32 = 6;

// D[tmem] (+)= A[tmem] * B[smem]
VLA_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

VLA_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
VLA_DEVINL void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
VLA_DEVINL void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
VLA_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

VLA_DEVINL float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

VLA_DEVINL void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Softmax of this thread's query row over ITS HALF of one 128-key score tile (NLIVE live chunks of 32 keys; the
// partner warp of the same TMEM lane quarter owns the other 64 keys).  Scores from TMEM, mask, row max exchanged
// with the partner through shared memory, running max with lazy O rescale, exp2, partial row sum, P (bf16) back
// over the score columns.
template <int HD, int NLIVE, bool MASKED, bool SPLIT>
VLA_DEVINL void fa_softmax_tile(const FaDev& p, uint32_t tSh, uint32_t tPh, uint32_t tO, int k0h, int grow, int j, int half,
                                float sl2, float& m_ref, float& l, float* xch_mine, const float* xch_other, int bar_id,
                                uint32_t turn_wait, uint32_t turn_parity, uint32_t turn_arrive, uint32_t s_free_bar,
                                uint32_t pv_done_bar, uint32_t pv_parity) {
  uint32_t v[NLIVE > 0 ? NLIVE : 1][32];
  float mloc = -INFINITY;
  if (NLIVE > 0) {
    if (FA_DBG(p, 64)) {  // timing experiment: no TMEM loads
#pragma unroll
      for (int c = 0; c < NLIVE; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) v[c][i] = 0x3f000000u + i;
    } else {
#pragma unroll
      for (int c = 0; c < NLIVE; ++c) tmem_ld_32x32b_x32(tSh + c * 32, v[c]);
      tmem_ld_wait();
    }
  }
  if (s_free_bar) {  // the scores are in registers: the MMA warp may overwrite S with the next tile's Q K^T
    tc_fence_before();
    mbar_arrive(s_free_bar);
  }
  if (NLIVE > 0) {
    if (MASKED) {  // the half touches the causal diagonal or the end of the keys: straight-line select on every element
      int last = p.Skv - 1;                        // last key this row may attend to ...
      if (p.causal && grow < last) last = grow;
      last -= k0h;                                 // ... relative to this half tile
#pragma unroll
      for (int c = 0; c < NLIVE; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) v[c][i] = (c * 32 + i > last) ? 0xff800000u : v[c][i];  // -inf
    }
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int c = 0; c < NLIVE; ++c) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(v[c][i]));
        mx1 = fmaxf(mx1, __uint_as_float(v[c][i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(v[c][i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(v[c][i + 3]));
      }
    }
    mloc = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
  }
  // row max over the whole tile: one float per row each way, one 64-thread named barrier
  float m_new = fmaxf(m_ref, mloc);
  if (SPLIT && !(FA_DBG(p, 256))) {
    *xch_mine = mloc;
    named_bar_sync(bar_id, 64);
    m_new = fmaxf(m_new, *xch_other);
  }
  if (pv_done_bar) {  // PV of the previous tile has retired: O is stable and the P columns may be rewritten
    mbar_wait(pv_done_bar, pv_parity);
    tc_fence_after();
  }
  if (j == 0) {
    m_ref = m_new;
  } else {
    // Lazy rescale: the reference max only moves when it would otherwise let exp2 exceed 2^8.  Both warps of the
    // quarter see the same m_new / m_ref, so they take this branch together; each rescales its half of O's columns.
    const bool need = (m_new - m_ref) * sl2 > 8.0f;
    if (__any_sync(0xffffffffu, need)) {
      const float alpha = need ? ex2f((m_ref - m_new) * sl2) : 1.0f;
      if (need) m_ref = m_new;
      l *= alpha;
      // s_full(j) was committed after PV(j-1), so O is stable here and PV(j) has not been issued yet.
      constexpr int G = (HD + 7) / 8;
#pragma unroll 1
      for (int c = (SPLIT && half) ? G / 2 : 0; c < ((SPLIT && !half) ? G / 2 : G); ++c) {
        uint32_t o[8];
        tmem_ld_32x32b_x8(tO + c * 8, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st_32x32b_x8(tO + c * 8, o);
      }
    }
  }
  // The exponentials of the two slots take turns: while one slot owns the MUFU, the other's P -> PV -> next QK ->
  // TMEM load -> row max chain runs on the tensor pipe, instead of both slots doing each phase in lockstep.
  if (turn_wait) {
    mbar_wait(turn_wait, turn_parity);
  }
  if (NLIVE > 0) {
    const float mb = m_ref * sl2;
    const float2 sl2v = make_float2(sl2, sl2), nmb = make_float2(-mb, -mb);
    float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NLIVE; ++c) {
      // phase 1: all 32 exponentials of the chunk in flight (packed FFMA2 for the scale/shift)
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[c][2 * i]), __uint_as_float(v[c][2 * i + 1])), sl2v, nmb);
        if (FA_DBG(p, 32)) {  // timing experiment: no MUFU work
          v[c][2 * i] = __float_as_uint(t.x);
          v[c][2 * i + 1] = __float_as_uint(t.y);
        } else {
          v[c][2 * i] = __float_as_uint(ex2f(t.x));
          v[c][2 * i + 1] = __float_as_uint(ex2f(t.y));
        }
      }
      // phase 2: row sum (packed FADD2) and bf16 packing
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 e = make_float2(__uint_as_float(v[c][2 * i]), __uint_as_float(v[c][2 * i + 1]));
        sum2 = __fadd2_rn(sum2, e);
        pk[i] = pack_bf16(e.x, e.y);
      }
      if (!(FA_DBG(p, 128))) tmem_st_32x32b_x16(tPh + c * 16, pk);
    }
    l += sum2.x + sum2.y;
  }
  if (turn_arrive) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(turn_arrive);
  }
}

template <int HD, int BN>
struct FaSmem {
  static constexpr bool TAIL = HD > 64;
  // BN = 128: one CTA per SM, 8 softmax warps per slot (two per TMEM lane quarter split the 128 keys).
  // BN = 64 : two CTAs per SM (four slots, four independent QK -> softmax -> PV chains per SM), 4 softmax warps
  //           per slot, 256 TMEM columns per CTA with P aliased over S.
  static constexpr bool SPLIT = BN == 128;
  static constexpr int CTAS_PER_SM = SPLIT ? 1 : 2;
  static constexpr int SM_WARPS = SPLIT ? 8 : 4;             // softmax warps per slot
  static constexpr int THREADS = 128 + 2 * SM_WARPS * 32;   // warpgroup 0 (TMA, MMA, 2 idle warps) + softmax warps
  static constexpr int SM_THREADS = SM_WARPS * 32;
  static constexpr int STAGES = TAIL ? 2 : 3;  // head dim 72: two 40 KB stages leave room for the epilogue staging
  static constexpr int QBUFS = 4;  // 2 slots x 2 buffers: the next item's Q tiles land while this item computes
  static constexpr uint32_t KV_TILE = BN * 128;  // BN keys x 64 bf16
  static constexpr uint32_t KV_TAIL = BN * 32;   // BN keys x 16 bf16
  static constexpr uint32_t Q_BYTES = FA_TILE_BYTES + (TAIL ? FA_TAIL_BYTES : 0);
  static constexpr uint32_t KV_BYTES = 2 * KV_TILE + (TAIL ? 2 * KV_TAIL : 0);
  // main tiles first (1024-byte aligned), then the 32B-swizzled tails, then barriers
  static constexpr uint32_t OFF_Q = 0;                               // QBUFS
  static constexpr uint32_t OFF_K = QBUFS * FA_TILE_BYTES;           // STAGES
  static constexpr uint32_t OFF_V = OFF_K + STAGES * KV_TILE;        // STAGES
  static constexpr uint32_t OFF_QT = OFF_V + STAGES * KV_TILE;
  static constexpr uint32_t OFF_KT = OFF_QT + QBUFS * FA_TAIL_BYTES;
  static constexpr uint32_t OFF_VT = OFF_KT + STAGES * KV_TAIL;
  static constexpr uint32_t OFF_BAR = TAIL ? OFF_VT + STAGES * KV_TAIL : OFF_QT;
  static constexpr uint32_t OFF_XCH = OFF_BAR + 384;  // row-max / row-sum exchange: [slot][parity][half][128] floats
  static constexpr uint32_t XCH_END = OFF_XCH + (SPLIT ? 2 * 2 * 2 * 128 * 4 : 0);
  // Head dim 64, 128-key tiles: the epilogue leaves through shared memory and one TMA store per warp (32 rows x 32
  // columns, 64B-swizzled, 2 KB per softmax warp) - the per-thread 16-byte row stores took ~3k cycles per item during
  // which the slot's next item could not start.  Head dim 72: columns [0, 64) leave that way, the warp of the
  // second half stores the last 8 columns itself.
  static constexpr bool TMA_EPI = SPLIT;
  static constexpr uint32_t OFF_STG = (XCH_END + 1023u) & ~1023u;
  static constexpr uint32_t TOTAL = TMA_EPI ? OFF_STG + 2 * SM_WARPS * 2048 : XCH_END;  // dynamic array is __align__(1024)
  // TMEM columns: S_A [0,BN), S_B [BN,2BN) (P aliases the first half of S), then O_A, O_B
  // Head dim 64 has room for P OUTSIDE the score columns (S_A S_B | P_A P_B | O_A O_B = 256 + 128 + 128): the next
  // tile's QK^T is issued as soon as the softmax warps have read the scores, and PV of this tile runs beside the
  // next tile's softmax.  Head dim 72 (2 x 80 accumulator columns) keeps P aliased over S.
  static constexpr bool DEALIAS = !TAIL && SPLIT;
  static constexpr uint32_t P_OFF = DEALIAS ? 2 * BN : 0;           // + P_STRIDE * slot
  static constexpr uint32_t P_STRIDE = DEALIAS ? 64 : BN;
  static constexpr uint32_t O_OFF = DEALIAS ? 2 * BN + 128 : 2 * BN;
  static constexpr uint32_t O_STRIDE = TAIL ? 128 : 64;
  static_assert(SPLIT || !TAIL, "64-key tiles: head dim 64 only (2 x 64 S + 2 x 64 O = 256 TMEM columns)");
  static constexpr uint32_t TMEM_COLS = SPLIT ? 512 : 256;
};

template <int HD, int BN>
__global__ void __launch_bounds__((FaSmem<HD, BN>::THREADS), (FaSmem<HD, BN>::CTAS_PER_SM))
fa_tcgen05_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                  const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapQt,
                  const __grid_constant__ CUtensorMap mapKt, const __grid_constant__ CUtensorMap mapVt,
                  const __grid_constant__ CUtensorMap mapO, const __grid_constant__ FaDev p) {
  using L = FaSmem<HD, BN>;
  constexpr bool TAIL = L::TAIL;
  constexpr int NS = L::STAGES;

  extern __shared__ __align__(1024) uint8_t fa_smem_raw[];
  const uint32_t raw_addr = smem_u32(fa_smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = fa_smem_raw + pad;
  const uint32_t sbase = raw_addr + pad;
  const uint32_t bar_base = sbase + L::OFF_BAR;
  auto q_full = [&](int qb) { return bar_base + 8u * qb; };          // qb = slot * 2 + buffer
  auto q_empty = [&](int qb) { return bar_base + 8u * (4 + qb); };
  auto s_full = [&](int x) { return bar_base + 8u * (8 + x); };
  auto p_ready = [&](int x) { return bar_base + 8u * (10 + x); };
  auto o_full = [&](int x) { return bar_base + 8u * (12 + x); };
  auto o_empty = [&](int x) { return bar_base + 8u * (14 + x); };
  auto turn = [&](int x) { return bar_base + 8u * (16 + x); };        // exp-phase turn taking between the slots
  auto s_free = [&](int x) { return bar_base + 8u * (18 + x); };      // scores read into registers (DEALIAS)
  auto pv_done = [&](int x) { return bar_base + 8u * (20 + x); };     // PV of the previous tile retired (DEALIAS)
  auto kv_full = [&](int s) { return bar_base + 8u * (22 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (22 + NS + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::OFF_BAR + 8 * (22 + 2 * NS));
  // watchdog (common.cuh): warp 3 is the monitor, parked on done_bar; every other warp arrives there at its end
  constexpr int N_WARPS = L::THREADS / 32;
  const uint32_t done_bar = bar_base + 8u * (23 + 2 * NS);

  // shfl-broadcast warp index: the role branches are then provably warp-uniform, so the MMA warp's descriptor
  // arithmetic stays on the uniform datapath and tcgen05.mma issues at the hardware rate (scripts/ubench/mma3.cu)
  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    for (int qb = 0; qb < 4; ++qb) {
      mbar_init(q_full(qb), 1);
      mbar_init(q_empty(qb), 1);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(turn(x), L::SM_WARPS);
      mbar_init(s_free(x), L::SM_THREADS);
      mbar_init(pv_done(x), 1);
      mbar_init(s_full(x), 1);
      mbar_init(p_ready(x), L::SM_THREADS);
      mbar_init(o_full(x), 1);
      mbar_init(o_empty(x), L::SM_THREADS);
    }
    for (int s = 0; s < NS; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 2);  // one release per slot's MMA warp
    }
    mbar_init(done_bar, N_WARPS - 1);
    mbar_fence_init();
    fence_proxy_async();
  }
  if (warp_idx == 1) {
    tmem_alloc(smem_u32(tmem_slot), L::TMEM_COLS);
    tmem_relinquish();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  // Everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel's tail.
  pdl_wait();
  pdl_launch_dependents();

  if (warp_idx < 4) {
  // register budget: SPLIT 128*32 + 512*112 == 640*96, else 128*32 + 256*104 == 384*80 (the launch allocations)
  asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  if (warp_idx == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp waits, one lane issues)
    uint32_t kv_cnt = 0, q_cnt[2] = {0, 0};
    unsigned int tr_cnt = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const FaItem it = fa_decode(p, item);
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (!it.n[x]) continue;
        const int qb = x * 2 + static_cast<int>(q_cnt[x] & 1u);
        mbar_wait_relaxed(q_empty(qb), ((q_cnt[x] >> 1) & 1u) ^ 1u);
        ++q_cnt[x];
        if (elect_one()) {
          mbar_arrive_expect_tx(q_full(qb), L::Q_BYTES);
          tma_load_4d(sbase + L::OFF_Q + qb * FA_TILE_BYTES, &mapQ, q_full(qb), 0, it.h[x], it.qb[x] * FA_BM, it.b);
          if (TAIL)
            tma_load_4d(sbase + L::OFF_QT + qb * FA_TAIL_BYTES, &mapQt, q_full(qb), 64, it.h[x], it.qb[x] * FA_BM, it.b);
        }
        __syncwarp();
      }
      for (int j = 0; j < it.nmax; ++j) {
        const int s = static_cast<int>(kv_cnt % NS);
        const uint32_t ph = (kv_cnt / NS) & 1u;
        ++kv_cnt;
        mbar_wait_relaxed(kv_empty(s), ph ^ 1u);
        if (elect_one()) {
          fa_trace(p, 0, tr_cnt, 100 + j);
          mbar_arrive_expect_tx(kv_full(s), L::KV_BYTES);
          tma_load_4d(sbase + L::OFF_K + s * L::KV_TILE, &mapK, kv_full(s), 0, it.g, j * BN, it.b);
          tma_load_4d(sbase + L::OFF_V + s * L::KV_TILE, &mapV, kv_full(s), 0, it.g, j * BN, it.b);
          if (TAIL) {
            tma_load_4d(sbase + L::OFF_KT + s * L::KV_TAIL, &mapKt, kv_full(s), 64, it.g, j * BN, it.b);
            tma_load_4d(sbase + L::OFF_VT + s * L::KV_TAIL, &mapVt, kv_full(s), 64, it.g, j * BN, it.b);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp_idx == 1 || warp_idx == 2) {
    // ------------------------------------------------------------ MMA issuers: one warp per slot (whole warp waits,
    // one lane issues).  A single issuer serialises the waits of both slots' QK -> softmax -> PV chains - its
    // wait / issue / commit loop alone was the pacing stage of the kernel's skeleton.
    const int x = warp_idx - 1;
    uint32_t kv_cnt = 0, q_cnt = 0, p_cnt = 0, o_cnt = 0, f_cnt = 0;
    unsigned int tr_cnt = 0;
    const uint32_t idesc_pv = make_idesc_bf16(128, 64) | (1u << 16);
    const uint32_t idesc_pvt = make_idesc_bf16(128, 16) | (1u << 16);
    const uint32_t tS = tmem_base + static_cast<uint32_t>(BN) * x;
    const uint32_t tP = tmem_base + L::P_OFF + L::P_STRIDE * x;  // P columns (over S unless de-aliased)
    const uint32_t tO = tmem_base + L::O_OFF + L::O_STRIDE * x;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const FaItem it = fa_decode(p, item);
      const int n_x = it.n[x];
      const uint32_t kv0 = kv_cnt;  // ring position of this item's key tile 0
      kv_cnt += static_cast<uint32_t>(it.nmax);
      int kv_waited = 0;            // key tiles [0, kv_waited) of this item are known to have landed
      auto stage_of = [&](int j) { return static_cast<int>((kv0 + j) % NS); };
      auto need_kv = [&](int j) {
        while (kv_waited <= j) {
          const uint32_t c = kv0 + kv_waited;
          mbar_wait(kv_full(static_cast<int>(c % NS)), (c / NS) & 1u);
          ++kv_waited;
        }
        tc_fence_after();
      };
      auto n16_of = [&](int j) {
        int nvalid = p.Skv - j * BN;
        if (nvalid > BN) nvalid = BN;
        return (nvalid + 15) >> 4;  // key columns actually computed, in units of 16
      };
      int qbuf = 0;
      // S_x = Q_x K_j^T   (M = 128 query rows, N = n16*16 keys, K = head dim)
      auto issue_qk = [&](int j) {
        need_kv(j);
        const int s = stage_of(j);
        const uint32_t idesc_qk = make_idesc_bf16(128, static_cast<uint32_t>(n16_of(j) * 16));
        const uint32_t sq = sbase + L::OFF_Q + qbuf * FA_TILE_BYTES, sk = sbase + L::OFF_K + s * L::KV_TILE;
        if (elect_one()) {
          if (!(FA_DBG(p, 4))) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tS, make_smem_desc(sq + k * 32, 1024, LAYOUT_SW128),
                        make_smem_desc(sk + k * 32, 1024, LAYOUT_SW128), idesc_qk, k != 0 ? 1u : 0u);
            if (TAIL)
              umma_bf16(tS, make_smem_desc(sbase + L::OFF_QT + qbuf * FA_TAIL_BYTES, 256, LAYOUT_SW32),
                        make_smem_desc(sbase + L::OFF_KT + s * L::KV_TAIL, 256, LAYOUT_SW32), idesc_qk, 1u);
          }
          umma_commit(s_full(x));
          if (j == n_x - 1) umma_commit(q_empty(qbuf));  // last use of this Q tile
        }
        __syncwarp();
      };
      // O_x (+)= P_x V_j   (A = P from TMEM, B = V MN-major: K = keys, N = head dim)
      auto issue_pv = [&](int j) {
        const int s = stage_of(j);
        const int n16 = n16_of(j);
        const uint32_t sv = sbase + L::OFF_V + s * L::KV_TILE;
        if (elect_one()) {
          if (!(FA_DBG(p, 2))) {
            for (int kk = 0; kk < n16; ++kk)
              umma_bf16_ts(tO, tP + kk * 8, make_smem_desc(sv + kk * 2048, 1024, LAYOUT_SW128), idesc_pv,
                           (j | kk) != 0 ? 1u : 0u);
            if (TAIL) {
              const uint32_t svt = sbase + L::OFF_VT + s * L::KV_TAIL;
              for (int kk = 0; kk < n16; ++kk)
                umma_bf16_ts(tO + 64, tP + kk * 8, make_smem_desc(svt + kk * 512, 256, LAYOUT_SW32), idesc_pvt,
                             (j | kk) != 0 ? 1u : 0u);
            }
          }
        }
        __syncwarp();
      };
      if (n_x) {
        qbuf = x * 2 + static_cast<int>(q_cnt & 1u);
        mbar_wait(q_full(qbuf), (q_cnt >> 1) & 1u);
        ++q_cnt;
        tc_fence_after();
        issue_qk(0);
      }
      for (int j = 0; j < it.nmax; ++j) {
        if (j < n_x) {
          if (L::DEALIAS) {
            // the next tile's scores as soon as this tile's have been read: S(j+1) is ready before softmax(j) ends
            mbar_wait(s_free(x), f_cnt & 1u);
            ++f_cnt;
            tc_fence_after();
            if (j + 1 < n_x && !(FA_DBG(p, 16))) issue_qk(j + 1);
          }
          mbar_wait(p_ready(x), p_cnt & 1u);
          if (lane == 0) fa_trace(p, 1, tr_cnt, 300 + x * 10 + j);
          ++p_cnt;
          if (j == 0) {  // previous item's epilogue has drained O_x
            mbar_wait(o_empty(x), (o_cnt & 1u) ^ 1u);
          }
          tc_fence_after();
          issue_pv(j);
          if (j + 1 < n_x) {
            if (L::DEALIAS) {
              if (elect_one()) umma_commit(pv_done(x));  // P_x / O_x may be touched again by the softmax warps
              __syncwarp();
              if (FA_DBG(p, 16)) issue_qk(j + 1);  // (experiment: no early issue)
            } else {
              issue_qk(j + 1);  // in-order after PV(j): S_x / P_x is free again
            }
          } else {
            if (elect_one()) umma_commit(o_full(x));
            __syncwarp();
            ++o_cnt;
          }
          // this slot's MMAs on key tile j have all been issued: its half of the release of the K/V ring slot
          if (elect_one()) umma_commit(kv_empty(stage_of(j)));
          __syncwarp();
        } else {
          // this slot is idle (or already past its causal extent) for key tile j: release its half right away - but
          // only once the tile has landed, so the arrival cannot fall into an earlier phase of the same slot
          need_kv(j);
          if (elect_one()) mbar_arrive(kv_empty(stage_of(j)));
          __syncwarp();
        }
      }
    }
  } else if (warp_idx == 3) {
#ifndef VLA_NO_WATCHDOG
    // ------------------------------------------------------------ watchdog monitor (see common.cuh)
    // minimal form: this warp lives in the 32-register warpgroup, and a barrier dump inlined here made ptxas spill
    // inside the MMA-issuing warps' loops (230 us instead of 134 us per DINOv2 layer)
    wd_monitor_min(done_bar, HD == 64 ? WD_K_FA64 : WD_K_FA72);
#endif
  }
  } else {
    if (L::SPLIT) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ------------------------------------------------------------ softmax + epilogue
    // 8 warps per slot: the two warps of one TMEM lane quarter own the same 32 query rows (one thread per row) and
    // split the tile's 128 keys (and later O's columns) in halves - four softmax warps per SM sub-partition keep the
    // MUFU busy where one warp per sub-partition reaches only ~57 % of its rate (scripts/ubench/expmix.cu).
    const int x = (warp_idx - 4) / L::SM_WARPS;   // slot
    const int quarter = warp_idx & 3;    // TMEM lane quarter this warp may touch
    const int half = L::SPLIT ? ((warp_idx - 4) >> 2) & 1 : 0;  // which 64 keys of a tile / which half of O's columns
    const int row = quarter * 32 + lane;
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(BN) * x;
    const uint32_t tO = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + L::O_OFF + L::O_STRIDE * x;
    const uint32_t tP = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + L::P_OFF + L::P_STRIDE * x;
    const uint32_t tSh = tS + 64u * half, tPh = tP + 32u * half;
    float* xch = reinterpret_cast<float*>(smem + L::OFF_XCH) + x * 512;  // [parity][half][128]
    const int bar_id = 1 + x * 4 + quarter;
    const float sl2 = p.scale_log2;
    uint32_t s_cnt = 0, o_cnt = 0, x_cnt = 0, t_cnt = 0, d_cnt = 0;
    unsigned int tr_cnt = 0;
    if (FA_TURNS && x == 1 && lane == 0) mbar_arrive(turn(0));  // slot A's first turn is pre-paid (see do_turn below)
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const FaItem it = fa_decode(p, item);
      const int n_it = it.n[x];
      if (!n_it) continue;
      const int q0 = it.qb[x] * FA_BM;
      const int grow = q0 + row;
      const int wrow0 = q0 + quarter * 32;            // first query row of this warp
      const bool warp_active = wrow0 < p.Sq && !(FA_DBG(p, 1));  // all-padding warps only keep the barriers moving
      float m_ref = -INFINITY, l = 0.f;
      // turn taking needs both slots busy with all their warps (padding-only warps would break the arrival counts)
      const bool pp = FA_TURNS && !(FA_DBG(p, 8)) && it.n[0] && it.n[1] && it.qb[0] * FA_BM + FA_BM <= p.Sq &&
                      it.qb[1] * FA_BM + FA_BM <= p.Sq && !(FA_DBG(p, 1));
      const int m_pp = it.n[0] < it.n[1] ? it.n[0] : it.n[1];
      for (int j = 0; j < n_it; ++j) {
        const int k0 = j * BN;
        int nvalid = p.Skv - k0;
        if (nvalid > BN) nvalid = BN;
        const int nch = (nvalid + 31) >> 5;           // 32-key chunks holding valid keys
        mbar_wait(s_full(x), s_cnt & 1u);
        ++s_cnt;
        tc_fence_after();
        if (quarter == 0 && half == 0 && lane == 0) fa_trace(p, 2 + x, tr_cnt, 400 + x * 10 + j);
        if (warp_active) {
          // causal: chunks entirely above the diagonal for every row of this warp carry no probability mass
          int nlive = nch;
          if (p.causal) {
            const int lim = (wrow0 + 31 - k0) / 32 + 1;  // chunks with a key <= the warp's last row
            nlive = lim < nch ? (lim < 0 ? 0 : lim) : nch;
          }
          // this warp's half: chunks [2*half, 2*half + 2) of the tile
          int my_live = nlive - 2 * half, my_all = nch - 2 * half;
          my_live = my_live < 0 ? 0 : (my_live > 2 ? 2 : my_live);
          my_all = my_all < 0 ? 0 : (my_all > 2 ? 2 : my_all);
          const int k0h = k0 + 64 * half;
          // warp-uniform: does a live chunk of this half touch the causal diagonal or the end of the keys?
          const bool masked = (k0h + my_live * 32 > p.Skv) || (p.causal && k0h + my_live * 32 - 1 > wrow0);
          float* xm = xch + (x_cnt & 1u) * 256 + half * 128 + row;
          const float* xo = xch + (x_cnt & 1u) * 256 + (half ^ 1) * 128 + row;
          ++x_cnt;
          // exp-phase turn taking while both slots are busy: A(j) -> B(j) -> A(j+1) ...  The protocol is symmetric:
          // each slot waits for the other's hand-over before EVERY common tile and hands over after it; slot B's
          // hand-over after the last common tile of an item is the one slot A collects at tile 0 of its next
          // turn-taking item (the very first one is pre-paid before the item loop).  Before this, A did not wait at
          // tile 0: it could finish its item and announce tile 0 of the next one while a stalled warp of B had not yet
          // observed A's previous announcement - the barrier was then two phases ahead of that warp's parity and the
          // two slots waited for each other forever (an intermittent hang on SM-contended multi-GPU runs).
          const bool do_turn = pp && j < m_pp;
          const uint32_t t_wait = do_turn ? turn(x) : 0u, t_par = t_cnt & 1u, t_arr = do_turn ? turn(x ^ 1) : 0u;
          if (do_turn) ++t_cnt;
          // de-aliased P: announce "scores read" after the TMEM loads, wait for PV(j-1) before touching O / P
          const uint32_t f_bar = L::DEALIAS ? s_free(x) : 0u;
          const uint32_t d_bar = (L::DEALIAS && j > 0) ? pv_done(x) : 0u, d_par = d_cnt & 1u;
          if (L::DEALIAS && j > 0) ++d_cnt;
          if (my_live == 2) {
            if (masked) fa_softmax_tile<HD, 2, true, L::SPLIT>(p, tSh, tPh, tO, k0h, grow, j, half, sl2, m_ref, l, xm, xo, bar_id, t_wait, t_par, t_arr, f_bar, d_bar, d_par);
            else fa_softmax_tile<HD, 2, false, L::SPLIT>(p, tSh, tPh, tO, k0h, grow, j, half, sl2, m_ref, l, xm, xo, bar_id, t_wait, t_par, t_arr, f_bar, d_bar, d_par);
          } else if (my_live == 1) {
            if (masked) fa_softmax_tile<HD, 1, true, L::SPLIT>(p, tSh, tPh, tO, k0h, grow, j, half, sl2, m_ref, l, xm, xo, bar_id, t_wait, t_par, t_arr, f_bar, d_bar, d_par);
            else fa_softmax_tile<HD, 1, false, L::SPLIT>(p, tSh, tPh, tO, k0h, grow, j, half, sl2, m_ref, l, xm, xo, bar_id, t_wait, t_par, t_arr, f_bar, d_bar, d_par);
          } else {
            fa_softmax_tile<HD, 0, false, L::SPLIT>(p, tSh, tPh, tO, k0h, grow, j, half, sl2, m_ref, l, xm, xo, bar_id, t_wait, t_par, t_arr, f_bar, d_bar, d_par);
          }
          if (my_live < my_all) {  // causal chunks above the diagonal: P = 0
            uint32_t z[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = 0u;
            for (int c = my_live; c < my_all; ++c) tmem_st_32x32b_x16(tPh + c * 16, z);
          }
          tmem_st_wait();
        } else if (L::DEALIAS) {
          // Padding-only warps keep the barriers moving, in step: with early QK issue s_full(j+1) can complete before
          // the working warps have finished tile j, and an arrival on p_ready for tile j+1 made before phase j
          // completes would be counted for phase j.  Waiting for PV(j-1) (issued after p_ready(j-1) completed)
          // pins the order.
          mbar_arrive(s_free(x));
          if (j > 0) {
            mbar_wait(pv_done(x), d_cnt & 1u);
            ++d_cnt;
          }
        }
        tc_fence_before();
        if (quarter == 0 && half == 0 && lane == 0) fa_trace(p, 2 + x, tr_cnt, 500 + x * 10 + j);
        mbar_arrive(p_ready(x));
      }
      // ---- epilogue: O / l -> bf16 -> global; each warp takes its half of the head's columns of its 32 rows
      float l_tot = l;
      if (warp_active && L::SPLIT) {  // total row sum = this warp's keys + the partner's
        float* xm = xch + (x_cnt & 1u) * 256 + half * 128 + row;
        const float* xo = xch + (x_cnt & 1u) * 256 + (half ^ 1) * 128 + row;
        ++x_cnt;
        *xm = l;
        named_bar_sync(bar_id, 64);
        l_tot = l + *xo;
      }
      mbar_wait_relaxed(o_full(x), o_cnt & 1u);
      ++o_cnt;
      tc_fence_after();
      if (quarter == 0 && half == 0 && lane == 0) fa_trace(p, 2 + x, tr_cnt, 600 + x);
      constexpr int G = (HD + 7) / 8;                  // 8-column groups of the head
      constexpr int G0 = L::SPLIT ? G / 2 : G;         // half 0: groups [0, G0), half 1: [G0, G)
      constexpr int GMAX = L::SPLIT ? G - G0 : G;
      uint32_t o[GMAX][8];
      const int g_lo = half ? G0 : 0, g_n = half ? G - G0 : G0;
      if (warp_active && !(FA_DBG(p, 1024))) {
#pragma unroll
        for (int c = 0; c < GMAX; ++c)
          if (c < g_n) tmem_ld_32x32b_x8(tO + (g_lo + c) * 8, o[c]);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(o_empty(x));  // O_x may be overwritten by the next item's first PV
      if constexpr (L::TMA_EPI) {
        if (warp_active && !(FA_DBG(p, 512))) {  // rows past Sq are clipped by the TMA unit
          const float inv = 1.0f / l_tot;
          const uint32_t stg = sbase + L::OFF_STG + static_cast<uint32_t>(warp_idx - 4) * 2048u;
          if (lane == 0) tma_store_wait_read<0>();  // the previous item's store has left this buffer
          __syncwarp();
          const uint32_t row_addr = stg + lane * 64;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t w0 = pack_bf16(__uint_as_float(o[c][0]) * inv, __uint_as_float(o[c][1]) * inv);
            const uint32_t w1 = pack_bf16(__uint_as_float(o[c][2]) * inv, __uint_as_float(o[c][3]) * inv);
            const uint32_t w2 = pack_bf16(__uint_as_float(o[c][4]) * inv, __uint_as_float(o[c][5]) * inv);
            const uint32_t w3 = pack_bf16(__uint_as_float(o[c][6]) * inv, __uint_as_float(o[c][7]) * inv);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((c ^ sw) << 4)), "r"(w0),
                         "r"(w1), "r"(w2), "r"(w3)
                         : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&mapO, stg, it.h[x] * HD + g_lo * 8, wrow0, it.b);
            tma_store_commit();
          }
          if constexpr (GMAX > 4) {  // head dim 72: columns [64, 72) of the row
            if (half && grow < p.Sq) {
              uint4 w;
              w.x = pack_bf16(__uint_as_float(o[4][0]) * inv, __uint_as_float(o[4][1]) * inv);
              w.y = pack_bf16(__uint_as_float(o[4][2]) * inv, __uint_as_float(o[4][3]) * inv);
              w.z = pack_bf16(__uint_as_float(o[4][4]) * inv, __uint_as_float(o[4][5]) * inv);
              w.w = pack_bf16(__uint_as_float(o[4][6]) * inv, __uint_as_float(o[4][7]) * inv);
              *reinterpret_cast<uint4*>(p.out + (static_cast<long long>(it.b) * p.q_rows + grow) * p.ld_out +
                                        it.h[x] * HD + 64) = w;
            }
          }
        }
      } else if (warp_active && grow < p.Sq && !(FA_DBG(p, 512))) {
        const float inv = 1.0f / l_tot;
        __nv_bfloat16* dst = p.out + (static_cast<long long>(it.b) * p.q_rows + grow) * p.ld_out + it.h[x] * HD + g_lo * 8;
#pragma unroll
        for (int c = 0; c < GMAX; ++c) {
          if (c < g_n) {
            uint4 w;
            w.x = pack_bf16(__uint_as_float(o[c][0]) * inv, __uint_as_float(o[c][1]) * inv);
            w.y = pack_bf16(__uint_as_float(o[c][2]) * inv, __uint_as_float(o[c][3]) * inv);
            w.z = pack_bf16(__uint_as_float(o[c][4]) * inv, __uint_as_float(o[c][5]) * inv);
            w.w = pack_bf16(__uint_as_float(o[c][6]) * inv, __uint_as_float(o[c][7]) * inv);
            *reinterpret_cast<uint4*>(dst + c * 8) = w;
          }
        }
      }
    }
  }

  if (L::TMA_EPI && warp_idx >= 4 && lane == 0) tma_store_wait<0>();  // output stores have left shared memory
#ifndef VLA_NO_WATCHDOG
  if (warp_idx != 3 && lane == 0) mbar_arrive(done_bar);  // this warp's role is complete
#endif
  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn fa_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// Per-head 4-D view (d, head, row, sample) of a [samples*rows, ld] bf16 matrix whose head h sits at column h*hd.
bool make_head_map(CUtensorMap* m, const void* base, int hd, int heads, int rows, int samples, int ld, int box_d,
                   int box_rows, CUtensorMapSwizzle swz, int sample_rows = 0) {
  if (!sample_rows) sample_rows = rows;
  EncodeTiledFn fn = fa_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(hd), static_cast<cuuint64_t>(heads), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(samples)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(hd) * 2, static_cast<cuuint64_t>(ld) * 2,
                           static_cast<cuuint64_t>(sample_rows) * ld * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_d), 1, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int fa_num_sms() { return device_num_sms(); }

// Work list of one (sample, kv head): every (query head of the group, query tile) unit, heaviest first, paired
// two by two into the slots of one work item.  Returns the number of items, or -1 if the table is too small.
int build_items(int Sq, int Skv, int group, int causal, int bn, uint32_t* items) {
  const int nq = (Sq + FA_BM - 1) / FA_BM, n_all = (Skv + bn - 1) / bn;
  struct Unit { int h, q, cost; };
  std::vector<Unit> units;
  for (int q = 0; q < nq; ++q) {
    const int rows = std::min(FA_BM, Sq - q * FA_BM);
    const int n_c = ((q + 1) * FA_BM + bn - 1) / bn;
    const int nkv = (causal && n_c < n_all) ? n_c : n_all;
    for (int h = 0; h < group; ++h) units.push_back({h, q, nkv * ((rows + 31) / 32)});
  }
  std::stable_sort(units.begin(), units.end(), [](const Unit& a, const Unit& b) { return a.cost > b.cost; });
  const int n_items = (static_cast<int>(units.size()) + 1) / 2;
  if (n_items > FA_MAX_ITEMS || group > 254 || nq > 255) return -1;
  for (int i = 0; i < n_items; ++i) {
    const Unit& a = units[2 * i];
    uint32_t w = static_cast<uint32_t>(a.h) | (static_cast<uint32_t>(a.q) << 8);
    if (2 * i + 1 < static_cast<int>(units.size())) {
      const Unit& b = units[2 * i + 1];
      w |= (static_cast<uint32_t>(b.h) << 16) | (static_cast<uint32_t>(b.q) << 24);
    } else {
      w |= 0xffu << 16;
    }
    items[i] = w;
  }
  return n_items;
}

template <int HD, int BN>
int launch_fa(const __nv_bfloat16* q, int ld_q, int Sq, const __nv_bfloat16* k, const __nv_bfloat16* v, int ld_kv,
              int Skv, int B, int n_heads, int group, int causal, __nv_bfloat16* out, int ld_out, int q_rows,
              cudaStream_t s, const char** err) {
  using L = FaSmem<HD, BN>;
  static PerDeviceFlag attr_flag;  // the shared-memory opt-in is per device
  bool& attr_set = attr_flag.here();
  if (!attr_set) {
    if (cudaFuncSetAttribute(fa_tcgen05_kernel<HD, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) !=
        cudaSuccess) {
      if (err) *err = "attention: cudaFuncSetAttribute failed";
      return -4;
    }
    attr_set = true;
  }
  FaDev p;
  const int kv_heads = n_heads / group;
  const int per_bg = build_items(Sq, Skv, group, causal, BN, p.items);
  if (per_bg < 0) return 1;
  CUtensorMap mQ, mK, mV, mQt, mKt, mVt;
  bool ok = make_head_map(&mQ, q, HD, n_heads, Sq, B, ld_q, 64, FA_BM, CU_TENSOR_MAP_SWIZZLE_128B, q_rows) &&
            make_head_map(&mK, k, HD, kv_heads, Skv, B, ld_kv, 64, BN, CU_TENSOR_MAP_SWIZZLE_128B) &&
            make_head_map(&mV, v, HD, kv_heads, Skv, B, ld_kv, 64, BN, CU_TENSOR_MAP_SWIZZLE_128B);
  if (ok && L::TAIL) {
    ok = make_head_map(&mQt, q, HD, n_heads, Sq, B, ld_q, 16, FA_BM, CU_TENSOR_MAP_SWIZZLE_32B, q_rows) &&
         make_head_map(&mKt, k, HD, kv_heads, Skv, B, ld_kv, 16, BN, CU_TENSOR_MAP_SWIZZLE_32B) &&
         make_head_map(&mVt, v, HD, kv_heads, Skv, B, ld_kv, 16, BN, CU_TENSOR_MAP_SWIZZLE_32B);
  } else if (ok) {
    mQt = mQ;
    mKt = mK;
    mVt = mV;
  }
  CUtensorMap mO = mQ;
  if (ok && L::TMA_EPI) {  // (column, row, sample) view of the output, 32 x 32 boxes in the staging layout
    EncodeTiledFn fn = fa_encode_fn();
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(n_heads) * HD, static_cast<cuuint64_t>(Sq), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld_out) * 2, static_cast<cuuint64_t>(q_rows) * ld_out * 2};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    ok = fn(&mO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  }
  if (!ok) {
    if (err) *err = "attention: cuTensorMapEncodeTiled failed";
    return -4;
  }
  p.bn = BN;
  p.Sq = Sq;
  p.q_rows = q_rows;
  p.Skv = Skv;
  p.group = group;
  p.kv_heads = kv_heads;
  p.causal = causal;
  p.n_bg = B * kv_heads;
  p.n_items = p.n_bg * per_bg;
  p.scale_log2 = (1.0f / sqrtf(static_cast<float>(HD))) * 1.4426950408889634f;
  p.out = out;
  p.ld_out = ld_out;
  {
    const char* dbg = getenv("VLA_FA_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;  // read by the kernel in the trace build only (FA_DBG)
  }
  const char* trace_path = getenv("VLA_FA_TRACE");
  p.trace = nullptr;
  if (trace_path) {
    cudaMalloc(&p.trace, 32004 * sizeof(unsigned int));
    cudaMemsetAsync(p.trace, 0, 32004 * sizeof(unsigned int), s);
  }
  const int grid = std::min(p.n_items, L::CTAS_PER_SM * fa_num_sms());
  launch_kernel(fa_tcgen05_kernel<HD, BN>, dim3(grid), dim3(L::THREADS), L::TOTAL, s, mQ, mK, mV, mQt, mKt, mVt, mO, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = cudaGetErrorString(e);
    return -4;
  }
  if (trace_path) {  // debugging aid: dump CTA 0's event log of this launch
    std::vector<unsigned int> h(32004);
    cudaStreamSynchronize(s);
    cudaMemcpy(h.data(), p.trace, h.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int role = 0; role < 4; ++role) {
        const unsigned int n = h[role] < 4000 ? h[role] : 4000;
        for (unsigned int i = 0; i < n; ++i)
          fprintf(f, "%d %u %u\n", role, h[4 + role * 8000 + 2 * i], h[5 + role * 8000 + 2 * i]);
      }
      fclose(f);
    }
  }
  ops_count_launch();
  return 0;
}

}  // namespace

cudaError_t fa_set_watchdog(WdBuf* dev_ptr, unsigned long long timeout_ms) {
  cudaError_t e = dev_ptr ? wd_set_buffer_this_tu(dev_ptr) : cudaSuccess;
  if (e == cudaSuccess && timeout_ms) e = wd_set_limit_this_tu(timeout_ms);
  return e;
}

// Returns 1 when the shape is not served by this kernel (caller falls back to the mma.sync kernel), 0 on launch.
int attention_tc_launch(const __nv_bfloat16* q, int ld_q, int Sq, const __nv_bfloat16* k, const __nv_bfloat16* v,
                        int ld_kv, int Skv, int B, int n_heads, int group, int hd, int causal, __nv_bfloat16* out,
                        int ld_out, int q_rows, cudaStream_t s, const char** err) {
  if (q_rows < Sq) q_rows = Sq;
  if ((ld_out & 7) || (reinterpret_cast<uintptr_t>(out) & 15)) return 1;
  static int bn64 = -1;  // VLA_FA_BN64=1: 64-key tiles, two CTAs per SM (experiment switch)
  if (bn64 < 0) {
    const char* e = getenv("VLA_FA_BN64");
    bn64 = e ? atoi(e) : 0;
  }
  if (hd == 64 && bn64) return launch_fa<64, 64>(q, ld_q, Sq, k, v, ld_kv, Skv, B, n_heads, group, causal, out, ld_out, q_rows, s, err);
  if (hd == 64) return launch_fa<64, 128>(q, ld_q, Sq, k, v, ld_kv, Skv, B, n_heads, group, causal, out, ld_out, q_rows, s, err);
  if (hd == 72) return launch_fa<72, 128>(q, ld_q, Sq, k, v, ld_kv, Skv, B, n_heads, group, causal, out, ld_out, q_rows, s, err);
  return 1;
}

}  // namespace vla
