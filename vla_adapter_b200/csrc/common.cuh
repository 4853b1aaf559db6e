// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers,
// bf16 packing, warp reductions.  Everything here is inline PTX for sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>

#define VLA_DEVINL __device__ __forceinline__

namespace vla {

// ---------------------------------------------------------------- misc
VLA_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
VLA_DEVINL uint32_t lane_id() { return threadIdx.x & 31; }

VLA_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

VLA_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
VLA_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

VLA_DEVINL uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
VLA_DEVINL float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
VLA_DEVINL float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

VLA_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// x * sigmoid(x) with two MUFU ops (ex2, rcp) and three FMA-pipe ops; an IEEE division here made the SwiGLU epilogue
// the pacing stage of the gate/up GEMM once the MMAs ran on CTA pairs.  |relative error| < 4e-7, far below bf16.
VLA_DEVINL float silu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return x * r;
}

// Exact-erf GELU on two lanes at once with Blackwell's packed fp32 pipe (FFMA2 / FMUL2):
//   gelu(x) = relu(x) - 0.5 |x| erfc(|x| / sqrt 2),   erfc(z) = t (a1 + t (a2 + ...)) exp(-z^2), t = 1/(1 + p z)
// (Abramowitz-Stegun 7.1.26, |erfc error| <= 1.5e-7; measured max |gelu error| 3.4e-7 over [-12, 12], i.e. far
// below one bf16 ulp, and no cancellation in the negative tail).  11 packed FMA-pipe ops + 4 MUFU per pair.
VLA_DEVINL float2 gelu_erf2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 d = __ffma2_rn(ax, make_float2(0.23164189f, 0.23164189f), make_float2(1.f, 1.f));
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
  float2 p = __ffma2_rn(t, make_float2(1.061405429f, 1.061405429f), make_float2(-1.453152027f, -1.453152027f));
  p = __ffma2_rn(p, t, make_float2(1.421413741f, 1.421413741f));
  p = __ffma2_rn(p, t, make_float2(-0.284496736f, -0.284496736f));
  p = __ffma2_rn(p, t, make_float2(0.254829592f, 0.254829592f));
  p = __fmul2_rn(p, t);
  const float2 zs = __fmul2_rn(ax, make_float2(0.849321800f, 0.849321800f));  // |x|/sqrt(2) * sqrt(log2 e)
  const float2 u = __fmul2_rn(zs, zs);
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(-u.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(-u.y));
  const float2 pe = __fmul2_rn(p, e);
  const float2 h = __fmul2_rn(ax, make_float2(-0.5f, -0.5f));
  return __ffma2_rn(h, pe, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}

// ---------------------------------------------------------------- programmatic dependent launch
// Blocks until the grids this one depends on have completed and their writes are visible (no-op without PDL).
VLA_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Lets the next kernel in the stream start being scheduled (it still waits in its own pdl_wait()).
VLA_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier
VLA_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
VLA_DEVINL void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
VLA_DEVINL void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
VLA_DEVINL void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
VLA_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
VLA_DEVINL bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
VLA_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// Latency-tolerant wait (e.g. epilogue warps waiting a whole mainloop for their accumulator): back off between polls
// so the idle warps do not burn issue slots - and power, which is what caps the clocks in this workload.
VLA_DEVINL void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(100);
}
VLA_DEVINL void mbar_wait_short(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(20);
}

// ---------------------------------------------------------------- watchdog
// A protocol deadlock inside a warp-specialised kernel must not hang the process: every tcgen05 kernel carries one
// MONITOR warp that does nothing but wait (suspended in mbarrier.try_wait, no issue slots) for the CTA's `done`
// barrier, on which every working warp arrives when its role loop ends.  If the CTA is still not done after the time
// limit (10 s; VLA_WATCHDOG_MS / vla_watchdog_set_timeout_ms) the monitor writes, for every warp of the CTA, the note
// that warp left before its last wait (which barrier, which parity, which step) plus the raw state of that mbarrier to
// a host-mapped buffer and traps: the hang becomes a CUDA launch failure whose message names the barrier
// (vla_watchdog_report).  The working warps pay one shared-memory store per wait (wd_note) and no registers - bounding
// every wait in place was tried first and spilled in the attention kernel's 32-register warps.
struct WdRecord {
  uint32_t magic, kernel, block, smid, warp, note, bar_lo, bar_hi;
};
constexpr unsigned int WD_MAX_RECORDS = 127;
struct WdBuf {
  unsigned int pad[8];
  WdRecord rec[WD_MAX_RECORDS];
};
constexpr uint32_t WD_MAGIC = 0x57444f47u;  // "WDOG"
enum : uint32_t { WD_K_GEMM1 = 1, WD_K_GEMM2 = 2, WD_K_FA64 = 3, WD_K_FA72 = 4, WD_K_POLICY = 5, WD_K_SELFTEST = 0x7f };
// note = barrier kind [31:24] | barrier index [23:16] | parity [15] | step [14:0]
#define VLA_WD_NOTE(kind, idx, parity, step)                                                \
  ((static_cast<uint32_t>(kind) << 24) | ((static_cast<uint32_t>(idx) & 0xffu) << 16) |   \
   ((static_cast<uint32_t>(parity) & 1u) << 15) | (static_cast<uint32_t>(step) & 0x7fffu))

// One copy per translation unit (the library is built without -rdc); every TU with a monitored kernel exports a
// setter built on these and the engine calls them all for its device (watchdog.cu).
static __device__ WdBuf* g_wd_buf = nullptr;
static __device__ unsigned int g_wd_slots = 0;  // slot allocation stays in device memory (no PCIe atomics needed)
static __constant__ uint32_t g_wd_limit_ticks = 9537;  // time limit in units of 2^20 ns (10 s)
static inline cudaError_t wd_set_buffer_this_tu(WdBuf* host_mapped) {
  return cudaMemcpyToSymbol(g_wd_buf, &host_mapped, sizeof(host_mapped));
}
static inline cudaError_t wd_set_limit_this_tu(unsigned long long timeout_ms) {
  unsigned long long t = (timeout_ms * 1000000ull) >> 20;
  const uint32_t ticks = t < 2 ? 2u : (t > 0x7ffffull ? 0x7ffffu : static_cast<uint32_t>(t));
  return cudaMemcpyToSymbol(g_wd_limit_ticks, &ticks, sizeof(ticks));
}

VLA_DEVINL unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Working warps: "I am about to wait on this barrier" (all lanes store the same word: one instruction, no branch).
// -DVLA_NO_WATCHDOG compiles the whole mechanism out (A/B measurements of its cost).
VLA_DEVINL void wd_note(uint32_t slot_addr, uint32_t note) {
#ifndef VLA_NO_WATCHDOG
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(slot_addr), "r"(note) : "memory");
#endif
}
// Cold end of wd_monitor_min: ONE 16-byte store {magic, kernel, CTA, 0xffffffff} at a slot derived from the CTA index,
// a system fence, the trap.  It has to stay this small: struct stores through a generic pointer, an atomic slot counter
// or a grace loop here made ptxas spill throughout the attention kernel's 32-register warpgroup (12 -> 380 bytes).
VLA_DEVINL void wd_record_min(uint32_t kernel) {
  WdBuf* w = g_wd_buf;
  if (w) {
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};\n\tfence.acq_rel.sys;"
                 ::"l"(&w->rec[blockIdx.x % WD_MAX_RECORDS]), "r"(WD_MAGIC), "r"(kernel), "r"(blockIdx.x), "r"(0xffffffffu)
                 : "memory");
  }
  __trap();
}

// Monitor warp (all 32 lanes enter; lane 0 does the waiting): returns when `done_bar` completes.  On timeout lane 0
// dumps `n_warps` notes (slots_addr[w]) with the barrier each note names (bar_of(kind, idx) -> shared address or 0),
// then the raw words of the `n_bars` mbarriers starting at `bars_addr` (records with warp = 0x100 + index), and traps.
// The wait is a bare try_wait loop: the hardware suspends the lane until the barrier completes or its own time limit
// expires, so the warp costs no issue slots and wakes at once - with a 200 ns back-off per poll the monitor's wake-up
// latency at the end of every kernel added 0.5 ms to the 800-kernel bs=1 forward.  The exit is warp-uniform
// (__syncwarp), as the .aligned barriers behind it require.
template <class BarOf>
VLA_DEVINL void wd_monitor(uint32_t done_bar, uint32_t slots_addr, int n_warps, uint32_t kernel, BarOf bar_of,
                           uint32_t bars_addr = 0, int n_bars = 0) {
  if ((threadIdx.x & 31u) == 0) {
    uint32_t polls = 0, t0 = 0;
    while (!mbar_try_wait(done_bar, 0)) {
      if ((++polls & 0xffu) != 0) continue;
      const uint32_t now = static_cast<uint32_t>(global_timer_ns() >> 20) | 1u;
      if (!t0) {
        t0 = now;
        continue;
      }
      if (now - t0 <= g_wd_limit_ticks) continue;
      WdBuf* w = g_wd_buf;
      if (w) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        for (int i = 0; i < n_warps + n_bars; ++i) {
          uint32_t note = 0, lo = 0, hi = 0, who;
          if (i < n_warps) {
            who = static_cast<uint32_t>(i);
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(note) : "r"(slots_addr + 4u * i));
            const uint32_t bar = bar_of(note >> 24, (note >> 16) & 0xffu);
            if (bar) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(bar));
          } else {
            who = 0x100u + static_cast<uint32_t>(i - n_warps);
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(bars_addr + 8u * (i - n_warps)));
          }
          const unsigned int slot = atomicAdd(&g_wd_slots, 1u);
          if (slot < WD_MAX_RECORDS) {
            volatile WdRecord* r = &w->rec[slot];
            r->kernel = kernel; r->block = blockIdx.x; r->smid = smid; r->warp = who;
            r->note = note; r->bar_lo = lo; r->bar_hi = hi;
            __threadfence_system();
            r->magic = WD_MAGIC;
          }
        }
        __threadfence_system();
      }
      // grace period: monitors of other stuck CTAs get to write their records before the trap takes the context down
      const unsigned long long t1 = global_timer_ns();
      while (global_timer_ns() - t1 < 100000000ull) __nanosleep(1000);
      __trap();
    }
  }
  __syncwarp();
}

// The same monitor without the barrier dump, for warps that have no registers to spare (the attention kernel's
// monitor lives in the 32-register warpgroup: with the full dump inlined there ptxas spilled inside the MMA-issuing
// warps' loops and the kernel ran 230 us instead of 134 us).  The record carries the kernel, CTA and SM only.
VLA_DEVINL void wd_monitor_min(uint32_t done_bar, uint32_t kernel) {
  if ((threadIdx.x & 31u) == 0) {
    uint32_t polls = 0, t0 = 0;
    while (!mbar_try_wait(done_bar, 0)) {
      if ((++polls & 0xffu) != 0) continue;
      const uint32_t now = static_cast<uint32_t>(global_timer_ns() >> 20) | 1u;
      if (!t0) {
        t0 = now;
        continue;
      }
      if (now - t0 <= g_wd_limit_ticks) continue;
      wd_record_min(kernel);
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------- TMA
VLA_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 3-D tiled load, completes on an mbarrier (bytes).  Coordinates are (inner, row, batch).
VLA_DEVINL void tma_load_3d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
VLA_DEVINL void tma_load_3d_hint(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0,
                                 int c1, int c2, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
      "l"(hint)
      : "memory");
}
// 3-D tiled store smem -> global (bulk group completion)
VLA_DEVINL void tma_store_3d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
VLA_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
VLA_DEVINL void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
VLA_DEVINL void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
VLA_DEVINL void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
VLA_DEVINL void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
VLA_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
VLA_DEVINL void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
VLA_DEVINL void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; bf16 inputs, fp32 accumulate.  One thread issues.
VLA_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
VLA_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
VLA_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of fp32 -> 32 registers per thread (thread i = lane i of the warp's quarter)
VLA_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------- CTA pairs (cta_group::2) and clusters
VLA_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
VLA_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
VLA_DEVINL uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
VLA_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-pair bit of a shared address: the even (leader) CTA
// 3-D tiled load issued by either CTA of a pair; the transaction bytes are credited to the LEADER's mbarrier.
VLA_DEVINL void tma_load_3d_cg2(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
VLA_DEVINL void tmem_alloc_cg2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
VLA_DEVINL void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
VLA_DEVINL void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 per CTA] * B[N rows: N/2 per CTA]^T; issued by the leader CTA only.
VLA_DEVINL void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the mbarrier at this offset in every CTA of `cta_mask` once the pair's MMAs issued so far retire.
VLA_DEVINL void umma_commit_cg2(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask)
               : "memory");
}

// Shared-memory matrix descriptor for a K-major bf16 tile written by TMA with 128B swizzle:
// rows are 128 bytes (64 bf16), 8-row groups are 1024 bytes apart (SBO), version=1 (sm_100),
// layout_type=2 (SWIZZLE_128B).  Bit layout per cute::UMMA::SmemDescriptor.
VLA_DEVINL uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);  // start address  [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO = 1024 B   [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // version = 1    [46,48)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B   [61,64)
  return d;
}

// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32, A and B both K-major.
// Bit layout per cute::UMMA::InstrDescriptor.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4)           // c_format = F32
         | (1u << 7)         // a_format = BF16
         | (1u << 10)        // b_format = BF16
         | ((N >> 3) << 17)  // n_dim
         | ((M >> 4) << 24); // m_dim
}

}  // namespace vla
