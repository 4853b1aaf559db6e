// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[b, r, n] = epilogue( sum_k A[b, r, k] * W[n, k] )     bf16 x bf16 -> fp32 (TMEM) -> bf16
//
// Every dense contraction of the predict_action path runs through this kernel: the ViT blocks'
// qkv/proj/fc1/fc2 (timm Block, restated in film_vit_wrapper.py:69-75), the fused projector
// (modeling_prismatic.py:261-273), Qwen2's q/k/v/o/gate/up/down, and the Bridge-Attention policy's
// q/k/v/o/ffn projections (action_heads.py:247-254, 355-367).
//
// Structure (one CTA per SM, 352 threads):
//   warp 0   : TMA producer  - cp.async.bulk.tensor (3-D maps, 128B swizzle) into a STAGES-deep ring
//   warp 1   : MMA issuer    - one thread issues tcgen05.mma (M=128, N=BN, K=16) from smem descriptors
//   warps 2-9: epilogue      - tcgen05.ld the fp32 accumulator out of TMEM, bias/act/scale/residual,
//                              cast to bf16, swizzled smem staging, TMA store / reduce-add
//   warp 10  : watchdog monitor (common.cuh): parked on the CTA's `done` barrier; dumps every warp's last wait and
//              traps if the CTA does not finish within the time limit
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1.
#include "common.cuh"
#include "launch.cuh"
#include "gemm.cuh"
#include "watchdog.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace vla {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
#ifdef VLA_NO_WATCHDOG  // bisecting aid: the kernel without its monitor warp
constexpr int GEMM_THREADS = 320;
#else
constexpr int GEMM_THREADS = 352;          // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue, warp 10 watchdog monitor
#endif
constexpr int GEMM_WORK_WARPS = 10;
constexpr int EPI_WARPS = 8;
constexpr uint32_t A_BYTES = BM * BK * 2;  // 16 KB per stage
// Shared memory: [operand ring | epilogue staging | barriers].  Ring + staging are 224 KB in both layouts:
//   plain epilogue          : 192 KB ring + 8 warps x 2 boxes (32 rows x 64 B) - a box is written, then TMA-stored
//   staged-residual epilogue: 160 KB ring + 8 warps x 4 boxes - the residual box is TMA-LOADED into the box two chunks
//                             ahead, the epilogue adds in fp32 in place and TMA-stores the same box
constexpr uint32_t RING_BYTES = 192 * 1024;
constexpr uint32_t RING_BYTES_STAGED = 160 * 1024;
constexpr uint32_t STAGING_BYTES = EPI_WARPS * 2 * 2048;
constexpr uint32_t BAR_OFFSET = RING_BYTES + STAGING_BYTES;  // = RING_BYTES_STAGED + EPI_WARPS * 4 * 2048
constexpr uint32_t SMEM_BYTES = BAR_OFFSET + 1024 /*align slack*/ + 512 /*barriers*/;
constexpr int RES_BUFS = 4;      // staging boxes per epilogue warp in the staged-residual layout
constexpr int RES_AHEAD = 2;     // residual boxes in flight ahead of the chunk being finished
constexpr uint32_t TMEM_COLS = 512;        // two accumulators of up to 256 columns
constexpr int MAX_STAGES = 8;

struct GemmDev {
  int rows, batches, N, K;
  int mt_per_batch, tiles_m, tiles_n;
  int pack;        // > 0: short row views (rows < 128, 128 % rows == 0): one M tile = `pack` consecutive batches
  int bn;          // tile width: 64 / 128 / 192 / 224 / 256
  int stages;      // smem ring depth = min(8, 192 KB / stage bytes)
  const float* bias;
  const float* colscale;
  int act;
  int accumulate;  // 1: C += epi(...) through TMA reduce-add (C holds the residual)
  // residual that lives in a DIFFERENT buffer than C: read by the epilogue and added in fp32 (no pre-copy kernel)
  const __nv_bfloat16* resid;
  long long r_bs;
  int ldr;
  const uint32_t* rope_cs;  // RoPE fused into the epilogue (see GemmArgs)
  int rope_cols, rope_S, rope_ld;
  const float2* row_stats;  // normalisation of A folded into the epilogue (see GemmArgs)
  const float* colsum;
  int resid_tma;            // 1: the residual arrives through mapR into the staging boxes (staged-residual epilogue)
  uint32_t ring_bytes;      // operand ring size of the layout in use
  // row statistics without a statistics kernel (see GemmArgs::stat_out / stat_in)
  float2* stat_out;
  const float2* stat_in;
  float stat_inv_dim, stat_eps;
  int stat_rms;
  // profiling only (vla_profile_gemm): CTA 0's monitor warp writes {globaltimer, clock64} at its start and end, which
  // gives the kernel's duration AND the SM clock it ran at inside a real step (ncu serialises and cannot show that)
  unsigned long long* prof;
};

VLA_DEVINL void tma_reduce_add_3d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile(
      "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
      ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// One 32-row x 32-column output box of a warp: 16 packed bf16x2 words per thread (thread = row) go to the
// warp's staging buffer in the 64-byte-swizzled layout the C tensor map expects, then one lane issues the
// TMA store (or reduce-add).  OOB rows / columns are clipped by the TMA unit.
VLA_DEVINL void store_box(const CUtensorMap* mapC, uint32_t buf_addr, int lane, const uint32_t (&o)[16], int c0,
                          int r0, int b, int accumulate) {
  if (lane == 0) tma_store_wait_read<1>();  // the store that used this buffer two boxes ago has read it
  __syncwarp();
  const uint32_t row_addr = buf_addr + lane * 64;
  const int sw = (lane >> 1) & 3;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint32_t addr = row_addr + ((u ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[4 * u]), "r"(o[4 * u + 1]),
                 "r"(o[4 * u + 2]), "r"(o[4 * u + 3])
                 : "memory");
  }
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    if (accumulate) tma_reduce_add_3d(mapC, buf_addr, c0, r0, b);
    else tma_store_3d(mapC, buf_addr, c0, r0, b);
    tma_store_commit();
  }
}

// The watchdog monitor of a GEMM CTA (common.cuh), deliberately NOT inlined and fed plain values only: with the record
// dump inlined into the kernel body (a lambda capturing the barrier-address lambdas by reference) the kernel hung on
// its first launch on B200 - even in builds whose monitor branch never executed - while this form and a dump-less
// inline loop both run (bisected on the GPU, round 2).
__device__ __noinline__ void gemm_monitor(uint32_t done_bar, uint32_t note_base, uint32_t bar_base, uint32_t kernel) {
  wd_monitor(done_bar, note_base, GEMM_WORK_WARPS, kernel, [bar_base](uint32_t kind, uint32_t idx) {
    return kind == 1   ? bar_base + 8u * (MAX_STAGES + idx)          // empty(stage)
           : kind == 2 ? bar_base + 8u * idx                         // full(stage)
           : kind == 3 ? bar_base + 8u * (2 * MAX_STAGES + 2 + idx)  // tmem_empty(acc)
           : kind == 4 ? bar_base + 8u * (2 * MAX_STAGES + idx)      // tmem_full(acc)
                       : 0u;
  });
}

// CG = 1: one CTA per 128 x bn output tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x bn tile -
// each CTA stages its own 128 rows of A and HALF of the B tile, the leader issues 256 x bn x 16 MMAs that read both
// CTAs' shared memory and write each CTA's 128 rows into its own TMEM; per FLOP that halves the B traffic through L2,
// TMA and shared memory (the GEMM is power-capped on B200: less data movement = higher clocks).
template <int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                         const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapR,
                         const GemmDev p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_addr + pad;

  const int STAGES = p.stages;
  const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;  // 0 = leader of the pair
  const uint32_t b_bytes = static_cast<uint32_t>(p.bn / CG) * BK * 2;  // this CTA's share of the B tile
  const uint32_t stage_bytes = A_BYTES + b_bytes;
  const uint32_t staging_base = smem_base + p.ring_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFFSET);
  const uint32_t bar_base = smem_base + BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  const uint32_t done_bar = bar_base + 8u * (2 * MAX_STAGES + 5);        // watchdog: every working warp arrives at its end
  const uint32_t note_base = bar_base + 8u * (2 * MAX_STAGES + 6);       // watchdog: one "last wait" word per warp
  auto res_bar = [&](int e, int i) { return bar_base + 256u + 8u * (e * RES_BUFS + i); };  // residual box landed

  // Warp index broadcast with shfl so the compiler knows the role branches are warp-uniform: the MMA warp's
  // descriptor arithmetic then stays on the uniform datapath and tcgen05.mma issues at the hardware rate
  // (measured: 128 cycles per 128x256x16, against ~160 when one diverged lane computes descriptors).
  const int warp_idx = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    tma_prefetch_desc(&mapC);
    if (p.resid_tma) tma_prefetch_desc(&mapR);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), EPI_WARPS * CG);  // the leader's barrier collects the epilogue warps of both CTAs
    }
    mbar_init(done_bar, GEMM_WORK_WARPS);
    if (p.resid_tma)
      for (int i = 0; i < EPI_WARPS * RES_BUFS; ++i) mbar_init(res_bar(0, i), 1);
    mbar_fence_init();
    fence_proxy_async();
  }
  if (warp_idx == 1) {
    if (CG == 2) {
      tmem_alloc_cg2(smem_u32(tmem_slot), TMEM_COLS);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
      tmem_relinquish();
    }
    tc_fence_before();
  }
  if (CG == 2) cluster_sync_all();  // the peer's barriers are initialised before anything is signalled across
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  // Everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel's tail.
  pdl_wait();
  pdl_launch_dependents();

  const int num_kb = (p.K + BK - 1) / BK;
  const int total_tiles = p.tiles_m * p.tiles_n;  // tiles of (BM * CG) rows x bn columns
  const int tile0 = static_cast<int>(blockIdx.x) / CG, tile_step = static_cast<int>(gridDim.x) / CG;
  const uint32_t my_note = note_base + 4u * static_cast<uint32_t>(warp_idx);

  if (warp_idx == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp waits, one lane issues)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const int n_idx = tile % p.tiles_n;
      const int m_idx = tile / p.tiles_n;
      // packed short views: the A box is (64, rows, pack) - `pack` whole batches land as 128 dense smem rows
      const int b = p.pack ? m_idx * p.pack : m_idx / p.mt_per_batch;
      const int r0 = p.pack ? 0 : (m_idx - b * p.mt_per_batch) * (BM * CG) + static_cast<int>(crank) * BM;
      const int n0 = n_idx * p.bn + static_cast<int>(crank) * (p.bn / CG);
      for (int kb = 0; kb < num_kb; ++kb) {
        wd_note(my_note, VLA_WD_NOTE(1, stage, phase ^ 1u, kb));
        mbar_wait_relaxed(empty_bar(stage), phase ^ 1u);  // the ring is deep: the producer mostly waits, politely
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * stage_bytes;
          if (CG == 2) {
            // both CTAs' loads are credited to the leader's barrier, which expects the bytes of the whole pair
            if (crank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * stage_bytes);
            tma_load_3d_cg2(sa, &mapA, full_bar(stage), kb * BK, r0, b);
            tma_load_3d_cg2(sa + A_BYTES, &mapB, full_bar(stage), kb * BK, n0, 0);
          } else {
            mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
            tma_load_3d(sa, &mapA, full_bar(stage), kb * BK, r0, b);
            tma_load_3d(sa + A_BYTES, &mapB, full_bar(stage), kb * BK, n0, 0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp waits, one lane issues)
    const uint32_t idesc = make_idesc_bf16(BM * CG, static_cast<uint32_t>(p.bn));
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (crank == 0)  // only the leader of a pair issues MMAs
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      wd_note(my_note, VLA_WD_NOTE(3, acc, acc_phase ^ 1u, tile));
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
      for (int kb = 0; kb < num_kb; ++kb) {
        wd_note(my_note, VLA_WD_NOTE(2, stage, phase, kb));
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_sw128_kmajor_desc(sa + k * 32);
            const uint64_t bdesc = make_sw128_kmajor_desc(sb + k * 32);
            if (CG == 2) umma_bf16_cg2(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs retire
          if (CG == 2) umma_commit_cg2(empty_bar(stage), 3);
          else umma_commit(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (elect_one()) {  // accumulator complete -> epilogue (of both CTAs)
        if (CG == 2) umma_commit_cg2(tfull_bar(acc), 3);
        else umma_commit(tfull_bar(acc));
      }
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else if (warp_idx == GEMM_WORK_WARPS) {
    // ------------------------------------------------------------ watchdog monitor (see common.cuh)
    const bool prof = p.prof != nullptr && blockIdx.x == 0 && lane == 0;
    if (prof) {
      p.prof[0] = global_timer_ns();
      p.prof[1] = static_cast<unsigned long long>(clock64());
    }
    gemm_monitor(done_bar, note_base, bar_base, CG == 2 ? WD_K_GEMM2 : WD_K_GEMM1);
    if (prof) {
      p.prof[2] = global_timer_ns();
      p.prof[3] = static_cast<unsigned long long>(clock64());
    }
  } else {
    // ------------------------------------------------------------ epilogue: 8 warps, each 32 rows x bn/2 columns
    const int e = warp_idx - 2;
    const int quarter = warp_idx & 3;  // TMEM lane quarter this warp may access
    const int half = e >> 2;           // which half of the tile's columns
    // accumulator columns per warp, in whole 32-column chunks: bn / 2 each, except tile width 224 = 128 + 96
    const int w0 = ((p.bn >> 1) + 31) & ~31;
    const int wcols = half ? p.bn - w0 : w0;
    const int wcol0 = half ? w0 : 0;
    const uint32_t buf0 = staging_base + static_cast<uint32_t>(e) * (p.resid_tma ? RES_BUFS * 2048u : 4096u);
    int buf = 0;
    uint32_t res_g = 0;  // staged residual: running chunk counter of this warp (box = res_g % RES_BUFS, parity from res_g)
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool swiglu = p.act == ACT_SWIGLU;
    const uint32_t tempty_leader0 = CG == 2 ? mapa_shared(tempty_bar(0), 0) : 0u;
    for (int tile = tile0; tile < total_tiles; tile += tile_step) {
      const int n_idx = tile % p.tiles_n;
      const int m_idx = tile / p.tiles_n;
      int b = m_idx / p.mt_per_batch;
      int r0 = (m_idx - b * p.mt_per_batch) * (BM * CG) + static_cast<int>(crank) * BM + quarter * 32;
      if (p.pack) {  // this warp's 32 tile rows = rows [t0 % rows, ...) of batch m_idx*pack + t0 / rows (and the next ones)
        const int t0 = quarter * 32;
        b = m_idx * p.pack + t0 / p.rows;
        r0 = t0 % p.rows;
      }
      const bool live = p.pack ? b < p.batches : r0 < p.rows;
      const int n0 = n_idx * p.bn + wcol0;
      float2 rst = make_float2(1.f, 0.f);  // this thread's row: (rstd, -mean * rstd)
      if (p.row_stats && r0 + lane < p.rows) rst = __ldg(p.row_stats + r0 + lane);
      if (p.stat_in && r0 + lane < p.rows) {
        // the producing GEMM left STAT_SLOTS partial (sum x, sum x^2) pairs per row: finish the statistics here
        const float4* sp = reinterpret_cast<const float4*>(p.stat_in + static_cast<long long>(r0 + lane) * STAT_SLOTS);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < STAT_SLOTS / 2; ++k) {
          const float4 t = __ldg(sp + k);
          s1 += t.x + t.z;
          s2 += t.y + t.w;
        }
        if (p.stat_rms) {
          rst = make_float2(rsqrtf(s2 * p.stat_inv_dim + p.stat_eps), 0.f);
        } else {
          const float mean = s1 * p.stat_inv_dim;
          const float rstd = rsqrtf(fmaxf(s2 * p.stat_inv_dim - mean * mean, 0.f) + p.stat_eps);
          rst = make_float2(rstd, -mean * rstd);
        }
      }
      // staged residual: chunk i of this tile uses box (res_g + i) % RES_BUFS; the first RES_AHEAD boxes are requested
      // before the wait for the accumulator, so they land under the tile's mainloop
      const bool res_live = p.resid_tma && live && n0 < p.N;
      auto res_fetch = [&](int i) {  // lane 0: residual box of chunk i -> its staging box
        const int c0 = n0 + i * 32;
        if (i * 32 < wcols && c0 < p.N) {
          const uint32_t g = res_g + static_cast<uint32_t>(i);
          tma_store_wait_read<1>();  // the store that last used this box (RES_BUFS chunks ago) has read it
          mbar_arrive_expect_tx(res_bar(e, g % RES_BUFS), 2048u);
          tma_load_3d(buf0 + (g % RES_BUFS) * 2048u, &mapR, res_bar(e, g % RES_BUFS), c0, r0, b);
        }
      };
      if (res_live && lane == 0) {
#pragma unroll
        for (int i = 0; i < RES_AHEAD; ++i) res_fetch(i);
      }
      float2 st_sum = make_float2(0.f, 0.f), st_sq = make_float2(0.f, 0.f);  // this row's partial sums (stat_out)

      wd_note(my_note, VLA_WD_NOTE(4, acc, acc_phase, tile));
      mbar_wait_relaxed(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * 256 + wcol0);
      if (live && n0 < p.N) {  // warp-uniform: sub-tiles entirely out of bounds are skipped
        if (swiglu) {
#pragma unroll 1
          for (int ch = 0; ch < wcols; ch += 64) {
            const int c0 = n0 + ch;
            if (c0 >= p.N) break;
            uint32_t o[16];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              uint32_t v[32];
              tmem_ld_32x32b_x32(t_row + ch + hh * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float g0 = rst.x * __uint_as_float(v[2 * j]), g1 = rst.x * __uint_as_float(v[2 * j + 1]);
                const float u0 = rst.x * __uint_as_float(v[16 + 2 * j]), u1 = rst.x * __uint_as_float(v[16 + 2 * j + 1]);
                o[hh * 8 + j] = pack_bf16(silu(g0) * u0, silu(g1) * u1);
              }
            }
            store_box(&mapC, buf0 + buf * 2048u, lane, o, c0 >> 1, r0, b, 0);
            buf ^= 1;
          }
        } else if (p.rope_cols > 0) {
          // q/k projection with RoPE: a 64-wide head = two 32-column chunks (lo, hi) of this warp; out_lo = lo*cos -
          // hi*sin, out_hi = hi*cos + lo*sin, every product and sum rounded to bf16 like the eager reference.
          // This row's 32 (cos, sin) pairs are loaded once per tile (the same for every head).
          uint32_t cs[32];
          if (n0 < p.rope_cols) {
            const uint32_t* tab = p.rope_cs + (r0 + lane) % p.rope_S;
#pragma unroll
            for (int j = 0; j < 32; ++j) cs[j] = __ldg(tab + j * p.rope_ld);
          }
#pragma unroll 1
          for (int ch = 0; ch < wcols; ch += 64) {
            const int c0 = n0 + ch;
            if (c0 >= p.N) break;
            uint32_t vl[32], vh[32];
            tmem_ld_32x32b_x32(t_row + ch, vl);
            tmem_ld_32x32b_x32(t_row + ch + 32, vh);
            tmem_ld_wait();
            const bool rot = c0 < p.rope_cols;
            uint32_t ol[16], oh[16];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 bl = make_float4(0.f, 0.f, 0.f, 0.f), bh = bl;
              if (p.bias) {
                if (c0 + q * 4 < p.N) bl = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + q * 4));
                if (c0 + 32 + q * 4 < p.N) bh = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + 32 + q * 4));
              }
              const float bls[4] = {bl.x, bl.y, bl.z, bl.w}, bhs[4] = {bh.x, bh.y, bh.z, bh.w};
              float rl[4], rh[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float a = bf16_round(fmaf(rst.x, __uint_as_float(vl[4 * q + t]), bls[t]));   // the projection output, bf16
                const float bq = bf16_round(fmaf(rst.x, __uint_as_float(vh[4 * q + t]), bhs[t]));
                if (rot) {
                  const float cc = __uint_as_float(cs[4 * q + t] << 16), sn = __uint_as_float(cs[4 * q + t] & 0xffff0000u);
                  rl[t] = bf16_round(bf16_round(a * cc) + bf16_round(-bq * sn));
                  rh[t] = bf16_round(bf16_round(bq * cc) + bf16_round(a * sn));
                } else {
                  rl[t] = a;
                  rh[t] = bq;
                }
              }
              ol[2 * q] = pack_bf16(rl[0], rl[1]);
              ol[2 * q + 1] = pack_bf16(rl[2], rl[3]);
              oh[2 * q] = pack_bf16(rh[0], rh[1]);
              oh[2 * q + 1] = pack_bf16(rh[2], rh[3]);
            }
            store_box(&mapC, buf0 + buf * 2048u, lane, ol, c0, r0, b, 0);
            buf ^= 1;
            if (c0 + 32 < p.N) {
              store_box(&mapC, buf0 + buf * 2048u, lane, oh, c0 + 32, r0, b, 0);
              buf ^= 1;
            }
          }
        } else {
#pragma unroll 1
          for (int ch = 0; ch < wcols; ch += 32) {
            const int c0 = n0 + ch;
            if (c0 >= p.N) break;
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_row + ch, v);
            float2 f[16];
            if (p.bias) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + q * 4 < p.N) bv = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + q * 4));
                f[2 * q] = make_float2(bv.x, bv.y);
                f[2 * q + 1] = make_float2(bv.z, bv.w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = make_float2(0.f, 0.f);
            }
            if (p.colsum) {  // mean term of a folded LayerNorm
              const float2 nm = make_float2(rst.y, rst.y);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + q * 4 < p.N) cv = __ldg(reinterpret_cast<const float4*>(p.colsum + c0 + q * 4));
                f[2 * q] = __ffma2_rn(nm, make_float2(cv.x, cv.y), f[2 * q]);
                f[2 * q + 1] = __ffma2_rn(nm, make_float2(cv.z, cv.w), f[2 * q + 1]);
              }
            }
            tmem_ld_wait();
            {
              const float2 rs2 = make_float2(rst.x, rst.x);
#pragma unroll
              for (int j = 0; j < 16; ++j)
                f[j] = __ffma2_rn(rs2, make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), f[j]);
            }
            if (p.act == ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = gelu_erf2(f[j]);
            } else if (p.act == ACT_RELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = make_float2(fmaxf(f[j].x, 0.0f), fmaxf(f[j].y, 0.0f));
            }
            if (p.colscale) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c0 + q * 4 < p.N) sv = __ldg(reinterpret_cast<const float4*>(p.colscale + c0 + q * 4));
                f[2 * q] = __fmul2_rn(f[2 * q], make_float2(sv.x, sv.y));
                f[2 * q + 1] = __fmul2_rn(f[2 * q + 1], make_float2(sv.z, sv.w));
              }
            }
            if (p.resid_tma) {
              // staged residual: the box landed in shared memory (64B-swizzled like the output box); add in fp32
              const uint32_t g = res_g + static_cast<uint32_t>(ch >> 5);
              const uint32_t box = buf0 + (g % RES_BUFS) * 2048u;
              if (lane == 0) res_fetch((ch >> 5) + RES_AHEAD);
              __syncwarp();
              mbar_wait(res_bar(e, g % RES_BUFS), (g / RES_BUFS) & 1u);
              const uint32_t row_addr = box + lane * 64;
              const int sw = (lane >> 1) & 3;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint32_t r0w, r1w, r2w, r3w;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(r0w), "=r"(r1w), "=r"(r2w), "=r"(r3w)
                             : "r"(row_addr + ((q ^ sw) << 4)));
                f[4 * q] = __fadd2_rn(f[4 * q], unpack_bf16(r0w));
                f[4 * q + 1] = __fadd2_rn(f[4 * q + 1], unpack_bf16(r1w));
                f[4 * q + 2] = __fadd2_rn(f[4 * q + 2], unpack_bf16(r2w));
                f[4 * q + 3] = __fadd2_rn(f[4 * q + 3], unpack_bf16(r3w));
              }
            } else if (p.resid) {  // out-of-place residual: this thread's row, 32 columns = four 16-byte loads
              const int rr = r0 + lane;
              if (rr < p.rows) {
                const __nv_bfloat16* rp = p.resid + b * p.r_bs + static_cast<long long>(rr) * p.ldr + c0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  if (c0 + q * 8 < p.N) {
                    const uint4 rv = __ldg(reinterpret_cast<const uint4*>(rp + q * 8));
                    const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                    for (int t = 0; t < 4; ++t) f[4 * q + t] = __fadd2_rn(f[4 * q + t], unpack_bf16(rw[t]));
                  }
                }
              }
            }
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = pack_bf16(f[j].x, f[j].y);
            if (p.stat_out) {  // partial row sums of what was just produced (columns past N hold zeros: bias / TMA fill)
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const bool in_n = c0 + 2 * j < p.N;
                const float2 u = in_n ? f[j] : make_float2(0.f, 0.f);
                st_sum = __fadd2_rn(st_sum, u);
                st_sq = __ffma2_rn(u, u, st_sq);
              }
            }
            if (p.resid_tma) {
              // in place: the box the residual arrived in is rewritten and stored (no other store is pending on it)
              const uint32_t g = res_g + static_cast<uint32_t>(ch >> 5);
              const uint32_t box = buf0 + (g % RES_BUFS) * 2048u;
              const uint32_t row_addr = box + lane * 64;
              const int sw = (lane >> 1) & 3;
              __syncwarp();
#pragma unroll
              for (int u = 0; u < 4; ++u)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + ((u ^ sw) << 4)), "r"(o[4 * u]),
                             "r"(o[4 * u + 1]), "r"(o[4 * u + 2]), "r"(o[4 * u + 3])
                             : "memory");
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&mapC, box, c0, r0, b);
                tma_store_commit();
              }
            } else {
              store_box(&mapC, buf0 + buf * 2048u, lane, o, c0, r0, b, p.accumulate);
              buf ^= 1;
            }
          }
          if (p.resid_tma) {  // chunks of this tile that were consumed
            int used = 0;
            for (int ch = 0; ch < wcols && n0 + ch < p.N; ch += 32) ++used;
            res_g += static_cast<uint32_t>(used);
          }
        }
      }
      if (p.stat_out && r0 + lane < p.rows) {
        // slot of this (column tile, half); a sub-tile that lies past N writes zeros; the first warp of a row also
        // clears the slots no column tile owns, so that the consumer can always sum STAT_SLOTS of them
        float2* sp = p.stat_out + static_cast<long long>(r0 + lane) * STAT_SLOTS;
        sp[n_idx * 2 + half] = make_float2(st_sum.x + st_sum.y, st_sq.x + st_sq.y);
        if (n_idx == 0 && half == 0)
          for (int k = p.tiles_n * 2; k < STAT_SLOTS; ++k) sp[k] = make_float2(0.f, 0.f);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty_leader0 + 8u * acc);  // the leader's MMA warp owns the accumulators
        else mbar_arrive(tempty_bar(acc));
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
    if (lane == 0) tma_store_wait_read<0>();  // staging smem must outlive the last bulk reads
  }
#ifndef VLA_NO_WATCHDOG
  if (warp_idx < GEMM_WORK_WARPS) {  // this warp's role is complete
    wd_note(my_note, 0xffffffffu);
    if (lane == 0) mbar_arrive(done_bar);
  }
#endif

  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // neither CTA may leave while the pair's MMAs / remote arrivals can touch it
  else __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// dst view = src view (bf16 rows of `cols` elements, 16-byte vectors); src batch stride may be 0 (broadcast)
__global__ void __launch_bounds__(256)
copy_view_kernel(const __nv_bfloat16* __restrict__ src, long long s_bs, int lds, __nv_bfloat16* __restrict__ dst,
                 long long d_bs, int ldd, int rows, int batches, int cols) {
  pdl_wait();  // programmatic dependent launch: predecessors complete, their writes visible
  pdl_launch_dependents();
  const int nvec = cols >> 3;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(batches) * rows * nvec;
  if (idx >= total) return;
  const int v = static_cast<int>(idx % nvec);
  const long long t = idx / nvec;
  const int r = static_cast<int>(t % rows);
  const long long b = t / rows;
  *reinterpret_cast<uint4*>(dst + b * d_bs + static_cast<long long>(r) * ldd + v * 8) =
      *reinterpret_cast<const uint4*>(src + b * s_bs + static_cast<long long>(r) * lds + v * 8);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
  });
  return fn;
}

// 3-D bf16 map: dims (inner, rows, batches), box (box_inner, box_rows, 1), zero OOB fill.
// Operand maps use 64-element (128 B) boxes with 128B swizzle; the output map uses 32-element (64 B) boxes
// with 64B swizzle (the epilogue's staging layout).
bool make_map_3d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t batches,
                 uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_inner, uint32_t box_rows,
                 CUtensorMapSwizzle swz, uint32_t box_batches = 1) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {inner, rows, batches};
  cuuint64_t strides[2] = {row_stride_elems * 2, batch_stride_elems * 2};
  cuuint32_t box[3] = {box_inner, box_rows, box_batches};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

std::atomic<long long> g_launches{0};

// Optional per-launch timing (bench.py's roofline leg): cudaEvents recorded around every GEMM launch on
// the launching stream.  Off by default; never used under graph capture.
struct ProfRec {
  cudaEvent_t e0, e1;
  int rows, batches, N, K, bn, act;
  int slot;  // index into g_prof_dev (4 words per launch), -1 = none
};
constexpr int PROF_SLOTS = 8192;
unsigned long long* g_prof_dev = nullptr;
FILE* g_prof_csv = nullptr;
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::mutex g_prof_mu;

int num_sms() { return device_num_sms(); }

}  // namespace

long long gemm_launch_count() { return g_launches.load(); }

cudaError_t gemm_set_watchdog(WdBuf* dev_ptr, unsigned long long timeout_ms) {
  cudaError_t e = dev_ptr ? wd_set_buffer_this_tu(dev_ptr) : cudaSuccess;
  if (e == cudaSuccess && timeout_ms) e = wd_set_limit_this_tu(timeout_ms);
  return e;
}

int copy_view_launch(const __nv_bfloat16* src, long long s_bs, int lds, __nv_bfloat16* dst, long long d_bs, int ldd,
                     int rows, int batches, int cols, cudaStream_t stream, const char** err) {
  if ((cols & 7) || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15) || (lds & 7) ||
      (ldd & 7) || (s_bs & 7) || (d_bs & 7)) {
    if (err) *err = "copy_view: 16-byte alignment required";
    return -1;
  }
  const long long total = static_cast<long long>(batches) * rows * (cols >> 3);
  launch_kernel(copy_view_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, src, s_bs, lds,
                dst, d_bs, ldd, rows, batches, cols);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = cudaGetErrorString(e);
    return -4;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

bool gemm_profile_enabled() { return g_prof_on; }

void gemm_profile_enable(bool on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on;
  // Optional per-launch CSV (rows,batches,N,K,bn,act,ms,TFLOP/s) for tuning: VLA_GEMM_PROF_CSV=<path>
  const char* path = getenv("VLA_GEMM_PROF_CSV");
  if (on && path && !g_prof_csv) {
    g_prof_csv = fopen(path, "a");
  }
}

// Synchronises the recorded events, returns the summed GEMM kernel time and clears the records.
int gemm_profile_read(double* total_ms, long long* launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double tot = 0.0;
  std::vector<unsigned long long> dev(g_prof.size() * 4, 0ull);
  if (g_prof_dev && !g_prof.empty()) {
    cudaDeviceSynchronize();
    const size_t n = g_prof.size() < static_cast<size_t>(PROF_SLOTS) ? g_prof.size() : static_cast<size_t>(PROF_SLOTS);
    cudaMemcpy(dev.data(), g_prof_dev, n * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  }
  for (auto& r : g_prof) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) tot += ms;
    if (g_prof_csv) {
      // columns: rows,batches,N,K,bn,act, event ms, TFLOP/s, CTA-0 lifetime [us] and the SM clock [MHz] it ran at
      const double fl = 2.0 * r.rows * r.batches * r.N * r.K;
      double us = 0.0, mhz = 0.0;
      if (r.slot >= 0) {
        const unsigned long long* d = dev.data() + 4 * r.slot;
        if (d[2] > d[0]) {
          us = (d[2] - d[0]) * 1e-3;
          mhz = static_cast<double>(d[3] - d[1]) / us;
        }
      }
      fprintf(g_prof_csv, "%d,%d,%d,%d,%d,%d,%.5f,%.1f,%.2f,%.0f\n", r.rows, r.batches, r.N, r.K, r.bn, r.act, ms,
              fl / (ms * 1e-3) / 1e12, us, mhz);
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = static_cast<long long>(g_prof.size());
  g_prof.clear();
  if (g_prof_csv) fflush(g_prof_csv);
  return 0;
}

int gemm_launch(const GemmArgs& a, cudaStream_t stream, const char** err) {
  if (!a.A || !a.W || !a.C || a.rows <= 0 || a.batches <= 0 || a.N <= 0 || a.K <= 0) {
    if (err) *err = "gemm: null pointer or non-positive shape";
    return -1;
  }
  if ((a.lda & 7) || (a.ldw & 7) || (a.N & 7) || (a.ldc & 7) || (a.a_batch_stride & 7) ||
      (a.c_batch_stride & 7) || (reinterpret_cast<uintptr_t>(a.A) & 15) ||
      (reinterpret_cast<uintptr_t>(a.W) & 15) || (reinterpret_cast<uintptr_t>(a.C) & 15)) {
    if (err) *err = "gemm: strides/N must be multiples of 8 elements and pointers 16-byte aligned";
    return -1;
  }
  if (a.resid && ((a.ldr & 7) || (a.r_batch_stride & 7) || (reinterpret_cast<uintptr_t>(a.resid) & 15))) {
    if (err) *err = "gemm: residual view must be 16-byte aligned";
    return -1;
  }
  const bool swiglu = a.act == ACT_SWIGLU;
  if (swiglu && ((a.N & 63) || a.resid)) {
    if (err) *err = "gemm: SwiGLU epilogue needs N % 64 == 0 and takes no residual";
    return -1;
  }
  if ((a.row_stats || a.colsum || a.stat_in) &&
      (a.batches != 1 || (!a.row_stats && !a.stat_in) || (a.row_stats && a.stat_in) || (swiglu && a.colsum) ||
       (reinterpret_cast<uintptr_t>(a.row_stats) & 7) || (reinterpret_cast<uintptr_t>(a.stat_in) & 15) ||
       (a.stat_in && a.stat_dim <= 0))) {
    if (err) *err = "gemm: a folded norm needs row_stats or stat_in (not both), one row view, and no mean term with SwiGLU";
    return -1;
  }
  if (a.stat_out && (a.batches != 1 || swiglu || a.rope_cols > 0 || (reinterpret_cast<uintptr_t>(a.stat_out) & 15))) {
    if (err) *err = "gemm: stat_out needs one row view and the plain epilogue";
    return -1;
  }
  static PerDeviceFlag attr_flag;  // the shared-memory opt-in is per device
  bool& attr_set = attr_flag.here();
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SMEM_BYTES) != cudaSuccess) {
      if (err) *err = "gemm: cudaFuncSetAttribute(max dynamic smem) failed";
      return -4;
    }
    attr_set = true;
  }

  // CTA pairs (256-row tiles) whenever a row view is at least one pair tall; VLA_GEMM_CG=1/2 forces the choice.
  static int cg_env = -1;
  if (cg_env < 0) {
    const char* e = getenv("VLA_GEMM_CG");
    cg_env = e ? atoi(e) : 0;
  }
  const int cg = cg_env == 1 ? 1 : (cg_env == 2 ? 2 : (a.rows >= 2 * BM ? 2 : 1));
  // Short row views (the policy's 8 / 64 rows per sample): pack 128 / rows samples into one M tile instead of
  // spending a 128-row tile on each of them.  VLA_GEMM_PACK=0 switches it off.
  static int pack_env = -1;
  if (pack_env < 0) {
    const char* e = getenv("VLA_GEMM_PACK");
    pack_env = e ? atoi(e) : 1;
  }
  int pack = 0;
  if (pack_env && cg == 1 && a.batches > 1 && a.rows < BM && (BM % a.rows) == 0 && !a.resid && a.rope_cols == 0 &&
      !swiglu && (a.rows >= 32 || (32 % a.rows) == 0))
    pack = BM / a.rows;
  const int mt_per_batch = (a.rows + BM * cg - 1) / (BM * cg);
  const int tiles_m = pack ? (a.batches + pack - 1) / pack : mt_per_batch * a.batches;
  const int sms = num_sms() / cg;  // scheduling units: CTAs or CTA pairs
  // Tile-width heuristic: fewest (rounds over the SMs) x (tile width + fixed per-tile cost).
  int bn = a.force_bn;
  if (bn != 64 && bn != 128 && bn != 192 && bn != 224 && bn != 256) {
    long long best = -1;
    // 224 = 896 / 4: Qwen's hidden size (o / down projections, projector fc2 / fc3, every policy projection) tiles
    // without the 7 % (5 x 192) or 12.5 % (4 x 256) of padded columns
    const int cands[5] = {256, 224, 192, 128, 64};
    for (int i = 0; i < 5; ++i) {
      const int c = cands[i];
      if (swiglu && (c & 127)) continue;  // a SwiGLU output box needs 64 accumulator columns per warp
      if (a.stat_out && 2 * ((a.N + c - 1) / c) > STAT_SLOTS) continue;  // one statistics slot per column tile half
      const long long tiles = static_cast<long long>(tiles_m) * ((a.N + c - 1) / c);
      const long long rounds = (tiles + sms - 1) / sms;
      // measured fixed cost per tile, in columns: ~96 for single CTAs, ~64 for CTA pairs (sig.fc2 / llm.down then take
      // 192-wide tiles that divide N = 1152 / 896 better than 256)
      const long long cost = rounds * (c + (cg == 2 ? 64 : 96));
      if (best < 0 || cost < best) {
        best = cost;
        bn = c;
      }
    }
  }
  if (swiglu && (bn & 127)) bn = 128;
  const bool rope = a.rope_cols > 0;
  if (rope) {
    if (!a.rope_cs || a.colsum || a.rope_S <= 0 || (a.rope_cols & 63) || (a.N & 63) || a.rope_cols > a.N || swiglu ||
        a.act != ACT_NONE || a.colscale || a.resid || a.batches != 1) {
      if (err) *err = "gemm: fused RoPE needs plain bias epilogue, one row view, N and rope_cols multiples of 64";
      return -1;
    }
    if (bn != 128 && bn != 256) bn = 128;  // a head (64 columns) must sit inside one epilogue warp's column range
  }

  // Residual: the epilogue adds into C with a TMA reduce-add, so C must hold the residual first.
  // Residual: in place (resid aliases C) the epilogue adds into C with a TMA reduce-add; out of place it reads the
  // residual rows itself and adds in fp32 before the bf16 store.  (A broadcast residual - batch stride 0, the ViT
  // position embedding - is pre-copied: its rows are shared by every batch.)
  const bool in_place = a.resid && a.resid == a.C && a.ldr == a.ldc && (a.batches == 1 || a.r_batch_stride == a.c_batch_stride);
  // Staged residual: in place or out of place, the residual box comes in by TMA and is added in fp32 - one rounding
  // instead of the reduce-add's two, no strided per-thread row loads, and the epilogue sees the FINAL values (which is
  // what stat_out needs).  Used when asked for (resid_staged = 1, stat_out) or with VLA_GEMM_STAGED_RESID=1; the default
  // stays the reduce-add / direct-read epilogue, which measures the same at bs=64 and is shorter at bs=1 (one tile per
  // CTA: the epilogue's latency is exposed there).
  static const int staged_env = [] {
    const char* e = getenv("VLA_GEMM_STAGED_RESID");
    return e ? atoi(e) : 0;
  }();
  const bool can_stage = a.resid && !pack && !swiglu && a.rope_cols == 0 && (a.batches == 1 || a.r_batch_stride != 0);
  const bool staged = can_stage && (a.resid_staged < 0 ? (staged_env != 0 || a.stat_out != nullptr) : a.resid_staged != 0);
  if (a.stat_out && a.resid && !staged && (in_place || a.batches != 1)) {
    // (out of place, the plain epilogue reads the residual rows itself and sees the final values too)
    if (err) *err = "gemm: stat_out with an in-place residual needs the staged-residual epilogue";
    return -1;
  }
  if (a.stat_out && 2 * ((a.N + bn - 1) / bn) > STAT_SLOTS) {
    if (err) *err = "gemm: stat_out: too many column tiles for STAT_SLOTS (tile width too small for this N)";
    return -1;
  }
  const bool fused_resid = a.resid && !staged && !in_place && (a.batches == 1 || a.r_batch_stride != 0);
  bool accumulate = in_place && !staged;
  if (a.resid && !staged && !in_place && !fused_resid) {
    const long long total = static_cast<long long>(a.batches) * a.rows * (a.N >> 3);
    launch_kernel(copy_view_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, 
        a.resid, a.r_batch_stride, a.ldr, a.C, a.c_batch_stride, a.ldc, a.rows, a.batches, a.N);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    accumulate = true;
  }

  GemmDev p;
  p.rows = a.rows;
  p.batches = a.batches;
  p.N = a.N;
  p.K = a.K;
  p.mt_per_batch = mt_per_batch;
  p.pack = pack;
  p.tiles_m = tiles_m;
  p.tiles_n = (a.N + bn - 1) / bn;
  p.bn = bn;
  const uint32_t stage_bytes = A_BYTES + static_cast<uint32_t>(bn / cg) * BK * 2;
  p.ring_bytes = staged ? RING_BYTES_STAGED : RING_BYTES;
  p.resid_tma = staged ? 1 : 0;
  p.stages = static_cast<int>(p.ring_bytes / stage_bytes);
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  p.bias = a.bias;
  p.colscale = a.colscale;
  p.act = a.act;
  p.accumulate = accumulate ? 1 : 0;
  p.resid = fused_resid ? a.resid : nullptr;
  p.r_bs = a.batches > 1 ? a.r_batch_stride : 0;
  p.ldr = a.ldr;
  p.rope_cs = a.rope_cs;
  p.rope_cols = rope ? a.rope_cols : 0;
  p.rope_S = a.rope_S;
  p.rope_ld = a.rope_ld > 0 ? a.rope_ld : a.rope_S;
  p.row_stats = reinterpret_cast<const float2*>(a.row_stats);
  p.colsum = a.colsum;
  p.prof = nullptr;
  p.stat_out = reinterpret_cast<float2*>(a.stat_out);
  p.stat_in = reinterpret_cast<const float2*>(a.stat_in);
  p.stat_inv_dim = a.stat_dim > 0 ? 1.0f / static_cast<float>(a.stat_dim) : 0.f;
  p.stat_eps = a.stat_eps;
  p.stat_rms = a.stat_rms;

  CUtensorMap mA, mB, mC, mR;
  const uint64_t a_bs = a.batches > 1 ? static_cast<uint64_t>(a.a_batch_stride)
                                      : static_cast<uint64_t>(a.rows) * a.lda;
  const uint64_t c_bs = a.batches > 1 ? static_cast<uint64_t>(a.c_batch_stride)
                                      : static_cast<uint64_t>(a.rows) * a.ldc;
  const uint64_t n_out = swiglu ? a.N / 2 : a.N;
  const uint32_t c_box_rows = pack ? static_cast<uint32_t>(a.rows < 32 ? a.rows : 32) : 32u;
  const uint32_t c_box_batches = pack ? 32u / c_box_rows : 1u;
  if (!make_map_3d(&mA, a.A, a.K, a.rows, a.batches, a.lda, a_bs, BK, pack ? a.rows : BM, CU_TENSOR_MAP_SWIZZLE_128B,
                   pack ? pack : 1) ||
      !make_map_3d(&mB, a.W, a.K, a.N, 1, a.ldw, static_cast<uint64_t>(a.N) * a.ldw, BK, bn / cg,
                   CU_TENSOR_MAP_SWIZZLE_128B) ||
      !make_map_3d(&mC, a.C, n_out, a.rows, a.batches, a.ldc, c_bs, 32, c_box_rows, CU_TENSOR_MAP_SWIZZLE_64B,
                   c_box_batches)) {
    if (err) *err = "gemm: cuTensorMapEncodeTiled failed";
    return -4;
  }
  mR = mC;
  if (staged && !in_place) {  // the residual lives elsewhere: its own (column, row, batch) view, same 32 x 32 boxes
    const uint64_t r_bs = a.batches > 1 ? static_cast<uint64_t>(a.r_batch_stride) : static_cast<uint64_t>(a.rows) * a.ldr;
    if (!make_map_3d(&mR, a.resid, n_out, a.rows, a.batches, a.ldr, r_bs, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, 1)) {
      if (err) *err = "gemm: cuTensorMapEncodeTiled failed (residual view)";
      return -4;
    }
  }

  const int total = p.tiles_m * p.tiles_n;
  const int grid = (total < sms ? total : sms) * cg;
  ProfRec rec{};
  rec.rows = p.rows; rec.batches = p.batches; rec.N = p.N; rec.K = p.K; rec.bn = bn; rec.act = p.act;
  const bool prof = g_prof_on;
  rec.slot = -1;
  if (prof) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_dev && cudaMalloc(&g_prof_dev, PROF_SLOTS * 4 * sizeof(unsigned long long)) != cudaSuccess) {
      cudaGetLastError();
      g_prof_dev = nullptr;
    }
    if (g_prof_dev && g_prof.size() < static_cast<size_t>(PROF_SLOTS)) {
      rec.slot = static_cast<int>(g_prof.size());
      p.prof = g_prof_dev + 4 * rec.slot;
    }
    cudaEventCreate(&rec.e0);
    cudaEventCreate(&rec.e1);
    cudaEventRecord(rec.e0, stream);
  }
  if (cg == 2) launch_kernel_cluster(2, gemm_bf16_tcgen05_kernel<2>, dim3(grid), dim3(GEMM_THREADS), SMEM_BYTES, stream, mA, mB, mC, mR, p);
  else launch_kernel(gemm_bf16_tcgen05_kernel<1>, dim3(grid), dim3(GEMM_THREADS), SMEM_BYTES, stream, mA, mB, mC, mR, p);
  if (prof) {
    cudaEventRecord(rec.e1, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(rec);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = cudaGetErrorString(e);
    return -4;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // namespace vla
