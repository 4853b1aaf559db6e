// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[b, r, n] = epilogue( sum_k A[b, r, k] * W[n, k] )     bf16 x bf16 -> fp32 (TMEM) -> bf16
//
// Every dense contraction of the predict_action path runs through this kernel: the ViT blocks'
// qkv/proj/fc1/fc2 (timm Block, restated in film_vit_wrapper.py:69-75), the fused projector
// (modeling_prismatic.py:261-273), Qwen2's q/k/v/o/gate/up/down, and the Bridge-Attention policy's
// q/k/v/o/ffn projections (action_heads.py:247-254, 355-367).
//
// Structure (one CTA per SM, 192 threads):
//   warp 0   : TMA producer  - cp.async.bulk.tensor (3-D maps, 128B swizzle) into a STAGES-deep ring
//   warp 1   : MMA issuer    - one thread issues tcgen05.mma (M=128, N=BN, K=16) from smem descriptors
//   warps 2-5: epilogue      - tcgen05.ld the fp32 accumulator out of TMEM, bias/act/scale/residual,
//                              cast to bf16, vectorised global stores
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1.
#include "common.cuh"
#include "gemm.cuh"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace vla {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;

struct GemmDev {
  int rows, batches, N, K;
  int mt_per_batch, tiles_m, tiles_n;
  __nv_bfloat16* C;
  long long c_bs;
  int ldc;
  const float* bias;
  const float* colscale;
  const __nv_bfloat16* resid;
  long long r_bs;
  int ldr;
  int act;
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr uint32_t A_BYTES = BM * BK * 2;
  static constexpr uint32_t B_BYTES = BN * BK * 2;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 2 * BN;  // 512 / 256 / 128: powers of two
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap mapA,
                         const __grid_constant__ CUtensorMap mapB, const GemmDev p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_addr + pad;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    mbar_fence_init();
    fence_proxy_async();
  }
  if (warp_idx == 1) {
    tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
    tmem_relinquish();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = (p.K + BK - 1) / BK;
  const int total_tiles = p.tiles_m * p.tiles_n;

  if (warp_idx == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_idx = tile % p.tiles_n;
        const int m_idx = tile / p.tiles_n;
        const int b = m_idx / p.mt_per_batch;
        const int r0 = (m_idx - b * p.mt_per_batch) * BM;
        const int n0 = n_idx * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          tma_load_3d(sa, &mapA, full_bar(stage), kb * BK, r0, b);
          tma_load_3d(sa + Cfg::A_BYTES, &mapB, full_bar(stage), kb * BK, n0, 0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t adesc = make_sw128_kmajor_desc(sa + k * 32);
            const uint64_t bdesc = make_sw128_kmajor_desc(sb + k * 32);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (4 warps, 128 rows)
    const int quarter = warp_idx & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int n_idx = tile % p.tiles_n;
      const int m_idx = tile / p.tiles_n;
      const int b = m_idx / p.mt_per_batch;
      const int r0 = (m_idx - b * p.mt_per_batch) * BM;
      const int n0 = n_idx * BN;
      const int r = r0 + quarter * 32 + lane;
      const bool row_ok = r < p.rows;

      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(acc * BN);
      if (p.act == ACT_SWIGLU) {
        __nv_bfloat16* crow = p.C + static_cast<long long>(b) * p.c_bs + static_cast<long long>(r) * p.ldc;
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          const int c0 = n0 + ch * 32;
          if (c0 >= p.N) break;
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + ch * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            uint32_t o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float g0 = __uint_as_float(v[2 * j]), g1 = __uint_as_float(v[2 * j + 1]);
              const float u0 = __uint_as_float(v[16 + 2 * j]), u1 = __uint_as_float(v[16 + 2 * j + 1]);
              o[j] = pack_bf16(silu(g0) * u0, silu(g1) * u1);
            }
            uint4* dst = reinterpret_cast<uint4*>(crow + (c0 >> 1));
            dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
            dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
          }
        }
      } else {
        __nv_bfloat16* crow = p.C + static_cast<long long>(b) * p.c_bs + static_cast<long long>(r) * p.ldc;
        const __nv_bfloat16* rrow =
            p.resid ? p.resid + static_cast<long long>(b) * p.r_bs + static_cast<long long>(r) * p.ldr
                    : nullptr;
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          const int c0 = n0 + ch * 32;
          if (c0 >= p.N) break;
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_row + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int c = c0 + q * 8;
            if (c < p.N && row_ok) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[q * 8 + j]);
              if (p.bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c + 4));
                f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
                f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
              }
              if (p.act == ACT_GELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = gelu_erf(f[j]);
              } else if (p.act == ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.0f);
              }
              if (p.colscale) {
                const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.colscale + c));
                const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.colscale + c + 4));
                f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
                f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
              }
              if (rrow) {
                const uint4 rv = *reinterpret_cast<const uint4*>(rrow + c);
                const float2 r0v = unpack_bf16(rv.x), r1v = unpack_bf16(rv.y);
                const float2 r2v = unpack_bf16(rv.z), r3v = unpack_bf16(rv.w);
                f[0] += r0v.x; f[1] += r0v.y; f[2] += r1v.x; f[3] += r1v.y;
                f[4] += r2v.x; f[5] += r2v.y; f[6] += r3v.x; f[7] += r3v.y;
              }
              *reinterpret_cast<uint4*>(crow + c) =
                  make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                             pack_bf16(f[6], f[7]));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
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

// 3-D bf16 map: dims (inner, rows, batches), box (64, box_rows, 1), 128B swizzle, zero OOB fill.
bool make_map_3d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t rows, uint64_t batches,
                 uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t dims[3] = {inner, rows, batches};
  cuuint64_t strides[2] = {row_stride_elems * 2, batch_stride_elems * 2};
  cuuint32_t box[3] = {BK, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

std::atomic<long long> g_launches{0};

// Optional per-launch timing (bench.py's roofline leg): cudaEvents recorded around every GEMM launch on
// the launching stream.  Off by default; never used under graph capture.
struct ProfRec {
  cudaEvent_t e0, e1;
  int rows, batches, N, K, bn, act;
};
FILE* g_prof_csv = nullptr;
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::mutex g_prof_mu;

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN>
int launch_bn(const CUtensorMap& mA, const CUtensorMap& mB, const GemmDev& p, cudaStream_t stream,
              const char** err) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg::SMEM_BYTES) != cudaSuccess) {
      if (err) *err = "gemm: cudaFuncSetAttribute(max dynamic smem) failed";
      return -4;
    }
    attr_set = true;
  }
  const int total = p.tiles_m * p.tiles_n;
  const int grid = total < num_sms() ? total : num_sms();
  ProfRec rec{};
  rec.rows = p.rows; rec.batches = p.batches; rec.N = p.N; rec.K = p.K; rec.bn = BN; rec.act = p.act;
  const bool prof = g_prof_on;
  if (prof) {
    cudaEventCreate(&rec.e0);
    cudaEventCreate(&rec.e1);
    cudaEventRecord(rec.e0, stream);
  }
  gemm_bf16_tcgen05_kernel<BN><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(mA, mB, p);
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

}  // namespace

long long gemm_launch_count() { return g_launches.load(); }

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
  for (auto& r : g_prof) {
    cudaEventSynchronize(r.e1);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) tot += ms;
    if (g_prof_csv) {
      const double fl = 2.0 * r.rows * r.batches * r.N * r.K;
      fprintf(g_prof_csv, "%d,%d,%d,%d,%d,%d,%.5f,%.1f\n", r.rows, r.batches, r.N, r.K, r.bn, r.act, ms,
              fl / (ms * 1e-3) / 1e12);
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
  if (a.act == ACT_SWIGLU && (a.N & 31)) {
    if (err) *err = "gemm: SwiGLU epilogue needs N % 32 == 0";
    return -1;
  }

  const int mt_per_batch = (a.rows + BM - 1) / BM;
  const int tiles_m = mt_per_batch * a.batches;
  // Tile-width heuristic: fewest "waves x tile width" over the SM count.
  int bn = a.force_bn;
  if (bn != 64 && bn != 128 && bn != 256) {
    const int sms = num_sms();
    long long best = -1;
    const int cands[3] = {256, 128, 64};
    for (int i = 0; i < 3; ++i) {
      const int c = cands[i];
      const long long tiles = static_cast<long long>(tiles_m) * ((a.N + c - 1) / c);
      const long long waves = (tiles + sms - 1) / sms;
      const long long cost = waves * (c + 24);
      if (best < 0 || cost < best) {
        best = cost;
        bn = c;
      }
    }
  }

  GemmDev p;
  p.rows = a.rows;
  p.batches = a.batches;
  p.N = a.N;
  p.K = a.K;
  p.mt_per_batch = mt_per_batch;
  p.tiles_m = tiles_m;
  p.tiles_n = (a.N + bn - 1) / bn;
  p.C = a.C;
  p.c_bs = a.c_batch_stride;
  p.ldc = a.ldc;
  p.bias = a.bias;
  p.colscale = a.colscale;
  p.resid = a.resid;
  p.r_bs = a.r_batch_stride;
  p.ldr = a.ldr;
  p.act = a.act;

  CUtensorMap mA, mB;
  const uint64_t a_bs = a.batches > 1 ? static_cast<uint64_t>(a.a_batch_stride)
                                      : static_cast<uint64_t>(a.rows) * a.lda;
  if (!make_map_3d(&mA, a.A, a.K, a.rows, a.batches, a.lda, a_bs, BM) ||
      !make_map_3d(&mB, a.W, a.K, a.N, 1, a.ldw, static_cast<uint64_t>(a.N) * a.ldw, bn)) {
    if (err) *err = "gemm: cuTensorMapEncodeTiled failed";
    return -4;
  }
  if (bn == 256) return launch_bn<256>(mA, mB, p, stream, err);
  if (bn == 128) return launch_bn<128>(mA, mB, p, stream, err);
  return launch_bn<64>(mA, mB, p, stream, err);
}

}  // namespace vla
