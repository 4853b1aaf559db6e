// The engine behind include/vla_b200.h: weight intake/repacking, workspace, and the batched
// predict_action forward (vision towers -> projector -> Qwen2.5 prefill with per-layer taps ->
// Bridge-Attention policy -> un-normalised action chunks) as one stream of own kernels.
//
// Reference call stack being replaced: OpenVLAForActionPrediction.predict_action
// (prismatic/extern/hf/modeling_prismatic.py:892-972) -> _process_vision_features (:463) ->
// _regression_or_discrete_prediction (:808-889) -> L1RegressionActionHead.predict_action
// (prismatic/models/action_heads.py:43-81) -> _unnormalize_actions (:786-805).
#include "../../include/vla_b200.h"
#include "gemm.cuh"
#include "launch.cuh"
#include "ops.cuh"
#include "policy_fused.cuh"
#include "watchdog.cuh"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

using bf16 = __nv_bfloat16;

namespace {

constexpr int D_DINO = 1024, F_DINO = 4096, TOK_DINO = 261, PRE_DINO = 5;
constexpr int D_SIG = 1152, F_SIG = 4304, TOK_SIG = 256;
constexpr int D_VIS = D_DINO + D_SIG;  // 2176
constexpr int D_PROJ1 = 4 * D_VIS;     // 8704
constexpr int D_LLM = 896, I_LLM = 4864, QKV_LLM = 1152, HQ = 14, HKV = 2;
constexpr int N_AQ = 64;
constexpr int KP = 592;  // patch vector (588) padded to a multiple of 8
constexpr int PKV = 1792;
constexpr float LLM_EPS = 1e-6f, VIT_EPS = 1e-6f, HEAD_EPS = 1e-5f, ROPE_THETA = 1e6f;

// A loaded tensor, kept in the caller's dtype until vla_finalize repacks it (then freed).
struct Master {
  void* d = nullptr;
  int dtype = VLA_F32;
  std::vector<int64_t> shape;
  size_t n = 0;
};

struct VitBlock {
  float *ln1w, *ln1b, *bqkv, *bproj, *ls1, *ln2w, *ln2b, *bfc1, *bfc2, *ls2;
  bf16 *wqkv, *wproj, *wfc1, *wfc2;
  float *cs_qkv = nullptr, *cs_fc1 = nullptr;  // column sums of the norm-folded wqkv / wfc1 (fold_norms)
};
struct Tower {
  int D, F, heads, hd, tokens, prefix, depth;
  bf16* wpatch;
  float* bpatch;
  bf16* pos;
  bf16* prefix_rows;
  std::vector<VitBlock> blocks;
};
struct LlmLayer {
  float *ln1, *bqkv, *ln2;
  bf16 *wqkv, *wo, *wgu, *wdown;
};
struct HeadBlock {
  bf16 *wq, *wkv_self, *wkv_cond, *wkv_vis, *wo, *wffn;
  float *bq, *bkv_self, *bkv_cond, *bkv_vis, *bo, *ffn_lnw, *ffn_lnb, *bffn;
  float* gatevec;  // [1792]: tanh(gating_factor) on the K half, 1 on the V half (epilogue column scale)
  float gate;
};

// One repack job of vla_finalize: dst[(r / group) * stride + offset + r % group, c] = cast(src[r, c]).  All jobs of a
// finalize run in ONE kernel launch (pack_jobs_kernel), split into chunks of PACK_CHUNK elements.
struct PackJob {
  const void* src;
  void* dst;
  int sdtype;      // vla_dtype of src
  int dst_f32;     // 1: fp32 destination (biases, norm weights), 0: bf16 (GEMM operands)
  int rows, cols, dst_ld, group, stride, offset;
};
constexpr unsigned int PACK_CHUNK = 16384;

__global__ void __launch_bounds__(256)
pack_jobs_kernel(const PackJob* __restrict__ jobs, const uint2* __restrict__ chunks) {
  const uint2 ch = chunks[blockIdx.x];  // (job, first element / PACK_CHUNK)
  const PackJob j = jobs[ch.x];
  const size_t total = static_cast<size_t>(j.rows) * j.cols;
  const size_t lo = static_cast<size_t>(ch.y) * PACK_CHUNK;
  const size_t hi = lo + PACK_CHUNK < total ? lo + PACK_CHUNK : total;
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float v;
    if (j.sdtype == VLA_BF16) v = __bfloat162float(static_cast<const bf16*>(j.src)[i]);
    else if (j.sdtype == VLA_F16) v = __half2float(static_cast<const __half*>(j.src)[i]);
    else v = static_cast<const float*>(j.src)[i];
    const int r = static_cast<int>(i / j.cols), c = static_cast<int>(i % j.cols);
    const size_t o = static_cast<size_t>((r / j.group) * j.stride + j.offset + r % j.group) * j.dst_ld + c;
    if (j.dst_f32) static_cast<float*>(j.dst)[o] = v;
    else static_cast<bf16*>(j.dst)[o] = __float2bfloat16_rn(v);
  }
}

}  // namespace

struct vla_engine {
  vla_cfg cfg;
  std::string err;
  bool finalized = false;
  int device = 0;  // the CUDA device this engine lives on: made current at every entry point
  std::unordered_map<std::string, Master> masters;  // loaded tensors in their source dtype; freed by vla_finalize
  std::vector<PackJob> jobs;
  std::vector<void*> allocs;

  // derived
  int NP = 0, T = 0, A = 0, P = 0;
  Tower dino, sig;
  bf16 *pj_w1, *pj_w2, *pj_w3;
  float *pj_b1, *pj_b2, *pj_b3;
  bf16 *embed, *aq_table;
  std::vector<LlmLayer> llm;
  float* llm_norm;
  float *rope_cos = nullptr, *rope_sin = nullptr;
  uint32_t* rope_cs = nullptr;  // the same table transposed and packed for the GEMM's RoPE epilogue
  // head
  std::vector<HeadBlock> head;
  bf16 *x0, *head_fc2_w, *pp_w1, *pp_w2;
  bf16* wkv_cond_all = nullptr;   // [24 * 1792, 896]: every block's K|V projection of the cond rows, for the proprio row
  float* bkv_cond_all = nullptr;  // [24 * 1792]
  bf16* h_pkv = nullptr;          // [B, 24 * 1792]: K|V of the proprio row for all 24 blocks (one GEMM per forward)
  float *head_ln2w, *head_ln2b, *head_fc2_b, *pp_b1, *pp_b2;
  float *prope_cos = nullptr, *prope_sin = nullptr;
  bf16* img_lut = nullptr;  // [2 towers][3 channels][256]: ToTensor + Normalize + bf16 of a uint8 pixel
  float crop_scale = 0.f;   // > 0: uint8 frames are centre-cropped on the device first (vla_set_center_crop)
  uint8_t* crop_buf = nullptr;
  float *st_hi = nullptr, *st_lo = nullptr;
  uint8_t* st_mask = nullptr;
  bool stats_set = false;

  // workspace
  int maxB = 0, maxL = 0, maxS = 0;
  struct TowerWs {
    bf16 *col, *x, *xn, *qkv, *attn, *h;
    float* stats;  // (rstd, -mean * rstd) per row for the GEMM that follows a folded norm
  };
  TowerWs tw[2];  // one workspace per vision tower: at small batch the towers run concurrently on two streams
  bf16 *w_patches, *w_ph1, *w_ph2;
  std::vector<bf16*> hid;  // 25 LLM states
  bf16 *l_tmp, *l_xn, *l_qkv, *l_attn, *l_act;
  bf16 *h_p1, *h_p, *h_kv, *h_q, *h_ao, *h_y, *h_yn;
  // prompts of different lengths in one batch (vla_predict's prompt_len): the 64 h_a rows of every sample are gathered
  // into a dense buffer per policy block, because their row offset then differs from sample to sample
  std::vector<bf16*> h_ha;
  // small-batch mode (B <= small_B): every policy block has its own K|V buffer so that the K|V projections of the
  // LLM states run on a side stream as soon as each LLM layer finishes, off the policy's critical path
  int small_B = 0;
  std::vector<bf16*> h_kv_blk;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<cudaEvent_t> ev_layer, ev_kv;
  std::vector<bf16*> head_x;  // 25 policy states
  // small-batch mode: the 24 policy blocks run as ONE cluster kernel (policy_fused.cu); per-block pointers on the device
  std::vector<vla::PolicyBlockW> pol_blocks;
  int* pol_progress = nullptr;
  long long* pol_prof = nullptr;  // VLA_POLICY_PROF=1: phase stamps of the fused policy kernel (printed by vla_destroy)
  int policy_fused = 1;  // VLA_NO_POLICY_FUSED=1 keeps the per-block launches
  // The fused policy kernel is launched in groups of `policy_group` blocks on its own stream, each as soon as the K|V
  // rows of its last block are projected: the policy runs BESIDE the LLM prefill and only the last group is left when
  // the prefill ends.  All dependencies are ordinary stream / graph edges (no kernel ever waits for a later kernel).
  // VLA_POLICY_GROUP=24 is the single launch after the prefill.
  int policy_group = 2;  // measured at bs=1: 2 -> 4.07 ms, 3 -> 4.12, 4 -> 4.13, 6 -> 4.19, 8 -> 4.24, 24 (one launch) -> 4.72
  cudaStream_t pol = nullptr;
  cudaEvent_t ev_pol = nullptr;
  int* err_flag = nullptr;
  // pinned/dev staging for vla_predict_host
  void *pin_pix = nullptr, *pin_ids = nullptr, *pin_aq = nullptr, *pin_prop = nullptr, *pin_out = nullptr,
       *pin_ha = nullptr;
  void *pin_u8 = nullptr, *dev_u8 = nullptr;
  void *dev_pix = nullptr, *dev_ids = nullptr, *dev_aq = nullptr, *dev_prop = nullptr, *dev_out = nullptr,
       *dev_ha = nullptr, *dev_len = nullptr;

  // last call
  int lastB = 0, lastL = 0;
  long long last_launches = 0;

  // CUDA graphs of forward(): one per distinct (B, L, buffer pointers), captured on the second call with that
  // key (the first runs eagerly and performs the one-time cudaFuncSetAttribute calls).  Replayed on an internal
  // stream fenced to the caller's stream with events, so the legacy default stream works too.
  struct GraphKey {
    int B, L, u8;
    const void *pix, *ids, *aq, *prop, *len, *out_norm, *out_unnorm, *out_ha;
    bool operator==(const GraphKey& o) const {
      return B == o.B && L == o.L && u8 == o.u8 && pix == o.pix && ids == o.ids && aq == o.aq && prop == o.prop &&
             len == o.len && out_norm == o.out_norm && out_unnorm == o.out_unnorm && out_ha == o.out_ha;
    }
  };
  struct GraphEntry {
    GraphKey key;
    int seen = 0;
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;
  };
  std::vector<GraphEntry> graphs;
  cudaStream_t gstream = nullptr;
  cudaEvent_t gev_in = nullptr, gev_out = nullptr;
  int use_graphs = 1;
  // LayerNorm / RMSNorm in front of a GEMM folded into that GEMM (weights pre-scaled at finalize, per-row statistics
  // applied in its epilogue): the normalised activations are never written.  VLA_NO_NORM_FOLD=1 keeps the norm kernels.
  int fold_norms = 1;
  // ... and, optionally (VLA_STAT_FUSE=1), the row statistics of those norms taken from partial sums the PRODUCING
  // GEMM leaves in its epilogue (gemm.cuh: stat_out / stat_in) instead of a statistics kernel per norm: 143 of the 146
  // statistics launches of a step disappear.  OFF by default: measured on B200 at bs=64 (same box, A/B/A/B) the step
  // takes 83.8 ms either way - the chip is ENERGY-limited there (every GEMM's in-kernel SM clock drops from ~1250 to
  // ~1130 MHz once the low-power statistics kernels no longer give the power budget a rest), the staged-residual
  // epilogue it needs costs what the statistics kernels cost, and at bs=1 its longer epilogue adds 0.37 ms.
  int stat_fuse = 0;
  float* l_stats = nullptr;
  // segment timing (vla_segment_times): events at the subsystem boundaries of the last EAGER forward
  int seg_on = 0;
  cudaEvent_t seg_ev[4] = {nullptr, nullptr, nullptr, nullptr};

  int fail(int code, const std::string& m) {
    err = m;
    if (code == VLA_ERR_CUDA) {  // a trapped kernel: say which barrier the watchdog caught
      const std::string wd = vla::watchdog_report();
      if (!wd.empty()) err += "\n" + wd;
    }
    return code;
  }
  template <class Tp>
  Tp* dalloc(size_t count) {
    void* p = nullptr;
    size_t bytes = count * sizeof(Tp);
    if (bytes == 0) bytes = 16;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      throw std::runtime_error("cudaMalloc of " + std::to_string(bytes) + " bytes failed");
    }
    allocs.push_back(p);
    return static_cast<Tp*>(p);
  }
  const Master& need(const std::string& name, size_t numel) {
    auto it = masters.find(name);
    if (it == masters.end()) throw std::runtime_error("missing tensor: " + name);
    if (it->second.n != numel)
      throw std::runtime_error("tensor " + name + " has " + std::to_string(it->second.n) + " elements, expected " +
                               std::to_string(numel));
    return it->second;
  }
  // ---- repack jobs: queued while vla_finalize walks the architecture, executed by ONE kernel launch (flush_jobs)
  void enqueue(const Master& m, void* dst, bool dst_f32, int rows, int cols, int dst_ld, int group, int stride,
               int offset) {
    PackJob j;
    j.src = m.d; j.dst = dst; j.sdtype = m.dtype; j.dst_f32 = dst_f32 ? 1 : 0;
    j.rows = rows; j.cols = cols; j.dst_ld = dst_ld; j.group = group; j.stride = stride; j.offset = offset;
    jobs.push_back(j);
  }
  void flush_jobs() {
    if (jobs.empty()) return;
    std::vector<uint2> chunks;
    for (size_t i = 0; i < jobs.size(); ++i) {
      const size_t total = static_cast<size_t>(jobs[i].rows) * jobs[i].cols;
      for (size_t c = 0; c * PACK_CHUNK < total; ++c) chunks.push_back(make_uint2(static_cast<unsigned>(i), static_cast<unsigned>(c)));
    }
    PackJob* dj = nullptr;
    uint2* dc = nullptr;
    cudaError_t ce = cudaMalloc(&dj, jobs.size() * sizeof(PackJob));
    if (ce == cudaSuccess) ce = cudaMalloc(&dc, chunks.size() * sizeof(uint2));
    if (ce == cudaSuccess) ce = cudaMemcpy(dj, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(dc, chunks.data(), chunks.size() * sizeof(uint2), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) {
      pack_jobs_kernel<<<static_cast<unsigned>(chunks.size()), 256>>>(dj, dc);
      ce = cudaDeviceSynchronize();
    }
    if (dj) cudaFree(dj);
    if (dc) cudaFree(dc);
    jobs.clear();
    if (ce != cudaSuccess) {
      cudaGetLastError();
      throw std::runtime_error(std::string("weight repack failed: ") + cudaGetErrorString(ce));
    }
  }
  // fp32 copy of a (small) tensor: biases, norm weights, LayerScale.  A fresh buffer per call - the norm fold adds into
  // it, and a retried vla_finalize must start from the loaded values again.
  float* f32(const std::string& name, size_t numel) {
    const Master& m = need(name, numel);
    float* dst = dalloc<float>(numel);
    enqueue(m, dst, true, 1, static_cast<int>(numel), static_cast<int>(numel), 1, 0, 0);
    return dst;
  }
  // pack one [rows, cols] tensor into dst (bf16) at a row mapping
  void pack_into(bf16* dst, int dst_ld, const std::string& name, int rows, int cols, int group, int stride,
                 int offset) {
    enqueue(need(name, static_cast<size_t>(rows) * cols), dst, false, rows, cols, dst_ld, group, stride, offset);
  }
  bf16* pack(const std::string& name, int rows, int cols, int dst_ld = 0) {
    if (!dst_ld) dst_ld = cols;
    bf16* dst = dalloc<bf16>(static_cast<size_t>(rows) * dst_ld);
    if (dst_ld != cols) cudaMemset(dst, 0, static_cast<size_t>(rows) * dst_ld * sizeof(bf16));
    pack_into(dst, dst_ld, name, rows, cols, rows, 0, 0);
    return dst;
  }
  // concatenates [name_i (rows_i x cols)] along rows
  bf16* pack_cat(const std::vector<std::pair<std::string, int>>& parts, int cols) {
    int total = 0;
    for (auto& p : parts) total += p.second;
    bf16* dst = dalloc<bf16>(static_cast<size_t>(total) * cols);
    int off = 0;
    for (auto& p : parts) {
      pack_into(dst, cols, p.first, p.second, cols, p.second, 0, off);
      off += p.second;
    }
    return dst;
  }
  float* cat_f32(const std::vector<std::pair<std::string, int>>& parts) {
    int total = 0;
    for (auto& p : parts) total += p.second;
    float* dst = dalloc<float>(total);
    int off = 0;
    for (auto& p : parts) {
      enqueue(need(p.first, p.second), dst + off, true, 1, p.second, p.second, 1, 0, 0);
      off += p.second;
    }
    return dst;
  }
  void free_masters() {
    for (auto& kv : masters)
      if (kv.second.d) cudaFree(kv.second.d);
    masters.clear();
  }
};

namespace {

bool ignored_name(const std::string& n) {
  static const char* pats[] = {"lm_head.", "film_gen.", "attn_pool.", "featurizer.norm.", "inv_freq",
                               "noisy_action", "rope."};
  for (const char* p : pats)
    if (n.find(p) != std::string::npos) return true;
  return false;
}

void build_tower(vla_engine* e, Tower& t, const std::string& pfx, bool is_dino, int depth) {
  t.D = is_dino ? D_DINO : D_SIG;
  t.F = is_dino ? F_DINO : F_SIG;
  t.heads = 16;
  t.hd = t.D / 16;
  t.tokens = is_dino ? TOK_DINO : TOK_SIG;
  t.prefix = is_dino ? PRE_DINO : 0;
  t.depth = depth;
  const int D = t.D, F = t.F;
  t.wpatch = e->pack(pfx + "patch_embed.proj.weight", D, 588, KP);
  t.bpatch = e->f32(pfx + "patch_embed.proj.bias", D);
  t.pos = e->pack(pfx + "pos_embed", 256, D);
  t.prefix_rows = nullptr;
  if (is_dino) t.prefix_rows = e->pack_cat({{pfx + "cls_token", 1}, {pfx + "reg_token", 4}}, D);
  // Only blocks 0 .. depth-2 contribute to the output (modeling_prismatic.py:141-142).
  for (int i = 0; i <= depth - 2; ++i) {
    const std::string b = pfx + "blocks." + std::to_string(i) + ".";
    VitBlock k;
    k.ln1w = e->f32(b + "norm1.weight", D);
    k.ln1b = e->f32(b + "norm1.bias", D);
    k.wqkv = e->pack(b + "attn.qkv.weight", 3 * D, D);
    k.bqkv = e->f32(b + "attn.qkv.bias", 3 * D);
    k.wproj = e->pack(b + "attn.proj.weight", D, D);
    k.bproj = e->f32(b + "attn.proj.bias", D);
    k.ls1 = is_dino ? e->f32(b + "ls1.scale_factor", D) : nullptr;
    k.ln2w = e->f32(b + "norm2.weight", D);
    k.ln2b = e->f32(b + "norm2.bias", D);
    k.wfc1 = e->pack(b + "mlp.fc1.weight", F, D);
    k.bfc1 = e->f32(b + "mlp.fc1.bias", F);
    k.wfc2 = e->pack(b + "mlp.fc2.weight", D, F);
    k.bfc2 = e->f32(b + "mlp.fc2.bias", D);
    k.ls2 = is_dino ? e->f32(b + "ls2.scale_factor", D) : nullptr;
    t.blocks.push_back(k);
  }
}

#define CK(call)                                   \
  do {                                             \
    const char* _err = nullptr;                    \
    int _rc = (call);                              \
    if (_rc) return e->fail(_rc, _err ? _err : "kernel launch failed"); \
  } while (0)

// Where the row statistics of a folded norm come from:
//   STATS_KERNEL  one row_stats launch per norm; the residual GEMMs add in place by TMA reduce-add (large batches:
//                 the cheapest in energy, which is what bounds the bs=64 step)
//   STATS_STAGED  partial sums from the producing GEMM with the staged-residual epilogue, in place (VLA_STAT_FUSE=1):
//                 143 launches fewer, the same 83.8 ms at bs=64, +0.37 ms at bs=1
// (A third form - partial sums from the PLAIN epilogue with the residual GEMMs writing out of place, x -> x' -> x, so
// that the epilogue sees final values without staging - was built and measured for the launch-bound small batches:
// bs=1 p50 5.27 ms against 4.93 ms with the statistics kernels.  A 2 us statistics kernel under PDL is cheaper than the
// longer producer epilogue; removed.)
enum StatsMode { STATS_KERNEL = 0, STATS_STAGED = 1 };
static StatsMode stats_mode(const vla_engine* e, int) {
  return (e->fold_norms && e->stat_fuse) ? STATS_STAGED : STATS_KERNEL;
}

int run_tower(vla_engine* e, const Tower& t, const bf16* pix, const uint8_t* pix_u8, int B, int tower_idx,
              cudaStream_t s) {
  const int n = e->cfg.n_images, slabs = B * n, D = t.D, F = t.F;
  const int M = slabs * t.tokens;
  const vla_engine::TowerWs& ws = e->tw[tower_idx];
  bf16* x = ws.x;
  const StatsMode sm = stats_mode(e, B);
  if (pix_u8) CK(vla::im2col_u8_launch(pix_u8, B, n, tower_idx, e->img_lut, ws.col, s, &_err));
  else CK(vla::im2col_launch(pix, B, n, tower_idx, ws.col, s, &_err));
  if (t.prefix) CK(vla::prefix_tokens_launch(x, slabs, static_cast<long long>(t.tokens) * D, D, t.prefix_rows, t.prefix, s, &_err));
  {
    // patch embed (+bias +pos_embed) written behind the prefix rows of every image slab
    vla::GemmArgs g;
    g.A = ws.col; g.a_batch_stride = 256LL * KP; g.lda = KP; g.rows = 256; g.batches = slabs;
    g.W = t.wpatch; g.ldw = KP; g.N = D; g.K = KP;
    g.C = x + static_cast<long long>(t.prefix) * D; g.c_batch_stride = static_cast<long long>(t.tokens) * D; g.ldc = D;
    g.bias = t.bpatch; g.resid = t.pos; g.r_batch_stride = 0; g.ldr = D;
    CK(vla::gemm_launch(g, s, &_err));
  }
  const int nblk = static_cast<int>(t.blocks.size());
  for (int i = 0; i < nblk; ++i) {
    const VitBlock& k = t.blocks[i];
    const bool last = (i == nblk - 1);
    vla::GemmArgs g;
    if (e->fold_norms && sm != STATS_KERNEL) {
      // norm1 lives in wqkv / bqkv / cs_qkv; the row statistics come from the partial sums the previous block's fc2
      // GEMM left (block 0: from one statistics pass over the patch-embed output)
      if (i == 0) CK(vla::row_stats_launch(x, M, D, D, 0, VIT_EPS, ws.stats, s, &_err, vla::STAT_SLOTS));
      g.A = x; g.stat_in = ws.stats; g.stat_dim = D; g.stat_eps = VIT_EPS; g.colsum = k.cs_qkv;
    } else if (e->fold_norms) {  // norm1 lives in wqkv / bqkv / cs_qkv; only the row statistics are computed here
      CK(vla::row_stats_launch(x, M, D, D, 0, VIT_EPS, ws.stats, s, &_err));
      g.A = x; g.row_stats = ws.stats; g.colsum = k.cs_qkv;
    } else {
      CK(vla::layernorm_launch(x, M, D, D, k.ln1w, k.ln1b, VIT_EPS, ws.xn, D, s, &_err));
      g.A = ws.xn;
    }
    g.lda = D; g.rows = M; g.W = k.wqkv; g.ldw = D; g.N = 3 * D; g.K = D;
    g.C = ws.qkv; g.ldc = 3 * D; g.bias = k.bqkv;
    CK(vla::gemm_launch(g, s, &_err));
    CK(vla::attention_launch(ws.qkv, 3 * D, 0, D, 2 * D, slabs, t.tokens, t.heads, 1, t.hd, 0, ws.attn, D, s, &_err));
    g = vla::GemmArgs();
    g.A = ws.attn; g.lda = D; g.rows = M; g.W = k.wproj; g.ldw = D; g.N = D; g.K = D;
    g.C = x; g.ldc = D; g.bias = k.bproj; g.colscale = k.ls1; g.resid = x; g.ldr = D;
    if (sm != STATS_KERNEL) g.stat_out = ws.stats;  // statistics of the new x for norm2, from this GEMM's epilogue
    CK(vla::gemm_launch(g, s, &_err));
    g = vla::GemmArgs();
    if (e->fold_norms && sm != STATS_KERNEL) {
      g.A = x; g.stat_in = ws.stats; g.stat_dim = D; g.stat_eps = VIT_EPS; g.colsum = k.cs_fc1;
    } else if (e->fold_norms) {
      CK(vla::row_stats_launch(x, M, D, D, 0, VIT_EPS, ws.stats, s, &_err));
      g.A = x; g.row_stats = ws.stats; g.colsum = k.cs_fc1;
    } else {
      CK(vla::layernorm_launch(x, M, D, D, k.ln2w, k.ln2b, VIT_EPS, ws.xn, D, s, &_err));
      g.A = ws.xn;
    }
    g.lda = D; g.rows = M; g.W = k.wfc1; g.ldw = D; g.N = F; g.K = D;
    g.C = ws.h; g.ldc = F; g.bias = k.bfc1; g.act = vla::ACT_GELU;
    CK(vla::gemm_launch(g, s, &_err));
    g = vla::GemmArgs();
    g.W = k.wfc2; g.ldw = F; g.N = D; g.K = F; g.bias = k.bfc2; g.colscale = k.ls2;
    if (!last) {
      g.A = ws.h; g.lda = F; g.rows = M;
      g.C = x; g.ldc = D; g.resid = x; g.ldr = D;
      if (sm != STATS_KERNEL) g.stat_out = ws.stats;  // ... and of the block's output for the next block's norm1
      } else {
      // Output block: only the patch rows (prefix stripped, film_vit_wrapper.py:162) go to the
      // feature-concatenated buffer (modeling_prismatic.py:233, 237).
      g.A = ws.h + static_cast<long long>(t.prefix) * F; g.a_batch_stride = static_cast<long long>(t.tokens) * F;
      g.lda = F; g.rows = 256; g.batches = slabs;
      g.resid = x + static_cast<long long>(t.prefix) * D; g.r_batch_stride = static_cast<long long>(t.tokens) * D; g.ldr = D;
      g.C = e->w_patches + (tower_idx == 0 ? 0 : D_DINO); g.c_batch_stride = 256LL * D_VIS; g.ldc = D_VIS;
    }
    CK(vla::gemm_launch(g, s, &_err));
  }
  return 0;
}

// K|V projections of one policy block that depend only on LLM hidden state i+1: the 64 ActionQuery rows -> kv rows
// [T, T+64) and the NP raw rows -> kv rows [T+65, NK), K columns of the latter scaled by tanh(gating_factor)
// (AH:269 / AH:391).
int policy_kv_gemms(vla_engine* e, int i, const bf16* hs, int B, int L, const int32_t* prompt_len, bf16* kvb,
                    cudaStream_t s) {
  const int NP = e->NP, T = e->T;
  const int S = NP + L + N_AQ + 1;
  const int ha_row0 = NP + L - 1;  // MP:855 with NUM_PROMPT_TOKENS = L-1 (MP:927)
  const long long kv_bs = static_cast<long long>(T + N_AQ + 1 + NP) * PKV;
  const HeadBlock& w = e->head[i];
  vla::GemmArgs g;
  if (prompt_len) {  // per-sample offsets: rows [NP + L_b - 1, NP + L_b + 63) of sample b, gathered first
    CK(vla::gather_rows_launch(hs, static_cast<long long>(S) * D_LLM, D_LLM, ha_row0, N_AQ, B, D_LLM, e->h_ha[i], s, &_err,
                               prompt_len, L, e->err_flag));
    g.A = e->h_ha[i]; g.a_batch_stride = static_cast<long long>(N_AQ) * D_LLM;
  } else {
    g.A = hs + static_cast<long long>(ha_row0) * D_LLM; g.a_batch_stride = static_cast<long long>(S) * D_LLM;
  }
  g.lda = D_LLM;
  g.rows = N_AQ; g.batches = B; g.W = w.wkv_cond; g.ldw = D_LLM; g.N = PKV; g.K = D_LLM;
  g.C = kvb + static_cast<long long>(T) * PKV; g.c_batch_stride = kv_bs; g.ldc = PKV; g.bias = w.bkv_cond;
  CK(vla::gemm_launch(g, s, &_err));
  g = vla::GemmArgs();
  g.A = hs; g.a_batch_stride = static_cast<long long>(S) * D_LLM; g.lda = D_LLM; g.rows = NP; g.batches = B;
  g.W = w.wkv_vis; g.ldw = D_LLM; g.N = PKV; g.K = D_LLM;
  g.C = kvb + static_cast<long long>(T + N_AQ + 1) * PKV; g.c_batch_stride = kv_bs; g.ldc = PKV;
  g.bias = w.bkv_vis; g.colscale = w.gatevec;
  CK(vla::gemm_launch(g, s, &_err));
  return 0;
}

// Small-batch mode: the same projections on the side stream, ordered after LLM layer i by an event on the main stream
// and announced to the policy loop by ev_kv[i].
int policy_kv_from_llm(vla_engine* e, int i, const bf16* hs, int B, int L, const int32_t* prompt_len, cudaStream_t s,
                       cudaStream_t s2, bool fused_pro) {
  if (cudaEventRecord(e->ev_layer[i], s) != cudaSuccess || cudaStreamWaitEvent(s2, e->ev_layer[i], 0) != cudaSuccess)
    return e->fail(VLA_ERR_CUDA, "policy K|V fork failed");
  const int rc = policy_kv_gemms(e, i, hs, B, L, prompt_len, e->h_kv_blk[i], s2);
  if (rc) return rc;
  if (fused_pro) {  // the fused policy kernel rotates q and the self keys itself; these rows are rotated here, once
    const char* _err = nullptr;
    const int rr = vla::policy_rope_launch(nullptr, e->h_kv_blk[i], B, e->T, e->NP, e->prope_cos, e->prope_sin, s2, &_err, 1);
    if (rr) return e->fail(rr, _err ? _err : "policy rope failed");
  }
  if (cudaEventRecord(e->ev_kv[i], s2) != cudaSuccess) return e->fail(VLA_ERR_CUDA, "policy K|V event failed");
  return 0;
}

// prompt_len (device, nullable): per-sample prompt lengths L_b <= L.  The LLM sequences are then RIGHT-padded to the
// common length NP + L + 65: sample b holds [tok0 | patches | tok1..tok_{L_b-1} | AQ0..63 | stop | padding].  Causal
// attention never lets a real row see the padding behind it, so every real row equals the row of the un-padded run;
// only the policy's h_a window (and the returned last-layer states) sits at a per-sample offset.
int forward(vla_engine* e, const bf16* pix, const uint8_t* pix_u8, const int64_t* ext_ids, const int32_t* aq_index, const float* proprio,
            const int32_t* prompt_len, int B, int L, float* out_norm, float* out_unnorm, bf16* out_last_ha, cudaStream_t s) {
  const int NP = e->NP, T = e->T, A = e->A, P = e->P;
  vla::PdlScope pdl(B <= 8);  // programmatic dependent launch pays off only when kernels are a few microseconds long
  const int Lext = L + N_AQ + 1;
  const int S = NP + Lext;
  const int M = B * S;
  const int NL = e->cfg.llm_layers;
  const StatsMode sm = stats_mode(e, B);

  const bool seg = e->seg_on && e->seg_ev[0];  // graph capture is off while segment timing is on
  if (seg) cudaEventRecord(e->seg_ev[0], s);
  if (pix_u8 && e->crop_scale > 0.f) {  // center_crop_image (OU:616-648) on the device, in front of the patch gather
    CK(vla::center_crop_u8_launch(pix_u8, e->crop_buf, static_cast<long long>(B) * e->cfg.n_images, 224, 224, 224,
                                  e->crop_scale, s, &_err));
    pix_u8 = e->crop_buf;
  }
  // Small batches leave most SMs idle inside every kernel: independent work goes to a side stream (the fork / join
  // are events, so inside a captured CUDA graph they become parallel branches).
  const bool small = e->small_B > 0 && B <= e->small_B;
  cudaStream_t s2 = small ? e->side : s;

  // ---------------- vision towers + projector (MP:196-237, 261-273)
  const bool pro = e->cfg.variant == VLA_HEAD_PRO;
  const int NK = T + N_AQ + 1 + NP;  // policy keys per sample: self | h_a ++ p | h_t
  const long long kv_bs = static_cast<long long>(NK) * PKV;
  // Small batches with a chunk of at most 16 rows: the 24 policy blocks are ONE cluster kernel (policy_fused.cu).
  const bool fused = small && e->policy_fused && !e->pol_blocks.empty() && e->pol_blocks.size() <= vla::POLICY_FUSED_MAX_BLOCKS && T <= 16;
  if (small) {
    if (cudaEventRecord(e->ev_fork, s) != cudaSuccess || cudaStreamWaitEvent(s2, e->ev_fork, 0) != cudaSuccess)
      return e->fail(VLA_ERR_CUDA, "side stream fork failed");
  }
  const int NB = static_cast<int>(e->head.size());
  const bool grouped = fused && e->pol && e->policy_group < NB;
  int pol_next = 0;  // first policy block not launched yet (grouped mode)
  // one group of the fused policy kernel: blocks [first, last], after the K|V rows of block `last` (and, by stream
  // order on s2, of every earlier block) are projected
  auto launch_policy_group = [&](int first, int last, cudaStream_t ps, int prefetch_clusters) -> int {
    vla::PolicyFusedArgs pa;
    for (int i = first; i <= last; ++i) pa.blocks[i - first] = e->pol_blocks[i];
    pa.n_blocks = last - first + 1; pa.x0 = e->head_x[first]; pa.ao = e->h_ao; pa.y = e->h_y;
    pa.T = T; pa.NK = NK; pa.pro = pro ? 1 : 0; pa.rope_cos = e->prope_cos; pa.rope_sin = e->prope_sin;
    pa.scale_log2 = (1.0f / sqrtf(112.0f)) * 1.4426950408889634f; pa.ln_eps = HEAD_EPS;
    pa.B = B; pa.progress = e->pol_progress; pa.prof = first == 0 ? e->pol_prof : nullptr;
    const char* perr = nullptr;
    const int prc = vla::policy_fused_launch(pa, B, ps, &perr, prefetch_clusters);
    if (prc) return e->fail(prc, perr ? perr : "policy_fused launch failed");
    return 0;
  };
  if (grouped) {
    if (cudaStreamWaitEvent(e->pol, e->ev_fork, 0) != cudaSuccess) return e->fail(VLA_ERR_CUDA, "policy stream fork failed");
    CK(vla::broadcast_row_launch(e->x0, D_LLM, B * T, e->head_x[0], e->pol, &_err));
  }
  if (fused) {
    // everything of the policy that depends on the proprio vector only runs first, on the side stream: the projector
    // (PJ:19-24), the proprio row's K|V for all 24 blocks, and its copy into every block's key/value buffer (row T+64)
    CK(vla::skinny_linear_launch(proprio, 1, P, B, P, e->pp_w1, P, D_LLM, e->pp_b1, 1, e->h_p1, D_LLM, nullptr, s2, &_err));
    CK(vla::skinny_linear_launch(e->h_p1, 0, D_LLM, B, D_LLM, e->pp_w2, D_LLM, D_LLM, e->pp_b2, 0, e->h_p, D_LLM, nullptr, s2, &_err));
    vla::GemmArgs g;
    g.A = e->h_p; g.lda = D_LLM; g.rows = B; g.W = e->wkv_cond_all; g.ldw = D_LLM; g.N = 24 * PKV; g.K = D_LLM;
    g.C = e->h_pkv; g.ldc = 24 * PKV; g.bias = e->bkv_cond_all;
    CK(vla::gemm_launch(g, s2, &_err));
    for (int i = 0; i < 24; ++i)
      CK(vla::copy_view_launch(e->h_pkv + static_cast<long long>(i) * PKV, 24LL * PKV, PKV,
                               e->h_kv_blk[i] + static_cast<long long>(T + N_AQ) * PKV, kv_bs, PKV, 1, B, PKV, s2, &_err));
  }
  int rc = run_tower(e, e->dino, pix, pix_u8, B, 0, s2);  // the shorter tower rides on the side stream
  if (rc) return rc;
  rc = run_tower(e, e->sig, pix, pix_u8, B, 1, s);
  if (rc) return rc;
  if (small) {
    if (cudaEventRecord(e->ev_join, s2) != cudaSuccess || cudaStreamWaitEvent(s, e->ev_join, 0) != cudaSuccess)
      return e->fail(VLA_ERR_CUDA, "side stream join failed");
  }
  {
    vla::GemmArgs g;
    g.A = e->w_patches; g.lda = D_VIS; g.rows = B * NP; g.W = e->pj_w1; g.ldw = D_VIS; g.N = D_PROJ1; g.K = D_VIS;
    g.C = e->w_ph1; g.ldc = D_PROJ1; g.bias = e->pj_b1; g.act = vla::ACT_GELU;
    CK(vla::gemm_launch(g, s, &_err));
    g = vla::GemmArgs();
    g.A = e->w_ph1; g.lda = D_PROJ1; g.rows = B * NP; g.W = e->pj_w2; g.ldw = D_PROJ1; g.N = D_LLM; g.K = D_PROJ1;
    g.C = e->w_ph2; g.ldc = D_LLM; g.bias = e->pj_b2; g.act = vla::ACT_GELU;
    CK(vla::gemm_launch(g, s, &_err));
    // fc3 writes straight into rows 1..NP of every sample's LLM input (MP:500-502)
    g = vla::GemmArgs();
    g.A = e->w_ph2; g.a_batch_stride = static_cast<long long>(NP) * D_LLM; g.lda = D_LLM; g.rows = NP; g.batches = B;
    g.W = e->pj_w3; g.ldw = D_LLM; g.N = D_LLM; g.K = D_LLM;
    g.C = e->hid[0] + D_LLM; g.c_batch_stride = static_cast<long long>(S) * D_LLM; g.ldc = D_LLM; g.bias = e->pj_b3;
    CK(vla::gemm_launch(g, s, &_err));
  }

  if (seg) cudaEventRecord(e->seg_ev[1], s);
  // ---------------- LLM input assembly + prefill (MP:418-454, 500-502, 834-845)
  CK(vla::assemble_launch(e->hid[0], B, S, NP, Lext, D_LLM, ext_ids, aq_index, e->embed, e->cfg.vocab_size,
                          e->aq_table, N_AQ, e->err_flag, s, &_err));
  for (int l = 0; l < NL; ++l) {
    const LlmLayer& w = e->llm[l];
    const bf16* xin = e->hid[l];
    bf16* xout = (l == NL - 1) ? e->l_tmp : e->hid[l + 1];
    vla::GemmArgs g;
    if (e->fold_norms && sm != STATS_KERNEL) {
      // sum x^2 of every row comes from the previous layer's down-projection epilogue (layer 0: one statistics pass)
      if (l == 0) CK(vla::row_stats_launch(xin, M, D_LLM, D_LLM, 1, LLM_EPS, e->l_stats, s, &_err, vla::STAT_SLOTS));
      g.A = xin; g.stat_in = e->l_stats; g.stat_dim = D_LLM; g.stat_eps = LLM_EPS; g.stat_rms = 1;
    } else if (e->fold_norms) {
      CK(vla::row_stats_launch(xin, M, D_LLM, D_LLM, 1, LLM_EPS, e->l_stats, s, &_err));
      g.A = xin; g.row_stats = e->l_stats;
    } else {
      CK(vla::rmsnorm_launch(xin, M, D_LLM, D_LLM, w.ln1, LLM_EPS, e->l_xn, D_LLM, s, &_err));
      g.A = e->l_xn;
    }
    g.lda = D_LLM; g.rows = M; g.W = w.wqkv; g.ldw = D_LLM; g.N = QKV_LLM; g.K = D_LLM;
    g.C = e->l_qkv; g.ldc = QKV_LLM; g.bias = w.bqkv;
    // Small batch: RoPE of the q and k heads rides in this GEMM's epilogue (one kernel less on the critical path); the
    // row's cos/sin pairs come from the transposed packed table, one coalesced load per frequency and tile.  At bs=64
    // the fused form and GEMM + stand-alone kernel measure the same (prefill 28.5 ms either way), so the large-batch
    // path keeps the two kernels; VLA_ROPE_FUSE=1 fuses at every batch size.
    static const bool rope_fuse_all = getenv("VLA_ROPE_FUSE") != nullptr;
    const bool fuse_rope = small || rope_fuse_all;
    if (fuse_rope) {
      g.rope_cs = e->rope_cs; g.rope_ld = e->maxS; g.rope_cols = (HQ + HKV) * 64; g.rope_S = S;
    }
    CK(vla::gemm_launch(g, s, &_err));
    if (!fuse_rope) CK(vla::rope_apply_launch(e->l_qkv, QKV_LLM, 0, HQ + HKV, B, S, e->rope_cos, e->rope_sin, s, &_err));
    CK(vla::attention_launch(e->l_qkv, QKV_LLM, 0, HQ * 64, (HQ + HKV) * 64, B, S, HQ, HQ / HKV, 64, e->cfg.causal,
                             e->l_attn, D_LLM, s, &_err));
    g = vla::GemmArgs();
    g.A = e->l_attn; g.lda = D_LLM; g.rows = M; g.W = w.wo; g.ldw = D_LLM; g.N = D_LLM; g.K = D_LLM;
    g.C = xout; g.ldc = D_LLM; g.resid = xin; g.ldr = D_LLM;
    if (sm != STATS_KERNEL) g.stat_out = e->l_stats;
    CK(vla::gemm_launch(g, s, &_err));
    g = vla::GemmArgs();
    if (e->fold_norms && sm != STATS_KERNEL) {
      g.A = xout; g.stat_in = e->l_stats; g.stat_dim = D_LLM; g.stat_eps = LLM_EPS; g.stat_rms = 1;
    } else if (e->fold_norms) {
      CK(vla::row_stats_launch(xout, M, D_LLM, D_LLM, 1, LLM_EPS, e->l_stats, s, &_err));
      g.A = xout; g.row_stats = e->l_stats;
    } else {
      CK(vla::rmsnorm_launch(xout, M, D_LLM, D_LLM, w.ln2, LLM_EPS, e->l_xn, D_LLM, s, &_err));
      g.A = e->l_xn;
    }
    g.lda = D_LLM; g.rows = M; g.W = w.wgu; g.ldw = D_LLM; g.N = 2 * I_LLM; g.K = D_LLM;
    g.C = e->l_act; g.ldc = I_LLM; g.act = vla::ACT_SWIGLU;
    CK(vla::gemm_launch(g, s, &_err));
    g = vla::GemmArgs();
    g.A = e->l_act; g.lda = I_LLM; g.rows = M; g.W = w.wdown; g.ldw = I_LLM; g.N = D_LLM; g.K = I_LLM;
    g.C = xout; g.ldc = D_LLM; g.resid = xout; g.ldr = D_LLM;
    if (sm != STATS_KERNEL && l + 1 < NL) g.stat_out = e->l_stats;
    CK(vla::gemm_launch(g, s, &_err));
    if (small && l + 1 < NL) {  // hid[l+1] is final: its policy K|V projections start now, beside the next LLM layer
      rc = policy_kv_from_llm(e, l, e->hid[l + 1], B, L, prompt_len, s, s2, fused && pro);
      if (rc) return rc;
      if (grouped && l + 1 - pol_next >= e->policy_group) {
        vla::PdlScope no_pdl(false);  // a group that started early would hold its SMs idle beside the prefill
        if (cudaStreamWaitEvent(e->pol, e->ev_kv[l], 0) != cudaSuccess) return e->fail(VLA_ERR_CUDA, "policy K|V join failed");
        rc = launch_policy_group(pol_next, l, e->pol, 2);
        if (rc) return rc;
        pol_next = l + 1;
      }
    }
  }
  // hidden_states[-1] is the post-final-norm state (HF output_hidden_states semantics)
  CK(vla::rmsnorm_launch(e->l_tmp, M, D_LLM, D_LLM, e->llm_norm, LLM_EPS, e->hid[NL], D_LLM, s, &_err));
  if (small) {
    rc = policy_kv_from_llm(e, NL - 1, e->hid[NL], B, L, prompt_len, s, s2, fused && pro);
    if (rc) return rc;
  }

  if (seg) cudaEventRecord(e->seg_ev[2], s);
  // ---------------- Bridge-Attention policy (AH:43-81, 111-121, 218-283 / 337-410)
  const int ha_row0 = NP + L - 1;  // MP:855 with NUM_PROMPT_TOKENS = L-1 (MP:927)
  if (grouped) {
    // the last group: nothing else is running any more, so it gets the full set of L2 prefetchers
    vla::PdlScope no_pdl(false);
    if (cudaStreamWaitEvent(e->pol, e->ev_kv[NB - 1], 0) != cudaSuccess) return e->fail(VLA_ERR_CUDA, "policy K|V join failed");
    rc = launch_policy_group(pol_next, NB - 1, e->pol, -1);
    if (rc) return rc;
    if (cudaEventRecord(e->ev_pol, e->pol) != cudaSuccess || cudaStreamWaitEvent(s, e->ev_pol, 0) != cudaSuccess)
      return e->fail(VLA_ERR_CUDA, "policy stream join failed");
  } else if (fused) {
    // every block's cond / vision K|V rows are on their way on the side stream; the last event orders them all
    if (cudaStreamWaitEvent(s, e->ev_kv[NB - 1], 0) != cudaSuccess) return e->fail(VLA_ERR_CUDA, "policy K|V join failed");
    CK(vla::broadcast_row_launch(e->x0, D_LLM, B * T, e->head_x[0], s, &_err));
    rc = launch_policy_group(0, NB - 1, s, -1);
    if (rc) return rc;
  } else {
  CK(vla::skinny_linear_launch(proprio, 1, P, B, P, e->pp_w1, P, D_LLM, e->pp_b1, 1, e->h_p1, D_LLM, nullptr, s, &_err));
  CK(vla::skinny_linear_launch(e->h_p1, 0, D_LLM, B, D_LLM, e->pp_w2, D_LLM, D_LLM, e->pp_b2, 0, e->h_p, D_LLM, nullptr, s, &_err));
  {
    vla::GemmArgs g;  // K|V of the proprio row for all 24 blocks at once
    g.A = e->h_p; g.lda = D_LLM; g.rows = B; g.W = e->wkv_cond_all; g.ldw = D_LLM; g.N = 24 * PKV; g.K = D_LLM;
    g.C = e->h_pkv; g.ldc = 24 * PKV; g.bias = e->bkv_cond_all;
    CK(vla::gemm_launch(g, s, &_err));
  }
  CK(vla::broadcast_row_launch(e->x0, D_LLM, B * T, e->head_x[0], s, &_err));
  for (int i = 0; i < NB; ++i) {
    const HeadBlock& w = e->head[i];
    const bf16* hs = e->hid[i + 1];  // AH:118: block i reads hidden state i+1
    bf16* kvb = small ? e->h_kv_blk[i] : e->h_kv;
    vla::GemmArgs g;
    if (small) {
      // the LLM-state K|V rows of this block were projected on the side stream (policy_kv_from_llm)
      if (cudaStreamWaitEvent(s, e->ev_kv[i], 0) != cudaSuccess) return e->fail(VLA_ERR_CUDA, "policy K|V join failed");
    } else {
      rc = policy_kv_gemms(e, i, hs, B, L, prompt_len, kvb, s);
      if (rc) return rc;
    }
    // K|V of the proprio row -> kv row T+64 (projected up front for all blocks: one row copy per sample here)
    CK(vla::copy_view_launch(e->h_pkv + static_cast<long long>(i) * PKV, 24LL * PKV, PKV,
                             kvb + static_cast<long long>(T + N_AQ) * PKV, kv_bs, PKV, 1, B, PKV, s, &_err));
    // q and self K|V of the current x -> kv rows [0, T)
    g = vla::GemmArgs();
    g.A = e->head_x[i]; g.lda = D_LLM; g.rows = B * T; g.W = w.wq; g.ldw = D_LLM; g.N = D_LLM; g.K = D_LLM;
    g.C = e->h_q; g.ldc = D_LLM; g.bias = w.bq;
    CK(vla::gemm_launch(g, s, &_err));
    g = vla::GemmArgs();
    g.A = e->head_x[i]; g.a_batch_stride = static_cast<long long>(T) * D_LLM; g.lda = D_LLM; g.rows = T; g.batches = B;
    g.W = w.wkv_self; g.ldw = D_LLM; g.N = PKV; g.K = D_LLM;
    g.C = kvb; g.c_batch_stride = kv_bs; g.ldc = PKV; g.bias = w.bkv_self;
    CK(vla::gemm_launch(g, s, &_err));
    if (pro) CK(vla::policy_rope_launch(e->h_q, kvb, B, T, NP, e->prope_cos, e->prope_sin, s, &_err));
    CK(vla::cross_attention_launch(e->h_q, D_LLM, T, kvb, kvb + D_LLM, PKV, NK, B, 8, 1, 112, 0, e->h_ao,
                                   D_LLM, s, &_err));
    g = vla::GemmArgs();
    g.A = e->h_ao; g.lda = D_LLM; g.rows = B * T; g.W = w.wo; g.ldw = D_LLM; g.N = D_LLM; g.K = D_LLM;
    g.C = e->h_y; g.ldc = D_LLM; g.bias = w.bo; g.resid = e->head_x[i]; g.ldr = D_LLM;
    CK(vla::gemm_launch(g, s, &_err));
    CK(vla::layernorm_launch(e->h_y, B * T, D_LLM, D_LLM, w.ffn_lnw, w.ffn_lnb, HEAD_EPS, e->h_yn, D_LLM, s, &_err));
    g = vla::GemmArgs();
    g.A = e->h_yn; g.lda = D_LLM; g.rows = B * T; g.W = w.wffn; g.ldw = D_LLM; g.N = D_LLM; g.K = D_LLM;
    g.C = e->head_x[i + 1]; g.ldc = D_LLM; g.bias = w.bffn; g.act = vla::ACT_RELU;
    CK(vla::gemm_launch(g, s, &_err));
  }
  }  // !fused
  CK(vla::head_out_launch(e->head_x[NB], B * T, e->head_ln2w, e->head_ln2b, e->head_fc2_w, e->head_fc2_b, A,
                          e->st_hi, e->st_lo, e->st_mask, out_norm, out_unnorm, s, &_err));
  if (seg) cudaEventRecord(e->seg_ev[3], s);
  if (out_last_ha)
    CK(vla::gather_rows_launch(e->hid[NL], static_cast<long long>(S) * D_LLM, D_LLM, ha_row0, N_AQ, B, D_LLM,
                               out_last_ha, s, &_err, prompt_len, L, e->err_flag));
  return 0;
}

}  // namespace

extern "C" {

int vla_create(const vla_cfg* cfg, vla_engine** out) {
  if (!cfg || !out) return VLA_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return VLA_ERR_CUDA;  // no CPU fallback
  }
  vla_engine* e = new vla_engine();
  e->cfg = *cfg;
  *out = e;
  if (cudaGetDevice(&e->device) != cudaSuccess) return e->fail(VLA_ERR_CUDA, "cudaGetDevice failed");
  {
    const char* werr = nullptr;
    if (vla::watchdog_install_current_device(&werr)) return e->fail(VLA_ERR_CUDA, werr ? werr : "watchdog install failed");
  }
  if (const char* ng = getenv("VLA_NO_GRAPH")) e->use_graphs = atoi(ng) ? 0 : 1;
  if (const char* nf = getenv("VLA_NO_NORM_FOLD")) e->fold_norms = atoi(nf) ? 0 : 1;
  if (const char* sf = getenv("VLA_STAT_FUSE")) e->stat_fuse = atoi(sf) ? 1 : 0;
  if (const char* pf = getenv("VLA_NO_POLICY_FUSED")) e->policy_fused = atoi(pf) ? 0 : 1;
  if (const char* pg = getenv("VLA_POLICY_GROUP")) {
    const int v = atoi(pg);
    if (v >= 1) e->policy_group = v;
  }
  if (!e->fold_norms) e->stat_fuse = 0;
  const vla_cfg& c = e->cfg;
  if (c.n_images < 1 || c.n_images > 3) return e->fail(VLA_ERR_INVALID, "n_images must be 1..3");
  if (c.chunk_len < 1 || c.chunk_len > 32) return e->fail(VLA_ERR_INVALID, "chunk_len must be 1..32");
  if (c.action_dim < 1 || c.action_dim > 64) return e->fail(VLA_ERR_INVALID, "action_dim must be 1..64");
  if (c.proprio_dim < 1 || c.proprio_dim > 1024) return e->fail(VLA_ERR_INVALID, "proprio_dim must be 1..1024");
  if (c.variant != VLA_HEAD_BASE && c.variant != VLA_HEAD_PRO) return e->fail(VLA_ERR_INVALID, "unknown head variant");
  if (c.dino_depth < 2 || c.siglip_depth < 2) return e->fail(VLA_ERR_INVALID, "ViT depth must be >= 2");
  if (c.llm_layers != 24)
    return e->fail(VLA_ERR_INVALID, "llm_layers must be 24: MLPResNet has 24 blocks, block i reads state i+1 (AH:36,118)");
  if (c.vocab_size < 3) return e->fail(VLA_ERR_INVALID, "vocab_size must cover the placeholder/stop ids");
  if (c.max_batch < 1 || c.max_prompt_len < 1) return e->fail(VLA_ERR_INVALID, "max_batch/max_prompt_len must be >= 1");
  e->NP = 256 * c.n_images;
  e->T = c.chunk_len;
  e->A = c.action_dim;
  e->P = c.proprio_dim;
  return VLA_OK;
}

int vla_load_tensor(vla_engine* e, const char* name, const void* ptr, int dtype, int ndim, const int64_t* shape) {
  if (!e) return VLA_ERR_INVALID;
  if (!name || !ptr || ndim < 0 || ndim > 8 || (ndim && !shape)) return e->fail(VLA_ERR_INVALID, "load_tensor: bad argument");
  if (e->finalized) return e->fail(VLA_ERR_INVALID, "load_tensor after finalize");
  std::string n(name);
  if (n.rfind("vla.", 0) != 0 && n.rfind("head.", 0) != 0 && n.rfind("proprio.", 0) != 0)
    return e->fail(VLA_ERR_INVALID, "unknown tensor name (expected vla./head./proprio. prefix): " + n);
  if (dtype != VLA_BF16 && dtype != VLA_F32 && dtype != VLA_F16) return e->fail(VLA_ERR_DTYPE, "unsupported dtype for " + n);
  if (ignored_name(n)) return VLA_OK;
  size_t numel = 1;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] < 0) return e->fail(VLA_ERR_INVALID, "negative dimension in " + n);
    numel *= static_cast<size_t>(shape[i]);
  }
  if (numel == 0) return e->fail(VLA_ERR_INVALID, "empty tensor " + n);
  const size_t esz = dtype == VLA_F32 ? 4 : 2;
  vla::DeviceGuard guard(e->device);
  // The tensor is kept in ITS dtype (one copy, no kernel): vla_finalize converts and repacks everything in one launch.
  Master m;
  m.n = numel;
  m.dtype = dtype;
  m.shape.assign(shape, shape + ndim);
  auto old = e->masters.find(n);
  if (old != e->masters.end() && old->second.n * (old->second.dtype == VLA_F32 ? 4 : 2) == numel * esz) {
    m.d = old->second.d;
  } else {
    if (old != e->masters.end()) {
      cudaFree(old->second.d);
      e->masters.erase(old);
    }
    if (cudaMalloc(&m.d, numel * esz) != cudaSuccess) {
      cudaGetLastError();
      return e->fail(VLA_ERR_CUDA, "cudaMalloc failed for " + n);
    }
  }
  // cudaMemcpyDefault: `ptr` may be host or device memory.  The stream synchronize makes the copy complete before
  // the call returns (a device-to-device cudaMemcpy is asynchronous): the caller may free `ptr` right away.
  cudaError_t ce = cudaMemcpyAsync(m.d, ptr, numel * esz, cudaMemcpyDefault, 0);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(0);
  e->masters[n] = m;
  if (ce != cudaSuccess) return e->fail(VLA_ERR_CUDA, "copy failed for " + n + ": " + cudaGetErrorString(ce));
  return VLA_OK;
}

int vla_set_action_stats(vla_engine* e, const double* hi, const double* lo, const uint8_t* mask) {
  if (!e) return VLA_ERR_INVALID;
  if (!hi || !lo) return e->fail(VLA_ERR_INVALID, "set_action_stats: null statistics");
  vla::DeviceGuard guard(e->device);
  const int A = e->A;
  std::vector<float> fh(A), fl(A);
  std::vector<uint8_t> fm(A, 1);
  for (int i = 0; i < A; ++i) {
    fh[i] = static_cast<float>(hi[i]);
    fl[i] = static_cast<float>(lo[i]);
    if (mask) fm[i] = mask[i] ? 1 : 0;
  }
  try {
    if (!e->st_hi) {
      e->st_hi = e->dalloc<float>(A);
      e->st_lo = e->dalloc<float>(A);
      e->st_mask = e->dalloc<uint8_t>(A);
    }
  } catch (const std::exception& ex) {
    return e->fail(VLA_ERR_CUDA, ex.what());
  }
  cudaMemcpy(e->st_hi, fh.data(), A * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(e->st_lo, fl.data(), A * sizeof(float), cudaMemcpyHostToDevice);
  cudaMemcpy(e->st_mask, fm.data(), A, cudaMemcpyHostToDevice);
  e->stats_set = true;
  return VLA_OK;
}

int vla_finalize(vla_engine* e) {
  if (!e) return VLA_ERR_INVALID;
  if (e->finalized) return VLA_OK;
  vla::DeviceGuard guard(e->device);
  const vla_cfg& c = e->cfg;
  size_t mark = e->allocs.size();
  // A failed attempt (typically VLA_ERR_MISSING) leaves the engine as it was before the call: the caller may load the
  // missing tensors and call again.
  auto rollback = [&]() {
    for (size_t i = mark; i < e->allocs.size(); ++i) cudaFree(e->allocs[i]);
    e->allocs.resize(mark);
    e->jobs.clear();
    e->dino.blocks.clear(); e->sig.blocks.clear(); e->llm.clear(); e->head.clear();
    e->hid.clear(); e->head_x.clear(); e->h_kv_blk.clear(); e->h_ha.clear();
    e->dev_pix = nullptr;
  };
  try {
    if (!e->stats_set) {
      std::vector<double> hi(e->A, 1.0), lo(e->A, -1.0);
      int rc = vla_set_action_stats(e, hi.data(), lo.data(), nullptr);
      if (rc) return rc;
    }
    if (!e->img_lut) {  // preprocessor_config.json: DINOv2 = ImageNet statistics, SigLIP = 0.5 / 0.5
      const float mean[6] = {0.485f, 0.456f, 0.406f, 0.5f, 0.5f, 0.5f}, stdv[6] = {0.229f, 0.224f, 0.225f, 0.5f, 0.5f, 0.5f};
      int rc = vla_set_image_norm(e, mean, stdv);
      if (rc) return rc;
    }
    mark = e->allocs.size();  // the statistics / image table above survive a failed attempt
    build_tower(e, e->dino, "vla.vision_backbone.featurizer.", true, c.dino_depth);
    build_tower(e, e->sig, "vla.vision_backbone.fused_featurizer.", false, c.siglip_depth);
    e->pj_w1 = e->pack("vla.projector.fc1.weight", D_PROJ1, D_VIS);
    e->pj_b1 = e->f32("vla.projector.fc1.bias", D_PROJ1);
    e->pj_w2 = e->pack("vla.projector.fc2.weight", D_LLM, D_PROJ1);
    e->pj_b2 = e->f32("vla.projector.fc2.bias", D_LLM);
    e->pj_w3 = e->pack("vla.projector.fc3.weight", D_LLM, D_LLM);
    e->pj_b3 = e->f32("vla.projector.fc3.bias", D_LLM);

    const std::string lm = "vla.language_model.model.";
    e->embed = e->pack(lm + "embed_tokens.weight", c.vocab_size, D_LLM);
    e->aq_table = e->pack("vla.action_queries.weight", N_AQ, D_LLM);
    for (int l = 0; l < c.llm_layers; ++l) {
      const std::string b = lm + "layers." + std::to_string(l) + ".";
      LlmLayer w;
      w.ln1 = e->f32(b + "input_layernorm.weight", D_LLM);
      w.wqkv = e->pack_cat({{b + "self_attn.q_proj.weight", HQ * 64}, {b + "self_attn.k_proj.weight", HKV * 64}, {b + "self_attn.v_proj.weight", HKV * 64}}, D_LLM);
      w.bqkv = e->cat_f32({{b + "self_attn.q_proj.bias", HQ * 64}, {b + "self_attn.k_proj.bias", HKV * 64}, {b + "self_attn.v_proj.bias", HKV * 64}});
      w.wo = e->pack(b + "self_attn.o_proj.weight", D_LLM, D_LLM);
      w.ln2 = e->f32(b + "post_attention_layernorm.weight", D_LLM);
      // gate/up rows interleaved in groups of 16 for the SwiGLU epilogue
      w.wgu = e->dalloc<bf16>(static_cast<size_t>(2 * I_LLM) * D_LLM);
      e->pack_into(w.wgu, D_LLM, b + "mlp.gate_proj.weight", I_LLM, D_LLM, 16, 32, 0);
      e->pack_into(w.wgu, D_LLM, b + "mlp.up_proj.weight", I_LLM, D_LLM, 16, 32, 16);
      w.wdown = e->pack(b + "mlp.down_proj.weight", D_LLM, I_LLM);
      e->llm.push_back(w);
    }
    e->llm_norm = e->f32(lm + "norm.weight", D_LLM);

    // ---- policy head
    const bool pro = c.variant == VLA_HEAD_PRO;
    const std::string hm = "head.model.";
    const int in_dim = e->A * D_LLM;  // MLPResNet input_dim = input_dim * ACTION_DIM (AH:37)
    float* gates_dev = e->dalloc<float>(24);
    e->wkv_cond_all = e->dalloc<bf16>(static_cast<size_t>(24) * PKV * D_LLM);
    e->bkv_cond_all = e->dalloc<float>(static_cast<size_t>(24) * PKV);
    for (int i = 0; i < 24; ++i) {
      const std::string b = hm + "mlp_resnet_blocks." + std::to_string(i) + ".";
      HeadBlock w;
      const std::string ks = pro ? "k_self" : "k_proj", vs = pro ? "v_self" : "v_proj";
      const std::string kc = pro ? "k_adapter" : "k_proj", vc = pro ? "v_adapter" : "v_proj";
      const std::string kv = pro ? "k_task" : "k_proj", vv = pro ? "v_task" : "v_proj";
      w.wq = e->pack(b + "q_proj.weight", D_LLM, D_LLM);
      w.bq = e->f32(b + "q_proj.bias", D_LLM);
      w.wkv_cond = e->pack_cat({{b + kc + ".weight", D_LLM}, {b + vc + ".weight", D_LLM}}, D_LLM);
      w.bkv_cond = e->cat_f32({{b + kc + ".bias", D_LLM}, {b + vc + ".bias", D_LLM}});
      if (pro) {
        w.wkv_self = e->pack_cat({{b + ks + ".weight", D_LLM}, {b + vs + ".weight", D_LLM}}, D_LLM);
        w.bkv_self = e->cat_f32({{b + ks + ".bias", D_LLM}, {b + vs + ".bias", D_LLM}});
        w.wkv_vis = e->pack_cat({{b + kv + ".weight", D_LLM}, {b + vv + ".weight", D_LLM}}, D_LLM);
        w.bkv_vis = e->cat_f32({{b + kv + ".bias", D_LLM}, {b + vv + ".bias", D_LLM}});
      } else {  // base: one shared k_proj / v_proj for all three segments (AH:247-254)
        w.wkv_self = w.wkv_vis = w.wkv_cond;
        w.bkv_self = w.bkv_vis = w.bkv_cond;
      }
      w.wo = e->pack(b + "o_proj.weight", D_LLM, D_LLM);
      w.bo = e->f32(b + "o_proj.bias", D_LLM);
      w.ffn_lnw = e->f32(b + "ffn.0.weight", D_LLM);
      w.ffn_lnb = e->f32(b + "ffn.0.bias", D_LLM);
      w.wffn = e->pack(b + "ffn.1.weight", D_LLM, D_LLM);
      w.bffn = e->f32(b + "ffn.1.bias", D_LLM);
      e->enqueue(e->need(b + "gating_factor", 1), gates_dev + i, true, 1, 1, 1, 1, 0, 0);
      // the proprio row is the same for all 24 blocks: its K|V projections become ONE GEMM against the stacked weights
      e->pack_into(e->wkv_cond_all, D_LLM, b + kc + ".weight", D_LLM, D_LLM, D_LLM, 0, i * PKV);
      e->pack_into(e->wkv_cond_all, D_LLM, b + vc + ".weight", D_LLM, D_LLM, D_LLM, 0, i * PKV + D_LLM);
      e->enqueue(e->need(b + kc + ".bias", D_LLM), e->bkv_cond_all + static_cast<size_t>(i) * PKV, true, 1, D_LLM, D_LLM, 1, 0, 0);
      e->enqueue(e->need(b + vc + ".bias", D_LLM), e->bkv_cond_all + static_cast<size_t>(i) * PKV + D_LLM, true, 1, D_LLM, D_LLM, 1, 0, 0);
      w.gatevec = e->dalloc<float>(PKV);
      e->head.push_back(w);
    }
    e->head_ln2w = e->f32(hm + "layer_norm2.weight", D_LLM);
    e->head_ln2b = e->f32(hm + "layer_norm2.bias", D_LLM);
    e->head_fc2_w = e->pack(hm + "fc2.weight", e->A, D_LLM);
    e->head_fc2_b = e->f32(hm + "fc2.bias", e->A);
    e->pp_w1 = e->pack("proprio.fc1.weight", D_LLM, e->P);
    e->pp_b1 = e->f32("proprio.fc1.bias", D_LLM);
    e->pp_w2 = e->pack("proprio.fc2.weight", D_LLM, D_LLM);
    e->pp_b2 = e->f32("proprio.fc2.bias", D_LLM);
    bf16* wfc1 = e->pack(hm + "fc1.weight", D_LLM, in_dim);
    e->need(hm + "layer_norm1.weight", in_dim);
    float* ln1_bias = e->f32(hm + "layer_norm1.bias", in_dim);
    float* fc1_bias = e->f32(hm + "fc1.bias", D_LLM);

    // ---- every conversion / repack queued above runs now, in one kernel launch
    e->flush_jobs();

    {  // ratio_g = tanh(g) evaluated in the parameter dtype (bf16), AH:225 / AH:344
      float gates[24];
      if (cudaMemcpy(gates, gates_dev, sizeof(gates), cudaMemcpyDeviceToHost) != cudaSuccess)
        throw std::runtime_error("reading the gating factors failed");
      std::vector<float> gv(PKV, 1.0f);
      for (int i = 0; i < 24; ++i) {
        HeadBlock& w = e->head[i];
        const float gb = __bfloat162float(__float2bfloat16_rn(gates[i]));
        w.gate = __bfloat162float(__float2bfloat16_rn(std::tanh(gb)));
        for (int c2 = 0; c2 < D_LLM; ++c2) gv[c2] = w.gate;
        cudaMemcpy(w.gatevec, gv.data(), PKV * sizeof(float), cudaMemcpyHostToDevice);
      }
    }
    // x0 = ReLU(fc1(LayerNorm(0))) = ReLU(fc1.W @ bf16(ln1.bias) + fc1.b): input independent (AH:60-67, 114-116)
    {
      e->x0 = e->dalloc<bf16>(D_LLM);
      const char* err = nullptr;
      int rc = vla::skinny_linear_launch(ln1_bias, 1, in_dim, 1, in_dim, wfc1, in_dim, D_LLM, fc1_bias, 2, e->x0, D_LLM,
                                         nullptr, 0, &err);
      if (rc) {
        rollback();
        return e->fail(rc, err ? err : "x0 precompute failed");
      }
    }

    // ---- norms in front of GEMMs folded into the GEMM weights (once; see fold_norms)
    if (e->fold_norms) {
      const char* err = nullptr;
      int rc = 0;
      for (Tower* t : {&e->dino, &e->sig}) {
        for (VitBlock& k : t->blocks) {
          k.cs_qkv = e->dalloc<float>(3 * t->D);
          k.cs_fc1 = e->dalloc<float>(t->F);
          if (!rc) rc = vla::fold_norm_launch(k.wqkv, 3 * t->D, t->D, t->D, k.ln1w, k.ln1b, k.bqkv, k.cs_qkv, nullptr, &err);
          if (!rc) rc = vla::fold_norm_launch(k.wfc1, t->F, t->D, t->D, k.ln2w, k.ln2b, k.bfc1, k.cs_fc1, nullptr, &err);
        }
      }
      for (LlmLayer& w : e->llm) {
        if (!rc) rc = vla::fold_norm_launch(w.wqkv, QKV_LLM, D_LLM, D_LLM, w.ln1, nullptr, nullptr, nullptr, nullptr, &err);
        if (!rc) rc = vla::fold_norm_launch(w.wgu, 2 * I_LLM, D_LLM, D_LLM, w.ln2, nullptr, nullptr, nullptr, nullptr, &err);
      }
      if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = VLA_ERR_CUDA;
      if (rc) {
        rollback();
        return e->fail(rc, err ? err : "norm fold failed");
      }
    }

    // ---- workspace
    const int B = c.max_batch, L = c.max_prompt_len, n = c.n_images;
    const int S = e->NP + L + N_AQ + 1;
    e->maxB = B; e->maxL = L; e->maxS = S;
    const size_t slabs = static_cast<size_t>(B) * n;
    const size_t Mv = slabs * TOK_DINO;
    for (int t = 0; t < 2; ++t) {
      const size_t D = t == 0 ? D_DINO : D_SIG, F = t == 0 ? F_DINO : F_SIG;
      e->tw[t].col = e->dalloc<bf16>(slabs * 256 * KP);
      e->tw[t].x = e->dalloc<bf16>(Mv * D);
      e->tw[t].xn = e->dalloc<bf16>(Mv * D);
      e->tw[t].stats = e->dalloc<float>(2 * vla::STAT_SLOTS * Mv);
      e->tw[t].qkv = e->dalloc<bf16>(Mv * 3 * D);
      e->tw[t].attn = e->dalloc<bf16>(Mv * D);
      e->tw[t].h = e->dalloc<bf16>(Mv * F);
    }
    e->w_patches = e->dalloc<bf16>(slabs * 256 * D_VIS);
    e->w_ph1 = e->dalloc<bf16>(slabs * 256 * D_PROJ1);
    e->w_ph2 = e->dalloc<bf16>(slabs * 256 * D_LLM);
    const size_t Ml = static_cast<size_t>(B) * S;
    for (int i = 0; i <= c.llm_layers; ++i) e->hid.push_back(e->dalloc<bf16>(Ml * D_LLM));
    e->l_tmp = e->dalloc<bf16>(Ml * D_LLM);
    e->l_xn = e->dalloc<bf16>(Ml * D_LLM);
    e->l_stats = e->dalloc<float>(2 * vla::STAT_SLOTS * Ml);
    e->l_qkv = e->dalloc<bf16>(Ml * QKV_LLM);
    e->l_attn = e->dalloc<bf16>(Ml * D_LLM);
    e->l_act = e->dalloc<bf16>(Ml * I_LLM);
    const size_t BT = static_cast<size_t>(B) * e->T;
    e->h_p1 = e->dalloc<bf16>(static_cast<size_t>(B) * D_LLM);
    e->h_p = e->dalloc<bf16>(static_cast<size_t>(B) * D_LLM);
    e->h_kv = e->dalloc<bf16>(static_cast<size_t>(B) * (e->T + N_AQ + 1 + e->NP) * PKV);
    e->small_B = B < 8 ? B : 8;
    if (getenv("VLA_NO_SIDE_STREAM")) e->small_B = 0;
    if (e->small_B) {
      for (int i = 0; i < 24; ++i)
        e->h_kv_blk.push_back(e->dalloc<bf16>(static_cast<size_t>(e->small_B) * (e->T + N_AQ + 1 + e->NP) * PKV));
      if (!e->side && cudaStreamCreateWithFlags(&e->side, cudaStreamNonBlocking) != cudaSuccess)
        throw std::runtime_error("side stream creation failed");
      if (!e->pol && cudaStreamCreateWithFlags(&e->pol, cudaStreamNonBlocking) != cudaSuccess)
        throw std::runtime_error("policy stream creation failed");
      if (!e->ev_pol) cudaEventCreateWithFlags(&e->ev_pol, cudaEventDisableTiming);
      if (!e->ev_fork) cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming);
      if (!e->ev_join) cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming);
      for (int i = static_cast<int>(e->ev_layer.size()); i < 24; ++i) {
        cudaEvent_t a = nullptr, b = nullptr;
        cudaEventCreateWithFlags(&a, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&b, cudaEventDisableTiming);
        e->ev_layer.push_back(a);
        e->ev_kv.push_back(b);
      }
    }
    for (int i = 0; i < 24; ++i) e->h_ha.push_back(e->dalloc<bf16>(static_cast<size_t>(B) * N_AQ * D_LLM));
    e->h_pkv = e->dalloc<bf16>(static_cast<size_t>(B) * 24 * PKV);
    e->h_q = e->dalloc<bf16>(BT * D_LLM);
    e->h_ao = e->dalloc<bf16>(BT * D_LLM);
    e->h_y = e->dalloc<bf16>(BT * D_LLM);
    e->h_yn = e->dalloc<bf16>(BT * D_LLM);
    for (int i = 0; i <= 24; ++i) e->head_x.push_back(e->dalloc<bf16>(BT * D_LLM));
    if (e->small_B) {  // pointer table of the fused small-batch policy kernel
      std::vector<vla::PolicyBlockW> pb(24);
      for (int i = 0; i < 24; ++i) {
        const HeadBlock& w = e->head[i];
        pb[i].wq = w.wq; pb[i].wkvs = w.wkv_self; pb[i].wo = w.wo; pb[i].wffn = w.wffn;
        pb[i].bq = w.bq; pb[i].bkvs = w.bkv_self; pb[i].bo = w.bo; pb[i].lnw = w.ffn_lnw; pb[i].lnb = w.ffn_lnb;
        pb[i].bffn = w.bffn; pb[i].kv = e->h_kv_blk[i]; pb[i].x_out = e->head_x[i + 1];
      }
      e->pol_blocks = pb;
      e->pol_progress = e->dalloc<int>(1);
      cudaMemset(e->pol_progress, 0, sizeof(int));
      if (getenv("VLA_POLICY_PROF")) {
        e->pol_prof = e->dalloc<long long>(24 * 8);
        cudaMemset(e->pol_prof, 0, 24 * 8 * sizeof(long long));
      }
    }
    e->err_flag = e->dalloc<int>(1);
    cudaMemset(e->err_flag, 0, sizeof(int));
    e->rope_cos = e->dalloc<float>(static_cast<size_t>(S) * 32);
    e->rope_sin = e->dalloc<float>(static_cast<size_t>(S) * 32);
    const char* err = nullptr;
    int rc = vla::rope_table_launch(e->rope_cos, e->rope_sin, S, 32, ROPE_THETA, 0, &err);
    e->rope_cs = e->dalloc<uint32_t>(static_cast<size_t>(S) * 32);
    if (!rc) rc = vla::rope_pack_launch(e->rope_cos, e->rope_sin, S, e->rope_cs, 0, &err);
    if (rc) throw std::runtime_error(err ? err : "rope table failed");
    const int max_pos = e->NP > 65 ? e->NP : 65;
    e->prope_cos = e->dalloc<float>(static_cast<size_t>(max_pos) * 112);
    e->prope_sin = e->dalloc<float>(static_cast<size_t>(max_pos) * 112);
    rc = vla::policy_rope_table_launch(e->prope_cos, e->prope_sin, max_pos, 0, &err);
    if (rc) throw std::runtime_error(err ? err : "policy rope table failed");
    cudaError_t ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) throw std::runtime_error(std::string("finalize: ") + cudaGetErrorString(ce));
  } catch (const std::exception& ex) {
    const std::string m = ex.what();
    rollback();
    return e->fail(m.rfind("missing tensor", 0) == 0 ? VLA_ERR_MISSING : (m.rfind("tensor ", 0) == 0 ? VLA_ERR_INVALID : VLA_ERR_CUDA), m);
  }
  e->free_masters();  // every weight now lives in its repacked form only
  e->finalized = true;
  return VLA_OK;
}

static int predict_impl(vla_engine* e, const void* pixel_values, int is_u8, const int64_t* ext_ids,
                        const int32_t* aq_index, const float* proprio, const int32_t* prompt_len, int B, int L,
                        float* out_norm, float* out_unnorm, void* out_last_ha, void* stream) {
  if (!e) return VLA_ERR_INVALID;
  if (!e->finalized) return e->fail(VLA_ERR_NOT_FINALIZED, "vla_predict before vla_finalize");
  if (!pixel_values || !ext_ids || !aq_index || !proprio || !out_norm)
    return e->fail(VLA_ERR_INVALID, "vla_predict: null input/output pointer");
  if (B < 1 || B > e->maxB) return e->fail(VLA_ERR_INVALID, "vla_predict: batch outside [1, max_batch]");
  if (L < 1 || L > e->maxL) return e->fail(VLA_ERR_INVALID, "vla_predict: prompt length outside [1, max_prompt_len]");
  vla::DeviceGuard guard(e->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bf16* pix = is_u8 ? nullptr : static_cast<const bf16*>(pixel_values);
  const uint8_t* pix8 = is_u8 ? static_cast<const uint8_t*>(pixel_values) : nullptr;
  bf16* ha = static_cast<bf16*>(out_last_ha);
  int rc = 0;
  vla_engine::GraphEntry* ge = nullptr;
  if (e->use_graphs && !vla::gemm_profile_enabled() && !e->seg_on) {
    const vla_engine::GraphKey key{B, L, is_u8, pixel_values, ext_ids, aq_index, proprio, prompt_len, out_norm, out_unnorm, out_last_ha};
    for (auto& g : e->graphs)
      if (g.key == key) ge = &g;
    if (!ge) {
      if (e->graphs.size() >= 16) {  // bounded cache: drop the oldest entry
        if (e->graphs.front().exec) cudaGraphExecDestroy(e->graphs.front().exec);
        e->graphs.erase(e->graphs.begin());
      }
      e->graphs.push_back(vla_engine::GraphEntry());
      ge = &e->graphs.back();
      ge->key = key;
    }
    ++ge->seen;
  }
  if (ge && ge->seen >= 2) {
    if (!e->gstream) {
      if (cudaStreamCreateWithFlags(&e->gstream, cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&e->gev_in, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&e->gev_out, cudaEventDisableTiming) != cudaSuccess)
        return e->fail(VLA_ERR_CUDA, "graph stream/event creation failed");
    }
    if (!ge->exec) {
      cudaGraph_t graph = nullptr;
      const long long before = vla::gemm_launch_count() + vla::ops_launch_count();
      if (cudaStreamBeginCapture(e->gstream, cudaStreamCaptureModeRelaxed) != cudaSuccess)
        return e->fail(VLA_ERR_CUDA, "cudaStreamBeginCapture failed");
      rc = forward(e, pix, pix8, ext_ids, aq_index, proprio, prompt_len, B, L, out_norm, out_unnorm, ha, e->gstream);
      const cudaError_t ce = cudaStreamEndCapture(e->gstream, &graph);
      ge->launches = vla::gemm_launch_count() + vla::ops_launch_count() - before;
      if (rc) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
      }
      if (ce != cudaSuccess || !graph) return e->fail(VLA_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
      const cudaError_t ci = cudaGraphInstantiate(&ge->exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ci != cudaSuccess) return e->fail(VLA_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(ci));
    }
    cudaError_t ce = cudaEventRecord(e->gev_in, s);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(e->gstream, e->gev_in, 0);
    if (ce == cudaSuccess) ce = cudaGraphLaunch(ge->exec, e->gstream);
    if (ce == cudaSuccess) ce = cudaEventRecord(e->gev_out, e->gstream);
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s, e->gev_out, 0);
    if (ce != cudaSuccess) return e->fail(VLA_ERR_CUDA, std::string("graph launch: ") + cudaGetErrorString(ce));
    e->last_launches = ge->launches;
  } else {
    const long long before = vla::gemm_launch_count() + vla::ops_launch_count();
    rc = forward(e, pix, pix8, ext_ids, aq_index, proprio, prompt_len, B, L, out_norm, out_unnorm, ha, s);
    e->last_launches = vla::gemm_launch_count() + vla::ops_launch_count() - before;
  }
  e->lastB = B;
  e->lastL = L;
  return rc;
}

int vla_predict(vla_engine* e, const void* pixel_values, const int64_t* ext_ids, const int32_t* aq_index,
                const int32_t* prompt_len, const float* proprio, int B, int L, float* out_norm, float* out_unnorm,
                void* out_last_ha, void* stream) {
  return predict_impl(e, pixel_values, 0, ext_ids, aq_index, proprio, prompt_len, B, L, out_norm, out_unnorm, out_last_ha,
                      stream);
}

int vla_predict_u8(vla_engine* e, const uint8_t* images, const int64_t* ext_ids, const int32_t* aq_index,
                   const int32_t* prompt_len, const float* proprio, int B, int L, float* out_norm, float* out_unnorm,
                   void* out_last_ha, void* stream) {
  return predict_impl(e, images, 1, ext_ids, aq_index, proprio, prompt_len, B, L, out_norm, out_unnorm, out_last_ha, stream);
}

int vla_set_center_crop(vla_engine* e, float crop_scale) {
  if (!e) return VLA_ERR_INVALID;
  if (crop_scale < 0.f || crop_scale > 1.f) return e->fail(VLA_ERR_INVALID, "set_center_crop: crop_scale outside [0, 1]");
  if (!e->finalized) return e->fail(VLA_ERR_NOT_FINALIZED, "vla_set_center_crop before vla_finalize");
  vla::DeviceGuard guard(e->device);
  if (crop_scale > 0.f && !e->crop_buf) {
    try {
      e->crop_buf = e->dalloc<uint8_t>(static_cast<size_t>(e->maxB) * e->cfg.n_images * 224 * 224 * 3);
    } catch (const std::exception& ex) {
      return e->fail(VLA_ERR_CUDA, ex.what());
    }
  }
  if (crop_scale != e->crop_scale) {  // captured graphs hold the old choice
    cudaDeviceSynchronize();
    for (auto& g : e->graphs)
      if (g.exec) cudaGraphExecDestroy(g.exec);
    e->graphs.clear();
  }
  e->crop_scale = crop_scale;
  return VLA_OK;
}

int vla_set_image_norm(vla_engine* e, const float* mean, const float* stdv) {
  if (!e) return VLA_ERR_INVALID;
  if (!mean || !stdv) return e->fail(VLA_ERR_INVALID, "set_image_norm: null statistics");
  vla::DeviceGuard guard(e->device);
  // ToTensor (u8 / 255), Normalize ((x - mean) / std) in fp32 like torchvision, then the bf16 cast of the caller
  std::vector<bf16> lut(2 * 3 * 256);
  for (int t = 0; t < 2; ++t)
    for (int c = 0; c < 3; ++c) {
      if (!(stdv[t * 3 + c] > 0.f)) return e->fail(VLA_ERR_INVALID, "set_image_norm: std must be positive");
      for (int v = 0; v < 256; ++v) {
        volatile float x = static_cast<float>(v) / 255.0f;
        volatile float y = x - mean[t * 3 + c];
        volatile float z = y / stdv[t * 3 + c];
        lut[(t * 3 + c) * 256 + v] = __float2bfloat16_rn(z);
      }
    }
  try {
    if (!e->img_lut) e->img_lut = e->dalloc<bf16>(lut.size());
  } catch (const std::exception& ex) {
    return e->fail(VLA_ERR_CUDA, ex.what());
  }
  if (cudaMemcpy(e->img_lut, lut.data(), lut.size() * sizeof(bf16), cudaMemcpyHostToDevice) != cudaSuccess)
    return e->fail(VLA_ERR_CUDA, "set_image_norm: copy failed");
  return VLA_OK;
}

static int predict_host_impl(vla_engine* e, const void* pixel_values, int is_u8, const int64_t* ext_ids,
                             const int32_t* aq_index, const int32_t* prompt_len, const float* proprio, int B, int L,
                             float* out_norm, float* out_unnorm, void* out_last_ha, void* stream) {
  if (!e) return VLA_ERR_INVALID;
  if (!e->finalized) return e->fail(VLA_ERR_NOT_FINALIZED, "vla_predict_host before vla_finalize");
  if (!pixel_values || !ext_ids || !aq_index || !proprio || !out_norm)
    return e->fail(VLA_ERR_INVALID, "vla_predict_host: null input/output pointer");
  if (B < 1 || B > e->maxB) return e->fail(VLA_ERR_INVALID, "vla_predict_host: batch outside [1, max_batch]");
  if (L < 1 || L > e->maxL) return e->fail(VLA_ERR_INVALID, "vla_predict_host: prompt length outside [1, max_prompt_len]");
  vla::DeviceGuard guard(e->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t pix_b = static_cast<size_t>(e->maxB) * 6 * e->cfg.n_images * 224 * 224 * 2;
  const size_t ids_b = static_cast<size_t>(e->maxB) * (e->maxL + N_AQ + 1) * 8;
  const size_t aq_b = ids_b / 2;
  const size_t prop_b = static_cast<size_t>(e->maxB) * e->P * 4;
  const size_t out_b = static_cast<size_t>(e->maxB) * e->T * e->A * 4;
  const size_t ha_b = static_cast<size_t>(e->maxB) * N_AQ * D_LLM * 2;
  if (!e->dev_pix) {
    try {
      e->dev_pix = e->dalloc<uint8_t>(pix_b);
      e->dev_ids = e->dalloc<uint8_t>(ids_b);
      e->dev_aq = e->dalloc<uint8_t>(aq_b);
      e->dev_prop = e->dalloc<uint8_t>(prop_b);
      e->dev_len = e->dalloc<uint8_t>(static_cast<size_t>(e->maxB) * 4);
      e->dev_out = e->dalloc<uint8_t>(2 * out_b);
      e->dev_ha = e->dalloc<uint8_t>(ha_b);
    } catch (const std::exception& ex) {
      return e->fail(VLA_ERR_CUDA, ex.what());
    }
  }
  const int Lext = L + N_AQ + 1;
  const size_t n_out = static_cast<size_t>(B) * e->T * e->A;
  cudaError_t ce = cudaSuccess;
  // uint8 frames are a quarter of the bf16 pixel_values (3 bytes per pixel instead of 2 towers x 3 channels x 2 bytes)
  const size_t in_bytes = is_u8 ? static_cast<size_t>(B) * e->cfg.n_images * 224 * 224 * 3
                                : static_cast<size_t>(B) * 6 * e->cfg.n_images * 224 * 224 * 2;
  ce = cudaMemcpyAsync(e->dev_pix, pixel_values, in_bytes, cudaMemcpyHostToDevice, s);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->dev_ids, ext_ids, static_cast<size_t>(B) * Lext * 8, cudaMemcpyHostToDevice, s);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->dev_aq, aq_index, static_cast<size_t>(B) * Lext * 4, cudaMemcpyHostToDevice, s);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->dev_prop, proprio, static_cast<size_t>(B) * e->P * 4, cudaMemcpyHostToDevice, s);
  if (prompt_len) {  // host array: validated here, so a bad length never reaches the device
    for (int b = 0; b < B; ++b)
      if (prompt_len[b] < 1 || prompt_len[b] > L)
        return e->fail(VLA_ERR_INVALID, "vla_predict_host: prompt_len[" + std::to_string(b) + "] outside [1, L]");
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(e->dev_len, prompt_len, static_cast<size_t>(B) * 4, cudaMemcpyHostToDevice, s);
  }
  if (ce != cudaSuccess) return e->fail(VLA_ERR_CUDA, std::string("H2D: ") + cudaGetErrorString(ce));
  float* d_norm = static_cast<float*>(e->dev_out);
  float* d_un = d_norm + static_cast<size_t>(e->maxB) * e->T * e->A;
  int rc = predict_impl(e, e->dev_pix, is_u8, static_cast<const int64_t*>(e->dev_ids), static_cast<const int32_t*>(e->dev_aq),
                       static_cast<const float*>(e->dev_prop), prompt_len ? static_cast<const int32_t*>(e->dev_len) : nullptr,
                       B, L, d_norm, d_un, out_last_ha ? e->dev_ha : nullptr, stream);
  if (rc) return rc;
  ce = cudaMemcpyAsync(out_norm, d_norm, n_out * 4, cudaMemcpyDeviceToHost, s);
  if (ce == cudaSuccess && out_unnorm) ce = cudaMemcpyAsync(out_unnorm, d_un, n_out * 4, cudaMemcpyDeviceToHost, s);
  if (ce == cudaSuccess && out_last_ha)
    ce = cudaMemcpyAsync(out_last_ha, e->dev_ha, static_cast<size_t>(B) * N_AQ * D_LLM * 2, cudaMemcpyDeviceToHost, s);
  int flag = 0;
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(&flag, e->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  if (ce != cudaSuccess) return e->fail(VLA_ERR_CUDA, std::string("predict: ") + cudaGetErrorString(ce));
  if (flag) {
    cudaMemset(e->err_flag, 0, sizeof(int));
    return e->fail(VLA_ERR_INVALID, flag == 1 ? "token id outside [0, vocab_size)"
                                    : (flag == 2 ? "ActionQuery index outside [0, 64)" : "prompt length outside [1, L]"));
  }
  return VLA_OK;
}

int vla_predict_host(vla_engine* e, const void* pixel_values, const int64_t* ext_ids, const int32_t* aq_index,
                     const int32_t* prompt_len, const float* proprio, int B, int L, float* out_norm, float* out_unnorm,
                     void* out_last_ha, void* stream) {
  return predict_host_impl(e, pixel_values, 0, ext_ids, aq_index, prompt_len, proprio, B, L, out_norm, out_unnorm,
                           out_last_ha, stream);
}

int vla_predict_host_u8(vla_engine* e, const uint8_t* images, const int64_t* ext_ids, const int32_t* aq_index,
                        const int32_t* prompt_len, const float* proprio, int B, int L, float* out_norm,
                        float* out_unnorm, void* out_last_ha, void* stream) {
  return predict_host_impl(e, images, 1, ext_ids, aq_index, prompt_len, proprio, B, L, out_norm, out_unnorm, out_last_ha,
                           stream);
}

int vla_get_tap(vla_engine* e, const char* name, void* dst, size_t capacity, size_t* bytes) {
  if (!e) return VLA_ERR_INVALID;
  if (!name || !e->lastB) return e->fail(VLA_ERR_INVALID, "get_tap: no forward has run yet");
  vla::DeviceGuard guard(e->device);
  const std::string n(name);
  const int B = e->lastB, L = e->lastL, NP = e->NP;
  const int S = NP + L + N_AQ + 1;
  const bf16* src = nullptr;
  size_t count = 0;
  bool strided_proj = false;
  if (n == "patches") { src = e->w_patches; count = static_cast<size_t>(B) * NP * D_VIS; }
  else if (n == "projected") { strided_proj = true; count = static_cast<size_t>(B) * NP * D_LLM; }
  else if (n == "llm_in") { src = e->hid[0]; count = static_cast<size_t>(B) * S * D_LLM; }
  else if (n.rfind("hidden.", 0) == 0) {
    const int i = atoi(n.c_str() + 7);
    if (i < 0 || i > e->cfg.llm_layers) return e->fail(VLA_ERR_INVALID, "get_tap: bad hidden index");
    src = e->hid[i]; count = static_cast<size_t>(B) * S * D_LLM;
  } else if (n.rfind("head_x.", 0) == 0) {
    const int i = atoi(n.c_str() + 7);
    if (i < 0 || i > 24) return e->fail(VLA_ERR_INVALID, "get_tap: bad head_x index");
    src = e->head_x[i]; count = static_cast<size_t>(B) * e->T * D_LLM;
  } else return e->fail(VLA_ERR_INVALID, "get_tap: unknown tap " + n);
  if (bytes) *bytes = count * 2;
  if (!dst) return VLA_OK;
  if (capacity < count * 2) return e->fail(VLA_ERR_INVALID, "get_tap: destination too small");
  cudaError_t ce;
  cudaDeviceSynchronize();
  if (strided_proj) {
    ce = cudaMemcpy2D(dst, static_cast<size_t>(NP) * D_LLM * 2, e->hid[0] + D_LLM, static_cast<size_t>(S) * D_LLM * 2,
                      static_cast<size_t>(NP) * D_LLM * 2, B, cudaMemcpyDefault);
  } else {
    ce = cudaMemcpy(dst, src, count * 2, cudaMemcpyDefault);
  }
  if (ce != cudaSuccess) return e->fail(VLA_ERR_CUDA, std::string("get_tap: ") + cudaGetErrorString(ce));
  return VLA_OK;
}

int vla_segment_timing(vla_engine* e, int enable) {
  if (!e) return VLA_ERR_INVALID;
  vla::DeviceGuard guard(e->device);
  if (enable && !e->seg_ev[0])
    for (int i = 0; i < 4; ++i)
      if (cudaEventCreate(&e->seg_ev[i]) != cudaSuccess) return e->fail(VLA_ERR_CUDA, "segment event creation failed");
  e->seg_on = enable ? 1 : 0;
  return VLA_OK;
}

int vla_segment_times(vla_engine* e, float* ms3) {
  if (!e) return VLA_ERR_INVALID;
  if (!ms3 || !e->seg_ev[0]) return e->fail(VLA_ERR_INVALID, "segment_times: timing was never enabled");
  vla::DeviceGuard guard(e->device);
  if (cudaEventSynchronize(e->seg_ev[3]) != cudaSuccess) return e->fail(VLA_ERR_CUDA, "segment_times: no timed forward yet");
  for (int i = 0; i < 3; ++i)
    if (cudaEventElapsedTime(&ms3[i], e->seg_ev[i], e->seg_ev[i + 1]) != cudaSuccess)
      return e->fail(VLA_ERR_CUDA, "segment_times: events not recorded");
  return VLA_OK;
}

// The device-pointer calls (vla_predict, vla_predict_u8) are asynchronous and cannot report what the kernels find:
// this call waits for `stream`, then reports (and clears) the device-side error flag - an out-of-range token id or
// ActionQuery index (the offending row of the LLM input was zero-filled) - and any CUDA error of the forward.
int vla_check_errors(vla_engine* e, void* stream) {
  if (!e) return VLA_ERR_INVALID;
  if (!e->finalized) return e->fail(VLA_ERR_NOT_FINALIZED, "vla_check_errors before vla_finalize");
  vla::DeviceGuard guard(e->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int flag = 0;
  cudaError_t ce = cudaMemcpyAsync(&flag, e->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
  if (ce != cudaSuccess) return e->fail(VLA_ERR_CUDA, std::string("forward failed: ") + cudaGetErrorString(ce));
  if (flag) {
    cudaMemsetAsync(e->err_flag, 0, sizeof(int), s);
    return e->fail(VLA_ERR_INVALID, flag == 1 ? "token id outside [0, vocab_size)"
                                    : (flag == 2 ? "ActionQuery index outside [0, 64)" : "prompt length outside [1, L]"));
  }
  return VLA_OK;
}

long long vla_last_launch_count(const vla_engine* e) { return e ? e->last_launches : 0; }

const char* vla_last_error(const vla_engine* e) { return e ? e->err.c_str() : "null engine"; }

void vla_destroy(vla_engine* e) {
  if (!e) return;
  vla::DeviceGuard guard(e->device);
  cudaDeviceSynchronize();
  if (e->pol_prof) {  // VLA_POLICY_PROF=1: phase stamps (clocks since the block's start) of the last fused policy launch
    long long h[24 * 8];
    if (cudaMemcpy(h, e->pol_prof, sizeof(h), cudaMemcpyDeviceToHost) == cudaSuccess) {
      fprintf(stderr, "policy_fused phases (clocks): blk  proj  attn  merge+ao  csync  oproj  csync+ln  ffn  csync+reload\n");
      for (int b = 0; b < 24; ++b) {
        const long long* t = h + b * 8;
        const long long next = b + 1 < 24 ? h[(b + 1) * 8] : t[7];
        fprintf(stderr, "policy_fused %2d  %lld %lld %lld %lld %lld %lld %lld %lld\n", b, t[1] - t[0], t[2] - t[1], t[3] - t[2],
                t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6], next - t[7]);
      }
    }
  }
  for (auto& g : e->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (e->gev_in) cudaEventDestroy(e->gev_in);
  if (e->gev_out) cudaEventDestroy(e->gev_out);
  for (int i = 0; i < 4; ++i)
    if (e->seg_ev[i]) cudaEventDestroy(e->seg_ev[i]);
  if (e->gstream) cudaStreamDestroy(e->gstream);
  for (cudaEvent_t ev : e->ev_layer) cudaEventDestroy(ev);
  for (cudaEvent_t ev : e->ev_kv) cudaEventDestroy(ev);
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  if (e->side) cudaStreamDestroy(e->side);
  if (e->pol) cudaStreamDestroy(e->pol);
  if (e->ev_pol) cudaEventDestroy(e->ev_pol);
  for (void* p : e->allocs) cudaFree(p);
  e->free_masters();
  delete e;
}

}  // extern "C"
