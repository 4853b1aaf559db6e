// Host side of the device watchdog (common.cuh): one host-mapped record buffer per process, installed into every
// translation unit that waits on mbarriers, for every device an engine is created on; a report formatter; and a test
// kernel that deadlocks on purpose (tests/test_watchdog_gpu.py runs it in a child process).
#include "../../include/vla_b200.h"
#include "common.cuh"
#include "watchdog.cuh"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

namespace vla {

namespace {

std::mutex g_mu;
WdBuf* g_host = nullptr;        // cudaHostAlloc(portable | mapped): readable by the host after a device trap
bool g_installed[64] = {false};
unsigned long long g_timeout_ms = 10000;

const char* kernel_name(uint32_t k) {
  switch (k) {
    case WD_K_GEMM1: return "gemm_bf16_tcgen05_kernel<1>";
    case WD_K_GEMM2: return "gemm_bf16_tcgen05_kernel<2>";
    case WD_K_FA64: return "fa_tcgen05_kernel<64>";
    case WD_K_FA72: return "fa_tcgen05_kernel<72>";
    case WD_K_POLICY: return "policy_block_kernel";
    case WD_K_SELFTEST: return "wd_selftest_kernel";
    default: return "unknown kernel";
  }
}
const char* role_name(uint32_t k, uint32_t warp) {
  if (k == WD_K_GEMM1 || k == WD_K_GEMM2) return warp == 0 ? "TMA producer" : (warp == 1 ? "MMA issuer" : "epilogue");
  if (k == WD_K_FA64 || k == WD_K_FA72)
    return warp == 0 ? "TMA producer" : (warp == 1 ? "MMA issuer A" : (warp == 2 ? "MMA issuer B" : (warp < 12 ? "softmax A" : "softmax B")));
  return "worker";
}
const char* bar_name(uint32_t k, uint32_t kind) {
  static const char* gemm[] = {"?", "empty", "full", "tmem_empty", "tmem_full"};
  static const char* fa[] = {"?", "q_empty", "kv_empty", "kv_full", "q_full", "s_free", "p_ready", "o_empty", "s_full",
                             "pv_done", "turn", "o_full"};
  if ((k == WD_K_GEMM1 || k == WD_K_GEMM2) && kind < 5) return gemm[kind];
  if ((k == WD_K_FA64 || k == WD_K_FA72) && kind < 12) return fa[kind];
  return "barrier";
}

// A CTA of one working warp that waits on a barrier nobody arrives at, and the monitor warp that must catch it.
__global__ void wd_selftest_kernel() {
  __shared__ uint64_t bars[2];
  __shared__ uint32_t notes[1];
  const uint32_t stuck = smem_u32(&bars[0]), done = smem_u32(&bars[1]), note = smem_u32(&notes[0]);
  if (threadIdx.x == 0) {
    mbar_init(stuck, 1);
    mbar_init(done, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    wd_note(note, VLA_WD_NOTE(1, 0, 0, 7));
    mbar_wait(stuck, 0);
    if (threadIdx.x == 0) mbar_arrive(done);
  } else {
    wd_monitor(done, note, 1, WD_K_SELFTEST, [&](uint32_t, uint32_t) { return stuck; });
  }
}

}  // namespace

int watchdog_install_current_device(const char** err) {
  std::lock_guard<std::mutex> lk(g_mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    if (err) *err = "watchdog: cudaGetDevice failed";
    return -4;
  }
  if (!g_host) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, sizeof(WdBuf), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
      cudaGetLastError();
      if (err) *err = "watchdog: cudaHostAlloc failed";
      return -4;
    }
    std::memset(p, 0, sizeof(WdBuf));
    g_host = static_cast<WdBuf*>(p);
    if (const char* ms = getenv("VLA_WATCHDOG_MS")) {
      const unsigned long long v = strtoull(ms, nullptr, 10);
      if (v) g_timeout_ms = v;
    }
  }
  if (g_installed[dev]) return 0;
  WdBuf* dptr = nullptr;
  if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&dptr), g_host, 0) != cudaSuccess) {
    cudaGetLastError();
    if (err) *err = "watchdog: cudaHostGetDevicePointer failed";
    return -4;
  }
  if (gemm_set_watchdog(dptr, g_timeout_ms) != cudaSuccess || fa_set_watchdog(dptr, g_timeout_ms) != cudaSuccess ||
      wd_set_buffer_this_tu(dptr) != cudaSuccess || wd_set_limit_this_tu(g_timeout_ms) != cudaSuccess) {
    cudaGetLastError();
    if (err) *err = "watchdog: installing the record buffer failed";
    return -4;
  }
  g_installed[dev] = true;
  return 0;
}

int watchdog_set_timeout_ms(unsigned long long ms) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!ms) return -1;
  g_timeout_ms = ms;  // devices installed later pick it up; the current one is updated now
  if (gemm_set_watchdog(nullptr, ms) != cudaSuccess || fa_set_watchdog(nullptr, ms) != cudaSuccess ||
      wd_set_limit_this_tu(ms) != cudaSuccess) {
    cudaGetLastError();
    return -4;
  }
  return 0;
}

// Human-readable list of the records written so far ("" when there are none).
std::string watchdog_report() {
  std::lock_guard<std::mutex> lk(g_mu);
  std::string out;
  if (!g_host) return out;
  for (unsigned int i = 0; i < WD_MAX_RECORDS; ++i) {
    const volatile WdRecord& r = g_host->rec[i];
    if (r.magic != WD_MAGIC) continue;
    const uint32_t k = r.kernel, note = r.note;
    char line[320];
    if (r.smid == 0xffffffffu) {  // minimal record (wd_monitor_min)
      snprintf(line, sizeof(line), "watchdog: %s CTA %u did not finish within the time limit\n", kernel_name(k), r.block);
    } else if (r.warp >= 0x100u) {
      static const char* fa_bars[] = {"q_full[0]", "q_full[1]", "q_full[2]", "q_full[3]", "q_empty[0]", "q_empty[1]",
                                      "q_empty[2]", "q_empty[3]", "s_full[A]", "s_full[B]", "p_ready[A]", "p_ready[B]",
                                      "o_full[A]", "o_full[B]", "o_empty[A]", "o_empty[B]", "turn[A]", "turn[B]",
                                      "s_free[A]", "s_free[B]", "pv_done[A]", "pv_done[B]"};
      const uint32_t i = r.warp - 0x100u;
      char name[32];
      if ((k == WD_K_FA64 || k == WD_K_FA72) && i < 22) snprintf(name, sizeof(name), "%s", fa_bars[i]);
      else if (k == WD_K_FA64 || k == WD_K_FA72) snprintf(name, sizeof(name), "kv ring barrier %u", i - 22);
      else snprintf(name, sizeof(name), "barrier %u", i);
      snprintf(line, sizeof(line), "watchdog: %s CTA %u (SM %u): %s word 0x%08x%08x\n", kernel_name(k), r.block, r.smid, name,
               r.bar_hi, r.bar_lo);
    } else if (note == 0xffffffffu) {
      snprintf(line, sizeof(line), "watchdog: %s CTA %u (SM %u) warp %u (%s): finished its role\n", kernel_name(k), r.block,
               r.smid, r.warp, role_name(k, r.warp));
    } else if (note == 0) {
      snprintf(line, sizeof(line), "watchdog: %s CTA %u (SM %u) warp %u (%s): never reached a wait\n", kernel_name(k),
               r.block, r.smid, r.warp, role_name(k, r.warp));
    } else {
      snprintf(line, sizeof(line),
               "watchdog: %s CTA %u (SM %u) warp %u (%s): last wait on %s[%u] parity %u at step %u; barrier word 0x%08x%08x\n",
               kernel_name(k), r.block, r.smid, r.warp, role_name(k, r.warp), bar_name(k, note >> 24), (note >> 16) & 0xffu,
               (note >> 15) & 1u, note & 0x7fffu, r.bar_hi, r.bar_lo);
    }
    out += line;
  }
  return out;
}

}  // namespace vla

extern "C" {

int vla_watchdog_report(char* buf, size_t capacity) {
  const std::string s = vla::watchdog_report();
  if (buf && capacity) {
    const size_t n = s.size() < capacity - 1 ? s.size() : capacity - 1;
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return static_cast<int>(s.size());
}

int vla_watchdog_set_timeout_ms(int ms) {
  if (ms <= 0) return VLA_ERR_INVALID;
  const char* err = nullptr;
  if (vla::watchdog_install_current_device(&err)) return VLA_ERR_CUDA;
  return vla::watchdog_set_timeout_ms(static_cast<unsigned long long>(ms)) ? VLA_ERR_CUDA : VLA_OK;
}

int vla_watchdog_selftest(void* stream) {
  const char* err = nullptr;
  if (vla::watchdog_install_current_device(&err)) return VLA_ERR_CUDA;
  vla::wd_selftest_kernel<<<2, 64, 0, static_cast<cudaStream_t>(stream)>>>();
  // the kernel can only end by trapping: the synchronize reports the launch failure
  return cudaStreamSynchronize(static_cast<cudaStream_t>(stream)) == cudaSuccess ? VLA_OK : VLA_ERR_CUDA;
}

}  // extern "C"
