// Kernel launch helper: every forward kernel is launched with programmatic dependent launch (PDL) allowed, so that the
// next kernel's CTAs are scheduled - and run their prologue (barrier init, TMEM allocation, tensor-map prefetch) - while
// the previous kernel drains.  Each kernel calls pdl_wait() before its first access to global memory (it then sees
// everything its predecessors wrote) and pdl_launch_dependents() right after.  VLA_PDL=0 switches the attribute off.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

namespace vla {

// Measured on B200: PDL shortens the bs=1 forward by ~5 % (prologues overlap the 6-8 us kernels) but costs ~4 % at
// bs=64 (early-resident CTAs of the next kernel compete with the draining one), so the engine chooses per forward:
// on for small batches, off for large ones.  The choice is the calling thread's (PdlScope in forward()), not a
// process global: engines on different devices / threads do not see each other's setting.  VLA_PDL=0/1 forces it.
inline int& pdl_flag() {
  static thread_local int v = 0;
  return v;
}
inline int pdl_forced() {
  static const int forced = [] {
    const char* e = getenv("VLA_PDL");
    return e ? (atoi(e) != 0 ? 1 : 0) : 2;
  }();
  return forced;
}
struct PdlScope {
  int saved;
  explicit PdlScope(bool on) : saved(pdl_flag()) {
    const int f = pdl_forced();
    pdl_flag() = f == 2 ? (on ? 1 : 0) : f;
  }
  ~PdlScope() { pdl_flag() = saved; }
  PdlScope(const PdlScope&) = delete;
  PdlScope& operator=(const PdlScope&) = delete;
};
inline bool pdl_enabled() { return pdl_flag() != 0; }

// Per-device one-time state (cudaFuncSetAttribute is per device, and so is the SM count): indexed by the CURRENT
// device of the calling thread.  The engine makes its device current at every C-ABI entry point.
constexpr int VLA_MAX_DEVICES = 64;
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= VLA_MAX_DEVICES) dev = 0;
  return dev;
}
struct PerDeviceFlag {
  bool done[VLA_MAX_DEVICES] = {};
  bool& here() { return done[current_device()]; }
};
inline int device_num_sms() {
  static int n[VLA_MAX_DEVICES] = {};
  const int dev = current_device();
  if (!n[dev]) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    n[dev] = v > 0 ? v : 148;
  }
  return n[dev];
}
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// cluster_x > 1 launches thread-block clusters of that many CTAs along x (CTA pairs for cta_group::2 kernels).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(int cluster_x, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args... args) {
  return launch_kernel_cluster(1, kernel, grid, block, smem, stream, args...);
}

}  // namespace vla
