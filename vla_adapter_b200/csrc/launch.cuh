// Kernel launch helper: every forward kernel is launched with programmatic dependent launch (PDL) allowed, so that the
// next kernel's CTAs are scheduled - and run their prologue (barrier init, TMEM allocation, tensor-map prefetch) - while
// the previous kernel drains.  Each kernel calls pdl_wait() before its first access to global memory (it then sees
// everything its predecessors wrote) and pdl_launch_dependents() right after.  VLA_PDL=0 switches the attribute off.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

namespace vla {

// Measured on B200: PDL shortens the bs=1 forward by ~5 % (prologues overlap the 6-8 us kernels) but costs ~4 % at
// bs=64 (early-resident CTAs of the next kernel compete with the draining one), so the engine switches it per call:
// on for small batches, off for large ones.  VLA_PDL=0/1 forces it.
inline int& pdl_flag() {
  static int v = 1;
  return v;
}
inline void pdl_set(bool on) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("VLA_PDL");
    forced = e ? (atoi(e) != 0 ? 1 : 0) : 2;
  }
  pdl_flag() = forced == 2 ? (on ? 1 : 0) : forced;
}
inline bool pdl_enabled() { return pdl_flag() != 0; }

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
