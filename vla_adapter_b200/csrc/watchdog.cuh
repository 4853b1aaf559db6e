// Host interface of the device watchdog (see common.cuh for the device side, watchdog.cu for the host side).
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace vla {

struct WdBuf;

// Installs the process-wide record buffer into every kernel translation unit for the CURRENT device (idempotent).
int watchdog_install_current_device(const char** err);
int watchdog_set_timeout_ms(unsigned long long ms);  // current device
std::string watchdog_report();

// per-TU setters (each TU has its own copy of the device-side pointer: the library is built without -rdc)
// (dev_ptr == nullptr: only the time limit is updated)
cudaError_t gemm_set_watchdog(WdBuf* dev_ptr, unsigned long long timeout_ms);
cudaError_t fa_set_watchdog(WdBuf* dev_ptr, unsigned long long timeout_ms);

}  // namespace vla
