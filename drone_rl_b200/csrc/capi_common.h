// Shared helpers of the C-ABI translation units: error text, CUDA error check, device guard.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "../../include/dronecu.h"

namespace dronecu {

void set_error(const std::string& s);

inline int fail(int code, const std::string& msg) {
  set_error(msg);
  return code;
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
};

}  // namespace dronecu

#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      cudaGetLastError();                                                                    \
      return ::dronecu::fail(DRONECU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    }                                                                                        \
  } while (0)
