// Internal definition of the env handle, shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dronecu.h"
#include "env_kernels.cuh"

struct dronecu_env {
  dronecu_config cfg;
  dronecu::EnvParams P;
  int device;
  int64_t n;
  uint64_t t;          // global step index (Philox counter for ACTION / NOISE streams)
  float4* planes;      // 5 * n quads
  dronecu::StatePlanes sp;
  dronecu::StatSlot* stats;
  uint64_t launches;
  uint64_t env_steps;
  // device + stream used by the *_host entry points (lazily created)
  bool io_ready;             // every staging buffer below exists (set only after ALL allocations succeeded)
  cudaStream_t io_stream;
  cudaStream_t io_stream2;   // second stream: chunked step_host overlaps H2D, kernel and D2H
  float *d_act, *d_obs, *d_rew, *d_term;
  uint8_t *d_done, *d_trunc, *d_mask;
  float* d_ep_r;
  int32_t* d_ep_l;
  float *d_view_f;     // 16 floats per env scratch for get/set_state_host
  int32_t* d_view_i;   // 3 ints per env
};

