// C ABI of the environment half of libdronecu.so (include/dronecu.h): handle management,
// parameter digestion, kernel dispatch, host-buffer entry points.  No torch types, no CPU
// fallback: every entry point needs a CUDA device and reports a CUDA error otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/dronecu.h"
#include "capi_common.h"
#include "env_kernels.cuh"
#include "env_handle.h"

using namespace dronecu;

namespace dronecu {
thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }
}  // namespace dronecu

extern "C" int dronecu_version(void) { return DRONECU_VERSION; }
extern "C" const char* dronecu_last_error(void) { return g_last_error.c_str(); }

static void fill_common(dronecu_config* c) {
  std::memset(c, 0, sizeof(*c));
  c->dt = 0.02; c->mass = 1.0; c->gravity = 9.81;
  c->inertia[0] = 0.005; c->inertia[1] = 0.005; c->inertia[2] = 0.01;
  c->arm_length = 0.5; c->k_yaw = 0.01;
  c->reward_scale = 0.01; c->bonus = 1.0; c->z_floor = 0.0; c->r_max = 50.0;
  c->fixed_target[0] = 0.0; c->fixed_target[1] = 0.0; c->fixed_target[2] = 10.0;
  c->fixed_start[0] = 0.1; c->fixed_start[1] = 0.1; c->fixed_start[2] = 0.1;
  c->start_z = 1.0; c->target_z = 1.0; c->curriculum_step = 0.1; c->curriculum_period = 2000;
}

extern "C" void dronecu_config_single(dronecu_config* c) {
  fill_common(c);
  c->bonus_radius = 0.05; c->max_steps = 200; c->obs_dim = 15;
  c->flags = DRONECU_RANDOMIZED | DRONECU_AUTORESET;
}

extern "C" void dronecu_config_vector(dronecu_config* c) {
  fill_common(c);
  c->bonus_radius = 1.0; c->max_steps = 1000; c->obs_dim = 12;
  c->flags = 0;
}

static EnvParams digest(const dronecu_config& c, uint64_t seed, int64_t env_offset) {
  EnvParams P;
  std::memset(&P, 0, sizeof(P));
  P.dt = (float)c.dt;
  P.inv_mass = (float)(1.0 / c.mass);
  P.gravity = (float)c.gravity;
  P.lever = (float)(c.arm_length / std::sqrt(2.0));     // drone.py:114
  P.k_yaw = (float)c.k_yaw;
  P.dI_roll = (float)(c.inertia[1] - c.inertia[2]);      // drone.py:136
  P.dI_pitch = (float)(c.inertia[2] - c.inertia[0]);     // drone.py:137
  P.dI_yaw = (float)(c.inertia[0] - c.inertia[1]);       // drone.py:138
  for (int k = 0; k < 3; ++k) {
    P.invI[k] = (float)(1.0 / c.inertia[k]);
    P.fixed_target[k] = (float)c.fixed_target[k];
    P.fixed_start[k] = (float)c.fixed_start[k];
  }
  P.reward_scale = (float)c.reward_scale;
  P.bonus_radius = (float)c.bonus_radius;
  P.bonus = (float)c.bonus;
  P.z_floor = (float)c.z_floor;
  P.r_max = (float)c.r_max;
  P.start_z = (float)c.start_z;
  P.target_z = (float)c.target_z;
  P.motor_max = (float)(3.0 * c.mass * c.gravity / 4.0);  // drone.py:263
  P.curriculum_step = c.curriculum_step;
  P.curriculum_period = c.curriculum_period;
  P.max_steps = c.max_steps;
  P.inv_curriculum_period = (float)(1.0 / (double)c.curriculum_period);
  P.r_max_sq = (float)(c.r_max * c.r_max);
  P.seed = seed;
  philox_expand_key(seed, P.keys);
  P.env_offset = (uint64_t)env_offset;
  return P;
}

// CTA size: 256 threads when there are enough envs to fill the 148 SMs several times over;
// small batches (configs[1]: 4096 envs) use one warp per CTA so they spread over all SMs.
static inline unsigned block_for(int64_t n) { return n >= (int64_t)148 * kBlock * 2 ? kBlock : (n >= 148 * 64 * 2 ? 64 : 32); }
static inline unsigned grid_for(int64_t n, unsigned block = kBlock) { return (unsigned)((n + block - 1) / block); }

template <int D, bool R>
static void launch_reset_t(dronecu_env* e, const uint8_t* mask, float* obs, int zero_first, cudaStream_t st) {
  const unsigned blk = block_for(e->n);
  reset_kernel<D, R><<<grid_for(e->n, blk), blk, 0, st>>>(e->sp, e->P, e->n, mask, obs, zero_first);
}

static int launch_reset(dronecu_env* e, const uint8_t* mask, float* obs, int zero_first, cudaStream_t st) {
  const bool R = e->cfg.flags & DRONECU_RANDOMIZED;
  if (e->cfg.obs_dim == 15) { R ? launch_reset_t<15, true>(e, mask, obs, zero_first, st) : launch_reset_t<15, false>(e, mask, obs, zero_first, st); }
  else { R ? launch_reset_t<12, true>(e, mask, obs, zero_first, st) : launch_reset_t<12, false>(e, mask, obs, zero_first, st); }
  e->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return DRONECU_OK;
}

template <int D, bool R, bool AR>
static void launch_rollout_t(const RolloutArgs& a, int mode, bool record, unsigned grid, unsigned blk, cudaStream_t st) {
  if (mode == DRONECU_ACTIONS_STREAMED) {
    if (record) rollout_kernel<D, R, AR, 0, true><<<grid, blk, 0, st>>>(a);
    else rollout_kernel<D, R, AR, 0, false><<<grid, blk, 0, st>>>(a);
  } else {
    if (record) rollout_kernel<D, R, AR, 1, true><<<grid, blk, 0, st>>>(a);
    else rollout_kernel<D, R, AR, 1, false><<<grid, blk, 0, st>>>(a);
  }
}

static int launch_rollout(dronecu_env* e, const RolloutArgs& a, int mode, cudaStream_t st) {
  const bool R = e->cfg.flags & DRONECU_RANDOMIZED, AR = e->cfg.flags & DRONECU_AUTORESET;
  const unsigned blk = block_for(e->n), g = grid_for(e->n, blk);
  const int D = e->cfg.obs_dim;
  // the "rollout record" specialisation: exactly next_obs + reward + done (+ the in-kernel action)
  const bool record = a.next_obs && a.reward && a.done && !a.truncated && !a.terminal_obs && !a.episode_r &&
                      !a.episode_l && ((mode == DRONECU_ACTIONS_UNIFORM) == (a.out_actions != nullptr));
#define DISPATCH(DD, RR, AA) if (D == DD && R == RR && AR == AA) launch_rollout_t<DD, RR, AA>(a, mode, record, g, blk, st);
  DISPATCH(15, true, true) DISPATCH(15, true, false) DISPATCH(15, false, true) DISPATCH(15, false, false)
  DISPATCH(12, true, true) DISPATCH(12, true, false) DISPATCH(12, false, true) DISPATCH(12, false, false)
#undef DISPATCH
  e->launches += 1;
  CUDA_TRY(cudaGetLastError());
  e->t += (uint64_t)a.K;
  e->env_steps += (uint64_t)a.K * (uint64_t)e->n;
  return DRONECU_OK;
}

extern "C" int dronecu_create(const dronecu_config* cfg, int device, int64_t n_envs, int64_t env_offset,
                              uint64_t seed, dronecu_env** out) {
  if (!cfg || !out || n_envs <= 0 || env_offset < 0) return fail(DRONECU_ERR_INVALID, "dronecu_create: bad argument");
  if (cfg->obs_dim != 12 && cfg->obs_dim != 15) return fail(DRONECU_ERR_INVALID, "obs_dim must be 12 or 15");
  if (cfg->max_steps <= 0 || cfg->curriculum_period <= 0 || !(cfg->mass > 0) || !(cfg->dt > 0))
    return fail(DRONECU_ERR_INVALID, "dronecu_create: non-positive max_steps/curriculum_period/mass/dt");
  if (n_envs > ((int64_t)1 << 30)) return fail(DRONECU_ERR_INVALID, "n_envs too large (at most 2^30 envs per handle: 32-bit thread indices)");
  int ndev = 0;
  cudaError_t err = cudaGetDeviceCount(&ndev);
  if (err != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(DRONECU_ERR_CUDA, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(err));
  }
  if (device < 0 || device >= ndev) return fail(DRONECU_ERR_INVALID, "device index out of range");
  DeviceGuard guard(device);
  dronecu_env* e = new (std::nothrow) dronecu_env();
  if (!e) return fail(DRONECU_ERR_ALLOC, "host allocation failed");
  std::memset(e, 0, sizeof(*e));
  e->cfg = *cfg; e->device = device; e->n = n_envs;
  e->P = digest(*cfg, seed, env_offset);
  err = cudaMalloc(&e->planes, sizeof(float4) * 5 * (size_t)n_envs);
  if (err != cudaSuccess) { delete e; cudaGetLastError(); return fail(DRONECU_ERR_ALLOC, std::string("cudaMalloc state: ") + cudaGetErrorString(err)); }
  err = cudaMalloc(&e->stats, sizeof(StatSlot) * kStatSlots);
  if (err != cudaSuccess) { cudaFree(e->planes); delete e; cudaGetLastError(); return fail(DRONECU_ERR_ALLOC, "cudaMalloc stats"); }
  for (int k = 0; k < 5; ++k) e->sp.q[k] = e->planes + (size_t)k * n_envs;
  cudaMemsetAsync(e->stats, 0, sizeof(StatSlot) * kStatSlots, 0);
  // the constructor's reset: drone.py:46 / vectorized_drone.py:36  (ep_num 0 -> 1)
  int rc = launch_reset(e, nullptr, nullptr, /*zero_first=*/1, 0);
  if (rc == DRONECU_OK) { err = cudaStreamSynchronize(0); if (err != cudaSuccess) rc = fail(DRONECU_ERR_CUDA, cudaGetErrorString(err)); }
  if (rc != DRONECU_OK) { cudaFree(e->planes); cudaFree(e->stats); delete e; return rc; }
  *out = e;
  return DRONECU_OK;
}

static void free_io(dronecu_env* e);

extern "C" int dronecu_destroy(dronecu_env* e) {
  if (!e) return DRONECU_OK;
  DeviceGuard guard(e->device);
  cudaDeviceSynchronize();
  cudaFree(e->planes); cudaFree(e->stats);
  free_io(e);
  if (e->io_stream) cudaStreamDestroy(e->io_stream);
  if (e->io_stream2) cudaStreamDestroy(e->io_stream2);
  cudaGetLastError();
  delete e;
  return DRONECU_OK;
}

extern "C" int64_t dronecu_num_envs(const dronecu_env* e) { return e ? e->n : 0; }
extern "C" int dronecu_obs_dim(const dronecu_env* e) { return e ? e->cfg.obs_dim : 0; }
extern "C" int64_t dronecu_global_step(const dronecu_env* e) { return e ? (int64_t)e->t : 0; }
extern "C" int dronecu_set_global_step(dronecu_env* e, int64_t t) {
  if (!e || t < 0) return fail(DRONECU_ERR_INVALID, "bad argument");
  e->t = (uint64_t)t;
  return DRONECU_OK;
}
extern "C" double dronecu_motor_max(const dronecu_env* e) { return e ? 3.0 * e->cfg.mass * e->cfg.gravity / 4.0 : 0.0; }
extern "C" uint64_t dronecu_launch_count(const dronecu_env* e) { return e ? e->launches : 0; }

extern "C" int dronecu_reset(dronecu_env* e, const uint8_t* d_mask, float* d_obs, void* stream) {
  if (!e) return fail(DRONECU_ERR_INVALID, "null handle");
  DeviceGuard guard(e->device);
  return launch_reset(e, d_mask, d_obs, 0, (cudaStream_t)stream);
}

extern "C" int dronecu_rollout(dronecu_env* e, int K, int action_mode, const float* d_actions,
                               const dronecu_rollout_out* out, void* stream) {
  if (!e) return fail(DRONECU_ERR_INVALID, "null handle");
  if (K <= 0) return fail(DRONECU_ERR_INVALID, "K must be positive");
  if (action_mode != DRONECU_ACTIONS_STREAMED && action_mode != DRONECU_ACTIONS_UNIFORM)
    return fail(DRONECU_ERR_INVALID, "unknown action_mode");
  if (action_mode == DRONECU_ACTIONS_STREAMED && !d_actions) return fail(DRONECU_ERR_INVALID, "streamed actions need d_actions");
  if (d_actions && (reinterpret_cast<uintptr_t>(d_actions) & 15)) return fail(DRONECU_ERR_INVALID, "d_actions must be 16-byte aligned");
  if (out && out->d_actions && (reinterpret_cast<uintptr_t>(out->d_actions) & 15)) return fail(DRONECU_ERR_INVALID, "out->d_actions must be 16-byte aligned");
  DeviceGuard guard(e->device);
  RolloutArgs a;
  std::memset(&a, 0, sizeof(a));
  a.state = e->sp; a.P = e->P; a.n = e->n; a.K = K; a.t0 = e->t; a.stats = e->stats;
  a.actions = reinterpret_cast<const float4*>(d_actions);
  if (out) {
    a.obs0 = out->d_obs0; a.next_obs = out->d_next_obs;
    a.out_actions = reinterpret_cast<float4*>(out->d_actions);
    a.reward = out->d_reward; a.done = out->d_done; a.truncated = out->d_truncated;
  }
  return launch_rollout(e, a, action_mode, (cudaStream_t)stream);
}

extern "C" int dronecu_step(dronecu_env* e, const float* d_actions, const dronecu_step_out* out, void* stream) {
  if (!e) return fail(DRONECU_ERR_INVALID, "null handle");
  if (!d_actions) return fail(DRONECU_ERR_INVALID, "dronecu_step: d_actions is NULL");
  if (reinterpret_cast<uintptr_t>(d_actions) & 15) return fail(DRONECU_ERR_INVALID, "d_actions must be 16-byte aligned");
  DeviceGuard guard(e->device);
  RolloutArgs a;
  std::memset(&a, 0, sizeof(a));
  a.state = e->sp; a.P = e->P; a.n = e->n; a.K = 1; a.t0 = e->t; a.stats = e->stats;
  a.actions = reinterpret_cast<const float4*>(d_actions);
  if (out) {
    a.next_obs = out->d_obs; a.reward = out->d_reward; a.done = out->d_done; a.truncated = out->d_truncated;
    a.terminal_obs = out->d_terminal_obs; a.episode_r = out->d_episode_r; a.episode_l = out->d_episode_l;
  }
  return launch_rollout(e, a, DRONECU_ACTIONS_STREAMED, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------------------------
// The *_host entry points run on the handle's own stream.  They first wait for everything
// already queued on the device (work the caller enqueued on other streams through the
// device-pointer entry points), and they return only after their own copies have landed.
static void free_io(dronecu_env* e) {
  cudaFree(e->d_act); cudaFree(e->d_obs); cudaFree(e->d_rew); cudaFree(e->d_term);
  cudaFree(e->d_done); cudaFree(e->d_trunc); cudaFree(e->d_mask); cudaFree(e->d_ep_r); cudaFree(e->d_ep_l);
  cudaFree(e->d_view_f); cudaFree(e->d_view_i);
  e->d_act = e->d_obs = e->d_rew = e->d_term = e->d_ep_r = e->d_view_f = nullptr;
  e->d_done = e->d_trunc = e->d_mask = nullptr;
  e->d_ep_l = e->d_view_i = nullptr;
  cudaGetLastError();
}

static int alloc_io(dronecu_env* e) {
  const size_t n = (size_t)e->n, D = (size_t)e->cfg.obs_dim;
  if (!e->io_stream) CUDA_TRY(cudaStreamCreateWithFlags(&e->io_stream, cudaStreamNonBlocking));
  if (!e->io_stream2) CUDA_TRY(cudaStreamCreateWithFlags(&e->io_stream2, cudaStreamNonBlocking));
  CUDA_TRY(cudaMalloc(&e->d_act, n * 4 * sizeof(float)));
  CUDA_TRY(cudaMalloc(&e->d_obs, n * D * sizeof(float)));
  CUDA_TRY(cudaMalloc(&e->d_term, n * D * sizeof(float)));
  CUDA_TRY(cudaMalloc(&e->d_rew, n * sizeof(float)));
  CUDA_TRY(cudaMalloc(&e->d_done, n));
  CUDA_TRY(cudaMalloc(&e->d_trunc, n));
  CUDA_TRY(cudaMalloc(&e->d_mask, n));
  CUDA_TRY(cudaMalloc(&e->d_ep_r, n * sizeof(float)));
  CUDA_TRY(cudaMalloc(&e->d_ep_l, n * sizeof(int32_t)));
  CUDA_TRY(cudaMalloc(&e->d_view_f, n * 16 * sizeof(float)));
  CUDA_TRY(cudaMalloc(&e->d_view_i, n * 3 * sizeof(int32_t)));
  return DRONECU_OK;
}

static int ensure_io(dronecu_env* e) {
  CUDA_TRY(cudaDeviceSynchronize());
  if (e->io_ready) return DRONECU_OK;
  // a failed allocation (out of memory at large n) must not leave a half-built set behind: the next call would
  // otherwise run with NULL staging buffers.  Free whatever exists, report, and let the caller retry.
  const int rc = alloc_io(e);
  if (rc != DRONECU_OK) { free_io(e); return rc; }
  e->io_ready = true;
  return DRONECU_OK;
}

// One env step over the sub-range [first, first + count) of the handle's envs (K = 1 only: every
// [n, .] output row block is addressed by env index, so a range is a pointer offset).  Does not advance
// the handle's step counters -- the caller does that once per full step.
static int launch_step_range(dronecu_env* e, const float* d_actions, const dronecu_step_out& d, int64_t first,
                             int64_t count, cudaStream_t st) {
  const int64_t D = e->cfg.obs_dim;
  RolloutArgs a;
  std::memset(&a, 0, sizeof(a));
  a.state = e->sp;
  for (int k = 0; k < 5; ++k) a.state.q[k] += first;
  a.P = e->P; a.P.env_offset += (uint64_t)first;
  a.n = count; a.K = 1; a.t0 = e->t; a.stats = e->stats;
  a.actions = reinterpret_cast<const float4*>(d_actions) + first;
  a.next_obs = d.d_obs ? d.d_obs + first * D : nullptr;
  a.reward = d.d_reward ? d.d_reward + first : nullptr;
  a.done = d.d_done ? d.d_done + first : nullptr;
  a.truncated = d.d_truncated ? d.d_truncated + first : nullptr;
  a.terminal_obs = d.d_terminal_obs ? d.d_terminal_obs + first * D : nullptr;
  a.episode_r = d.d_episode_r ? d.d_episode_r + first : nullptr;
  a.episode_l = d.d_episode_l ? d.d_episode_l + first : nullptr;
  const uint64_t t_keep = e->t, s_keep = e->env_steps;
  dronecu_env tmp_view = *e;          // launch_rollout reads n / cfg from the handle: use a shallow view
  tmp_view.n = count;
  int rc = launch_rollout(&tmp_view, a, DRONECU_ACTIONS_STREAMED, st);
  e->launches += 1;
  e->t = t_keep; e->env_steps = s_keep;
  return rc;
}

extern "C" int dronecu_step_host(dronecu_env* e, const float* h_actions, const dronecu_step_out* h) {
  if (!e || !h_actions || !h) return fail(DRONECU_ERR_INVALID, "dronecu_step_host: null argument");
  DeviceGuard guard(e->device);
  int rc = ensure_io(e);
  if (rc) return rc;
  const int64_t n = e->n, D = e->cfg.obs_dim;
  dronecu_step_out d;
  d.d_obs = h->d_obs ? e->d_obs : nullptr;
  d.d_reward = h->d_reward ? e->d_rew : nullptr;
  d.d_done = h->d_done ? e->d_done : nullptr;
  d.d_truncated = h->d_truncated ? e->d_trunc : nullptr;
  d.d_terminal_obs = h->d_terminal_obs ? e->d_term : nullptr;
  d.d_episode_r = h->d_episode_r ? e->d_ep_r : nullptr;
  d.d_episode_l = h->d_episode_l ? e->d_ep_l : nullptr;
  // Large batches are cut into chunks that alternate between two streams, so the H2D copy of chunk
  // c+1 and the kernel of chunk c run under the D2H copy of chunk c-1 (PCIe is full duplex; the D2H of
  // the observations, 60 of the 82 bytes per env-step, is the long pole).
  const int64_t kChunk = 1 << 20;
  const int n_chunks = n >= 2 * kChunk ? (int)((n + kChunk - 1) / kChunk) : 1;
  cudaStream_t streams[2] = {e->io_stream, e->io_stream2};
  for (int c = 0; c < n_chunks; ++c) {
    const int64_t first = (int64_t)c * kChunk, count = std::min<int64_t>(kChunk, n - first);
    const size_t cnt = (size_t)count;
    cudaStream_t st = streams[c & 1];
    CUDA_TRY(cudaMemcpyAsync(e->d_act + first * 4, h_actions + first * 4, cnt * 4 * sizeof(float), cudaMemcpyHostToDevice, st));
    rc = launch_step_range(e, e->d_act, d, first, count, st);
    if (rc) return rc;
#define COPY_OUT(field, stride, esize) if (h->field) CUDA_TRY(cudaMemcpyAsync(h->field + first * (stride), d.field + first * (stride), cnt * (stride) * (esize), cudaMemcpyDeviceToHost, st));
    COPY_OUT(d_obs, D, sizeof(float)) COPY_OUT(d_reward, 1, sizeof(float)) COPY_OUT(d_done, 1, 1)
    COPY_OUT(d_truncated, 1, 1) COPY_OUT(d_terminal_obs, D, sizeof(float))
    COPY_OUT(d_episode_r, 1, sizeof(float)) COPY_OUT(d_episode_l, 1, sizeof(int32_t))
#undef COPY_OUT
  }
  e->t += 1;
  e->env_steps += (uint64_t)n;
  CUDA_TRY(cudaStreamSynchronize(e->io_stream));
  if (n_chunks > 1) CUDA_TRY(cudaStreamSynchronize(e->io_stream2));
  return DRONECU_OK;
}

extern "C" int dronecu_reset_host(dronecu_env* e, const uint8_t* h_mask, float* h_obs) {
  if (!e) return fail(DRONECU_ERR_INVALID, "null handle");
  DeviceGuard guard(e->device);
  int rc = ensure_io(e);
  if (rc) return rc;
  const size_t n = (size_t)e->n, D = (size_t)e->cfg.obs_dim;
  cudaStream_t st = e->io_stream;
  if (h_mask) CUDA_TRY(cudaMemcpyAsync(e->d_mask, h_mask, n, cudaMemcpyHostToDevice, st));
  rc = launch_reset(e, h_mask ? e->d_mask : nullptr, h_obs ? e->d_obs : nullptr, 0, st);
  if (rc) return rc;
  if (h_obs) CUDA_TRY(cudaMemcpyAsync(h_obs, e->d_obs, n * D * sizeof(float), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return DRONECU_OK;
}

static StateView to_view(const dronecu_state_view* v) {
  StateView s;
  s.pos = v->d_pos; s.vel = v->d_vel; s.euler = v->d_euler; s.omega = v->d_omega; s.target = v->d_target;
  s.ep_ret = v->d_ep_ret; s.step = v->d_step; s.ep_num = v->d_ep_num; s.ep_len = v->d_ep_len;
  return s;
}

extern "C" int dronecu_get_state(dronecu_env* e, const dronecu_state_view* v, void* stream) {
  if (!e || !v) return fail(DRONECU_ERR_INVALID, "null argument");
  DeviceGuard guard(e->device);
  get_state_kernel<<<grid_for(e->n), kBlock, 0, (cudaStream_t)stream>>>(e->sp, e->n, to_view(v));
  e->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return DRONECU_OK;
}

extern "C" int dronecu_set_state(dronecu_env* e, const dronecu_state_view* v, void* stream) {
  if (!e || !v) return fail(DRONECU_ERR_INVALID, "null argument");
  DeviceGuard guard(e->device);
  set_state_kernel<<<grid_for(e->n), kBlock, 0, (cudaStream_t)stream>>>(e->sp, e->n, to_view(v));
  e->launches += 1;
  CUDA_TRY(cudaGetLastError());
  return DRONECU_OK;
}

// host views: stage through d_view_f ([5][n,3] + ep_ret[n]) and d_view_i ([3][n])
static dronecu_state_view device_scratch_view(dronecu_env* e, const dronecu_state_view* h) {
  const size_t n = (size_t)e->n;
  dronecu_state_view d;
  std::memset(&d, 0, sizeof(d));
  if (h->d_pos) d.d_pos = e->d_view_f;
  if (h->d_vel) d.d_vel = e->d_view_f + 3 * n;
  if (h->d_euler) d.d_euler = e->d_view_f + 6 * n;
  if (h->d_omega) d.d_omega = e->d_view_f + 9 * n;
  if (h->d_target) d.d_target = e->d_view_f + 12 * n;
  if (h->d_ep_ret) d.d_ep_ret = e->d_view_f + 15 * n;
  if (h->d_step) d.d_step = e->d_view_i;
  if (h->d_ep_num) d.d_ep_num = e->d_view_i + n;
  if (h->d_ep_len) d.d_ep_len = e->d_view_i + 2 * n;
  return d;
}

extern "C" int dronecu_get_state_host(dronecu_env* e, const dronecu_state_view* h) {
  if (!e || !h) return fail(DRONECU_ERR_INVALID, "null argument");
  DeviceGuard guard(e->device);
  int rc = ensure_io(e);
  if (rc) return rc;
  const size_t n = (size_t)e->n;
  cudaStream_t st = e->io_stream;
  dronecu_state_view d = device_scratch_view(e, h);
  rc = dronecu_get_state(e, &d, st);
  if (rc) return rc;
#define COPY_OUT(field, bytes) if (h->field) CUDA_TRY(cudaMemcpyAsync(h->field, d.field, bytes, cudaMemcpyDeviceToHost, st));
  COPY_OUT(d_pos, n * 12) COPY_OUT(d_vel, n * 12) COPY_OUT(d_euler, n * 12) COPY_OUT(d_omega, n * 12)
  COPY_OUT(d_target, n * 12) COPY_OUT(d_ep_ret, n * 4) COPY_OUT(d_step, n * 4) COPY_OUT(d_ep_num, n * 4) COPY_OUT(d_ep_len, n * 4)
#undef COPY_OUT
  CUDA_TRY(cudaStreamSynchronize(st));
  return DRONECU_OK;
}

extern "C" int dronecu_set_state_host(dronecu_env* e, const dronecu_state_view* h) {
  if (!e || !h) return fail(DRONECU_ERR_INVALID, "null argument");
  DeviceGuard guard(e->device);
  int rc = ensure_io(e);
  if (rc) return rc;
  const size_t n = (size_t)e->n;
  cudaStream_t st = e->io_stream;
  dronecu_state_view d = device_scratch_view(e, h);
#define COPY_IN(field, bytes) if (h->field) CUDA_TRY(cudaMemcpyAsync(d.field, h->field, bytes, cudaMemcpyHostToDevice, st));
  COPY_IN(d_pos, n * 12) COPY_IN(d_vel, n * 12) COPY_IN(d_euler, n * 12) COPY_IN(d_omega, n * 12)
  COPY_IN(d_target, n * 12) COPY_IN(d_ep_ret, n * 4) COPY_IN(d_step, n * 4) COPY_IN(d_ep_num, n * 4) COPY_IN(d_ep_len, n * 4)
#undef COPY_IN
  rc = dronecu_set_state(e, &d, st);
  if (rc) return rc;
  CUDA_TRY(cudaStreamSynchronize(st));
  return DRONECU_OK;
}

extern "C" int dronecu_episode_stats(dronecu_env* e, dronecu_stats* out, int reset) {
  if (!e || !out) return fail(DRONECU_ERR_INVALID, "null argument");
  DeviceGuard guard(e->device);
  CUDA_TRY(cudaDeviceSynchronize());
  StatSlot host[kStatSlots];
  CUDA_TRY(cudaMemcpy(host, e->stats, sizeof(host), cudaMemcpyDeviceToHost));
  std::memset(out, 0, sizeof(*out));
  for (int k = 0; k < kStatSlots; ++k) {
    out->episodes += host[k].episodes;
    out->terminated += host[k].terminated;
    out->length_sum += host[k].length_sum;
    out->return_sum += host[k].return_sum;
  }
  out->truncated = out->episodes - out->terminated;
  out->env_steps = e->env_steps;
  if (reset) {
    CUDA_TRY(cudaMemset(e->stats, 0, sizeof(StatSlot) * kStatSlots));
    e->env_steps = 0;
  }
  return DRONECU_OK;
}
