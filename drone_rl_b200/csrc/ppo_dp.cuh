// Data-parallel PPO update over NVLink peer memory: gradient exchange + clip_grad_norm_ + Adam in ONE kernel.
//
// The only exchange on the path is PPO's (reference call site train.py:63-68 -- the reference is
// single-process, so this has no counterpart there): per optimiser step the flat gradient + statistics
// (kGradLen = 10,705 float32 = 42.8 KB) must be summed over the ranks.  Round 1 did reduce -> torch.distributed
// all_reduce (NCCL) -> apply: three host-side operations per minibatch, not CUDA-graph capturable with the process
// group in between, and measured 0.37 ms per collective at 8 GPUs (18 ms per iteration) for 43 KB -- launch and
// straggler serialisation, not bandwidth.  Here every rank owns a MAILBOX in its own HBM that all peers map
// (cudaIpc* across processes, plain peer access inside one process):
//
//     float    data[2][world][kDpSlot]      slot [parity][r] is written only by rank r
//     uint32_t flag[2][kDpMaxWorld][8]      one 32-byte sector per flag; flag [parity][r] = sequence number of r's last push
//
// One exchange (sequence number s = ++seq, parity = s & 1), executed by ONE CTA per rank:
//   push   the local vector goes to slot [parity][rank] of EVERY rank's mailbox (posted 128-bit stores over NVLink;
//          own mailbox included), __threadfence_system, then st.release.sys of s into flag [parity][rank] everywhere;
//   wait   spin with ld.acquire.sys on the OWN mailbox's flags (local HBM / L2, no NVLink round trips) until all
//          `world` flags of this parity carry s (bounded by a timeout that raises a status instead of hanging the GPU);
//   sum    slots 0 .. world-1 in FIXED rank order -> every rank forms bit-identical sums, so the replicas of the
//          parameters never drift apart (asserted by the 2-rank test);
//   apply  clip_grad_norm_ + Adam on the summed gradient (the body of ppo_apply_kernel), scaled by 1 / (summed sample count).
// Two parities suffice: a rank can run at most one exchange ahead of the slowest rank (it cannot pass the wait of
// exchange s+1 before every peer has pushed s+1, i.e. has finished reading the slots of s).
// All of it is plain kernel launches on the caller's stream, so a whole epoch (grad -> reduce -> exchange+apply per
// minibatch) is one CUDA graph on every rank.
#pragma once
#include "ppo_update.cuh"

namespace dronecu {

constexpr int kDpMaxWorld = 16;
constexpr int kDpSlot = 10752;                    // floats per slot: kGradLen rounded up to a multiple of 32 (128-byte lines)
constexpr int kDpFlagStride = 8;                  // uint32 per flag: one 32-byte sector each
constexpr int kDpBlock = 1024;
static_assert(kDpSlot >= kGradLen && kDpSlot % 32 == 0, "slot size");

__host__ __device__ inline size_t dp_mailbox_bytes(int world) {
  return sizeof(float) * 2 * (size_t)world * kDpSlot + sizeof(uint32_t) * 2 * kDpMaxWorld * kDpFlagStride;
}

struct DpArgs {
  float* mail[kDpMaxWorld];          // mail[r]: rank r's mailbox as mapped in THIS process (mail[rank] = the local one)
  int rank, world;
  unsigned long long* seq;           // device-resident exchange counter of this rank (CUDA-graph capturable)
  int* status;                       // device: 0 = ok, 1 = an exchange timed out (results are garbage from there on)
  unsigned long long timeout_ns;
};

__device__ __forceinline__ float* dp_slot(float* mail, int world, int parity, int r) {
  return mail + ((size_t)parity * world + r) * kDpSlot;
}
__device__ __forceinline__ uint32_t* dp_flag(float* mail, int world, int parity, int r) {
  return reinterpret_cast<uint32_t*>(mail + 2 * (size_t)world * kDpSlot) + ((size_t)parity * kDpMaxWorld + r) * kDpFlagStride;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// push `n` floats (local, 16-byte aligned) into every rank's slot of this rank, publish, wait for all ranks.
// Whole CTA; returns with the peers' data visible to every thread of the CTA.
__device__ __forceinline__ void dp_exchange(const DpArgs& D, const float* __restrict__ src, const int n, const uint32_t s) {
  const int parity = (int)(s & 1u), tid = threadIdx.x, n4 = n >> 2;
  const float4* src4 = reinterpret_cast<const float4*>(src);
  for (int i = tid; i < n4; i += kDpBlock) {
    const float4 v = src4[i];
    for (int k = 0; k < D.world; ++k) {                       // start at the next rank: spreads the NVLink traffic
      int r = D.rank + 1 + k; if (r >= D.world) r -= D.world;
      reinterpret_cast<float4*>(dp_slot(D.mail[r], D.world, parity, D.rank))[i] = v;
    }
  }
  for (int i = 4 * n4 + tid; i < n; i += kDpBlock) {
    const float v = src[i];
    for (int r = 0; r < D.world; ++r) dp_slot(D.mail[r], D.world, parity, D.rank)[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (tid < D.world) st_release_sys(dp_flag(D.mail[tid], D.world, parity, D.rank), s);
  if (tid < D.world) {
    const uint32_t* f = dp_flag(D.mail[D.rank], D.world, parity, tid);
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(ld_acquire_sys(f) - s) < 0) {
      if (global_timer_ns() - t0 > D.timeout_ns) { atomicExch(D.status, 1); break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

// generic small all-reduce (sum) of float64 values in place: the per-epoch advantage statistics, episode statistics.
// n <= kDpSlot / 2.  One CTA.
__global__ void __launch_bounds__(kDpBlock) dp_allreduce_f64_kernel(const __grid_constant__ DpArgs D, double* __restrict__ buf, const int n) {
  const uint32_t s = (uint32_t)(*D.seq + 1ull);
  dp_exchange(D, reinterpret_cast<const float*>(buf), 2 * n, s);
  const int parity = (int)(s & 1u);
  for (int i = threadIdx.x; i < n; i += kDpBlock) {
    double acc = 0.0;
    for (int r = 0; r < D.world; ++r) acc += __ldcg(reinterpret_cast<const double*>(dp_slot(D.mail[D.rank], D.world, parity, r)) + i);
    buf[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) *D.seq = *D.seq + 1ull;
}

// exchange + clip_grad_norm_ + Adam.  A.grad: the LOCAL sum-form gradient (kGradLen floats: output of ppo_reduce_*), replaced
// by the sum over the ranks; A.inv_count is ignored: the denominator is the summed sample count grad[kParams + 4].
__global__ void __launch_bounds__(kDpBlock) ppo_apply_dp_kernel(const __grid_constant__ DpArgs D, const AdamArgs A, float* __restrict__ grad_io) {
  __shared__ float red[32];
  __shared__ float scal[3];
  __shared__ float stats_s[kGradLen - kParams];  // the summed statistics (sample count at [4])
  static_assert(kDpBlock == kApplyBlock, "one thread layout for both apply kernels");
  const uint32_t s = (uint32_t)(*D.seq + 1ull);
  dp_exchange(D, grad_io, kGradLen, s);
  const int parity = (int)(s & 1u);
  float g[kApplyPer];
#pragma unroll
  for (int k = 0; k < kApplyPer; ++k) {
    const int i = threadIdx.x + k * kDpBlock;
    float acc = 0.f;
    if (i < kGradLen) {
      acc = __ldcg(dp_slot(D.mail[D.rank], D.world, parity, 0) + i);
      for (int r = 1; r < D.world; ++r) acc += __ldcg(dp_slot(D.mail[D.rank], D.world, parity, r) + i);
      grad_io[i] = acc;
      if (i >= kParams) stats_s[i - kParams] = acc;
    }
    g[k] = (i < kParams) ? acc : 0.f;
  }
  __syncthreads();
  clip_adam_apply(A, g, stats_s, 1.0f / stats_s[4], red, scal);
  if (threadIdx.x == 0) *D.seq = *D.seq + 1ull;
}

}  // namespace dronecu
