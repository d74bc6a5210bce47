// PPO minibatch gradient on the 5th-generation tensor cores (tcgen05 + TMEM, tf32 operands, fp32
// accumulation): same contract as ppo_grad_kernel (ppo_update.cuh), i.e. what SB3's PPO.train() does
// per minibatch (call site /root/reference/train.py:63-68) -- "parity unpinned".
//
// Grid (x, 2): blockIdx.y selects the tower (0 = policy, 1 = value) -- the towers share nothing but the
// observation.  One CTA per SM, two warpgroups per CTA; a warpgroup owns one tile of 128 samples at a
// time (thread i <-> sample i <-> TMEM lane i).  All matrix products run as tcgen05.mma (kind::tf32),
// issued by the warpgroup's first thread:
//   D1  [128 x 64]  = X . W1^T       A = X in TMEM (K = 16: 15 observations + a ones column carrying b1)
//   D2  [128 x 64]  = H1 . W2^T      A = H1 = tanh(D1), written back to TMEM in place
//   dW3 [ 64 x  8] += H2^T . G       G = (dL/d head output (4) | 1 0 0 0) per sample
//   dH1 [128 x 64]  = dZ2 . W2       A = dZ2 in TMEM (over D2), B = a transposed copy of W2
//   dW2 [ 64 x 64] += dZ2^T . H1 ,   db2 [64 x 8] += dZ2^T . G   (column 4 = bias gradient)
//   dW1 [ 64 x 16] += dZ1^T . X      (column 15, the ones column, = db1)
// The weight-gradient products reduce over the SAMPLES, so their accumulators (M = 64: 16 TMEM lanes
// per subpartition) stay resident in TMEM for the whole life of the CTA and are read out once at the
// end.  TMEM per warpgroup: P 64 | Q 64 | dW2 64 | dW1 16 | dW3 8 | db2 8 = 224 of its 256 columns.
//
// Operand staging for the weight gradients (K = samples).  The "transposed" operands H2^T, dZ2^T, H1,
// dZ1^T are MN-major: a thread writes its sample's 64-float row once, as two 128-byte rows (32 features
// each), and the tensor core reads the buffer as [64 features] x [K = 128 samples].  For tf32 the only
// MN-major shared-memory layout the hardware accepts is SWIZZLE_128B_BASE32B (cute
// Layout_MN_SW128_32B_Atom: atoms of 4 K-rows x 128 bytes, 32-byte chunks XOR-ed with the row index;
// probed on the B200: the 16-byte-base 128B swizzle and the no-swizzle MN-major forms return zeros).
// The narrow operands X (16 wide) and G (8 wide) are scattered transposed into a K-major no-swizzle
// tile whose K-chunk stride is padded to 144 bytes so that the 32 lanes of a warp hit 32 banks.
// tanh is MUFU.TANH; tanh' = 1 - h^2 and the four-wide heads run on the CUDA cores.
#pragma once
#include "ppo_update.cuh"
#include "tc_mlp.cuh"

namespace dronecu {
namespace tcu {

using tc::fence_after;
using tc::fence_before;
using tc::mbar_init;
using tc::mbar_wait;
using tc::mma_commit;
using tc::mma_tf32_ts;
using tc::smem_addr;
using tc::tanh_mufu;
using tc::tmem_ld16;
using tc::tmem_st16;
using tc::to_tf32;
using tc::to_tf32_fast;
using tc::umma_off;
using tc::wait_st;

constexpr int kWG = 2;                       // warpgroups (= sample tiles in flight) per CTA
constexpr int kThreads = 128 * kWG;
constexpr int kHalfBytes = 128 * 128;        // one 32-feature half of a [128 samples x 64] fp32 operand buffer
constexpr int kColP = 0, kColQ = 64;         // P: D1 -> H1 (A of layer 2) -> dH1 ;  Q: X (A of layer 1) -> D2 -> dZ2 (A of dH1)
constexpr int kAccW2 = 128, kAccW1 = 192, kAccW3 = 208, kAccB2 = 216;
constexpr int kColsPerWG = 256;
constexpr int kTmemCols = kColsPerWG * kWG;  // 512
constexpr int kNG = 16;                      // per-warpgroup scalar accumulators (see G below)
// K-major no-swizzle tile of the narrow operands: rows 0..15 = X^T (x0..x14, 1), rows 16..23 = G^T
constexpr int kXaLbo = 144;                  // bytes between consecutive 4-sample K chunks (128 + 16 padding)
constexpr int kXaSbo = 32 * kXaLbo;          // bytes between 8-row groups (128 samples = 32 chunks)
constexpr int kXaBytes = 3 * kXaSbo;         // 13,824

struct alignas(1024) Smem {
  unsigned char bufA[kWG][2 * kHalfBytes];   // H1, later dZ1         (MN-major, SWIZZLE_128B_BASE32B)
  unsigned char bufB[kWG][2 * kHalfBytes];   // H2, later dZ2
  unsigned char XA[kWG][kXaBytes];
  float W1[kHid * 16];                       // this CTA's tower.  canonical no-swizzle K-major [out][k], k = 15 holds b1
  float W2[kHid * kHid];                     // canonical no-swizzle K-major [out][in]   (B of layer 2)
  float W2T[kHid * kHid];                    // canonical no-swizzle K-major [in][out]   (B of dH1 = dZ2 . W2)
  float b2[kHid];
  float W3T[kHid][kAct];                     // head weights [j][o]; the value tower uses o = 0 only
  float b3[kAct];
  float log_std[kAct];
  float wsum[kWG][4][12];                    // per-warp partial sums
  // pi: [0..3] d log_std | [4..7] db3 | [8] policy loss | [9] kl | [10] clip fraction ; vf: [0] db3 | [1] value loss | [2] count
  float G[kWG][kNG];
  alignas(8) unsigned long long mbar[kWG];
  uint32_t tmem_base, pad1[3];
};

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 @0, LBO >> 4 @16, SBO >> 4 @32,
// version 1 @46, layout type @61 (0 = no swizzle, 1 = 128-byte swizzle with 32-byte base)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
// MN-major [64 features x 8 samples] slice s of an activation buffer: MN atoms (32 features) 16 KB apart, K atoms (4 rows) 512 B apart
__device__ __forceinline__ uint64_t desc_mn(uint32_t buf, int s) { return make_desc(buf + 1024 * s, kHalfBytes, 512, 1); }
// K-major slice s (8 samples) of the narrow tile, starting at row group g (0: X rows 0..15, 2: G rows 16..23)
__device__ __forceinline__ uint64_t desc_xa(uint32_t xa, int g, int s) { return make_desc(xa + g * kXaSbo + 2 * kXaLbo * s, kXaLbo, kXaSbo, 0); }
// canonical no-swizzle K-major weight tile [64 x K]: slice s of 8 k
__device__ __forceinline__ uint64_t desc_w(uint32_t w, int K, int s) { return make_desc(w + 256 * s, 128, K * 32, 0); }

// instruction descriptor: D = f32, A = B = tf32, major-ness bits 15 / 16 (1 = MN-major), N >> 3 @17, M >> 4 @24
__host__ __device__ constexpr uint32_t idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t id, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(id), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void wg_barrier(int wg) { asm volatile("bar.sync %0, 128;" :: "r"(wg + 1) : "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// float4 number c4 (features 4 c4 .. 4 c4 + 3, c4 in 0..15) of sample row r in an MN-major activation buffer
__device__ __forceinline__ float4* mn_quad(unsigned char* buf, int r, int c4) {
  return reinterpret_cast<float4*>(buf + (c4 >> 3) * kHalfBytes + r * 128 + ((((c4 >> 1) & 3) ^ (r & 3)) << 5) + ((c4 & 1) << 4));
}
// element (row n, sample s) of the narrow K-major tile
__device__ __forceinline__ float* xa_elem(unsigned char* xa, int n, int s) {
  return reinterpret_cast<float*>(xa + (n >> 3) * kXaSbo + (s >> 2) * kXaLbo + ((n & 7) << 4) + ((s & 3) << 2));
}

// hand-over of the warpgroup's generic-proxy smem writes / TMEM accesses to its MMA-issuing thread
__device__ __forceinline__ void publish(int wg) {
  proxy_fence();
  fence_before();
  wg_barrier(wg);
}

__device__ __forceinline__ void setup(Smem& S, const float* __restrict__ theta, const int tw) {
  const int tid = threadIdx.x;
  const int oW1 = tw ? O_VF_W1 : O_PI_W1, oB1 = tw ? O_VF_B1 : O_PI_B1, oW2 = tw ? O_VF_W2 : O_PI_W2;
  for (int q = tid; q < kHid * 16; q += kThreads) {
    const int j = q / 16, k = q % 16;
    S.W1[umma_off(j, k, 16)] = to_tf32((k < kObs) ? theta[oW1 + j * kObs + k] : theta[oB1 + j]);
  }
  for (int q = tid; q < kHid * kHid; q += kThreads) {
    const int j = q / kHid, i = q % kHid;
    const float v = to_tf32(theta[oW2 + q]);
    S.W2[umma_off(j, i, kHid)] = v;
    S.W2T[umma_off(i, j, kHid)] = v;
  }
  for (int q = tid; q < kHid; q += kThreads) S.b2[q] = theta[(tw ? O_VF_B2 : O_PI_B2) + q];
  for (int q = tid; q < kAct * kHid; q += kThreads) {
    const int o = q / kHid, j = q % kHid;
    S.W3T[j][o] = tw ? (o == 0 ? theta[O_VF_W3 + j] : 0.f) : theta[O_PI_W3 + q];
  }
  if (tid < kAct) {
    S.b3[tid] = tw ? (tid == 0 ? theta[O_VF_B3] : 0.f) : theta[O_PI_B3 + tid];
    S.log_std[tid] = theta[O_LOGSTD + tid];
  }
  for (int q = tid; q < kWG * kNG; q += kThreads) S.G[q / kNG][q % kNG] = 0.f;
  if (tid == 0) {
    for (int w = 0; w < kWG; ++w) mbar_init(&S.mbar[w], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_addr(&S.tmem_base)), "r"((uint32_t)kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  proxy_fence();
  fence_before();
  __syncthreads();
  fence_after();
}

}  // namespace tcu

constexpr size_t kTcUpdSmem = sizeof(tcu::Smem) + 1024;     // + slack to align the dynamic segment to 1024 B

// partials: [2 towers][kWG * gridDim.x][kGradLen]; a CTA writes only its tower's entries (ppo_reduce_tc_kernel)
__global__ void __launch_bounds__(tcu::kThreads, 1) ppo_grad_tc_kernel(const __grid_constant__ UpdArgs A) {
  using namespace tcu;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  Smem& S = *reinterpret_cast<Smem*>(smem_dyn + ((1024u - (smem_addr(smem_dyn) & 1023u)) & 1023u));
  const int tid = threadIdx.x, wg = tid >> 7, r = tid & 127, lane = tid & 31, wq = r >> 5;
  const int tw = blockIdx.y;

  setup(S, A.theta, tw);

  float std_inv[kAct], logstd_sum = 0.f;
#pragma unroll
  for (int o = 0; o < kAct; ++o) { std_inv[o] = expf(-S.log_std[o]); logstd_sum += S.log_std[o]; }
  float adv_mean = A.adv_mean, adv_inv_std = A.adv_inv_std;
  if (A.adv_stats != nullptr) {        // SB3: (adv - adv.mean()) / (adv.std() + 1e-8), torch.std is unbiased
    const double cnt = A.adv_stats[2], mu = A.adv_stats[0] / cnt;
    const double var = (A.adv_stats[1] - A.adv_stats[0] * mu) / (cnt - 1.0);
    adv_mean = (float)mu;
    adv_inv_std = (float)(1.0 / (sqrt(fmax(var, 0.0)) + 1e-8));
  }

  unsigned char* const bufA = S.bufA[wg];
  unsigned char* const bufB = S.bufB[wg];
  unsigned char* const XA = S.XA[wg];
  unsigned long long* const bar = &S.mbar[wg];
  const uint32_t aA = smem_addr(bufA), aB = smem_addr(bufB), aX = smem_addr(XA);
  const uint32_t aW1 = smem_addr(S.W1), aW2 = smem_addr(S.W2), aW2T = smem_addr(S.W2T);
  const uint32_t tmem = S.tmem_base + wg * kColsPerWG;                   // lane 0, first column of the warpgroup
  const uint32_t tL = tmem + ((uint32_t)(wq * 32) << 16);                // this thread's subpartition
  uint32_t phase = 0;

  const int64_t n_tiles = (A.m + 127) / 128;
  int it = 0;
  for (int64_t tile = (int64_t)blockIdx.x * kWG + wg; tile < n_tiles; tile += (int64_t)gridDim.x * kWG, ++it) {
    const int64_t pos = tile * 128 + r;
    const bool live = pos < A.m;
    const int64_t row = live ? (A.index ? (int64_t)A.index[pos] : A.first + pos) : 0;
    float x[16];
#pragma unroll
    for (int i = 0; i < kObs; ++i) x[i] = live ? to_tf32_fast(A.obs[row * kObs + i]) : 0.f;
    x[15] = live ? 1.0f : 0.f;
    float4 act = make_float4(0.f, 0.f, 0.f, 0.f);
    float old_logp = 0.f, adv_raw = 0.f, ret = 0.f;
    if (live) {
      if (tw == 0) { act = A.actions[row]; old_logp = A.old_logp[row]; adv_raw = A.adv[row]; }
      else ret = A.ret[row];
    }
    const bool first = (it == 0);

    if (!first) { mbar_wait(bar, phase); phase ^= 1; fence_after(); }    // the previous tile's dW1 has read bufA / XA
    tmem_st16(tL + kColQ, x);
#pragma unroll
    for (int i = 0; i < 16; ++i) *xa_elem(XA, i, r) = x[i];
    wait_st();
    // ---------------- layer 1 ----------------
    publish(wg);
    if (r == 0) {
      fence_after();
#pragma unroll
      for (int s = 0; s < 2; ++s) mma_tf32_ts(tmem + kColP, tmem + kColQ + 8 * s, desc_w(aW1, 16, s), idesc(128, 64, 0, 0), s > 0);
      mma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    fence_after();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v[16];
      tmem_ld16(tL + kColP + 16 * c, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = to_tf32_fast(tanh_mufu(v[i]));
      tmem_st16(tL + kColP + 16 * c, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) *mn_quad(bufA, r, 4 * c + q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    wait_st();
    // ---------------- layer 2 + head ----------------
    publish(wg);
    if (r == 0) {
      fence_after();
#pragma unroll
      for (int s = 0; s < 8; ++s) mma_tf32_ts(tmem + kColQ, tmem + kColP + 8 * s, desc_w(aW2, kHid, s), idesc(128, 64, 0, 0), s > 0);
      mma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    fence_after();
    float out[kAct];
#pragma unroll
    for (int o = 0; o < kAct; ++o) out[o] = S.b3[o];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v[16];
      tmem_ld16(tL + kColQ + 16 * c, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 b = reinterpret_cast<const float4*>(S.b2 + 16 * c)[q];
        v[4 * q] += b.x; v[4 * q + 1] += b.y; v[4 * q + 2] += b.z; v[4 * q + 3] += b.w;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = tanh_mufu(v[i]);
        if (tw == 0) {
          const float4 w = reinterpret_cast<const float4*>(S.W3T[16 * c + i])[0];
          out[0] = fmaf(w.x, a, out[0]); out[1] = fmaf(w.y, a, out[1]);
          out[2] = fmaf(w.z, a, out[2]); out[3] = fmaf(w.w, a, out[3]);
        } else {
          out[0] = fmaf(S.W3T[16 * c + i][0], a, out[0]);
        }
        v[i] = to_tf32_fast(a);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) *mn_quad(bufB, r, 4 * c + q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    // ---------------- loss gradient at the head ----------------
    float g3[kAct] = {0.f, 0.f, 0.f, 0.f};
    if (tw == 0) {
      float g_ls[kAct] = {0.f, 0.f, 0.f, 0.f}, st_pl = 0.f, st_kl = 0.f, st_cf = 0.f;
      if (live) {
        const float av[4] = {act.x, act.y, act.z, act.w};
        float z[kAct], sq = 0.f;
#pragma unroll
        for (int o = 0; o < kAct; ++o) { z[o] = (av[o] - out[o]) * std_inv[o]; sq = fmaf(z[o], z[o], sq); }
        const float logp = -0.5f * sq - logstd_sum - kAct * kHalfLog2Pi;
        const float log_ratio = logp - old_logp;
        const float ratio = expf(log_ratio);
        const float adv = (adv_raw - adv_mean) * adv_inv_std;
        const float lo = 1.0f - A.clip, hi = 1.0f + A.clip;
        const float s1 = adv * ratio, s2 = adv * fminf(fmaxf(ratio, lo), hi);
        const bool inside = (ratio >= lo) && (ratio <= hi);
        const float dl_dlogp = (inside || s1 < s2) ? -adv * ratio : 0.f;     // d(-min(s1,s2)) / d logp
#pragma unroll
        for (int o = 0; o < kAct; ++o) {
          g3[o] = dl_dlogp * z[o] * std_inv[o];
          g_ls[o] = dl_dlogp * (z[o] * z[o] - 1.0f) - A.ent_coef;
        }
        st_pl = -fminf(s1, s2);
        st_kl = (ratio - 1.0f) - log_ratio;
        st_cf = (fabsf(ratio - 1.0f) > A.clip) ? 1.0f : 0.f;
      }
      float v[11] = {g_ls[0], g_ls[1], g_ls[2], g_ls[3], g3[0], g3[1], g3[2], g3[3], st_pl, st_kl, st_cf};
#pragma unroll
      for (int q = 0; q < 11; ++q) v[q] = warp_sum(v[q]);
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 11; ++q) S.wsum[wg][wq][q] = v[q];
      }
    } else {
      float st_vl = 0.f;
      if (live) {
        const float diff = out[0] - ret;
        g3[0] = 2.0f * A.vf_coef * diff;                   // d(vf_coef * (ret - v)^2) / dv
        st_vl = diff * diff;
      }
      const float v0 = warp_sum(g3[0]), v1 = warp_sum(st_vl), v2 = warp_sum(live ? 1.0f : 0.f);
      if (lane == 0) { S.wsum[wg][wq][0] = v0; S.wsum[wg][wq][1] = v1; S.wsum[wg][wq][2] = v2; }
    }
#pragma unroll
    for (int o = 0; o < kAct; ++o) *xa_elem(XA, 16 + o, r) = to_tf32_fast(g3[o]);
    *xa_elem(XA, 20, r) = live ? 1.0f : 0.f;
#pragma unroll
    for (int o = 21; o < 24; ++o) *xa_elem(XA, o, r) = 0.f;
    // ---------------- dW3 += H2^T . G ----------------
    publish(wg);
    if (r == 0) {
      fence_after();
#pragma unroll
      for (int s = 0; s < 16; ++s) mma_ss(tmem + kAccW3, desc_mn(aB, s), desc_xa(aX, 2, s), idesc(64, 8, 1, 0), (!first) || s > 0);
      mma_commit(bar);
    }
    if (r < (tw == 0 ? 11 : 3))       // scalar accumulators, fixed order over the four warps (deterministic)
      S.G[wg][r] += ((S.wsum[wg][0][r] + S.wsum[wg][1][r]) + S.wsum[wg][2][r]) + S.wsum[wg][3][r];
    mbar_wait(bar, phase); phase ^= 1;
    fence_after();
    // ---------------- dZ2 = (g3 . W3) * (1 - H2^2): over H2 in shared memory, and into TMEM (A of dH1) ----------------
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v[16];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4* p = mn_quad(bufB, r, 4 * c + q);
        const float4 h = *p;
        const float h4[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = 16 * c + 4 * q + e;
          float up;
          if (tw == 0) {
            const float4 w = reinterpret_cast<const float4*>(S.W3T[j])[0];
            up = g3[0] * w.x + g3[1] * w.y + g3[2] * w.z + g3[3] * w.w;
          } else {
            up = g3[0] * S.W3T[j][0];
          }
          v[4 * q + e] = to_tf32_fast(up * (1.0f - h4[e] * h4[e]));
        }
        *p = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
      tmem_st16(tL + kColQ + 16 * c, v);
    }
    wait_st();
    // ---------------- dH1 = dZ2 . W2 ; dW2 += dZ2^T . H1 ; db2 += dZ2^T . 1 ----------------
    publish(wg);
    if (r == 0) {
      fence_after();
#pragma unroll
      for (int s = 0; s < 8; ++s) mma_tf32_ts(tmem + kColP, tmem + kColQ + 8 * s, desc_w(aW2T, kHid, s), idesc(128, 64, 0, 0), s > 0);
#pragma unroll
      for (int s = 0; s < 16; ++s) mma_ss(tmem + kAccW2, desc_mn(aB, s), desc_mn(aA, s), idesc(64, 64, 1, 1), (!first) || s > 0);
#pragma unroll
      for (int s = 0; s < 16; ++s) mma_ss(tmem + kAccB2, desc_mn(aB, s), desc_xa(aX, 2, s), idesc(64, 8, 1, 0), (!first) || s > 0);
      mma_commit(bar);
    }
    mbar_wait(bar, phase); phase ^= 1;
    fence_after();
    // ---------------- dZ1 = dH1 * (1 - H1^2), over H1 ----------------
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v[16];
      tmem_ld16(tL + kColP + 16 * c, v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4* p = mn_quad(bufA, r, 4 * c + q);
        const float4 h = *p;
        *p = make_float4(to_tf32_fast(v[4 * q] * (1.0f - h.x * h.x)), to_tf32_fast(v[4 * q + 1] * (1.0f - h.y * h.y)),
                         to_tf32_fast(v[4 * q + 2] * (1.0f - h.z * h.z)), to_tf32_fast(v[4 * q + 3] * (1.0f - h.w * h.w)));
      }
    }
    // ---------------- dW1 (+ db1 in column 15) += dZ1^T . X ----------------
    publish(wg);
    if (r == 0) {
      fence_after();
#pragma unroll
      for (int s = 0; s < 16; ++s) mma_ss(tmem + kAccW1, desc_mn(aA, s), desc_xa(aX, 0, s), idesc(64, 16, 1, 0), (!first) || s > 0);
      mma_commit(bar);                        // waited for at the top of the next tile / after the loop
    }
  }
  if (it > 0) { mbar_wait(bar, phase); phase ^= 1; }
  fence_after();

  if (A.dbg != nullptr && tw == 0) {       // debugging aid: this thread's TMEM lane, all 256 columns of the warpgroup
    float* d = A.dbg + (((size_t)blockIdx.x * kWG + wg) * 128 + r) * kColsPerWG;
#pragma unroll 1
    for (int c = 0; c < kColsPerWG / 16; ++c) {
      float v[16];
      tmem_ld16(tL + 16 * c, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) d[16 * c + i] = v[i];
    }
  }

  // ---------------- read the accumulators out: one partial vector per (tower, CTA, warpgroup) ----------------
  float* const outv = A.partials + (((size_t)tw * gridDim.x + blockIdx.x) * kWG + wg) * kGradLen;
  const int oW1 = tw ? O_VF_W1 : O_PI_W1, oB1 = tw ? O_VF_B1 : O_PI_B1, oW2 = tw ? O_VF_W2 : O_PI_W2;
  const int oB2 = tw ? O_VF_B2 : O_PI_B2, oW3 = tw ? O_VF_W3 : O_PI_W3, oB3 = tw ? O_VF_B3 : O_PI_B3;
  const int nB3 = tw ? 1 : kAct;
  if (it == 0) {         // this warpgroup saw no tile: contribute zeros
    for (int idx = r; idx < oB3 + nB3 - oW1; idx += 128) outv[oW1 + idx] = 0.f;
    if (tw == 0 && r < kAct) outv[O_LOGSTD + r] = 0.f;
    if (r < kStats) outv[kParams + r] = 0.f;
  } else {
    const int j = 16 * wq + lane;               // accumulator row held by lanes 0..15 of each subpartition
    const bool own = lane < 16;
    float v[16];
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      tmem_ld16(tL + kAccW2 + 16 * c, v);
      if (own) {
#pragma unroll
        for (int i = 0; i < 16; ++i) outv[oW2 + j * kHid + 16 * c + i] = v[i];
      }
    }
    tmem_ld16(tL + kAccW1, v);
    if (own) {
#pragma unroll
      for (int i = 0; i < kObs; ++i) outv[oW1 + j * kObs + i] = v[i];
      outv[oB1 + j] = v[15];
    }
    tmem_ld16(tL + kAccW3, v);                 // columns 208..223: dW3 (8) | db2 (8)
    if (own) {
      if (tw == 0) {
#pragma unroll
        for (int o = 0; o < kAct; ++o) outv[oW3 + o * kHid + j] = v[o];
      } else {
        outv[oW3 + j] = v[0];
      }
      outv[oB2 + j] = v[8 + 4];
    }
    if (tw == 0) {
      if (r < kAct) { outv[O_LOGSTD + r] = S.G[wg][r]; outv[O_PI_B3 + r] = S.G[wg][4 + r]; }
      // statistics block: [policy loss, value loss, kl, clip fraction, count, 0, 0, 0]; the policy tower owns 0, 2, 3
      if (r < kStats) outv[kParams + r] = (r == 0) ? S.G[wg][8] : (r == 2) ? S.G[wg][9] : (r == 3) ? S.G[wg][10] : 0.f;
    } else {
      if (r == 0) outv[O_VF_B3] = S.G[wg][0];
      if (r < kStats) outv[kParams + r] = (r == 1) ? S.G[wg][1] : (r == 4) ? S.G[wg][2] : 0.f;
    }
  }

  fence_before();
  __syncthreads();
  if (tid < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(S.tmem_base), "r"((uint32_t)tcu::kTmemCols) : "memory");
}

// fixed-order sum of the partial vectors of ppo_grad_tc_kernel: element idx belongs to one tower (the value tower
// owns its parameter block and the statistics 1 (value loss) and 4 (count)), whose np partials are summed
__global__ void ppo_reduce_tc_kernel(const float* __restrict__ partials, int np, float* __restrict__ grad) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kGradLen) return;
  int tw = (idx >= O_VF_W1 && idx < O_LOGSTD) ? 1 : 0;
  if (idx >= kParams) { const int q = idx - kParams; tw = (q == 1 || q == 4) ? 1 : 0; }
  const float* p = partials + (size_t)tw * np * kGradLen + idx;
  float sum = 0.f;
  for (int k = 0; k < np; ++k) sum += p[(size_t)k * kGradLen];
  grad[idx] = sum;
}

}  // namespace dronecu
