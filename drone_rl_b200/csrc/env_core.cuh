// Per-env device logic of the quadcopter environment: one Euler step of the rigid body,
// reward, termination, observation, reset with curriculum target.
//
// WHAT it computes is fixed by /root/reference/drone.py:48-159 (DroneEnv) and
// /root/reference/vectorized_drone.py:38-216 (VectorizedDroneEnv); SURVEY.md appendix B has
// the maths in one place.  HOW is ours: float32 state held in registers, only the third
// column of the rotation matrix is formed (the reference multiplies the full matrix by
// [0,0,T]), the three sincos are shared between the translational and the Euler-rate update
// (the reference evaluates sin/cos(roll) three times), tan/sec come from one reciprocal.
// Terms that are multiplied by an exact zero in the reference are kept where they decide
// inf/NaN propagation (0*inf = NaN in T(roll,pitch) @ omega and in (Ixx-Iyy)*p*q).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "philox.cuh"

namespace dronecu {

// Device copy of dronecu_config, pre-digested on the host in float64 and rounded once.
struct EnvParams {
  float dt, inv_mass, gravity, lever, k_yaw;
  float dI_roll, dI_pitch, dI_yaw;  // (Iyy-Izz), (Izz-Ixx), (Ixx-Iyy)   drone.py:136-138
  float invI[3];
  float reward_scale, bonus_radius, bonus, z_floor, r_max;
  float fixed_target[3], fixed_start[3];
  float start_z, target_z, motor_max;
  double curriculum_step;
  int32_t curriculum_period, max_steps;
  float inv_curriculum_period, r_max_sq;
  uint64_t seed, env_offset;
  PhiloxKeys keys;   // round keys of `seed`
};

struct EnvState {
  float px, py, pz, vx, vy, vz;
  float roll, pitch, yaw, wp, wq, wr;
  float tx, ty, tz;
  int32_t step;     // DroneEnv.current_step (drone.py:66,155)
  int32_t ep_num;   // DroneEnv.ep_num       (drone.py:61)
  int32_t ep_len;   // VecMonitor episode length
  float ep_ret;     // VecMonitor episode return (float32, as SB3 keeps it)
};

// HBM layout: five planes of 16-byte quads, one quad per env per plane, so every access of a
// warp is one fully coalesced 512-byte 128-bit transaction:
//   plane 0: px py pz vx | plane 1: vy vz roll pitch | plane 2: yaw p q r
//   plane 3: tx ty tz step(bits) | plane 4: ep_num(bits) ep_len(bits) ep_ret 0
struct StatePlanes {
  float4* q[5];
};

__device__ __forceinline__ float4 ld_quad(const float4* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_quad_nc(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_quad(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ EnvState load_state(const StatePlanes& sp, int64_t i) {
  const float4 a = ld_quad(sp.q[0] + i), b = ld_quad(sp.q[1] + i), c = ld_quad(sp.q[2] + i);
  const float4 d = ld_quad(sp.q[3] + i), e = ld_quad(sp.q[4] + i);
  EnvState s;
  s.px = a.x; s.py = a.y; s.pz = a.z; s.vx = a.w;
  s.vy = b.x; s.vz = b.y; s.roll = b.z; s.pitch = b.w;
  s.yaw = c.x; s.wp = c.y; s.wq = c.z; s.wr = c.w;
  s.tx = d.x; s.ty = d.y; s.tz = d.z; s.step = __float_as_int(d.w);
  s.ep_num = __float_as_int(e.x); s.ep_len = __float_as_int(e.y); s.ep_ret = e.z;
  return s;
}

__device__ __forceinline__ void store_state(const StatePlanes& sp, int64_t i, const EnvState& s) {
  st_quad(sp.q[0] + i, make_float4(s.px, s.py, s.pz, s.vx));
  st_quad(sp.q[1] + i, make_float4(s.vy, s.vz, s.roll, s.pitch));
  st_quad(sp.q[2] + i, make_float4(s.yaw, s.wp, s.wq, s.wr));
  st_quad(sp.q[3] + i, make_float4(s.tx, s.ty, s.tz, __int_as_float(s.step)));
  st_quad(sp.q[4] + i, make_float4(__int_as_float(s.ep_num), __int_as_float(s.ep_len), s.ep_ret, 0.f));
}

// ---------------------------------------------------------------------------------------------
// reset: drone.py:48-75 (RANDOMIZED) / vectorized_drone.py:38-57 (fixed start, fixed target)
// ---------------------------------------------------------------------------------------------
template <bool RANDOMIZED>
__device__ __forceinline__ void reset_env(EnvState& s, const EnvParams& P, uint64_t env_id) {
  s.vx = s.vy = s.vz = 0.f;
  s.roll = s.pitch = s.yaw = 0.f;
  s.wp = s.wq = s.wr = 0.f;
  s.ep_num += 1;          // drone.py:61
  s.step = 0;             // drone.py:66 / vectorized_drone.py:56
  s.ep_len = 0;
  s.ep_ret = 0.f;
  if constexpr (RANDOMIZED) {
    // exactly five uniforms, in the reference's order pos.x pos.y tgt.x tgt.y tgt.z (drone.py:57,73),
    // out of ONE Philox call: 4 x 24 high bits + the 3 low bytes of words 0..2 (oracle/philox.py)
    const uint4 w = env_stream(P.keys, env_id, (uint64_t)(uint32_t)s.ep_num, STREAM_RESET);
    s.px = u01(w.x) - 0.5f;   // exact: u is a multiple of 2^-24
    s.py = u01(w.y) - 0.5f;
    s.pz = P.start_z;
    // eps: the reference ACCUMULATES eps += 0.1 in float64 every curriculum_period episodes
    // (drone.py:68-70) -> 0.1, 0.2, 0.30000000000000004 ...; replay the accumulation.
    // bumps = ep_num / period without an integer division: float estimate, exact fix-up
    // (|estimate - true| < 1 for every 31-bit ep_num)
    int bumps = 0;
    if (s.ep_num >= P.curriculum_period) {
      bumps = __float2int_rd((float)s.ep_num * P.inv_curriculum_period);
      const int r = s.ep_num - bumps * P.curriculum_period;
      bumps += (r >= P.curriculum_period) ? 1 : ((r < 0) ? -1 : 0);
    }
    if (bumps == 0) {         // eps == 0: target = [0, 0, target_z] exactly
      s.tx = 0.f; s.ty = 0.f; s.tz = P.target_z;
    } else {
      double eps = 0.0;
      for (int k = 0; k < bumps; ++k) eps += P.curriculum_step;
      const uint32_t lo = (w.x & 255u) | ((w.y & 255u) << 8) | ((w.z & 255u) << 16);
      s.tx = (float)(eps * (double)u01(w.z));
      s.ty = (float)(eps * (double)u01(w.w));
      s.tz = (float)(eps * (double)((float)lo * 5.9604644775390625e-8f) + (double)P.target_z);
    }
  } else {
    s.px = P.fixed_start[0]; s.py = P.fixed_start[1]; s.pz = P.fixed_start[2];
    s.tx = P.fixed_target[0]; s.ty = P.fixed_target[1]; s.tz = P.fixed_target[2];
  }
}

// ---------------------------------------------------------------------------------------------
// one physics step + reward + termination flags.  drone.py:101-157 / vectorized_drone.py:151-213
// ---------------------------------------------------------------------------------------------
// sin and cos of the three Euler angles.  One range check for all three (the reference never wraps its
// angles, so huge arguments must stay exact: they take libm's Payne-Hanek path, as do inf -> NaN); the common
// case is a three-term Cody-Waite reduction by pi/2 (valid below 105615, as in CUDA's own sincosf) with the
// quadrant taken from the low mantissa bits of a magic-number rounding, and the Cephes single-precision
// minimax polynomials on [-pi/4, pi/4] (< 1.5 ulp).  NaN arguments fall through the check (fmaxf drops
// NaN) and propagate through the arithmetic.
#ifndef DRONECU_SINCOS_COMPACT
#define DRONECU_SINCOS_COMPACT 0      // 1: compact the huge (lane, angle) pairs over the warp (measured: c2 -17 %, c4 -5 %; profiles/README.md)
#endif
__device__ __forceinline__ void sincos_reduced(float x, float& s, float& c) {
  const float j = fmaf(x, 0.636619772367581343f, 12582912.0f);      // 1.5 * 2^23: round to nearest integer
  const uint32_t q = __float_as_uint(j);
  const float n = j - 12582912.0f;
  float r = fmaf(n, -1.5707962512969970703f, x);
  r = fmaf(n, -7.5497894158615963534e-08f, r);
  r = fmaf(n, -5.3903029534742383927e-15f, r);
  const float r2 = r * r;
  float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(ps, r2, -1.6666654611e-1f);
  const float sn = fmaf(ps, r2 * r, r);
  float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(pc, r2, 4.166664568298827e-2f);
  pc = fmaf(pc, r2, -0.5f);
  const float cs = fmaf(pc, r2, 1.0f);
  const bool odd = q & 1u;
  const float s0 = odd ? cs : sn, c0 = odd ? sn : cs;
  s = __uint_as_float(__float_as_uint(s0) ^ ((q << 30) & 0x80000000u));          // quadrants 2, 3: sin < 0
  c = __uint_as_float(__float_as_uint(c0) ^ (((q + 1u) << 30) & 0x80000000u));   // quadrants 1, 2: cos < 0
}

// EXPERIMENT, off (DRONECU_SINCOS_COMPACT): measured slower -- in configs[1] the drones of a warp tumble together, so most lanes
// hold huge angles at the same time (nothing to compact), and the out-of-line call costs the common path registers.
// The rare path of sincos3: at least one lane of the warp holds an angle that is huge (>= 105615), inf or NaN.  Only THOSE
// angles need libm's Payne-Hanek reduction (~200 instructions each, executed by the whole warp whoever needs it): the
// (lane, angle) pairs that need it are compacted over the warp -- one pair per lane -- so that one sincosf call serves up to 32
// of them, instead of every affected lane walking through three calls one after the other.  The reference's vectorized env
// never resets (vectorized_drone.py:135-216): by step 1000 of configs[1] most warps hold a drone that has spun that far, and the
// three serial calls were ~60 % of all executed instructions there (profiles/r02_c2_chain_model.md).
// Full warps only (every kernel calls step_env with all 32 lanes converged); results are those of sincosf on the same argument.
static __device__ __noinline__ void sincos3_huge(const float a, const float b, const float g, const bool ha, const bool hb, const bool hg,
                                          float& sa, float& ca, float& sb, float& cb, float& sg, float& cg) {
  constexpr unsigned kFull = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u, below = (1u << lane) - 1u;
  const unsigned ma = __ballot_sync(kFull, ha), mb = __ballot_sync(kFull, hb), mg = __ballot_sync(kFull, hg);
  const int na = __popc(ma), nb = __popc(mb), total = na + nb + __popc(mg);
  const int ia = __popc(ma & below), ib = na + __popc(mb & below), ig = na + nb + __popc(mg & below);   // this lane's items
  for (int base = 0; base < total; base += 32) {
    const int item = base + (int)lane;
    const bool mine = item < total;
    const int which = item < na ? 0 : (item < na + nb ? 1 : 2);
    const unsigned msk = which == 0 ? ma : (which == 1 ? mb : mg);
    const int nth = item - (which == 0 ? 0 : (which == 1 ? na : na + nb));
    const int src = mine ? (int)__fns(msk, 0u, nth + 1) : 0;
    const float xa = __shfl_sync(kFull, a, src), xb = __shfl_sync(kFull, b, src), xg = __shfl_sync(kFull, g, src);
    float sx = 0.f, cx = 0.f;
    if (mine) sincosf(which == 0 ? xa : (which == 1 ? xb : xg), &sx, &cx);
    const int pa = ia - base, pb = ib - base, pg = ig - base;
    const float s_a = __shfl_sync(kFull, sx, pa & 31), c_a = __shfl_sync(kFull, cx, pa & 31);
    const float s_b = __shfl_sync(kFull, sx, pb & 31), c_b = __shfl_sync(kFull, cx, pb & 31);
    const float s_g = __shfl_sync(kFull, sx, pg & 31), c_g = __shfl_sync(kFull, cx, pg & 31);
    if (ha && pa >= 0 && pa < 32) { sa = s_a; ca = c_a; }
    if (hb && pb >= 0 && pb < 32) { sb = s_b; cb = c_b; }
    if (hg && pg >= 0 && pg < 32) { sg = s_g; cg = c_g; }
  }
}

__device__ __forceinline__ void sincos3(float a, float b, float g, float& sa, float& ca, float& sb, float& cb,
                                        float& sg, float& cg) {
  constexpr float kLim = 105615.0f;
  const bool small = fmaxf(fmaxf(fabsf(a), fabsf(b)), fabsf(g)) < kLim;       // false for huge, inf and NaN arguments
#if DRONECU_SINCOS_COMPACT
  const unsigned am = __activemask();
  if (am == 0xffffffffu) {
    sincos_reduced(a, sa, ca);
    sincos_reduced(b, sb, cb);
    sincos_reduced(g, sg, cg);
    if (__all_sync(0xffffffffu, small)) return;
    sincos3_huge(a, b, g, !(fabsf(a) < kLim), !(fabsf(b) < kLim), !(fabsf(g) < kLim), sa, ca, sb, cb, sg, cg);
    return;
  }
#endif
  if (small) {
    sincos_reduced(a, sa, ca);
    sincos_reduced(b, sb, cb);
    sincos_reduced(g, sg, cg);
  } else {
    sincosf(a, &sa, &ca);
    sincosf(b, &sb, &cb);
    sincosf(g, &sg, &cg);
  }
}

__device__ __forceinline__ float rcp_approx(float x) {     // 1 ulp; 0 -> inf, inf -> 0, NaN -> NaN like 1.0f / x
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {    // 1 ulp; 0 -> 0, inf -> inf, NaN -> NaN like sqrtf
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct StepResult {
  float reward;
  bool crashed;   // z < z_floor or |pos| > r_max          (drone.py:154)
  bool timeout;   // current_step >= max_steps             (drone.py:156)
};

__device__ __forceinline__ StepResult step_env(EnvState& s, const EnvParams& P, const float4 f) {
  // rotor mixer (drone.py:106-117); np.sum adds left to right
  const float thrust = ((f.x + f.y) + f.z) + f.w;
  const float tau_roll = P.lever * (((f.x + f.y) - f.z) - f.w);
  const float tau_pitch = P.lever * (((-f.x + f.y) + f.z) - f.w);
  const float tau_yaw = P.k_yaw * (((f.x - f.y) + f.z) - f.w);

  float sr, cr, sp, cp, sy, cy;
  sincos3(s.roll, s.pitch, s.yaw, sr, cr, sp, cp, sy, cy);

  // third column of R = Rz Ry Rx (drone.py:170-172) times thrust / mass, plus gravity (drone.py:124)
  const float tm = thrust * P.inv_mass;
  const float ax = (cy * sp * cr + sy * sr) * tm;
  const float ay = (sy * sp * cr - cy * sr) * tm;
  const float az = -P.gravity + (cp * cr) * tm;

  // Euler rates from the OLD angles and OLD body rates (drone.py:131, :181-186)
  const float sec_p = rcp_approx(cp);
  const float tan_p = sp * sec_p;
  const float roll_dot = s.wp + (sr * tan_p) * s.wq + (cr * tan_p) * s.wr;
  const float pitch_dot = 0.0f * s.wp + cr * s.wq + (-sr) * s.wr;
  const float yaw_dot = 0.0f * s.wp + (sr * sec_p) * s.wq + (cr * sec_p) * s.wr;

  // body-rate dynamics from the OLD rates (drone.py:135-138)
  const float wp_dot = (tau_roll - (P.dI_roll * s.wq) * s.wr) * P.invI[0];
  const float wq_dot = (tau_pitch - (P.dI_pitch * s.wp) * s.wr) * P.invI[1];
  const float wr_dot = (tau_yaw - (P.dI_yaw * s.wp) * s.wq) * P.invI[2];

  // semi-implicit Euler for the translation (drone.py:127-128), explicit for the rest
  s.vx += ax * P.dt; s.vy += ay * P.dt; s.vz += az * P.dt;
  s.px += s.vx * P.dt; s.py += s.vy * P.dt; s.pz += s.vz * P.dt;
  s.roll += roll_dot * P.dt; s.pitch += pitch_dot * P.dt; s.yaw += yaw_dot * P.dt;
  s.wp += wp_dot * P.dt; s.wq += wq_dot * P.dt; s.wr += wr_dot * P.dt;

  // reward (drone.py:142-148) and termination (drone.py:154-157); NaN compares false
  const float dx = s.px - s.tx, dy = s.py - s.ty, dz = s.pz - s.tz;
  const float dist = sqrt_approx(dx * dx + dy * dy + dz * dz);
  StepResult r;
  r.reward = -P.reward_scale * dist;
  if (dist < P.bonus_radius) r.reward += P.bonus;
  // |pos| > r_max tested on the squares (sqrt is monotonic; NaN still compares false)
  const float rad_sq = s.px * s.px + s.py * s.py + s.pz * s.pz;
  r.crashed = (s.pz < P.z_floor) || (rad_sq > P.r_max_sq);
  s.step += 1;
  r.timeout = s.step >= P.max_steps;
  s.ep_ret += r.reward;   // VecMonitor
  s.ep_len += 1;
  return r;
}

// observation: drone.py:77-79 (15: ..., target - pos) / vectorized_drone.py:59-61 (12)
template <int OBS_DIM>
__device__ __forceinline__ void write_obs(float* o, const EnvState& s) {
  o[0] = s.px; o[1] = s.py; o[2] = s.pz; o[3] = s.vx; o[4] = s.vy; o[5] = s.vz;
  o[6] = s.roll; o[7] = s.pitch; o[8] = s.yaw; o[9] = s.wp; o[10] = s.wq; o[11] = s.wr;
  if constexpr (OBS_DIM == 15) {
    o[12] = s.tx - s.px; o[13] = s.ty - s.py; o[14] = s.tz - s.pz;
  }
}

}  // namespace dronecu
