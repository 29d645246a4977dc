// Mode C ("acro"): the inner loop BASELINE.json's north_star names -- stick -> rate set-point, acro rate PID, motor
// mixer, per-motor thrust / torque from the T-Motor F80 bench curve (shared-memory LUT, linear interpolation) -- feeding
// the reference's translational model (drag, gravity, ground spring / crash, semi-explicit Euler).
// PARITY UNPINNED: the reference has no such model; the definition and the only oracle are
// oracle/acro_oracle.py.  Pieces with a reference counterpart cite it below.
// One env per thread, 7 float4 planes, K substeps in registers; the motor curve is read 4 times per substep from the
// shared-memory table.
#pragma once
#include "../../include/fpv_api.h"
#include "vec.cuh"
#include "drone_kernels.cuh"

namespace fpv {

struct AcroK {
  float dt, inv_dt;
  int substeps;
  float max_rates, rtr, one_minus_rtr, ttr, one_minus_ttr;  // components.py:185-194
  float deg2rad;
  float gains[3][3];
  float i_lim[3];         // integrator clamp = integral_limit / kI
  float mix[4][3];        // motor throttle += mix . (roll, pitch, yaw PID sums)
  float motor_xy[4][2];   // components.py:123-125
  float spin_kappa[4];    // spin_m * kappa
  float inertia[3], inv_inertia[3];
  float u_min, u_max;
  float poly[4];          // 4-motor bench cubic (components.py:136); per motor = / 4
  float lut_scale;
  int lut_n;
  float rc_cen[3], rc_span[3], rc_expo[3];   // stick curve: centre sensitivity, max(0, max - centre), expo (FPV_F_RATE_CURVE)
  float kd[3];            // k_drag, kinematics.py:36
  float kd_a, kd_b;       // kd[1] - kd[0], kd[2] - kd[0] (column form of the drag, see acro_body)
  float wind[3];
  float grav_z;           // -g m
  float inv_mass;
  float motor_radius, spring_k;
  unsigned flags;
};

// Table index without float<->int conversions (F2I / I2F / FRND run on the quarter-rate XU pipe, and mode C does four
// lookups per substep): t = (x - 0.5) + 1.5 * 2^23 rounds x - 0.5 to the nearest integer INTO THE MANTISSA of t, so the
// index is t's low bits and its float value is t - 1.5 * 2^23.  The rounding is to nearest-even, so the index is floor(x)
// or -- at an exact knot -- floor(x) - 1 with f = 1: the interpolation is continuous, the result is the same.  u is
// clamped to [u_min, u_max] within [-1, 1] (checked by the host), so 0 <= x <= lut_n - 1 and the index can reach
// lut_n - 1 (with f = 0): the staged table carries one padding entry behind its last.
#define FPV_LUT_MAGIC 12582912.f
template <class V> __device__ __forceinline__ V acro_thrust_lut(const AcroK& k, const float* lut_s, V u);
template <> __device__ __forceinline__ float acro_thrust_lut<float>(const AcroK& k, const float* lut_s, float u) {
  const float x = fmaf(u, k.lut_scale, k.lut_scale);
  const float t = (x - 0.5f) + FPV_LUT_MAGIC;
  const float f = x - (t - FPV_LUT_MAGIC);
  const int i = __float_as_int(t) & 0x3fffff;
  const float a = lut_s[i], b = lut_s[i + 1];
  return 0.25f * fmaf(f, b - a, a);
}
template <> __device__ __forceinline__ F2 acro_thrust_lut<F2>(const AcroK& k, const float* lut_s, F2 u) {
  // the index arithmetic and the interpolation packed, the two lookups per lane
  const F2 x = vfma(u, S<F2>(k.lut_scale), S<F2>(k.lut_scale));
  const F2 t = (x - S<F2>(0.5f)) + S<F2>(FPV_LUT_MAGIC);
  const F2 f = x - (t - S<F2>(FPV_LUT_MAGIC));
  float t0, t1;
  f2_unpack(t, t0, t1);
  const int i0 = __float_as_int(t0) & 0x3fffff, i1 = __float_as_int(t1) & 0x3fffff;
  const F2 a = f2_pack(lut_s[i0], lut_s[i1]), b = f2_pack(lut_s[i0 + 1], lut_s[i1 + 1]);
  return vfma(f, b - a, a) * S<F2>(0.25f);
}
// 4-motor bench cubic / 4 (components.py:136)
template <class V> __device__ __forceinline__ V acro_thrust_poly(const AcroK& k, V u) {
  const V pct = vfma(u, S<V>(50.f), S<V>(50.f));
  V p = vfma(S<V>(k.poly[0]), pct, S<V>(k.poly[1]));
  p = vfma(p, pct, S<V>(k.poly[2]));
  return vfma(p, pct, S<V>(k.poly[3])) * S<V>(0.25f);
}

// the motor-curve table as both step kernels stage it: lut_n entries + one copy of the last (see above)
__device__ __forceinline__ void acro_stage_lut(const AcroK& k, const float* lut, float* lut_s, int tid, int nthreads) {
  for (int i = tid; i <= k.lut_n; i += nthreads) lut_s[i] = lut[min(i, k.lut_n - 1)];
}

__global__ void acro_reset_kernel(float4* state, long long n, long long stride, const float* pos, const float* vel,
                                  const float* rpy_deg, const unsigned char* mask) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (mask && !mask[e]) return;
  const double d2r = 0.017453292519943295 * 0.5;
  double sr, cr, sp, cp, sy, cy;
  sincos((double)rpy_deg[3 * e] * d2r, &sr, &cr);
  sincos((double)rpy_deg[3 * e + 1] * d2r, &sp, &cp);
  sincos((double)rpy_deg[3 * e + 2] * d2r, &sy, &cy);
  double w = cy * cp * cr + sy * sp * sr, x = cy * cp * sr - sy * sp * cr;
  double y = cy * sp * cr + sy * cp * sr, z = sy * cp * cr - cy * sp * sr;
  if (w < 0.0) { w = -w; x = -x; y = -y; z = -z; }
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  state[e] = make_float4(pos[3 * e], pos[3 * e + 1], pos[3 * e + 2], -1.f);  // .w: filtered throttle, motors off
  state[stride + e] = make_float4(vel[3 * e], vel[3 * e + 1], vel[3 * e + 2], __int_as_float(0));
  state[2 * stride + e] = make_float4((float)w, (float)x, (float)y, (float)z);
  state[3 * stride + e] = make_float4(0.f, 0.f, 0.f, 1.f);                   // .w: PID first-call flag
  state[4 * stride + e] = zero;
  state[5 * stride + e] = zero;
  state[6 * stride + e] = zero;
}

// sin(a)/a and cos(a) of the half rotation angle from a^2.  |a| < 0.1 (|omega| < 200 rad/s at dt = 1 ms): series to a^6
// (truncation < 3e-13); beyond that the accurate functions, chosen PER LANE so that an env's result never depends on
// which other env shares its thread.
__device__ __forceinline__ void sinc_cos_small(float a2, float& sinc, float& cs) {
  sinc = fmaf(a2, fmaf(a2, fmaf(a2, -1.f / 5040.f, 1.f / 120.f), -1.f / 6.f), 1.f);
  cs = fmaf(a2, fmaf(a2, fmaf(a2, -1.f / 720.f, 1.f / 24.f), -0.5f), 1.f);
}
__device__ __forceinline__ void sinc_cos_any(float a2, float& sinc, float& cs) {
  if (a2 < 0.01f) { sinc_cos_small(a2, sinc, cs); return; }
  const float a = sqrtf(a2);
  float sn;
  sincosf(a, &sn, &cs);
  sinc = sn / a;
}
template <class V> __device__ __forceinline__ void sinc_cos(V a2, V& sinc, V& cs);
template <> __device__ __forceinline__ void sinc_cos<float>(float a2, float& sinc, float& cs) { sinc_cos_any(a2, sinc, cs); }
template <> __device__ __forceinline__ void sinc_cos<F2>(F2 a2, F2& sinc, F2& cs) {
  if (!vany(vle(S<F2>(0.01f), a2))) {   // common case: both lanes small -> packed series
    const F2 ps = vfma(a2, vfma(a2, S<F2>(-1.f / 5040.f), S<F2>(1.f / 120.f)), S<F2>(-1.f / 6.f));
    sinc = vfma(a2, ps, S<F2>(1.f));
    const F2 pc = vfma(a2, vfma(a2, S<F2>(-1.f / 720.f), S<F2>(1.f / 24.f)), S<F2>(-0.5f));
    cs = vfma(a2, pc, S<F2>(1.f));
    return;
  }
  float x, y, s0, c0, s1, c1;
  f2_unpack(a2, x, y);
  sinc_cos_any(x, s0, c0);
  sinc_cos_any(y, s1, c1);
  sinc = f2_pack(s0, s1);
  cs = f2_pack(c0, c1);
}

// The control steps of the L envs one thread owns (1 = float, 2 = packed F2: FFMA2/FMUL2/FADD2), shared by the plain
// kernel (acro_step_kernel: T >= 1 control steps per launch, slot l of thread t = env tile*TILE + l*THREADS + t) and the
// ring form (AcroMode: T = 1, slot l of lane t = env chunk*64 + 32 l + t).  q = the 7 state rows, act = the first step's
// actions; everything after the loads happens here, stores included.
// T control steps per launch (T = 1: fpv_acro_step; T > 1: fpv_acro_rollout, the open-loop form -- the state stays in
// registers from the first step to the last, step t reads actions[t * act_stride + env] and writes
// done_seq[t * done_stride + env]; crashes restart from the snapshot IN REGISTERS, so the result is bit-identical to
// T launches with T = 1).
template <class V, int SLOT_STRIDE>
__device__ __forceinline__ void acro_body(const AcroK& k, float4* state, long long n, long long stride, const float4* actions,
                                          const float* lut_s, unsigned char* done_out, float4* motor_out,
                                          const float4* reset_state, TileStats& st, const int T, const long long act_stride,
                                          unsigned char* done_seq, const long long done_stride,
                                          const float4 (&q)[FPV_ACRO_PLANES][Lane<V>::N], float4 (&act)[Lane<V>::N],
                                          const long long (&ei)[Lane<V>::N], const long long base) {
  constexpr int L = Lane<V>::N;
  using M = typename Lane<V>::Mask;
  float4 act_next[L];
  V px = Pack<V>::x(q[0]), py = Pack<V>::y(q[0]), pz = Pack<V>::z(q[0]), thr = Pack<V>::w(q[0]);
  V vx = Pack<V>::x(q[1]), vy = Pack<V>::y(q[1]), vz = Pack<V>::z(q[1]);
  V qw = Pack<V>::x(q[2]), qx = Pack<V>::y(q[2]), qy = Pack<V>::z(q[2]), qz = Pack<V>::w(q[2]);
  V sp[3] = {Pack<V>::x(q[3]), Pack<V>::y(q[3]), Pack<V>::z(q[3])};
  V w[3] = {Pack<V>::x(q[4]), Pack<V>::y(q[4]), Pack<V>::z(q[4])};
  V ie[3] = {Pack<V>::x(q[5]), Pack<V>::y(q[5]), Pack<V>::z(q[5])};
  V le[3] = {Pack<V>::x(q[6]), Pack<V>::y(q[6]), Pack<V>::z(q[6])};
  int epi[L];
  float nf_[2], first_flag[2];
#pragma unroll
  for (int l = 0; l < L; ++l) { epi[l] = __float_as_int(q[1][l].w); nf_[l] = q[3][l].w != 0.f ? 0.f : 1.f; first_flag[l] = q[3][l].w; }
  V notfirst = Lane<V>::make(nf_[0], nf_[L - 1]);   // 0 on a PID's first call: no derivative term (racer_drone_test.py:28)
  const V zero = S<V>(0.f), one = S<V>(1.f);
  const V mr = S<V>(k.max_rates), rtr = S<V>(k.rtr);
  const V omr = S<V>(k.one_minus_rtr), omt = S<V>(k.one_minus_ttr), dt = S<V>(k.dt), inv_dt = S<V>(k.inv_dt);
  const V d2r = S<V>(k.deg2rad), hdt = S<V>(0.5f * k.dt), s_m = S<V>(k.inv_mass * k.dt);
  V fm[4] = {zero, zero, zero, zero};
  for (int t = 0; t < T; ++t) {
  if (t + 1 < T) {   // next step's sticks travel while this step computes
#pragma unroll
    for (int l = 0; l < L; ++l) act_next[l] = ldg_stream(actions + (long long)(t + 1) * act_stride + ei[l]);
  }
  V cmd[3];
  {
    const V a3[3] = {Pack<V>::x(act), Pack<V>::y(act), Pack<V>::z(act)};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (k.flags & FPV_F_RATE_CURVE) {   // "actual rates": s c + span |s| (s^5 e + s (1 - e)),  s = -stick in [-1, 1]
        const V sx = vmin(vmax(vneg(a3[i]), S<V>(-1.f)), one);
        const V s2 = sx * sx;
        const V s5 = (s2 * s2) * sx;
        const V e = S<V>(k.rc_expo[i]);
        const V ex = vabs(sx) * vfma(s5, e, sx * (one - e));
        cmd[i] = vfma(S<V>(k.rc_span[i]), ex, sx * S<V>(k.rc_cen[i])) * rtr;
      } else {
        cmd[i] = vmin(vmax(vneg(a3[i]) * mr, vneg(mr)), mr) * rtr;   // components.py:185
      }
    }
  }
  const V thr_in = Pack<V>::w(act) * S<V>(k.ttr);
  M done = vlt(one, zero);
#pragma unroll 1
  for (int it = 0; it < k.substeps; ++it) {
    // ---- stick -> rate set-point / collective throttle, low-passed (components.py:185-194)
    thr = vfma(thr, omt, thr_in);
    V pid[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      sp[i] = vfma(sp[i], omr, cmd[i]);
      // ---- rate PID (racer_drone_test.py:22-32) with an integrator clamp
      const V err = vfma(sp[i], d2r, vneg(w[i]));
      ie[i] = vmin(vmax(vfma(err, dt, ie[i]), S<V>(-k.i_lim[i])), S<V>(k.i_lim[i]));
      const V de = ((err - le[i]) * inv_dt) * notfirst;
      le[i] = err;
      pid[i] = vfma(S<V>(k.gains[i][0]), err, vfma(S<V>(k.gains[i][1]), ie[i], S<V>(k.gains[i][2]) * de));
    }
    notfirst = one;
    // ---- mixer in throttle units, per-motor saturation, bench curve (shared-memory LUT) -> per-motor thrust
    V tx = zero, ty = zero, tz = zero, fsum = zero;
    V u4[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const V u = vfma(S<V>(k.mix[m][0]), pid[0], vfma(S<V>(k.mix[m][1]), pid[1], vfma(S<V>(k.mix[m][2]), pid[2], thr)));
      u4[m] = vmin(vmax(u, S<V>(k.u_min)), S<V>(k.u_max));
    }
    // ONE branch on the table flag around all four motors, so that the eight table reads of a lane are in flight together
    if (k.flags & FPV_F_THRUST_LUT) {
#pragma unroll
      for (int m = 0; m < 4; ++m) fm[m] = acro_thrust_lut<V>(k, lut_s, u4[m]);
    } else {
#pragma unroll
      for (int m = 0; m < 4; ++m) fm[m] = acro_thrust_poly<V>(k, u4[m]);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const V f = fm[m];
      fsum = fsum + f;
      tx = vfma(S<V>(k.motor_xy[m][1]), f, tx);     // arm x thrust: roll torque  =  sum y_m f_m
      ty = vfma(S<V>(-k.motor_xy[m][0]), f, ty);    //               pitch torque = -sum x_m f_m
      tz = vfma(S<V>(k.spin_kappa[m]), f, tz);      // rotor reaction torque
    }
    // ---- Euler's rigid-body equation, diagonal inertia
    const V Iw0 = S<V>(k.inertia[0]) * w[0], Iw1 = S<V>(k.inertia[1]) * w[1], Iw2 = S<V>(k.inertia[2]) * w[2];
    const V wd0 = (tx - vfma(w[1], Iw2, vneg(w[2] * Iw1))) * S<V>(k.inv_inertia[0]);
    const V wd1 = (ty - vfma(w[2], Iw0, vneg(w[0] * Iw2))) * S<V>(k.inv_inertia[1]);
    const V wd2 = (tz - vfma(w[0], Iw1, vneg(w[1] * Iw0))) * S<V>(k.inv_inertia[2]);
    // ---- the reference's force model on the current attitude (components.py:233-243), in mode A's column form
    //      (drone_kernels.cuh): R diag(k)|u| R^T u = |u| (k0 u + (k1-k0)(c1.u) c1 + (k2-k0)(c2.u) c2) needs only the columns
    //      c1, c2 of R(q) (plus R[2][0] for the motor heights); the thrust rides on the c2 coefficient.  Every "2 q_a q_b"
    //      of the matrix is one product with the doubled component (2 q is exact, so K substeps in one step and K steps
    //      of one substep still give the same bits).
    const V nqw = vneg(qw);
    const V dqx = qx + qx, dqy = qy + qy, dqz = qz + qz;
    const V xz = qx * dqz, yz = qy * dqz, xy = qx * dqy;
    const V r02 = vfma(qw, dqy, xz), r20 = vfma(nqw, dqy, xz);
    const V r21 = vfma(qw, dqx, yz), r12 = vfma(nqw, dqx, yz);
    const V r01 = vfma(nqw, dqz, xy);
    const V t1 = vfma(vneg(qx), dqx, one);
    const V r11 = vfma(vneg(qz), dqz, t1), r22 = vfma(vneg(qy), dqy, t1);
    const V ux = vx + S<V>(k.wind[0]), uy = vy + S<V>(k.wind[1]), uz = vz + S<V>(k.wind[2]);   // kinematics.py:34 (PLUS wind)
    const V nrm = vsqrt_fast(vfma(ux, ux, vfma(uy, uy, uz * uz)));
    const V d1 = vfma(r01, ux, vfma(r11, uy, r21 * uz));
    const V d2 = vfma(r02, ux, vfma(r12, uy, r22 * uz));
    const V ks = S<V>(k.kd[0]) * nrm;
    const V g1 = (S<V>(k.kd_a) * nrm) * d1;
    const V g2 = vfma(S<V>(k.kd_b) * nrm, d2, fsum);
    const V Fx = vfma(g2, r02, vfma(g1, r01, ks * ux));
    const V Fy = vfma(g2, r12, vfma(g1, r11, ks * uy));
    V Fz = vfma(g2, r22, vfma(g1, r21, vfma(ks, uz, S<V>(k.grav_z))));
    if (k.flags & FPV_F_GROUND) {   // components.py:198-214, :239 with the plane z = 0
      M crashed = vlt(one, zero);
      V comp = zero;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const V mz = vfma(S<V>(k.motor_xy[m][0]), r20, vfma(S<V>(k.motor_xy[m][1]), r21, pz));
        crashed = vor(crashed, vlt(mz, zero));
        comp = comp + vmax(S<V>(k.motor_radius) - mz, zero);   // spring compression of motor m
      }
      Fz = vfma(S<V>(k.spring_k), vsel(crashed, zero, comp), Fz);
      done = vor(done, crashed);
    }
    // ---- translation, kinematics.py:21-22 (old velocity first)
    px = vfma(vx, dt, px); py = vfma(vy, dt, py); pz = vfma(vz, dt, pz);
    vx = vfma(Fx, s_m, vx); vy = vfma(Fy, s_m, vy); vz = vfma(Fz, s_m, vz);
    // ---- rotation: omega first, then q <- q (x) exp(omega dt / 2)
    w[0] = vfma(wd0, dt, w[0]); w[1] = vfma(wd1, dt, w[1]); w[2] = vfma(wd2, dt, w[2]);
    const V hx = hdt * w[0], hy = hdt * w[1], hz = hdt * w[2];
    const V ang2 = vfma(hx, hx, vfma(hy, hy, hz * hz));
    V sinc, cs;
    sinc_cos<V>(ang2, sinc, cs);
    const V dx = hx * sinc, dy = hy * sinc, dz = hz * sinc;
    const V nw = vfma(qw, cs, vneg(vfma(qx, dx, vfma(qy, dy, qz * dz))));
    const V nx = vfma(qw, dx, vfma(qx, cs, vfma(qy, dz, vneg(qz * dy))));
    const V ny = vfma(qw, dy, vfma(qy, cs, vfma(qz, dx, vneg(qx * dz))));
    const V nz = vfma(qw, dz, vfma(qz, cs, vfma(qx, dy, vneg(qy * dx))));
    const V inv = vrsqrt_fast(vfma(nw, nw, vfma(nx, nx, vfma(ny, ny, nz * nz))));
    qw = nw * inv; qx = nx * inv; qy = ny * inv; qz = nz * inv;
  }
  // ---- per-step bookkeeping on registers: flags, statistics, restart from the snapshot
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const long long e = base + (long long)l * SLOT_STRIDE;
    if (e >= n) break;
    const bool d = mask_get(done, l);
    epi[l] += 1;
    first_flag[l] = 0.f;   // the PIDs have been called
    if (done_seq) done_seq[(long long)t * done_stride + e] = d ? 1 : 0;
    if (done_out && t == T - 1) done_out[e] = d ? 1 : 0;
    if (d) {   // rare events: accumulated per thread, flushed once per launch (stats_warp_flush)
      st.crash += 1.f;
      if (k.flags & FPV_F_AUTO_RESET) { st.epi += 1.f; st.len += (float)epi[l]; }
    }
    if (d && (k.flags & FPV_F_AUTO_RESET)) {
      float4 v[FPV_ACRO_PLANES];
#pragma unroll
      for (int p = 0; p < FPV_ACRO_PLANES; ++p) v[p] = ldg_stream(reset_state + p * stride + e);
      px = lane_set<V>(px, l, v[0].x); py = lane_set<V>(py, l, v[0].y); pz = lane_set<V>(pz, l, v[0].z); thr = lane_set<V>(thr, l, v[0].w);
      vx = lane_set<V>(vx, l, v[1].x); vy = lane_set<V>(vy, l, v[1].y); vz = lane_set<V>(vz, l, v[1].z);
      qw = lane_set<V>(qw, l, v[2].x); qx = lane_set<V>(qx, l, v[2].y); qy = lane_set<V>(qy, l, v[2].z); qz = lane_set<V>(qz, l, v[2].w);
      sp[0] = lane_set<V>(sp[0], l, v[3].x); sp[1] = lane_set<V>(sp[1], l, v[3].y); sp[2] = lane_set<V>(sp[2], l, v[3].z);
      w[0] = lane_set<V>(w[0], l, v[4].x); w[1] = lane_set<V>(w[1], l, v[4].y); w[2] = lane_set<V>(w[2], l, v[4].z);
      ie[0] = lane_set<V>(ie[0], l, v[5].x); ie[1] = lane_set<V>(ie[1], l, v[5].y); ie[2] = lane_set<V>(ie[2], l, v[5].z);
      le[0] = lane_set<V>(le[0], l, v[6].x); le[1] = lane_set<V>(le[1], l, v[6].y); le[2] = lane_set<V>(le[2], l, v[6].z);
      first_flag[l] = v[3].w;
      notfirst = lane_set<V>(notfirst, l, v[3].w != 0.f ? 0.f : 1.f);
      epi[l] = 0;
    }
  }
#pragma unroll
  for (int l = 0; l < L; ++l) act[l] = act_next[l];
  }  // control steps
  // ---- the state goes back once
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const long long e = base + (long long)l * SLOT_STRIDE;
    if (e >= n) break;
    if (motor_out)
      stg_stream(motor_out + e, make_float4(Lane<V>::get(fm[0], l), Lane<V>::get(fm[1], l), Lane<V>::get(fm[2], l), Lane<V>::get(fm[3], l)));
    const auto g = [&](V v) { return Lane<V>::get(v, l); };
    stg_stream(state + e, make_float4(g(px), g(py), g(pz), g(thr)));
    stg_stream(state + stride + e, make_float4(g(vx), g(vy), g(vz), __int_as_float(epi[l])));
    stg_stream(state + 2 * stride + e, make_float4(g(qw), g(qx), g(qy), g(qz)));
    stg_stream(state + 3 * stride + e, make_float4(g(sp[0]), g(sp[1]), g(sp[2]), first_flag[l]));   // .w: PID first-call flag
    stg_stream(state + 4 * stride + e, make_float4(g(w[0]), g(w[1]), g(w[2]), 0.f));
    stg_stream(state + 5 * stride + e, make_float4(g(ie[0]), g(ie[1]), g(ie[2]), 0.f));
    stg_stream(state + 6 * stride + e, make_float4(g(le[0]), g(le[1]), g(le[2]), 0.f));
  }
}

// One thread owns L envs; CTA tile = THREADS*L envs, slot l of thread t is env tile*TILE + l*THREADS + t (every 128-bit
// access of a warp is one contiguous 512 B run).  Used for rollouts (T > 1), the scalar cross-check and as the fallback.
template <class V, int THREADS>
__global__ void __launch_bounds__(THREADS) acro_step_kernel(const __grid_constant__ AcroK k, float4* state, long long n,
                                                            long long stride, const float4* actions, const float* lut,
                                                            unsigned char* done_out, float4* motor_out,
                                                            const float4* reset_state, fpv_stats_t* stats, const int T,
                                                            const long long act_stride, unsigned char* done_seq,
                                                            const long long done_stride) {
  constexpr int L = Lane<V>::N;
  constexpr int TILE = THREADS * L;
  extern __shared__ float lut_s[];
  if (k.flags & FPV_F_THRUST_LUT) {
    acro_stage_lut(k, lut, lut_s, threadIdx.x, THREADS);
    __syncthreads();
  }
  const long long base = (long long)blockIdx.x * TILE + threadIdx.x;
  TileStats st = {0.f, 0.f, 0.f, 0.f, 0.f};
  if (base < n) {
    long long ei[L];
#pragma unroll
    for (int l = 0; l < L; ++l) ei[l] = min(base + (long long)l * THREADS, n - 1);   // a slot past the end is never stored
    float4 q[FPV_ACRO_PLANES][L], act[L];
#pragma unroll
    for (int p = 0; p < FPV_ACRO_PLANES; ++p)
#pragma unroll
      for (int l = 0; l < L; ++l) q[p][l] = ldg_stream(state + p * stride + ei[l]);
#pragma unroll
    for (int l = 0; l < L; ++l) act[l] = ldg_stream(actions + ei[l]);
    acro_body<V, THREADS>(k, state, n, stride, actions, lut_s, done_out, motor_out, reset_state, st, T, act_stride, done_seq,
                          done_stride, q, act, ei, base);
  }
  if (stats) stats_warp_flush(stats, st);
}

// The ring form of one control step (fpv_acro_step, packed): 7 state planes + actions per chunk, LUT staged once per CTA.
struct AcroIO {
  float4* state;
  long long n, stride;
  const float4* actions;
  const float* lut;
  unsigned char* done;
  float4* motor_out;
  const float4* reset_state;
  fpv_stats_t* stats;
  unsigned* work;
  unsigned* chunk_epoch;   // always null: mode C has no chained form
  unsigned epoch;
  unsigned* err;
  unsigned long long* trace;
};

template <class V_>
struct AcroMode {
  using V = V_;
  using K = AcroK;
  using IO = AcroIO;
  static constexpr int ROWS = FPV_ACRO_PLANES + 1;
  struct Ctx {
    TileStats st;
  };
  static __device__ __forceinline__ bool chained(const K&) { return false; }
  static __device__ __forceinline__ constexpr int row_bytes(int) { return 16; }
  static __device__ __forceinline__ const void* row_ptr(const IO& io, int r, long long first) {
    return (r < FPV_ACRO_PLANES ? io.state + r * io.stride : io.actions) + first;
  }
  static __device__ __forceinline__ void stage(const K& k, const IO& io, unsigned char* smem, int tid, int nthreads) {
    float* lut_s = reinterpret_cast<float*>(smem);
    if (k.flags & FPV_F_THRUST_LUT) acro_stage_lut(k, io.lut, lut_s, tid, nthreads);
  }
  static __device__ __forceinline__ Ctx begin(const K&, const IO&) { return Ctx{TileStats{0.f, 0.f, 0.f, 0.f, 0.f}}; }
  template <class PreStore>
  static __device__ __forceinline__ void tile(const K& k, const IO& io, const unsigned char* staged,
                                              const float4 (&rows)[ROWS][Lane<V>::N], const long long (&ei)[Lane<V>::N],
                                              long long base, Ctx& c, PreStore pre_store) {
    constexpr int L = Lane<V>::N;
    pre_store();
    if (base >= io.n) return;
    float4 q[FPV_ACRO_PLANES][L], act[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
#pragma unroll
      for (int p = 0; p < FPV_ACRO_PLANES; ++p) q[p][l] = rows[p][l];
      act[l] = rows[FPV_ACRO_PLANES][l];
    }
    acro_body<V, 32>(k, io.state, io.n, io.stride, io.actions, reinterpret_cast<const float*>(staged), io.done, io.motor_out,
                     io.reset_state, c.st, 1, 0, nullptr, 0, q, act, ei, base);
  }
  static __device__ __forceinline__ void finish(const K&, const IO& io, Ctx& c) {
    if (io.stats) stats_warp_flush(io.stats, c.st);
  }
};

}  // namespace fpv
