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
  float kd[3];            // k_drag, kinematics.py:36
  float wind[3];
  float grav_z;           // -g m
  float inv_mass;
  float motor_radius, spring_k;
  unsigned flags;
};

__device__ __forceinline__ float acro_motor_thrust(const AcroK& k, const float* lut_s, float u) {
  if (k.flags & FPV_F_THRUST_LUT) {
    float x = (u + 1.f) * k.lut_scale;
    int i = (int)floorf(x);
    i = max(0, min(i, k.lut_n - 2));
    const float f = x - (float)i;
    const float a = lut_s[i], b = lut_s[i + 1];
    return 0.25f * fmaf(f, b - a, a);
  }
  const float pct = fmaf(u, 50.f, 50.f);
  float p = fmaf(k.poly[0], pct, k.poly[1]);
  p = fmaf(p, pct, k.poly[2]);
  return 0.25f * fmaf(p, pct, k.poly[3]);
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

template <int THREADS>
__global__ void __launch_bounds__(THREADS) acro_step_kernel(const __grid_constant__ AcroK k, float4* state, long long n,
                                                            long long stride, const float4* actions, const float* lut,
                                                            unsigned char* done_out, float4* motor_out,
                                                            const float4* reset_state, fpv_stats_t* stats) {
  extern __shared__ float lut_s[];
  if (k.flags & FPV_F_THRUST_LUT) {
    for (int i = threadIdx.x; i < k.lut_n; i += THREADS) lut_s[i] = lut[i];
    __syncthreads();
  }
  const long long e = (long long)blockIdx.x * THREADS + threadIdx.x;
  if (e >= n) return;
  float4 p0 = ldg_stream(state + e), p1 = ldg_stream(state + stride + e), q = ldg_stream(state + 2 * stride + e);
  float4 p3 = ldg_stream(state + 3 * stride + e), p4 = ldg_stream(state + 4 * stride + e);
  float4 p5 = ldg_stream(state + 5 * stride + e), p6 = ldg_stream(state + 6 * stride + e);
  const float4 a = ldg_stream(actions + e);
  float sp_deg[3] = {p3.x, p3.y, p3.z};
  float w[3] = {p4.x, p4.y, p4.z}, ie[3] = {p5.x, p5.y, p5.z}, le[3] = {p6.x, p6.y, p6.z};
  bool first = p3.w != 0.f;
  float thr = p0.w;
  const float act[3] = {a.x, a.y, a.z};
  float cmd[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) cmd[i] = fminf(fmaxf(-act[i] * k.max_rates, -k.max_rates), k.max_rates) * k.rtr;
  const float thr_in = a.w * k.ttr;
  bool done = false;
  float fm[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
  for (int it = 0; it < k.substeps; ++it) {
    // ---- stick -> rate set-point / collective throttle, low-passed (components.py:185-194)
    thr = fmaf(thr, k.one_minus_ttr, thr_in);
    float pid[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      sp_deg[i] = fmaf(sp_deg[i], k.one_minus_rtr, cmd[i]);
      // ---- rate PID (racer_drone_test.py:22-32) with an integrator clamp
      const float err = fmaf(sp_deg[i], k.deg2rad, -w[i]);
      ie[i] = fminf(fmaxf(fmaf(err, k.dt, ie[i]), -k.i_lim[i]), k.i_lim[i]);
      const float de = first ? 0.f : (err - le[i]) * k.inv_dt;
      le[i] = err;
      pid[i] = fmaf(k.gains[i][0], err, fmaf(k.gains[i][1], ie[i], k.gains[i][2] * de));
    }
    first = false;
    // ---- mixer in throttle units, per-motor saturation, bench curve (shared-memory LUT) -> per-motor thrust
    float tx = 0.f, ty = 0.f, tz = 0.f, fsum = 0.f;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      float u = fmaf(k.mix[m][0], pid[0], fmaf(k.mix[m][1], pid[1], fmaf(k.mix[m][2], pid[2], thr)));
      u = fminf(fmaxf(u, k.u_min), k.u_max);
      const float f = acro_motor_thrust(k, lut_s, u);
      fm[m] = f;
      fsum += f;
      tx = fmaf(k.motor_xy[m][1], f, tx);        // arm x thrust: roll torque  =  sum y_m f_m
      ty = fmaf(-k.motor_xy[m][0], f, ty);       //               pitch torque = -sum x_m f_m
      tz = fmaf(k.spin_kappa[m], f, tz);         // rotor reaction torque
    }
    // ---- Euler's rigid-body equation, diagonal inertia
    const float Iw0 = k.inertia[0] * w[0], Iw1 = k.inertia[1] * w[1], Iw2 = k.inertia[2] * w[2];
    const float wd0 = (tx - (w[1] * Iw2 - w[2] * Iw1)) * k.inv_inertia[0];
    const float wd1 = (ty - (w[2] * Iw0 - w[0] * Iw2)) * k.inv_inertia[1];
    const float wd2 = (tz - (w[0] * Iw1 - w[1] * Iw0)) * k.inv_inertia[2];
    // ---- the reference's force model on the current attitude (components.py:233-243)
    float R[9];
    {
      const float qw = q.x, qx = q.y, qy = q.z, qz = q.w;
      R[0] = 1.f - 2.f * (qy * qy + qz * qz); R[1] = 2.f * (qx * qy - qz * qw); R[2] = 2.f * (qx * qz + qy * qw);
      R[3] = 2.f * (qx * qy + qz * qw); R[4] = 1.f - 2.f * (qx * qx + qz * qz); R[5] = 2.f * (qy * qz - qx * qw);
      R[6] = 2.f * (qx * qz - qy * qw); R[7] = 2.f * (qy * qz + qx * qw); R[8] = 1.f - 2.f * (qx * qx + qy * qy);
    }
    const float ux = p1.x + k.wind[0], uy = p1.y + k.wind[1], uz = p1.z + k.wind[2];   // kinematics.py:34 (PLUS wind)
    const float nrm = sqrtf(fmaf(ux, ux, fmaf(uy, uy, uz * uz)));
    const float b0 = k.kd[0] * nrm * fmaf(R[0], ux, fmaf(R[3], uy, R[6] * uz));
    const float b1 = k.kd[1] * nrm * fmaf(R[1], ux, fmaf(R[4], uy, R[7] * uz));
    const float b2 = k.kd[2] * nrm * fmaf(R[2], ux, fmaf(R[5], uy, R[8] * uz)) + fsum;  // thrust rides on body z
    float Fx = fmaf(R[0], b0, fmaf(R[1], b1, R[2] * b2));
    float Fy = fmaf(R[3], b0, fmaf(R[4], b1, R[5] * b2));
    float Fz = fmaf(R[6], b0, fmaf(R[7], b1, R[8] * b2)) + k.grav_z;
    if (k.flags & FPV_F_GROUND) {   // components.py:198-214, :239 with the plane z = 0
      bool crashed = false;
      float spring = 0.f;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float mz = fmaf(k.motor_xy[m][0], R[6], fmaf(k.motor_xy[m][1], R[7], p0.z));
        crashed |= mz < 0.f;
        const float pen = mz - k.motor_radius;
        spring += pen < 0.f ? -k.spring_k * pen : 0.f;
      }
      Fz += crashed ? 0.f : spring;
      done |= crashed;
    }
    // ---- translation, kinematics.py:21-22 (old velocity first)
    p0.x = fmaf(p1.x, k.dt, p0.x); p0.y = fmaf(p1.y, k.dt, p0.y); p0.z = fmaf(p1.z, k.dt, p0.z);
    const float s = k.inv_mass * k.dt;
    p1.x = fmaf(Fx, s, p1.x); p1.y = fmaf(Fy, s, p1.y); p1.z = fmaf(Fz, s, p1.z);
    // ---- rotation: omega first, then q <- q (x) exp(omega dt / 2)
    w[0] = fmaf(wd0, k.dt, w[0]); w[1] = fmaf(wd1, k.dt, w[1]); w[2] = fmaf(wd2, k.dt, w[2]);
    const float hx = 0.5f * k.dt * w[0], hy = 0.5f * k.dt * w[1], hz = 0.5f * k.dt * w[2];
    const float ang2 = fmaf(hx, hx, fmaf(hy, hy, hz * hz));
    float sn, cs;
    const float ang = sqrtf(ang2);
    sincosf(ang, &sn, &cs);
    const float sinc = ang > 1e-6f ? sn / ang : 1.f - ang2 * (1.f / 6.f);
    const float dw = cs, dx = hx * sinc, dy = hy * sinc, dz = hz * sinc;
    const float qw = q.x, qx = q.y, qy = q.z, qz = q.w;
    float nw = qw * dw - qx * dx - qy * dy - qz * dz;
    float nx = qw * dx + qx * dw + qy * dz - qz * dy;
    float ny = qw * dy - qx * dz + qy * dw + qz * dx;
    float nz = qw * dz + qx * dy - qy * dx + qz * dw;
    const float inv = rsqrtf(fmaf(nw, nw, fmaf(nx, nx, fmaf(ny, ny, nz * nz))));
    q = make_float4(nw * inv, nx * inv, ny * inv, nz * inv);
  }
  int ep = __float_as_int(p1.w) + 1;
  if (done_out) done_out[e] = done ? 1 : 0;
  if (motor_out) stg_stream(motor_out + e, make_float4(fm[0], fm[1], fm[2], fm[3]));
  if (done && stats) {
    atomicAdd(&stats->crashes, 1.0);
    if (k.flags & FPV_F_AUTO_RESET) { atomicAdd(&stats->episodes, 1.0); atomicAdd(&stats->episode_len_sum, (double)ep); }
  }
  if (done && (k.flags & FPV_F_AUTO_RESET)) {
    float4 v[FPV_ACRO_PLANES];
#pragma unroll
    for (int p = 0; p < FPV_ACRO_PLANES; ++p) v[p] = ldg_stream(reset_state + p * stride + e);
    v[1].w = __int_as_float(0);
#pragma unroll
    for (int p = 0; p < FPV_ACRO_PLANES; ++p) stg_stream(state + p * stride + e, v[p]);
    return;
  }
  p0.w = thr;
  p1.w = __int_as_float(ep);
  stg_stream(state + e, p0);
  stg_stream(state + stride + e, p1);
  stg_stream(state + 2 * stride + e, q);
  stg_stream(state + 3 * stride + e, make_float4(sp_deg[0], sp_deg[1], sp_deg[2], first ? 1.f : 0.f));
  stg_stream(state + 4 * stride + e, make_float4(w[0], w[1], w[2], 0.f));
  stg_stream(state + 5 * stride + e, make_float4(ie[0], ie[1], ie[2], 0.f));
  stg_stream(state + 6 * stride + e, make_float4(le[0], le[1], le[2], 0.f));
}

}  // namespace fpv
