// Reset, observation, stick front-end and the Racer (mode B) kernels.
#pragma once
#include "../../include/fpv_api.h"
#include "vec.cuh"

namespace fpv {

// ---------------------------------------------------------------------------------------------
// Drone.reset, components.py:150-169.  R = Rz(yaw) Ry(pitch) Rx(roll) from degrees
// (helper_functions.py:39-44).  Rare path: evaluated in double, stored as float.
// ---------------------------------------------------------------------------------------------
__global__ void drone_reset_kernel(float4* state, long long n, long long stride, const float* pos, const float* vel,
                                   const float* rpy_deg, const unsigned char* mask) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (mask && !mask[e]) return;
  const double d2r = 0.017453292519943295;
  const double roll = (double)rpy_deg[3 * e] * d2r, pitch = (double)rpy_deg[3 * e + 1] * d2r,
               yaw = (double)rpy_deg[3 * e + 2] * d2r;
  double sr, cr, sp, cp, sy, cy;
  sincos(roll, &sr, &cr);
  sincos(pitch, &sp, &cp);
  sincos(yaw, &sy, &cy);
  state[e] = make_float4(pos[3 * e], pos[3 * e + 1], pos[3 * e + 2], 0.f);
  state[stride + e] = make_float4(vel[3 * e], vel[3 * e + 1], vel[3 * e + 2], __int_as_float(0));
  state[2 * stride + e] = make_float4((float)(cy * cp), (float)(cy * sp * sr - sy * cr), (float)(cy * sp * cr + sy * sr), 0.f);
  state[3 * stride + e] = make_float4((float)(sy * cp), (float)(sy * sp * sr + cy * cr), (float)(sy * sp * cr - cy * sr), 0.f);
  state[4 * stride + e] = make_float4((float)(-sp), (float)(cp * sr), (float)(cp * cr), 0.f);
}

// ---------------------------------------------------------------------------------------------
// The tuple Drone.step returns, components.py:247-248:
//   R^T, euler_angles_to_rotation_matrix(*rates)  [rates are deg/s but consumed as radians], R @ acc
// ---------------------------------------------------------------------------------------------
__global__ void drone_observe_kernel(const float4* state, long long n, long long stride, const float4* acc, float* Rt,
                                     float* gyro, float* accel) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float4 r0 = state[2 * stride + e], r1 = state[3 * stride + e], r2 = state[4 * stride + e];
  if (Rt) {
    float* o = Rt + 9 * e;
    o[0] = r0.x; o[1] = r1.x; o[2] = r2.x;
    o[3] = r0.y; o[4] = r1.y; o[5] = r2.y;
    o[6] = r0.z; o[7] = r1.z; o[8] = r2.z;
  }
  if (gyro) {
    float sr, cr, sp, cp, sy, cy;
    sincosf(r0.w, &sr, &cr);
    sincosf(r1.w, &sp, &cp);
    sincosf(r2.w, &sy, &cy);
    float* o = gyro + 9 * e;
    o[0] = cy * cp; o[1] = cy * sp * sr - sy * cr; o[2] = cy * sp * cr + sy * sr;
    o[3] = sy * cp; o[4] = sy * sp * sr + cy * cr; o[5] = sy * sp * cr - cy * sr;
    o[6] = -sp;     o[7] = cp * sr;                o[8] = cp * cr;
  }
  if (accel && acc) {
    const float4 a = acc[e];
    accel[3 * e] = r0.x * a.x + r0.y * a.y + r0.z * a.z;
    accel[3 * e + 1] = r1.x * a.x + r1.y * a.y + r1.z * a.z;
    accel[3 * e + 2] = r2.x * a.x + r2.y * a.y + r2.z * a.z;
  }
}

// ---------------------------------------------------------------------------------------------
// Joystick.calib_read (get_sticks.py:245-265) + Drone.read_sticks (components.py:250-253)
// ---------------------------------------------------------------------------------------------
struct StickK {
  float min_v[6], inv_span2[6], sign[6];  // inv_span2 = 2/(max-min)
  int idx[4];
  float center[4], inv_lo[4], inv_hi[4];  // 1/(c+1), 1/(1-c)
};

__global__ void sticks_kernel(const __grid_constant__ StickK k, const int* raw, long long n, float4* actions,
                              float* calibrated) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float v[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    // mapFromTo(r, min, max, -1, 1) * sign_reverse
    v[i] = fmaf((float)raw[6 * e + i] - k.min_v[i], k.inv_span2[i], -1.f) * k.sign[i];
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      if (i == k.idx[s]) {
        const float x = v[i], c = k.center[s];
        // x <= c: mapFromTo(x,-1,c,-1,0) ; else mapFromTo(x,c,1,0,1)
        v[i] = (x <= c) ? fmaf(x + 1.f, k.inv_lo[s], -1.f) : (x - c) * k.inv_hi[s];
      }
    }
  }
  // throttle, roll, pitch, arm, _, yaw = calib[0..5]; action = [-roll, pitch, yaw, throttle]
  actions[e] = make_float4(-v[1], v[2], v[5], v[0]);
  if (calibrated) {
#pragma unroll
    for (int i = 0; i < 6; ++i) calibrated[6 * e + i] = v[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Mode B: Racer (tests/racer_drone_test.py)
// ---------------------------------------------------------------------------------------------
struct RacerK {
  float dt, inv_dt;
  int substeps;
  float inv_mass;
  float dt_over_I[3];
  float gains[3][3];
  float vel_decay;
};

__global__ void racer_reset_kernel(float4* state, long long n, long long stride, const unsigned char* mask) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (mask && !mask[e]) return;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  state[e] = make_float4(0.f, 0.f, 0.f, 1.f);  // .w = PID first-call flag (racer_drone_test.py:20)
  state[stride + e] = z;
  state[2 * stride + e] = make_float4(1.f, 0.f, 0.f, 0.f);
  state[3 * stride + e] = make_float4(0.f, 1.f, 0.f, 0.f);
  state[4 * stride + e] = make_float4(0.f, 0.f, 1.f, 0.f);
  state[5 * stride + e] = z;
  state[6 * stride + e] = z;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) racer_step_kernel(const __grid_constant__ RacerK k, float4* state, long long n,
                                                             long long stride, const float4* actions, float4* torque_out) {
  const long long e = (long long)blockIdx.x * THREADS + threadIdx.x;
  if (e >= n) return;
  float4 q0 = ldg_stream(state + e), q1 = ldg_stream(state + stride + e);
  float4 r0 = ldg_stream(state + 2 * stride + e), r1 = ldg_stream(state + 3 * stride + e),
         r2 = ldg_stream(state + 4 * stride + e);
  float4 qi = ldg_stream(state + 5 * stride + e), ql = ldg_stream(state + 6 * stride + e);
  const float4 a = ldg_stream(actions + e);
  float w[3] = {r0.w, r1.w, r2.w};
  float ie[3] = {qi.x, qi.y, qi.z}, le[3] = {ql.x, ql.y, ql.z};
  const float sp[3] = {a.x, a.y, a.z};
  bool first = q0.w != 0.f;
  float tq[3] = {0.f, 0.f, 0.f};
#pragma unroll 1
  for (int it = 0; it < k.substeps; ++it) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      // PID.step, racer_drone_test.py:22-32
      const float err = sp[i] - w[i];
      ie[i] = fmaf(err, k.dt, ie[i]);
      const float de = first ? 0.f : (err - le[i]) * k.inv_dt;
      le[i] = err;
      tq[i] = fmaf(k.gains[i][0], err, fmaf(k.gains[i][1], ie[i], k.gains[i][2] * de));
      w[i] = fmaf(tq[i], k.dt_over_I[i], w[i]);  // :98
    }
    first = false;
    // orientation <- orientation @ Rx(w0) Ry(w1) Rz(w2)   (scipy "XYZ" intrinsic, angle = omega, :99)
    float sa, ca, sb, cb, sc, cc;
    sincosf(w[0], &sa, &ca);
    sincosf(w[1], &sb, &cb);
    sincosf(w[2], &sc, &cc);
    const float e00 = cb * cc, e01 = -cb * sc, e02 = sb;
    const float e10 = fmaf(sa * sb, cc, ca * sc), e11 = fmaf(-sa * sb, sc, ca * cc), e12 = -sa * cb;
    const float e20 = fmaf(-ca * sb, cc, sa * sc), e21 = fmaf(ca * sb, sc, sa * cc), e22 = ca * cb;
    float t0, t1, t2;
    t0 = fmaf(r0.x, e00, fmaf(r0.y, e10, r0.z * e20)); t1 = fmaf(r0.x, e01, fmaf(r0.y, e11, r0.z * e21));
    t2 = fmaf(r0.x, e02, fmaf(r0.y, e12, r0.z * e22)); r0.x = t0; r0.y = t1; r0.z = t2;
    t0 = fmaf(r1.x, e00, fmaf(r1.y, e10, r1.z * e20)); t1 = fmaf(r1.x, e01, fmaf(r1.y, e11, r1.z * e21));
    t2 = fmaf(r1.x, e02, fmaf(r1.y, e12, r1.z * e22)); r1.x = t0; r1.y = t1; r1.z = t2;
    t0 = fmaf(r2.x, e00, fmaf(r2.y, e10, r2.z * e20)); t1 = fmaf(r2.x, e01, fmaf(r2.y, e11, r2.z * e21));
    t2 = fmaf(r2.x, e02, fmaf(r2.y, e12, r2.z * e22)); r2.x = t0; r2.y = t1; r2.z = t2;
    // force = thrust * body z, acc = F/m, v <- decay*v + a dt, x <- x + v_new dt  (:100-103)
    const float s = a.w * k.inv_mass * k.dt;
    q1.x = fmaf(k.vel_decay, q1.x, r0.z * s);
    q1.y = fmaf(k.vel_decay, q1.y, r1.z * s);
    q1.z = fmaf(k.vel_decay, q1.z, r2.z * s);
    q0.x = fmaf(q1.x, k.dt, q0.x);
    q0.y = fmaf(q1.y, k.dt, q0.y);
    q0.z = fmaf(q1.z, k.dt, q0.z);
  }
  q0.w = first ? 1.f : 0.f;
  r0.w = w[0]; r1.w = w[1]; r2.w = w[2];
  stg_stream(state + e, q0);
  stg_stream(state + stride + e, q1);
  stg_stream(state + 2 * stride + e, r0);
  stg_stream(state + 3 * stride + e, r1);
  stg_stream(state + 4 * stride + e, r2);
  stg_stream(state + 5 * stride + e, make_float4(ie[0], ie[1], ie[2], 0.f));
  stg_stream(state + 6 * stride + e, make_float4(le[0], le[1], le[2], 0.f));
  if (torque_out) stg_stream(torque_out + e, make_float4(tq[0], tq[1], tq[2], 0.f));
}

}  // namespace fpv
