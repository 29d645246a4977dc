// Reset, observation, stick front-end and the Racer (mode B) kernels.
#pragma once
#include "../../include/fpv_api.h"
#include "vec.cuh"

namespace fpv {

// ---------------------------------------------------------------------------------------------
// Quaternion <-> matrix (reference conventions: src/utils/helper_functions.py:65-80, :100-117)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void quat_to_matrix(float4 q, float (&R)[9]) {  // q = (w,x,y,z) in .x .y .z .w
  // explicit fused multiply-adds in a fixed order: the same bits in every kernel this is inlined in
  const float w = q.x, x = q.y, y = q.z, z = q.w;
  const float x2 = x + x, y2 = y + y, z2 = z + z;
  R[0] = __fmaf_rn(-y2, y, __fmaf_rn(-z2, z, 1.f)); R[1] = __fmaf_rn(x2, y, -(z2 * w));              R[2] = __fmaf_rn(x2, z, y2 * w);
  R[3] = __fmaf_rn(x2, y, z2 * w);                  R[4] = __fmaf_rn(-x2, x, __fmaf_rn(-z2, z, 1.f)); R[5] = __fmaf_rn(y2, z, -(x2 * w));
  R[6] = __fmaf_rn(x2, z, -(y2 * w));               R[7] = __fmaf_rn(y2, z, x2 * w);                  R[8] = __fmaf_rn(-x2, x, __fmaf_rn(-y2, y, 1.f));
}

// Robust matrix -> unit quaternion (Shepperd's branch on the largest diagonal term; the reference's
// rotation_matrix_to_quaternion, helper_functions.py:65-80, is the trace branch only and divides by zero at 180 deg).
// Evaluated in double and normalised; sign fixed to w >= 0.
__device__ __forceinline__ float4 matrix_to_quat(const float* R) {
  const double m00 = R[0], m01 = R[1], m02 = R[2], m10 = R[3], m11 = R[4], m12 = R[5], m20 = R[6], m21 = R[7], m22 = R[8];
  double w, x, y, z;
  const double tr = m00 + m11 + m22;
  if (tr > 0.0) {
    const double s = sqrt(tr + 1.0) * 2.0;
    w = 0.25 * s; x = (m21 - m12) / s; y = (m02 - m20) / s; z = (m10 - m01) / s;
  } else if (m00 > m11 && m00 > m22) {
    const double s = sqrt(1.0 + m00 - m11 - m22) * 2.0;
    w = (m21 - m12) / s; x = 0.25 * s; y = (m01 + m10) / s; z = (m02 + m20) / s;
  } else if (m11 > m22) {
    const double s = sqrt(1.0 + m11 - m00 - m22) * 2.0;
    w = (m02 - m20) / s; x = (m01 + m10) / s; y = 0.25 * s; z = (m12 + m21) / s;
  } else {
    const double s = sqrt(1.0 + m22 - m00 - m11) * 2.0;
    w = (m10 - m01) / s; x = (m02 + m20) / s; y = (m12 + m21) / s; z = 0.25 * s;
  }
  double n = rsqrt(w * w + x * x + y * y + z * z);
  if (w < 0.0) n = -n;
  return make_float4((float)(w * n), (float)(x * n), (float)(y * n), (float)(z * n));
}

// ---------------------------------------------------------------------------------------------
// Drone.reset, components.py:150-169.  R = Rz(yaw) Ry(pitch) Rx(roll) from degrees
// (helper_functions.py:39-44), stored as its quaternion.  Rare path: evaluated in double.
// ---------------------------------------------------------------------------------------------
__global__ void drone_reset_kernel(float4* state, long long n, long long stride, const float* pos, const float* vel,
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
  state[e] = make_float4(pos[3 * e], pos[3 * e + 1], pos[3 * e + 2], 0.f);
  state[stride + e] = make_float4(vel[3 * e], vel[3 * e + 1], vel[3 * e + 2], __int_as_float(0));
  state[2 * stride + e] = make_float4((float)w, (float)x, (float)y, (float)z);
  state[3 * stride + e] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// rotation_matrix attribute of the reference object (components.py:154): read / write through the quaternion plane
__global__ void drone_get_rotation_kernel(const float4* state, long long n, long long stride, float* R) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float m[9];
  quat_to_matrix(state[2 * stride + e], m);
#pragma unroll
  for (int i = 0; i < 9; ++i) R[9 * e + i] = m[i];
}
__global__ void drone_set_rotation_kernel(float4* state, long long n, long long stride, const float* R,
                                          const unsigned char* mask) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (mask && !mask[e]) return;
  state[2 * stride + e] = matrix_to_quat(R + 9 * e);
}
// free-standing conversion for override inputs: R[n][9] -> q[n] (float4)
__global__ void matrix_to_quat_kernel(const float* R, long long n, float4* q) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) q[e] = matrix_to_quat(R + 9 * e);
}

// ---------------------------------------------------------------------------------------------
// The tuple Drone.step returns, components.py:247-248:
//   R^T, euler_angles_to_rotation_matrix(*rates)  [rates are deg/s but consumed as radians], R @ acc
// ---------------------------------------------------------------------------------------------
__global__ void drone_observe_kernel(const float4* state, long long n, long long stride, const float4* acc, float* Rt,
                                     float* gyro, float* accel) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float R[9];
  quat_to_matrix(state[2 * stride + e], R);
  const float4 rates = state[3 * stride + e];
  if (Rt) {
    float* o = Rt + 9 * e;
    o[0] = R[0]; o[1] = R[3]; o[2] = R[6];
    o[3] = R[1]; o[4] = R[4]; o[5] = R[7];
    o[6] = R[2]; o[7] = R[5]; o[8] = R[8];
  }
  if (gyro) {
    float sr, cr, sp, cp, sy, cy;
    sincosf(rates.x, &sr, &cr);
    sincosf(rates.y, &sp, &cp);
    sincosf(rates.z, &sy, &cy);
    float* o = gyro + 9 * e;
    o[0] = cy * cp; o[1] = cy * sp * sr - sy * cr; o[2] = cy * sp * cr + sy * sr;
    o[3] = sy * cp; o[4] = sy * sp * sr + cy * cr; o[5] = sy * sp * cr - cy * sr;
    o[6] = -sp;     o[7] = cp * sr;                o[8] = cp * cr;
  }
  if (accel && acc) {
    const float4 a = acc[e];
    accel[3 * e] = R[0] * a.x + R[1] * a.y + R[2] * a.z;
    accel[3 * e + 1] = R[3] * a.x + R[4] * a.y + R[5] * a.z;
    accel[3 * e + 2] = R[6] * a.x + R[7] * a.y + R[8] * a.z;
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

// The same map for the compact transport format of fpv_drone_step_host_sticks: uint16[n][4] = the raw readings of
// axes 0, 1, 2 and 5 -- the four values Drone.read_sticks keeps of calib_read's six (components.py:251-252) -- as the
// joystick driver reports them (0..65535).  Identical arithmetic per axis, so the actions are bit-identical.
__global__ void sticks4_u16_kernel(const __grid_constant__ StickK k, const ushort4* raw, long long n, float4* actions) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const ushort4 r = raw[e];
  const int axis[4] = {0, 1, 2, 5};
  const float in[4] = {(float)r.x, (float)r.y, (float)r.z, (float)r.w};
  float v[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = axis[a];
    v[a] = fmaf(in[a] - k.min_v[i], k.inv_span2[i], -1.f) * k.sign[i];
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (axis[a] == k.idx[s]) {
        const float x = v[a], c = k.center[s];
        v[a] = (x <= c) ? fmaf(x + 1.f, k.inv_lo[s], -1.f) : (x - c) * k.inv_hi[s];
      }
    }
  }
  actions[e] = make_float4(-v[1], v[2], v[3], v[0]);
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
