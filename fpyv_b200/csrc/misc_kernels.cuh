// Reset, observation and stick front-end kernels.
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

// The same map for the compact transport formats of fpv_drone_step_host_sticks: the raw readings of axes 0, 1, 2 and 5 --
// the four values Drone.read_sticks keeps of calib_read's six (components.py:251-252) -- as the joystick driver reports
// them (0..65535).  Identical arithmetic per axis, so the actions are bit-identical to sticks_kernel's.
__device__ __forceinline__ float4 sticks4_to_action(const StickK& k, float r0, float r1, float r2, float r3) {
  const int axis[4] = {0, 1, 2, 5};
  const float in[4] = {r0, r1, r2, r3};
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
  return make_float4(-v[1], v[2], v[3], v[0]);
}

// FPV_STICKS_U16: uint16[n][4].
__global__ void sticks4_u16_kernel(const __grid_constant__ StickK k, const ushort4* raw, long long n, float4* actions) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const ushort4 r = raw[e];
  actions[e] = sticks4_to_action(k, (float)r.x, (float)r.y, (float)r.z, (float)r.w);
}

// FPV_STICKS_CRSF: what an RC link actually carries -- four 11-bit channels (CRSF / SBUS resolution) packed little-endian
// into 6 bytes per env (channel c = bits [11 c, 11 c + 11); the top 4 bits are unused).  An 11-bit value v is widened to
// the driver's 16-bit range by bit replication, raw = (v << 5) | (v >> 6) (0 -> 0, 2047 -> 65535), then calibrated as above.
__device__ __forceinline__ float4 crsf_to_action(const StickK& k, unsigned h0, unsigned h1, unsigned h2) {
  const unsigned long long bits = (unsigned long long)h0 | ((unsigned long long)h1 << 16) | ((unsigned long long)h2 << 32);
  float r[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const unsigned v = (unsigned)(bits >> (11 * c)) & 0x7ffu;
    r[c] = (float)((v << 5) | (v >> 6));
  }
  return sticks4_to_action(k, r[0], r[1], r[2], r[3]);
}
__global__ void sticks4_crsf_kernel(const __grid_constant__ StickK k, const unsigned short* raw, long long n, float4* actions) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  actions[e] = crsf_to_action(k, raw[3 * e], raw[3 * e + 1], raw[3 * e + 2]);
}

}  // namespace fpv
