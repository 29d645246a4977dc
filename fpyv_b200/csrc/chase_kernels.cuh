// The callers on either side of Drone.step in the reference's chase loop (src/core/simulator.py:98-110):
//   Camera.update / projection / depth splat      src/utils/components.py:501-503, :532-536, :558-568, :584-629
//   mean target pixel                             src/core/simulator.py:104-108
//   components.PID + point-and-shoot autopilot    src/utils/components.py:43-54, :258-304
// Everything here is evaluated in DOUBLE on the device: the splat turns geometry into integer pixel indices and bytes
// (trunc toward zero), which are only reproducible against the float64 reference if the projection itself is float64;
// the work is per point / per env and tiny next to the dynamics step, and nowhere near the FP64 pipe's limits.
// The images are HBM-bound byte work: one memset of the frame plus one 8-bit atomic max per visible point.
#pragma once
#include "../../include/fpv_api.h"
#include "misc_kernels.cuh"

namespace fpv {

struct CamK {
  double rel_rot[9];  // WORLD2CAM^T Rx(pitch)                      components.py:455
  double rel_pos[3];  // position_relative_to_frame                 components.py:452
  double fx, fy, cx, cy;  // intrinsic matrix                       components.py:469-470
  int W, H;
};

__device__ __forceinline__ void quat_to_matrix_d(float4 q, double (&R)[9]) {  // helper_functions.py:100-117
  const double w = q.x, x = q.y, y = q.z, z = q.w;
  R[0] = 1.0 - 2.0 * y * y - 2.0 * z * z; R[1] = 2.0 * x * y - 2.0 * z * w;       R[2] = 2.0 * x * z + 2.0 * y * w;
  R[3] = 2.0 * x * y + 2.0 * z * w;       R[4] = 1.0 - 2.0 * x * x - 2.0 * z * z; R[5] = 2.0 * y * z - 2.0 * x * w;
  R[6] = 2.0 * x * z - 2.0 * y * w;       R[7] = 2.0 * y * z + 2.0 * x * w;       R[8] = 1.0 - 2.0 * x * x - 2.0 * y * y;
}

// Camera.update, components.py:501-503: position = p + R rel_pos, rotation = R rel_rot.   pose[e] = {R_cam[9], t[3]}
__global__ void camera_update_kernel(const __grid_constant__ CamK k, const float4* state, long long n, long long stride,
                                     double* pose) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double R[9];
  quat_to_matrix_d(state[2 * stride + e], R);
  const float4 p = state[e];
  double* o = pose + 12 * e;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      o[3 * i + j] = R[3 * i] * k.rel_rot[j] + R[3 * i + 1] * k.rel_rot[3 + j] + R[3 * i + 2] * k.rel_rot[6 + j];
  const double pp[3] = {(double)p.x, (double)p.y, (double)p.z};
#pragma unroll
  for (int i = 0; i < 3; ++i)
    o[9 + i] = pp[i] + (R[3 * i] * k.rel_pos[0] + R[3 * i + 1] * k.rel_pos[1] + R[3 * i + 2] * k.rel_pos[2]);
}

// The same update from explicit float64 poses (Camera.update(drone_position, drone_rotation_matrix) with arrays):
// pos double[n][3], R double[n][9] row-major.
__global__ void camera_update_pose_kernel(const __grid_constant__ CamK k, const double* pos, const double* Rm, long long n,
                                          double* pose) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const double* R = Rm + 9 * e;
  double* o = pose + 12 * e;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      o[3 * i + j] = R[3 * i] * k.rel_rot[j] + R[3 * i + 1] * k.rel_rot[3 + j] + R[3 * i + 2] * k.rel_rot[6 + j];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    o[9 + i] = pos[3 * e + i] + (R[3 * i] * k.rel_pos[0] + R[3 * i + 1] * k.rel_pos[1] + R[3 * i + 2] * k.rel_pos[2]);
}

// K [R t]^-1 applied to a world point (components.py:532-536, :558-568): camera coordinates c = R^T (p - t), then
// u = fx cx' ... ; returns false when the point is not in front of the camera (depth <= 0).
__device__ __forceinline__ bool project_point(const CamK& k, const double* pose, double x, double y, double z, int& px, int& py,
                                              double& depth) {
  const double dx = x - pose[9], dy = y - pose[10], dz = z - pose[11];
  const double c0 = pose[0] * dx + pose[3] * dy + pose[6] * dz;
  const double c1 = pose[1] * dx + pose[4] * dy + pose[7] * dz;
  const double c2 = pose[2] * dx + pose[5] * dy + pose[8] * dz;
  depth = c2;
  if (!(c2 > 0.0)) return false;
  const double u = (k.fx * c0 + k.cx * c2) / c2, v = (k.fy * c1 + k.cy * c2) / c2;
  // astype(int): truncation toward zero; saturate far-away values so the cast is defined
  px = (int)fmin(fmax(u, -2.0e9), 2.0e9);
  py = (int)fmin(fmax(v, -2.0e9), 2.0e9);
  return true;
}

// Camera.pruned_objects_list, components.py:584-599: an object survives if some corner of its 3-D box is in front of
// the camera and the 2-D box of those corners overlaps the frame.  box[o] = {lo[3], hi[3]} (helper_functions.py:120-136)
__device__ __forceinline__ bool object_visible(const CamK& k, const double* pose, const double* box, const double* off) {
  int lo_x = 0x7fffffff, lo_y = 0x7fffffff, hi_x = -0x7fffffff - 1, hi_y = -0x7fffffff - 1;
  bool any = false;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const double x = box[(c & 4) ? 3 : 0] + (off ? off[0] : 0.0);
    const double y = box[(c & 1) ? 4 : 1] + (off ? off[1] : 0.0);
    const double z = box[(c & 2) ? 5 : 2] + (off ? off[2] : 0.0);
    int px, py;
    double d;
    if (project_point(k, pose, x, y, z, px, py, d)) {
      any = true;
      lo_x = min(lo_x, px); hi_x = max(hi_x, px);
      lo_y = min(lo_y, py); hi_y = max(hi_y, py);
    }
  }
  return any && hi_x > 0 && hi_y > 0 && lo_x < k.W && lo_y < k.H;
}

// byte value of a depth (components.py:626-628): clip to [0, max_depth], 255 (1 - z / max_depth), truncate
__device__ __forceinline__ unsigned depth_byte(double z, double max_depth) {
  const double c = fmin(fmax(z, 0.0), max_depth);
  return (unsigned)(255.0 * (1.0 - c / max_depth));
}

__global__ void camera_prune_kernel(const __grid_constant__ CamK k, const double* pose, long long n, const double* boxes,
                                    int n_objects, const double* obj_offset, unsigned char* keep) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * n_objects) return;
  const long long e = i / n_objects;
  const int o = (int)(i - e * n_objects);
  keep[i] = object_visible(k, pose + 12 * e, boxes + 6 * o, obj_offset ? obj_offset + 3 * i : nullptr) ? 1 : 0;
}

// Depth splat, components.py:614-629 (max_depth > 0) or binary splat :601-612 (max_depth <= 0).  The reference keeps
// the nearest depth per pixel and maps it to a byte with a non-increasing function, so the image is the per-pixel
// MAXIMUM of the byte values of the points that land there -- one 8-bit atomic max per visible point into a frame
// the host zeroed beforehand.  points[p] = {x, y, z, object index}; grid = (point tiles, envs).
__global__ void camera_splat_kernel(const __grid_constant__ CamK k, const double* pose, const double4* points, int n_points,
                                    int n_objects, const double* obj_offset, const unsigned char* keep, double max_depth,
                                    unsigned char* image) {
  const long long e = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_points) return;
  const double4 q = points[p];
  const int o = (int)q.w;
  if (!keep[e * n_objects + o]) return;
  const double* off = obj_offset ? obj_offset + 3 * (e * n_objects + o) : nullptr;
  int px, py;
  double depth;
  if (!project_point(k, pose + 12 * e, q.x + (off ? off[0] : 0.0), q.y + (off ? off[1] : 0.0), q.z + (off ? off[2] : 0.0), px,
                     py, depth))
    return;
  if (px < 0 || px >= k.W || py < 0 || py >= k.H) return;
  const unsigned v = max_depth > 0.0 ? depth_byte(depth, max_depth) : 1u;
  if (v == 0u) return;
  const long long idx = (e * k.H + py) * (long long)k.W + px;
  unsigned* word = reinterpret_cast<unsigned*>(image + (idx & ~3LL));
  const unsigned shift = (unsigned)(idx & 3LL) * 8u;
  unsigned old = *word;
  while (((old >> shift) & 0xffu) < v) {
    const unsigned assumed = old;
    old = atomicCAS(word, assumed, (assumed & ~(0xffu << shift)) | (v << shift));
    if (old == assumed) break;
  }
}

// Mean pixel of the target as the reference extracts it (simulator.py:104-108): the mean (x, y) over the DISTINCT
// non-zero pixels of render_depth_image([target]).  One CTA per env; the frame is a bitmap in shared memory
// (W*H bits), so the full byte image is never materialised.
__global__ void camera_target_pixel_kernel(const __grid_constant__ CamK k, const double* pose, const double4* points,
                                           int n_points, int n_objects, const double* boxes, const double* obj_offset,
                                           double max_depth, double* pixel, unsigned char* seen) {
  extern __shared__ unsigned bitmap[];
  __shared__ unsigned char keep_s[64];
  __shared__ unsigned long long acc[3];
  const long long e = blockIdx.x;
  const int words = (k.W * k.H + 31) / 32;
  for (int i = threadIdx.x; i < words; i += blockDim.x) bitmap[i] = 0u;
  if (threadIdx.x < 3) acc[threadIdx.x] = 0ull;
  for (int o = threadIdx.x; o < n_objects; o += blockDim.x)
    keep_s[o] = object_visible(k, pose + 12 * e, boxes + 6 * o, obj_offset ? obj_offset + 3 * (e * n_objects + o) : nullptr);
  __syncthreads();
  for (int p = threadIdx.x; p < n_points; p += blockDim.x) {
    const double4 q = points[p];
    const int o = (int)q.w;
    if (!keep_s[o]) continue;
    const double* off = obj_offset ? obj_offset + 3 * (e * n_objects + o) : nullptr;
    int px, py;
    double depth;
    if (!project_point(k, pose + 12 * e, q.x + (off ? off[0] : 0.0), q.y + (off ? off[1] : 0.0), q.z + (off ? off[2] : 0.0),
                       px, py, depth))
      continue;
    if (px < 0 || px >= k.W || py < 0 || py >= k.H) continue;
    if (depth_byte(depth, max_depth) == 0u) continue;
    const int bit = py * k.W + px;
    atomicOr(&bitmap[bit >> 5], 1u << (bit & 31));
  }
  __syncthreads();
  unsigned long long cnt = 0, sx = 0, sy = 0;
  for (int i = threadIdx.x; i < words; i += blockDim.x) {
    unsigned w = bitmap[i];
    while (w) {
      const int b = __ffs(w) - 1;
      w &= w - 1;
      const int bit = i * 32 + b;
      cnt += 1; sx += (unsigned)(bit % k.W); sy += (unsigned)(bit / k.W);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
  }
  if ((threadIdx.x & 31) == 0 && cnt) {
    atomicAdd(&acc[0], cnt); atomicAdd(&acc[1], sx); atomicAdd(&acc[2], sy);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long c = acc[0];
    seen[e] = c ? 1 : 0;
    pixel[2 * e] = c ? (double)acc[1] / (double)c : 0.0;
    pixel[2 * e + 1] = c ? (double)acc[2] / (double)c : 0.0;
  }
}

// Camera.pixel2direction, components.py:505-526, frames: 0 world, 1 drone, 2 camera.
__device__ __forceinline__ void pixel_ray(const CamK& k, const double* camR, double u, double v, int frame, double (&d)[3]) {
  const double r[3] = {(u - k.cx) / k.fx, (v - k.cy) / k.fy, 1.0};  // K^-1 [u v 1]
  if (frame == 2) {
    d[0] = r[0]; d[1] = r[1]; d[2] = r[2];
  } else {
    const double* M = frame == 0 ? camR : k.rel_rot;
#pragma unroll
    for (int i = 0; i < 3; ++i) d[i] = M[3 * i] * r[0] + M[3 * i + 1] * r[1] + M[3 * i + 2] * r[2];
  }
  const double inv = 1.0 / sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  d[0] *= inv; d[1] *= inv; d[2] *= inv;
}

__global__ void camera_rays_kernel(const __grid_constant__ CamK k, const double* pose, long long n, const double* pixel,
                                   int frame, double* dir) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  double d[3];
  pixel_ray(k, pose + 12 * e, pixel[2 * e], pixel[2 * e + 1], frame, d);
  dir[3 * e] = d[0]; dir[3 * e + 1] = d[1]; dir[3 * e + 2] = d[2];
}

// ---------------------------------------------------------------------------------------------------------------
// Drone.calculate_needed_force_orientation (components.py:258-304) with its components.PID (:43-54).
// pid[e] = {integral, prev_derivative, previous_error, is_first}.  Outputs the rotation to apply the force in
// (columns x, y, force, normalised), its quaternion for fpv_drone_io_t.override_q, and |force|.
// Envs with seen[e] == 0 are left alone: NaN thrust = "no override for this env" (fpv_drone_io_t.override_thrust).
// ---------------------------------------------------------------------------------------------------------------
struct AutopilotK {
  double mass, dt, vdrag_coef, vlift_coef, tof_dist, keep_distance, uwb_max;
  double kP, kI, kD, integral_clip, min_out, max_out, dtr;
  int ref_frame;  // 0 world, 1 drone
  int mode;       // 0 level, 1 frontarget
  double max_force;  // max_throttle_in_force: the limit of point_and_shoot's loop (components.py:355)
  int max_iter;
};

__device__ __forceinline__ void cross3(const double (&a)[3], const double (&b)[3], double (&c)[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// VARIANT 0: calculate_needed_force_orientation (:258-304).  VARIANT 1: point_and_shoot (:312-381) -- the pixel is
// shifted by the virtual-target part of `action` (:322-323), the PID regulates the pixel ROW against the on-screen
// target row (:348-350), the virtual lift scales with the sink rate (:345), and the multiplier is walked down until the
// force fits max_throttle_in_force (:355-363; the reference's loop has no exit when drag + lift + gravity alone exceed the
// limit -- capped at max_iter passes here).
template <int VARIANT>
__global__ void autopilot_kernel(const __grid_constant__ AutopilotK a, const __grid_constant__ CamK k, const float4* state,
                                 long long n, long long stride, const double* pixel, const unsigned char* seen,
                                 const double* target_pos, const double* target_radius, const double* action, double* pid,
                                 float* rot_out, float4* quat_out, float* force_out, double* pixel_out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (seen && !seen[e]) {
    if (force_out) force_out[e] = __int_as_float(0x7fc00000);
    if (quat_out) quat_out[e] = make_float4(1.f, 0.f, 0.f, 0.f);
    if (rot_out) {
#pragma unroll
      for (int i = 0; i < 9; ++i) rot_out[9 * e + i] = (i % 4 == 0) ? 1.f : 0.f;
    }
    return;
  }
  double R[9], camR[9];
  quat_to_matrix_d(state[2 * stride + e], R);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      camR[3 * i + j] = R[3 * i] * k.rel_rot[j] + R[3 * i + 1] * k.rel_rot[3 + j] + R[3 * i + 2] * k.rel_rot[6 + j];
  const float4 pf = state[e], vf = state[stride + e];
  const double pos[3] = {(double)pf.x, (double)pf.y, (double)pf.z};
  const double vel[3] = {(double)vf.x, (double)vf.y, (double)vf.z};
  double dir[3];
  double pu = pixel[2 * e], pv = pixel[2 * e + 1];
  if (VARIANT == 1) {  // virtual target relative to the real one, :322-323
    pu += action[4 * e + 2] * (0.5 * k.W);
    pv += action[4 * e + 3] * (0.5 * k.H);
    if (pixel_out) { pixel_out[2 * e] = pu; pixel_out[2 * e + 1] = pv; }
  }
  pixel_ray(k, camR, pu, pv, 0, dir);  // :268 / :332 (always the world-frame ray)
  const double speed = sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
  double grav[3] = {0.0, 0.0, -9.81 * a.mass};  // kinematics.gravity_vector(mass, g=9.81), :271
  double v[3] = {vel[0], vel[1], vel[2]};
  if (a.ref_frame == 1) {  // :274-277
    const double g0[3] = {grav[0], grav[1], grav[2]};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      grav[i] = R[3 * i] * g0[0] + R[3 * i + 1] * g0[1] + R[3 * i + 2] * g0[2];
      v[i] = R[3 * i] * vel[0] + R[3 * i + 1] * vel[1] + R[3 * i + 2] * vel[2];
    }
  }
  const double cosang = (v[0] / speed) * dir[0] + (v[1] / speed) * dir[1] + (v[2] / speed) * dir[2];
  const double dscale = -(cosang - 1.0) / 2.0;
  const bool low = pos[2] < a.tof_dist;  // virtual lift near the ground, :287
  const double lift_gain = VARIANT == 1 ? -fmin(vel[2], 0.0) : 1.0 + fabs(vel[2]);  // :345 / :287
  const double lift = low ? -(a.tof_dist - pos[2]) * a.vlift_coef * lift_gain : 0.0;
  double* ps = pid + 4 * e;
  double err;
  if (VARIANT == 1) {
    // on-screen target row, convert_action2position :383-387 (astype(int): truncation), PID on the pixel row :350
    const double row = (double)(long long)(0.5 * k.H * (1.0 + action[4 * e + 1]));
    err = pv - row;
  } else {
    // measured distance and the force-multiplier PID, :288-291 (+ Target.calculate_distance :770-771)
    const double tx = pos[0] - target_pos[3 * e], ty = pos[1] - target_pos[3 * e + 1], tz = pos[2] - target_pos[3 * e + 2];
    const double dist = fmin(sqrt(tx * tx + ty * ty + tz * tz) - target_radius[e], a.uwb_max);
    err = dist - a.keep_distance;
  }
  const double integ = fmin(fmax(0.99 * ps[0] + err * a.dt, -a.integral_clip), a.integral_clip);
  double der = fmin(fmax((ps[3] != 0.0 ? 0.0 : 1.0) * (err - ps[2]) / a.dt, -1.0), 1.0);
  der = (1.0 - a.dtr) * ps[1] + a.dtr * der;
  ps[0] = integ; ps[1] = der; ps[2] = err; ps[3] = 0.0;
  double mult = fmin(fmax(a.kP * err + a.kI * integ + a.kD * der, a.min_out), a.max_out);
  double rest[3], f[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {  // :293 / :352
    rest[i] = a.vdrag_coef * (dscale * -v[i] * speed) + lift * grav[i] - grav[i];
    f[i] = mult * dir[i] + rest[i];
  }
  double fn = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
  if (VARIANT == 1) {  // :353-363
    double crit = 0.9999;
    for (int it = 0; it < a.max_iter && fn > a.max_force; ++it) {
      mult = fmin(fmax(mult * crit, a.min_out), a.max_out);
#pragma unroll
      for (int i = 0; i < 3; ++i) f[i] = mult * dir[i] + rest[i];
      fn = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
      crit = a.max_force / fn;
    }
  }
  double yv[3], xv[3];
  if (a.mode == 0) cross3(f, grav, yv); else cross3(f, dir, yv);  // :295-300
  cross3(yv, f, xv);
  const double nx = 1.0 / sqrt(xv[0] * xv[0] + xv[1] * xv[1] + xv[2] * xv[2]);
  const double ny = 1.0 / sqrt(yv[0] * yv[0] + yv[1] * yv[1] + yv[2] * yv[2]);
  float m[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {  // columns: x, y, force (:303-304)
    m[3 * i] = (float)(xv[i] * nx); m[3 * i + 1] = (float)(yv[i] * ny); m[3 * i + 2] = (float)(f[i] / fn);
  }
  if (rot_out) {
#pragma unroll
    for (int i = 0; i < 9; ++i) rot_out[9 * e + i] = m[i];
  }
  if (quat_out) quat_out[e] = matrix_to_quat(m);
  if (force_out) force_out[e] = (float)fn;
}

}  // namespace fpv
