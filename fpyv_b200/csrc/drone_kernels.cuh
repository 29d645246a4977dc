// Mode A: the reference `Drone.step` (src/utils/components.py:220-248) for a batch of envs.
// One thread owns L envs (1 = float, 2 = packed F2), takes their state once per control step (4 x 128-bit per env),
// runs K substeps entirely in registers and stores the state once.
//
// Attitude is carried as a unit quaternion q = [w,x,y,z] (the reference's own convention,
// src/utils/helper_functions.py:65-117) instead of the 3x3 matrix the reference mutates: R(q1 (x) q2) = R(q1) R(q2),
// so the reference's  R <- R E^T E^T  (kinematics.py:27-30 applied twice, components.py:216-218) is
// q <- q (x) conj(qE)^2  with qE the quaternion of E = Rz Ry Rx.  Same rotation, 4 floats instead of 9 in HBM and
// ~30 fewer FP32 operations per substep.  q is re-normalised once per control step.
#pragma once
#include "../../include/fpv_api.h"
#include "vec.cuh"
#include "ring_kernels.cuh"
#include "misc_kernels.cuh"

namespace fpv {

// Launch-constant parameters, derived on the host in double precision (fpv_api.cu) and passed by value
// (kernel-parameter constant bank => uniform loads, no __constant__ global state).
struct DroneK {
  float dt;
  int substeps;
  float one_minus_rtr;      // 1 - rates_transition_rate            components.py:187-188
  float max_rates;
  float rtr;
  float ttr;
  float one_minus_ttr;      // components.py:192-193
  float kd0, kd_a, kd_b;    // k_drag[0], k_drag[1]-k_drag[0], k_drag[2]-k_drag[0]   (kinematics.py:36)
  float kd_max;             // max |k_drag[i]| * 1.001: bound on the drag force per |u|^2
  float neg_motor_xy[4][2]; // -motor offsets                       components.py:123-125
  float motor_xy[4][2];
  float motor_radius;
  float arm_reach;          // max |motor offset| + motor_radius + margin: beyond it an obstacle cannot touch any motor
  float bound[4];           // a sphere (centre xyz, radius) that contains every obstacle of the list: broad-phase culling
  float spring_k, spring_c; // components.py:198
  float poly[4];            // throttle% -> N, high->low; evaluated at 100*(x+1)/2   components.py:136
  float wind[3];
  float grav_force_z;       // -g*m                                  kinematics.py:41-45
  float inv_mass;
  float dt_over_mass;       // dt / m
  float half_ang_scale;     // 0.5 * deg2rad * dt                    kinematics.py:29
  float inv_half_ang_scale;
  float lut_scale;          // (lut_n-1)/2
  int lut_n;
  unsigned flags;
  int n_objects;
  fpv_object_t objects[FPV_MAX_OBJECTS];
};

struct DroneIO {
  float4* state;
  long long n, stride;
  const float4* actions;
  const void* sticks;          // raw stick readings instead of actions (FPV_STICKS_U16: uint16[n][4]; FPV_STICKS_CRSF: 6 B per env),
  StickK stick;                // calibrated in registers by the step itself; device memory or pinned host memory
  const float4* wind_env;
  const float* lut;
  unsigned char* done;
  unsigned* done_bits;           // [ceil(n / 32)]: the same flags as a bitmask (bit e % 32 of word e / 32), or null
  float4* acc_out;
  const float4* reset_state;
  const float4* override_q;      // [n]: rotation override as a quaternion (w,x,y,z)
  const float* override_thrust;  // [n]
  fpv_stats_t* stats;
  unsigned* work;              // [0] next-chunk counter, [1] finished-warp counter (dynamic scheduling), or null
  unsigned* chunk_epoch;       // [n_chunks] per-chunk step count (chained launches), or null
  unsigned epoch;              // value chunk_epoch[] holds before this launch; the launch publishes epoch + 1
  unsigned cta_cap;            // host only: cap on resident CTAs per SM (0 = none)
  unsigned* err;               // [0] += 1 when a chained wait timed out (fpv_drone_io_t.work[16]), or null
  unsigned long long* trace;
};

// Obstacle signed distance and contact normal for one motor point (GENERAL path only; warp-uniform object loop).
// The distance is evaluated for every motor of every obstacle in reach; the normal (divisions) only for motors that
// actually touch, which is rare.
template <class V>
__device__ __forceinline__ V object_distance(const fpv_object_t& o, V px, V py, V pz) {
  if (o.kind == FPV_OBJ_SPHERE) {  // Target.calculate_distance, components.py:773-774
    const V dx = px - S<V>(o.x), dy = py - S<V>(o.y), dz = pz - S<V>(o.z);
    return vsqrt_fast(vfma(dx, dx, vfma(dy, dy, dz * dz))) - S<V>(o.a);
  }
  // Cylinder.calculate_distance, components.py:710-716
  const V dx = px - S<V>(o.x), dy = py - S<V>(o.y);
  const V d2 = vsqrt_fast(vfma(dx, dx, dy * dy)) - S<V>(o.a);
  const V top = S<V>(o.z + o.b);
  const auto inside = vand(vlt(S<V>(o.z), pz), vlt(pz, top));
  if (!vany(vnot(inside))) return d2;   // every env of this thread is inside the height band: no cap distance needed
  const V dh = vmin(vabs(pz - S<V>(o.z)), vabs(pz - top));
  return vsel(inside, d2, vsqrt_fast(vfma(d2, d2, dh * dh)));
}
template <class V>
__device__ __forceinline__ void object_normal(const fpv_object_t& o, V px, V py, V pz, V& nx, V& ny, V& nz) {
  if (o.kind == FPV_OBJ_SPHERE) {  // Target.calculate_normal, components.py:776-777
    const V dx = px - S<V>(o.x), dy = py - S<V>(o.y), dz = pz - S<V>(o.z);
    const V r = vsqrt_fast(vfma(dx, dx, vfma(dy, dy, dz * dz)));
    nx = vdiv_fast(dx, r); ny = vdiv_fast(dy, r); nz = vdiv_fast(dz, r);
    return;
  }
  // Cylinder.calculate_normal, components.py:718-729: it first makes the point RELATIVE to the base (:719) and then
  // compares its z with the ABSOLUTE band (:720) and cap heights (:725) -- reproduced as written.
  const V dx = px - S<V>(o.x), dy = py - S<V>(o.y);
  const V rad = vsqrt_fast(vfma(dx, dx, dy * dy));
  const V top = S<V>(o.z + o.b);
  const V qz = pz - S<V>(o.z);
  const auto inside_n = vand(vlt(S<V>(o.z), qz), vlt(qz, top));
  const auto below = vlt(vabs(qz - S<V>(o.z)), vabs(qz - top));
  nx = vsel(inside_n, vdiv_fast(dx, rad), S<V>(0.f));
  ny = vsel(inside_n, vdiv_fast(dy, rad), S<V>(0.f));
  nz = vsel(inside_n, S<V>(0.f), vsel(below, S<V>(-1.f), S<V>(1.f)));
}

template <class V> struct DroneRegs {
  V px, py, pz, vx, vy, vz;
  V qw, qx, qy, qz;
  V pr0, pr1, pr2, pt;
  V ax, ay, az;  // last substep's acceleration
};

// K reference steps for the envs held in `s`.  Returns the OR of the per-step crash flags.
// ANG selects the sin/cos evaluation of the HALF Euler angles h: 0 = full-range sincosf, 1 = |h| <= 0.25 rad
// (degree-7/8 kernels, no range reduction), 2 = |h| <= 0.05 (degree-5/4, truncation < 2e-11), 3 = |h| <= 0.03
// (degree-3/2, truncation < 3.4e-8), 4 = |h| <= 0.008 (qE from its own series, no sin/cos at all).
// WIND = false drops the "+ wind" adds when the launch has no wind at all.
template <class V, int ANG, bool GENERAL, bool WIND>
__device__ __forceinline__ typename Lane<V>::Mask drone_substeps(const DroneK& k, DroneRegs<V>& s, V a0, V a1, V a2,
                                                                   V thrust_target, V wx, V wy, V wz,
                                                                   bool has_override, V o_thrust, V oqw, V oqx, V oqy,
                                                                   V oqz, const fpv_object_t* objs = nullptr) {
  using M = typename Lane<V>::Mask;
  if (GENERAL && objs == nullptr) objs = k.objects;   // kernels without a staged table read the kernel-parameter copy
  // GENERAL: which obstacles can any motor of this thread's envs REACH during this control step?  Decided once, outside
  // the substep loop.  Positions tested by the substeps are x_0 .. x_{K-1} = x_0 + dt * sum of earlier velocities, so an
  // obstacle whose surface is further from x_0 than  arm_reach + K dt V*  (V* = a bound on the speed over the step)
  // fails the per-substep reach test in EVERY substep and contributes exactly nothing: skipping it is bit-identical.
  // V*: |F| <= |thrust| + kmax (|v| + |wind|)^2 + m g + contact springs, the thrust filter is a convex mix of its state and
  // its target, an obstacle / the ground pushes with at most 4 (k r_m + c |v|); with G = 1.5 K dt a(|v0|) + 0.01,
  // K dt a(|v0| + G) <= G proves |v_i| <= |v0| + G for all substeps by induction.  An env whose bound does not close (or an
  // override, which replaces thrust and attitude) keeps every obstacle and is tested substep by substep as before.
  unsigned step_mask = 0u;
  const V zero_h = S<V>(0.f);
  if (GENERAL && k.n_objects > 0) {
    const V Kdt = S<V>((float)k.substeps * k.dt);
    const V v0 = vsqrt_fast(vfma(s.vx, s.vx, vfma(s.vy, s.vy, s.vz * s.vz))) * S<V>(1.0001f);
    const V wn = WIND ? vsqrt_fast(vfma(wx, wx, vfma(wy, wy, wz * wz))) * S<V>(1.0001f) : zero_h;
    V T = vmax(vabs(s.pt), vabs(thrust_target));
    if (has_override) T = vmax(T, vabs(o_thrust));
    const V spring = S<V>(4.f * (float)(k.n_objects + 1) * k.spring_k * k.motor_radius);
    const V damp = S<V>(4.f * (float)(k.n_objects + 1) * fabsf(k.spring_c));
    const V kmax = S<V>(k.kd_max);
    const V base_f = T + spring + S<V>(fabsf(k.grav_force_z));
    auto acc_bound = [&](V vb) {
      const V u = vb + wn;
      return vfma(kmax * u, u, vfma(damp, vb, base_f)) * S<V>(k.inv_mass * 1.001f);
    };
    const V G = vfma(Kdt * acc_bound(v0), S<V>(1.5f), S<V>(0.01f));
    const M closes = vle(Kdt * acc_bound(v0 + G), G);      // false for NaN as well
    const V travel = Kdt * (v0 + G);
    const V reach = S<V>(k.arm_reach) + travel;
    // broad phase: an env outside the sphere that contains ALL obstacles (grown by its reach) can reach none of them
    bool any_candidate = true;
    {
      const V bx = s.px - S<V>(k.bound[0]), by = s.py - S<V>(k.bound[1]), bz = s.pz - S<V>(k.bound[2]);
      const V lim = S<V>(k.bound[3]) + reach;
      const M outside = vand(vlt(lim * lim, vfma(bx, bx, vfma(by, by, bz * bz))), closes);
      any_candidate = vany(vnot(outside));
    }
    for (int o = 0; any_candidate && o < k.n_objects; ++o) {
      const fpv_object_t& ob = k.objects[o];               // warp-uniform index: kernel-parameter constant bank
      const V dx = s.px - S<V>(ob.x), dy = s.py - S<V>(ob.y);
      const V lim = S<V>(ob.a) + reach;
      M near;
      if (ob.kind == FPV_OBJ_SPHERE) {
        const V dz = s.pz - S<V>(ob.z);
        near = vlt(vfma(dx, dx, vfma(dy, dy, dz * dz)), lim * lim);
      } else {
        near = vand(vlt(vfma(dx, dx, dy * dy), lim * lim),
                    vand(vlt(S<V>(ob.z) - reach, s.pz), vlt(s.pz, S<V>(ob.z + ob.b) + reach)));
      }
      near = vor(near, vnot(closes));
      step_mask |= vany(near) ? (1u << o) : 0u;
    }
    if (has_override) step_mask = (k.n_objects >= 32) ? 0xffffffffu : ((1u << k.n_objects) - 1u);
  }

  // No obstacle in reach of ANY env of this warp during this control step, and the reference's own ground configuration:
  // the whole warp runs the hot loop.  The general loop below computes the very same bits for an env without contact (the
  // obstacle forces are added as exact zeros), so an env's result does not depend on which path its warp took.
  if (GENERAL) {
    const bool simple = !has_override && (k.flags & FPV_F_GROUND) != 0 && k.spring_c == 0.f;
    if (simple && !__any_sync(__activemask(), step_mask != 0u))
      return drone_substeps<V, ANG, false, WIND>(k, s, a0, a1, a2, thrust_target, wx, wy, wz, false, o_thrust, oqw, oqx, oqy, oqz,
                                                  nullptr);
  }
  // action2force invariants (the action is held for the whole control step), components.py:185-193
  // The rate filter runs on the half Euler angles h_i = rates_i * (deg2rad*dt/2) directly (same linear recurrence,
  // scaled), so no per-substep rescaling is needed; rates are recovered once after the loop.
  const V mr = S<V>(k.max_rates);
  const V cs = S<V>(k.rtr * k.half_ang_scale);
  const V c0 = vmin(vmax(vneg(a0) * mr, vneg(mr)), mr) * cs;
  const V c1 = vmin(vmax(vneg(a1) * mr, vneg(mr)), mr) * cs;
  const V c2 = vmin(vmax(vneg(a2) * mr, vneg(mr)), mr) * cs;
  const V tt = thrust_target * S<V>(k.ttr);
  const V omr = S<V>(k.one_minus_rtr), omt = S<V>(k.one_minus_ttr);
  const V dt = S<V>(k.dt), dt_m = S<V>(k.dt_over_mass);
  V h0 = s.pr0 * S<V>(k.half_ang_scale), h1 = s.pr1 * S<V>(k.half_ang_scale), h2 = s.pr2 * S<V>(k.half_ang_scale);
  {  // s = sqrt(2) q for the duration of the loop (see the R(q) entries below)
    const V r2 = S<V>(1.41421356237f);
    s.qw = s.qw * r2; s.qx = s.qx * r2; s.qy = s.qy * r2; s.qz = s.qz * r2;
  }
  const V zero = S<V>(0.f), one = S<V>(1.f);
  const bool ground = (k.flags & FPV_F_GROUND) != 0;
  M done = vlt(one, zero);  // all false
  V Fx = zero, Fy = zero, Fz = zero;

#pragma unroll(GENERAL ? 1 : 2)
  for (int it = 0; it < k.substeps; ++it) {
    // ---- low-pass filters on rates and thrust, components.py:187-194
    h0 = vfma(h0, omr, c0); h1 = vfma(h1, omr, c1); h2 = vfma(h2, omr, c2);
    V th = vfma(s.pt, omt, tt);
    s.pt = th;
    if (GENERAL && has_override) {  // components.py:230-232; a NaN thrust leaves that env on its stick command
      const V r2 = S<V>(1.41421356237f);
      const M ovr = vle(o_thrust, o_thrust);
      s.qw = vsel(ovr, oqw * r2, s.qw); s.qx = vsel(ovr, oqx * r2, s.qx);
      s.qy = vsel(ovr, oqy * r2, s.qy); s.qz = vsel(ovr, oqz * r2, s.qz);
      th = vsel(ovr, o_thrust, th);
    }
    // ---- the entries of R(q) this step reads: columns 1 and 2 and R[2][0]  (helper_functions.py:100-117).
    //      The loop carries s = sqrt(2) q, so every "2 q_a q_b" of the matrix is the plain product s_a s_b
    //      (the Hamilton update is linear in q, the scale rides along; undone by the normalisation after the loop).
    const V nqw = vneg(s.qw);
    const V xz = s.qx * s.qz, yz = s.qy * s.qz, xy = s.qx * s.qy;
    const V r02 = vfma(s.qw, s.qy, xz), r20 = vfma(nqw, s.qy, xz);
    const V r21 = vfma(s.qw, s.qx, yz), r12 = vfma(nqw, s.qx, yz);
    const V r01 = vfma(nqw, s.qz, xy);
    const V tx = vfma(vneg(s.qx), s.qx, one);  // 1 - 2x^2
    const V r11 = vfma(vneg(s.qz), s.qz, tx), r22 = vfma(vneg(s.qy), s.qy, tx);
    // ---- drag + thrust + gravity (components.py:242).  calculate_drag (kinematics.py:33-38) is
    //      R diag(k)|u| R^T u with u = v + wind; with R orthonormal that equals
    //      |u| (k0 u + (k1-k0)(c1.u) c1 + (k2-k0)(c2.u) c2), c_j the columns of R -- and the thrust
    //      R[:,2]*th (kinematics.py:48-49) rides on the c2 coefficient.
    V ux = s.vx, uy = s.vy, uz = s.vz;
    if (WIND) { ux = ux + wx; uy = uy + wy; uz = uz + wz; }
    const V nrm = vsqrt_fast(vfma(ux, ux, vfma(uy, uy, uz * uz)));
    const V d1 = vfma(r01, ux, vfma(r11, uy, r21 * uz));
    const V d2 = vfma(r02, ux, vfma(r12, uy, r22 * uz));
    const V ks = S<V>(k.kd0) * nrm;
    const V g1 = (S<V>(k.kd_a) * nrm) * d1;
    const V g2 = vfma(S<V>(k.kd_b) * nrm, d2, th);
    Fx = vfma(g2, r02, vfma(g1, r01, ks * ux));
    Fy = vfma(g2, r12, vfma(g1, r11, ks * uy));
    Fz = vfma(g2, r22, vfma(g1, r21, vfma(ks, uz, S<V>(k.grav_force_z))));
    // ---- motors, collisions, crash test (components.py:235-239, :198-214)
    M crashed;
    if (!GENERAL) {
      // Hot path = the reference's own configuration: ground plane in the object list, undamped spring
      // (components.py:198 passes damping_constant=0); anything else is routed to the GENERAL kernel by the host.
      // distance = z, normal = +z (components.py:674-680); only z of M_rel @ R^T matters.  t_m = motor_radius - z_m
      // is the spring compression; spring_force (kinematics.py:56-59) per motor is k*t_m where t_m > 0; a motor
      // below the plane (t_m > radius) is a crash with NO force (:207-210).
      const V h = S<V>(k.motor_radius) - s.pz;
      V tmax, pen_sum;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const V t = vfma(S<V>(k.neg_motor_xy[m][0]), r20, vfma(S<V>(k.neg_motor_xy[m][1]), r21, h));
        tmax = m == 0 ? t : vmax(tmax, t);
        pen_sum = m == 0 ? vmax(t, zero) : pen_sum + vmax(t, zero);
      }
      crashed = vlt(S<V>(k.motor_radius), tmax);
      Fz = vfma(S<V>(k.spring_k), vsel(crashed, zero, pen_sum), Fz);
    } else {
      // General path: obstacles (sphere / cylinder SDFs), damped contact spring, optional ground.  Reference order:
      // the extra objects first, the ground plane last; the first object any motor penetrates raises `done` and ends
      // the collision pass with the forces gathered so far (components.py:205-210).
      crashed = vlt(one, zero);
      // (1) reach tests for ALL obstacles, branch-free (independent, so their latencies overlap): every motor lies
      //     within the arm length of the drone's centre, so an obstacle whose surface is further away than
      //     arm + motor_radius contributes exactly nothing and skipping it leaves the result bit-identical.
      unsigned near_mask = 0u;
      if (step_mask) {
        const V reach = S<V>(k.arm_reach);
        for (unsigned rem = step_mask; rem;) {
          const int o = __ffs(rem) - 1;
          rem &= rem - 1u;
          const fpv_object_t ob = objs[o];
          const V dx = s.px - S<V>(ob.x), dy = s.py - S<V>(ob.y);
          const V lim = S<V>(ob.a) + reach;
          M near;
          if (ob.kind == FPV_OBJ_SPHERE) {
            const V dz = s.pz - S<V>(ob.z);
            near = vlt(vfma(dx, dx, vfma(dy, dy, dz * dz)), lim * lim);
          } else {  // cylinder: horizontally close AND inside the height band grown by the reach
            near = vand(vlt(vfma(dx, dx, dy * dy), lim * lim),
                        vand(vlt(S<V>(ob.z) - reach, s.pz), vlt(s.pz, S<V>(ob.z + ob.b) + reach)));
          }
          near_mask |= vany(near) ? (1u << o) : 0u;
        }
      }
      // (2) the obstacles in reach (rare): 4 motor distances each; normals and spring forces only where a motor is
      //     inside the contact shell
      if (near_mask) {
        // Code size matters here: this block is rarely executed but, fully unrolled (4 motors x sphere / cylinder distance and
        // normal code, twice if the substep loop is unrolled), it made the kernel ~130 KB of SASS and the no-contact path
        // stalled on instruction fetch.  One copy: the loop over the motors is NOT unrolled and every motor position is
        // recomputed from the (warp-uniform) offset table.
        const V r00 = vfma(vneg(s.qy), s.qy, vfma(vneg(s.qz), s.qz, one)), r10 = vfma(s.qw, s.qz, xy);
        V cfx = zero, cfy = zero, cfz = zero;
        while (near_mask) {
          const int o = __ffs(near_mask) - 1;
          near_mask &= near_mask - 1u;
          const fpv_object_t ob = objs[o];
          M hit = vlt(one, zero);
          V ox = zero, oy = zero, oz = zero;
#pragma unroll 1
          for (int m = 0; m < 4; ++m) {
            const V mx_ = S<V>(k.motor_xy[m][0]), my_ = S<V>(k.motor_xy[m][1]);
            const V wxm = vfma(mx_, r00, vfma(my_, r01, s.px));
            const V wym = vfma(mx_, r10, vfma(my_, r11, s.py));
            const V wzm = vfma(mx_, r20, vfma(my_, r21, s.pz));
            const V dm = object_distance<V>(ob, wxm, wym, wzm);
            hit = vor(hit, vlt(dm, zero));
            const V pen = dm - S<V>(k.motor_radius);
            // an env this object has already penetrated (an earlier motor, or this one) gets nothing from it (`live`
            // below), so it does not ask for the normals
            const M act = vand(vlt(pen, zero), vnot(hit));
            if (vany(act)) {   // spring force of a motor inside the contact shell, :207-214
              V nx, ny, nz;
              object_normal<V>(ob, wxm, wym, wzm, nx, ny, nz);
              const V vn = vfma(s.vx, nx, vfma(s.vy, ny, s.vz * nz));
              const V f = vneg(vfma(S<V>(k.spring_k), pen, S<V>(k.spring_c) * vn));
              ox = ox + vsel(act, f * nx, zero);
              oy = oy + vsel(act, f * ny, zero);
              oz = oz + vsel(act, f * nz, zero);
            }
          }
          hit = vand(hit, vnot(crashed));
          crashed = vor(crashed, hit);
          const M live = vnot(crashed);       // an object any motor penetrates contributes nothing and ends the pass (:205-210)
          cfx = cfx + vsel(live, ox, zero);
          cfy = cfy + vsel(live, oy, zero);
          cfz = cfz + vsel(live, oz, zero);
        }
        Fx = Fx + cfx; Fy = Fy + cfy; Fz = Fz + cfz;   // (only where something was in reach: the no-contact path adds nothing)
      }
      // (3) the ground plane, last in the list (distance = z, normal = +z, components.py:674-680), and the crash test
      //     on the motor heights that holds with or without it (:239)
      if (ground && k.spring_c == 0.f) {   // the reference's own ground configuration: the hot path's formula (same bits)
        const V h = S<V>(k.motor_radius) - s.pz;
        V tmax, pen_sum;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const V t = vfma(S<V>(k.neg_motor_xy[m][0]), r20, vfma(S<V>(k.neg_motor_xy[m][1]), r21, h));
          tmax = m == 0 ? t : vmax(tmax, t);
          pen_sum = m == 0 ? vmax(t, zero) : pen_sum + vmax(t, zero);
        }
        crashed = vor(crashed, vlt(S<V>(k.motor_radius), tmax));
        Fz = vfma(S<V>(k.spring_k), vsel(crashed, zero, pen_sum), Fz);   // the hot path's expression, bit for bit
      } else {
        V mz[4];
        V minz;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          mz[m] = vfma(S<V>(k.motor_xy[m][0]), r20, vfma(S<V>(k.motor_xy[m][1]), r21, s.pz));
          minz = m == 0 ? mz[m] : vmin(minz, mz[m]);
        }
        const M below = vlt(minz, zero);
        if (ground) {
          crashed = vor(crashed, vand(below, vnot(crashed)));
          const M live = vnot(crashed);
          V oz = zero;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const V pen = mz[m] - S<V>(k.motor_radius);
            const V f = vneg(vfma(S<V>(k.spring_k), pen, S<V>(k.spring_c) * s.vz));
            oz = oz + vsel(vlt(pen, zero), f, zero);
          }
          Fz = Fz + vsel(live, oz, zero);
        }
        crashed = vor(crashed, below);  // components.py:239
      }
    }
    done = vor(done, crashed);
    // ---- translation: x += v*dt with the OLD v, then v += (F/m)*dt, kinematics.py:21-22, components.py:243
    s.px = vfma(s.vx, dt, s.px); s.py = vfma(s.vy, dt, s.py); s.pz = vfma(s.vz, dt, s.pz);
    s.vx = vfma(Fx, dt_m, s.vx); s.vy = vfma(Fy, dt_m, s.vy); s.vz = vfma(Fz, dt_m, s.vz);
    // ---- attitude: E = Rz(yaw)Ry(pitch)Rx(roll) of deg2rad(rates)*dt, R <- R E^T E^T
    //      (rotate_body_by_rates kinematics.py:27-30 runs inside update_kinematic_step :23 AND again in
    //      Drone.update components.py:218).  Quaternion form: qE from the half angles, q <- q (x) conj(qE)^2.
    V ew, ex, ey, ez;  // qE = qz (x) qy (x) qx
    if (ANG == 4) {
      // |h| <= 0.008: the four components as their series to O(h^4) (dropped terms < 3.5e-9, far below fp32
      // resolution): e_v = h_i (1 - n/2 + h_i^2/3) -/+ h_j h_k,  e_w = 1 - n/2 + hx hy hz,  n = |h|^2
      const V sx = h0 * h0, sy = h1 * h1, sz = h2 * h2;
      const V g = vfma((sx + sy) + sz, S<V>(-0.5f), one);
      const V third = S<V>(0.333333333f);
      const V yz = h1 * h2, xz = h0 * h2, xy = h0 * h1;
      ex = vfma(h0, vfma(sx, third, g), vneg(yz));
      ey = vfma(h1, vfma(sy, third, g), xz);
      ez = vfma(h2, vfma(sz, third, g), vneg(xy));
      ew = vfma(h0, yz, g);
    } else {
      V sr, cr, sp, cp, sy, cy;
      vsincos<ANG>(h0, sr, cr);
      vsincos<ANG>(h1, sp, cp);
      vsincos<ANG>(h2, sy, cy);
      const V A = cy * cp, B = sy * sp, C = cy * sp, D = sy * cp;
      ew = vfma(A, cr, B * sr);
      ex = vfma(A, sr, vneg(B * cr));
      ey = vfma(C, cr, D * sr);
      ez = vfma(D, cr, vneg(C * sr));
    }
    // p = conj(qE)^2 = (ew^2 - |ev|^2, -2 ew ev)
    const V pw = vfma(ew, ew, vneg(vfma(ex, ex, vfma(ey, ey, ez * ez))));
    const V m2 = ew * S<V>(-2.f);
    const V px_ = m2 * ex, py_ = m2 * ey, pz_ = m2 * ez;
    // q <- q (x) p   (Hamilton product)
    const V nw = vfma(s.qw, pw, vneg(vfma(s.qx, px_, vfma(s.qy, py_, s.qz * pz_))));
    const V nx_ = vfma(s.qw, px_, vfma(s.qx, pw, vfma(s.qy, pz_, vneg(s.qz * py_))));
    const V ny_ = vfma(s.qw, py_, vfma(s.qy, pw, vfma(s.qz, px_, vneg(s.qx * pz_))));
    const V nz_ = vfma(s.qw, pz_, vfma(s.qz, pw, vfma(s.qx, py_, vneg(s.qy * px_))));
    s.qw = nw; s.qx = nx_; s.qy = ny_; s.qz = nz_;
  }
  s.pr0 = h0 * S<V>(k.inv_half_ang_scale); s.pr1 = h1 * S<V>(k.inv_half_ang_scale); s.pr2 = h2 * S<V>(k.inv_half_ang_scale);
  // acceleration of the last substep (Drone.acceleration, components.py:243)
  const V inv_m = S<V>(k.inv_mass);
  s.ax = Fx * inv_m; s.ay = Fy * inv_m; s.az = Fz * inv_m;
  // keep q on the unit sphere (the reference's R is orthonormal to 1e-14; fp32 products drift by ~1e-7 per step)
  const V n2 = vfma(s.qw, s.qw, vfma(s.qx, s.qx, vfma(s.qy, s.qy, s.qz * s.qz)));
  const V rn = vrsqrt_fast(n2);
  s.qw = s.qw * rn; s.qx = s.qx * rn; s.qy = s.qy * rn; s.qz = s.qz * rn;
  return done;
}

// throttle2thrust, components.py:136: cubic in percent = 100*(x+1)/2, or the shared-memory LUT.
template <class V> __device__ __forceinline__ V thrust_poly(const DroneK& k, V x) {
  const V pct = vfma(x, S<V>(50.f), S<V>(50.f));
  V p = vfma(S<V>(k.poly[0]), pct, S<V>(k.poly[1]));
  p = vfma(p, pct, S<V>(k.poly[2]));
  return vfma(p, pct, S<V>(k.poly[3]));
}
__device__ __forceinline__ float thrust_lut1(const DroneK& k, const float* lut, float x) {
  // table sampled uniformly on throttle in [-1,1]; linear interpolation, linear extrapolation outside
  float u = (x + 1.f) * k.lut_scale;
  int i = (int)floorf(u);
  i = max(0, min(i, k.lut_n - 2));
  const float f = u - (float)i;
  const float a = lut[i], b = lut[i + 1];
  return fmaf(f, b - a, a);
}

// Episode statistics: every counter is a RARE event (crash / episode end / non-finite / frozen), so the common
// path is one warp vote and no memory traffic; a warp that saw an event reduces with shuffles and issues one
// fire-and-forget red.global.add.f64 per non-zero counter.  No block barrier.  env_steps = n per launch (added by
// one thread of the grid) minus the frozen envs.
struct TileStats {
  float crash, epi, len, nf, frozen;
};
__device__ __forceinline__ void stats_warp_flush(fpv_stats_t* stats, const TileStats& t) {
  const bool any = (t.crash != 0.f) | (t.nf != 0.f) | (t.frozen != 0.f);
  if (!__any_sync(0xffffffffu, any)) return;
  float v[5] = {t.crash, t.epi, t.len, t.nf, t.frozen};
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  if ((threadIdx.x & 31) == 0) {
    if (v[0] != 0.f) atomicAdd(&stats->crashes, (double)v[0]);
    if (v[1] != 0.f) atomicAdd(&stats->episodes, (double)v[1]);
    if (v[2] != 0.f) atomicAdd(&stats->episode_len_sum, (double)v[2]);
    if (v[3] != 0.f) atomicAdd(&stats->nonfinite, (double)v[3]);
    if (v[4] != 0.f) atomicAdd(&stats->env_steps, -(double)v[4]);
  }
}

// One tile worth of work for this thread: unpack L envs from their float4 rows, run the substeps in registers,
// episode bookkeeping, stores.  q[p][l] = plane p of slot l; slot l is env base + l*SLOT_STRIDE (ei[l] = the same
// index clamped to n-1, used for the side inputs).
// Per-chunk epilogue hook of drone_tile (see DroneMode below): nothing.
struct NoPost {
  static constexpr bool enabled = false;
  struct Ctx {};
  template <class IO> static __device__ __forceinline__ Ctx begin(const IO&) { return Ctx{}; }
  template <class IO> static __device__ __forceinline__ void stage(const IO&, unsigned char*, int, int) {}
  template <class IO> static __device__ __forceinline__ int bytes(const IO&) { return 0; }
  template <class IO, int L> static __device__ __forceinline__ void prefetch(const IO&, const long long (&)[L], Ctx&) {}
  template <class IO, int L>
  static __device__ __forceinline__ void run(const IO&, const unsigned char*, const float4 (&)[FPV_DRONE_PLANES][L],
                                             const bool (&)[L], const bool (&)[L], const long long (&)[L], Ctx&) {}
  template <class IO> static __device__ __forceinline__ void finish(const IO&, Ctx&) {}
};

// `pre_store` runs between the arithmetic and the first store of the tile (used by the ring kernel to publish the
// PREVIOUS chunk's epoch once its stores have had a whole substep loop to land).  `Post` (NoPost, or the gate-race env
// step) runs after the stores on the FINAL rows of the thread's envs -- what the next step will read -- with all 32 lanes
// converged, so that it may use warp shuffles / votes.
template <class V, int ANG, bool GENERAL, int SLOT_STRIDE, class PreStore = NoHook, class Post = NoPost, class IO = DroneIO>
__device__ __forceinline__ void drone_tile(const DroneK& k, const IO& io, const float* lut_s, const fpv_object_t* objs,
                                           const float4 (&q)[FPV_DRONE_PLANES][Lane<V>::N],
                                           const float4 (&act)[Lane<V>::N], const long long (&ei)[Lane<V>::N],
                                           long long base, TileStats& st, const bool wind_on,
                                           PreStore pre_store = PreStore(), const unsigned char* staged = nullptr,
                                           typename Post::Ctx* post_ctx = nullptr) {
  constexpr int L = Lane<V>::N;
  if (Post::enabled) Post::prefetch(io, ei, *post_ctx);   // the hook's own per-env inputs travel during the substep loop
  DroneRegs<V> s;
  s.px = Pack<V>::x(q[0]); s.py = Pack<V>::y(q[0]); s.pz = Pack<V>::z(q[0]); s.pt = Pack<V>::w(q[0]);
  s.vx = Pack<V>::x(q[1]); s.vy = Pack<V>::y(q[1]); s.vz = Pack<V>::z(q[1]);
  s.qw = Pack<V>::x(q[2]); s.qx = Pack<V>::y(q[2]); s.qy = Pack<V>::z(q[2]); s.qz = Pack<V>::w(q[2]);
  s.pr0 = Pack<V>::x(q[3]); s.pr1 = Pack<V>::y(q[3]); s.pr2 = Pack<V>::z(q[3]);
  s.ax = S<V>(0.f); s.ay = S<V>(0.f); s.az = S<V>(0.f);
  int epi[L];
  float spare[L];
#pragma unroll
  for (int l = 0; l < L; ++l) { epi[l] = __float_as_int(q[1][l].w); spare[l] = q[3][l].w; }

  V wx = S<V>(k.wind[0]), wy = S<V>(k.wind[1]), wz = S<V>(k.wind[2]);
  if (io.wind_env) {
    float4 w[L];
#pragma unroll
    for (int l = 0; l < L; ++l) w[l] = ldg_stream(io.wind_env + ei[l]);
    wx = Pack<V>::x(w); wy = Pack<V>::y(w); wz = Pack<V>::z(w);
  }
  const V a0 = Pack<V>::x(act), a1 = Pack<V>::y(act), a2 = Pack<V>::z(act), a3 = Pack<V>::w(act);
  V target;
  if (k.flags & FPV_F_THRUST_LUT) {
    float t[2];
#pragma unroll
    for (int l = 0; l < L; ++l) t[l] = thrust_lut1(k, lut_s, act[l].w);
    target = Lane<V>::make(t[0], t[L - 1]);
  } else {
    target = thrust_poly<V>(k, a3);
  }
  V o_thrust = S<V>(0.f), oqw = S<V>(1.f), oqx = S<V>(0.f), oqy = S<V>(0.f), oqz = S<V>(0.f);
  const bool has_ovr = GENERAL && io.override_q != nullptr;
  if (has_ovr) {
    float4 oq[L];
    float ot[2];
#pragma unroll
    for (int l = 0; l < L; ++l) { oq[l] = ldg_stream(io.override_q + ei[l]); ot[l] = io.override_thrust[ei[l]]; }
    oqw = Pack<V>::x(oq); oqx = Pack<V>::y(oq); oqy = Pack<V>::z(oq); oqz = Pack<V>::w(oq);
    o_thrust = Lane<V>::make(ot[0], ot[L - 1]);
  }

  typename Lane<V>::Mask done;
  if (wind_on) done = drone_substeps<V, ANG, GENERAL, true>(k, s, a0, a1, a2, target, wx, wy, wz, has_ovr, o_thrust, oqw, oqx, oqy, oqz, objs);
  else done = drone_substeps<V, ANG, GENERAL, false>(k, s, a0, a1, a2, target, wx, wy, wz, has_ovr, o_thrust, oqw, oqx, oqy, oqz, objs);

  pre_store();
  // ---- the done flags as a bitmask: one ballot per slot, one 32-bit store per 32 envs (SLOT_STRIDE == 32: the ring
  //      layout, where slot l of lane t is env chunk*64 + 32 l + t, i.e. bit t of word chunk*2 + l)
  if (SLOT_STRIDE == 32 && io.done_bits) {
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const long long e = base + (long long)l * SLOT_STRIDE;
      const bool flag = e < io.n && (mask_get(done, l) || epi[l] < 0);
      const unsigned word = __ballot_sync(0xffffffffu, flag);
      if ((threadIdx.x & 31) == 0 && e < io.n) io.done_bits[e >> 5] = word;
    }
  }
  // ---- epilogue per env: episode bookkeeping, freeze / auto-reset, stores
  float4 fin[FPV_DRONE_PLANES][L];   // the rows the next step will read (only materialised when a Post hook consumes them)
  bool crashed_l[L], live_l[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const long long e = base + (long long)l * SLOT_STRIDE;
    crashed_l[l] = false;
    live_l[l] = e < io.n;
    if (Post::enabled) {
#pragma unroll
      for (int p = 0; p < FPV_DRONE_PLANES; ++p) fin[p][l] = q[p][l];
    }
    if (e >= io.n) continue;
    const bool d = mask_get(done, l);
    int ep = epi[l];
    if (ep < 0) {  // frozen after a crash (FPV_F_FREEZE_DONE): state in memory stays as it is, done is sticky
      if (io.done) io.done[e] = 1;
      st.frozen += 1.f;
      crashed_l[l] = true;
      continue;
    }
    crashed_l[l] = d;
    ep += 1;
    float4* const dst = io.state + e;
    if (io.done) io.done[e] = d ? 1 : 0;
    if (io.acc_out)
      stg_stream(io.acc_out + e, make_float4(Lane<V>::get(s.ax, l), Lane<V>::get(s.ay, l), Lane<V>::get(s.az, l), 0.f));
    if (d) {  // rare: crash this control step
      st.crash += 1.f;
      if (k.flags & FPV_F_AUTO_RESET) {  // restart from the reset snapshot
        st.epi += 1.f; st.len += (float)ep;
        const float4* const src = io.reset_state + e;
        float4 v[FPV_DRONE_PLANES];  // all loads in flight before the first store (one round trip, not four)
#pragma unroll
        for (int p = 0; p < FPV_DRONE_PLANES; ++p) v[p] = ldg_stream(src + p * io.stride);
        v[1].w = __int_as_float(0);
#pragma unroll
        for (int p = 0; p < FPV_DRONE_PLANES; ++p) stg_stream(dst + p * io.stride, v[p]);
        if (Post::enabled) {
#pragma unroll
          for (int p = 0; p < FPV_DRONE_PLANES; ++p) fin[p][l] = v[p];
        }
        continue;
      }
      if (k.flags & FPV_F_FREEZE_DONE) {
        st.epi += 1.f; st.len += (float)ep;
        ep = -ep - 1;
      }
    }
    const float px = Lane<V>::get(s.px, l), py = Lane<V>::get(s.py, l), pz = Lane<V>::get(s.pz, l);
    const float vx = Lane<V>::get(s.vx, l), vy = Lane<V>::get(s.vy, l), vz = Lane<V>::get(s.vz, l);
    // NaN/Inf guard on the position: a non-finite velocity or attitude reaches it within one more step
    if (!(fabsf((px + py) + pz) <= 3.0e38f)) st.nf += 1.f;
    const float4 o0 = make_float4(px, py, pz, Lane<V>::get(s.pt, l));
    const float4 o1 = make_float4(vx, vy, vz, __int_as_float(ep));
    const float4 o2 = make_float4(Lane<V>::get(s.qw, l), Lane<V>::get(s.qx, l), Lane<V>::get(s.qy, l), Lane<V>::get(s.qz, l));
    const float4 o3 = make_float4(Lane<V>::get(s.pr0, l), Lane<V>::get(s.pr1, l), Lane<V>::get(s.pr2, l), spare[l]);
    stg_stream(dst, o0);
    stg_stream(dst + io.stride, o1);
    stg_stream(dst + 2 * io.stride, o2);
    stg_stream(dst + 3 * io.stride, o3);
    if (Post::enabled) { fin[0][l] = o0; fin[1][l] = o1; fin[2][l] = o2; fin[3][l] = o3; }
  }
  if (Post::enabled) Post::run(io, staged, fin, crashed_l, live_l, ei, *post_ctx);
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 1 (general path: obstacles / overrides): persistent CTAs, state fetched with 128-bit streaming loads.
// gridDim.x CTAs (a multiple of the SM count chosen by the host) walk the tiles of THREADS*L envs with stride
// gridDim.x.  The motor-curve LUT is staged into shared memory once per CTA.
// ---------------------------------------------------------------------------------------------------------------
template <class V, int ANG, bool GENERAL, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) drone_step_kernel(const __grid_constant__ DroneK k, const DroneIO io) {
  constexpr int L = Lane<V>::N;
  constexpr int TILE = THREADS * L;
  extern __shared__ __align__(16) float lut_dyn[];
  if (k.flags & FPV_F_THRUST_LUT) {
    for (int i = threadIdx.x; i < k.lut_n; i += THREADS) lut_dyn[i] = io.lut[i];
    __syncthreads();
  }
  if (io.stats && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&io.stats->env_steps, (double)io.n);
  const long long n_tiles = (io.n + TILE - 1) / TILE;
  const bool wind_on = io.wind_env != nullptr || k.wind[0] != 0.f || k.wind[1] != 0.f || k.wind[2] != 0.f;
  TileStats st = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // thread t, slot l -> env tile*TILE + l*THREADS + t: every 128-bit access of a warp is one contiguous 512 B run
    const long long base = tile * TILE + threadIdx.x;
    if (base >= io.n) continue;
    long long ei[L];  // a slot past the end re-reads env n-1 and is never stored
#pragma unroll
    for (int l = 0; l < L; ++l) ei[l] = min(base + (long long)l * THREADS, io.n - 1);
    float4 q[FPV_DRONE_PLANES][L];
    float4 act[L];
#pragma unroll
    for (int p = 0; p < FPV_DRONE_PLANES; ++p)
#pragma unroll
      for (int l = 0; l < L; ++l) q[p][l] = ldg_stream(io.state + p * io.stride + ei[l]);
#pragma unroll
    for (int l = 0; l < L; ++l) act[l] = ldg_stream(io.actions + ei[l]);
    drone_tile<V, ANG, GENERAL, THREADS>(k, io, lut_dyn, nullptr, q, act, ei, base, st, wind_on);
  }
  if (io.stats) stats_warp_flush(io.stats, st);
}

// ---------------------------------------------------------------------------------------------------------------
// Kernel 2 (the ring form of the step): ring_step_kernel<DroneMode<...>> (ring_kernels.cuh).  Rows of a chunk = the four
// state planes + the actions; staged table = the motor-curve LUT.  GENERAL = false is the hot path (the reference's own
// configuration: ground plane, undamped spring); GENERAL = true adds obstacles / overrides / damped contact.  `Post` is the
// per-chunk epilogue run on the FINAL rows of the chunk (after auto-reset), all lanes converged: the gate-race env step
// plugs in there (env_kernels.cuh); NoPost for the plain step.
// ---------------------------------------------------------------------------------------------------------------
// STICKS: 0 = row 4 of a chunk are the float4 actions; FPV_STICKS_U16 / FPV_STICKS_CRSF = row 4 are raw stick readings
// (8 / 6 bytes per env), turned into actions in registers with the arithmetic of sticks4_u16_kernel / sticks4_crsf_kernel.
template <class V_, int ANG, bool GENERAL, class Post = NoPost, class IO_ = DroneIO, int STICKS = 0>
struct DroneMode {
  using V = V_;
  using K = DroneK;
  using IO = IO_;
  static constexpr int ROWS = FPV_DRONE_PLANES + 1;
  struct Ctx {
    TileStats st;
    bool wind_on;
    typename Post::Ctx post;
  };
  static __device__ __forceinline__ bool chained(const K& k) { return (k.flags & FPV_F_CHAINED) != 0; }
  static __device__ __forceinline__ constexpr int row_bytes(int r) {
    return r < FPV_DRONE_PLANES ? 16 : (STICKS == 0 ? 16 : (STICKS == FPV_STICKS_U16 ? 8 : 6));
  }
  static __device__ __forceinline__ const void* row_ptr(const IO& io, int r, long long first) {
    if (r < FPV_DRONE_PLANES) return io.state + r * io.stride + first;
    if (STICKS == 0) return io.actions + first;
    return static_cast<const char*>(io.sticks) + first * row_bytes(FPV_DRONE_PLANES);
  }
  static __device__ __forceinline__ int lut_bytes(const K& k) {
    return (k.flags & FPV_F_THRUST_LUT) ? (int)(((size_t)k.lut_n * sizeof(float) + 127) / 128 * 128) : 0;
  }
  static __device__ __forceinline__ void stage(const K& k, const IO& io, unsigned char* smem, int tid, int nthreads) {
    float* lut_s = reinterpret_cast<float*>(smem);
    if (k.flags & FPV_F_THRUST_LUT)
      for (int i = tid; i < k.lut_n; i += nthreads) lut_s[i] = io.lut[i];
    if (GENERAL) {   // obstacle table behind the LUT: indexed PER THREAD in the contact loop (a divergent index into the
                     // kernel-parameter constant bank would serialise every field load)
      float* ob = reinterpret_cast<float*>(smem + lut_bytes(k));
      const float* src = reinterpret_cast<const float*>(k.objects);
      for (int i = tid; i < k.n_objects * (int)(sizeof(fpv_object_t) / sizeof(float)); i += nthreads) ob[i] = src[i];
    }
    Post::stage(io, smem + post_off(k), tid, nthreads);
  }
  static __device__ __forceinline__ int post_off(const K& k) {
    return lut_bytes(k) + (GENERAL ? (k.n_objects * (int)sizeof(fpv_object_t) + 15) / 16 * 16 : 0);
  }
  static __device__ __forceinline__ Ctx begin(const K& k, const IO& io) {
    if (io.stats && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&io.stats->env_steps, (double)io.n);
    Ctx c;
    c.st = TileStats{0.f, 0.f, 0.f, 0.f, 0.f};
    c.wind_on = io.wind_env != nullptr || k.wind[0] != 0.f || k.wind[1] != 0.f || k.wind[2] != 0.f;
    c.post = Post::begin(io);
    return c;
  }
  template <class PreStore>
  static __device__ __forceinline__ void tile(const K& k, const IO& io, const unsigned char* staged,
                                              const float4 (&rows)[ROWS][Lane<V>::N], const long long (&ei)[Lane<V>::N],
                                              long long base, Ctx& c, PreStore pre_store) {
    constexpr int L = Lane<V>::N;
    float4 q[FPV_DRONE_PLANES][L], act[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
#pragma unroll
      for (int p = 0; p < FPV_DRONE_PLANES; ++p) q[p][l] = rows[p][l];
      const float4 a = rows[FPV_DRONE_PLANES][l];
      if (STICKS == 0) {
        act[l] = a;
      } else if (STICKS == FPV_STICKS_U16) {
        const unsigned lo = __float_as_uint(a.x), hi = __float_as_uint(a.y);
        act[l] = sticks4_to_action(io.stick, (float)(lo & 0xffffu), (float)(lo >> 16), (float)(hi & 0xffffu), (float)(hi >> 16));
      } else {
        act[l] = crsf_to_action(io.stick, __float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z));
      }
    }
    const fpv_object_t* objs = GENERAL ? reinterpret_cast<const fpv_object_t*>(staged + lut_bytes(k)) : nullptr;
    drone_tile<V, ANG, GENERAL, 32, PreStore, Post>(k, io, reinterpret_cast<const float*>(staged), objs, q, act, ei, base, c.st,
                                                    c.wind_on, pre_store, staged + post_off(k), &c.post);
  }
  static __device__ __forceinline__ void finish(const K&, const IO& io, Ctx& c) {
    if (io.stats) stats_warp_flush(io.stats, c.st);
    Post::finish(io, c.post);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Kernel 3 (open-loop rollouts): T control steps per launch with the state held in REGISTERS across the steps.
// Every env's state is read once and written once per launch; per control step only its 16-byte action is read and
// its done byte written (17 B/env/step instead of 145).  Arithmetic per control step is drone_substeps() exactly as in
// the step kernels -- the values pass through the same float32 registers they would pass through in memory -- so a
// rollout is BIT-IDENTICAL to T launches of fpv_drone_step, episode bookkeeping, auto-reset and statistics included.
// Warp-chunks of 32*L envs are pulled dynamically; the next step's actions are fetched while the current step computes.
// ---------------------------------------------------------------------------------------------------------------
template <class V, int ANG, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) drone_rollout_kernel(const __grid_constant__ DroneK k, const DroneIO io,
                                                                      const float4* actions_seq, const long long act_stride,
                                                                      const int T, unsigned char* done_seq,
                                                                      const long long done_stride) {
  constexpr int L = Lane<V>::N;
  constexpr int CHUNK = 32 * L;
  constexpr int WARPS = THREADS / 32;
  extern __shared__ __align__(16) float lut_dyn[];
  if (k.flags & FPV_F_THRUST_LUT) {
    for (int i = threadIdx.x; i < k.lut_n; i += THREADS) lut_dyn[i] = io.lut[i];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const long long n_chunks = (io.n + CHUNK - 1) / CHUNK;
  const long long total_warps = (long long)gridDim.x * WARPS;
  const bool wind_on = k.wind[0] != 0.f || k.wind[1] != 0.f || k.wind[2] != 0.f;
  const V wx = S<V>(k.wind[0]), wy = S<V>(k.wind[1]), wz = S<V>(k.wind[2]);
  const V zero = S<V>(0.f);
  if (io.stats && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&io.stats->env_steps, (double)io.n * (double)T);
  TileStats st = {0.f, 0.f, 0.f, 0.f, 0.f};
  long long cur = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5);   // first chunk static, the rest pulled
  while (cur < n_chunks) {
    const long long base = cur * CHUNK + lane;
    long long ei[L];
    bool live[L];
#pragma unroll
    for (int l = 0; l < L; ++l) {
      live[l] = base + (long long)l * 32 < io.n;
      ei[l] = min(base + (long long)l * 32, io.n - 1);   // a slot past the end recomputes env n-1 and is never stored
    }
    float4 q[FPV_DRONE_PLANES][L];
#pragma unroll
    for (int p = 0; p < FPV_DRONE_PLANES; ++p)
#pragma unroll
      for (int l = 0; l < L; ++l) q[p][l] = ldg_stream(io.state + p * io.stride + ei[l]);
    float4 act[L], act_next[L];
#pragma unroll
    for (int l = 0; l < L; ++l) act[l] = ldg_stream(actions_seq + ei[l]);
    DroneRegs<V> s;
    s.px = Pack<V>::x(q[0]); s.py = Pack<V>::y(q[0]); s.pz = Pack<V>::z(q[0]); s.pt = Pack<V>::w(q[0]);
    s.vx = Pack<V>::x(q[1]); s.vy = Pack<V>::y(q[1]); s.vz = Pack<V>::z(q[1]);
    s.qw = Pack<V>::x(q[2]); s.qx = Pack<V>::y(q[2]); s.qy = Pack<V>::z(q[2]); s.qz = Pack<V>::w(q[2]);
    s.pr0 = Pack<V>::x(q[3]); s.pr1 = Pack<V>::y(q[3]); s.pr2 = Pack<V>::z(q[3]);
    s.ax = zero; s.ay = zero; s.az = zero;
    int epi[L];
    float spare[L];
#pragma unroll
    for (int l = 0; l < L; ++l) { epi[l] = __float_as_int(q[1][l].w); spare[l] = q[3][l].w; }

    for (int t = 0; t < T; ++t) {
      if (t + 1 < T) {   // next step's sticks travel while this step computes
#pragma unroll
        for (int l = 0; l < L; ++l) act_next[l] = ldg_stream(actions_seq + (long long)(t + 1) * act_stride + ei[l]);
      }
      const V a0 = Pack<V>::x(act), a1 = Pack<V>::y(act), a2 = Pack<V>::z(act), a3 = Pack<V>::w(act);
      V target;
      if (k.flags & FPV_F_THRUST_LUT) {
        float tt[2];
#pragma unroll
        for (int l = 0; l < L; ++l) tt[l] = thrust_lut1(k, lut_dyn, act[l].w);
        target = Lane<V>::make(tt[0], tt[L - 1]);
      } else {
        target = thrust_poly<V>(k, a3);
      }
      typename Lane<V>::Mask done;
      if (wind_on) done = drone_substeps<V, ANG, false, true>(k, s, a0, a1, a2, target, wx, wy, wz, false, zero, zero, zero, zero, zero);
      else done = drone_substeps<V, ANG, false, false>(k, s, a0, a1, a2, target, wx, wy, wz, false, zero, zero, zero, zero, zero);
      // ---- per-step bookkeeping, exactly drone_tile's epilogue but on registers
#pragma unroll
      for (int l = 0; l < L; ++l) {
        if (!live[l]) continue;
        const bool d = mask_get(done, l);
        if (done_seq) done_seq[(long long)t * done_stride + ei[l]] = d ? 1 : 0;
        if (io.done && t == T - 1) io.done[ei[l]] = d ? 1 : 0;
        epi[l] += 1;
        if (d) {
          st.crash += 1.f;
          if (k.flags & FPV_F_AUTO_RESET) {   // restart from the reset snapshot: the registers take its rows
            st.epi += 1.f; st.len += (float)epi[l];
            float4 v[FPV_DRONE_PLANES];
#pragma unroll
            for (int p = 0; p < FPV_DRONE_PLANES; ++p) v[p] = ldg_stream(io.reset_state + p * io.stride + ei[l]);
            s.px = lane_set<V>(s.px, l, v[0].x); s.py = lane_set<V>(s.py, l, v[0].y); s.pz = lane_set<V>(s.pz, l, v[0].z);
            s.pt = lane_set<V>(s.pt, l, v[0].w);
            s.vx = lane_set<V>(s.vx, l, v[1].x); s.vy = lane_set<V>(s.vy, l, v[1].y); s.vz = lane_set<V>(s.vz, l, v[1].z);
            s.qw = lane_set<V>(s.qw, l, v[2].x); s.qx = lane_set<V>(s.qx, l, v[2].y); s.qy = lane_set<V>(s.qy, l, v[2].z);
            s.qz = lane_set<V>(s.qz, l, v[2].w);
            s.pr0 = lane_set<V>(s.pr0, l, v[3].x); s.pr1 = lane_set<V>(s.pr1, l, v[3].y); s.pr2 = lane_set<V>(s.pr2, l, v[3].z);
            spare[l] = v[3].w;
            epi[l] = 0;
            continue;
          }
        }
        if (!(fabsf((Lane<V>::get(s.px, l) + Lane<V>::get(s.py, l)) + Lane<V>::get(s.pz, l)) <= 3.0e38f)) st.nf += 1.f;
      }
#pragma unroll
      for (int l = 0; l < L; ++l) act[l] = act_next[l];
    }
    // ---- the state goes back once
#pragma unroll
    for (int l = 0; l < L; ++l) {
      if (!live[l]) continue;
      const long long e = base + (long long)l * 32;
      float4* const dst = io.state + e;
      if (io.acc_out)
        stg_stream(io.acc_out + e, make_float4(Lane<V>::get(s.ax, l), Lane<V>::get(s.ay, l), Lane<V>::get(s.az, l), 0.f));
      stg_stream(dst, make_float4(Lane<V>::get(s.px, l), Lane<V>::get(s.py, l), Lane<V>::get(s.pz, l), Lane<V>::get(s.pt, l)));
      stg_stream(dst + io.stride, make_float4(Lane<V>::get(s.vx, l), Lane<V>::get(s.vy, l), Lane<V>::get(s.vz, l), __int_as_float(epi[l])));
      stg_stream(dst + 2 * io.stride, make_float4(Lane<V>::get(s.qw, l), Lane<V>::get(s.qx, l), Lane<V>::get(s.qy, l), Lane<V>::get(s.qz, l)));
      stg_stream(dst + 3 * io.stride, make_float4(Lane<V>::get(s.pr0, l), Lane<V>::get(s.pr1, l), Lane<V>::get(s.pr2, l), spare[l]));
    }
    // next chunk
    unsigned v = 0;
    if (lane == 0) v = atomicAdd(io.work, 1u);
    v = __shfl_sync(0xffffffffu, v, 0);
    cur = total_warps + (long long)v;
  }
  if (lane == 0) {   // the last warp to finish puts the counters back
    const unsigned finished = atomicAdd(io.work + 1, 1u);
    if (finished == (unsigned)total_warps - 1u) { io.work[0] = 0u; io.work[1] = 0u; }
  }
  if (io.stats) stats_warp_flush(io.stats, st);
}

}  // namespace fpv
