// Mode A: the reference `Drone.step` (src/utils/components.py:220-248) for a batch of envs.
// One thread owns LANES envs (1 = float, 2 = packed F2), loads their state once (5 x 128-bit per env),
// runs K substeps entirely in registers and stores the state once.
#pragma once
#include "../../include/fpv_api.h"
#include "vec.cuh"

namespace fpv {

// Launch-constant parameters, derived on the host in double precision (api.cu) and passed by value
// (kernel-parameter constant bank => uniform loads, no __constant__ global state).
struct DroneK {
  float dt;
  int substeps;
  float one_minus_rtr;      // 1 - rates_transition_rate            components.py:187-188
  float rtr_max_rates;      // unused by the kernel maths; kept for debugging
  float max_rates;
  float rtr;
  float ttr;
  float one_minus_ttr;      // components.py:192-193
  float k_drag[3];          // kinematics.py:36
  float motor_xy[4][2];     // components.py:123-125
  float motor_radius;
  float spring_k, spring_c; // components.py:198
  float poly[4];            // throttle% -> N, high->low; evaluated at 100*(x+1)/2   components.py:136
  float wind[3];
  float grav_force_z;       // -g*m                                  kinematics.py:41-45
  float inv_mass;
  float mass;
  float ang_scale;          // deg2rad * dt                          kinematics.py:29
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
  const float4* wind_env;
  const float* lut;
  unsigned char* done;
  float4* acc_out;
  const float4* reset_state;
  const float4* override_R;
  fpv_stats_t* stats;
};

template <class V> struct Vec3 {
  V x, y, z;
};

// Obstacle SDF + normal for one motor point (GENERAL path only; warp-uniform object loop).
template <class V>
__device__ __forceinline__ void object_sdf(const fpv_object_t& o, V px, V py, V pz, V& d, V& nx, V& ny, V& nz) {
  if (o.kind == FPV_OBJ_SPHERE) {  // Target.calculate_distance/normal, components.py:773-777
    V dx = px - S<V>(o.x), dy = py - S<V>(o.y), dz = pz - S<V>(o.z);
    V r = vsqrt(vfma(dx, dx, vfma(dy, dy, dz * dz)));
    d = r - S<V>(o.a);
    nx = vdiv(dx, r);
    ny = vdiv(dy, r);
    nz = vdiv(dz, r);
  } else {  // Cylinder, components.py:710-729
    V dx = px - S<V>(o.x), dy = py - S<V>(o.y);
    V rad = vsqrt(vfma(dx, dx, dy * dy));
    V d2 = rad - S<V>(o.a);
    V top = S<V>(o.z + o.b);
    auto inside = vand(vlt(S<V>(o.z), pz), vlt(pz, top));
    V dlo = vabs(pz - S<V>(o.z)), dhi = vabs(pz - top);
    V dh = vmin(dlo, dhi);
    d = vsel(inside, d2, vsqrt(vfma(d2, d2, dh * dh)));
    // calculate_normal first makes the point RELATIVE to the base (:719) and then compares its z with
    // the ABSOLUTE band (:720) and cap heights (:725) -- reproduced as written.
    V qz = pz - S<V>(o.z);
    auto inside_n = vand(vlt(S<V>(o.z), qz), vlt(qz, top));
    auto below = vlt(vabs(qz - S<V>(o.z)), vabs(qz - top));
    nx = vsel(inside_n, vdiv(dx, rad), S<V>(0.f));
    ny = vsel(inside_n, vdiv(dy, rad), S<V>(0.f));
    nz = vsel(inside_n, S<V>(0.f), vsel(below, S<V>(-1.f), S<V>(1.f)));
  }
}

template <class V> struct DroneRegs {
  V px, py, pz, vx, vy, vz;
  V r00, r01, r02, r10, r11, r12, r20, r21, r22;
  V pr0, pr1, pr2, pt;
  V ax, ay, az;  // last substep's acceleration
};

// K reference steps for the envs held in `s`.  Returns the OR of the per-step crash flags.
template <class V, bool SMALL, bool GENERAL, bool FAST>
__device__ __forceinline__ typename Lane<V>::Mask drone_substeps(const DroneK& k, DroneRegs<V>& s, V a0, V a1, V a2,
                                                                   V thrust_target, V wx, V wy, V wz,
                                                                   bool has_override, V o_thrust,
                                                                   const DroneRegs<V>* ovr) {
  using M = typename Lane<V>::Mask;
  // action2force invariants (the action is held for the whole control step), components.py:185-193
  const V mr = S<V>(k.max_rates);
  const V c0 = vmin(vmax(vneg(a0) * mr, vneg(mr)), mr) * S<V>(k.rtr);
  const V c1 = vmin(vmax(vneg(a1) * mr, vneg(mr)), mr) * S<V>(k.rtr);
  const V c2 = vmin(vmax(vneg(a2) * mr, vneg(mr)), mr) * S<V>(k.rtr);
  const V tt = thrust_target * S<V>(k.ttr);
  const V omr = S<V>(k.one_minus_rtr), omt = S<V>(k.one_minus_ttr);
  const V dt = S<V>(k.dt), inv_m = S<V>(k.inv_mass), asc = S<V>(k.ang_scale);
  const V zero = S<V>(0.f);
  M done = vlt(S<V>(1.f), zero);  // all false

#pragma unroll 1
  for (int it = 0; it < k.substeps; ++it) {
    // ---- low-pass filters on rates and thrust, components.py:187-194
    const V w0 = vfma(s.pr0, omr, c0), w1 = vfma(s.pr1, omr, c1), w2 = vfma(s.pr2, omr, c2);
    s.pr0 = w0; s.pr1 = w1; s.pr2 = w2;
    V th = vfma(s.pt, omt, tt);
    s.pt = th;
    if (GENERAL && has_override) {  // components.py:230-232
      s.r00 = ovr->r00; s.r01 = ovr->r01; s.r02 = ovr->r02;
      s.r10 = ovr->r10; s.r11 = ovr->r11; s.r12 = ovr->r12;
      s.r20 = ovr->r20; s.r21 = ovr->r21; s.r22 = ovr->r22;
      th = o_thrust;
    }
    // ---- drag, kinematics.py:33-38: R * (k (.) (R^T (v + wind)) * |v + wind|)
    const V ux = s.vx + wx, uy = s.vy + wy, uz = s.vz + wz;
    const V n2 = vfma(ux, ux, vfma(uy, uy, uz * uz));
    const V nrm = FAST ? vsqrt_fast(n2) : vsqrt(n2);
    const V b0 = vfma(s.r00, ux, vfma(s.r10, uy, s.r20 * uz));
    const V b1 = vfma(s.r01, ux, vfma(s.r11, uy, s.r21 * uz));
    const V b2 = vfma(s.r02, ux, vfma(s.r12, uy, s.r22 * uz));
    const V f0 = (S<V>(k.k_drag[0]) * b0) * nrm, f1 = (S<V>(k.k_drag[1]) * b1) * nrm,
            f2 = (S<V>(k.k_drag[2]) * b2) * nrm;
    // ---- thrust + gravity + drag (components.py:242), thrust_vector kinematics.py:48-49
    V Fx = vfma(s.r00, f0, vfma(s.r01, f1, vfma(s.r02, f2, s.r02 * th)));
    V Fy = vfma(s.r10, f0, vfma(s.r11, f1, vfma(s.r12, f2, s.r12 * th)));
    V Fz = vfma(s.r20, f0, vfma(s.r21, f1, vfma(s.r22, f2, vfma(s.r22, th, S<V>(k.grav_force_z)))));
    // ---- motors, collisions, crash test (components.py:235-239, :198-214)
    M crashed = vlt(S<V>(1.f), zero);
    V cfx = zero, cfy = zero, cfz = zero;
    if (!GENERAL) {
      // ground only: distance = z, normal = +z (components.py:674-680); only z of M_rel @ R^T matters
      V minz = S<V>(3.0e38f);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const V mz = vfma(S<V>(k.motor_xy[m][0]), s.r20, vfma(S<V>(k.motor_xy[m][1]), s.r21, s.pz));
        minz = vmin(minz, mz);
        const V pen = mz - S<V>(k.motor_radius);
        // spring_force, kinematics.py:56-59: (-k*pen - c*(v.n)) * n
        const V f = vneg(vfma(S<V>(k.spring_k), pen, S<V>(k.spring_c) * s.vz));
        cfz = cfz + vsel(vlt(pen, zero), f, zero);
      }
      crashed = vlt(minz, zero);
      if (k.flags & FPV_F_GROUND) cfz = vsel(crashed, zero, cfz);  // early return with no force, :207-210
      else { cfz = zero; }
      // without the ground object the crash test of components.py:239 still applies
    } else {
      V mxw[4], myw[4], mzw[4];
      V minz = S<V>(3.0e38f);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const V ox = S<V>(k.motor_xy[m][0]), oy = S<V>(k.motor_xy[m][1]);
        mxw[m] = vfma(ox, s.r00, vfma(oy, s.r01, s.px));
        myw[m] = vfma(ox, s.r10, vfma(oy, s.r11, s.py));
        mzw[m] = vfma(ox, s.r20, vfma(oy, s.r21, s.pz));
        minz = vmin(minz, mzw[m]);
      }
      const int n_obj = k.n_objects + ((k.flags & FPV_F_GROUND) ? 1 : 0);
      for (int o = 0; o < n_obj; ++o) {  // object order: extra objects first, ground last
        V d[4], nx[4], ny[4], nz[4];
        M hit = vlt(S<V>(1.f), zero);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if (o < k.n_objects) object_sdf<V>(k.objects[o], mxw[m], myw[m], mzw[m], d[m], nx[m], ny[m], nz[m]);
          else { d[m] = mzw[m]; nx[m] = zero; ny[m] = zero; nz[m] = S<V>(1.f); }
          hit = vor(hit, vlt(d[m], zero));
        }
        hit = vand(hit, vnot(crashed));
        crashed = vor(crashed, hit);
        V ox = zero, oy = zero, oz = zero;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const V pen = d[m] - S<V>(k.motor_radius);
          const V vn = vfma(s.vx, nx[m], vfma(s.vy, ny[m], s.vz * nz[m]));
          const V f = vneg(vfma(S<V>(k.spring_k), pen, S<V>(k.spring_c) * vn));
          const M act = vlt(pen, zero);
          ox = ox + vsel(act, f * nx[m], zero);
          oy = oy + vsel(act, f * ny[m], zero);
          oz = oz + vsel(act, f * nz[m], zero);
        }
        const M live = vnot(crashed);
        cfx = cfx + vsel(live, ox, zero);
        cfy = cfy + vsel(live, oy, zero);
        cfz = cfz + vsel(live, oz, zero);
      }
      crashed = vor(crashed, vlt(minz, zero));  // components.py:239
    }
    done = vor(done, crashed);
    Fx = Fx + cfx; Fy = Fy + cfy; Fz = Fz + cfz;
    // ---- acceleration, components.py:243
    s.ax = Fx * inv_m; s.ay = Fy * inv_m; s.az = Fz * inv_m;
    // ---- translation: x += v*dt with the OLD v, then v += a*dt, kinematics.py:21-22
    s.px = vfma(s.vx, dt, s.px); s.py = vfma(s.vy, dt, s.py); s.pz = vfma(s.vz, dt, s.pz);
    s.vx = vfma(s.ax, dt, s.vx); s.vy = vfma(s.ay, dt, s.vy); s.vz = vfma(s.az, dt, s.vz);
    // ---- attitude: E = Rz(yaw)Ry(pitch)Rx(roll) of deg2rad(rates)*dt, R <- R E^T E^T
    //      (rotate_body_by_rates kinematics.py:27-30 runs inside update_kinematic_step :23 AND again in
    //      Drone.update components.py:218)
    V sr, cr, sp, cp, sy, cy;
    vsincos<SMALL>(w0 * asc, sr, cr);
    vsincos<SMALL>(w1 * asc, sp, cp);
    vsincos<SMALL>(w2 * asc, sy, cy);
    const V sysp = sy * sp, cysp = cy * sp;
    const V e00 = cy * cp, e01 = vfma(cysp, sr, vneg(sy * cr)), e02 = vfma(cysp, cr, sy * sr);
    const V e10 = sy * cp, e11 = vfma(sysp, sr, cy * cr), e12 = vfma(sysp, cr, vneg(cy * sr));
    const V e20 = vneg(sp), e21 = cp * sr, e22 = cp * cr;
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
      V t0, t1, t2;
      t0 = vfma(s.r00, e00, vfma(s.r01, e01, s.r02 * e02));
      t1 = vfma(s.r00, e10, vfma(s.r01, e11, s.r02 * e12));
      t2 = vfma(s.r00, e20, vfma(s.r01, e21, s.r02 * e22));
      s.r00 = t0; s.r01 = t1; s.r02 = t2;
      t0 = vfma(s.r10, e00, vfma(s.r11, e01, s.r12 * e02));
      t1 = vfma(s.r10, e10, vfma(s.r11, e11, s.r12 * e12));
      t2 = vfma(s.r10, e20, vfma(s.r11, e21, s.r12 * e22));
      s.r10 = t0; s.r11 = t1; s.r12 = t2;
      t0 = vfma(s.r20, e00, vfma(s.r21, e01, s.r22 * e02));
      t1 = vfma(s.r20, e10, vfma(s.r21, e11, s.r22 * e12));
      t2 = vfma(s.r20, e20, vfma(s.r21, e21, s.r22 * e22));
      s.r20 = t0; s.r21 = t1; s.r22 = t2;
    }
  }
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

template <class V> struct Pack;
template <> struct Pack<float> {
  static __device__ __forceinline__ float x(const float4* q) { return q[0].x; }
  static __device__ __forceinline__ float y(const float4* q) { return q[0].y; }
  static __device__ __forceinline__ float z(const float4* q) { return q[0].z; }
  static __device__ __forceinline__ float w(const float4* q) { return q[0].w; }
};
template <> struct Pack<F2> {
  static __device__ __forceinline__ F2 x(const float4* q) { return F2{make_float2(q[0].x, q[1].x)}; }
  static __device__ __forceinline__ F2 y(const float4* q) { return F2{make_float2(q[0].y, q[1].y)}; }
  static __device__ __forceinline__ F2 z(const float4* q) { return F2{make_float2(q[0].z, q[1].z)}; }
  static __device__ __forceinline__ F2 w(const float4* q) { return F2{make_float2(q[0].w, q[1].w)}; }
};

// Block-level accumulation of the episode statistics: warp shuffle -> shared -> one atomic per CTA.
__device__ __forceinline__ void stats_accumulate(fpv_stats_t* stats, float steps, float crashes, float episodes,
                                                 float len_sum, float nonfinite) {
  __shared__ float red[5][32];
  float v[5] = {steps, crashes, episodes, len_sum, nonfinite};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    if (lane == 0) red[i][warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      float t = lane < nw ? red[i][lane] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      v[i] = t;
    }
    if (lane == 0) {
      if (v[0] != 0.f) atomicAdd(&stats->env_steps, (double)v[0]);
      if (v[1] != 0.f) atomicAdd(&stats->crashes, (double)v[1]);
      if (v[2] != 0.f) atomicAdd(&stats->episodes, (double)v[2]);
      if (v[3] != 0.f) atomicAdd(&stats->episode_len_sum, (double)v[3]);
      if (v[4] != 0.f) atomicAdd(&stats->nonfinite, (double)v[4]);
    }
  }
}

template <class V, bool SMALL, bool GENERAL, bool FAST, int THREADS>
__global__ void __launch_bounds__(THREADS) drone_step_kernel(const __grid_constant__ DroneK k, const DroneIO io) {
  constexpr int L = Lane<V>::N;
  extern __shared__ float lut_s[];
  const bool use_lut = (k.flags & FPV_F_THRUST_LUT) != 0;
  if (use_lut) {  // stage the motor curve in shared memory once per CTA
    for (int i = threadIdx.x; i < k.lut_n; i += THREADS) lut_s[i] = io.lut[i];
    __syncthreads();
  }
  // thread t, slot l -> env blockBase + l*THREADS + t: every 128-bit access of a warp is one contiguous 512 B run
  const long long base = (long long)blockIdx.x * (THREADS * L) + threadIdx.x;
  const bool active = base < io.n;
  float st_steps = 0.f, st_crash = 0.f, st_epi = 0.f, st_len = 0.f, st_nf = 0.f;
  if (active) {
    long long ei[L];  // slot -> env index; a slot past the end re-reads env n-1 and is never stored
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

    DroneRegs<V> s;
    s.px = Pack<V>::x(q[0]); s.py = Pack<V>::y(q[0]); s.pz = Pack<V>::z(q[0]); s.pt = Pack<V>::w(q[0]);
    s.vx = Pack<V>::x(q[1]); s.vy = Pack<V>::y(q[1]); s.vz = Pack<V>::z(q[1]);
    s.r00 = Pack<V>::x(q[2]); s.r01 = Pack<V>::y(q[2]); s.r02 = Pack<V>::z(q[2]); s.pr0 = Pack<V>::w(q[2]);
    s.r10 = Pack<V>::x(q[3]); s.r11 = Pack<V>::y(q[3]); s.r12 = Pack<V>::z(q[3]); s.pr1 = Pack<V>::w(q[3]);
    s.r20 = Pack<V>::x(q[4]); s.r21 = Pack<V>::y(q[4]); s.r22 = Pack<V>::z(q[4]); s.pr2 = Pack<V>::w(q[4]);
    s.ax = S<V>(0.f); s.ay = S<V>(0.f); s.az = S<V>(0.f);
    int epi[L];
#pragma unroll
    for (int l = 0; l < L; ++l) epi[l] = __float_as_int(q[1][l].w);

    V wx = S<V>(k.wind[0]), wy = S<V>(k.wind[1]), wz = S<V>(k.wind[2]);
    if (io.wind_env) {
      float4 w[L];
#pragma unroll
      for (int l = 0; l < L; ++l) w[l] = ldg_stream(io.wind_env + ei[l]);
      wx = Pack<V>::x(w); wy = Pack<V>::y(w); wz = Pack<V>::z(w);
    }
    const V a0 = Pack<V>::x(act), a1 = Pack<V>::y(act), a2 = Pack<V>::z(act), a3 = Pack<V>::w(act);
    V target;
    if (use_lut) {
      float t[2];
#pragma unroll
      for (int l = 0; l < L; ++l) t[l] = thrust_lut1(k, lut_s, act[l].w);
      target = Lane<V>::make(t[0], t[L - 1]);
    } else {
      target = thrust_poly<V>(k, a3);
    }
    DroneRegs<V> ovr;
    V o_thrust = S<V>(0.f);
    const bool has_ovr = GENERAL && io.override_R != nullptr;
    if (has_ovr) {
      float4 r0[L], r1[L], r2[L];
#pragma unroll
      for (int l = 0; l < L; ++l) {
        r0[l] = ldg_stream(io.override_R + ei[l]);
        r1[l] = ldg_stream(io.override_R + io.n + ei[l]);
        r2[l] = ldg_stream(io.override_R + 2 * io.n + ei[l]);
      }
      ovr.r00 = Pack<V>::x(r0); ovr.r01 = Pack<V>::y(r0); ovr.r02 = Pack<V>::z(r0); o_thrust = Pack<V>::w(r0);
      ovr.r10 = Pack<V>::x(r1); ovr.r11 = Pack<V>::y(r1); ovr.r12 = Pack<V>::z(r1);
      ovr.r20 = Pack<V>::x(r2); ovr.r21 = Pack<V>::y(r2); ovr.r22 = Pack<V>::z(r2);
    }

    auto done = drone_substeps<V, SMALL, GENERAL, FAST>(k, s, a0, a1, a2, target, wx, wy, wz, has_ovr, o_thrust, &ovr);

    // ---- epilogue per env: episode bookkeeping, freeze / auto-reset, stores
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const long long e = base + (long long)l * THREADS;
      if (e >= io.n) break;
      bool d = mask_get(done, l);
      int ep = epi[l];
      if (ep < 0) {  // frozen after a crash (FPV_F_FREEZE_DONE): state in memory stays as it is, done is sticky
        if (io.done) io.done[e] = 1;
        continue;
      }
      float4 o0, o1, o2, o3, o4;
      o0 = make_float4(Lane<V>::get(s.px, l), Lane<V>::get(s.py, l), Lane<V>::get(s.pz, l), Lane<V>::get(s.pt, l));
      o1 = make_float4(Lane<V>::get(s.vx, l), Lane<V>::get(s.vy, l), Lane<V>::get(s.vz, l), 0.f);
      o2 = make_float4(Lane<V>::get(s.r00, l), Lane<V>::get(s.r01, l), Lane<V>::get(s.r02, l), Lane<V>::get(s.pr0, l));
      o3 = make_float4(Lane<V>::get(s.r10, l), Lane<V>::get(s.r11, l), Lane<V>::get(s.r12, l), Lane<V>::get(s.pr1, l));
      o4 = make_float4(Lane<V>::get(s.r20, l), Lane<V>::get(s.r21, l), Lane<V>::get(s.r22, l), Lane<V>::get(s.pr2, l));
      ep += 1;
      st_steps += 1.f;
      const bool fin = isfinite(o0.x) && isfinite(o0.y) && isfinite(o0.z) && isfinite(o1.x) && isfinite(o1.y) &&
                       isfinite(o1.z);
      if (!fin) st_nf += 1.f;
      if (d) {
        st_crash += 1.f;
        if ((k.flags & FPV_F_AUTO_RESET) && io.reset_state) {
          st_epi += 1.f; st_len += (float)ep;
          o0 = io.reset_state[e]; o1 = io.reset_state[io.stride + e]; o2 = io.reset_state[2 * io.stride + e];
          o3 = io.reset_state[3 * io.stride + e]; o4 = io.reset_state[4 * io.stride + e];
          ep = 0;
        } else if (k.flags & FPV_F_FREEZE_DONE) {
          st_epi += 1.f; st_len += (float)ep;
          ep = -ep - 1;
        }
      }
      o1.w = __int_as_float(ep);
      stg_stream(io.state + e, o0);
      stg_stream(io.state + io.stride + e, o1);
      stg_stream(io.state + 2 * io.stride + e, o2);
      stg_stream(io.state + 3 * io.stride + e, o3);
      stg_stream(io.state + 4 * io.stride + e, o4);
      if (io.done) io.done[e] = d ? 1 : 0;
      if (io.acc_out)
        stg_stream(io.acc_out + e, make_float4(Lane<V>::get(s.ax, l), Lane<V>::get(s.ay, l), Lane<V>::get(s.az, l), 0.f));
    }
  }
  if (io.stats) stats_accumulate(io.stats, st_steps, st_crash, st_epi, st_len, st_nf);
}

}  // namespace fpv
