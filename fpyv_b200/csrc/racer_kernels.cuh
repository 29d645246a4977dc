// Mode B: the acro rate-PID drone of the reference's tests/racer_drone_test.py (`PID` :11-32, `Racer` :68-103) on the
// packed ring kernel: two envs per thread on FFMA2/FMUL2/FADD2, TMA-fed 64-env chunks, K substeps in registers.
// The reference mutates a 3x3 matrix  Rot <- Rot @ Rx(w0) Ry(w1) Rz(w2)  (scipy intrinsic "XYZ", ANGLE = omega, not omega*dt
// -- :99) and re-orthonormalises it on every step (Rotation.from_matrix); the same rotation is carried here as the unit
// quaternion q <- q (x) qx(w0/2) (x) qy(w1/2) (x) qz(w2/2), normalised every substep: 5 planes instead of 7.
//   plane 0: position xyz, PID first-call flag (1.0 / 0.0)     :20, :47-51
//   plane 1: velocity xyz, omega[0]
//   plane 2: orientation quaternion w x y z (helper_functions.py:65-117 convention)
//   plane 3: PID integral xyz, omega[1]
//   plane 4: PID last error xyz, omega[2]
#pragma once
#include "../../include/fpv_api.h"
#include "vec.cuh"
#include "ring_kernels.cuh"

namespace fpv {

struct RacerK {
  float dt, inv_dt;
  int substeps;
  float dt_over_m;        // dt / mass
  float dt_over_I[3];
  float gains[3][3];
  float vel_decay;
};

struct RacerIO {
  float4* state;
  long long n, stride;
  const float4* actions;
  float4* torque_out;
  unsigned* work;
  unsigned* chunk_epoch;   // unused (always null): mode B has no chained form
  unsigned epoch;
  unsigned* err;
  unsigned long long* trace;
};

__global__ void racer_reset_kernel(float4* state, long long n, long long stride, const unsigned char* mask) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (mask && !mask[e]) return;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  state[e] = make_float4(0.f, 0.f, 0.f, 1.f);  // .w = PID first-call flag (racer_drone_test.py:20)
  state[stride + e] = z;
  state[2 * stride + e] = make_float4(1.f, 0.f, 0.f, 0.f);
  state[3 * stride + e] = z;
  state[4 * stride + e] = z;
}

// Racer.orientation (the 3x3 matrix the reference keeps, :73) and angular_velocity read from the planes above
__global__ void racer_observe_kernel(const float4* state, long long n, long long stride, float* R, float* omega) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  if (R) {
    const float4 q = state[2 * stride + e];
    const float w = q.x, x = q.y, y = q.z, z = q.w;
    float* o = R + 9 * e;
    o[0] = 1.f - 2.f * (y * y + z * z); o[1] = 2.f * (x * y - z * w);       o[2] = 2.f * (x * z + y * w);
    o[3] = 2.f * (x * y + z * w);       o[4] = 1.f - 2.f * (x * x + z * z); o[5] = 2.f * (y * z - x * w);
    o[6] = 2.f * (x * z - y * w);       o[7] = 2.f * (y * z + x * w);       o[8] = 1.f - 2.f * (x * x + y * y);
  }
  if (omega) {
    omega[3 * e] = state[stride + e].w;
    omega[3 * e + 1] = state[3 * stride + e].w;
    omega[3 * e + 2] = state[4 * stride + e].w;
  }
}

// sin/cos for ANY argument the Racer produces (|omega| reaches ~80 rad in the reference's own demo, so half angles of
// ~40 rad).  Reduction by PI, not pi/2: x = j pi + r with |r| <= pi/2, sin x = (-1)^j sin r, cos x = (-1)^j cos r -- the
// fix-up is ONE sign flip shared by both results (an XOR of the sign bits, done on the packed 64-bit value) instead of a
// per-lane quadrant swap.  Cody-Waite in three FMA steps (products exact for |j| < 2^15, i.e. |x| < 1e5 rad); the kernels
// are least-squares fits on [-pi/2, pi/2]: sin r = r + r^3 Q(r^2) (degree 9), cos r = 1 + r^2 R(r^2) (degree 10), both
// within 1.3e-7 absolute of the exact functions in float32 evaluation.  Arguments beyond 3e4 rad (never reached by a
// finite-gain PID at these step sizes) go through sincosf per lane.
template <class V> __device__ __forceinline__ void sincos_pi_kernels(V r, V& s, V& c) {
  const V r2 = r * r;
  V q = vfma(S<V>(2.6050197e-06f), r2, S<V>(-1.9808984e-04f));
  q = vfma(q, r2, S<V>(8.33305e-03f));
  q = vfma(q, r2, S<V>(-1.6666658e-01f));
  s = vfma(q, r2 * r, r);
  V p = vfma(S<V>(-2.6049608e-07f), r2, S<V>(2.4760055e-05f));
  p = vfma(p, r2, S<V>(-1.388836e-03f));
  p = vfma(p, r2, S<V>(4.1666634e-02f));
  p = vfma(p, r2, S<V>(-0.5f));
  c = vfma(p, r2, S<V>(1.0f));
}
template <class V> __device__ __forceinline__ void sincos_reduced(V x, V& s, V& c);
template <> __device__ __forceinline__ void sincos_reduced<float>(float x, float& s, float& c) {
  if (!(fabsf(x) <= 3.0e4f)) { sincosf(x, &s, &c); return; }
  const float t = fmaf(x, 0.318309886f, 12582912.f);
  const float j = t - 12582912.f;
  float r = fmaf(j, -3.14159203e+00f, x);
  r = fmaf(j, -6.27832947e-07f, r);
  r = fmaf(j, -1.07806051e-14f, r);
  sincos_pi_kernels<float>(r, s, c);
  const unsigned flip = (unsigned)__float_as_int(t) << 31;        // j odd: both signs flip
  s = __int_as_float(__float_as_int(s) ^ flip);
  c = __int_as_float(__float_as_int(c) ^ flip);
}
template <> __device__ __forceinline__ void sincos_reduced<F2>(F2 x, F2& s, F2& c) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  if (!(fmaxf(fabsf(x0), fabsf(x1)) <= 3.0e4f)) {
    float s0, c0, s1, c1;
    sincos_reduced<float>(x0, s0, c0);
    sincos_reduced<float>(x1, s1, c1);
    s = f2_pack(s0, s1);
    c = f2_pack(c0, c1);
    return;
  }
  const F2 magic = S<F2>(12582912.f);
  const F2 t = vfma(x, S<F2>(0.318309886f), magic);
  const F2 j = t - magic;
  F2 r = vfma(j, S<F2>(-3.14159203e+00f), x);
  r = vfma(j, S<F2>(-6.27832947e-07f), r);
  r = vfma(j, S<F2>(-1.07806051e-14f), r);
  sincos_pi_kernels<F2>(r, s, c);
  // bit 0 of each half of t (the parity of j) moved to that half's sign position
  const unsigned long long flip = (t.v << 31) & 0x8000000080000000ull;
  s.v ^= flip;
  c.v ^= flip;
}

template <class V_>
struct RacerMode {
  using V = V_;
  using K = RacerK;
  using IO = RacerIO;
  static constexpr int PLANES = 5;
  static constexpr int ROWS = PLANES + 1;
  struct Ctx {};
  static __device__ __forceinline__ bool chained(const K&) { return false; }
  static __device__ __forceinline__ constexpr int row_bytes(int) { return 16; }
  static __device__ __forceinline__ const void* row_ptr(const IO& io, int r, long long first) {
    return (r < PLANES ? io.state + r * io.stride : io.actions) + first;
  }
  static __device__ __forceinline__ void stage(const K&, const IO&, unsigned char*, int, int) {}
  static __device__ __forceinline__ Ctx begin(const K&, const IO&) { return Ctx{}; }
  static __device__ __forceinline__ void finish(const K&, const IO&, Ctx&) {}

  template <class PreStore>
  static __device__ __forceinline__ void tile(const K& k, const IO& io, const unsigned char*, const float4 (&rows)[ROWS][Lane<V>::N],
                                              const long long (&)[Lane<V>::N], long long base, Ctx&, PreStore pre_store) {
    constexpr int L = Lane<V>::N;
    V px = Pack<V>::x(rows[0]), py = Pack<V>::y(rows[0]), pz = Pack<V>::z(rows[0]);
    V vx = Pack<V>::x(rows[1]), vy = Pack<V>::y(rows[1]), vz = Pack<V>::z(rows[1]);
    V qw = Pack<V>::x(rows[2]), qx = Pack<V>::y(rows[2]), qy = Pack<V>::z(rows[2]), qz = Pack<V>::w(rows[2]);
    V ie[3] = {Pack<V>::x(rows[3]), Pack<V>::y(rows[3]), Pack<V>::z(rows[3])};
    V le[3] = {Pack<V>::x(rows[4]), Pack<V>::y(rows[4]), Pack<V>::z(rows[4])};
    V w[3] = {Pack<V>::w(rows[1]), Pack<V>::w(rows[3]), Pack<V>::w(rows[4])};
    const V sp[3] = {Pack<V>::x(rows[5]), Pack<V>::y(rows[5]), Pack<V>::z(rows[5])};
    const V thrust = Pack<V>::w(rows[5]);
    float nf[2];
#pragma unroll
    for (int l = 0; l < L; ++l) nf[l] = rows[0][l].w != 0.f ? 0.f : 1.f;
    V notfirst = Lane<V>::make(nf[0], nf[L - 1]);   // 0 on the PIDs' first call: no derivative term (:28)
    const V one = S<V>(1.f), half = S<V>(0.5f), dt = S<V>(k.dt), inv_dt = S<V>(k.inv_dt);
    const V s_m = thrust * S<V>(k.dt_over_m), decay = S<V>(k.vel_decay), two = S<V>(2.f);
    V tq[3] = {S<V>(0.f), S<V>(0.f), S<V>(0.f)};
#pragma unroll 1
    for (int it = 0; it < k.substeps; ++it) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {   // PID.step, racer_drone_test.py:22-32, then omega += torque dt / I (:98)
        const V err = sp[i] - w[i];
        ie[i] = vfma(err, dt, ie[i]);
        const V de = ((err - le[i]) * inv_dt) * notfirst;
        le[i] = err;
        tq[i] = vfma(S<V>(k.gains[i][0]), err, vfma(S<V>(k.gains[i][1]), ie[i], S<V>(k.gains[i][2]) * de));
        w[i] = vfma(tq[i], S<V>(k.dt_over_I[i]), w[i]);
      }
      notfirst = one;
      // orientation <- orientation @ Rx(w0) Ry(w1) Rz(w2)   (:99; angle = omega).  qE = qx (x) qy (x) qz of the half angles
      V sa, ca, sb, cb, sc, cc;
      sincos_reduced<V>(w[0] * half, sa, ca);
      sincos_reduced<V>(w[1] * half, sb, cb);
      sincos_reduced<V>(w[2] * half, sc, cc);
      const V aw = ca * cb, ax = sa * cb, ay = ca * sb, az = sa * sb;             // qx (x) qy
      const V ew = vfma(aw, cc, vneg(az * sc)), ex = vfma(ax, cc, ay * sc);
      const V ey = vfma(ay, cc, vneg(ax * sc)), ez = vfma(az, cc, aw * sc);
      const V nw = vfma(qw, ew, vneg(vfma(qx, ex, vfma(qy, ey, qz * ez))));
      const V nx = vfma(qw, ex, vfma(qx, ew, vfma(qy, ez, vneg(qz * ey))));
      const V ny = vfma(qw, ey, vfma(qy, ew, vfma(qz, ex, vneg(qx * ez))));
      const V nz = vfma(qw, ez, vfma(qz, ew, vfma(qx, ey, vneg(qy * ex))));
      const V inv = vrsqrt_fast(vfma(nw, nw, vfma(nx, nx, vfma(ny, ny, nz * nz))));
      qw = nw * inv; qx = nx * inv; qy = ny * inv; qz = nz * inv;
      // force = thrust * (new) body z, acc = F / m, v <- decay v + a dt, x <- x + v_new dt   (:100-103)
      const V bx = two * vfma(qx, qz, qw * qy), by = two * vfma(qy, qz, vneg(qw * qx));
      const V bz = vfma(vneg(two), vfma(qx, qx, qy * qy), one);
      vx = vfma(decay, vx, bx * s_m); vy = vfma(decay, vy, by * s_m); vz = vfma(decay, vz, bz * s_m);
      px = vfma(vx, dt, px); py = vfma(vy, dt, py); pz = vfma(vz, dt, pz);
    }
    pre_store();
#pragma unroll
    for (int l = 0; l < L; ++l) {
      const long long e = base + (long long)l * 32;
      if (e >= io.n) break;
      const auto g = [&](V v) { return Lane<V>::get(v, l); };
      float4* const dst = io.state + e;
      stg_stream(dst, make_float4(g(px), g(py), g(pz), 0.f));                         // the PIDs have been called
      stg_stream(dst + io.stride, make_float4(g(vx), g(vy), g(vz), g(w[0])));
      stg_stream(dst + 2 * io.stride, make_float4(g(qw), g(qx), g(qy), g(qz)));
      stg_stream(dst + 3 * io.stride, make_float4(g(ie[0]), g(ie[1]), g(ie[2]), g(w[1])));
      stg_stream(dst + 4 * io.stride, make_float4(g(le[0]), g(le[1]), g(le[2]), g(w[2])));
      if (io.torque_out) stg_stream(io.torque_out + e, make_float4(g(tq[0]), g(tq[1]), g(tq[2]), 0.f));
    }
  }
};

}  // namespace fpv
