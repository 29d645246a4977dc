// Lane value types for the dynamics kernels.
//   float : one env per thread.
//   F2    : two envs per thread in one 64-bit register pair, so that add/mul/fma issue as the
//           Blackwell packed-FP32 instructions (PTX add/mul/fma.rn.f32x2 -> SASS FADD2/FMUL2/FFMA2):
//           half the issue slots per env for the same FP32 lane throughput.
#pragma once
#include <cuda_runtime.h>

namespace fpv {

// F2 is ONE 64-bit value (an aligned register pair) so that the allocator never splits the halves; all
// arithmetic is inline PTX add/mul/fma.rn.f32x2 on .b64 operands.
struct F2 {
  unsigned long long v;
};
__device__ __forceinline__ F2 f2_pack(float lo, float hi) {
  F2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(F2 a, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
struct B2 {
  bool x, y;
};

// ---- construction / lanes
template <class V> struct Lane;
template <> struct Lane<float> {
  static constexpr int N = 1;
  using Mask = bool;
  static __device__ __forceinline__ float splat(float s) { return s; }
  static __device__ __forceinline__ float get(float v, int) { return v; }
  static __device__ __forceinline__ float make(float a, float) { return a; }
};
template <> struct Lane<F2> {
  static constexpr int N = 2;
  using Mask = B2;
  static __device__ __forceinline__ F2 splat(float s) { return f2_pack(s, s); }
  static __device__ __forceinline__ float get(F2 v, int i) {
    float lo, hi;
    f2_unpack(v, lo, hi);
    return i ? hi : lo;
  }
  static __device__ __forceinline__ F2 make(float a, float b) { return f2_pack(a, b); }
};

// ---- float
__device__ __forceinline__ float vfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float vneg(float a) { return -a; }
__device__ __forceinline__ float vmin(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float vmax(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float vsqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ float vsqrt_fast(float a) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float vrsqrt_fast(float a) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ bool vlt(float a, float b) { return a < b; }
__device__ __forceinline__ bool vle(float a, float b) { return a <= b; }
__device__ __forceinline__ bool vor(bool a, bool b) { return a || b; }
__device__ __forceinline__ bool vand(bool a, bool b) { return a && b; }
__device__ __forceinline__ bool vnot(bool a) { return !a; }
__device__ __forceinline__ bool vany(bool a) { return a; }
__device__ __forceinline__ float vsel(bool m, float a, float b) { return m ? a : b; }
__device__ __forceinline__ float vtrunc_i(float a) { return truncf(a); }
__device__ __forceinline__ float vabs(float a) { return fabsf(a); }
__device__ __forceinline__ float vdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float vdiv_fast(float a, float b) { return __fdividef(a, b); }   // 2 ulp, |b| < 2^126
__device__ __forceinline__ bool vfinite(float a) { return isfinite(a); }

// ---- F2 (packed pair)
__device__ __forceinline__ F2 operator+(F2 a, F2 b) {
  F2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ F2 operator*(F2 a, F2 b) {
  F2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ F2 vfma(F2 a, F2 b, F2 c) {
  F2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}
// negation as two scalar negs between an unpack and a pack: ptxas folds this into the consumer's operand
// modifier (FFMA2 R, -R.F32x2, ...), whereas a 64-bit XOR of the sign bits stays as two LOP3
__device__ __forceinline__ F2 vneg(F2 a) {
  float x, y;
  f2_unpack(a, x, y);
  return f2_pack(-x, -y);
}
__device__ __forceinline__ F2 operator-(F2 a, F2 b) {
  F2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ F2 operator-(F2 a) { return vneg(a); }
#define FPV_F2_MAP1(name, expr)                              \
  __device__ __forceinline__ F2 name(F2 a) {                 \
    float x, y;                                              \
    f2_unpack(a, x, y);                                      \
    float rx, ry;                                            \
    { const float v = x; rx = (expr); }                      \
    { const float v = y; ry = (expr); }                      \
    return f2_pack(rx, ry);                                  \
  }
#define FPV_F2_MAP2(name, expr)                              \
  __device__ __forceinline__ F2 name(F2 a, F2 b) {           \
    float ax, ay, bx, by;                                    \
    f2_unpack(a, ax, ay);                                    \
    f2_unpack(b, bx, by);                                    \
    float rx, ry;                                            \
    { const float u = ax, v = bx; rx = (expr); }             \
    { const float u = ay, v = by; ry = (expr); }             \
    return f2_pack(rx, ry);                                  \
  }
FPV_F2_MAP2(vmin, fminf(u, v))
FPV_F2_MAP2(vmax, fmaxf(u, v))
FPV_F2_MAP2(vdiv, __fdiv_rn(u, v))
FPV_F2_MAP2(vdiv_fast, __fdividef(u, v))
FPV_F2_MAP1(vsqrt, __fsqrt_rn(v))
FPV_F2_MAP1(vsqrt_fast, vsqrt_fast(v))
FPV_F2_MAP1(vrsqrt_fast, vrsqrt_fast(v))
FPV_F2_MAP1(vabs, fabsf(v))
__device__ __forceinline__ B2 vlt(F2 a, F2 b) {
  float ax, ay, bx, by;
  f2_unpack(a, ax, ay);
  f2_unpack(b, bx, by);
  return B2{ax < bx, ay < by};
}
__device__ __forceinline__ B2 vle(F2 a, F2 b) {
  float ax, ay, bx, by;
  f2_unpack(a, ax, ay);
  f2_unpack(b, bx, by);
  return B2{ax <= bx, ay <= by};
}
__device__ __forceinline__ B2 vor(B2 a, B2 b) { return B2{a.x || b.x, a.y || b.y}; }
__device__ __forceinline__ B2 vand(B2 a, B2 b) { return B2{a.x && b.x, a.y && b.y}; }
__device__ __forceinline__ B2 vnot(B2 a) { return B2{!a.x, !a.y}; }
__device__ __forceinline__ bool vany(B2 a) { return a.x || a.y; }
__device__ __forceinline__ F2 vsel(B2 m, F2 a, F2 b) {
  float ax, ay, bx, by;
  f2_unpack(a, ax, ay);
  f2_unpack(b, bx, by);
  return f2_pack(m.x ? ax : bx, m.y ? ay : by);
}
__device__ __forceinline__ B2 vfinite(F2 a) {
  float x, y;
  f2_unpack(a, x, y);
  return B2{isfinite(x), isfinite(y)};
}

__device__ __forceinline__ bool mask_get(bool m, int) { return m; }
__device__ __forceinline__ bool mask_get(B2 m, int i) { return i ? m.y : m.x; }

// scalar (warp-uniform parameter) with lane value
template <class V> __device__ __forceinline__ V S(float s) { return Lane<V>::splat(s); }

// ---- sin/cos
// Polynomial kernels on |x| <= pi/4 (the same minimax coefficients a reduced-argument sinf/cosf uses);
// max error < 1 ulp there.  Used directly when the host proves the angle bound (SMALL), otherwise
// after a Cody-Waite reduction by pi/2 (float path only needs accurate sincosf for |x| up to ~1e5).
template <class V> __device__ __forceinline__ void sincos_poly(V x, V& s, V& c) {
  const V x2 = x * x;
  V ps = vfma(S<V>(-1.95152959e-4f), x2, S<V>(8.33216087e-3f));
  ps = vfma(ps, x2, S<V>(-1.66666546e-1f));
  const V x3 = x2 * x;
  s = vfma(ps, x3, x);
  V pc = vfma(S<V>(2.44331571e-5f), x2, S<V>(-1.38873163e-3f));
  pc = vfma(pc, x2, S<V>(4.16666457e-2f));
  pc = vfma(pc, x2, S<V>(-0.5f));
  c = vfma(pc, x2, S<V>(1.0f));
}

// |x| <= 0.1 rad: x - x^3/6 + x^5/120 and 1 - x^2/2 + x^4/24 (next terms: 2e-11 and 1.4e-9 * x^6 -> < 1.4e-15)
template <class V> __device__ __forceinline__ void sincos_tiny(V x, V& s, V& c) {
  const V x2 = x * x;
  const V ps = vfma(x2, S<V>(8.3333333e-3f), S<V>(-1.6666667e-1f));
  s = vfma(ps, x2 * x, x);
  const V pc = vfma(x2, S<V>(4.1666668e-2f), S<V>(-0.5f));
  c = vfma(pc, x2, S<V>(1.0f));
}

// |x| <= 0.03 rad: x (1 - x^2/6) and 1 - x^2/2 (next terms: x^5/120 -> 7e-9 relative, x^4/24 -> 3.4e-8 absolute)
template <class V> __device__ __forceinline__ void sincos_micro(V x, V& s, V& c) {
  const V x2 = x * x;
  s = x * vfma(x2, S<V>(-1.6666667e-1f), S<V>(1.0f));
  c = vfma(x2, S<V>(-0.5f), S<V>(1.0f));
}

template <int ANG> __device__ __forceinline__ void vsincos(float x, float& s, float& c) {
  if (ANG == 3) sincos_micro<float>(x, s, c);
  else if (ANG == 2) sincos_tiny<float>(x, s, c);
  else if (ANG == 1) sincos_poly<float>(x, s, c);
  else sincosf(x, &s, &c);
}
template <int ANG> __device__ __forceinline__ void vsincos(F2 x, F2& s, F2& c) {
  if (ANG == 3) {
    sincos_micro<F2>(x, s, c);
  } else if (ANG == 2) {
    sincos_tiny<F2>(x, s, c);
  } else if (ANG == 1) {
    sincos_poly<F2>(x, s, c);
  } else {
    float x0, x1, s0, c0, s1, c1;
    f2_unpack(x, x0, x1);
    sincosf(x0, &s0, &c0);
    sincosf(x1, &s1, &c1);
    s = f2_pack(s0, s1);
    c = f2_pack(c0, c1);
  }
}

// ---- float4 rows <-> lane values: component c of the L float4 rows a thread holds (one per env), as one lane value
template <class V> struct Pack;
template <> struct Pack<float> {
  static __device__ __forceinline__ float x(const float4* q) { return q[0].x; }
  static __device__ __forceinline__ float y(const float4* q) { return q[0].y; }
  static __device__ __forceinline__ float z(const float4* q) { return q[0].z; }
  static __device__ __forceinline__ float w(const float4* q) { return q[0].w; }
};
template <> struct Pack<F2> {
  static __device__ __forceinline__ F2 x(const float4* q) { return f2_pack(q[0].x, q[1].x); }
  static __device__ __forceinline__ F2 y(const float4* q) { return f2_pack(q[0].y, q[1].y); }
  static __device__ __forceinline__ F2 z(const float4* q) { return f2_pack(q[0].z, q[1].z); }
  static __device__ __forceinline__ F2 w(const float4* q) { return f2_pack(q[0].w, q[1].w); }
};

// replace lane l of a lane value
template <class V> __device__ __forceinline__ V lane_set(V v, int l, float x);
template <> __device__ __forceinline__ float lane_set<float>(float, int, float x) { return x; }
template <> __device__ __forceinline__ F2 lane_set<F2>(F2 v, int l, float x) {
  float a, b;
  f2_unpack(v, a, b);
  return l ? f2_pack(a, x) : f2_pack(x, b);
}

// 128-bit global access.  State/action streams are touched exactly once per control step, so they
// bypass L1 allocation (streaming) -- L2 still serves the re-reads of small batches.
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

}  // namespace fpv
