// oxg_math.cuh — device arithmetic in the reference's operation order.
//
// Every function here restates a piece of oxmpl's f64 arithmetic with explicitly
// rounded, never-fused operations (__dadd_rn / __dsub_rn / __dmul_rn; the library is
// also compiled with -fmad=false).  sqrt(), '/' and fmod() on doubles are IEEE
// correctly-rounded / exact in CUDA, as they are in Rust's std on the CPU.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace oxg {

#define OXG_PI 3.14159265358979323846264338327950288
#define OXG_DEV __device__ __forceinline__

// Dimension that is a compile-time constant when DT > 0, else a runtime value.
template <int DT>
struct DimOf {
  OXG_DEV static int get(int rt) { return DT > 0 ? DT : rt; }
};

// fmod(x, c), bit for bit, without libdevice's long-division loop on the common path.  fmod is
// an exact operation: r = x - q*c with q = trunc(x / c) is always representable, so one FMA from
// the right q produces it without rounding.  q from the rounded quotient can be off by one; the
// exact value is then r - c < 0 or r + c >= c, which survives the FMA's single rounding as such
// and is repaired from the original x.  |x| < c (angles already in range) returns x itself.
// Anything unusual (c <= 0, huge ratio, inf, NaN) takes libdevice's fmod.
OXG_DEV double fmod_exact_dev(double x, double c) {
  const double ax = fabs(x);
  if (!(c > 0.0) || !(ax < __dmul_rn(c, 4503599627370496.0))) return fmod(x, c);
  if (ax < c) return x;
  double q = trunc(__ddiv_rn(ax, c));
  double r = __fma_rn(-q, c, ax);
  if (r < 0.0) {
    q = __dsub_rn(q, 1.0);
    r = __fma_rn(-q, c, ax);
  } else if (r >= c) {
    q = __dadd_rn(q, 1.0);
    r = __fma_rn(-q, c, ax);
  }
  return copysign(r, x);
}

// f64::rem_euclid (Rust core): r = self % rhs; if r < 0.0 { r + rhs.abs() } else { r }
OXG_DEV double rem_euclid_dev(double v, double rhs) {
  double r = fmod_exact_dev(v, rhs);
  if (r < 0.0) r = __dadd_rn(r, fabs(rhs));
  return r;
}

// SO2State::normalise — oxmpl/src/base/states/so2_state.rs:55-59
OXG_DEV double so2_normalise_dev(double v) {
  return __dsub_rn(rem_euclid_dev(__dadd_rn(v, OXG_PI), __dmul_rn(2.0, OXG_PI)), OXG_PI);
}

// SO2StateSpace::distance — oxmpl/src/base/spaces/so2_state_space.rs:97-101
OXG_DEV double so2_distance_dev(double a, double b) {
  double diff = __dsub_rn(a, b);
  diff = __dsub_rn(rem_euclid_dev(__dadd_rn(diff, OXG_PI), __dmul_rn(2.0, OXG_PI)), OXG_PI);
  return fabs(diff);
}

// SO2StateSpace::interpolate — so2_state_space.rs:107-122
OXG_DEV double so2_interpolate_dev(double from, double to, double t) {
  double diff = __dsub_rn(so2_normalise_dev(to), so2_normalise_dev(from));
  if (diff > OXG_PI) {
    diff = __dsub_rn(diff, __dmul_rn(2.0, OXG_PI));
  } else if (diff < -OXG_PI) {
    diff = __dadd_rn(diff, __dmul_rn(2.0, OXG_PI));
  }
  double v = __dadd_rn(from, __dmul_rn(diff, t));
  return so2_normalise_dev(v);
}

// Sum of squared differences, RealVectorStateSpace::distance's pre-sqrt value —
// oxmpl/src/base/spaces/real_vector_state_space.rs:148-153 (left fold; 0.0 + x*x == x*x).
template <int DT>
OXG_DEV double rv_sumsq_dev(const double* a, const double* b, int n_rt) {
  const int n = DimOf<DT>::get(n_rt);
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < n; ++i) {
    double d = __dsub_rn(a[i], b[i]);
    double m = __dmul_rn(d, d);
    s = (i == 0) ? m : __dadd_rn(s, m);
  }
  return s;
}

// Space distance for runtime descriptors (see include/oxg.h, oxg_space_desc).
OXG_DEV double space_distance_dev(uint32_t n_real, uint32_t n_so2, double w_real, double w_so2,
                                  const double* a, const double* b) {
  if (n_so2 == 0) return sqrt(rv_sumsq_dev<0>(a, b, (int)n_real));
  if (n_real == 0 && n_so2 == 1) return so2_distance_dev(a[0], b[0]);
  double acc = 0.0;
  bool have = false;
  if (n_real > 0) {
    acc = __dmul_rn(w_real, sqrt(rv_sumsq_dev<0>(a, b, (int)n_real)));
    have = true;
  }
  for (uint32_t c = 0; c < n_so2; ++c) {
    double term = __dmul_rn(w_so2, so2_distance_dev(a[n_real + c], b[n_real + c]));
    acc = have ? __dadd_rn(acc, term) : term;
    have = true;
  }
  return acc;
}

// Rust `x.ceil() as usize`: saturating, NaN -> 0.
OXG_DEV unsigned long long ceil_as_usize_dev(double x) {
  double c = ceil(x);
  if (!(c == c)) return 0ull;
  if (c <= 0.0) return 0ull;
  if (c >= 18446744073709551615.0) return 0xFFFFFFFFFFFFFFFFull;
  return (unsigned long long)c;
}

// next / previous representable double for x >= +0 (bit pattern arithmetic).
OXG_DEV double next_up_nonneg(double x) {
  return __longlong_as_double(__double_as_longlong(x) + 1);
}
OXG_DEV double next_down_pos(double x) {
  return __longlong_as_double(__double_as_longlong(x) - 1);
}

// max{x >= 0 : sqrt(x) <= s} for s >= 0 (s = +inf -> +inf).  "sqrt(d2) <= s" <=> "d2 <= U".
// The preimage of one double under the correctly-rounded sqrt holds 1-3 doubles, and
// fl(s*s) is within one of them, so both loops run a handful of iterations.
OXG_DEV double sqrt_le_threshold_dev(double s) {
  if (!(s == s) || s < 0.0) return -1.0;
  double x = __dmul_rn(s, s);
  if (x == __longlong_as_double(0x7FF0000000000000ll)) {
    if (s == x) return x;                       // s = +inf
    x = __longlong_as_double(0x7FEFFFFFFFFFFFFFll);  // DBL_MAX
  }
  while (sqrt(x) > s && x > 0.0) x = next_down_pos(x);
  for (;;) {
    double nx = next_up_nonneg(x);
    if (nx == nx && sqrt(nx) <= s)
      x = nx;
    else
      break;
  }
  return x;
}

// ----------------------------------------------------------------- validity predicates
// World layouts are described in oxg_internal.cuh (oxg_world_s).

// sphere: valid iff sqrt(sum (x_i - c_i)^2) > r   (README.md:147-150)  <=>  sumsq > U
template <int DT>
OXG_DEV bool sphere_hit_dev(const double* x, const double* __restrict__ sph, int dim_rt) {
  const int dim = DimOf<DT>::get(dim_rt);
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < dim; ++i) {
    double d = __dsub_rn(x[i], __ldg(sph + i));
    double m = __dmul_rn(d, d);
    s = (i == 0) ? m : __dadd_rn(s, m);
  }
  return !(s > __ldg(sph + dim));
}

// aabb: invalid iff all_i lo_i <= x_i <= hi_i   (oxmpl/tests/rrt_rvss_tests.rs:24-36)
template <int DT>
OXG_DEV bool aabb_hit_dev(const double* x, const double* __restrict__ box, int dim_rt) {
  const int dim = DimOf<DT>::get(dim_rt);
  bool inside = true;
#pragma unroll
  for (int i = 0; i < dim; ++i) {
    double lo = __ldg(box + i), hi = __ldg(box + dim + i);
    inside = inside && (x[i] >= lo) && (x[i] <= hi);
  }
  return inside;
}

// capsule: valid iff dist(x, segment) > r, operation order as in oracle/oxg_oracle.cpp
// (capsule_valid).  ab_i and den are precomputed at world creation with the same roundings.
template <int DT>
OXG_DEV bool capsule_hit_dev(const double* x, const double* __restrict__ cap, int dim_rt) {
  const int dim = DimOf<DT>::get(dim_rt);
  double num = 0.0;
#pragma unroll
  for (int i = 0; i < dim; ++i) {
    double ap = __dsub_rn(x[i], __ldg(cap + i));
    double m = __dmul_rn(ap, __ldg(cap + dim + i));
    num = (i == 0) ? m : __dadd_rn(num, m);
  }
  double den = __ldg(cap + 2 * dim);
  double t = (den > 0.0) ? __ddiv_rn(num, den) : 0.0;
  if (t < 0.0) t = 0.0;
  if (t > 1.0) t = 1.0;
  double d2 = 0.0;
#pragma unroll
  for (int i = 0; i < dim; ++i) {
    double c = __dadd_rn(__ldg(cap + i), __dmul_rn(__ldg(cap + dim + i), t));
    double e = __dsub_rn(x[i], c);
    double m = __dmul_rn(e, e);
    d2 = (i == 0) ? m : __dadd_rn(d2, m);
  }
  return !(d2 > __ldg(cap + 2 * dim + 1));
}

// ----------------------------------------------------------------- certified FP32 pre-filters
// "maybe hit" tests in FP32 from the world's f32 mirrors: they may only answer NO when a
// rigorous bound on the FP32 evaluation error proves the exact FP64 predicate would answer NO
// too; a YES is always followed by the exact predicate above.  Per-edge constants (EdgeTol)
// carry the part of the bound that depends on the magnitude of the state's coordinates.
//
// Sphere (difference form, as the k-NN filter): with u = 2^-24, M >= |x_d|, |c_d|,
// a = 4uM: d2_32 <= [S + 2a sqrt(D S) + 2 D a^2](1 + D 2^-22) whenever the exact sum S <= U.
// Capsule: t32 is some point of [0,1], so the exact-arithmetic value at t32 is >= the true
// minimum and exceeds it by |ab|^2 (t32 - t*)^2 <= (D+8)^2 u^2 (|ap| + M)^2; the evaluation of
// e_d = x_d - a_d - ab_d t adds at most 8uM per coordinate.  Both are folded into
// T = U1 + a2 * sq + ce with U1 >= U(1+1e-4), sq >= sqrt(D U1) from the mirror and a2, ce
// from EdgeTol, every operation rounded up.
struct EdgeTol {
  float a2_s, ce_s;  // spheres
  float a2_c, ce_c;  // capsules
};

// ms = max |coordinate| of the states this tolerance covers (both endpoints of an edge).
OXG_DEV EdgeTol edge_tol_dev(double ms, double wmax_s, double wmax_c, int dim) {
  EdgeTol t;
  const float inf = __int_as_float(0x7F800000);
  const double Ms = fmax(ms, wmax_s), Mc = fmax(ms, wmax_c);
  if (!(Ms < 1.0e15) || !(ms == ms)) {
    t.a2_s = inf;
    t.ce_s = inf;
  } else {
    const double a = Ms * 2.384185791015625e-07 + 1e-37;  // 2^-22 M = 4uM
    t.a2_s = __double2float_ru(2.0 * a * 1.0001);
    t.ce_s = __double2float_ru((2.0 * dim * a * a + 1e-35) * 1.0001);
  }
  if (!(Mc < 1.0e15) || !(ms == ms)) {
    t.a2_c = inf;
    t.ce_c = inf;
  } else {
    const double a = Mc * 4.76837158203125e-07 + 1e-37;  // 2^-21 M = 8uM
    t.a2_c = __double2float_ru(2.0 * a * 1.0001);
    t.ce_c = __double2float_ru((2.0 * dim * a * a + 4e-10 * dim * Mc * Mc + 1e-35) * 1.0001);
  }
  return t;
}

template <int DT>
OXG_DEV bool sphere_maybe_hit32(const float* xf, const float* __restrict__ sph, float a2, float ce,
                                int dim_rt) {
  const int dim = DimOf<DT>::get(dim_rt);
  if (DT > 0) {
    constexpr int NV = ((DT > 0 ? DT : 1) + 2 + 3) / 4;  // row stride is a multiple of 4 floats
    float c[NV * 4];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(sph) + v);
      c[4 * v] = t4.x;
      c[4 * v + 1] = t4.y;
      c[4 * v + 2] = t4.z;
      c[4 * v + 3] = t4.w;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      const float d = xf[i] - c[i];
      s = __fmaf_rn(d, d, s);
    }
    const float T = __fmaf_ru(a2, c[DT + 1], __fadd_ru(c[DT], ce));
    return !(s > T);
  }
  float s = 0.f;
  for (int i = 0; i < dim; ++i) {
    const float d = xf[i] - __ldg(sph + i);
    s = __fmaf_rn(d, d, s);
  }
  const float T = __fmaf_ru(a2, __ldg(sph + dim + 1), __fadd_ru(__ldg(sph + dim), ce));
  return !(s > T);
}

template <int DT>
OXG_DEV bool aabb_maybe_hit32(const float* xf, const float* __restrict__ box, int dim_rt) {
  const int dim = DimOf<DT>::get(dim_rt);
  bool inside = true;
#pragma unroll
  for (int i = 0; i < dim; ++i)
    inside = inside && !(xf[i] < __ldg(box + i)) && !(xf[i] > __ldg(box + dim + i));
  return inside;
}

template <int DT>
OXG_DEV bool capsule_maybe_hit32(const float* xf, const float* __restrict__ cap, float a2, float ce,
                                 int dim_rt) {
  const int dim = DimOf<DT>::get(dim_rt);
  if (DT > 0) {
    // one aligned 16-byte load per four mirror values (the row stride is a multiple of 4 floats)
    constexpr int NV = (2 * (DT > 0 ? DT : 1) + 3 + 3) / 4;
    float c[NV * 4];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(cap) + v);
      c[4 * v] = t4.x;
      c[4 * v + 1] = t4.y;
      c[4 * v + 2] = t4.z;
      c[4 * v + 3] = t4.w;
    }
    float ap[DT > 0 ? DT : 1];
    float num = 0.f;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      ap[i] = xf[i] - c[i];
      num = __fmaf_rn(ap[i], c[DT + i], num);
    }
    float t = num * c[2 * DT];
    t = fminf(fmaxf(t, 0.f), 1.f);  // NaN -> 0: any t in [0,1] keeps the bound valid
    float d2 = 0.f;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
      const float e = __fmaf_rn(-t, c[DT + i], ap[i]);
      d2 = __fmaf_rn(e, e, d2);
    }
    const float T = __fmaf_ru(a2, c[2 * DT + 2], __fadd_ru(c[2 * DT + 1], ce));
    return !(d2 > T);
  }
  float ap[32];
  float num = 0.f;
  for (int i = 0; i < dim; ++i) {
    ap[i] = xf[i] - __ldg(cap + i);
    num = __fmaf_rn(ap[i], __ldg(cap + dim + i), num);
  }
  float t = num * __ldg(cap + 2 * dim);
  t = fminf(fmaxf(t, 0.f), 1.f);
  float d2 = 0.f;
  for (int i = 0; i < dim; ++i) {
    const float e = __fmaf_rn(-t, __ldg(cap + dim + i), ap[i]);
    d2 = __fmaf_rn(e, e, d2);
  }
  const float T = __fmaf_ru(a2, __ldg(cap + 2 * dim + 2), __fadd_ru(__ldg(cap + 2 * dim + 1), ce));
  return !(d2 > T);
}

// Variants reading the mirror row from shared memory (32-bit shared address, 16-byte loads):
// the validity kernels stage the world's f32 mirrors once per CTA, which turns the per-obstacle
// row reads from warp-uniform global loads into broadcast LDS.128.
__device__ __forceinline__ float4 lds128_row(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}

template <int DT>
OXG_DEV bool sphere_maybe_hit32_s(const float* xf, uint32_t row, float a2, float ce) {
  constexpr int DN = DT > 0 ? DT : 1;  // (never called with DT == 0; keeps the arrays non-empty)
  constexpr int NV = (DN + 2 + 3) / 4;
  float c[NV * 4];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 t4 = lds128_row(row + 16u * v);
    c[4 * v] = t4.x;
    c[4 * v + 1] = t4.y;
    c[4 * v + 2] = t4.z;
    c[4 * v + 3] = t4.w;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    const float d = xf[i] - c[i];
    s = __fmaf_rn(d, d, s);
  }
  const float T = __fmaf_ru(a2, c[DT + 1], __fadd_ru(c[DT], ce));
  return !(s > T);
}

template <int DT>
OXG_DEV bool capsule_maybe_hit32_s(const float* xf, uint32_t row, float a2, float ce) {
  constexpr int DN = DT > 0 ? DT : 1;
  constexpr int NV = (2 * DN + 3 + 3) / 4;
  float c[NV * 4];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 t4 = lds128_row(row + 16u * v);
    c[4 * v] = t4.x;
    c[4 * v + 1] = t4.y;
    c[4 * v + 2] = t4.z;
    c[4 * v + 3] = t4.w;
  }
  float ap[DN];
  float num = 0.f;
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    ap[i] = xf[i] - c[i];
    num = __fmaf_rn(ap[i], c[DT + i], num);
  }
  float t = num * c[2 * DT];
  t = fminf(fmaxf(t, 0.f), 1.f);
  float d2 = 0.f;
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    const float e = __fmaf_rn(-t, c[DT + i], ap[i]);
    d2 = __fmaf_rn(e, e, d2);
  }
  const float T = __fmaf_ru(a2, c[2 * DT + 2], __fadd_ru(c[2 * DT + 1], ce));
  return !(d2 > T);
}

// Two capsules per call with packed FP32 arithmetic (FADD2 / FFMA2 / FMUL2): `row` is the shared
// address of a pair row of the interleaved mirror (oxg_world_s::d_capsules32p).  Returns bit 0 /
// bit 1 = "maybe hit" for the first / second capsule of the pair.
template <int DT>
OXG_DEV unsigned capsule_pair_maybe_hit32_s(const float* xf, uint32_t row, float a2, float ce) {
  constexpr int DN = DT > 0 ? DT : 1;
  constexpr int NV = (2 * (2 * DN + 3) + 3) / 4;
  float2 c[NV * 2];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 t4 = lds128_row(row + 16u * v);
    c[2 * v] = make_float2(t4.x, t4.y);
    c[2 * v + 1] = make_float2(t4.z, t4.w);
  }
  float2 ap[DN];
  float2 num = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < DN; ++i) {
    ap[i] = __fadd2_rn(make_float2(xf[i], xf[i]), make_float2(-c[i].x, -c[i].y));
    num = __ffma2_rn(ap[i], c[DN + i], num);
  }
  float2 t = __fmul2_rn(num, c[2 * DN]);
  t.x = fminf(fmaxf(t.x, 0.f), 1.f);  // NaN -> 0: any t in [0,1] keeps the bound valid
  t.y = fminf(fmaxf(t.y, 0.f), 1.f);
  const float2 nt = make_float2(-t.x, -t.y);
  float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < DN; ++i) {
    const float2 e = __ffma2_rn(nt, c[DN + i], ap[i]);
    d2 = __ffma2_rn(e, e, d2);
  }
  const float T0 = __fmaf_ru(a2, c[2 * DN + 2].x, __fadd_ru(c[2 * DN + 1].x, ce));
  const float T1 = __fmaf_ru(a2, c[2 * DN + 2].y, __fadd_ru(c[2 * DN + 1].y, ce));
  return (!(d2.x > T0) ? 1u : 0u) | (!(d2.y > T1) ? 2u : 0u);
}

// ----------------------------------------------------------------- whole-edge culling
// "Could ANY state within rho of the point m hit this obstacle?"  Every interpolated state of
// check_motion(from, to) lies within half the edge length of the midpoint, so one such test per
// (edge, obstacle) replaces one test per (step, obstacle) for the obstacles it rules out.  Same
// certified bound as the per-state filters with the obstacle grown by rho: a state x with
// |x - m| <= rho' (rho' the exact half length plus the interpolation roundings; rho >= rho'
// (1 + 1e-5) + 1e-6 M) can only hit if the exact sum at m is <= (sqrt(U) + rho')^2 =: S_fat, and
// then d2_32(m) <= [S_fat + 2a sqrt(D S_fat) + 2 D a^2](1 + D 2^-22) <= T_fat below, where
// rr >= sqrt(U1) + rho is taken from the mirror's sq >= sqrt(D U1).
struct FatTol {
  float rho;        // ball radius around the midpoint, rounded up
  float sqrt_d;     // >= sqrt(D)
  float inv_sqrt_d; // >= 1 / sqrt(D)
};

OXG_DEV float fat_threshold(float sq, float a2, float ce, const FatTol& f) {
  const float rr = __fmaf_ru(sq, f.inv_sqrt_d, f.rho);
  return __fmaf_ru(rr, rr, __fmaf_ru(a2, __fmul_ru(f.sqrt_d, rr), ce));
}

template <int DT>
OXG_DEV bool sphere_fat_maybe_hit32_s(const float* mf, uint32_t row, float a2, float ce, const FatTol& f) {
  constexpr int DN = DT > 0 ? DT : 1;
  constexpr int NV = (DN + 2 + 3) / 4;
  float c[NV * 4];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 t4 = lds128_row(row + 16u * v);
    c[4 * v] = t4.x;
    c[4 * v + 1] = t4.y;
    c[4 * v + 2] = t4.z;
    c[4 * v + 3] = t4.w;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < DT; ++i) {
    const float d = mf[i] - c[i];
    s = __fmaf_rn(d, d, s);
  }
  return !(s > fat_threshold(c[DT + 1], a2, ce, f));
}

// AABB mirror row in global memory: [lo.. rounded down, hi.. rounded up].
template <int DT>
OXG_DEV bool aabb_fat_maybe_hit32(const float* mf, const float* __restrict__ box, const FatTol& f) {
  bool near = true;
#pragma unroll
  for (int i = 0; i < DT; ++i)
    near = near && !(mf[i] < __fsub_rd(__ldg(box + i), f.rho)) && !(mf[i] > __fadd_ru(__ldg(box + DT + i), f.rho));
  return near;
}

// Pair row of the interleaved capsule mirror (see capsule_pair_maybe_hit32_s): bit 0 / bit 1.
template <int DT>
OXG_DEV unsigned capsule_pair_fat_maybe_hit32_s(const float* mf, uint32_t row, float a2, float ce,
                                                const FatTol& f) {
  constexpr int DN = DT > 0 ? DT : 1;
  constexpr int NV = (2 * (2 * DN + 3) + 3) / 4;
  float2 c[NV * 2];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 t4 = lds128_row(row + 16u * v);
    c[2 * v] = make_float2(t4.x, t4.y);
    c[2 * v + 1] = make_float2(t4.z, t4.w);
  }
  float2 ap[DN];
  float2 num = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < DN; ++i) {
    ap[i] = __fadd2_rn(make_float2(mf[i], mf[i]), make_float2(-c[i].x, -c[i].y));
    num = __ffma2_rn(ap[i], c[DN + i], num);
  }
  float2 t = __fmul2_rn(num, c[2 * DN]);
  t.x = fminf(fmaxf(t.x, 0.f), 1.f);  // NaN -> 0: any t in [0,1] keeps the bound valid
  t.y = fminf(fmaxf(t.y, 0.f), 1.f);
  const float2 nt = make_float2(-t.x, -t.y);
  float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < DN; ++i) {
    const float2 e = __ffma2_rn(nt, c[DN + i], ap[i]);
    d2 = __ffma2_rn(e, e, d2);
  }
  const float T0 = fat_threshold(c[2 * DN + 2].x, a2, ce, f);
  const float T1 = fat_threshold(c[2 * DN + 2].y, a2, ce, f);
  return (!(d2.x > T0) ? 1u : 0u) | (!(d2.y > T1) ? 2u : 0u);
}

// One capsule of a pair row (half = 0 / 1): the per-state filter for the few capsules that
// survive the edge cull.
template <int DT>
OXG_DEV bool capsule_half_maybe_hit32_s(const float* xf, uint32_t row, unsigned half, float a2, float ce) {
  constexpr int DN = DT > 0 ? DT : 1;
  constexpr int NV = (2 * (2 * DN + 3) + 3) / 4;
  float c[NV * 2];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float4 t4 = lds128_row(row + 16u * v);
    c[2 * v] = half ? t4.y : t4.x;
    c[2 * v + 1] = half ? t4.w : t4.z;
  }
  float ap[DN];
  float num = 0.f;
#pragma unroll
  for (int i = 0; i < DN; ++i) {
    ap[i] = xf[i] - c[i];
    num = __fmaf_rn(ap[i], c[DN + i], num);
  }
  float t = num * c[2 * DN];
  t = fminf(fmaxf(t, 0.f), 1.f);
  float d2 = 0.f;
#pragma unroll
  for (int i = 0; i < DN; ++i) {
    const float e = __fmaf_rn(-t, c[DN + i], ap[i]);
    d2 = __fmaf_rn(e, e, d2);
  }
  const float T = __fmaf_ru(a2, c[2 * DN + 2], __fadd_ru(c[2 * DN + 1], ce));
  return !(d2 > T);
}

}  // namespace oxg
