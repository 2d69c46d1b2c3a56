// csrc/imp_fast.cuh -- certified fast arithmetic for the Gauss-Seidel impulse chain.
//
// The impulse sweep is one dependency chain (response.cu); per hop more than half of its latency is fp64 division
// (125 cycles on sm_100a) and square root (220 cycles) -- tools/ubench.cu.  The helpers below produce the SAME
// correctly rounded results with shorter dependent chains:
//   * a quotient a / b whose divisor is known early (or is a constant) is formed from an early reciprocal estimate
//     with three dependent multiply-adds (Markstein's correction);
//   * a square root and the reciprocal of the rounded root come from one coupled Newton iteration;
//   * comparisons of a square root against a constant are done on the radicand against a host-computed threshold
//     (sqrt is monotone and correctly rounded, so `sqrt(x) < c` is exactly `x < min{x : sqrt(x) >= c}`).
// Every fast result carries a CERTIFICATE: the exact residual (one fused multiply-add) proves that the candidate is
// the round-to-nearest value of the real quotient / root, with a 0.1 % safety margin against the residual's own
// rounding and with exponent-range and power-of-two guards.  A result that cannot be certified clears `ok`; the
// caller then recomputes the whole element with the plain IEEE operations (the reference arithmetic).  The fast path
// therefore never changes a bit -- it is identical when certified and absent otherwise.  Explicit __fma_rn is used
// only inside these correction / certification steps (never for the reference's own a * b + c expressions).
#pragma once
#include "dev_math.cuh"

namespace cg {

// values whose binary exponent lies in [-400, 400]: products of two such values and their ulps neither overflow nor
// underflow
__device__ __forceinline__ bool exp_safe(double x) {
  const unsigned e = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu;
  return e - 623u <= 800u;
}
// unit in the last place of a normal x (spacing to the next larger magnitude)
__device__ __forceinline__ double ulp_of(double x) { return __hiloint2double((__double2hiint(x) & 0x7ff00000) - (52 << 20), 0); }
// x is a power of two (the spacing below it is half the spacing above: excluded from certification)
__device__ __forceinline__ bool mant_zero(double x) { return ((__double2hiint(x) & 0x000fffff) | __double2loint(x)) == 0; }

// estimate of 1 / b, accurate to about one ulp (NOT necessarily correctly rounded)
__device__ __forceinline__ double rcp_newton(double b) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
  double e = __fma_rn(-b, y, 1.0);
  y = __fma_rn(y, e, y);
  e = __fma_rn(-b, y, 1.0);
  y = __fma_rn(y, e, y);
  return y;
}

// RN(a / b) from an estimate y of 1 / b.  Certificate: r = a - b q (exact up to one rounding) with
// |r| < |b| ulp(q) / 2 means a / b lies strictly inside the rounding interval of q (a tie is impossible: a / b is
// never the midpoint of two neighbouring doubles unless b is a power of two, where the quotient is exact).
__device__ __forceinline__ double div_cert(double a, double b, double y, bool& ok) {
  const double q0 = a * y;
  const double r0 = __fma_rn(-b, q0, a);
  const double q1 = __fma_rn(r0, y, q0);
  const double r1 = __fma_rn(-b, q1, a);
  ok = ok && exp_safe(q1) && exp_safe(b) && !mant_zero(q1) && fabs(r1) < 0.4995 * (fabs(b) * ulp_of(q1));
  return q1;
}
// RN(1 / b) from an estimate y
__device__ __forceinline__ double rcp_cert(double b, double y, bool& ok) {
  const double e = __fma_rn(-b, y, 1.0);
  const double y1 = __fma_rn(e, y, y);
  const double e1 = __fma_rn(-b, y1, 1.0);
  ok = ok && exp_safe(b) && !mant_zero(y1) && fabs(e1) < 0.4995 * (fabs(b) * ulp_of(y1));
  return y1;
}
// g = RN(sqrt(s)); h ~ 1 / (2 g) comes out of the same iteration.  Certificate: d = s - g^2 (exact up to one rounding)
// with |d| < g ulp(g) puts s strictly between the squares of the two rounding boundaries g -+ ulp(g) / 2.
__device__ __forceinline__ double sqrt_cert(double s, double& h_out, bool& ok) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  double g = s * y, h = 0.5 * y;
  double r = __fma_rn(-g, h, 0.5);
  g = __fma_rn(g, r, g);
  h = __fma_rn(h, r, h);
  r = __fma_rn(-g, h, 0.5);
  g = __fma_rn(g, r, g);
  h = __fma_rn(h, r, h);
  double d = __fma_rn(-g, g, s);
  g = __fma_rn(d, h, g);
  d = __fma_rn(-g, g, s);
  ok = ok && exp_safe(s) && !mant_zero(g) && fabs(d) < 0.999 * (g * ulp_of(g));
  h_out = h;
  return g;
}
// normalized(a) (vector3d.h:97-103): a * (1.0 / sqrt(a.a))
__device__ __forceinline__ V3 normalized_cert(V3 a, bool& ok) {
  const double s = a.x * a.x + a.y * a.y + a.z * a.z;
  double h;
  const double g = sqrt_cert(s, h, ok);
  const double invmag = rcp_cert(g, h + h, ok);
  return v3(a.x * invmag, a.y * invmag, a.z * invmag);
}

// launch constants of an impulse sweep (host computed, see make_imp_const in response.cu)
struct ImpConst {
  double delta, stiffness, dt, mass;
  double inv_mass;   // RN(1 / mass)
  double thr_ccd;    // 0.1 * delta / dt
  double par2;       // smallest x with sqrt(x) >= 1e-10 : length(v) < 1e-10  <=>  v.v < par2
  double dist2;      // largest x with sqrt(x) <= delta  : length(v) > delta  <=>  v.v > dist2
  int fast;          // 0: reference arithmetic only (CLOTHGPU_IMP_FAST=0)
  int lanes;         // working lanes per warp of the data-flow kernel (CLOTHGPU_IMP_LANES)
  int sleep_ns;      // pause between two polls of a predecessor's flag (CLOTHGPU_IMP_SLEEP)
};

// FindCoplanerInterp (bvh_aabb.cpp:598-662) as in coplanar_time_inl<true>; the roots of the derivative come from the
// certified square root and quotients (the reciprocal of 3 k3 is formed while the root is computed).  The branch
// ladder only selects the bracket; the bisection runs once at the end (one copy of its code instead of nine).
static __device__ __forceinline__ double coplanar_time_cert(V3 s0, V3 s1, V3 s2, V3 s3, V3 e0, V3 e1, V3 e2, V3 e3, bool& ok) {
  const V3 x1 = s1 - s0, x2 = s2 - s0, x3 = s3 - s0;
  const V3 v1 = (e1 - e0) - x1, v2 = (e2 - e0) - x2, v3v = (e3 - e0) - x3;
  const double k0 = triple(x3, x1, x2);
  const double k1 = triple(v3v, x1, x2) + triple(x3, v1, x2) + triple(x3, x1, v2);
  const double k2 = triple(v3v, v1, x2) + triple(v3v, x1, v2) + triple(x3, v1, v2);
  const double k3 = triple(v3v, v1, v2);
  const double r0 = -0.0, r1 = +1.0;
  const double f0 = cubic(r0, k0, k1, k2, k3), f1 = cubic(r1, k0, k1, k2, k3);
  const double det = k2 * k2 - 3 * k1 * k3;
  double ba = 0, bb = 0, bf = 0;  // bracket [ba, bb] and the value at its left end
  bool found = false;
#define CG_BRACKET(A_, B_, F_) do { ba = (A_); bb = (B_); bf = (F_); found = true; } while (0)
  if (fabs(k3) < 1.0e-10 && fabs(k2) > 1.0e-10) {
    const double r2 = -k1 / (2 * k2);
    const double f2 = cubic(r2, k0, k1, k2, k3);
    if (r2 > 0 && r2 < 1) {
      if (f0 * f2 < 0) CG_BRACKET(r0, r2, f0);
      else if (f2 * f1 < 0) CG_BRACKET(r2, r1, f2);
    }
  }
  if (!found && det > 0 && fabs(k3) > 1.0e-10) {
    const double b3 = 3 * k3;
    const double y3 = rcp_newton(b3);
    double hh;
    const double sq = sqrt_cert(det, hh, ok);
    const double r3 = div_cert(-k2 - sq, b3, y3, ok);
    const double f3 = cubic(r3, k0, k1, k2, k3);
    if (r3 > 0 && r3 < 1) {
      if (f0 * f3 < 0) CG_BRACKET(r0, r3, f0);
      else if (f3 * f1 < 0) CG_BRACKET(r3, r1, f3);
    }
    if (!found) {
      const double r4 = div_cert(-k2 + sq, b3, y3, ok);
      const double f4 = cubic(r4, k0, k1, k2, k3);
      if (r3 > 0 && r3 < 1 && r4 > 0 && r4 < 1) {
        if (f3 * f4 < 0) CG_BRACKET(r3, r4, f3);
      }
      if (!found && r4 > 0 && r4 < 1) {
        if (f0 * f4 < 0) CG_BRACKET(r0, r4, f0);
        else if (f4 * f1 < 0) CG_BRACKET(r4, r1, f4);
      }
    }
  }
  if (!found) {
    if (f0 * f1 > 0) return -1;
    CG_BRACKET(r0, r1, f0);
  }
#undef CG_BRACKET
  return cubic_root_spec(ba, bb, bf, k0, k1, k2, k3);
}

}  // namespace cg
