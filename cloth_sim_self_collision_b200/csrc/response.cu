// csrc/response.cu -- collision response: repulsion / CCD impulses and rigid impact zones.
//
// Replaces SelfCollisionImpulse_Proximity (self_collision_cloth.cpp:20-112), SelfCollisionImpulse_CCD (:116-212),
// MakeRigidImpactZone (:220-274) and ApplyRigidImpactZone (:298-348).
//
// Impulses: the reference applies the elements one after the other in std::set order and every element reads the
// velocities written by the earlier ones (Gauss-Seidel).  Two elements only interact through a shared vertex, so
// the serial order is equivalent to any order that keeps, for every vertex, the relative order of the elements
// touching it.  The predecessor of element k on each of its 4 vertices (the previous element touching the vertex) comes
// from one stable radix sort of (vertex, element * 4 + corner) pairs; k_imp_flow (the default) starts an element as
// soon as its predecessors have finished, k_imp_run levels the dependency DAG and runs level after level
// (CLOTHGPU_IMPULSE=levels).  The divisions / square roots on the dependency chain use the certified fast forms of
// imp_fast.cuh.  Bit-identical to the serial loop.
//
// Impact zones: the reference's zone list equals the connected components of the member vertices under
// "same contact element" and "mesh-adjacent and both members" (order independent, verified against the
// reference); a lock-free union-find in global memory builds them.  ApplyRigidImpactZone sums over each zone in
// ascending vertex id; a stable sort by zone label lays the zones out as segments in that order and each zone is
// reduced sequentially, so the sums are bit-identical too.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>

#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cmath>

#include "ctx.h"
#include "dev_math.cuh"
#include "imp_fast.cuh"

namespace cg {

// v[i] += n * a * b evaluated as (n*a)*b, like `-norm.x*imp_mod*w0` in the reference
__device__ __forceinline__ void add3(double* v, int i, V3 n, double a, double b) {
  v[i * 3 + 0] += n.x * a * b;
  v[i * 3 + 1] += n.y * a * b;
  v[i * 3 + 2] += n.z * a * b;
}
// the same updates from velocities that are already in registers (the four vertices of an element are distinct, so
// the value in memory is the one loaded at the start of the element): no reload behind the previous vertex's stores
__device__ __forceinline__ void upd3(double* v, int i, V3 old, V3 n, double a, double b) {
  v[i * 3 + 0] = old.x + n.x * a * b;
  v[i * 3 + 1] = old.y + n.y * a * b;
  v[i * 3 + 2] = old.z + n.z * a * b;
}
__device__ __forceinline__ void upd3s(double* v, int i, V3 old, V3 n, double a) {
  v[i * 3 + 0] = old.x + n.x * a;
  v[i * 3 + 1] = old.y + n.y * a;
  v[i * 3 + 2] = old.z + n.z * a;
}
__device__ __forceinline__ V3 upd3v(V3 old, V3 n, double a, double b) { return v3(old.x + n.x * a * b, old.y + n.y * a * b, old.z + n.z * a * b); }
__device__ __forceinline__ V3 upd3sv(V3 old, V3 n, double a) { return v3(old.x + n.x * a, old.y + n.y * a, old.z + n.z * a); }
__device__ __forceinline__ void st3(double* v, int i, V3 a) { v[i * 3 + 0] = a.x; v[i * 3 + 1] = a.y; v[i * 3 + 2] = a.z; }
__device__ __forceinline__ void add3s(double* v, int i, V3 n, double a) {
  v[i * 3 + 0] += n.x * a;
  v[i * 3 + 1] += n.y * a;
  v[i * 3 + 2] += n.z * a;
}

// one element of SelfCollisionImpulse_Proximity, self_collision_cloth.cpp:31-110, in two parts: everything that only
// depends on the positions (closest-point weights, normal, penetration depth) and the velocity update.  The data-flow
// kernel evaluates the first part while it waits for the element's predecessors.
struct PrxGeom {
  bool ok, fv;
  double wa, wb, wc;  // FV: w0, w1, w2 ; EE: w01, w23 (wc unused)
  V3 norm;
  double p_depth;
  // velocity-independent pieces of the update, evaluated while the element waits for its predecessors
  double thr;     // 0.1 * p_depth / dt
  double imp_el;  // dt * stiffness * p_depth
  double den;     // 1 + w0^2 + w1^2 + w2^2  /  w01^2 + (1 - w01)^2 + w23^2 + (1 - w23)^2
  double yden;    // estimate of 1 / den for the certified quotient
};
__device__ __forceinline__ void impulse_prx_geom(const int* n, const double* xyz, double delta, PrxGeom& G) {
  const V3 p0 = ld3(xyz, n[0]), p1 = ld3(xyz, n[1]), p2 = ld3(xyz, n[2]), p3 = ld3(xyz, n[3]);
  G.ok = false;
  G.fv = n[4] != 0;
  if (G.fv) {
    double w0, w1;
    const double dist = dist_face_vertex(p0, p1, p2, p3, w0, w1);
    if (w0 < 0 || w0 > 1) return;
    if (w1 < 0 || w1 > 1) return;
    if (dist > delta) return;
    const double w2 = 1.0 - w0 - w1;
    const V3 pc = (w0 * p0 + w1 * p1) + w2 * p2;
    G.norm = normalized(p3 - pc);
    G.p_depth = delta - dot(p3 - pc, G.norm);
    G.wa = w0; G.wb = w1; G.wc = w2;
  } else {
    double w01, w23;
    const double dist = dist_edge_edge(p0, p1, p2, p3, w01, w23);
    if (w01 < 0 || w01 > 1) return;
    if (w23 < 0 || w23 > 1) return;
    if (dist > delta) return;
    const V3 c01 = (1 - w01) * p0 + w01 * p1;
    const V3 c23 = (1 - w23) * p2 + w23 * p3;
    G.norm = normalized(c23 - c01);
    G.p_depth = delta - length(c23 - c01);
    G.wa = w01; G.wb = w23; G.wc = 0.0;
  }
  G.ok = true;
}
// the velocity-independent scalars of impulse_prx_apply (same expressions, same operand order)
__device__ __forceinline__ void impulse_prx_geom_more(PrxGeom& G, double stiffness, double dt) {
  if (!G.ok) return;
  G.thr = 0.1 * G.p_depth / dt;
  G.imp_el = dt * stiffness * G.p_depth;
  if (G.fv) G.den = 1 + G.wa * G.wa + G.wb * G.wb + G.wc * G.wc;
  else G.den = G.wa * G.wa + (1 - G.wa) * (1 - G.wa) + G.wb * G.wb + (1 - G.wb) * (1 - G.wb);
  G.yden = rcp_newton(G.den);
}
// impulse_prx_apply on velocities held in registers, with the two divisions of the chain (2 imp / den, / mass) taken
// from early reciprocals; a quotient that cannot be certified is recomputed with the plain divisions (certified =
// false).  v[0..3] in: the current velocities of the element's vertices, out: the updated ones.  Returns true if the
// element changes the velocities.
__device__ __forceinline__ bool impulse_prx_apply_fast_v(const PrxGeom& G, V3* v, const ImpConst& K, bool& certified) {
  certified = true;
  if (!G.ok) return false;
  const V3 v0 = v[0], v1 = v[1], v2 = v[2], v3v = v[3];
  const V3 norm = G.norm;
  double rel_v;
  if (G.fv) rel_v = dot(((v3v - G.wa * v0) - G.wb * v1) - G.wc * v2, norm);
  else rel_v = dot((((1 - G.wb) * v2 + G.wb * v3v) - (1 - G.wa) * v0) - G.wa * v1, norm);
  if (rel_v > G.thr) return false;
  const double imp_ie = K.mass * (G.thr - rel_v);
  const double imp_min = (G.imp_el < imp_ie) ? G.imp_el : imp_ie;
  bool ok = true;
  double imp_mod = div_cert(2 * imp_min, G.den, G.yden, ok);
  imp_mod = div_cert(imp_mod, K.mass, K.inv_mass, ok);
  if (!ok) {
    imp_mod = 2 * imp_min / G.den;
    imp_mod /= K.mass;
    certified = false;
  }
  imp_mod *= 0.25;
  if (G.fv) {
    v[0] = upd3v(v0, -norm, imp_mod, G.wa);
    v[1] = upd3v(v1, -norm, imp_mod, G.wb);
    v[2] = upd3v(v2, -norm, imp_mod, G.wc);
    v[3] = upd3sv(v3v, norm, imp_mod);
  } else {
    v[0] = upd3v(v0, -norm, imp_mod, 1 - G.wa);
    v[1] = upd3v(v1, -norm, imp_mod, G.wa);
    v[2] = upd3v(v2, norm, imp_mod, 1 - G.wb);
    v[3] = upd3v(v3v, norm, imp_mod, G.wb);
  }
  return true;
}
// the same on velocities in global memory.  Returns false if the plain divisions had to be used.
__device__ __forceinline__ bool impulse_prx_apply_fast(const int* n, const PrxGeom& G, double* uvw, const ImpConst& K) {
  if (!G.ok) return true;
  V3 v[4] = {ld3(uvw, n[0]), ld3(uvw, n[1]), ld3(uvw, n[2]), ld3(uvw, n[3])};
  bool certified;
  if (impulse_prx_apply_fast_v(G, v, K, certified)) {
#pragma unroll
    for (int a = 0; a < 4; a++) st3(uvw, n[a], v[a]);
  }
  return certified;
}
__device__ __forceinline__ void impulse_prx_apply(const int* n, const PrxGeom& G, double* uvw, double stiffness, double dt,
                                                  double mass) {
  if (!G.ok) return;
  const V3 v0 = ld3(uvw, n[0]), v1 = ld3(uvw, n[1]), v2 = ld3(uvw, n[2]), v3v = ld3(uvw, n[3]);
  const V3 norm = G.norm;
  const double p_depth = G.p_depth;
  if (G.fv) {
    const double w0 = G.wa, w1 = G.wb, w2 = G.wc;
    const double rel_v = dot(((v3v - w0 * v0) - w1 * v1) - w2 * v2, norm);
    if (rel_v > 0.1 * p_depth / dt) return;
    const double imp_el = dt * stiffness * p_depth;
    const double imp_ie = mass * (0.1 * p_depth / dt - rel_v);
    const double imp_min = (imp_el < imp_ie) ? imp_el : imp_ie;
    double imp_mod = 2 * imp_min / (1 + w0 * w0 + w1 * w1 + w2 * w2);
    imp_mod /= mass;
    imp_mod *= 0.25;
    add3(uvw, n[0], -norm, imp_mod, w0);
    add3(uvw, n[1], -norm, imp_mod, w1);
    add3(uvw, n[2], -norm, imp_mod, w2);
    add3s(uvw, n[3], norm, imp_mod);
  } else {
    const double w01 = G.wa, w23 = G.wb;
    const double rel_v = dot((((1 - w23) * v2 + w23 * v3v) - (1 - w01) * v0) - w01 * v1, norm);
    if (rel_v > 0.1 * p_depth / dt) return;
    const double imp_el = dt * stiffness * p_depth;
    const double imp_ie = mass * (0.1 * p_depth / dt - rel_v);
    const double imp_min = (imp_el < imp_ie) ? imp_el : imp_ie;
    double imp_mod = 2 * imp_min / (w01 * w01 + (1 - w01) * (1 - w01) + w23 * w23 + (1 - w23) * (1 - w23));
    imp_mod /= mass;
    imp_mod *= 0.25;
    add3(uvw, n[0], -norm, imp_mod, 1 - w01);
    add3(uvw, n[1], -norm, imp_mod, w01);
    add3(uvw, n[2], norm, imp_mod, 1 - w23);
    add3(uvw, n[3], norm, imp_mod, w23);
  }
}
__device__ void impulse_prx_one(const int* n, const double* xyz, double* uvw, double delta, double stiffness, double dt,
                                double mass) {
  PrxGeom G;
  impulse_prx_geom(n, xyz, delta, G);
  impulse_prx_apply(n, G, uvw, stiffness, dt, mass);
}

// one element of SelfCollisionImpulse_CCD, self_collision_cloth.cpp:127-210.  Reference quirk (:141): the impact
// time is re-solved with the UN-scaled velocity (end points p + v, not p + dt*v).
__device__ __forceinline__ void impulse_ccd_apply(const int* n, V3 p0, V3 p1, V3 p2, V3 p3, double* uvw, double delta, double dt,
                                  double mass);
__device__ void impulse_ccd_one(const int* n, const double* xyz, double* uvw, double delta, double dt, double mass) {
  impulse_ccd_apply(n, ld3(xyz, n[0]), ld3(xyz, n[1]), ld3(xyz, n[2]), ld3(xyz, n[3]), uvw, delta, dt, mass);
}
// positions are passed in: the data-flow kernel loads them while it waits for the element's predecessors
__device__ __forceinline__ void impulse_ccd_apply(const int* n, V3 p0, V3 p1, V3 p2, V3 p3, double* uvw, double delta, double dt,
                                  double mass) {
  const V3 v0 = ld3(uvw, n[0]), v1 = ld3(uvw, n[1]), v2 = ld3(uvw, n[2]), v3v = ld3(uvw, n[3]);
  const double t = coplanar_time_inl<true>(p0, p1, p2, p3, p0 + v0, p1 + v1, p2 + v2, p3 + v3v);
  if (t < 0 || t > 1) return;
  const V3 p0m = p0 + t * v0, p1m = p1 + t * v1, p2m = p2 + t * v2, p3m = p3 + t * v3v;
  if (n[4]) {
    double w0, w1;
    {
      const double dist = dist_face_vertex(p0m, p1m, p2m, p3m, w0, w1);
      if (w0 < 0 || w0 > 1) return;
      if (w1 < 0 || w1 > 1) return;
      if (dist > delta) return;
    }
    const double w2 = 1.0 - w0 - w1;
    const V3 pc = (w0 * p0 + w1 * p1) + w2 * p2;
    const V3 norm = normalized(p3 - pc);
    const double rel_v = dot(((v3v - w0 * v0) - w1 * v1) - w2 * v2, norm);
    if (rel_v > 0.1 * delta / dt) return;
    const double imp = mass * (0.1 * delta / dt - rel_v);
    double imp_mod = 2 * imp / (1.0 + w0 * w0 + w1 * w1 + w2 * w2);
    imp_mod /= mass;
    imp_mod *= 0.1;
    add3(uvw, n[0], -norm, imp_mod, w0);
    add3(uvw, n[1], -norm, imp_mod, w1);
    add3(uvw, n[2], -norm, imp_mod, w2);
    add3s(uvw, n[3], norm, imp_mod);
  } else {
    double w01, w23;
    {
      const double dist = dist_edge_edge(p0m, p1m, p2m, p3m, w01, w23);
      if (w01 < 0 || w01 > 1) return;
      if (w23 < 0 || w23 > 1) return;
      if (dist > delta) return;
    }
    const V3 c01 = (1 - w01) * p0 + w01 * p1;
    const V3 c23 = (1 - w23) * p2 + w23 * p3;
    const V3 norm = normalized(c23 - c01);
    const double rel_v = dot((((1 - w23) * v2 + w23 * v3v) - (1 - w01) * v0) - w01 * v1, norm);
    if (rel_v > 0.1 * delta / dt) return;
    const double imp = mass * (0.1 * delta / dt - rel_v);
    double imp_mod = 2 * imp / (w01 * w01 + (1 - w01) * (1 - w01) + w23 * w23 + (1 - w23) * (1 - w23));
    imp_mod /= mass;
    imp_mod *= 0.1;
    add3(uvw, n[0], -norm, imp_mod, 1 - w01);
    add3(uvw, n[1], -norm, imp_mod, w01);
    add3(uvw, n[2], norm, imp_mod, 1 - w23);
    add3(uvw, n[3], norm, imp_mod, w23);
  }
}

// impulse_ccd_apply with certified fast arithmetic (imp_fast.cuh): same expressions, same operand order; the square
// roots / quotients of the chain are replaced by their certified forms and the two distance gates by comparisons of
// the squared distance.  Returns 0: certified, nothing written; 1: certified, velocities written; 2: not certified,
// nothing written -- the caller runs impulse_ccd_apply.
__device__ __forceinline__ int impulse_ccd_fast_v(int is_fv, V3 p0, V3 p1, V3 p2, V3 p3, V3* v, const ImpConst& K) {
  bool ok = true;
  const V3 v0 = v[0], v1 = v[1], v2 = v[2], v3v = v[3];
  const double t = coplanar_time_cert(p0, p1, p2, p3, p0 + v0, p1 + v1, p2 + v2, p3 + v3v, ok);
  if (t < 0 || t > 1) return ok ? 0 : 2;
  const V3 p0m = p0 + t * v0, p1m = p1 + t * v1, p2m = p2 + t * v2, p3m = p3 + t * v3v;
  const double mass = K.mass;
  if (is_fv) {
    double w0, w1;
    {  // DistanceFaceVertex (dev_math.cuh: dist_face_vertex), distance gate on the square
      const V3 v20 = p0m - p2m, v21 = p1m - p2m;
      const double t0 = dot(v20, v20), t1 = dot(v21, v21), t2 = dot(v20, v21);
      const double t3 = dot(v20, p3m - p2m), t4 = dot(v21, p3m - p2m);
      const double det = t0 * t1 - t2 * t2;
      const double invdet = rcp_cert(det, rcp_newton(det), ok);
      w0 = (+t1 * t3 - t2 * t4) * invdet;
      w1 = (-t2 * t3 + t0 * t4) * invdet;
      const double w2m = 1 - w0 - w1;
      const V3 pw = (w0 * p0m + w1 * p1m) + w2m * p2m;
      const V3 dd = pw - p3m;
      const double d2 = dd.x * dd.x + dd.y * dd.y + dd.z * dd.z;
      if (w0 < 0 || w0 > 1) return ok ? 0 : 2;
      if (w1 < 0 || w1 > 1) return ok ? 0 : 2;
      if (d2 > K.dist2) return ok ? 0 : 2;
    }
    const double w2 = 1.0 - w0 - w1;
    const V3 pc = (w0 * p0 + w1 * p1) + w2 * p2;
    const double den = 1.0 + w0 * w0 + w1 * w1 + w2 * w2;
    const double yden = rcp_newton(den);
    const V3 norm = normalized_cert(p3 - pc, ok);
    const double rel_v = dot(((v3v - w0 * v0) - w1 * v1) - w2 * v2, norm);
    if (rel_v > K.thr_ccd) return ok ? 0 : 2;
    const double imp = mass * (K.thr_ccd - rel_v);
    double imp_mod = div_cert(2 * imp, den, yden, ok);
    imp_mod = div_cert(imp_mod, mass, K.inv_mass, ok);
    imp_mod *= 0.1;
    if (!ok) return 2;
    v[0] = upd3v(v0, -norm, imp_mod, w0);
    v[1] = upd3v(v1, -norm, imp_mod, w1);
    v[2] = upd3v(v2, -norm, imp_mod, w2);
    v[3] = upd3sv(v3v, norm, imp_mod);
  } else {
    double w01, w23;
    {  // DistanceEdgeEdge (dev_math.cuh: dist_edge_edge); the (nearly) parallel configuration goes to the reference path
      const V3 vp = p1m - p0m, vq = p3m - p2m;
      const V3 cr = cross(vp, vq);
      if (cr.x * cr.x + cr.y * cr.y + cr.z * cr.z < K.par2) return 2;
      const double t0 = dot(vp, vp), t1 = dot(vq, vq), t2 = dot(vp, vq);
      const double t3 = dot(vp, p2m - p0m), t4 = dot(vq, p2m - p0m);
      const double det = t0 * t1 - t2 * t2;
      const double invdet = rcp_cert(det, rcp_newton(det), ok);
      w01 = (+t1 * t3 - t2 * t4) * invdet;
      w23 = (+t2 * t3 - t0 * t4) * invdet;
      const V3 pc = p0m + w01 * vp;
      const V3 qc = p2m + w23 * vq;
      const V3 dd = pc - qc;
      const double d2 = dd.x * dd.x + dd.y * dd.y + dd.z * dd.z;
      if (w01 < 0 || w01 > 1) return ok ? 0 : 2;
      if (w23 < 0 || w23 > 1) return ok ? 0 : 2;
      if (d2 > K.dist2) return ok ? 0 : 2;
    }
    const V3 c01 = (1 - w01) * p0 + w01 * p1;
    const V3 c23 = (1 - w23) * p2 + w23 * p3;
    const double den = w01 * w01 + (1 - w01) * (1 - w01) + w23 * w23 + (1 - w23) * (1 - w23);
    const double yden = rcp_newton(den);
    const V3 norm = normalized_cert(c23 - c01, ok);
    const double rel_v = dot((((1 - w23) * v2 + w23 * v3v) - (1 - w01) * v0) - w01 * v1, norm);
    if (rel_v > K.thr_ccd) return ok ? 0 : 2;
    const double imp = mass * (K.thr_ccd - rel_v);
    double imp_mod = div_cert(2 * imp, den, yden, ok);
    imp_mod = div_cert(imp_mod, mass, K.inv_mass, ok);
    imp_mod *= 0.1;
    if (!ok) return 2;
    v[0] = upd3v(v0, -norm, imp_mod, 1 - w01);
    v[1] = upd3v(v1, -norm, imp_mod, w01);
    v[2] = upd3v(v2, norm, imp_mod, 1 - w23);
    v[3] = upd3v(v3v, norm, imp_mod, w23);
  }
  return 1;
}

// the same on velocities in global memory
__device__ __forceinline__ int impulse_ccd_fast(const int* n, V3 p0, V3 p1, V3 p2, V3 p3, double* uvw, const ImpConst& K) {
  V3 v[4] = {ld3(uvw, n[0]), ld3(uvw, n[1]), ld3(uvw, n[2]), ld3(uvw, n[3])};
  const int how = impulse_ccd_fast_v(n[4], p0, p1, p2, p3, v, K);
  if (how == 1) {
#pragma unroll
    for (int a = 0; a < 4; a++) st3(uvw, n[a], v[a]);
  }
  return how;
}

// (vertex, element*4+corner) pairs in element order; all scenes in one launch (scene = blockIdx.y, stride cap4)
__global__ void k_imp_pairs(long long cap_elem, long long cap4, const int* __restrict__ elems_all, unsigned* __restrict__ keys_all,
                            int* __restrict__ vals_all, const int* __restrict__ cnt) {
  const int s = blockIdx.y;
  const int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int n4 = c[CNT_ELEM] * 4;
  const int* elems = elems_all + (size_t)s * cap_elem * 5;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    const int k = i >> 2, a = i & 3;
    keys_all[(size_t)s * cap4 + i] = (unsigned)elems[k * 5 + a];
    vals_all[(size_t)s * cap4 + i] = i;
  }
}
// segment offsets of the scenes' (vertex, element) slices for the segmented sort; inactive scenes get empty segments
__global__ void k_imp_segments(int B, long long cap4, const int* __restrict__ cnt, int* __restrict__ seg) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= B) return;
  const int* c = cnt + s * NCNT;
  const int n4 = c[CNT_ACTIVE] ? c[CNT_ELEM] * 4 : 0;
  seg[s] = (int)(s * cap4);
  seg[B + s] = (int)(s * cap4) + n4;
}
// predecessor of (element, corner): the previous element touching the same vertex and the corner the vertex has
// there, as element * 4 + corner, or -1
__global__ void k_imp_pred(long long cap_elem, long long cap4, const unsigned* __restrict__ keys_all,
                           const int* __restrict__ vals_all, int* __restrict__ pred_all, const int* __restrict__ cnt) {
  const int s = blockIdx.y;
  const int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int n4 = c[CNT_ELEM] * 4;
  const unsigned* keys = keys_all + (size_t)s * cap4;
  const int* vals = vals_all + (size_t)s * cap4;
  int* pred = pred_all + (size_t)s * cap_elem * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x)
    pred[vals[i]] = (i > 0 && keys[i - 1] == keys[i]) ? vals[i - 1] : -1;
}

// one CTA per scene: level the dependency DAG, bucket the elements by level, apply level after level
__global__ void __launch_bounds__(1024) k_imp_run(int ccd, int np, long long cap_elem, const int* __restrict__ elems_all,
                                                   const int* __restrict__ pred_all, int* level_all, int* order_all,
                                                   int* lptr_all, const double* __restrict__ xyz_all, double* uvw_all,
                                                   double delta, double stiffness, double dt, double mass,
                                                   const int* __restrict__ cnt) {
  const int s = blockIdx.x;
  const int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int ne = c[CNT_ELEM];
  if (ne == 0) return;
  const int* elems = elems_all + (size_t)s * cap_elem * 5;
  const int* pred = pred_all + (size_t)s * cap_elem * 4;
  int* level = level_all + (size_t)s * cap_elem;
  int* order = order_all + (size_t)s * cap_elem;
  int* lptr = lptr_all + (size_t)s * (cap_elem + 2) * 2;  // [0, cap+2): level offsets, [cap+2, 2cap+4): cursors
  int* cursor = lptr + cap_elem + 2;
  const double* xyz = xyz_all + (size_t)s * np * 3;
  double* uvw = uvw_all + (size_t)s * np * 3;
  __shared__ int sh_changed, sh_nlev;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int k = tid; k < ne; k += nt) level[k] = 0;
  if (tid == 0) sh_nlev = 0;
  __syncthreads();
  // chaotic relaxation towards the longest-path levels (monotone, unique fixed point)
  for (;;) {
    if (tid == 0) sh_changed = 0;
    __syncthreads();
    for (int k = tid; k < ne; k += nt) {
      int l = 0;
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const int pk = pred[k * 4 + a];
        if (pk >= 0) l = max(l, ((volatile int*)level)[pk >> 2] + 1);
      }
      if (l > level[k]) { level[k] = l; sh_changed = 1; }
    }
    __syncthreads();
    if (!sh_changed) break;
    __syncthreads();
  }
  for (int k = tid; k < ne; k += nt) atomicMax(&sh_nlev, level[k] + 1);
  __syncthreads();
  const int nlev = sh_nlev;
  for (int l = tid; l <= nlev; l += nt) lptr[l] = 0;
  __syncthreads();
  for (int k = tid; k < ne; k += nt) atomicAdd(&lptr[level[k] + 1], 1);
  __syncthreads();
  if (tid == 0) {
    for (int l = 0; l < nlev; l++) lptr[l + 1] += lptr[l];
  }
  __syncthreads();
  for (int l = tid; l < nlev; l += nt) cursor[l] = lptr[l];
  __syncthreads();
  for (int k = tid; k < ne; k += nt) order[atomicAdd(&cursor[level[k]], 1)] = k;
  __syncthreads();
  for (int l = 0; l < nlev; l++) {
    const int b = lptr[l], e = lptr[l + 1];
    for (int i = b + tid; i < e; i += nt) {
      const int* n = elems + (size_t)order[i] * 5;
      if (ccd) impulse_ccd_one(n, xyz, uvw, delta, dt, mass);
      else impulse_prx_one(n, xyz, uvw, delta, stiffness, dt, mass);
    }
    __syncthreads();
  }
}

// Data-flow variant of the same schedule (the default): no levels, no barriers.  One CTA per scene; thread t owns the
// elements t, t + T, ... in ascending order and starts an element as soon as the (at most four) predecessors on its
// vertices have finished (one "done" byte per element in shared memory).  The smallest unfinished element is always
// runnable and its owner is working on it, so the loop cannot dead-lock; independent dependency chains overlap
// instead of meeting at a barrier per level, and geometry (positions, indices) is fetched before the wait.
// Consecutive elements (which usually share their first vertex) are spread over different warps.
#ifdef IMP_TRACE
// development build (tools/build_variant.sh trace -DIMP_TRACE): per element {clock when its predecessors were done,
// clock when it was done, outcome}; launch_impulse prints a summary to stderr
__device__ long long* g_imp_trace;
#endif
// release / acquire around the "done" flags of k_imp_flow (IMP_FENCE_SC: the sequentially consistent __threadfence_block)
#ifdef IMP_FENCE_SC
#define IMP_FENCE() __threadfence_block()
#else
#define IMP_FENCE() asm volatile("fence.acq_rel.cta;" ::: "memory")
#endif
__global__ void __launch_bounds__(256) k_imp_flow(int ccd, int np, long long cap_elem, const int* __restrict__ elems_all,
                                                    const int* __restrict__ pred_all, const double* __restrict__ xyz_all,
                                                    double* uvw_all, const ImpConst K, int* cnt) {
  extern __shared__ unsigned char done_sm[];
  const double delta = K.delta, stiffness = K.stiffness, dt = K.dt, mass = K.mass;
  const int s = blockIdx.x;
  int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int ne = c[CNT_ELEM];
  if (ne == 0) return;
  const int* elems = elems_all + (size_t)s * cap_elem * 5;
  const int4* pred = reinterpret_cast<const int4*>(pred_all + (size_t)s * cap_elem * 4);
  const double* xyz = xyz_all + (size_t)s * np * 3;
  double* uvw = uvw_all + (size_t)s * np * 3;
  volatile unsigned char* done = done_sm;
  for (int k = threadIdx.x; k < ne; k += blockDim.x) done_sm[k] = 0;
  __syncthreads();
  // thread <-> element: element k of a round goes to warp (k % nwarp), lane (k / nwarp).  The sweep is nearly serial
  // (chain depth ~0.9 ne), so the thread count hardly matters: 256 measured best (512: +3 %, 32: +28 % dense window)
  // Only K.lanes lanes of every warp work: a waiting lane polls, and the polls of the lanes of one warp are issued one
  // divergent group after the other, in the same instruction stream as the lane that is working on the chain.
  const int nwarp = blockDim.x >> 5;
  if ((int)(threadIdx.x & 31) >= K.lanes) return;
  const int slot = (threadIdx.x & 31) * nwarp + (threadIdx.x >> 5);
  for (int k = slot; k < ne; k += nwarp * K.lanes) {
    const int4 p = pred[k];
    int n[5];
#pragma unroll
    for (int a = 0; a < 5; a++) n[a] = elems[(size_t)k * 5 + a];
    PrxGeom G;
    V3 q0 = v3(0, 0, 0), q1 = q0, q2 = q0, q3 = q0;
    if (!ccd) {  // velocity independent: in the shadow of the wait
      impulse_prx_geom(n, xyz, delta, G);
      if (K.fast) impulse_prx_geom_more(G, stiffness, dt);
    } else {
      q0 = ld3(xyz, n[0]); q1 = ld3(xyz, n[1]); q2 = ld3(xyz, n[2]); q3 = ld3(xyz, n[3]);
    }
    // waiting lanes sleep between polls: a tight poll loop keeps the whole warp in "branch resolving" stalls (ncu: the
    // top stall of this kernel) and slows the one lane of the warp that is working on the chain.  Measured: 0 ns
    // 30.4 ms, 20..200 ns 29.6-29.8 ms, 500 ns 30.8 ms, 1000 ns 33.3 ms per step in the dense window
    const unsigned sleep_ns = (unsigned)K.sleep_ns;
#define IMP_WAIT(P_) if ((P_) >= 0) while (!done[(P_) >> 2]) { if (sleep_ns > 0) __nanosleep(sleep_ns); }
    IMP_WAIT(p.x) IMP_WAIT(p.y) IMP_WAIT(p.z) IMP_WAIT(p.w)
#undef IMP_WAIT
    IMP_FENCE();
#ifdef IMP_TRACE
    const long long tr0 = clock64();
    int tr_how = 3;
#endif
    // certified fast arithmetic first (imp_fast.cuh); what cannot be certified is recomputed with the reference's
    // plain IEEE operations (CNT_IMP_SLOW counts those elements)
    if (ccd) {
      const int how = K.fast ? impulse_ccd_fast(n, q0, q1, q2, q3, uvw, K) : 2;
#ifdef IMP_TRACE
      tr_how = how;
#endif
      if (how == 2) {
        if (K.fast) atomicAdd(&c[CNT_IMP_SLOW], 1);
        impulse_ccd_apply(n, q0, q1, q2, q3, uvw, delta, dt, mass);
      }
    } else if (K.fast) {
      if (!impulse_prx_apply_fast(n, G, uvw, K)) atomicAdd(&c[CNT_IMP_SLOW], 1);
    } else {
      impulse_prx_apply(n, G, uvw, stiffness, dt, mass);
    }
    IMP_FENCE();
    done[k] = 1;
#ifdef IMP_TRACE
    if (s == 0) { g_imp_trace[k * 3 + 0] = tr0; g_imp_trace[k * 3 + 1] = clock64(); g_imp_trace[k * 3 + 2] = tr_how; }
#endif
  }
}

// smallest x with sqrt(x) >= c / largest x with sqrt(x) <= c (the host sqrt is the correctly rounded IEEE operation)
static double sqrt_ge_threshold(double c) {
  double x = c * c;
  while (x > 0 && std::sqrt(x) >= c) x = std::nextafter(x, -INFINITY);
  while (std::sqrt(x) < c) x = std::nextafter(x, INFINITY);
  return x;
}
static double sqrt_le_threshold(double c) {
  double x = c * c;
  while (std::sqrt(x) <= c) x = std::nextafter(x, INFINITY);
  while (x > 0 && std::sqrt(x) > c) x = std::nextafter(x, -INFINITY);
  return x;
}
static ImpConst make_imp_const(const Ctx* c, double delta, double stiffness, double dt, double mass) {
  ImpConst K;
  K.delta = delta; K.stiffness = stiffness; K.dt = dt; K.mass = mass;
  K.inv_mass = 1.0 / mass;
  K.thr_ccd = 0.1 * delta / dt;
  K.par2 = sqrt_ge_threshold(1.0e-10);
  K.dist2 = delta > 0 ? sqrt_le_threshold(delta) : 0.0;
  // the certified path needs finite positive parameters; anything else keeps the reference arithmetic
  K.lanes = c->imp_lanes;
  K.sleep_ns = c->imp_sleep_ns;
  K.fast = c->imp_fast && delta > 0 && mass > 0 && dt > 0 && std::isfinite(K.inv_mass) && std::isfinite(K.thr_ccd);
  return K;
}

int launch_impulse(Ctx* c, const double* xyz, double* uvw, double delta, double stiffness, double dt, double mass, int ccd) {
  // element counts are on the host already (the reference's loop control needs them).  Pairs and predecessors of all
  // scenes in one launch each; the library radix sort in between runs scene by scene on the scene's slice.
  {
    int ne_all = 0;
    for (int s = 0; s < c->B; s++)
      if (c->h_cnt[s * NCNT + CNT_ACTIVE]) ne_all = std::max(ne_all, c->h_cnt[s * NCNT + CNT_ELEM]);
    if (ne_all > 0) {
      const long long cap4 = (long long)ne_all * 4;
      CG_CHECK(ensure_sort_capacity(c, cap4 * c->B));
      unsigned* keys = (unsigned*)c->sort_keys;
      unsigned* keys2 = (unsigned*)c->sort_keys2;
      int end_bit = 1;
      while ((1LL << end_bit) < c->np) end_bit++;
      size_t tb = 0;
      CG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, c->sort_vals, c->sort_vals2, (int)cap4, 0, end_bit, c->stream));
      CG_CHECK(ensure_scratch(c, tb));
      dim3 g(std::min(div_up((int)cap4, 256), 148 * 4), c->B);
      k_imp_pairs<<<g, 256, 0, KS(c, "k_imp_pairs")>>>(c->cap_elem, cap4, c->elems, keys, c->sort_vals, c->cnt);
      CG_LAUNCH_CHECK(c);
      if (c->B >= 4 && cap4 * c->B < (1LL << 31)) {
        // batches: one segmented sort (segment = scene) instead of a device-wide sort per scene
        int* seg = c->imp_level;  // 2 B ints of scratch: begin / end offsets of the scenes' slices
        k_imp_segments<<<div_up(c->B, 128), 128, 0, KS(c, "k_imp_segments")>>>(c->B, cap4, c->cnt, seg);
        CG_LAUNCH_CHECK(c);
        size_t tbs = 0;
        CG_CUDA(cub::DeviceSegmentedRadixSort::SortPairs(nullptr, tbs, keys, keys2, c->sort_vals, c->sort_vals2, (int)(cap4 * c->B), c->B,
                                                         seg, seg + c->B, 0, end_bit, c->stream));
        CG_CHECK(ensure_scratch(c, tbs));
        CG_CUDA(cub::DeviceSegmentedRadixSort::SortPairs(c->scratch, tbs, keys, keys2, c->sort_vals, c->sort_vals2, (int)(cap4 * c->B),
                                                         c->B, seg, seg + c->B, 0, end_bit, c->stream));
        c->launches += 3;
      } else {
        for (int s = 0; s < c->B; s++) {
          if (!c->h_cnt[s * NCNT + CNT_ACTIVE]) continue;
          const int n4 = c->h_cnt[s * NCNT + CNT_ELEM] * 4;
          if (n4 == 0) continue;
          size_t tbs = tb;
          CG_CUDA(cub::DeviceRadixSort::SortPairs(c->scratch, tbs, keys + s * cap4, keys2 + s * cap4, c->sort_vals + s * cap4,
                                                  c->sort_vals2 + s * cap4, n4, 0, end_bit, c->stream));
          c->launches += 3;
        }
      }
      k_imp_pred<<<g, 256, 0, KS(c, "k_imp_pred")>>>(c->cap_elem, cap4, keys2, c->sort_vals2, c->imp_pred, c->cnt);
      CG_LAUNCH_CHECK(c);
    }
  }
  int ne_max = 0;
  for (int s = 0; s < c->B; s++)
    if (c->h_cnt[s * NCNT + CNT_ACTIVE]) ne_max = std::max(ne_max, c->h_cnt[s * NCNT + CNT_ELEM]);
  if (ne_max == 0) return CLOTHGPU_OK;
  static const bool use_levels = getenv("CLOTHGPU_IMPULSE") && strcmp(getenv("CLOTHGPU_IMPULSE"), "levels") == 0;
  if (!use_levels && ne_max <= 200 * 1024) {
    const int smem = (ne_max + 127) / 128 * 128;
    CG_CUDA(cudaFuncSetAttribute(k_imp_flow, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(smem, 48 * 1024)));
    const int imp_threads = c->imp_threads;
#ifdef IMP_TRACE
    static long long* tr_dev = nullptr;
    const int tr_cap = 1 << 20;
    if (!tr_dev) {
      CG_CUDA(cudaMalloc((void**)&tr_dev, sizeof(long long) * 3 * tr_cap));
      CG_CUDA(cudaMemcpyToSymbol(g_imp_trace, &tr_dev, sizeof(tr_dev)));
    }
#endif
    k_imp_flow<<<c->B, imp_threads, smem, KS(c, "k_imp_flow")>>>(ccd, c->np, c->cap_elem, c->elems, c->imp_pred, xyz, uvw,
                                                  make_imp_const(c, delta, stiffness, dt, mass), c->cnt);
#ifdef IMP_TRACE
    {
      const int ne = std::min(c->h_cnt[CNT_ELEM], tr_cap);
      std::vector<long long> tr((size_t)ne * 3);
      std::vector<int> pred((size_t)ne * 4);
      CG_CUDA(cudaStreamSynchronize(c->stream));
      CG_CUDA(cudaMemcpy(tr.data(), tr_dev, sizeof(long long) * 3 * ne, cudaMemcpyDeviceToHost));
      CG_CUDA(cudaMemcpy(pred.data(), c->imp_pred, sizeof(int) * 4 * ne, cudaMemcpyDeviceToHost));
      double comp = 0, hand = 0;
      long long nh = 0, how[4] = {0, 0, 0, 0}, tmin = LLONG_MAX, tmax = 0;
      for (int k = 0; k < ne; k++) {
        comp += (double)(tr[k * 3 + 1] - tr[k * 3 + 0]);
        how[tr[k * 3 + 2] & 3]++;
        long long last = -1;
        for (int a = 0; a < 4; a++) if (pred[k * 4 + a] >= 0) last = std::max(last, tr[(size_t)(pred[k * 4 + a] >> 2) * 3 + 1]);
        if (last >= 0) { hand += (double)(tr[k * 3 + 0] - last); nh++; }
        tmin = std::min(tmin, tr[k * 3 + 0]); tmax = std::max(tmax, tr[k * 3 + 1]);
      }
      fprintf(stderr, "imptrace ccd=%d ne=%d span=%lld cyc  compute/elem=%.0f  handoff/elem=%.0f (n=%lld)  nowrite=%lld write=%lld slow=%lld prx=%lld\n",
              ccd, ne, ne ? tmax - tmin : 0, ne ? comp / ne : 0.0, nh ? hand / nh : 0.0, nh, how[0], how[1], how[2], how[3]);
    }
#endif
  } else {
    k_imp_run<<<c->B, 1024, 0, KS(c, "k_imp_run")>>>(ccd, c->np, c->cap_elem, c->elems, c->imp_pred, c->imp_level, c->imp_order,
                                            c->imp_lptr, xyz, uvw, delta, stiffness, dt, mass, c->cnt);
  }
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}

// ------------------------------------------------------------------------------------------------ impact zones
__device__ __forceinline__ int uf_find(int* parent, int v) {
  for (;;) {
    const int p = ((volatile int*)parent)[v];
    if (p == v) return v;
    v = p;
  }
}
// link the larger root under the smaller one: the root of a component is its smallest vertex id
__device__ void uf_union(int* parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    const int hi = a > b ? a : b, lo = a > b ? b : a;
    if (atomicCAS(&parent[hi], hi, lo) == hi) return;
  }
}

// new members: the 4 vertices of every contact element (MakeRigidImpactZone inserts n[0..3], :244-271)
__global__ void k_riz_mark(int np, long long cap_elem, const int* __restrict__ elems_all, int* zone_all,
                           const int* __restrict__ cnt) {
  const int s = blockIdx.y;
  const int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int ne = c[CNT_ELEM];
  const int* elems = elems_all + (size_t)s * cap_elem * 5;
  int* zone = zone_all + (size_t)s * np;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ne; k += gridDim.x * blockDim.x)
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int v = elems[k * 5 + a];
      atomicCAS(&zone[v], -1, v);
    }
}
// merges: the vertices of one element, and every element vertex with its mesh neighbours that are members (:229-241)
__global__ void k_riz_union(int np, long long cap_elem, const int* __restrict__ elems_all,
                            const int* __restrict__ edge_index, const int* __restrict__ edge_array, int* zone_all,
                            const int* __restrict__ cnt) {
  const int s = blockIdx.y;
  const int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int ne = c[CNT_ELEM];
  const int* elems = elems_all + (size_t)s * cap_elem * 5;
  int* zone = zone_all + (size_t)s * np;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ne; k += gridDim.x * blockDim.x) {
    const int n0 = elems[k * 5 + 0];
#pragma unroll
    for (int a = 1; a < 4; a++) uf_union(zone, n0, elems[k * 5 + a]);
    for (int a = 0; a < 4; a++) {
      const int v = elems[k * 5 + a];
      for (int j = edge_index[v]; j < edge_index[v + 1]; j++) {
        const int u = edge_array[j];
        if (((volatile int*)zone)[u] != -1) uf_union(zone, v, u);
      }
    }
  }
}
// label = root (smallest vertex id of the zone) or -1; sort key = label, non-members last
__global__ void k_riz_label(int np, int* zone_all, int* __restrict__ label_all, int* cnt) {
  const int s = blockIdx.y;
  int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  int* zone = zone_all + (size_t)s * np;
  int* label = label_all + (size_t)s * np;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= np) return;
  const int p = zone[v];
  label[v] = (p == -1) ? -1 : uf_find(zone, v);
}
__global__ void k_riz_keys(int np, const int* __restrict__ label, unsigned* __restrict__ keys, int* __restrict__ vals) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= np) return;
  keys[v] = label[v] < 0 ? 0x7fffffffu : (unsigned)label[v];
  vals[v] = v;
}

// ApplyRigidImpactZone, self_collision_cloth.cpp:298-348, one thread per zone (segment of the label-sorted vertex
// list, ascending vertex id inside the segment = std::set<int> order).
// Reference quirk (:327-329): the "inertia tensor" diagonal uses v.v (velocity) where q.q (position) is meant.
// one rigid impact zone: the members vals[i .. e) in ascending vertex id (ApplyRigidImpactZone, self_collision_cloth.cpp:298-348)
// STAGED: xyz / uvw0 are per-member copies in shared memory, indexed by the member's position in vals (a serial loop
// over global memory pays one L2 round trip per member and pass)
template <bool STAGED>
__device__ __forceinline__ void riz_zone(int i, int e, const int* __restrict__ vals, const double* xyz,
                                         const double* uvw0, double* __restrict__ uvw) {
#define RZ_LD(A_, J_) ld3((A_), STAGED ? (J_) : vals[(J_)])
  V3 gc = v3(0, 0, 0), av = v3(0, 0, 0);
  for (int j = i; j < e; j++) {
    gc = gc + RZ_LD(xyz, j);
    av = av + RZ_LD(uvw0, j);
  }
  gc = div_guarded(gc, (double)(e - i));
  av = div_guarded(av, (double)(e - i));
  V3 L = v3(0, 0, 0);
  double I[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = i; j < e; j++) {
    const V3 p = RZ_LD(xyz, j);
    const V3 v = RZ_LD(uvw0, j);
    L = L + cross(p - gc, v - av);
    const V3 q = p - gc;
    const double vv = dot(v, v);
    I[0] += vv - q.x * q.x; I[1] += -q.x * q.y;     I[2] += -q.x * q.z;
    I[3] += -q.y * q.x;     I[4] += vv - q.y * q.y; I[5] += -q.y * q.z;
    I[6] += -q.z * q.x;     I[7] += -q.z * q.y;     I[8] += vv - q.z * q.z;
  }
  double Iinv[9];
  {  // CalcInvMat3, self_collision_cloth.cpp:278-296
    const double det = +I[0] * I[4] * I[8] + I[3] * I[7] * I[2] + I[6] * I[1] * I[5] - I[0] * I[7] * I[5] -
                       I[6] * I[4] * I[2] - I[3] * I[1] * I[8];
    const double inv_det = 1.0 / det;
    Iinv[0] = inv_det * (I[4] * I[8] - I[5] * I[7]);
    Iinv[1] = inv_det * (I[2] * I[7] - I[1] * I[8]);
    Iinv[2] = inv_det * (I[1] * I[5] - I[2] * I[4]);
    Iinv[3] = inv_det * (I[5] * I[6] - I[3] * I[8]);
    Iinv[4] = inv_det * (I[0] * I[8] - I[2] * I[6]);
    Iinv[5] = inv_det * (I[2] * I[3] - I[0] * I[5]);
    Iinv[6] = inv_det * (I[3] * I[7] - I[4] * I[6]);
    Iinv[7] = inv_det * (I[1] * I[6] - I[0] * I[7]);
    Iinv[8] = inv_det * (I[0] * I[4] - I[1] * I[3]);
  }
  V3 omg;
  omg.x = Iinv[0] * L.x + Iinv[1] * L.y + Iinv[2] * L.z;
  omg.y = Iinv[3] * L.x + Iinv[4] * L.y + Iinv[5] * L.z;
  omg.z = Iinv[6] * L.x + Iinv[7] * L.y + Iinv[8] * L.z;
  for (int j = i; j < e; j++) {
    const int v = vals[j];
    const V3 rot = -cross(RZ_LD(xyz, j) - gc, omg);
    uvw[v * 3 + 0] = av.x + rot.x;
    uvw[v * 3 + 1] = av.y + rot.y;
    uvw[v * 3 + 2] = av.z + rot.z;
  }
#undef RZ_LD
}

__global__ void k_riz_apply(int np, const unsigned* __restrict__ keys, const int* __restrict__ vals,
                            const double* __restrict__ xyz, const double* __restrict__ uvw0, double* __restrict__ uvw,
                            int* cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const unsigned key = keys[i];
  if (key == 0x7fffffffu) return;
  atomicAdd(&cnt[CNT_RIZ_NODES], 1);
  if (i > 0 && keys[i - 1] == key) return;
  int e = i;
  while (e < np && keys[e] == key) e++;
  riz_zone<false>(i, e, vals, xyz, uvw0, uvw);
}
__global__ void k_zero_cnt(int* cnt, int slot) { cnt[blockIdx.x * NCNT + slot] = 0; }

// ApplyRigidImpactZone for all scenes in ONE launch (one CTA per scene) instead of, per scene, a key kernel, a library
// radix sort of all np vertices and the zone kernel: the members (a few hundred vertices, typically) are collected,
// sorted by (zone label, vertex id) with a bitonic network -- in shared memory up to RIZ_SMEM members, in a global
// scratch array beyond -- and every zone is then reduced by one thread exactly as in k_riz_apply (riz_zone).
enum { RIZ_T = 1024, RIZ_SMEM = 4096, RIZ_STAGE = 1024 };
__global__ void __launch_bounds__(RIZ_T) k_riz_apply_all(int np, const int* __restrict__ label_all,
                                                          const double* __restrict__ xyz_all, const double* __restrict__ uvw0_all,
                                                          double* __restrict__ uvw_all, unsigned long long* __restrict__ gkeys_all,
                                                          int* __restrict__ gvals_all, int* cnt) {
  extern __shared__ __align__(16) double riz_stage[];  // [RIZ_STAGE][3] positions, then [RIZ_STAGE][3] velocities
  __shared__ unsigned long long skeys[RIZ_SMEM];
  __shared__ int svals_n, s_total;
  const int s = blockIdx.x;
  int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int* label = label_all + (size_t)s * np;
  const double* xyz = xyz_all + (size_t)s * np * 3;
  const double* uvw0 = uvw0_all + (size_t)s * np * 3;
  double* uvw = uvw_all + (size_t)s * np * 3;
  if (threadIdx.x == 0) { svals_n = 0; s_total = 0; }
  __syncthreads();
  int mine = 0;
  for (int v = threadIdx.x; v < np; v += blockDim.x) mine += label[v] >= 0;
  if (mine) atomicAdd(&s_total, mine);
  __syncthreads();
  const int n = s_total;
  if (threadIdx.x == 0) c[CNT_RIZ_NODES] = n;
  if (n == 0) return;
  int N = 1;
  while (N < n) N <<= 1;
  unsigned long long* keys = (N <= RIZ_SMEM) ? skeys : gkeys_all + (size_t)s * 2 * np;  // (N < 2 np)
  for (int v = threadIdx.x; v < np; v += blockDim.x) {
    const int l = label[v];
    if (l >= 0) keys[atomicAdd(&svals_n, 1)] = ((unsigned long long)(unsigned)l << 32) | (unsigned)v;
  }
  for (int i = n + threadIdx.x; i < N; i += blockDim.x) keys[i] = ~0ULL;
  __syncthreads();
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], b = keys[l];
          const bool up = (i & k) == 0;
          if (up ? (b < a) : (a < b)) { keys[i] = b; keys[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  // vertex ids in sorted order (the member list riz_zone walks); the keys stay for the zone boundaries
  int* vals = gvals_all + (size_t)s * np;
  for (int i = threadIdx.x; i < n; i += blockDim.x) vals[i] = (int)(keys[i] & 0xffffffffULL);
  __syncthreads();
  const bool staged = n <= RIZ_STAGE;
  if (staged) {  // the members' positions and start velocities, in member order, for the serial zone loops
    for (int t = threadIdx.x; t < n * 3; t += blockDim.x) {
      const int i = t / 3, k = t - i * 3;
      riz_stage[t] = xyz[(size_t)vals[i] * 3 + k];
      riz_stage[RIZ_STAGE * 3 + t] = uvw0[(size_t)vals[i] * 3 + k];
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned zl = (unsigned)(keys[i] >> 32);
    if (i > 0 && (unsigned)(keys[i - 1] >> 32) == zl) continue;
    int e = i;
    while (e < n && (unsigned)(keys[e] >> 32) == zl) e++;
    if (staged) riz_zone<true>(i, e, vals, riz_stage, riz_stage + RIZ_STAGE * 3, uvw);
    else riz_zone<false>(i, e, vals, xyz, uvw0, uvw);
  }
}

int launch_riz_reset(Ctx* c) {
  CG_CUDA(cudaMemsetAsync(c->zone, 0xff, sizeof(int) * (size_t)c->B * c->np, c->stream));
  CG_CUDA(cudaMemsetAsync(c->zone_label, 0xff, sizeof(int) * (size_t)c->B * c->np, c->stream));
  return CLOTHGPU_OK;
}

int launch_riz_make(Ctx* c) {
  dim3 g(148, c->B);
  k_riz_mark<<<g, 256, 0, KS(c, "k_riz_mark")>>>(c->np, c->cap_elem, c->elems, c->zone, c->cnt);
  CG_LAUNCH_CHECK(c);
  k_riz_union<<<g, 256, 0, KS(c, "k_riz_union")>>>(c->np, c->cap_elem, c->elems, c->edge_index, c->edge_array, c->zone, c->cnt);
  CG_LAUNCH_CHECK(c);
  dim3 g2(div_up(c->np, 256), c->B);
  k_riz_label<<<g2, 256, 0, KS(c, "k_riz_label")>>>(c->np, c->zone, c->zone_label, c->cnt);
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}

int launch_riz_apply(Ctx* c, const double* xyz, double* uvw) {
  if (c->riz_fused) {
    // scratch per scene: 2 np keys (a power of two >= the member count) + np vertex ids
    CG_CHECK(ensure_sort_capacity(c, 2LL * c->np * c->B));
    const int smem = RIZ_STAGE * 6 * (int)sizeof(double);
    CG_CUDA(cudaFuncSetAttribute(k_riz_apply_all, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_riz_apply_all<<<c->B, RIZ_T, smem, KS(c, "k_riz_apply")>>>(c->np, c->zone_label, xyz, c->uvw0, uvw, c->sort_keys, c->sort_vals, c->cnt);
    CG_LAUNCH_CHECK(c);
    return CLOTHGPU_OK;
  }
  k_zero_cnt<<<c->B, 1, 0, KS(c, "k_zero_cnt")>>>(c->cnt, CNT_RIZ_NODES);
  CG_LAUNCH_CHECK(c);
  CG_CHECK(ensure_sort_capacity(c, c->np));
  unsigned* keys = (unsigned*)c->sort_keys;
  unsigned* keys2 = (unsigned*)c->sort_keys2;
  size_t tb = 0;
  CG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, c->sort_vals, c->sort_vals2, c->np, 0, 32, c->stream));
  CG_CHECK(ensure_scratch(c, tb));
  for (int s = 0; s < c->B; s++) {
    if (!c->h_cnt[s * NCNT + CNT_ACTIVE]) continue;
    k_riz_keys<<<div_up(c->np, 256), 256, 0, KS(c, "k_riz_keys")>>>(c->np, c->zone_label + (size_t)s * c->np, keys, c->sort_vals);
    CG_LAUNCH_CHECK(c);
    CG_CUDA(cub::DeviceRadixSort::SortPairs(c->scratch, tb, keys, keys2, c->sort_vals, c->sort_vals2, c->np, 0, 32, c->stream));
    c->launches += 3;
    k_riz_apply<<<div_up(c->np, 128), 128, 0, KS(c, "k_riz_apply")>>>(c->np, keys2, c->sort_vals2, xyz + (size_t)s * c->ndof,
                                                           c->uvw0 + (size_t)s * c->ndof, uvw + (size_t)s * c->ndof,
                                                           c->cnt + s * NCNT);
    CG_LAUNCH_CHECK(c);
  }
  return CLOTHGPU_OK;
}

// ------------------------------------------------------------------------------------------------ self test
// The certified helpers of imp_fast.cuh against the plain IEEE operations on pseudo-random operands (random mantissas,
// exponents in [-60, 60], half of the quotients placed next to a rounding boundary): a certified result must equal the
// hardware result bit for bit.  out[2k] = mismatches, out[2k+1] = uncertified samples of test k: 0 quotient with a
// Newton reciprocal, 1 quotient with the correctly rounded reciprocal, 2 reciprocal, 3 square root, 4 reciprocal of
// the root from the coupled iteration.
__device__ __forceinline__ unsigned long long splitmix(unsigned long long& x) {
  x += 0x9e3779b97f4a7c15ULL;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double rnd_double(unsigned long long& st) {
  const unsigned long long m = splitmix(st);
  const unsigned long long e = 1023 - 60 + (splitmix(st) % 121);
  return __longlong_as_double((long long)((m & 0x800fffffffffffffULL) | (e << 52)));
}
__global__ void k_cert_selftest(long long per_thread, unsigned long long seed, unsigned long long* out) {
  unsigned long long st = seed + 0x1234567ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x);
  unsigned long long bad[5] = {0, 0, 0, 0, 0}, unc[5] = {0, 0, 0, 0, 0};
  for (long long i = 0; i < per_thread; i++) {
    double a = rnd_double(st);
    const double b = rnd_double(st);
    if (i & 1) {  // a / b next to the midpoint of two neighbouring doubles
      const double q = rnd_double(st);
      a = __fma_rn(b, q, 0.5 * (b * ulp_of(q)));
    }
    { bool ok = true; const double q = div_cert(a, b, rcp_newton(b), ok); if (!ok) unc[0]++; else if (q != a / b) bad[0]++; }
    { bool ok = true; const double q = div_cert(a, b, 1.0 / b, ok); if (!ok) unc[1]++; else if (q != a / b) bad[1]++; }
    { bool ok = true; const double y = rcp_cert(b, rcp_newton(b), ok); if (!ok) unc[2]++; else if (y != 1.0 / b) bad[2]++; }
    {
      bool ok = true;
      double h;
      const double s = fabs(a);
      const double g = sqrt_cert(s, h, ok);
      if (!ok) unc[3]++; else if (g != sqrt(s)) bad[3]++;
      bool ok2 = ok;
      const double y = rcp_cert(g, h + h, ok2);
      if (!ok2) unc[4]++; else if (y != 1.0 / sqrt(s)) bad[4]++;
    }
  }
  for (int k = 0; k < 5; k++) {
    atomicAdd(&out[2 * k], bad[k]);
    atomicAdd(&out[2 * k + 1], unc[k]);
  }
}

}  // namespace cg

extern "C" int clothgpu_selftest_sqrt_thresholds(double c, double* ge, double* le) {
  if (!(c > 0) || !ge || !le) { cg::set_error("clothgpu_selftest_sqrt_thresholds: bad argument"); return CLOTHGPU_ERR_ARG; }
  *ge = cg::sqrt_ge_threshold(c);
  *le = cg::sqrt_le_threshold(c);
  return CLOTHGPU_OK;
}

extern "C" int clothgpu_selftest_certified_math(int device, long long samples, unsigned long long seed, long long* out10) {
  using namespace cg;
  if (!out10 || samples <= 0) { set_error("clothgpu_selftest_certified_math: bad argument"); return CLOTHGPU_ERR_ARG; }
  CG_CUDA(cudaSetDevice(device));
  unsigned long long* d = nullptr;
  CG_CUDA(cudaMalloc((void**)&d, sizeof(unsigned long long) * 10));
  CG_CUDA(cudaMemset(d, 0, sizeof(unsigned long long) * 10));
  const int blocks = 148 * 4, threads = 128;
  const long long per_thread = (samples + (long long)blocks * threads - 1) / ((long long)blocks * threads);
  k_cert_selftest<<<blocks, threads>>>(per_thread, seed, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out10, d, sizeof(unsigned long long) * 10, cudaMemcpyDeviceToHost);
  cudaFree(d);
  CG_CUDA(e);
  return CLOTHGPU_OK;
}
