// csrc/dev_math.cuh -- fp64 3-vector helpers and geometric predicates on the device.
//
// Parity contract: every expression keeps the operand order and association of the reference (file:line cited),
// and the library is compiled with -fmad=false so that no multiply-add is contracted (the reference trajectories
// fork under FMA contraction).  IEEE fp64 add/mul/div/sqrt are correctly rounded on sm_100a.
#pragma once
#include <cuda_runtime.h>

namespace cg {

struct V3 {
  double x, y, z;
};

__device__ __forceinline__ V3 v3(double x, double y, double z) {
  V3 r;
  r.x = x; r.y = y; r.z = z;
  return r;
}
__device__ __forceinline__ V3 ld3(const double* __restrict__ a, int i) { return v3(a[i * 3 + 0], a[i * 3 + 1], a[i * 3 + 2]); }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
// CVector3D operator*(double, vec) and (vec, double): component * scalar (vector3d.h:111-135)
__device__ __forceinline__ V3 operator*(double s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
// vector3d.h:214-217
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// vector3d.h:219-226
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
// vector3d.h:94
__device__ __forceinline__ double length(V3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
// vector3d.h:97-103
__device__ __forceinline__ V3 normalized(V3 a) {
  const double invmag = 1.0 / length(a);
  return v3(a.x * invmag, a.y * invmag, a.z * invmag);
}
// ScalarTripleProduct, vector3d.h:228-230
__device__ __forceinline__ double triple(V3 a, V3 b, V3 c) {
  return a.x * (b.y * c.z - b.z * c.y) + a.y * (b.z * c.x - b.x * c.z) + a.z * (b.x * c.y - b.y * c.x);
}
// CVector3D::operator/=, vector3d.h:63-67: silently skipped for |d| < 1e-16
__device__ __forceinline__ V3 div_guarded(V3 a, double d) {
  if (fabs(d) < 1.0e-16) return a;
  return v3(a.x / d, a.y / d, a.z / d);
}

// ---- axis-aligned boxes: {xmin,xmax,ymin,ymax,zmin,zmax}
struct Box {
  double lo[3], hi[3];
};
__device__ __forceinline__ Box ld_box(const double* __restrict__ boxes, int i) {
  Box b;
  const double* p = boxes + (size_t)i * 6;
  b.lo[0] = p[0]; b.hi[0] = p[1];
  b.lo[1] = p[2]; b.hi[1] = p[3];
  b.lo[2] = p[4]; b.hi[2] = p[5];
  return b;
}
// CAABB3D::IsIntersect, aabb.h:73-84 (closed intervals)
__device__ __forceinline__ bool box_intersect(const Box& a, const Box& b) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (a.hi[k] < b.lo[k]) return false;
    if (a.lo[k] > b.hi[k]) return false;
  }
  return true;
}
// CAABB3D::IsInside, aabb.h:137-144
__device__ __forceinline__ bool box_inside(const Box& b, V3 p) {
  return p.x >= b.lo[0] && p.x <= b.hi[0] && p.y >= b.lo[1] && p.y <= b.hi[1] && p.z >= b.lo[2] && p.z <= b.hi[2];
}
// CAABB3D::AddPoint, aabb.h:102-122, first point
__device__ __forceinline__ void box_init(Box& b, V3 p, double eps) {
  b.lo[0] = p.x - eps; b.hi[0] = p.x + eps;
  b.lo[1] = p.y - eps; b.hi[1] = p.y + eps;
  b.lo[2] = p.z - eps; b.hi[2] = p.z + eps;
}
__device__ __forceinline__ void box_add(Box& b, V3 p, double eps) {
  const double c[3] = {p.x, p.y, p.z};
#pragma unroll
  for (int a = 0; a < 3; a++) {
    b.lo[a] = (b.lo[a] < c[a] - eps) ? b.lo[a] : c[a] - eps;
    b.hi[a] = (b.hi[a] > c[a] + eps) ? b.hi[a] : c[a] + eps;
  }
}

// DistanceFaceVertex, bvh_aabb.cpp:321-340
__device__ __forceinline__ double dist_face_vertex(V3 p0, V3 p1, V3 p2, V3 p3, double& w0, double& w1) {
  const V3 v20 = p0 - p2, v21 = p1 - p2;
  const double t0 = dot(v20, v20), t1 = dot(v21, v21), t2 = dot(v20, v21);
  const double t3 = dot(v20, p3 - p2), t4 = dot(v21, p3 - p2);
  const double det = t0 * t1 - t2 * t2;
  const double invdet = 1.0 / det;
  w0 = (+t1 * t3 - t2 * t4) * invdet;
  w1 = (-t2 * t3 + t0 * t4) * invdet;
  const double w2 = 1 - w0 - w1;
  const V3 pw = (w0 * p0 + w1 * p1) + w2 * p2;
  return length(pw - p3);
}

// DistanceEdgeEdge, bvh_aabb.cpp:343-384 (parallel branch :350-371)
static __device__ __noinline__ double dist_edge_edge(V3 p0, V3 p1, V3 q0, V3 q1, double& rp, double& rq) {
  const V3 vp = p1 - p0, vq = q1 - q0;
  if (length(cross(vp, vq)) < 1.0e-10) {
    const V3 pq0 = p0 - q0;
    const V3 nvp = normalized(vp);
    const V3 vert = pq0 - dot(pq0, nvp) * nvp;
    const double dist = length(vert);
    const double lp0 = dot(p0, nvp), lp1 = dot(p1, nvp), lq0 = dot(q0, nvp), lq1 = dot(q1, nvp);
    const double p_min = (lp0 < lp1) ? lp0 : lp1, p_max = (lp0 > lp1) ? lp0 : lp1;
    const double q_min = (lq0 < lq1) ? lq0 : lq1, q_max = (lq0 > lq1) ? lq0 : lq1;
    double lm;
    if (p_max < q_min) lm = (p_max + q_min) * 0.5;
    else if (q_max < p_min) lm = (q_max + p_min) * 0.5;
    else if (p_max < q_max) lm = (p_max + q_min) * 0.5;
    else lm = (q_max + p_min) * 0.5;
    rp = (lm - lp0) / (lp1 - lp0);
    rq = (lm - lq0) / (lq1 - lq0);
    return dist;
  }
  const double t0 = dot(vp, vp), t1 = dot(vq, vq), t2 = dot(vp, vq);
  const double t3 = dot(vp, q0 - p0), t4 = dot(vq, q0 - p0);
  const double det = t0 * t1 - t2 * t2;
  const double invdet = 1.0 / det;
  rp = (+t1 * t3 - t2 * t4) * invdet;
  rq = (+t2 * t3 - t0 * t4) * invdet;
  const V3 pc = p0 + rp * vp;
  const V3 qc = q0 + rq * vq;
  return length(pc - qc);
}

// Height, vector3d.h:254-267
__device__ __forceinline__ double tet_height(V3 v1, V3 v2, V3 v3p, V3 v4) {
  double nx = (v2.y - v1.y) * (v3p.z - v1.z) - (v2.z - v1.z) * (v3p.y - v1.y);
  double ny = (v2.z - v1.z) * (v3p.x - v1.x) - (v2.x - v1.x) * (v3p.z - v1.z);
  double nz = (v2.x - v1.x) * (v3p.y - v1.y) - (v2.y - v1.y) * (v3p.x - v1.x);
  const double s = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
  nx *= s; ny *= s; nz *= s;
  return (v4.x - v1.x) * nx + (v4.y - v1.y) * ny + (v4.z - v1.z) * nz;
}

// IsContact_FV_Proximity, bvh_aabb.cpp:392-412 (mgn forced to 0 at :407)
__device__ __forceinline__ bool is_fv_prx(int i0, int i1, int i2, int i3, V3 p0, V3 p1, V3 p2, V3 p3, const Box& bb,
                                          double delta) {
  if (i3 == i0 || i3 == i1 || i3 == i2) return false;
  if (!box_inside(bb, p3)) return false;
  if (fabs(tet_height(p0, p1, p2, p3)) > delta) return false;
  double w0, w1;
  const double dist = dist_face_vertex(p0, p1, p2, p3, w0, w1);
  const double w2 = 1 - w0 - w1;
  if (dist > delta) return false;
  if (w0 < -0.0 || w0 > 1 + 0.0) return false;
  if (w1 < -0.0 || w1 > 1 + 0.0) return false;
  if (w2 < -0.0 || w2 > 1 + 0.0) return false;
  return true;
}

// IsContact_EE_Proximity, bvh_aabb.cpp:415-438
__device__ __forceinline__ bool is_ee_prx(int i0, int i1, int j0, int j1, V3 p0, V3 p1, V3 q0, V3 q1, double delta) {
  if (i0 == j0 || i0 == j1 || i1 == j0 || i1 == j1) return false;
  if (q0.x + delta < p0.x && q0.x + delta < p1.x && q1.x + delta < p0.x && q1.x + delta < p1.x) return false;
  if (q0.x - delta > p0.x && q0.x - delta > p1.x && q1.x - delta > p0.x && q1.x - delta > p1.x) return false;
  if (q0.y + delta < p0.y && q0.y + delta < p1.y && q1.y + delta < p0.y && q1.y + delta < p1.y) return false;
  if (q0.y - delta > p0.y && q0.y - delta > p1.y && q1.y - delta > p0.y && q1.y - delta > p1.y) return false;
  if (q0.z + delta < p0.z && q0.z + delta < p1.z && q1.z + delta < p0.z && q1.z + delta < p1.z) return false;
  if (q0.z - delta > p0.z && q0.z - delta > p1.z && q1.z - delta > p0.z && q1.z - delta > p1.z) return false;
  double rp, rq;
  const double dist = dist_edge_edge(p0, p1, q0, q1, rp, rq);
  if (dist > delta) return false;
  if (rp < 0) return false;
  if (rp > 1) return false;
  if (rq < 0) return false;
  if (rq > 1) return false;
  const V3 pm = (1 - rp) * p0 + rp * p1;
  const V3 qm = (1 - rq) * q0 + rq * q1;
  if (length(pm - qm) > delta) return false;
  return true;
}

// EvaluateCubic, bvh_aabb.cpp:563-568
__device__ __forceinline__ double cubic(double r, double k0, double k1, double k2, double k3) {
  return k0 + k1 * r + k2 * r * r + k3 * r * r * r;
}
// FindRootCubic + BisectRangeCubicRoot, bvh_aabb.cpp:571-595: the counter is pre-decremented (15 -> 14 halvings)
// and the sign test always uses the value at the ORIGINAL left end.
__device__ __forceinline__ double cubic_root(double r0, double r1, double v0, double k0, double k1, double k2, double k3) {
  for (int n = 0; n < 14; n++) {
    const double r2 = 0.5 * (r0 + r1);
    const double v2 = cubic(r2, k0, k1, k2, k3);
    if (v0 * v2 < 0) r1 = r2; else r0 = r2;
  }
  return 0.5 * (r0 + r1);
}
// The same bisection with the next level evaluated speculatively: both candidate midpoints of halving n+1 (and their
// cubic values) are computed while the sign test of halving n resolves, so two halvings cost one dependent
// evaluation plus a select.  Identical arithmetic on the path taken, hence identical result; used where the chain
// latency counts (Gauss-Seidel impulses), not in the throughput-bound narrow phase (twice the flops).
__device__ __forceinline__ double cubic_root_spec(double r0, double r1, double v0, double k0, double k1, double k2, double k3) {
#pragma unroll
  for (int n = 0; n < 14; n += 2) {
    const double r2 = 0.5 * (r0 + r1);
    const double ra = 0.5 * (r0 + r2), rb = 0.5 * (r2 + r1);  // next midpoint if the root is left / right of r2
    const double v2 = cubic(r2, k0, k1, k2, k3);
    const double va = cubic(ra, k0, k1, k2, k3), vb = cubic(rb, k0, k1, k2, k3);
    const bool left = v0 * v2 < 0;
    if (left) r1 = r2; else r0 = r2;
    const double r3 = left ? ra : rb, v3 = left ? va : vb;
    if (v0 * v3 < 0) r1 = r3; else r0 = r3;
  }
  return 0.5 * (r0 + r1);
}
// FindCoplanerInterp, bvh_aabb.cpp:598-662
// (body inlined where latency counts -- the Gauss-Seidel impulse chain --, a call in the throughput-bound narrow phase)
template <bool SPEC>
static __device__ __forceinline__ double coplanar_time_inl(V3 s0, V3 s1, V3 s2, V3 s3, V3 e0, V3 e1, V3 e2, V3 e3) {
#define cubic_root(A_, B_, C_, D_, E_, F_, G_) (SPEC ? cubic_root_spec(A_, B_, C_, D_, E_, F_, G_) : cubic_root(A_, B_, C_, D_, E_, F_, G_))
  const V3 x1 = s1 - s0, x2 = s2 - s0, x3 = s3 - s0;
  const V3 v1 = (e1 - e0) - x1, v2 = (e2 - e0) - x2, v3v = (e3 - e0) - x3;
  const double k0 = triple(x3, x1, x2);
  const double k1 = triple(v3v, x1, x2) + triple(x3, v1, x2) + triple(x3, x1, v2);
  const double k2 = triple(v3v, v1, x2) + triple(v3v, x1, v2) + triple(x3, v1, v2);
  const double k3 = triple(v3v, v1, v2);
  const double r0 = -0.0, r1 = +1.0;
  const double f0 = cubic(r0, k0, k1, k2, k3), f1 = cubic(r1, k0, k1, k2, k3);
  const double det = k2 * k2 - 3 * k1 * k3;
  if (fabs(k3) < 1.0e-10 && fabs(k2) > 1.0e-10) {
    const double r2 = -k1 / (2 * k2);
    const double f2 = cubic(r2, k0, k1, k2, k3);
    if (r2 > 0 && r2 < 1) {
      if (f0 * f2 < 0) return cubic_root(r0, r2, f0, k0, k1, k2, k3);
      else if (f2 * f1 < 0) return cubic_root(r2, r1, f2, k0, k1, k2, k3);
    }
  }
  if (det > 0 && fabs(k3) > 1.0e-10) {
    const double r3 = (-k2 - sqrt(det)) / (3 * k3);
    const double f3 = cubic(r3, k0, k1, k2, k3);
    if (r3 > 0 && r3 < 1) {
      if (f0 * f3 < 0) return cubic_root(r0, r3, f0, k0, k1, k2, k3);
      else if (f3 * f1 < 0) return cubic_root(r3, r1, f3, k0, k1, k2, k3);
    }
    const double r4 = (-k2 + sqrt(det)) / (3 * k3);
    const double f4 = cubic(r4, k0, k1, k2, k3);
    if (r3 > 0 && r3 < 1 && r4 > 0 && r4 < 1) {
      if (f3 * f4 < 0) return cubic_root(r3, r4, f3, k0, k1, k2, k3);
    }
    if (r4 > 0 && r4 < 1) {
      if (f0 * f4 < 0) return cubic_root(r0, r4, f0, k0, k1, k2, k3);
      else if (f4 * f1 < 0) return cubic_root(r4, r1, f4, k0, k1, k2, k3);
    }
  }
  if (f0 * f1 > 0) return -1;
  return cubic_root(r0, r1, f0, k0, k1, k2, k3);
}

#undef cubic_root
static __device__ __noinline__ double coplanar_time(V3 s0, V3 s1, V3 s2, V3 s3, V3 e0, V3 e1, V3 e2, V3 e3) {
  return coplanar_time_inl<false>(s0, s1, s2, s3, e0, e1, e2, e3);
}

// GetMinDist_LineSegPoint, vector3d.h:282-295
__device__ __forceinline__ V3 closest_on_segment(V3 p, V3 s, V3 e) {
  const V3 d = e - s;
  const double a = dot(d, d);
  const double b = dot(d, s - p);
  double t = -b / a;
  if (t < 0) t = 0;
  if (t > 1) t = 1;
  return s + t * d;
}

// IsContact_FV_CCD, bvh_aabb.cpp:666-719
__device__ __forceinline__ bool is_fv_ccd(int i0, int i1, int i2, int i3, V3 p0, V3 p1, V3 p2, V3 p3, V3 q0, V3 q1,
                                          V3 q2, V3 q3, const Box& bb) {
  const double eps = 1.0e-10;
  if (i3 == i0 || i3 == i1 || i3 == i2) return false;
  {
    Box bbp;
    box_init(bbp, p3, eps);
    box_add(bbp, q3, eps);
    if (!box_intersect(bb, bbp)) return false;
  }
  {
    const V3 n = cross(p1 - p0, p2 - p0);
    const double t0 = dot(p0 - p3, n), t1 = dot(q0 - q3, n);
    const double t2 = dot(q1 - q3, n), t3 = dot(q2 - q3, n);
    if (t0 * t1 > 0 && t0 * t2 > 0 && t0 * t3 > 0) return false;
  }
  double r0, r1;
  const double dist = dist_face_vertex(p0, p1, p2, p3, r0, r1);
  {
    const double vn0 = length(p0 - q0), vn1 = length(p1 - q1), vn2 = length(p2 - q2), vn3 = length(p3 - q3);
    double vnt = (vn0 > vn1) ? vn0 : vn1;
    vnt = (vn2 > vnt) ? vn2 : vnt;
    const double max_app = vnt + vn3;
    const double r2 = 1 - r0 - r1;
    if (dist > max_app) return false;
    if (r0 < 0 || r0 > 1 || r1 < 0 || r1 > 1 || r2 < 0 || r2 > 1) {
      const double d01 = length(closest_on_segment(p3, p0, p1) - p3);
      const double d12 = length(closest_on_segment(p3, p1, p2) - p3);
      const double d20 = length(closest_on_segment(p3, p2, p0) - p3);
      if (d01 > max_app && d12 > max_app && d20 > max_app) return false;
    }
  }
  const double t = coplanar_time(p0, p1, p2, p3, q0, q1, q2, q3);
  if (t < 0 || t > 1) return false;
  double w0, w1;
  {
    const V3 p0m = (1 - t) * p0 + t * q0;
    const V3 p1m = (1 - t) * p1 + t * q1;
    const V3 p2m = (1 - t) * p2 + t * q2;
    const V3 p3m = (1 - t) * p3 + t * q3;
    dist_face_vertex(p0m, p1m, p2m, p3m, w0, w1);
  }
  const double w2 = 1 - w0 - w1;
  if (w0 < 0 || w0 > 1) return false;
  if (w1 < 0 || w1 > 1) return false;
  if (w2 < 0 || w2 > 1) return false;
  return true;
}

// IsContact_EE_CCD, bvh_aabb.cpp:723-793 (CSAT / max-approach culls are commented out in the reference; the
// distance gate is the literal 1e-2)
__device__ __forceinline__ bool is_ee_ccd(int i0, int i1, int j0, int j1, V3 p0s, V3 p1s, V3 q0s, V3 q1s, V3 p0e,
                                          V3 p1e, V3 q0e, V3 q1e) {
  const double eps = 1.0e-10;
  if (i0 == j0 || i0 == j1 || i1 == j0 || i1 == j1) return false;
  {
    Box bbq, bbp;
    box_init(bbq, q0s, eps); box_add(bbq, q1s, eps); box_add(bbq, q0e, eps); box_add(bbq, q1e, eps);
    box_init(bbp, p0s, eps); box_add(bbp, p1s, eps); box_add(bbp, p0e, eps); box_add(bbp, p1e, eps);
    if (!box_intersect(bbp, bbq)) return false;
  }
  const double t = coplanar_time(p0s, p1s, q0s, q1s, p0e, p1e, q0e, q1e);
  if (t < 0 || t > 1) return false;
  double w0, w1, dist;
  {
    const V3 p0m = (1 - t) * p0s + t * p0e;
    const V3 p1m = (1 - t) * p1s + t * p1e;
    const V3 q0m = (1 - t) * q0s + t * q0e;
    const V3 q1m = (1 - t) * q1s + t * q1e;
    dist = dist_edge_edge(p0m, p1m, q0m, q1m, w0, w1);
  }
  if (w0 < 0 || w0 > 1) return false;
  if (w1 < 0 || w1 > 1) return false;
  if (dist > 1.0e-2) return false;
  return true;
}

// ---- cheap prefixes of the four predicates above (shared-vertex test and the box / slab / sign culls, in the same
// order and with the same arithmetic): a conservative filter -- whatever survives is re-tested with the full
// predicate, so the contact sets are unchanged
__device__ __forceinline__ bool maybe_fv_prx(int i0, int i1, int i2, int i3, V3 p3, const Box& bb) {
  if (i3 == i0 || i3 == i1 || i3 == i2) return false;
  return box_inside(bb, p3);
}
__device__ __forceinline__ bool maybe_ee_prx(int i0, int i1, int j0, int j1, V3 p0, V3 p1, V3 q0, V3 q1, double delta) {
  if (i0 == j0 || i0 == j1 || i1 == j0 || i1 == j1) return false;
  if (q0.x + delta < p0.x && q0.x + delta < p1.x && q1.x + delta < p0.x && q1.x + delta < p1.x) return false;
  if (q0.x - delta > p0.x && q0.x - delta > p1.x && q1.x - delta > p0.x && q1.x - delta > p1.x) return false;
  if (q0.y + delta < p0.y && q0.y + delta < p1.y && q1.y + delta < p0.y && q1.y + delta < p1.y) return false;
  if (q0.y - delta > p0.y && q0.y - delta > p1.y && q1.y - delta > p0.y && q1.y - delta > p1.y) return false;
  if (q0.z + delta < p0.z && q0.z + delta < p1.z && q1.z + delta < p0.z && q1.z + delta < p1.z) return false;
  if (q0.z - delta > p0.z && q0.z - delta > p1.z && q1.z - delta > p0.z && q1.z - delta > p1.z) return false;
  return true;
}
__device__ __forceinline__ bool maybe_fv_ccd(int i0, int i1, int i2, int i3, V3 p0, V3 p1, V3 p2, V3 p3, V3 q0, V3 q1,
                                             V3 q2, V3 q3, const Box& bb) {
  const double eps = 1.0e-10;
  if (i3 == i0 || i3 == i1 || i3 == i2) return false;
  Box bbp;
  box_init(bbp, p3, eps);
  box_add(bbp, q3, eps);
  if (!box_intersect(bb, bbp)) return false;
  const V3 n = cross(p1 - p0, p2 - p0);
  const double t0 = dot(p0 - p3, n), t1 = dot(q0 - q3, n);
  const double t2 = dot(q1 - q3, n), t3 = dot(q2 - q3, n);
  if (t0 * t1 > 0 && t0 * t2 > 0 && t0 * t3 > 0) return false;
  return true;
}
__device__ __forceinline__ bool maybe_ee_ccd(int i0, int i1, int j0, int j1, V3 p0s, V3 p1s, V3 q0s, V3 q1s, V3 p0e,
                                             V3 p1e, V3 q0e, V3 q1e) {
  const double eps = 1.0e-10;
  if (i0 == j0 || i0 == j1 || i1 == j0 || i1 == j1) return false;
  Box bbq, bbp;
  box_init(bbq, q0s, eps); box_add(bbq, q1s, eps); box_add(bbq, q0e, eps); box_add(bbq, q1e, eps);
  box_init(bbp, p0s, eps); box_add(bbp, p1s, eps); box_add(bbp, p0e, eps); box_add(bbp, p1e, eps);
  return box_intersect(bbp, bbq);
}

}  // namespace cg
