// include/clothgpu_compat/utility.h -- the double[3] helpers and the camera quaternions the reference demos use
// (utility.h:15-123: scene set-up, MakeNormal, view rotation).  Host-side convenience for callers; the per-step
// arithmetic of the library lives in the CUDA kernels.
#ifndef CLOTHGPU_COMPAT_UTILITY_H
#define CLOTHGPU_COMPAT_UTILITY_H

#include <cmath>

inline double Dot3D(const double a[], const double b[]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void Cross3D(double r[3], const double a[3], const double b[3]) {
  for (int i = 0; i < 3; i++) {
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    r[i] = a[j] * b[k] - b[j] * a[k];
  }
}
inline double Length3D(const double v[3]) { return std::sqrt(Dot3D(v, v)); }
inline double Distance3D(const double a[3], const double b[3]) {
  const double d[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
  return Length3D(d);
}
namespace clothgpu_compat_detail {
// (v2 - v1) x (v3 - v1)
inline void TriCross(double n[3], const double v1[3], const double v2[3], const double v3[3]) {
  const double e[3] = {v2[0] - v1[0], v2[1] - v1[1], v2[2] - v1[2]};
  const double f[3] = {v3[0] - v1[0], v3[1] - v1[1], v3[2] - v1[2]};
  Cross3D(n, e, f);
}
}  // namespace clothgpu_compat_detail
inline double TriArea3D(const double v1[3], const double v2[3], const double v3[3]) {
  double n[3];
  clothgpu_compat_detail::TriCross(n, v1, v2, v3);
  return Length3D(n) * 0.5;
}
// unit normal n and area a of the triangle (v1, v2, v3)
inline void UnitNormalAreaTri3D(double n[3], double& a, const double v1[3], const double v2[3], const double v3[3]) {
  clothgpu_compat_detail::TriCross(n, v1, v2, v3);
  a = Length3D(n) * 0.5;
  const double s = 0.5 / a;
  n[0] *= s; n[1] *= s; n[2] *= s;
}

// quaternions (w, x, y, z) of the demo camera
inline void QuatMult(double r[], const double p[], const double q[]) {
  r[0] = p[0] * q[0] - (p[1] * q[1] + p[2] * q[2] + p[3] * q[3]);
  double c[3];
  Cross3D(c, p + 1, q + 1);
  for (int i = 0; i < 3; i++) r[i + 1] = p[0] * q[i + 1] + q[0] * p[i + 1] + c[i];
}
inline void QuatCopy(double r[], const double p[]) {
  for (int i = 0; i < 4; i++) r[i] = p[i];
}
// column-major 4x4 rotation matrix (glMultMatrixd layout) of the unit quaternion q
inline void QuatRot(double r[], const double q[]) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  for (int i = 0; i < 16; i++) r[i] = 0.0;
  r[0] = 1.0 - 2.0 * (y * y + z * z); r[1] = 2.0 * (x * y + z * w);       r[2] = 2.0 * (z * x - y * w);
  r[4] = 2.0 * (x * y - z * w);       r[5] = 1.0 - 2.0 * (z * z + x * x); r[6] = 2.0 * (y * z + x * w);
  r[8] = 2.0 * (z * x + y * w);       r[9] = 2.0 * (y * z - x * w);       r[10] = 1.0 - 2.0 * (x * x + y * y);
  r[15] = 1.0;
}

#endif  // CLOTHGPU_COMPAT_UTILITY_H
