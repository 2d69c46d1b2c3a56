// include/clothgpu_compat/vector3d.h -- the reference's 3-vector value type (vector3d.h:33-295) for callers of
// DistanceFaceVertex / DistanceEdgeEdge / FindCoplanerInterp (bvh_aabb.h:49-61).  Plain host value type: it carries
// coordinates across the API and is not part of the per-step computation.  Semantics kept: public x, y, z; a virtual
// destructor (sizeof == 32, as in the reference); operator/= leaves the vector unchanged when |d| < 1e-16;
// operator== compares with that tolerance; operator* of two vectors is the dot product, operator^ the cross product.
#ifndef CLOTHGPU_COMPAT_VECTOR3D_H
#define CLOTHGPU_COMPAT_VECTOR3D_H

#include <cassert>
#include <cmath>
#include <iostream>

#ifndef NEARLY_ZERO
#define NEARLY_ZERO 1.e-16
#endif

class CVector3D {
public:
  CVector3D() : x(0.0), y(0.0), z(0.0) {}
  CVector3D(double vx, double vy, double vz) : x(vx), y(vy), z(vz) {}
  CVector3D(const CVector3D& r) : x(r.x), y(r.y), z(r.z) {}
  virtual ~CVector3D() {}

  void SetVector(double vx, double vy, double vz) { x = vx; y = vy; z = vz; }
  void SetZero() { x = y = z = 0.0; }
  CVector3D& operator=(const CVector3D& r) { x = r.x; y = r.y; z = r.z; return *this; }
  CVector3D& operator+=(const CVector3D& r) { x += r.x; y += r.y; z += r.z; return *this; }
  CVector3D& operator-=(const CVector3D& r) { x -= r.x; y -= r.y; z -= r.z; return *this; }
  CVector3D& operator*=(double d) { x *= d; y *= d; z *= d; return *this; }
  CVector3D& operator/=(double d) {
    if (std::fabs(d) < NEARLY_ZERO) return *this;  // the reference's silent no-op (vector3d.h:63-67)
    x /= d; y /= d; z /= d;
    return *this;
  }
  CVector3D operator+() const { return *this; }
  CVector3D operator-() const { return CVector3D(-x, -y, -z); }
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : 0.0)); }
  double& operator[](int i) {
    assert(i >= 0 && i < 3);
    return i == 0 ? x : (i == 1 ? y : z);
  }
  double DLength() const { return x * x + y * y + z * z; }
  double Length() const { return std::sqrt(DLength()); }
  void SetNormalizedVector() {
    const double s = 1.0 / Length();
    x *= s; y *= s; z *= s;
  }

public:
  double x, y, z;
};

inline CVector3D operator+(const CVector3D& a, const CVector3D& b) { CVector3D t(a); t += b; return t; }
inline CVector3D operator-(const CVector3D& a, const CVector3D& b) { CVector3D t(a); t -= b; return t; }
inline CVector3D operator*(double d, const CVector3D& a) { CVector3D t(a); t *= d; return t; }
inline CVector3D operator*(const CVector3D& a, double d) { CVector3D t(a); t *= d; return t; }
inline CVector3D operator/(const CVector3D& a, double d) { CVector3D t(a); t /= d; return t; }

inline double Dot(const CVector3D& a, const CVector3D& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline CVector3D Cross(const CVector3D& a, const CVector3D& b) {
  return CVector3D(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline double operator*(const CVector3D& a, const CVector3D& b) { return Dot(a, b); }
inline CVector3D operator^(const CVector3D& a, const CVector3D& b) { return Cross(a, b); }
inline bool operator==(const CVector3D& a, const CVector3D& b) {
  return std::fabs(a.x - b.x) < NEARLY_ZERO && std::fabs(a.y - b.y) < NEARLY_ZERO && std::fabs(a.z - b.z) < NEARLY_ZERO;
}
inline bool operator!=(const CVector3D& a, const CVector3D& b) { return !(a == b); }
inline std::ostream& operator<<(std::ostream& os, const CVector3D& v) {
  os.setf(std::ios::scientific);
  return os << v.x << " " << v.y << " " << v.z;
}

inline double Length(const CVector3D& p) { return p.Length(); }
inline double SquareDistance(const CVector3D& a, const CVector3D& b) { return (b - a).DLength(); }
inline double Distance(const CVector3D& a, const CVector3D& b) { return std::sqrt(SquareDistance(a, b)); }
inline double ScalarTripleProduct(const CVector3D& a, const CVector3D& b, const CVector3D& c) {
  return a.x * (b.y * c.z - b.z * c.y) + a.y * (b.z * c.x - b.x * c.z) + a.z * (b.x * c.y - b.y * c.x);
}
// six times the signed volume, over six
inline double TetVolume(const CVector3D& v1, const CVector3D& v2, const CVector3D& v3, const CVector3D& v4) {
  return ScalarTripleProduct(v2 - v1, v3 - v1, v4 - v1) / 6.0;
}
inline void UnitNormal(CVector3D& n, const CVector3D& v1, const CVector3D& v2, const CVector3D& v3) {
  n = Cross(v2 - v1, v3 - v1);
  n.SetNormalizedVector();
}
// signed distance of v4 from the plane of (v1, v2, v3)
inline double Height(const CVector3D& v1, const CVector3D& v2, const CVector3D& v3, const CVector3D& v4) {
  CVector3D n;
  UnitNormal(n, v1, v2, v3);
  return Dot(v4 - v1, n);
}
// closest point to p on the line through s with direction d / on the segment s-e
inline CVector3D GetMinDist_LinePoint(const CVector3D& p, const CVector3D& s, const CVector3D& d) {
  const double t = -Dot(d, s - p) / Dot(d, d);
  return s + t * d;
}
inline CVector3D GetMinDist_LineSegPoint(const CVector3D& p, const CVector3D& s, const CVector3D& e) {
  const CVector3D d = e - s;
  double t = -Dot(d, s - p) / Dot(d, d);
  t = t < 0 ? 0 : (t > 1 ? 1 : t);
  return s + t * d;
}

#endif  // CLOTHGPU_COMPAT_VECTOR3D_H
