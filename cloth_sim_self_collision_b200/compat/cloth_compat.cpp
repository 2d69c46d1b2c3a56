// compat/cloth_compat.cpp -- bodies of the reference-named C++ API (include/clothgpu_compat/cloth_compat.h).
//
// Host code only: marshals the caller's std::vectors into the C ABI of libclothgpu.so (include/clothgpu.h).  Every
// numerical operation of the hot path runs in the CUDA library; nothing here computes on the CPU except the
// setup-side helpers the reference also runs once at start-up (adjacency lists, BVH topology) and the evaluation of
// the caller's own collider callback.
#include "cloth_compat.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>

#include "../../include/clothgpu.h"

namespace {

bool g_verbose = false;
int g_last_iter = 0;
double g_last_conv = 0.0;

void check(int rc, const char* what) {
  if (rc != CLOTHGPU_OK)
    throw std::runtime_error(std::string(what) + ": clothgpu error " + std::to_string(rc) + ": " + clothgpu_last_error());
}

// one cached device context per mesh (keyed by the address of the caller's triangle array), recreated when a call
// brings topology the context was not built with (e.g. the first contact call after the first implicit step)
struct Scene {
  clothgpu_ctx* ctx = nullptr;
  int np = 0;
  std::vector<int> tri, quad, bc, rowptr, colidx, bvh, eidx, earr;
  std::vector<double> xyz0;
  bool dirty = true;
};
std::map<const void*, Scene>& scenes() {
  static std::map<const void*, Scene> s;
  return s;
}

template <class T>
bool assign_if_changed(std::vector<T>& dst, const T* src, size_t n) {
  if (dst.size() == n && (n == 0 || memcmp(dst.data(), src, n * sizeof(T)) == 0)) return false;
  dst.assign(src, src + n);
  return true;
}

void pattern_from_quads(const std::vector<int>& quad, int np, std::vector<int>& rowptr, std::vector<int>& colidx) {
  CJaggedArray crs;
  crs.SetEdgeOfElem(quad, (int)quad.size() / 4, 4, np, false);
  crs.Sort();
  rowptr = crs.index;
  colidx = crs.array;
}

clothgpu_ctx* ensure_ctx(Scene& s) {
  if (s.ctx && !s.dirty) return s.ctx;
  if (s.ctx) { clothgpu_destroy(s.ctx); s.ctx = nullptr; }
  static const int dummy_i[4] = {0, 0, 0, 0};
  if (s.bc.empty()) s.bc.assign(s.np, 0);
  if (s.xyz0.empty()) s.xyz0.assign((size_t)s.np * 3, 0.0);
  if (s.rowptr.empty()) {
    if (!s.quad.empty()) pattern_from_quads(s.quad, s.np, s.rowptr, s.colidx);
    else s.rowptr.assign(s.np + 1, 0);
  }
  const int nnode = (int)s.bvh.size() / 3;
  check(clothgpu_create(&s.ctx, -1, 1, s.np, (int)s.tri.size() / 3, (int)s.quad.size() / 4,
                        s.tri.empty() ? dummy_i : s.tri.data(), s.quad.empty() ? dummy_i : s.quad.data(), s.bc.data(),
                        s.xyz0.data(), s.rowptr.data(), s.colidx.empty() ? dummy_i : s.colidx.data(), nnode,
                        nnode ? s.bvh.data() : nullptr, nnode ? s.eidx.data() : nullptr,
                        nnode ? (s.earr.empty() ? dummy_i : s.earr.data()) : nullptr),
        "clothgpu_create");
  s.dirty = false;
  return s.ctx;
}

Scene& scene_for(const std::vector<int>& aTri, int np) {
  Scene& s = scenes()[aTri.data()];
  if (s.np != np) { s.np = np; s.dirty = true; }
  if (assign_if_changed(s.tri, aTri.data(), aTri.size())) s.dirty = true;
  return s;
}
void add_internal(Scene& s, const std::vector<double>& aXYZ0, const std::vector<int>& aBCFlag,
                  const std::vector<int>& aQuad, const CMatrixSquareSparse* mat) {
  if (assign_if_changed(s.quad, aQuad.data(), aQuad.size())) { s.dirty = true; s.rowptr.clear(); s.colidx.clear(); }
  if (assign_if_changed(s.bc, aBCFlag.data(), aBCFlag.size())) s.dirty = true;
  if (assign_if_changed(s.xyz0, aXYZ0.data(), aXYZ0.size())) s.dirty = true;
  if (mat && mat->m_colInd) {
    if (assign_if_changed(s.rowptr, mat->m_colInd, (size_t)mat->m_nblk + 1)) s.dirty = true;
    if (assign_if_changed(s.colidx, mat->m_rowPtr, (size_t)mat->m_ncrs)) s.dirty = true;
  }
}
void add_contact(Scene& s, const std::vector<CNodeBVH>& nodes, const CJaggedArray* edge) {
  static_assert(sizeof(CNodeBVH) == 3 * sizeof(int), "CNodeBVH must be three ints");
  if (assign_if_changed(s.bvh, reinterpret_cast<const int*>(nodes.data()), nodes.size() * 3)) s.dirty = true;
  if (edge) {
    if (assign_if_changed(s.eidx, edge->index.data(), edge->index.size())) s.dirty = true;
    if (assign_if_changed(s.earr, edge->array.data(), edge->array.size())) s.dirty = true;
  } else if (s.eidx.empty()) {  // fine-grained BVH calls do not need the adjacency: an empty one
    s.eidx.assign(s.np + 1, 0);
    s.dirty = true;
  }
}

clothgpu_params make_params(double dt, double lambda, double myu, double stiff_bend, const double* gravity,
                            double mass_point, double stiff_contact, double clearance) {
  clothgpu_params p;
  memset(&p, 0, sizeof(p));
  p.dt = dt; p.lambda = lambda; p.myu = myu; p.stiff_bend = stiff_bend;
  if (gravity) { p.gravity[0] = gravity[0]; p.gravity[1] = gravity[1]; p.gravity[2] = gravity[2]; }
  p.mass_point = mass_point; p.stiff_contact = stiff_contact; p.contact_clearance = clearance;
  p.collider_kind = CLOTHGPU_COLLIDER_NONE;
  return p;
}

// the caller's collider is a host function pointer (solve_internal_sparse.h:274): evaluate it per vertex on the host
// and hand the (pd, n) table to the device -- the generic path of SURVEY.md 7.3-6
void upload_collider(clothgpu_ctx* c, clothgpu_params& p, const std::vector<double>& aXYZ,
                     void (*penetrationDepth)(double&, double*, const double*)) {
  if (!penetrationDepth) return;
  const int np = (int)aXYZ.size() / 3;
  std::vector<double> table((size_t)np * 4);
  for (int ip = 0; ip < np; ip++) {
    double pd = 0, n[3] = {0, 0, 0};
    penetrationDepth(pd, n, &aXYZ[(size_t)ip * 3]);
    table[(size_t)ip * 4 + 0] = pd; table[(size_t)ip * 4 + 1] = n[0];
    table[(size_t)ip * 4 + 2] = n[1]; table[(size_t)ip * 4 + 3] = n[2];
  }
  check(clothgpu_write(c, CLOTHGPU_BUF_COLLIDER_TABLE, 0, table.data(), (long long)table.size()), "collider table");
  p.collider_kind = CLOTHGPU_COLLIDER_TABLE;
}

void boxes_to_vec(clothgpu_ctx* c, size_t nnode, std::vector<CAABB3D>& aBB) {
  std::vector<double> b(nnode * 6);
  check(clothgpu_read(c, CLOTHGPU_BUF_BOXES, 0, b.data(), (long long)b.size()), "read boxes");
  aBB.resize(nnode);
  for (size_t i = 0; i < nnode; i++) {
    CAABB3D& o = aBB[i];
    o.x_min = b[i * 6 + 0]; o.x_max = b[i * 6 + 1]; o.y_min = b[i * 6 + 2]; o.y_max = b[i * 6 + 3];
    o.z_min = b[i * 6 + 4]; o.z_max = b[i * 6 + 5];
    o.isnt_empty = o.x_min <= o.x_max;
  }
}
void vec_to_boxes(clothgpu_ctx* c, const std::vector<CAABB3D>& aBB) {
  std::vector<double> b(aBB.size() * 6);
  for (size_t i = 0; i < aBB.size(); i++) {
    const CAABB3D& o = aBB[i];
    if (o.isnt_empty) {
      b[i * 6 + 0] = o.x_min; b[i * 6 + 1] = o.x_max; b[i * 6 + 2] = o.y_min; b[i * 6 + 3] = o.y_max;
      b[i * 6 + 4] = o.z_min; b[i * 6 + 5] = o.z_max;
    } else {  // an empty box intersects nothing
      b[i * 6 + 0] = b[i * 6 + 2] = b[i * 6 + 4] = 1.0;
      b[i * 6 + 1] = b[i * 6 + 3] = b[i * 6 + 5] = -1.0;
    }
  }
  check(clothgpu_write(c, CLOTHGPU_BUF_BOXES, 0, b.data(), (long long)b.size()), "write boxes");
}
void elems_to_set(clothgpu_ctx* c, int n, std::set<CContactElement>& out) {
  std::vector<int> e((size_t)n * 5);
  if (n > 0) check(clothgpu_read(c, CLOTHGPU_BUF_ELEMS, 0, e.data(), (long long)e.size()), "read elems");
  for (int i = 0; i < n; i++) {
    const int* q = &e[(size_t)i * 5];
    out.insert(CContactElement(q[4] != 0, q[0], q[1], q[2], q[3]));
  }
}

// pattern-only context of a matrix (fine-grained linear algebra calls)
clothgpu_ctx* matrix_ctx(const CMatrixSquareSparse& m) {
  if (m.m_dev) return m.m_dev;
  if (m.m_len != 3) throw std::runtime_error("CMatrixSquareSparse: only 3x3 blocks run on the GPU path");
  static const int dummy_i[4] = {0, 0, 0, 0};
  std::vector<int> bc(m.m_nblk, 0);
  std::vector<double> x0((size_t)m.m_nblk * 3, 0.0);
  check(clothgpu_create(&m.m_dev, -1, 1, m.m_nblk, 0, 0, dummy_i, dummy_i, bc.data(), x0.data(), m.m_colInd,
                        m.m_ncrs ? m.m_rowPtr : dummy_i, 0, nullptr, nullptr, nullptr),
        "clothgpu_create (matrix)");
  return m.m_dev;
}
void upload_matrix(clothgpu_ctx* c, const CMatrixSquareSparse& m, int buf_dia, int buf_crs) {
  check(clothgpu_write(c, buf_dia, 0, m.m_valDia, (long long)m.m_nblk * 9), "write diagonal blocks");
  if (m.m_ncrs) check(clothgpu_write(c, buf_crs, 0, m.m_valCrs, (long long)m.m_ncrs * 9), "write off-diagonal blocks");
}

}  // namespace

namespace clothgpu_compat {
void SetVerbose(bool on) { g_verbose = on; }
int LastIteration() { return g_last_iter; }
double LastConvRatio() { return g_last_conv; }
void ReleaseAll() {
  for (auto& kv : scenes())
    if (kv.second.ctx) clothgpu_destroy(kv.second.ctx);
  scenes().clear();
}
}  // namespace clothgpu_compat

// ------------------------------------------------------------------------------------------------ CJaggedArray
void CJaggedArray::InitializeSize(int size) {
  index.assign(size + 1, 0);
  array.clear();
}
int CJaggedArray::Size() const { return index.empty() ? 0 : (int)index.size() - 1; }
void CJaggedArray::SetEdgeOfElem(const std::vector<int>& aElem, int nelem, int nnoel, int nnode, bool is_dia) {
  std::vector<std::vector<int> > adj(nnode);
  for (int e = 0; e < nelem; e++)
    for (int a = 0; a < nnoel; a++)
      for (int b = 0; b < nnoel; b++) {
        const int i = aElem[e * nnoel + a], j = aElem[e * nnoel + b];
        if (i == j && !is_dia) continue;
        adj[i].push_back(j);
      }
  index.assign(nnode + 1, 0);
  array.clear();
  for (int i = 0; i < nnode; i++) {
    std::vector<int>& v = adj[i];
    // first-seen order with duplicates removed (the hot path treats rows as sets; Sort() makes them ascending)
    std::vector<int> seen;
    for (size_t k = 0; k < v.size(); k++)
      if (std::find(seen.begin(), seen.end(), v[k]) == seen.end()) seen.push_back(v[k]);
    array.insert(array.end(), seen.begin(), seen.end());
    index[i + 1] = (int)array.size();
  }
}
void CJaggedArray::Sort() {
  for (int i = 0; i + 1 < (int)index.size(); i++) std::sort(array.begin() + index[i], array.begin() + index[i + 1]);
}

// ------------------------------------------------------------------------------------------------ CAABB3D
CAABB3D::CAABB3D() : x_min(0), x_max(0), y_min(0), y_max(0), z_min(0), z_max(0), isnt_empty(false) {}
CAABB3D::CAABB3D(double x0, double x1, double y0, double y1, double z0, double z1)
    : x_min(x0), x_max(x1), y_min(y0), y_max(y1), z_min(z0), z_max(z1), isnt_empty(true) {}
bool CAABB3D::IsIntersect(const CAABB3D& bb) const {
  if (!isnt_empty || !bb.isnt_empty) return false;
  return !(x_max < bb.x_min || x_min > bb.x_max || y_max < bb.y_min || y_min > bb.y_max || z_max < bb.z_min ||
           z_min > bb.z_max);
}
CAABB3D& CAABB3D::operator+=(const CAABB3D& bb) {
  if (!bb.isnt_empty) return *this;
  if (!isnt_empty) { *this = bb; return *this; }
  x_min = std::min(x_min, bb.x_min); x_max = std::max(x_max, bb.x_max);
  y_min = std::min(y_min, bb.y_min); y_max = std::max(y_max, bb.y_max);
  z_min = std::min(z_min, bb.z_min); z_max = std::max(z_max, bb.z_max);
  return *this;
}
bool CAABB3D::AddPoint(double x, double y, double z, double eps) {
  if (eps <= 0) return false;
  CAABB3D b(x - eps, x + eps, y - eps, y + eps, z - eps, z + eps);
  *this += b;
  return true;
}
bool CAABB3D::IsInside(double x, double y, double z) const {
  return isnt_empty && x >= x_min && x <= x_max && y >= y_min && y <= y_max && z >= z_min && z <= z_max;
}

// ------------------------------------------------------------------------------------------------ CContactElement
CContactElement::CContactElement(bool fv, int j0, int j1, int j2, int j3) : is_fv(fv) {
  if (fv) {  // sorted triangle, then the vertex
    int t[3] = {j0, j1, j2};
    std::sort(t, t + 3);
    ino0 = t[0]; ino1 = t[1]; ino2 = t[2]; ino3 = j3;
  } else {   // the edge holding the smallest index first (smallest first), the other edge ascending
    int a0 = j0, a1 = j1, b0 = j2, b1 = j3;
    if (std::min(b0, b1) < std::min(a0, a1)) { std::swap(a0, b0); std::swap(a1, b1); }
    if (a1 < a0) std::swap(a0, a1);
    if (b1 < b0) std::swap(b0, b1);
    ino0 = a0; ino1 = a1; ino2 = b0; ino3 = b1;
  }
}
bool CContactElement::operator<(const CContactElement& p) const {
  if (ino0 != p.ino0) return ino0 < p.ino0;
  if (ino1 != p.ino1) return ino1 < p.ino1;
  if (ino2 != p.ino2) return ino2 < p.ino2;
  return ino3 < p.ino3;
}

// ------------------------------------------------------------------------------------------------ element level
void MakePositiveDefinite_Sim22(const double s2[3], double s3[3]) {
  check(clothgpu_elem_psd_sym22(1, s2, s3), "clothgpu_elem_psd_sym22");
}
void WdWddW_CST(double& W, double dW[3][3], double ddW[3][3][3][3], const double C[3][3], const double c[3][3],
                const double lambda, const double myu) {
  check(clothgpu_elem_cst(1, &C[0][0], &c[0][0], lambda, myu, &W, &dW[0][0], &ddW[0][0][0][0]), "clothgpu_elem_cst");
}
void WdWddW_Bend(double& W, double dW[4][3], double ddW[4][4][3][3], const double C[4][3], const double c[4][3],
                 double stiff) {
  double w = 0;
  check(clothgpu_elem_bend(1, &C[0][0], &c[0][0], stiff, &w, &dW[0][0], &ddW[0][0][0][0]), "clothgpu_elem_bend");
  W += w;  // cloth_internal_physics.h:235 accumulates
}
void WdWddW_Contact(double& W, double dW[3], double ddW[3][3], const double c[3], double stiff_contact,
                    double contact_clearance, void (*penetrationDepth)(double&, double*, const double*)) {
  double pdn[4] = {0, 0, 0, 0};
  penetrationDepth(pdn[0], pdn + 1, c);  // the caller's callback is host code
  check(clothgpu_elem_contact(1, pdn, stiff_contact, contact_clearance, &W, dW, &ddW[0][0]), "clothgpu_elem_contact");
}

namespace {
void pack4(double* p, const CVector3D& a, const CVector3D& b, const CVector3D& c, const CVector3D& d) {
  const CVector3D* v[4] = {&a, &b, &c, &d};
  for (int i = 0; i < 4; i++) { p[i * 3] = v[i]->x; p[i * 3 + 1] = v[i]->y; p[i * 3 + 2] = v[i]->z; }
}
}  // namespace
double DistanceFaceVertex(const CVector3D& p0, const CVector3D& p1, const CVector3D& p2, const CVector3D& p3,
                          double& w0, double& w1) {
  double p[12], d = 0;
  pack4(p, p0, p1, p2, p3);
  check(clothgpu_dist_face_vertex(1, p, &d, &w0, &w1), "clothgpu_dist_face_vertex");
  return d;
}
double DistanceEdgeEdge(const CVector3D& p0, const CVector3D& p1, const CVector3D& q0, const CVector3D& q1,
                        double& ratio_p, double& ratio_q) {
  double p[12], d = 0;
  pack4(p, p0, p1, q0, q1);
  check(clothgpu_dist_edge_edge(1, p, &d, &ratio_p, &ratio_q), "clothgpu_dist_edge_edge");
  return d;
}
double FindCoplanerInterp(const CVector3D& s0, const CVector3D& s1, const CVector3D& s2, const CVector3D& s3,
                          const CVector3D& e0, const CVector3D& e1, const CVector3D& e2, const CVector3D& e3) {
  double s[12], e[12], t = 0;
  pack4(s, s0, s1, s2, s3);
  pack4(e, e0, e1, e2, e3);
  check(clothgpu_coplanar_interp(1, s, e, &t), "clothgpu_coplanar_interp");
  return t;
}

// ------------------------------------------------------------------------------------------------ BVH
// MakeBVHTopology_TopDown (bvh_aabb.cpp:294-313): the reference's own node array (libclothgpu's host builder,
// csrc/bvhbuild.cu); where the reference crashes on exactly square grids it falls back to a longest-axis split
int MakeBVHTopology_TopDown(const std::vector<int>& aTri, const std::vector<double>& aXYZ,
                            std::vector<CNodeBVH>& aNodeBVH) {
  const int ntri = (int)aTri.size() / 3;
  aNodeBVH.clear();
  if (ntri == 0) return -1;
  std::vector<int> nodes((size_t)(2 * ntri - 1) * 3);
  const int n = clothgpu_bvh_build_topdown(ntri, aTri.data(), (int)aXYZ.size() / 3, aXYZ.data(), nodes.data(),
                                           2 * ntri - 1, nullptr);
  if (n < 0) check(-n, "clothgpu_bvh_build_topdown");
  aNodeBVH.resize(n);
  for (int i = 0; i < n; i++) {
    aNodeBVH[i].iroot = nodes[(size_t)i * 3];
    aNodeBVH[i].ichild[0] = nodes[(size_t)i * 3 + 1];
    aNodeBVH[i].ichild[1] = nodes[(size_t)i * 3 + 2];
  }
  return 0;
}

void BuildBoundingBoxChild_Prx(int, double delta, const std::vector<double>& aXYZ, const std::vector<int>& aTri,
                               const std::vector<CNodeBVH>& aNodeBVH, std::vector<CAABB3D>& aBB) {
  Scene& s = scene_for(aTri, (int)aXYZ.size() / 3);
  add_contact(s, aNodeBVH, nullptr);
  clothgpu_ctx* c = ensure_ctx(s);
  check(clothgpu_write(c, CLOTHGPU_BUF_XYZ, 0, aXYZ.data(), (long long)aXYZ.size()), "write xyz");
  check(clothgpu_bvh_refit_prx(c, CLOTHGPU_BUF_XYZ, delta), "clothgpu_bvh_refit_prx");
  boxes_to_vec(c, aNodeBVH.size(), aBB);
}
void BuildBoundingBoxChild_CCD(int, double dt, const std::vector<double>& aXYZ, const std::vector<double>& aUVW,
                               const std::vector<int>& aTri, const std::vector<CNodeBVH>& aNodeBVH,
                               std::vector<CAABB3D>& aBB) {
  Scene& s = scene_for(aTri, (int)aXYZ.size() / 3);
  add_contact(s, aNodeBVH, nullptr);
  clothgpu_ctx* c = ensure_ctx(s);
  check(clothgpu_upload_state(c, 0, aXYZ.data(), aUVW.data()), "upload state");
  check(clothgpu_bvh_refit_ccd(c, CLOTHGPU_BUF_XYZ, CLOTHGPU_BUF_UVW, dt), "clothgpu_bvh_refit_ccd");
  boxes_to_vec(c, aNodeBVH.size(), aBB);
}
void GetContactElement_Proximity(std::set<CContactElement>& out, double delta, const std::vector<double>& aXYZ,
                                 const std::vector<int>& aTri, int, const std::vector<CNodeBVH>& aBVH,
                                 const std::vector<CAABB3D>& aBB) {
  Scene& s = scene_for(aTri, (int)aXYZ.size() / 3);
  add_contact(s, aBVH, nullptr);
  clothgpu_ctx* c = ensure_ctx(s);
  check(clothgpu_write(c, CLOTHGPU_BUF_XYZ, 0, aXYZ.data(), (long long)aXYZ.size()), "write xyz");
  vec_to_boxes(c, aBB);
  int n = 0;
  check(clothgpu_contacts_prx(c, CLOTHGPU_BUF_XYZ, delta, &n), "clothgpu_contacts_prx");
  elems_to_set(c, n, out);
}
void GetContactElement_CCD(std::set<CContactElement>& out, double dt, double delta, const std::vector<double>& aXYZ,
                           const std::vector<double>& aUVW, const std::vector<int>& aTri, int,
                           const std::vector<CNodeBVH>& aBVH, const std::vector<CAABB3D>& aBB) {
  Scene& s = scene_for(aTri, (int)aXYZ.size() / 3);
  add_contact(s, aBVH, nullptr);
  clothgpu_ctx* c = ensure_ctx(s);
  check(clothgpu_upload_state(c, 0, aXYZ.data(), aUVW.data()), "upload state");
  vec_to_boxes(c, aBB);
  int n = 0;
  check(clothgpu_contacts_ccd(c, CLOTHGPU_BUF_XYZ, CLOTHGPU_BUF_UVW, dt, delta, &n), "clothgpu_contacts_ccd");
  elems_to_set(c, n, out);
}

// ------------------------------------------------------------------------------------------------ CMatrixSquareSparse
CMatrixSquareSparse::CMatrixSquareSparse(int nblk, int len)
    : m_nblk(0), m_len(0), m_ncrs(0), m_colInd(nullptr), m_rowPtr(nullptr), m_valCrs(nullptr), m_valDia(nullptr), m_dev(nullptr) {
  Initialize(nblk, len);
}
CMatrixSquareSparse::CMatrixSquareSparse()
    : m_nblk(0), m_len(0), m_ncrs(0), m_colInd(nullptr), m_rowPtr(nullptr), m_valCrs(nullptr), m_valDia(nullptr), m_dev(nullptr) {}
CMatrixSquareSparse::~CMatrixSquareSparse() {
  delete[] m_colInd; delete[] m_rowPtr; delete[] m_valCrs; delete[] m_valDia;
  if (m_dev) clothgpu_destroy(m_dev);
}
void CMatrixSquareSparse::Initialize(int nblk, int len) {
  delete[] m_colInd; delete[] m_rowPtr; delete[] m_valCrs; delete[] m_valDia;
  if (m_dev) { clothgpu_destroy(m_dev); m_dev = nullptr; }
  m_nblk = nblk; m_len = len; m_ncrs = 0;
  m_colInd = new int[nblk + 1];
  for (int i = 0; i <= nblk; i++) m_colInd[i] = 0;
  m_rowPtr = nullptr; m_valCrs = nullptr;
  m_valDia = new double[(size_t)nblk * len * len];
  for (size_t i = 0; i < (size_t)nblk * len * len; i++) m_valDia[i] = 0.0;
}
void CMatrixSquareSparse::operator=(const CMatrixSquareSparse& m) {
  Initialize(m.m_nblk, m.m_len);
  m_ncrs = m.m_ncrs;
  for (int i = 0; i <= m_nblk; i++) m_colInd[i] = m.m_colInd[i];
  m_rowPtr = new int[std::max(m_ncrs, 1)];
  for (int i = 0; i < m_ncrs; i++) m_rowPtr[i] = m.m_rowPtr[i];
  const size_t bs = (size_t)m_len * m_len;
  m_valCrs = new double[std::max<size_t>(m_ncrs * bs, 1)];
  if (m.m_valCrs) for (size_t i = 0; i < m_ncrs * bs; i++) m_valCrs[i] = m.m_valCrs[i];
  for (size_t i = 0; i < m_nblk * bs; i++) m_valDia[i] = m.m_valDia[i];
}
void CMatrixSquareSparse::SetPattern(const std::vector<int>& colind, const std::vector<int>& rowptr) {
  if ((int)colind.size() != m_nblk + 1) throw std::runtime_error("SetPattern: row pointer has the wrong length");
  if (m_dev) { clothgpu_destroy(m_dev); m_dev = nullptr; }
  m_ncrs = colind[m_nblk];
  delete[] m_rowPtr; delete[] m_valCrs;
  for (int i = 0; i <= m_nblk; i++) m_colInd[i] = colind[i];
  m_rowPtr = new int[std::max(m_ncrs, 1)];
  for (int i = 0; i < m_ncrs; i++) m_rowPtr[i] = rowptr[i];
  const size_t n = (size_t)m_ncrs * m_len * m_len;
  m_valCrs = new double[std::max<size_t>(n, 1)];
  for (size_t i = 0; i < n; i++) m_valCrs[i] = 0.0;
}
bool CMatrixSquareSparse::SetZero() {
  const size_t bs = (size_t)m_len * m_len;
  for (size_t i = 0; i < m_nblk * bs; i++) m_valDia[i] = 0.0;
  for (size_t i = 0; i < m_ncrs * bs; i++) m_valCrs[i] = 0.0;
  return true;
}
// host-side scatter-add of an element matrix into the mirrors (matrix_square_sparse.cpp:88-138): blocks outside the
// pattern are dropped.  The GPU step assembles on the device and never calls this; it exists for callers that build
// their own systems and then use MatVec / Solve_PCG.
bool CMatrixSquareSparse::Mearge(int nrow, const int* rows, int ncol, const int* cols, int blksize, const double* emat,
                                 std::vector<int>& tmp) {
  if (blksize != m_len * m_len) return false;
  if ((int)tmp.size() < m_nblk) tmp.assign(m_nblk, -1);
  for (int a = 0; a < nrow; a++) {
    const int i = rows[a];
    for (int j = m_colInd[i]; j < m_colInd[i + 1]; j++) tmp[m_rowPtr[j]] = j;
    for (int b = 0; b < ncol; b++) {
      const int jc = cols[b];
      const double* e = emat + (size_t)(a * ncol + b) * blksize;
      double* dst = nullptr;
      if (jc == i) dst = m_valDia + (size_t)i * blksize;
      else if (tmp[jc] >= 0) dst = m_valCrs + (size_t)tmp[jc] * blksize;
      if (!dst) continue;
      for (int k = 0; k < blksize; k++) dst[k] += e[k];
    }
    for (int j = m_colInd[i]; j < m_colInd[i + 1]; j++) tmp[m_rowPtr[j]] = -1;
  }
  return true;
}
void CMatrixSquareSparse::MatVec(double alpha, const std::vector<double>& x, double beta, std::vector<double>& y) const {
  clothgpu_ctx* c = matrix_ctx(*this);
  upload_matrix(c, *this, CLOTHGPU_BUF_VAL_DIA, CLOTHGPU_BUF_VAL_CRS);
  check(clothgpu_write(c, CLOTHGPU_BUF_VEC_P, 0, x.data(), (long long)x.size()), "write x");
  check(clothgpu_write(c, CLOTHGPU_BUF_VEC_Z, 0, y.data(), (long long)y.size()), "write y");
  check(clothgpu_matvec(c, alpha, CLOTHGPU_BUF_VEC_P, beta, CLOTHGPU_BUF_VEC_Z), "clothgpu_matvec");
  check(clothgpu_read(c, CLOTHGPU_BUF_VEC_Z, 0, y.data(), (long long)y.size()), "read y");
}
void CMatrixSquareSparse::SetBoundaryCondition(const std::vector<int>& bc) {  // matrix_square_sparse.cpp:272-302
  // one flag per block row: a fixed row becomes the identity, a fixed column is cleared
  const size_t bs = (size_t)m_len * m_len;
  for (int i = 0; i < m_nblk; i++) {
    if (!bc[i]) continue;
    for (size_t k = 0; k < bs; k++) m_valDia[i * bs + k] = 0.0;
    for (int d = 0; d < m_len; d++) m_valDia[i * bs + d * m_len + d] = 1.0;
    for (int j = m_colInd[i]; j < m_colInd[i + 1]; j++)
      for (size_t k = 0; k < bs; k++) m_valCrs[j * bs + k] = 0.0;
  }
  for (int j = 0; j < m_ncrs; j++) {
    if (!bc[m_rowPtr[j]]) continue;
    for (size_t k = 0; k < bs; k++) m_valCrs[j * bs + k] = 0.0;
  }
}

double InnerProduct(std::vector<double>& r_vec, std::vector<double>& u_vec) {
  // a pattern-free BLAS-1 call has no context to run in: use a throw-away diagonal matrix context of that length
  CMatrixSquareSparse m((int)r_vec.size() / 3, 3);
  std::vector<int> rp(m.m_nblk + 1, 0), ci;
  m.SetPattern(rp, ci);
  clothgpu_ctx* c = matrix_ctx(m);
  check(clothgpu_write(c, CLOTHGPU_BUF_VEC_P, 0, r_vec.data(), (long long)r_vec.size()), "write r");
  check(clothgpu_write(c, CLOTHGPU_BUF_VEC_Z, 0, u_vec.data(), (long long)u_vec.size()), "write u");
  double d = 0;
  check(clothgpu_dot(c, CLOTHGPU_BUF_VEC_P, CLOTHGPU_BUF_VEC_Z, &d), "clothgpu_dot");
  return d;
}
void AXPY(double a, const std::vector<double>& x, std::vector<double>& y) {
  CMatrixSquareSparse m((int)x.size() / 3, 3);
  std::vector<int> rp(m.m_nblk + 1, 0), ci;
  m.SetPattern(rp, ci);
  clothgpu_ctx* c = matrix_ctx(m);
  check(clothgpu_write(c, CLOTHGPU_BUF_VEC_P, 0, x.data(), (long long)x.size()), "write x");
  check(clothgpu_write(c, CLOTHGPU_BUF_VEC_Z, 0, y.data(), (long long)y.size()), "write y");
  check(clothgpu_axpy(c, a, CLOTHGPU_BUF_VEC_P, CLOTHGPU_BUF_VEC_Z), "clothgpu_axpy");
  check(clothgpu_read(c, CLOTHGPU_BUF_VEC_Z, 0, y.data(), (long long)y.size()), "read y");
}
void Solve_CG(double& conv_ratio, int& iteration, const CMatrixSquareSparse& mat, std::vector<double>& r_vec,
              std::vector<double>& u_vec) {
  clothgpu_ctx* c = matrix_ctx(mat);
  upload_matrix(c, mat, CLOTHGPU_BUF_VAL_DIA, CLOTHGPU_BUF_VAL_CRS);
  check(clothgpu_write(c, CLOTHGPU_BUF_VEC_R, 0, r_vec.data(), (long long)r_vec.size()), "write r");
  check(clothgpu_cg(c, &conv_ratio, &iteration), "clothgpu_cg");
  u_vec.resize(r_vec.size());
  check(clothgpu_read(c, CLOTHGPU_BUF_VEC_X, 0, u_vec.data(), (long long)u_vec.size()), "read x");
  check(clothgpu_read(c, CLOTHGPU_BUF_VEC_R, 0, r_vec.data(), (long long)r_vec.size()), "read r");
}

// ------------------------------------------------------------------------------------------------ CPreconditionerILU
CPreconditionerILU::CPreconditionerILU() : m_diaInd(nullptr) {}
CPreconditionerILU::~CPreconditionerILU() { delete[] m_diaInd; }
void CPreconditionerILU::Initialize_ILU0(const CMatrixSquareSparse& m) {  // ilu_sparse.cpp:114-134
  mat = m;
  delete[] m_diaInd;
  m_diaInd = new int[std::max(m.m_nblk, 1)];
  for (int i = 0; i < m.m_nblk; i++) {
    m_diaInd[i] = mat.m_colInd[i + 1];
    for (int j = mat.m_colInd[i]; j < mat.m_colInd[i + 1]; j++)
      if (mat.m_rowPtr[j] > i) { m_diaInd[i] = j; break; }
  }
}
void CPreconditionerILU::SetValueILU(const CMatrixSquareSparse& m) {  // ilu_sparse.cpp:352-384
  const size_t bs = (size_t)m.m_len * m.m_len;
  for (size_t i = 0; i < m.m_nblk * bs; i++) mat.m_valDia[i] = m.m_valDia[i];
  for (size_t i = 0; i < m.m_ncrs * bs; i++) mat.m_valCrs[i] = m.m_valCrs[i];
}
void CPreconditionerILU::DoILUDecomp() {
  clothgpu_ctx* c = matrix_ctx(mat);
  upload_matrix(c, mat, CLOTHGPU_BUF_ILU_DIA, CLOTHGPU_BUF_ILU_CRS);
  check(clothgpu_ilu_factor(c), "clothgpu_ilu_factor");
  check(clothgpu_read(c, CLOTHGPU_BUF_ILU_DIA, 0, mat.m_valDia, (long long)mat.m_nblk * 9), "read factor");
  if (mat.m_ncrs) check(clothgpu_read(c, CLOTHGPU_BUF_ILU_CRS, 0, mat.m_valCrs, (long long)mat.m_ncrs * 9), "read factor");
}
void CPreconditionerILU::Solve(std::vector<double>& vec) const {
  clothgpu_ctx* c = matrix_ctx(mat);
  upload_matrix(c, mat, CLOTHGPU_BUF_ILU_DIA, CLOTHGPU_BUF_ILU_CRS);
  check(clothgpu_write(c, CLOTHGPU_BUF_VEC_Z, 0, vec.data(), (long long)vec.size()), "write vec");
  check(clothgpu_ilu_solve(c, CLOTHGPU_BUF_VEC_Z), "clothgpu_ilu_solve");
  check(clothgpu_read(c, CLOTHGPU_BUF_VEC_Z, 0, vec.data(), (long long)vec.size()), "read vec");
}
void Solve_PCG(double& conv_ratio, int& iteration, const CMatrixSquareSparse& mat, const CPreconditionerILU& ilu,
               std::vector<double>& r_vec, std::vector<double>& u_vec) {
  clothgpu_ctx* c = matrix_ctx(mat);
  upload_matrix(c, mat, CLOTHGPU_BUF_VAL_DIA, CLOTHGPU_BUF_VAL_CRS);
  upload_matrix(c, ilu.mat, CLOTHGPU_BUF_ILU_DIA, CLOTHGPU_BUF_ILU_CRS);
  check(clothgpu_write(c, CLOTHGPU_BUF_VEC_R, 0, r_vec.data(), (long long)r_vec.size()), "write r");
  check(clothgpu_pcg(c, &conv_ratio, &iteration), "clothgpu_pcg");
  u_vec.resize(r_vec.size());
  check(clothgpu_read(c, CLOTHGPU_BUF_VEC_X, 0, u_vec.data(), (long long)u_vec.size()), "read x");
  check(clothgpu_read(c, CLOTHGPU_BUF_VEC_R, 0, r_vec.data(), (long long)r_vec.size()), "read r");
}

// ------------------------------------------------------------------------------------------------ step functions
static void step_internal(bool ilu, std::vector<double>& aXYZ, std::vector<double>& aUVW, CMatrixSquareSparse& mat_A,
                          const std::vector<double>& aXYZ0, const std::vector<int>& aBCFlag,
                          const std::vector<int>& aTri, const std::vector<int>& aQuad, double dt, double lambda,
                          double myu, double stiff_bend, const double gravity[3], double mass_point,
                          double stiff_contact, double contact_clearance,
                          void (*penetrationDepth)(double&, double*, const double*)) {
  Scene& s = scene_for(aTri, (int)aXYZ.size() / 3);
  add_internal(s, aXYZ0, aBCFlag, aQuad, &mat_A);
  clothgpu_ctx* c = ensure_ctx(s);
  clothgpu_params p = make_params(dt, lambda, myu, stiff_bend, gravity, mass_point, stiff_contact, contact_clearance);
  upload_collider(c, p, aXYZ, penetrationDepth);
  check(clothgpu_upload_state(c, 0, aXYZ.data(), aUVW.data()), "upload state");
  double conv = 0;
  int iter = 0;
  if (ilu) check(clothgpu_step_internal_ilu(c, &p, &conv, &iter), "clothgpu_step_internal_ilu");
  else check(clothgpu_step_internal_cg(c, &p, &conv, &iter), "clothgpu_step_internal_cg");
  check(clothgpu_download_state(c, 0, aXYZ.data(), aUVW.data()), "download state");
  g_last_iter = iter;
  g_last_conv = conv;
  if (g_verbose) printf("  conv_ratio:%g  iteration:%d\n", conv, iter);
}
void StepTime_InternalDynamics(std::vector<double>& aXYZ, std::vector<double>& aUVW, CMatrixSquareSparse& mat_A,
                               const std::vector<double>& aXYZ0, const std::vector<int>& aBCFlag,
                               const std::vector<int>& aTri, const std::vector<int>& aQuad, const double dt,
                               double lambda, double myu, double stiff_bend, const double gravity[3],
                               double mass_point, double stiff_contact, double contact_clearance,
                               void (*penetrationDepth)(double&, double*, const double*)) {
  step_internal(false, aXYZ, aUVW, mat_A, aXYZ0, aBCFlag, aTri, aQuad, dt, lambda, myu, stiff_bend, gravity, mass_point,
                stiff_contact, contact_clearance, penetrationDepth);
}
void StepTime_InternalDynamicsILU(std::vector<double>& aXYZ, std::vector<double>& aUVW, CMatrixSquareSparse& mat_A,
                                  CPreconditionerILU&, const std::vector<double>& aXYZ0,
                                  const std::vector<int>& aBCFlag, const std::vector<int>& aTri,
                                  const std::vector<int>& aQuad, const double dt, double lambda, double myu,
                                  double stiff_bend, const double gravity[3], double mass_point, double stiff_contact,
                                  double contact_clearance,
                                  void (*penetrationDepth)(double&, double*, const double*)) {
  step_internal(true, aXYZ, aUVW, mat_A, aXYZ0, aBCFlag, aTri, aQuad, dt, lambda, myu, stiff_bend, gravity, mass_point,
                stiff_contact, contact_clearance, penetrationDepth);
}
void UpdateIntermidiateVelocity(std::vector<double>& aUVW, const std::vector<double>& aXYZ,
                                const std::vector<double>& aXYZ0, const std::vector<int>& aBCFlag,
                                const std::vector<int>& aTri, const std::vector<int>& aQuad, const double dt,
                                double lambda, double myu, double stiff_bend, const double gravity[3],
                                double mass_point) {
  Scene& s = scene_for(aTri, (int)aXYZ.size() / 3);
  add_internal(s, aXYZ0, aBCFlag, aQuad, nullptr);
  clothgpu_ctx* c = ensure_ctx(s);
  clothgpu_params p = make_params(dt, lambda, myu, stiff_bend, gravity, mass_point, 0.0, 0.0);
  check(clothgpu_upload_state(c, 0, aXYZ.data(), aUVW.data()), "upload state");
  check(clothgpu_update_intermediate_velocity(c, &p), "clothgpu_update_intermediate_velocity");
  std::vector<double> x(aXYZ.size());
  check(clothgpu_download_state(c, 0, x.data(), aUVW.data()), "download state");
}
void GetIntermidiateVelocityContactResolved(std::vector<double>& aUVWm, bool& is_impulse_applied, double dt,
                                            double contact_clearance, double mass_point,
                                            double cloth_contact_stiffness, const std::vector<double>& aXYZ,
                                            const std::vector<int>& aTri, const CJaggedArray& aEdge, int,
                                            const std::vector<CNodeBVH>& aNodeBVH, std::vector<CAABB3D>& aBB) {
  Scene& s = scene_for(aTri, (int)aXYZ.size() / 3);
  add_contact(s, aNodeBVH, &aEdge);
  clothgpu_ctx* c = ensure_ctx(s);
  clothgpu_params p = make_params(dt, 0, 0, 0, nullptr, mass_point, cloth_contact_stiffness, contact_clearance);
  check(clothgpu_write(c, CLOTHGPU_BUF_XYZ_PREV, 0, aXYZ.data(), (long long)aXYZ.size()), "write positions");
  check(clothgpu_write(c, CLOTHGPU_BUF_UVW, 0, aUVWm.data(), (long long)aUVWm.size()), "write velocities");
  int applied = 0;
  std::vector<int> stats(CLOTHGPU_NSTATS, 0);
  check(clothgpu_resolve_self_contact(c, &p, &applied, stats.data()), "clothgpu_resolve_self_contact");
  check(clothgpu_read(c, CLOTHGPU_BUF_UVW, 0, aUVWm.data(), (long long)aUVWm.size()), "read velocities");
  boxes_to_vec(c, aNodeBVH.size(), aBB);
  is_impulse_applied = applied != 0;
  if (g_verbose) printf("  contact elem prox:%d  ccd passes:%d  riz rounds:%d\n", stats[1], stats[2], stats[3]);
}
