// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Flat extern "C" wrappers around the UNMODIFIED reference library, compiled from the sources
// where they lie under /root/reference (see oracle/Makefile; nothing is copied into this repo).
// The resulting oracle/_ref/libclothref.so is (1) the authority the C restatement in
// oracle/cloth_oracle.c is validated against, (2) the generator of tests/golden/*.npz and
// (3) the "reference" CPU baseline timed by bench.py.
//
// Only the demo-level glue that lives in the GLUT main.cpp files (colliders, StepTime
// orchestration; self_contact_sparse/main.cpp:110-133,213-263, internal_cloth_sparse/main.cpp:107-132)
// is restated here, because those translation units need <GL/glut.h>, which is absent.
#include <cstdlib>
#include <cstring>
#include <new>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#include <iostream>

#include "utility.h"
#include "solve_internal_sparse.h"   // includes matrix_square_sparse.h, ilu_sparse.h, cloth_internal_physics.h
#include "self_collision_cloth.h"    // jagged_array.h, aabb.h, bvh_aabb.h

// non-static reference functions that are not declared in any reference header
void SelfCollisionImpulse_Proximity(std::vector<double>& aUVWm, double delta, double stiffness, double dt,
                                    double mass, const std::vector<double>& aXYZ, const std::vector<int>& aTri,
                                    const std::vector<CContactElement>& aContactElem);
void SelfCollisionImpulse_CCD(std::vector<double>& aUVWm, double delta, double stiffness, double dt,
                              double mass, const std::vector<double>& aXYZ, const std::vector<int>& aTri,
                              const std::vector<CContactElement>& aContactElem);
void MakeRigidImpactZone(std::vector<std::set<int> >& aRIZ, const std::vector<CContactElement>& aContactElem,
                         const CJaggedArray& aEdge);
void ApplyRigidImpactZone(std::vector<double>& aUVWm, const std::vector<std::set<int> >& aRIZ,
                          const std::vector<double>& aXYZ, const std::vector<double>& aUVWm0);
bool IsContact_FV_Proximity(int, int, int, int, const CVector3D&, const CVector3D&, const CVector3D&,
                            const CVector3D&, const CAABB3D&, const double);
bool IsContact_EE_Proximity(int, int, int, int, const CVector3D&, const CVector3D&, const CVector3D&,
                            const CVector3D&, const double);
bool IsContact_FV_CCD(int, int, int, int, const CVector3D&, const CVector3D&, const CVector3D&, const CVector3D&,
                      const CVector3D&, const CVector3D&, const CVector3D&, const CVector3D&, const CAABB3D&);
bool IsContact_EE_CCD(int, int, int, int, const CVector3D&, const CVector3D&, const CVector3D&, const CVector3D&,
                      const CVector3D&, const CVector3D&, const CVector3D&, const CVector3D&);

namespace {

// ---- colliders (restated demo glue). kind: 0 none, 1 plane z=c0, 2 sphere obstacle, 3 sphere container
int g_col_kind = 0;
double g_col[4] = {0, 0, 0, 0};

void col_none(double& pd, double*, const double*) { pd = -100; }  // internal_cloth_sparse/main.cpp:107-110
void col_plane(double& pd, double* n, const double* p) {           // self_contact_sparse/main.cpp:110-114
  n[0] = 0.0; n[1] = 0.0; n[2] = 1.0;
  pd = g_col[0] - p[2];
}
void col_sphere(double& pd, double* n, const double* p) {          // internal_cloth_sparse/main.cpp:120-132
  n[0] = p[0] - g_col[0]; n[1] = p[1] - g_col[1]; n[2] = p[2] - g_col[2];
  double len = Length3D(n);
  n[0] /= len; n[1] /= len; n[2] /= len;
  pd = g_col[3] - len;
}
void col_sphere_in(double& pd, double* n, const double* p) {       // self_contact_sparse/main.cpp:117-133
  col_sphere(pd, n, p);
  n[0] *= -1; n[1] *= -1; n[2] *= -1;
  pd *= -1;
}
typedef void (*col_fn)(double&, double*, const double*);
col_fn pick_collider(int kind, const double* par) {
  g_col_kind = kind;
  for (int i = 0; i < 4; i++) g_col[i] = par ? par[i] : 0.0;
  switch (kind) {
    case 1: return col_plane;
    case 2: return col_sphere;
    case 3: return col_sphere_in;
    default: return col_none;
  }
}

// The reference library reports only through std::cout; capture it so that the counts it prints
// (PCG iterations, contact elements per pass) become machine-readable and tests stay quiet.
struct CoutCapture {
  std::streambuf* old;
  std::ostringstream ss;
  CoutCapture() { old = std::cout.rdbuf(ss.rdbuf()); }
  ~CoutCapture() { std::cout.rdbuf(old); }
};

CPreconditionerILU* new_ilu() {  // ctor leaves m_diaInd uninitialised (ilu_sparse.cpp:102-105): pre-zero the storage
  void* mem = std::calloc(1, sizeof(CPreconditionerILU));
  return new (mem) CPreconditionerILU();
}
void delete_ilu(CPreconditionerILU* p) {
  p->~CPreconditionerILU();
  std::free(p);
}

void set_matrix(CMatrixSquareSparse& m, int np, const int* rowptr, const int* colidx) {
  m.Initialize(np, 3);
  std::vector<int> ci(rowptr, rowptr + np + 1), rp(colidx, colidx + rowptr[np]);
  m.SetPattern(ci, rp);
  m.SetZero();
}

std::vector<CNodeBVH> to_nodes(int nnode, const int* nodes) {
  std::vector<CNodeBVH> v(nnode);
  for (int i = 0; i < nnode; i++) {
    v[i].iroot = nodes[i * 3 + 0];
    v[i].ichild[0] = nodes[i * 3 + 1];
    v[i].ichild[1] = nodes[i * 3 + 2];
  }
  return v;
}
std::vector<CAABB3D> to_boxes(int nnode, const double* boxes) {
  std::vector<CAABB3D> v(nnode);
  for (int i = 0; i < nnode; i++) {
    v[i].x_min = boxes[i * 6 + 0]; v[i].x_max = boxes[i * 6 + 1];
    v[i].y_min = boxes[i * 6 + 2]; v[i].y_max = boxes[i * 6 + 3];
    v[i].z_min = boxes[i * 6 + 4]; v[i].z_max = boxes[i * 6 + 5];
    v[i].isnt_empty = true;
  }
  return v;
}
void from_boxes(const std::vector<CAABB3D>& v, double* boxes) {
  for (size_t i = 0; i < v.size(); i++) {
    boxes[i * 6 + 0] = v[i].x_min; boxes[i * 6 + 1] = v[i].x_max;
    boxes[i * 6 + 2] = v[i].y_min; boxes[i * 6 + 3] = v[i].y_max;
    boxes[i * 6 + 4] = v[i].z_min; boxes[i * 6 + 5] = v[i].z_max;
  }
}
std::vector<CContactElement> to_elems(int n, const int* e) {
  std::vector<CContactElement> v;
  v.reserve(n);
  for (int i = 0; i < n; i++) {
    // elements arrive canonical already; build through the ctor with a layout it leaves unchanged
    CContactElement ce(true, 0, 1, 2, 3);
    ce.is_fv = e[i * 5 + 4] != 0;
    ce.ino0 = e[i * 5 + 0]; ce.ino1 = e[i * 5 + 1]; ce.ino2 = e[i * 5 + 2]; ce.ino3 = e[i * 5 + 3];
    v.push_back(ce);
  }
  return v;
}
int from_set(const std::set<CContactElement>& s, int* out, int cap) {
  int n = 0;
  for (std::set<CContactElement>::const_iterator it = s.begin(); it != s.end(); ++it, ++n) {
    if (n < cap) {
      out[n * 5 + 0] = it->ino0; out[n * 5 + 1] = it->ino1; out[n * 5 + 2] = it->ino2; out[n * 5 + 3] = it->ino3;
      out[n * 5 + 4] = it->is_fv ? 1 : 0;
    }
  }
  return n;
}

// parse "<tag><int>" occurrences out of captured stdout
void parse_ints(const std::string& s, const char* tag, std::vector<int>& out) {
  size_t pos = 0, tl = std::strlen(tag);
  while ((pos = s.find(tag, pos)) != std::string::npos) {
    pos += tl;
    out.push_back(std::atoi(s.c_str() + pos));
  }
}

// collect the leaf pairs the reference traversal reaches, in its DFS order (bvh_aabb.cpp:441-557 control flow only)
void leaf_pairs_pair(std::vector<int>& out, int i0, int i1, const std::vector<CNodeBVH>& n, const std::vector<CAABB3D>& bb) {
  if (!bb[i0].IsIntersect(bb[i1])) return;
  const bool l0 = n[i0].ichild[1] == -1, l1 = n[i1].ichild[1] == -1;
  if (!l0 && !l1) {
    leaf_pairs_pair(out, n[i0].ichild[0], n[i1].ichild[0], n, bb);
    leaf_pairs_pair(out, n[i0].ichild[1], n[i1].ichild[0], n, bb);
    leaf_pairs_pair(out, n[i0].ichild[0], n[i1].ichild[1], n, bb);
    leaf_pairs_pair(out, n[i0].ichild[1], n[i1].ichild[1], n, bb);
  } else if (!l0 && l1) {
    leaf_pairs_pair(out, n[i0].ichild[0], i1, n, bb);
    leaf_pairs_pair(out, n[i0].ichild[1], i1, n, bb);
  } else if (l0 && !l1) {
    leaf_pairs_pair(out, i0, n[i1].ichild[0], n, bb);
    leaf_pairs_pair(out, i0, n[i1].ichild[1], n, bb);
  } else {
    out.push_back(i0);
    out.push_back(i1);
  }
}
void leaf_pairs_self(std::vector<int>& out, int i, const std::vector<CNodeBVH>& n, const std::vector<CAABB3D>& bb) {
  if (n[i].ichild[1] == -1) return;
  leaf_pairs_pair(out, n[i].ichild[0], n[i].ichild[1], n, bb);
  leaf_pairs_self(out, n[i].ichild[0], n, bb);
  leaf_pairs_self(out, n[i].ichild[1], n, bb);
}

}  // namespace

extern "C" {

// ------------------------------------------------------------------ topology / setup
// MakeBVHTopology_TopDown (bvh_aabb.cpp:294-313). nodes: (2*ntri-1) x {iroot, ichild0, ichild1}. Returns node count.
int ref_bvh_build(int ntri, const int* tri, int np, const double* xyz, int* nodes, int cap) {
  std::vector<int> aTri(tri, tri + ntri * 3);
  std::vector<double> aXYZ(xyz, xyz + np * 3);
  std::vector<CNodeBVH> aNode;
  MakeBVHTopology_TopDown(aTri, aXYZ, aNode);
  int n = (int)aNode.size();
  for (int i = 0; i < n && i < cap; i++) {
    nodes[i * 3 + 0] = aNode[i].iroot;
    nodes[i * 3 + 1] = aNode[i].ichild[0];
    nodes[i * 3 + 2] = aNode[i].ichild[1];
  }
  return n;
}

// CJaggedArray::SetEdgeOfElem (+Sort) (jagged_array.h:158-207, 71-90). Returns nnz; fills index[nnode+1], array[<=cap].
int ref_edge_of_elem(const int* elem, int nelem, int nnoel, int nnode, int do_sort, int* index, int* array, int cap) {
  std::vector<int> e(elem, elem + nelem * nnoel);
  CJaggedArray ja;
  ja.SetEdgeOfElem(e, nelem, nnoel, nnode, false);
  if (do_sort) ja.Sort();
  for (int i = 0; i <= nnode; i++) index[i] = ja.index[i];
  int nnz = ja.index[nnode];
  for (int i = 0; i < nnz && i < cap; i++) array[i] = ja.array[i];
  return nnz;
}

// Initialize_ILU0's m_diaInd (ilu_sparse.cpp:114-134)
void ref_ilu_diaind(int np, const int* rowptr, const int* colidx, int* diaind) {
  CMatrixSquareSparse m;
  set_matrix(m, np, rowptr, colidx);
  CPreconditionerILU* ilu = new_ilu();
  ilu->Initialize_ILU0(m);
  for (int i = 0; i < np; i++) diaind[i] = ilu->m_diaInd[i];
  delete_ilu(ilu);
}

// ------------------------------------------------------------------ element energies
void ref_cst(const double* C, const double* c, double lambda, double myu, double* W, double* dW, double* ddW) {
  double Cm[3][3], cm[3][3], dw[3][3], ddw[3][3][3][3];
  std::memcpy(Cm, C, sizeof(Cm)); std::memcpy(cm, c, sizeof(cm));
  WdWddW_CST(*W, dw, ddw, Cm, cm, lambda, myu);
  std::memcpy(dW, dw, sizeof(dw)); std::memcpy(ddW, ddw, sizeof(ddw));
}
void ref_bend(const double* C, const double* c, double stiff, double* W, double* dW, double* ddW) {
  double Cm[4][3], cm[4][3], dw[4][3], ddw[4][4][3][3];
  std::memcpy(Cm, C, sizeof(Cm)); std::memcpy(cm, c, sizeof(cm));
  *W = 0;  // the reference accumulates into the caller's variable (cloth_internal_physics.h:235)
  WdWddW_Bend(*W, dw, ddw, Cm, cm, stiff);
  std::memcpy(dW, dw, sizeof(dw)); std::memcpy(ddW, ddw, sizeof(ddw));
}
void ref_contact(const double* c, double stiff, double clearance, int col_kind, const double* col_par,
                 double* W, double* dW, double* ddW) {
  double ddw[3][3];
  WdWddW_Contact(*W, dW, ddw, c, stiff, clearance, pick_collider(col_kind, col_par));
  std::memcpy(ddW, ddw, sizeof(ddw));
}

// ------------------------------------------------------------------ implicit step pieces
// Left/right-hand side exactly as StepTime_InternalDynamicsILU builds them (solve_internal_sparse.h:281-313).
void ref_assemble(int np, int ntri, int nquad, const int* tri, const int* quad, const int* bc,
                  const double* xyz0, const double* xyz, const double* uvw,
                  const int* rowptr, const int* colidx,
                  double dt, double lambda, double myu, double stiff_bend, const double* gravity,
                  double mass_point, double stiff_contact, double clearance, int col_kind, const double* col_par,
                  double* valDia, double* valCrs, double* vec_b) {
  CoutCapture cap;
  std::vector<double> aXYZ(xyz, xyz + np * 3), aXYZ0(xyz0, xyz0 + np * 3), aUVW(uvw, uvw + np * 3);
  std::vector<int> aTri(tri, tri + ntri * 3), aQuad(quad, quad + nquad * 4), aBC(bc, bc + np);
  CMatrixSquareSparse mat_A;
  set_matrix(mat_A, np, rowptr, colidx);
  const int nDof = np * 3;
  double W = 0;
  std::vector<double> b(nDof, 0);
  mat_A.SetZero();
  std::vector<int> tmp_buffer(np, -1);
  AddWdWddW_Cloth(W, b, mat_A, tmp_buffer, aXYZ, aXYZ0, aTri, aQuad, lambda, myu, stiff_bend);
  AddWdWddW_Contact(W, b, mat_A, tmp_buffer, aXYZ, stiff_contact, clearance, pick_collider(col_kind, col_par));
  AddWdW_Gravity(W, b, aXYZ, gravity, mass_point);
  for (int i = 0; i < nDof; i++) { b[i] = -b[i] + mass_point * aUVW[i] / dt; }
  for (int ip = 0; ip < np; ip++) {
    mat_A.m_valDia[ip * 9 + 0 * 3 + 0] += mass_point / (dt * dt);
    mat_A.m_valDia[ip * 9 + 1 * 3 + 1] += mass_point / (dt * dt);
    mat_A.m_valDia[ip * 9 + 2 * 3 + 2] += mass_point / (dt * dt);
  }
  mat_A.SetBoundaryCondition(aBC);
  for (int ip = 0; ip < np; ip++) {
    if (aBC[ip] == 0) continue;
    b[ip * 3 + 0] = 0; b[ip * 3 + 1] = 0; b[ip * 3 + 2] = 0;
  }
  std::memcpy(valDia, mat_A.m_valDia, sizeof(double) * np * 9);
  std::memcpy(valCrs, mat_A.m_valCrs, sizeof(double) * rowptr[np] * 9);
  std::memcpy(vec_b, b.data(), sizeof(double) * nDof);
}

void ref_matvec(int np, const int* rowptr, const int* colidx, const double* valDia, const double* valCrs,
                double alpha, const double* x, double beta, double* y) {
  CMatrixSquareSparse m;
  set_matrix(m, np, rowptr, colidx);
  std::memcpy(m.m_valDia, valDia, sizeof(double) * np * 9);
  std::memcpy(m.m_valCrs, valCrs, sizeof(double) * rowptr[np] * 9);
  std::vector<double> vx(x, x + np * 3), vy(y, y + np * 3);
  m.MatVec(alpha, vx, beta, vy);
  std::memcpy(y, vy.data(), sizeof(double) * np * 3);
}

double ref_dot(int n, const double* a, const double* b) {
  std::vector<double> va(a, a + n), vb(b, b + n);
  return InnerProduct(va, vb);
}

// SetValueILU + DoILUDecomp (ilu_sparse.cpp:352-384, 387-624); values are factored in place
void ref_ilu_factor(int np, const int* rowptr, const int* colidx, double* valDia, double* valCrs) {
  CoutCapture cap;
  CMatrixSquareSparse m;
  set_matrix(m, np, rowptr, colidx);
  std::memcpy(m.m_valDia, valDia, sizeof(double) * np * 9);
  std::memcpy(m.m_valCrs, valCrs, sizeof(double) * rowptr[np] * 9);
  CPreconditionerILU* ilu = new_ilu();
  ilu->Initialize_ILU0(m);
  ilu->SetValueILU(m);
  ilu->DoILUDecomp();
  std::memcpy(valDia, ilu->mat.m_valDia, sizeof(double) * np * 9);
  std::memcpy(valCrs, ilu->mat.m_valCrs, sizeof(double) * rowptr[np] * 9);
  delete_ilu(ilu);
}

namespace {
CPreconditionerILU* ilu_from_factored(int np, const int* rowptr, const int* colidx, const double* dia, const double* crs) {
  CMatrixSquareSparse m;
  set_matrix(m, np, rowptr, colidx);
  CPreconditionerILU* ilu = new_ilu();
  ilu->Initialize_ILU0(m);
  std::memcpy(ilu->mat.m_valDia, dia, sizeof(double) * np * 9);
  std::memcpy(ilu->mat.m_valCrs, crs, sizeof(double) * rowptr[np] * 9);
  return ilu;
}
}  // namespace

// CPreconditionerILU::Solve on already factored values (ilu_sparse.h:23-26)
void ref_ilu_solve(int np, const int* rowptr, const int* colidx, const double* iluDia, const double* iluCrs, double* vec) {
  CPreconditionerILU* ilu = ilu_from_factored(np, rowptr, colidx, iluDia, iluCrs);
  std::vector<double> v(vec, vec + np * 3);
  ilu->Solve(v);
  std::memcpy(vec, v.data(), sizeof(double) * np * 3);
  delete_ilu(ilu);
}

// Solve_PCG (ilu_sparse.cpp:628-709). conv_ratio/iteration are in/out exactly as in the reference.
void ref_pcg(int np, const int* rowptr, const int* colidx, const double* valDia, const double* valCrs,
             const double* iluDia, const double* iluCrs, double* r, double* x, double* conv_ratio, int* iteration) {
  CMatrixSquareSparse m;
  set_matrix(m, np, rowptr, colidx);
  std::memcpy(m.m_valDia, valDia, sizeof(double) * np * 9);
  std::memcpy(m.m_valCrs, valCrs, sizeof(double) * rowptr[np] * 9);
  CPreconditionerILU* ilu = ilu_from_factored(np, rowptr, colidx, iluDia, iluCrs);
  std::vector<double> vr(r, r + np * 3), vx;
  Solve_PCG(*conv_ratio, *iteration, m, *ilu, vr, vx);
  std::memcpy(r, vr.data(), sizeof(double) * np * 3);
  std::memcpy(x, vx.data(), sizeof(double) * np * 3);
  delete_ilu(ilu);
}

// Solve_CG (matrix_square_sparse.cpp:338-406)
void ref_cg(int np, const int* rowptr, const int* colidx, const double* valDia, const double* valCrs,
            double* r, double* x, double* conv_ratio, int* iteration) {
  CMatrixSquareSparse m;
  set_matrix(m, np, rowptr, colidx);
  std::memcpy(m.m_valDia, valDia, sizeof(double) * np * 9);
  std::memcpy(m.m_valCrs, valCrs, sizeof(double) * rowptr[np] * 9);
  std::vector<double> vr(r, r + np * 3), vx;
  Solve_CG(*conv_ratio, *iteration, m, vr, vx);
  std::memcpy(r, vr.data(), sizeof(double) * np * 3);
  std::memcpy(x, vx.data(), sizeof(double) * np * 3);
}

// UpdateIntermidiateVelocity (solve_internal_sparse.h:329-369)
void ref_update_intermediate_velocity(int np, int ntri, int nquad, const int* tri, const int* quad, const int* bc,
                                      const double* xyz0, const double* xyz, double* uvw, double dt, double lambda,
                                      double myu, double stiff_bend, const double* gravity, double mass_point) {
  std::vector<double> aXYZ(xyz, xyz + np * 3), aXYZ0(xyz0, xyz0 + np * 3), aUVW(uvw, uvw + np * 3);
  std::vector<int> aTri(tri, tri + ntri * 3), aQuad(quad, quad + nquad * 4), aBC(bc, bc + np);
  UpdateIntermidiateVelocity(aUVW, aXYZ, aXYZ0, aBC, aTri, aQuad, dt, lambda, myu, stiff_bend, gravity, mass_point);
  std::memcpy(uvw, aUVW.data(), sizeof(double) * np * 3);
}

// ------------------------------------------------------------------ BVH / contact detection
void ref_refit_prx(int nnode, const int* nodes, int ntri, const int* tri, int np, const double* xyz, double delta,
                   double* boxes) {
  std::vector<CNodeBVH> n = to_nodes(nnode, nodes);
  std::vector<int> aTri(tri, tri + ntri * 3);
  std::vector<double> aXYZ(xyz, xyz + np * 3);
  std::vector<CAABB3D> bb;
  BuildBoundingBoxChild_Prx(0, delta, aXYZ, aTri, n, bb);
  from_boxes(bb, boxes);
}
void ref_refit_ccd(int nnode, const int* nodes, int ntri, const int* tri, int np, const double* xyz,
                   const double* uvw, double dt, double* boxes) {
  std::vector<CNodeBVH> n = to_nodes(nnode, nodes);
  std::vector<int> aTri(tri, tri + ntri * 3);
  std::vector<double> aXYZ(xyz, xyz + np * 3), aUVW(uvw, uvw + np * 3);
  std::vector<CAABB3D> bb(nnode);
  BuildBoundingBoxChild_CCD(0, dt, aXYZ, aUVW, aTri, n, bb);
  from_boxes(bb, boxes);
}
// leaf pairs (BVH node ids) reached by the dual-tree self traversal, in DFS order. Returns the pair count.
int ref_leaf_pairs(int nnode, const int* nodes, const double* boxes, int* pairs, int cap) {
  std::vector<CNodeBVH> n = to_nodes(nnode, nodes);
  std::vector<CAABB3D> bb = to_boxes(nnode, boxes);
  std::vector<int> out;
  leaf_pairs_self(out, 0, n, bb);
  int np2 = (int)out.size() / 2;
  for (int i = 0; i < np2 && i < cap; i++) { pairs[i * 2] = out[i * 2]; pairs[i * 2 + 1] = out[i * 2 + 1]; }
  return np2;
}
// GetContactElement_Proximity on given boxes (bvh_aabb.cpp:540-557). elems: n x {ino0..3,is_fv} in std::set order.
int ref_contacts_prx(int nnode, const int* nodes, const double* boxes, int ntri, const int* tri, int np,
                     const double* xyz, double delta, int* elems, int cap) {
  std::vector<CNodeBVH> n = to_nodes(nnode, nodes);
  std::vector<CAABB3D> bb = to_boxes(nnode, boxes);
  std::vector<int> aTri(tri, tri + ntri * 3);
  std::vector<double> aXYZ(xyz, xyz + np * 3);
  std::set<CContactElement> s;
  GetContactElement_Proximity(s, delta, aXYZ, aTri, 0, n, bb);
  return from_set(s, elems, cap);
}
int ref_contacts_ccd(int nnode, const int* nodes, const double* boxes, int ntri, const int* tri, int np,
                     const double* xyz, const double* uvw, double dt, double delta, int* elems, int cap) {
  std::vector<CNodeBVH> n = to_nodes(nnode, nodes);
  std::vector<CAABB3D> bb = to_boxes(nnode, boxes);
  std::vector<int> aTri(tri, tri + ntri * 3);
  std::vector<double> aXYZ(xyz, xyz + np * 3), aUVW(uvw, uvw + np * 3);
  std::set<CContactElement> s;
  GetContactElement_CCD(s, dt, delta, aXYZ, aUVW, aTri, 0, n, bb);
  return from_set(s, elems, cap);
}

// geometric predicates, exposed for unit-level parity of the restatement and the device functions
double ref_dist_fv(const double* p, double* w0, double* w1) {
  return DistanceFaceVertex(CVector3D(p[0], p[1], p[2]), CVector3D(p[3], p[4], p[5]), CVector3D(p[6], p[7], p[8]),
                            CVector3D(p[9], p[10], p[11]), *w0, *w1);
}
double ref_dist_ee(const double* p, double* rp, double* rq) {
  return DistanceEdgeEdge(CVector3D(p[0], p[1], p[2]), CVector3D(p[3], p[4], p[5]), CVector3D(p[6], p[7], p[8]),
                          CVector3D(p[9], p[10], p[11]), *rp, *rq);
}
double ref_coplanar(const double* s, const double* e) {
  return FindCoplanerInterp(CVector3D(s[0], s[1], s[2]), CVector3D(s[3], s[4], s[5]), CVector3D(s[6], s[7], s[8]),
                            CVector3D(s[9], s[10], s[11]), CVector3D(e[0], e[1], e[2]), CVector3D(e[3], e[4], e[5]),
                            CVector3D(e[6], e[7], e[8]), CVector3D(e[9], e[10], e[11]));
}
int ref_is_fv_prx(const int* id, const double* p, const double* box, double delta) {
  CAABB3D bb(box[0], box[1], box[2], box[3], box[4], box[5]);
  return IsContact_FV_Proximity(id[0], id[1], id[2], id[3], CVector3D(p[0], p[1], p[2]), CVector3D(p[3], p[4], p[5]),
                                CVector3D(p[6], p[7], p[8]), CVector3D(p[9], p[10], p[11]), bb, delta);
}
int ref_is_ee_prx(const int* id, const double* p, double delta) {
  return IsContact_EE_Proximity(id[0], id[1], id[2], id[3], CVector3D(p[0], p[1], p[2]), CVector3D(p[3], p[4], p[5]),
                                CVector3D(p[6], p[7], p[8]), CVector3D(p[9], p[10], p[11]), delta);
}
int ref_is_fv_ccd(const int* id, const double* p, const double* q, const double* box) {
  CAABB3D bb(box[0], box[1], box[2], box[3], box[4], box[5]);
  return IsContact_FV_CCD(id[0], id[1], id[2], id[3], CVector3D(p[0], p[1], p[2]), CVector3D(p[3], p[4], p[5]),
                          CVector3D(p[6], p[7], p[8]), CVector3D(p[9], p[10], p[11]), CVector3D(q[0], q[1], q[2]),
                          CVector3D(q[3], q[4], q[5]), CVector3D(q[6], q[7], q[8]), CVector3D(q[9], q[10], q[11]), bb);
}
int ref_is_ee_ccd(const int* id, const double* p, const double* q) {
  return IsContact_EE_CCD(id[0], id[1], id[2], id[3], CVector3D(p[0], p[1], p[2]), CVector3D(p[3], p[4], p[5]),
                          CVector3D(p[6], p[7], p[8]), CVector3D(p[9], p[10], p[11]), CVector3D(q[0], q[1], q[2]),
                          CVector3D(q[3], q[4], q[5]), CVector3D(q[6], q[7], q[8]), CVector3D(q[9], q[10], q[11]));
}

// ------------------------------------------------------------------ contact response
void ref_impulse_prx(int np, double* uvw, double delta, double stiffness, double dt, double mass, const double* xyz,
                     int nelem, const int* elems) {
  std::vector<double> v(uvw, uvw + np * 3), x(xyz, xyz + np * 3);
  std::vector<int> aTri;
  SelfCollisionImpulse_Proximity(v, delta, stiffness, dt, mass, x, aTri, to_elems(nelem, elems));
  std::memcpy(uvw, v.data(), sizeof(double) * np * 3);
}
void ref_impulse_ccd(int np, double* uvw, double delta, double stiffness, double dt, double mass, const double* xyz,
                     int nelem, const int* elems) {
  std::vector<double> v(uvw, uvw + np * 3), x(xyz, xyz + np * 3);
  std::vector<int> aTri;
  SelfCollisionImpulse_CCD(v, delta, stiffness, dt, mass, x, aTri, to_elems(nelem, elems));
  std::memcpy(uvw, v.data(), sizeof(double) * np * 3);
}

// rigid impact zones: the zone list persists across rounds of one step (self_collision_cloth.cpp:427-455)
struct RefRIZ { std::vector<std::set<int> > zones; };
void* ref_riz_new() { return new RefRIZ(); }
void ref_riz_free(void* h) { delete (RefRIZ*)h; }
void ref_riz_make(void* h, int nelem, const int* elems, int np, const int* edge_index, const int* edge_array) {
  CJaggedArray aEdge;
  aEdge.index.assign(edge_index, edge_index + np + 1);
  aEdge.array.assign(edge_array, edge_array + edge_index[np]);
  MakeRigidImpactZone(((RefRIZ*)h)->zones, to_elems(nelem, elems), aEdge);
}
// zone_of_node[ip] = index of the zone holding ip, or -1. Returns the number of zones.
int ref_riz_get(void* h, int np, int* zone_of_node) {
  RefRIZ* r = (RefRIZ*)h;
  for (int i = 0; i < np; i++) zone_of_node[i] = -1;
  for (size_t z = 0; z < r->zones.size(); z++)
    for (std::set<int>::iterator it = r->zones[z].begin(); it != r->zones[z].end(); ++it) zone_of_node[*it] = (int)z;
  return (int)r->zones.size();
}
void ref_riz_apply(void* h, int np, double* uvw, const double* xyz, const double* uvw0) {
  std::vector<double> v(uvw, uvw + np * 3), x(xyz, xyz + np * 3), v0(uvw0, uvw0 + np * 3);
  ApplyRigidImpactZone(v, ((RefRIZ*)h)->zones, x, v0);
  std::memcpy(uvw, v.data(), sizeof(double) * np * 3);
}

// ------------------------------------------------------------------ whole-scene stepping (persistent handle)
struct RefSim {
  int np, ntri, nquad, nnode;
  std::vector<double> aXYZ0, aXYZ, aUVW;
  std::vector<int> aBC, aTri, aQuad;
  CMatrixSquareSparse mat_A;
  CPreconditionerILU* ilu_A;
  std::vector<CNodeBVH> aNodeBVH;
  std::vector<CAABB3D> aBB;
  CJaggedArray aEdge;
};

// init sequence of self_contact_sparse/main.cpp:561-583 with the BVH topology passed in (nnode may be 0)
void* ref_sim_new(int np, int ntri, int nquad, const int* tri, const int* quad, const int* bc, const double* xyz0,
                  int nnode, const int* nodes) {
  RefSim* s = new RefSim();
  s->np = np; s->ntri = ntri; s->nquad = nquad; s->nnode = nnode;
  s->aXYZ0.assign(xyz0, xyz0 + np * 3);
  s->aXYZ = s->aXYZ0;
  s->aUVW.assign(np * 3, 0.0);
  s->aBC.assign(bc, bc + np);
  s->aTri.assign(tri, tri + ntri * 3);
  s->aQuad.assign(quad, quad + nquad * 4);
  if (nnode > 0) s->aNodeBVH = to_nodes(nnode, nodes);
  s->aEdge.SetEdgeOfElem(s->aTri, ntri, 3, np, false);
  s->mat_A.Initialize(np, 3);
  CJaggedArray crs;
  crs.SetEdgeOfElem(s->aQuad, nquad, 4, np, false);
  crs.Sort();
  s->mat_A.SetPattern(crs.index, crs.array);
  s->ilu_A = new_ilu();
  s->ilu_A->Initialize_ILU0(s->mat_A);
  return s;
}
void ref_sim_free(void* h) {
  RefSim* s = (RefSim*)h;
  delete_ilu(s->ilu_A);
  delete s;
}
void ref_sim_set_state(void* h, const double* xyz, const double* uvw) {
  RefSim* s = (RefSim*)h;
  s->aXYZ.assign(xyz, xyz + s->np * 3);
  s->aUVW.assign(uvw, uvw + s->np * 3);
}
void ref_sim_get_state(void* h, double* xyz, double* uvw) {
  RefSim* s = (RefSim*)h;
  std::memcpy(xyz, s->aXYZ.data(), sizeof(double) * s->np * 3);
  std::memcpy(uvw, s->aUVW.data(), sizeof(double) * s->np * 3);
}

// StepTime_InternalDynamicsILU (solve_internal_sparse.h:255-326) == StepTime of internal_cloth_sparse/main.cpp:205-237
void ref_sim_step_internal(void* h, double dt, double lambda, double myu, double stiff_bend, const double* gravity,
                           double mass_point, double stiff_contact, double clearance, int col_kind,
                           const double* col_par, double* conv_ratio, int* iteration) {
  RefSim* s = (RefSim*)h;
  CoutCapture cap;
  StepTime_InternalDynamicsILU(s->aXYZ, s->aUVW, s->mat_A, *s->ilu_A, s->aXYZ0, s->aBC, s->aTri, s->aQuad, dt, lambda,
                               myu, stiff_bend, gravity, mass_point, stiff_contact, clearance,
                               pick_collider(col_kind, col_par));
  std::string out = cap.ss.str();
  std::vector<int> it;
  parse_ints(out, "iteration:", it);
  if (iteration) *iteration = it.empty() ? -1 : it[0];
  if (conv_ratio) {
    size_t p = out.find("conv_ratio:");
    *conv_ratio = p == std::string::npos ? -1.0 : std::atof(out.c_str() + p + 11);
  }
}

// StepTime_InternalDynamics (solve_internal_sparse.h:185-253): the unpreconditioned CG variant
void ref_sim_step_internal_cg(void* h, double dt, double lambda, double myu, double stiff_bend, const double* gravity,
                              double mass_point, double stiff_contact, double clearance, int col_kind,
                              const double* col_par, double* conv_ratio, int* iteration) {
  RefSim* s = (RefSim*)h;
  CoutCapture cap;
  StepTime_InternalDynamics(s->aXYZ, s->aUVW, s->mat_A, s->aXYZ0, s->aBC, s->aTri, s->aQuad, dt, lambda, myu, stiff_bend,
                            gravity, mass_point, stiff_contact, clearance, pick_collider(col_kind, col_par));
  std::string out = cap.ss.str();
  std::vector<int> it;
  parse_ints(out, "iteration:", it);
  if (iteration) *iteration = it.empty() ? -1 : it[0];
  if (conv_ratio) {
    size_t p = out.find("conv_ratio:");
    *conv_ratio = p == std::string::npos ? -1.0 : std::atof(out.c_str() + p + 11);
  }
}

// stats layout (ints): [0]=PCG iteration, [1]=#proximity elems, [2]=#CCD impulse passes run, [3]=#RIZ rounds run,
// [4]=is_impulse_applied, [5]=last "NNode In RIZ", [8..8+5)=CCD elems per impulse pass, [16..16+100)=CCD elems per RIZ round
enum { REF_NSTATS = 128 };

// GetIntermidiateVelocityContactResolved on the sim's BVH (self_collision_cloth.cpp:355-456)
int ref_sim_resolve_contacts(void* h, double* uvw, const double* xyz_prev, double dt, double clearance,
                             double mass_point, double stiffness, int* stats) {
  RefSim* s = (RefSim*)h;
  CoutCapture cap;
  std::vector<double> v(uvw, uvw + s->np * 3), x(xyz_prev, xyz_prev + s->np * 3);
  bool applied = false;
  GetIntermidiateVelocityContactResolved(v, applied, dt, clearance, mass_point, stiffness, x, s->aTri, s->aEdge, 0,
                                         s->aNodeBVH, s->aBB);
  std::memcpy(uvw, v.data(), sizeof(double) * s->np * 3);
  if (stats) {
    std::string out = cap.ss.str();
    for (int i = 0; i < REF_NSTATS; i++) stats[i] = 0;
    std::vector<int> prx, ccd, riz, nn;
    parse_ints(out, "Proximity      Contact Elem Size: ", prx);
    stats[1] = prx.empty() ? 0 : prx[0];
    // "  CCD iter: k    Contact Elem Size: n"
    size_t pos = 0;
    while ((pos = out.find("CCD iter: ", pos)) != std::string::npos) {
      size_t q = out.find("Contact Elem Size: ", pos);
      ccd.push_back(std::atoi(out.c_str() + q + 19));
      pos = q;
    }
    pos = 0;
    while ((pos = out.find("RIZ iter: ", pos)) != std::string::npos) {
      size_t q = out.find("Contact Elem Size: ", pos);
      riz.push_back(std::atoi(out.c_str() + q + 19));
      size_t r = out.find("NNode In RIZ: ", q);
      nn.push_back(std::atoi(out.c_str() + r + 14));
      pos = r;
    }
    stats[2] = (int)ccd.size();
    stats[3] = (int)riz.size();
    stats[4] = applied ? 1 : 0;
    stats[5] = nn.empty() ? 0 : nn.back();
    for (size_t i = 0; i < ccd.size() && i < 5; i++) stats[8 + i] = ccd[i];
    for (size_t i = 0; i < riz.size() && i < 100; i++) stats[16 + i] = riz[i];
  }
  return applied ? 1 : 0;
}

// StepTime of self_contact_sparse/main.cpp:213-263 (MakeNormal excluded: visualisation only)
void ref_sim_step_self_contact(void* h, double dt, double lambda, double myu, double stiff_bend, const double* gravity,
                               double mass_point, double stiff_contact, double clearance, int col_kind,
                               const double* col_par, int* stats) {
  RefSim* s = (RefSim*)h;
  std::vector<double> aXYZ1 = s->aXYZ;
  int iters = -1;
  ref_sim_step_internal(h, dt, lambda, myu, stiff_bend, gravity, mass_point, stiff_contact, clearance, col_kind,
                        col_par, 0, &iters);
  int applied = ref_sim_resolve_contacts(h, s->aUVW.data(), aXYZ1.data(), dt, clearance, mass_point, stiff_contact, stats);
  if (stats) stats[0] = iters;
  if (applied) {
    for (int ip = 0; ip < s->np; ip++) {
      s->aXYZ[ip * 3 + 0] = aXYZ1[ip * 3 + 0] + s->aUVW[ip * 3 + 0] * dt;
      s->aXYZ[ip * 3 + 1] = aXYZ1[ip * 3 + 1] + s->aUVW[ip * 3 + 1] * dt;
      s->aXYZ[ip * 3 + 2] = aXYZ1[ip * 3 + 2] + s->aUVW[ip * 3 + 2] * dt;
    }
    UpdateIntermidiateVelocity(s->aUVW, s->aXYZ, s->aXYZ0, s->aBC, s->aTri, s->aQuad, dt, lambda, myu, stiff_bend,
                               gravity, mass_point);
  }
}

}  // extern "C"
