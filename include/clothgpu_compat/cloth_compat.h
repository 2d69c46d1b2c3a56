// include/clothgpu_compat/cloth_compat.h -- source-compatible C++ host layer over the C ABI (include/clothgpu.h).
//
// The reference (nobuyuki83/cloth_sim_self_collision) has no plugin interface: its boundary is the set of C++
// classes and free functions that StepTime() / main() of self_contact_sparse/main.cpp:213-263,561-583 and
// internal_cloth_sparse/main.cpp:205-237,531-550 call.  This header declares those names with the reference's
// signatures; their bodies (cloth_sim_self_collision_b200/compat/cloth_compat.cpp -> libclothcompat.so) forward to
// the extern "C" clothgpu_* entry points, which are the only symbols that cross into CUDA code.  The per-reference
// header names (matrix_square_sparse.h, ilu_sparse.h, solve_internal_sparse.h, self_collision_cloth.h, bvh_aabb.h,
// jagged_array.h, aabb.h, cloth_internal_physics.h) in this directory just include this file; utility.h and vector3d.h
// are small host-side value helpers of their own.  tests/cpp/Makefile builds the reference's UNCHANGED demo sources
// (self_contact_sparse/main.cpp, internal_cloth_sparse/main.cpp) against this directory -- see INTEGRATION.md.
//
// Semantics kept from the reference: the caller owns all std::vectors; every call takes its inputs from them and
// writes its outputs back (positions/velocities cross the PCIe bus once per call -- the "e2e" path of bench.py).
// Device contexts are created lazily per mesh and cached; there is NO CPU fallback: without a CUDA device every
// computing call throws std::runtime_error carrying clothgpu_last_error().
#ifndef CLOTHGPU_COMPAT_H
#define CLOTHGPU_COMPAT_H

#include <set>
#include <vector>

#include "utility.h"   // utility.h:15-123 (double[3] helpers, camera quaternions)
#include "vector3d.h"  // vector3d.h:33-295 (CVector3D)

struct clothgpu_ctx;

// ---- cloth_internal_physics.h:15-280: the element energies, evaluated on the GPU (clothgpu_elem_*; one element per
// call here, batches through the C ABI).  Same values as the fused assembly computes.
void MakePositiveDefinite_Sim22(const double s2[3], double s3[3]);
void WdWddW_CST(double& W, double dW[3][3], double ddW[3][3][3][3], const double C[3][3], const double c[3][3],
                const double lambda, const double myu);
// (the reference ADDS the bending energy to its W argument, cloth_internal_physics.h:235; kept)
void WdWddW_Bend(double& W, double dW[4][3], double ddW[4][4][3][3], const double C[4][3], const double c[4][3],
                 double stiff);
void WdWddW_Contact(double& W, double dW[3], double ddW[3][3], const double c[3], double stiff_contact,
                    double contact_clearance, void (*penetrationDepth)(double&, double*, const double*));

// ---- jagged_array.h:21-262 (setup side: adjacency lists; the hot path consumes index/array)
class CJaggedArray {
public:
  CJaggedArray() {}
  void InitializeSize(int size);
  int Size() const;
  // vertex -> vertex adjacency over elements (jagged_array.h:158-207); is_dia adds the vertex itself
  void SetEdgeOfElem(const std::vector<int>& aElem, int nelem, int nnoel, int nnode, bool is_dia);
  void Sort();  // ascending per row (jagged_array.h:71-90)
public:
  std::vector<int> index;
  std::vector<int> array;
};

// ---- aabb.h:16-148
class CAABB3D {
public:
  CAABB3D();
  CAABB3D(double x_min0, double x_max0, double y_min0, double y_max0, double z_min0, double z_max0);
  bool IsIntersect(const CAABB3D& bb) const;   // closed intervals
  CAABB3D& operator+=(const CAABB3D& bb);
  bool AddPoint(double x, double y, double z, double eps);
  bool IsInside(double x, double y, double z) const;
public:
  double x_min, x_max, y_min, y_max, z_min, z_max;
  bool isnt_empty;
};

// ---- bvh_aabb.h:18-23
class CNodeBVH {
public:
  int iroot;
  int ichild[2];
};

// ---- bvh_aabb.h:65-106: canonical 4-tuple, ordered lexicographically, is_fv not part of the order
class CContactElement {
public:
  CContactElement(bool is_fv, int j0, int j1, int j2, int j3);
  bool operator<(const CContactElement& p2) const;
public:
  bool is_fv;
  int ino0, ino1, ino2, ino3;
};

// ---- bvh_aabb.h:49-61: the geometric predicates behind the narrow phase and the impulses, evaluated on the GPU
// (clothgpu_dist_face_vertex / _edge_edge / clothgpu_coplanar_interp with n = 1)
double DistanceFaceVertex(const CVector3D& p0, const CVector3D& p1, const CVector3D& p2, const CVector3D& p3,
                          double& w0, double& w1);
double DistanceEdgeEdge(const CVector3D& p0, const CVector3D& p1, const CVector3D& q0, const CVector3D& q1,
                        double& ratio_p, double& ratio_q);
double FindCoplanerInterp(const CVector3D& s0, const CVector3D& s1, const CVector3D& s2, const CVector3D& s3,
                          const CVector3D& e0, const CVector3D& e1, const CVector3D& e2, const CVector3D& e3);

// Setup: the BVH topology over the triangles (bvh_aabb.h:25-28, bvh_aabb.cpp:294-313): node for node the array the
// reference builds (clothgpu_bvh_build_topdown); where the reference crashes on exactly square grids
// the split falls back to the longest axis.  Returns 0 (the root index the demos store in iroot_bvh).
int MakeBVHTopology_TopDown(const std::vector<int>& aTri, const std::vector<double>& aXYZ,
                            std::vector<CNodeBVH>& aNodeBVH);
void BuildBoundingBoxChild_Prx(int ibvh, double delta, const std::vector<double>& aXYZ, const std::vector<int>& aTri,
                               const std::vector<CNodeBVH>& aNodeBVH, std::vector<CAABB3D>& aBB);
void BuildBoundingBoxChild_CCD(int ibvh, double dt, const std::vector<double>& aXYZ, const std::vector<double>& aUVW,
                               const std::vector<int>& aTri, const std::vector<CNodeBVH>& aNodeBVH,
                               std::vector<CAABB3D>& aBB);
void GetContactElement_Proximity(std::set<CContactElement>& aContactElem, double delta,
                                 const std::vector<double>& aXYZ, const std::vector<int>& aTri, int ibvh,
                                 const std::vector<CNodeBVH>& aBVH, const std::vector<CAABB3D>& aBB);
void GetContactElement_CCD(std::set<CContactElement>& aContactElem, double dt, double delta,
                           const std::vector<double>& aXYZ, const std::vector<double>& aUVW,
                           const std::vector<int>& aTri, int ibvh, const std::vector<CNodeBVH>& aBVH,
                           const std::vector<CAABB3D>& aBB);

// ---- matrix_square_sparse.h:14-63.  m_colInd is the ROW POINTER and m_rowPtr the COLUMN INDEX, as in the reference.
class CMatrixSquareSparse {
public:
  CMatrixSquareSparse(int nblk, int len);
  CMatrixSquareSparse();
  virtual ~CMatrixSquareSparse();
  void Initialize(int nblk, int len);
  void operator=(const CMatrixSquareSparse& m);
  void SetPattern(const std::vector<int>& colind, const std::vector<int>& rowptr);
  bool SetZero();
  bool Mearge(int nblkel_col, const int* blkel_col, int nblkel_row, const int* blkel_row, int blksize,
              const double* emat, std::vector<int>& m_marge_tmp_buffer);
  void MatVec(double alpha, const std::vector<double>& x, double beta, std::vector<double>& y) const;  // on the GPU
  void SetBoundaryCondition(const std::vector<int>& bc_flag);
public:
  int m_nblk;
  int m_len;
  int m_ncrs;
  int* m_colInd;
  int* m_rowPtr;
  double* m_valCrs;
  double* m_valDia;
  // extension: the device context bound to this pattern (created on demand, shared with the ILU object)
  mutable clothgpu_ctx* m_dev;
};
double InnerProduct(std::vector<double>& r_vec, std::vector<double>& u_vec);
void AXPY(double a, const std::vector<double>& x, std::vector<double>& y);
void Solve_CG(double& conv_ratio, int& iteration, const CMatrixSquareSparse& mat, std::vector<double>& r_vec,
              std::vector<double>& u_vec);

// ---- ilu_sparse.h:16-43
class CPreconditionerILU {
public:
  CPreconditionerILU();
  ~CPreconditionerILU();
  void Initialize_ILU0(const CMatrixSquareSparse& m);
  void SetValueILU(const CMatrixSquareSparse& m);
  void Solve(std::vector<double>& vec) const;  // forward + backward substitution on the GPU
  void DoILUDecomp();                          // on the GPU; mat.m_val* receive the factor
public:
  CMatrixSquareSparse mat;
  int* m_diaInd;
};
void Solve_PCG(double& conv_ratio, int& iteration, const CMatrixSquareSparse& mat, const CPreconditionerILU& ilu,
               std::vector<double>& r_vec, std::vector<double>& u_vec);

// ---- solve_internal_sparse.h:185-369
void StepTime_InternalDynamics(std::vector<double>& aXYZ, std::vector<double>& aUVW, CMatrixSquareSparse& mat_A,
                               const std::vector<double>& aXYZ0, const std::vector<int>& aBCFlag,
                               const std::vector<int>& aTri, const std::vector<int>& aQuad, const double dt,
                               double lambda, double myu, double stiff_bend, const double gravity[3],
                               double mass_point, double stiff_contact, double contact_clearance,
                               void (*penetrationDepth)(double&, double*, const double*));
void StepTime_InternalDynamicsILU(std::vector<double>& aXYZ, std::vector<double>& aUVW, CMatrixSquareSparse& mat_A,
                                  CPreconditionerILU& ilu_A, const std::vector<double>& aXYZ0,
                                  const std::vector<int>& aBCFlag, const std::vector<int>& aTri,
                                  const std::vector<int>& aQuad, const double dt, double lambda, double myu,
                                  double stiff_bend, const double gravity[3], double mass_point, double stiff_contact,
                                  double contact_clearance,
                                  void (*penetrationDepth)(double&, double*, const double*));
void UpdateIntermidiateVelocity(std::vector<double>& aUVW, const std::vector<double>& aXYZ,
                                const std::vector<double>& aXYZ0, const std::vector<int>& aBCFlag,
                                const std::vector<int>& aTri, const std::vector<int>& aQuad, const double dt,
                                double lambda, double myu, double stiff_bend, const double gravity[3],
                                double mass_point);

// ---- self_collision_cloth.h:21-34
void GetIntermidiateVelocityContactResolved(std::vector<double>& aUVWm, bool& is_impulse_applied, double dt,
                                            double contact_clearance, double mass_point,
                                            double cloth_contact_stiffness, const std::vector<double>& aXYZ,
                                            const std::vector<int>& aTri, const CJaggedArray& aEdge, int iroot_bvh,
                                            const std::vector<CNodeBVH>& aNodeBVH, std::vector<CAABB3D>& aBB);

// ---- extensions (not in the reference)
namespace clothgpu_compat {
// the reference prints "energy", "conv_ratio/iteration" and contact counts to stdout; off by default here
void SetVerbose(bool on);
// statistics of the last StepTime_InternalDynamics* / GetIntermidiateVelocityContactResolved call
int LastIteration();
double LastConvRatio();
// releases every cached device context
void ReleaseAll();
}  // namespace clothgpu_compat

#endif  // CLOTHGPU_COMPAT_H
