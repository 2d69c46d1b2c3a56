// tests/cpp/self_contact_headless.cpp -- the caller side of the drop-in boundary, written the way the reference's
// demo is: the init sequence of self_contact_sparse/main.cpp:561-583 and its StepTime() (:213-263, without the GL
// drawing and MakeNormal), compiled against include/clothgpu_compat/ instead of the reference headers.
//
//   self_contact_headless <ndiv_x> <ndiv_z> <nsteps> <out.bin> [--check-link]
// writes: int np, double checksum, np*3 positions, np*3 velocities.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "solve_internal_sparse.h"
#include "self_collision_cloth.h"

// scene constants of self_contact_sparse/main.cpp:138-156
static const double elem_length = 0.1;
static const double lambda = 0.0, myu = 30.0, stiff_bend = 1.0e-2, areal_density = 0.04;
static const double gravity[3] = {0, 0.05, -10};
static const double time_step_size = 0.01, stiff_contact = 300.0, contact_clearance = 0.01;

static std::vector<double> aXYZ0, aXYZ, aUVW;
static std::vector<int> aBCFlag, aTri, aQuad;
static std::vector<CNodeBVH> aNodeBVH;
static std::vector<CAABB3D> aBB_BVH;
static CJaggedArray aEdge;
static CMatrixSquareSparse mat_A;
static CPreconditionerILU ilu_A;
static int iroot_bvh = 0;
static double mass_point = 0;

// collider of the default scene: the plane z = -0.5 (self_contact_sparse/main.cpp:110-114)
static void penetrationDepth_Plane(double& pd, double* n, const double* p) {
  n[0] = 0.0; n[1] = 0.0; n[2] = 1.0;
  pd = -0.5 - p[2];
}

// a cloth of (ndiv_x+1) x (ndiv_z+1) vertices in the plane y = 0, vertex id = ix*(ndiv_z+1)+iz, two triangles per
// cell, one hinge per interior edge (the layout SetClothShape_Square produces, main.cpp:31-106)
static void make_cloth(int ndiv_x, int ndiv_z, double& total_area) {
  const int nz1 = ndiv_z + 1, np = (ndiv_x + 1) * nz1;
  aXYZ0.assign((size_t)np * 3, 0.0);
  for (int ix = 0; ix <= ndiv_x; ix++)
    for (int iz = 0; iz < nz1; iz++) {
      aXYZ0[(size_t)(ix * nz1 + iz) * 3 + 0] = ix * elem_length;
      aXYZ0[(size_t)(ix * nz1 + iz) * 3 + 2] = iz * elem_length;
    }
  aBCFlag.assign(np, 0);
  aTri.clear(); aQuad.clear();
  for (int ix = 0; ix < ndiv_x; ix++)
    for (int iz = 0; iz < ndiv_z; iz++) {
      const int v00 = ix * nz1 + iz, v10 = (ix + 1) * nz1 + iz, v01 = ix * nz1 + iz + 1, v11 = (ix + 1) * nz1 + iz + 1;
      const int t[6] = {v00, v10, v01, v11, v01, v10};
      aTri.insert(aTri.end(), t, t + 6);
    }
  for (int ix = 0; ix < ndiv_x; ix++)
    for (int iz = 0; iz < ndiv_z; iz++) {
      const int q[4] = {ix * nz1 + iz, (ix + 1) * nz1 + iz + 1, (ix + 1) * nz1 + iz, ix * nz1 + iz + 1};
      aQuad.insert(aQuad.end(), q, q + 4);
    }
  for (int ix = 0; ix < ndiv_x; ix++)
    for (int iz = 0; iz < ndiv_z - 1; iz++) {
      const int q[4] = {(ix + 1) * nz1 + iz, ix * nz1 + iz + 2, (ix + 1) * nz1 + iz + 1, ix * nz1 + iz + 1};
      aQuad.insert(aQuad.end(), q, q + 4);
    }
  for (int ix = 0; ix < ndiv_x - 1; ix++)
    for (int iz = 0; iz < ndiv_z; iz++) {
      const int q[4] = {ix * nz1 + iz + 1, (ix + 2) * nz1 + iz, (ix + 1) * nz1 + iz, (ix + 1) * nz1 + iz + 1};
      aQuad.insert(aQuad.end(), q, q + 4);
    }
  total_area = ndiv_x * elem_length * ndiv_z * elem_length;
}

static void StepTime() {
  void (*penetrationDepth)(double&, double*, const double*) = penetrationDepth_Plane;
  std::vector<double> aXYZ1 = aXYZ;
  ::StepTime_InternalDynamicsILU(aXYZ, aUVW, mat_A, ilu_A, aXYZ0, aBCFlag, aTri, aQuad, time_step_size, lambda, myu,
                                 stiff_bend, gravity, mass_point, stiff_contact, contact_clearance, penetrationDepth);
  bool is_impulse_applied;
  GetIntermidiateVelocityContactResolved(aUVW, is_impulse_applied, time_step_size, contact_clearance, mass_point,
                                         stiff_contact, aXYZ1, aTri, aEdge, iroot_bvh, aNodeBVH, aBB_BVH);
  if (is_impulse_applied) {
    for (unsigned int ip = 0; ip < aXYZ.size() / 3; ip++) {
      aXYZ[ip * 3 + 0] = aXYZ1[ip * 3 + 0] + aUVW[ip * 3 + 0] * time_step_size;
      aXYZ[ip * 3 + 1] = aXYZ1[ip * 3 + 1] + aUVW[ip * 3 + 1] * time_step_size;
      aXYZ[ip * 3 + 2] = aXYZ1[ip * 3 + 2] + aUVW[ip * 3 + 2] * time_step_size;
    }
    ::UpdateIntermidiateVelocity(aUVW, aXYZ, aXYZ0, aBCFlag, aTri, aQuad, time_step_size, lambda, myu, stiff_bend,
                                 gravity, mass_point);
  }
}

int main(int argc, char* argv[]) {
  if (argc < 5) {
    fprintf(stderr, "usage: %s ndiv_x ndiv_z nsteps out.bin [--check-link] [--bvh nodes.bin]\n", argv[0]);
    return 2;
  }
  const int ndiv_x = atoi(argv[1]), ndiv_z = atoi(argv[2]), nsteps = atoi(argv[3]);
  bool check_link = false;
  const char* bvh_file = nullptr;
  for (int i = 5; i < argc; i++) {
    if (!strcmp(argv[i], "--check-link")) check_link = true;
    if (!strcmp(argv[i], "--bvh") && i + 1 < argc) bvh_file = argv[++i];
  }
  {  // initialise data: main.cpp:561-583
    double total_area;
    make_cloth(ndiv_x, ndiv_z, total_area);
    const int np = (int)aXYZ0.size() / 3;
    mass_point = total_area * areal_density / (double)np;
    aXYZ = aXYZ0;
    aUVW.assign(np * 3, 0.0);
    iroot_bvh = MakeBVHTopology_TopDown(aTri, aXYZ, aNodeBVH);
    if (bvh_file) {  // a BVH topology supplied by the test (the one the golden trajectory was recorded with)
      FILE* f = fopen(bvh_file, "rb");
      if (!f || fread(aNodeBVH.data(), sizeof(CNodeBVH), aNodeBVH.size(), f) != aNodeBVH.size()) { fprintf(stderr, "bad bvh file\n"); return 2; }
      fclose(f);
      iroot_bvh = 0;
    }
    aEdge.SetEdgeOfElem(aTri, (int)aTri.size() / 3, 3, np, false);
    mat_A.Initialize(np, 3);
    CJaggedArray crs;
    crs.SetEdgeOfElem(aQuad, (int)aQuad.size() / 4, 4, np, false);
    crs.Sort();
    mat_A.SetPattern(crs.index, crs.array);
    ilu_A.Initialize_ILU0(mat_A);
  }
  if (check_link) {
    printf("link ok: np=%d ntri=%d nquad=%d ncrs=%d nbvh=%d\n", (int)aXYZ0.size() / 3, (int)aTri.size() / 3,
           (int)aQuad.size() / 4, mat_A.m_ncrs, (int)aNodeBVH.size());
    return 0;
  }
  try {
    for (int i = 0; i < nsteps; i++) StepTime();
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  const int np = (int)aXYZ.size() / 3;
  double chk = 0;
  for (size_t i = 0; i < aXYZ.size(); i++) chk += aXYZ[i] * (1.0 + (double)(i % 7));
  FILE* f = fopen(argv[4], "wb");
  if (!f) return 2;
  fwrite(&np, sizeof(int), 1, f);
  fwrite(&chk, sizeof(double), 1, f);
  fwrite(aXYZ.data(), sizeof(double), aXYZ.size(), f);
  fwrite(aUVW.data(), sizeof(double), aUVW.size(), f);
  fclose(f);
  printf("steps=%d checksum=%.15g last PCG iteration=%d\n", nsteps, chk, clothgpu_compat::LastIteration());
  clothgpu_compat::ReleaseAll();
  return 0;
}
