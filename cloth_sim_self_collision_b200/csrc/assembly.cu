// csrc/assembly.cu -- element energies and the deterministic gather assembly.
//
// Replaces AddWdWddW_Cloth / AddWdW_Cloth / AddWdWddW_Contact / AddWdW_Gravity and the right-hand-side, mass and
// boundary-condition code of StepTime_InternalDynamicsILU (solve_internal_sparse.h:19-183, 281-313) and
// UpdateIntermidiateVelocity (:329-369).
//
// Design: the reference scatters element matrices into the block-CRS matrix one element after the other
// (CMatrixSquareSparse::Mearge, matrix_square_sparse.cpp:88-138).  Here two element kernels evaluate all
// triangles / hinges in parallel into element buffers, and one gather kernel per destination (vertex, diagonal
// block, off-diagonal slot) sums its contributions in the reference's element order from lists precomputed at
// context creation: bit-identical to the serial scatter, no atomics, every matrix entry written exactly once
// (which also replaces SetZero, the mass term and SetBoundaryCondition).
#include "ctx.h"
#include "dev_math.cuh"
#include "elem_math.cuh"

namespace cg {

struct AsmParams {
  double dt, lambda, myu, stiff_bend, mass, stiff_contact, clearance;
  int col_kind;
  double col_par[4];
};

// WdWddW_CST, cloth_internal_physics.h:70-176, one thread per (scene, triangle).
// tri_dW: [scene][tri][3][3]; tri_ddW: [scene][tri][ino*3+jno][idim*3+jdim] (nullptr: gradient only)
#ifndef TRI_B
#define TRI_B 3
#endif
__global__ void __launch_bounds__(128, TRI_B) k_tri_elements(int ntri, int np, const int* __restrict__ tri,
                                                       const double* __restrict__ xyz0, const double* __restrict__ xyz,
                                                       double lambda, double myu, double* __restrict__ tri_dW,
                                                       double* __restrict__ tri_ddW) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntri) return;
  const int s = blockIdx.y;
  const double* x = xyz + (size_t)s * np * 3;
  double C[9], c[9];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const int n = tri[t * 3 + a];
#pragma unroll
    for (int i = 0; i < 3; i++) { C[a * 3 + i] = xyz0[n * 3 + i]; c[a * 3 + i] = x[n * 3 + i]; }
  }
  cst_element(C, c, lambda, myu, nullptr, tri_dW + ((size_t)s * ntri + t) * 9,
              tri_ddW ? tri_ddW + ((size_t)s * ntri + t) * 81 : nullptr);
}

// rest-shape part of WdWddW_Bend, cloth_internal_physics.h:180-217: cotangent weights K[4] and
// tmp0 = stiff/((A0+A1) L0^2).  Shared by all scenes; recomputed only when stiff_bend changes.
__global__ void k_quad_rest(int nquad, const int* __restrict__ quad, const double* __restrict__ xyz0, double stiff,
                            double* __restrict__ quadK, double* __restrict__ quadT) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nquad) return;
  double C[12];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int n = quad[q * 4 + a];
#pragma unroll
    for (int i = 0; i < 3; i++) C[a * 3 + i] = xyz0[n * 3 + i];
  }
  const double A0 = tri_area(C + 0, C + 6, C + 9);
  const double A1 = tri_area(C + 3, C + 9, C + 6);
  const double L0 = dist3(C + 6, C + 9);
  const double H0 = A0 * 2.0 / L0;
  const double H1 = A1 * 2.0 / L0;
  double e23[3], e02[3], e03[3], e12[3], e13[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    e23[i] = C[9 + i] - C[6 + i];
    e02[i] = C[6 + i] - C[0 + i];
    e03[i] = C[9 + i] - C[0 + i];
    e12[i] = C[6 + i] - C[3 + i];
    e13[i] = C[9 + i] - C[3 + i];
  }
  double cot023, cot032, cot123, cot132;
  { const double r2 = -dot3(e02, e23), r3 = +dot3(e03, e23); cot023 = r2 / H0; cot032 = r3 / H0; }
  { const double r2 = -dot3(e12, e23), r3 = +dot3(e13, e23); cot123 = r2 / H1; cot132 = r3 / H1; }
  quadT[q] = stiff / ((A0 + A1) * L0 * L0);
  quadK[q * 4 + 0] = -cot023 - cot032;
  quadK[q * 4 + 1] = -cot123 - cot132;
  quadK[q * 4 + 2] = cot032 + cot132;
  quadK[q * 4 + 3] = cot023 + cot123;
}

// gradient of WdWddW_Bend, cloth_internal_physics.h:218-237: dW[ino] = sum_jno (K_ino K_jno tmp0) c[jno]; the
// Hessian blocks are (K_ino K_jno tmp0) I3, so the zero off-diagonal entries drop out of the reference's 12-term sum
__global__ void k_quad_elements(int nquad, int np, const int* __restrict__ quad, const double* __restrict__ xyz,
                                const double* __restrict__ quadK, const double* __restrict__ quadT,
                                double* __restrict__ quad_dW) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nquad) return;
  const int s = blockIdx.y;
  const double* x = xyz + (size_t)s * np * 3;
  double c[12], K[4];
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int n = quad[q * 4 + a];
    K[a] = quadK[q * 4 + a];
#pragma unroll
    for (int i = 0; i < 3; i++) c[a * 3 + i] = x[n * 3 + i];
  }
  const double tmp0 = quadT[q];
  double* dW = quad_dW + ((size_t)s * nquad + q) * 12;
#pragma unroll
  for (int ino = 0; ino < 4; ino++) {
#pragma unroll
    for (int idim = 0; idim < 3; idim++) {
      double g = 0;
#pragma unroll
      for (int jno = 0; jno < 4; jno++) g += K[ino] * K[jno] * tmp0 * c[jno * 3 + idim];
      dW[ino * 3 + idim] = g;
    }
  }
}

// penetration depth and normal of the built-in colliders (host callbacks in the reference demos)
__device__ __forceinline__ void collider_eval(int kind, const double* par, const double* p, const double* table,
                                              double& pd, double* n) {
  if (kind == CLOTHGPU_COLLIDER_PLANE_Z) {  // self_contact_sparse/main.cpp:110-114
    n[0] = 0.0; n[1] = 0.0; n[2] = 1.0;
    pd = par[0] - p[2];
  } else if (kind == CLOTHGPU_COLLIDER_SPHERE || kind == CLOTHGPU_COLLIDER_SPHERE_INSIDE) {
    // internal_cloth_sparse/main.cpp:120-132, self_contact_sparse/main.cpp:117-133
    n[0] = p[0] - par[0]; n[1] = p[1] - par[1]; n[2] = p[2] - par[2];
    const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    n[0] /= len; n[1] /= len; n[2] /= len;
    pd = par[3] - len;
    if (kind == CLOTHGPU_COLLIDER_SPHERE_INSIDE) {
      n[0] *= -1; n[1] *= -1; n[2] *= -1;
      pd *= -1;
    }
  } else if (kind == CLOTHGPU_COLLIDER_TABLE) {
    pd = table[0]; n[0] = table[1]; n[1] = table[2]; n[2] = table[3];
  } else {  // internal_cloth_sparse/main.cpp:107-110
    n[0] = n[1] = n[2] = 0.0;
    pd = -100;
  }
}

// per vertex: right-hand side and diagonal block.  Order of the sums = the reference's: triangles, hinges,
// contact (AddWdWddW_Contact, :137-161), gravity (:165-183), then b = -dW + m v/dt (:299-301), mass on the diagonal
// (:302-306), boundary condition (:307-313, matrix_square_sparse.cpp:272-302)
__global__ void __launch_bounds__(128) k_gather_vertex(int np, int ntri, int nquad, AsmParams P,
                                                        const int* __restrict__ bc, const int* __restrict__ vt_ptr,
                                                        const int* __restrict__ vt_ent, const int* __restrict__ vq_ptr,
                                                        const int* __restrict__ vq_ent, const double* __restrict__ xyz,
                                                        const double* __restrict__ uvw, const double* __restrict__ gravity,
                                                        const double* __restrict__ col_table,
                                                        const double* __restrict__ tri_dW, const double* __restrict__ tri_ddW,
                                                        const double* __restrict__ quad_dW, const double* __restrict__ quadK,
                                                        const double* __restrict__ quadT, double* __restrict__ valDia,
                                                        double* __restrict__ vec_b) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const int s = blockIdx.y;
  const double* tdW = tri_dW + (size_t)s * ntri * 9;
  const double* tddW = tri_ddW + (size_t)s * ntri * 81;
  const double* qdW = quad_dW + (size_t)s * nquad * 12;
  double b[3] = {0, 0, 0};
  double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = vt_ptr[i]; k < vt_ptr[i + 1]; k++) {
    const int e = vt_ent[k];  // tri*3 + corner
    const int t = e / 3, a = e - t * 3;
    const double* g = tdW + (size_t)e * 3;
    b[0] += g[0]; b[1] += g[1]; b[2] += g[2];
    const double* h = tddW + ((size_t)t * 9 + a * 3 + a) * 9;
#pragma unroll
    for (int m = 0; m < 9; m++) A[m] += h[m];
  }
  for (int k = vq_ptr[i]; k < vq_ptr[i + 1]; k++) {
    const int e = vq_ent[k];  // quad*4 + corner
    const double* g = qdW + (size_t)e * 3;
    b[0] += g[0]; b[1] += g[1]; b[2] += g[2];
    const double kk = quadK[e] * quadK[e] * quadT[e >> 2];
    A[0] += kk; A[4] += kk; A[8] += kk;
  }
  {  // WdWddW_Contact, cloth_internal_physics.h:243-280
    const double* p = xyz + ((size_t)s * np + i) * 3;
    double pd, n[3];
    collider_eval(P.col_kind, P.col_par, p, col_table + ((size_t)s * np + i) * 4, pd, n);
    pd += P.clearance;
    if (!(pd < 0)) {
      b[0] += -P.stiff_contact * pd * n[0];
      b[1] += -P.stiff_contact * pd * n[1];
      b[2] += -P.stiff_contact * pd * n[2];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int q = 0; q < 3; q++) A[r * 3 + q] += P.stiff_contact * n[r] * n[q];
    }
  }
  const double* gv = gravity + s * 3;
  b[0] -= P.mass * gv[0]; b[1] -= P.mass * gv[1]; b[2] -= P.mass * gv[2];
  const double* v = uvw + ((size_t)s * np + i) * 3;
  b[0] = -b[0] + P.mass * v[0] / P.dt;
  b[1] = -b[1] + P.mass * v[1] / P.dt;
  b[2] = -b[2] + P.mass * v[2] / P.dt;
  A[0] += P.mass / (P.dt * P.dt);
  A[4] += P.mass / (P.dt * P.dt);
  A[8] += P.mass / (P.dt * P.dt);
  if (bc[i] != 0) {
#pragma unroll
    for (int m = 0; m < 9; m++) A[m] = 0.0;
    A[0] = 1.0; A[4] = 1.0; A[8] = 1.0;
    b[0] = 0; b[1] = 0; b[2] = 0;
  }
  double* od = valDia + ((size_t)s * np + i) * 9;
#pragma unroll
  for (int m = 0; m < 9; m++) od[m] = A[m];
  double* ob = vec_b + ((size_t)s * np + i) * 3;
  ob[0] = b[0]; ob[1] = b[1]; ob[2] = b[2];
}

// per off-diagonal slot: triangle blocks then hinge blocks in element order, then the boundary condition
__global__ void __launch_bounds__(128) k_gather_slot(int np, int ncrs, int ntri, const int* __restrict__ bc,
                                                      const int* __restrict__ rowidx_of_slot,
                                                      const int* __restrict__ colidx, const int* __restrict__ st_ptr,
                                                      const int* __restrict__ st_ent, const int* __restrict__ sq_ptr,
                                                      const int* __restrict__ sq_ent, const double* __restrict__ tri_ddW,
                                                      const double* __restrict__ quadK, const double* __restrict__ quadT,
                                                      double* __restrict__ valCrs) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncrs) return;
  const int s = blockIdx.y;
  const double* tddW = tri_ddW + (size_t)s * ntri * 81;
  double A[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = st_ptr[j]; k < st_ptr[j + 1]; k++) {
    const double* h = tddW + (size_t)st_ent[k] * 9;  // entry = tri*9 + ino*3 + jno
#pragma unroll
    for (int m = 0; m < 9; m++) A[m] += h[m];
  }
  for (int k = sq_ptr[j]; k < sq_ptr[j + 1]; k++) {
    const int e = sq_ent[k];  // quad*16 + ino*4 + jno
    const int q = e >> 4, a = (e >> 2) & 3, b2 = e & 3;
    const double kk = quadK[q * 4 + a] * quadK[q * 4 + b2] * quadT[q];
    A[0] += kk; A[4] += kk; A[8] += kk;
  }
  if (bc[rowidx_of_slot[j]] != 0 || bc[colidx[j]] != 0) {
#pragma unroll
    for (int m = 0; m < 9; m++) A[m] = 0.0;
  }
  double* o = valCrs + ((size_t)s * ncrs + j) * 9;
#pragma unroll
  for (int m = 0; m < 9; m++) o[m] = A[m];
}

// UpdateIntermidiateVelocity, solve_internal_sparse.h:329-369: gradient-only gather, v[fixed] = 0,
// v += -(0.5 dt / m) dW
__global__ void k_gather_velocity(int np, int ntri, int nquad, double dt, double mass, const int* __restrict__ bc,
                                  const int* __restrict__ vt_ptr, const int* __restrict__ vt_ent,
                                  const int* __restrict__ vq_ptr, const int* __restrict__ vq_ent,
                                  const double* __restrict__ gravity, const double* __restrict__ tri_dW,
                                  const double* __restrict__ quad_dW, double* __restrict__ uvw,
                                  const int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const int s = blockIdx.y;
  if (!cnt[s * NCNT + CNT_ACTIVE]) return;  // scenes without an applied impulse keep their velocities
  const double* tdW = tri_dW + (size_t)s * ntri * 9;
  const double* qdW = quad_dW + (size_t)s * nquad * 12;
  double b[3] = {0, 0, 0};
  for (int k = vt_ptr[i]; k < vt_ptr[i + 1]; k++) {
    const double* g = tdW + (size_t)vt_ent[k] * 3;
    b[0] += g[0]; b[1] += g[1]; b[2] += g[2];
  }
  for (int k = vq_ptr[i]; k < vq_ptr[i + 1]; k++) {
    const double* g = qdW + (size_t)vq_ent[k] * 3;
    b[0] += g[0]; b[1] += g[1]; b[2] += g[2];
  }
  const double* gv = gravity + s * 3;
  b[0] -= mass * gv[0]; b[1] -= mass * gv[1]; b[2] -= mass * gv[2];
  double* v = uvw + ((size_t)s * np + i) * 3;
  double v0 = v[0], v1 = v[1], v2 = v[2];
  if (bc[i] != 0) { v0 = 0; v1 = 0; v2 = 0; }
  v0 += -(0.5 * dt / mass) * b[0];
  v1 += -(0.5 * dt / mass) * b[1];
  v2 += -(0.5 * dt / mass) * b[2];
  v[0] = v0; v[1] = v1; v[2] = v2;
}

static int upload_gravity(Ctx* c, const clothgpu_params& p) {
  std::vector<double> g((size_t)c->B * 3);
  for (int s = 0; s < c->B; s++)
    for (int k = 0; k < 3; k++) {
      const double o = c->gravity_override[(size_t)s * 3 + k];
      g[(size_t)s * 3 + k] = (o == o) ? o : p.gravity[k];
    }
  // tiny (24 B per scene) and ordered on the stream; the source is copied to staging by the runtime before return
  CG_CUDA(cudaMemcpyAsync(c->gravity, g.data(), sizeof(double) * g.size(), cudaMemcpyHostToDevice, c->stream));
  CG_CUDA(cudaStreamSynchronize(c->stream));
  return CLOTHGPU_OK;
}

static int element_pass(Ctx* c, const clothgpu_params& p, bool with_hessian) {
  if (p.stiff_bend != c->quad_stiff) {
    k_quad_rest<<<div_up(c->nquad, 128), 128, 0, KS(c, "k_quad_rest")>>>(c->nquad, c->quad, c->xyz0, p.stiff_bend, c->quadK, c->quadT);
    CG_LAUNCH_CHECK(c);
    c->quad_stiff = p.stiff_bend;
  }
  if (c->ntri > 0) {
    dim3 g(div_up(c->ntri, 128), c->B);
    k_tri_elements<<<g, 128, 0, KS(c, "k_tri_elements")>>>(c->ntri, c->np, c->tri, c->xyz0, c->xyz, p.lambda, p.myu, c->tri_dW,
                                             with_hessian ? c->tri_ddW : nullptr);
    CG_LAUNCH_CHECK(c);
  }
  if (c->nquad > 0) {
    dim3 g(div_up(c->nquad, 128), c->B);
    k_quad_elements<<<g, 128, 0, KS(c, "k_quad_elements")>>>(c->nquad, c->np, c->quad, c->xyz, c->quadK, c->quadT, c->quad_dW);
    CG_LAUNCH_CHECK(c);
  }
  return CLOTHGPU_OK;
}

int launch_assemble(Ctx* c, const clothgpu_params& p, bool with_hessian) {
  (void)with_hessian;
  CG_CHECK(upload_gravity(c, p));
  CG_CHECK(element_pass(c, p, true));
  AsmParams P;
  P.dt = p.dt; P.lambda = p.lambda; P.myu = p.myu; P.stiff_bend = p.stiff_bend; P.mass = p.mass_point;
  P.stiff_contact = p.stiff_contact; P.clearance = p.contact_clearance; P.col_kind = p.collider_kind;
  for (int k = 0; k < 4; k++) P.col_par[k] = p.collider_par[k];
  {
    dim3 g(div_up(c->np, 128), c->B);
    k_gather_vertex<<<g, 128, 0, KS(c, "k_gather_vertex")>>>(c->np, c->ntri, c->nquad, P, c->bc, c->vt_ptr, c->vt_ent, c->vq_ptr,
                                              c->vq_ent, c->xyz, c->uvw, c->gravity, c->col_table, c->tri_dW,
                                              c->tri_ddW, c->quad_dW, c->quadK, c->quadT, c->valDia, c->vec_r);
    CG_LAUNCH_CHECK(c);
  }
  if (c->ncrs > 0) {
    dim3 g(div_up(c->ncrs, 128), c->B);
    k_gather_slot<<<g, 128, 0, KS(c, "k_gather_slot")>>>(c->np, c->ncrs, c->ntri, c->bc, c->slot_row, c->colidx, c->st_ptr,
                                            c->st_ent, c->sq_ptr, c->sq_ent, c->tri_ddW, c->quadK, c->quadT, c->valCrs);
    CG_LAUNCH_CHECK(c);
  }
  return CLOTHGPU_OK;
}

int launch_update_velocity(Ctx* c, const clothgpu_params& p) {
  CG_CHECK(upload_gravity(c, p));
  CG_CHECK(element_pass(c, p, false));
  dim3 g(div_up(c->np, 128), c->B);
  k_gather_velocity<<<g, 128, 0, KS(c, "k_gather_velocity")>>>(c->np, c->ntri, c->nquad, p.dt, p.mass_point, c->bc, c->vt_ptr,
                                              c->vt_ent, c->vq_ptr, c->vq_ent, c->gravity, c->tri_dW, c->quad_dW, c->uvw, c->cnt);
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}

}  // namespace cg
