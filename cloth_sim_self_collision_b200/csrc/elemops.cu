// csrc/elemops.cu -- per-element entry points of the reference's public headers, as batches on the GPU.
//
// The reference exposes its element energies (cloth_internal_physics.h: MakePositiveDefinite_Sim22 :15-67, WdWddW_CST
// :70-176, WdWddW_Bend :180-238, WdWddW_Contact :243-280) and three geometric predicates (bvh_aabb.h:49-59:
// DistanceFaceVertex, DistanceEdgeEdge, FindCoplanerInterp) as free functions; the fused step never calls them one at
// a time, but a drop-in has to offer them.  Each entry point here evaluates n independent elements with the same
// device functions the assembly and the narrow phase use (elem_math.cuh, dev_math.cuh), so the scalar forms of the
// C++ layer (n = 1) return bit-identical values to the fused path.  No context: the current device, the default
// stream, synchronous; without a device they fail with CLOTHGPU_ERR_CUDA like everything else (no CPU fallback).
//
// Also here: MakeNormal (self_contact_sparse/main.cpp:183-211), the vertex normals the demos hand to the renderer.
#include <vector>

#include "ctx.h"
#include "dev_math.cuh"
#include "elem_math.cuh"

namespace cg {

__global__ void k_elem_psd(int n, const double* __restrict__ s2, double* __restrict__ s3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[3] = {s2[i * 3], s2[i * 3 + 1], s2[i * 3 + 2]}, b[3];
  b[0] = a[0]; b[1] = a[1]; b[2] = a[2];
  psd_clamp_sym22(a, b);
  s3[i * 3] = b[0]; s3[i * 3 + 1] = b[1]; s3[i * 3 + 2] = b[2];
}

__global__ void __launch_bounds__(128) k_elem_cst(int n, const double* __restrict__ C, const double* __restrict__ c,
                                                   double lambda, double myu, double* __restrict__ W,
                                                   double* __restrict__ dW, double* __restrict__ ddW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double Cl[9], cl[9];
#pragma unroll
  for (int k = 0; k < 9; k++) { Cl[k] = C[(size_t)i * 9 + k]; cl[k] = c[(size_t)i * 9 + k]; }
  cst_element(Cl, cl, lambda, myu, W + i, dW + (size_t)i * 9, ddW + (size_t)i * 81);
}

// WdWddW_Bend, cloth_internal_physics.h:180-238 in the reference's own output layout: ddW[ino][jno][3][3] with the
// scalar K_ino K_jno stiff/((A0+A1) L0^2) on the block diagonals, dW = ddW c (12-term sums incl. the zero entries),
// W accumulated as sum dW.c
__global__ void k_elem_bend(int n, const double* __restrict__ Cg, const double* __restrict__ cg, double stiff,
                            double* __restrict__ W, double* __restrict__ dW, double* __restrict__ ddW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double C[12], c[12];
#pragma unroll
  for (int k = 0; k < 12; k++) { C[k] = Cg[(size_t)i * 12 + k]; c[k] = cg[(size_t)i * 12 + k]; }
  const double A0 = tri_area(C + 0, C + 6, C + 9);
  const double A1 = tri_area(C + 3, C + 9, C + 6);
  const double L0 = dist3(C + 6, C + 9);
  const double H0 = A0 * 2.0 / L0;
  const double H1 = A1 * 2.0 / L0;
  double e23[3], e02[3], e03[3], e12[3], e13[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    e23[k] = C[9 + k] - C[6 + k];
    e02[k] = C[6 + k] - C[0 + k];
    e03[k] = C[9 + k] - C[0 + k];
    e12[k] = C[6 + k] - C[3 + k];
    e13[k] = C[9 + k] - C[3 + k];
  }
  double cot023, cot032, cot123, cot132;
  { const double r2 = -dot3(e02, e23), r3 = +dot3(e03, e23); cot023 = r2 / H0; cot032 = r3 / H0; }
  { const double r2 = -dot3(e12, e23), r3 = +dot3(e13, e23); cot123 = r2 / H1; cot132 = r3 / H1; }
  const double tmp0 = stiff / ((A0 + A1) * L0 * L0);
  const double K[4] = {-cot023 - cot032, -cot123 - cot132, cot032 + cot132, cot023 + cot123};
  double* H = ddW + (size_t)i * 144;
  for (int k = 0; k < 144; k++) H[k] = 0.0;
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++) {
      const double t = K[a] * K[b] * tmp0;
      double* blk = H + (a * 4 + b) * 9;
      blk[0] = t; blk[4] = t; blk[8] = t;
    }
  double w = 0;
  for (int a = 0; a < 4; a++)
    for (int d = 0; d < 3; d++) {
      double g = 0;
      for (int b = 0; b < 4; b++)
        for (int e = 0; e < 3; e++) g += H[(a * 4 + b) * 9 + d * 3 + e] * c[b * 3 + e];
      dW[(size_t)i * 12 + a * 3 + d] = g;
      w += g * c[a * 3 + d];
    }
  W[i] = w;
}

// WdWddW_Contact, cloth_internal_physics.h:243-280; pdn: per element {pd, nx, ny, nz} as returned by the caller's
// penetrationDepth callback (the callback itself is host code and stays on the host)
__global__ void k_elem_contact(int n, const double* __restrict__ pdn, double stiff, double clearance,
                               double* __restrict__ W, double* __restrict__ dW, double* __restrict__ ddW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double pd = pdn[i * 4];
  const double nn[3] = {pdn[i * 4 + 1], pdn[i * 4 + 2], pdn[i * 4 + 3]};
  pd += clearance;
  if (pd < 0) {
    W[i] = 0;
    for (int k = 0; k < 3; k++) dW[i * 3 + k] = 0;
    for (int k = 0; k < 9; k++) ddW[i * 9 + k] = 0;
    return;
  }
  W[i] = 0.5 * stiff * pd * pd;
  for (int k = 0; k < 3; k++) dW[i * 3 + k] = -stiff * pd * nn[k];
  for (int r = 0; r < 3; r++)
    for (int q = 0; q < 3; q++) ddW[i * 9 + r * 3 + q] = stiff * nn[r] * nn[q];
}

// mode 0: DistanceFaceVertex (bvh_aabb.cpp:321-340) p = {p0,p1,p2,p3} -> out {dist, w0, w1}
// mode 1: DistanceEdgeEdge (:343-384)                p = {p0,p1,q0,q1} -> out {dist, ratio_p, ratio_q}
// mode 2: FindCoplanerInterp (:598-662)              p = {s0..s3}, e = {e0..e3} -> out {t}
__global__ void k_geom(int mode, int n, const double* __restrict__ p, const double* __restrict__ e,
                       double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* q = p + (size_t)i * 12;
  const V3 a0 = v3(q[0], q[1], q[2]), a1 = v3(q[3], q[4], q[5]), a2 = v3(q[6], q[7], q[8]), a3 = v3(q[9], q[10], q[11]);
  if (mode == 0) {
    double w0 = 0, w1 = 0;
    const double d = dist_face_vertex(a0, a1, a2, a3, w0, w1);
    out[i * 3] = d; out[i * 3 + 1] = w0; out[i * 3 + 2] = w1;
  } else if (mode == 1) {
    double rp = 0, rq = 0;
    const double d = dist_edge_edge(a0, a1, a2, a3, rp, rq);
    out[i * 3] = d; out[i * 3 + 1] = rp; out[i * 3 + 2] = rq;
  } else {
    const double* r = e + (size_t)i * 12;
    out[i] = coplanar_time(a0, a1, a2, a3, v3(r[0], r[1], r[2]), v3(r[3], r[4], r[5]), v3(r[6], r[7], r[8]),
                           v3(r[9], r[10], r[11]));
  }
}

// MakeNormal, self_contact_sparse/main.cpp:183-211: unit normal per triangle (UnitNormalAreaTri3D), summed per vertex
// in ascending triangle order (the gather list of the assembly), normalised
__global__ void k_tri_unit_normals(int ntri, int np, const int* __restrict__ tri, const double* __restrict__ xyz,
                                   double* __restrict__ tn) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntri) return;
  const int s = blockIdx.y;
  const double* x = xyz + (size_t)s * np * 3;
  double c[9], nrm[3];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int k = 0; k < 3; k++) c[a * 3 + k] = x[tri[t * 3 + a] * 3 + k];
  tri_normal(nrm, c, c + 3, c + 6);
  const double area = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]) * 0.5;
  const double invlen = 0.5 / area;
  double* o = tn + ((size_t)s * ntri + t) * 3;
  o[0] = nrm[0] * invlen; o[1] = nrm[1] * invlen; o[2] = nrm[2] * invlen;
}
__global__ void k_vertex_normals(int np, int ntri, const int* __restrict__ vt_ptr, const int* __restrict__ vt_ent,
                                 const double* __restrict__ tn, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const int s = blockIdx.y;
  const double* t = tn + (size_t)s * ntri * 3;
  double a[3] = {0, 0, 0};
  for (int k = vt_ptr[i]; k < vt_ptr[i + 1]; k++) {
    const double* g = t + (size_t)(vt_ent[k] / 3) * 3;
    a[0] += g[0]; a[1] += g[1]; a[2] += g[2];
  }
  const double sq = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
  const double inv = 1.0 / sqrt(sq);
  double* o = out + ((size_t)s * np + i) * 3;
  o[0] = a[0] * inv; o[1] = a[1] * inv; o[2] = a[2] * inv;
}

namespace {
// device copies of host arrays for one call; freed on scope exit
struct Tmp {
  std::vector<void*> ptrs;
  ~Tmp() { for (void* p : ptrs) cudaFree(p); }
  int in(const double* h, size_t n, double** d) {
    CG_CUDA(cudaMalloc((void**)d, std::max<size_t>(n, 1) * sizeof(double)));
    ptrs.push_back(*d);
    CG_CUDA(cudaMemcpy(*d, h, n * sizeof(double), cudaMemcpyHostToDevice));
    return CLOTHGPU_OK;
  }
  int out(size_t n, double** d) {
    CG_CUDA(cudaMalloc((void**)d, std::max<size_t>(n, 1) * sizeof(double)));
    ptrs.push_back(*d);
    return CLOTHGPU_OK;
  }
};
int back(double* h, const double* d, size_t n) {
  CG_CUDA(cudaMemcpy(h, d, n * sizeof(double), cudaMemcpyDeviceToHost));
  return CLOTHGPU_OK;
}
int finish(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return cuda_fail(e, what, __FILE__, __LINE__);
  return CLOTHGPU_OK;
}
}  // namespace

}  // namespace cg

using namespace cg;

#define NEED(cond)                                                     \
  do {                                                                 \
    if (!(cond)) { set_error("bad argument: " #cond); return CLOTHGPU_ERR_ARG; } \
  } while (0)

extern "C" {

int clothgpu_elem_psd_sym22(int n, const double* s2, double* s3) {
  NEED(n >= 0 && s2 && s3);
  if (n == 0) return CLOTHGPU_OK;
  Tmp t;
  double *a, *b;
  CG_CHECK(t.in(s2, (size_t)n * 3, &a)); CG_CHECK(t.out((size_t)n * 3, &b));
  k_elem_psd<<<div_up(n, 128), 128>>>(n, a, b);
  CG_CHECK(finish("k_elem_psd"));
  return back(s3, b, (size_t)n * 3);
}

int clothgpu_elem_cst(int n, const double* C, const double* c, double lambda, double myu, double* W, double* dW,
                      double* ddW) {
  NEED(n >= 0 && C && c && W && dW && ddW);
  if (n == 0) return CLOTHGPU_OK;
  Tmp t;
  double *dC, *dc, *w, *g, *h;
  CG_CHECK(t.in(C, (size_t)n * 9, &dC)); CG_CHECK(t.in(c, (size_t)n * 9, &dc));
  CG_CHECK(t.out(n, &w)); CG_CHECK(t.out((size_t)n * 9, &g)); CG_CHECK(t.out((size_t)n * 81, &h));
  k_elem_cst<<<div_up(n, 128), 128>>>(n, dC, dc, lambda, myu, w, g, h);
  CG_CHECK(finish("k_elem_cst"));
  CG_CHECK(back(W, w, n)); CG_CHECK(back(dW, g, (size_t)n * 9));
  return back(ddW, h, (size_t)n * 81);
}

int clothgpu_elem_bend(int n, const double* C, const double* c, double stiff, double* W, double* dW, double* ddW) {
  NEED(n >= 0 && C && c && W && dW && ddW);
  if (n == 0) return CLOTHGPU_OK;
  Tmp t;
  double *dC, *dc, *w, *g, *h;
  CG_CHECK(t.in(C, (size_t)n * 12, &dC)); CG_CHECK(t.in(c, (size_t)n * 12, &dc));
  CG_CHECK(t.out(n, &w)); CG_CHECK(t.out((size_t)n * 12, &g)); CG_CHECK(t.out((size_t)n * 144, &h));
  k_elem_bend<<<div_up(n, 64), 64>>>(n, dC, dc, stiff, w, g, h);
  CG_CHECK(finish("k_elem_bend"));
  CG_CHECK(back(W, w, n)); CG_CHECK(back(dW, g, (size_t)n * 12));
  return back(ddW, h, (size_t)n * 144);
}

int clothgpu_elem_contact(int n, const double* pd_normal, double stiff_contact, double contact_clearance, double* W,
                          double* dW, double* ddW) {
  NEED(n >= 0 && pd_normal && W && dW && ddW);
  if (n == 0) return CLOTHGPU_OK;
  Tmp t;
  double *p, *w, *g, *h;
  CG_CHECK(t.in(pd_normal, (size_t)n * 4, &p));
  CG_CHECK(t.out(n, &w)); CG_CHECK(t.out((size_t)n * 3, &g)); CG_CHECK(t.out((size_t)n * 9, &h));
  k_elem_contact<<<div_up(n, 128), 128>>>(n, p, stiff_contact, contact_clearance, w, g, h);
  CG_CHECK(finish("k_elem_contact"));
  CG_CHECK(back(W, w, n)); CG_CHECK(back(dW, g, (size_t)n * 3));
  return back(ddW, h, (size_t)n * 9);
}

static int geom(int mode, int n, const double* p, const double* e, double* o0, double* o1, double* o2) {
  if (n == 0) return CLOTHGPU_OK;
  Tmp t;
  double *dp, *de = nullptr, *out;
  const int w = mode == 2 ? 1 : 3;
  CG_CHECK(t.in(p, (size_t)n * 12, &dp));
  if (mode == 2) CG_CHECK(t.in(e, (size_t)n * 12, &de));
  CG_CHECK(t.out((size_t)n * w, &out));
  k_geom<<<div_up(n, 128), 128>>>(mode, n, dp, de, out);
  CG_CHECK(finish("k_geom"));
  std::vector<double> h((size_t)n * w);
  CG_CHECK(back(h.data(), out, h.size()));
  for (int i = 0; i < n; i++) {
    o0[i] = h[(size_t)i * w];
    if (w == 3) { o1[i] = h[(size_t)i * 3 + 1]; o2[i] = h[(size_t)i * 3 + 2]; }
  }
  return CLOTHGPU_OK;
}
int clothgpu_dist_face_vertex(int n, const double* p, double* dist, double* w0, double* w1) {
  NEED(n >= 0 && p && dist && w0 && w1);
  return geom(0, n, p, nullptr, dist, w0, w1);
}
int clothgpu_dist_edge_edge(int n, const double* p, double* dist, double* ratio_p, double* ratio_q) {
  NEED(n >= 0 && p && dist && ratio_p && ratio_q);
  return geom(1, n, p, nullptr, dist, ratio_p, ratio_q);
}
int clothgpu_coplanar_interp(int n, const double* s, const double* e, double* t) {
  NEED(n >= 0 && s && e && t);
  return geom(2, n, s, e, t, nullptr, nullptr);
}

int clothgpu_make_normal(clothgpu_ctx* c, int buf_xyz, int scene, double* normals) {
  if (!c || !normals || scene < 0 || scene >= c->B) { set_error("bad argument"); return CLOTHGPU_ERR_ARG; }
  cudaSetDevice(c->device);
  double* x = vec_buf(c, buf_xyz);
  if (!x) { set_error("not a vector buffer id"); return CLOTHGPU_ERR_ARG; }
  // scratch: the triangle-gradient buffer of the assembly (9 doubles per triangle, rewritten by the next step)
  dim3 g(div_up(c->ntri, 128), c->B);
  k_tri_unit_normals<<<g, 128, 0, KS(c, "k_tri_unit_normals")>>>(c->ntri, c->np, c->tri, x, c->tri_dW);
  CG_LAUNCH_CHECK(c);
  dim3 g2(div_up(c->np, 128), c->B);
  if (!c->normals) CG_CUDA(cudaMalloc((void**)&c->normals, sizeof(double) * (size_t)c->B * c->ndof));
  k_vertex_normals<<<g2, 128, 0, KS(c, "k_vertex_normals")>>>(c->np, c->ntri, c->vt_ptr, c->vt_ent, c->tri_dW, c->normals);
  CG_LAUNCH_CHECK(c);
  CG_CUDA(cudaMemcpyAsync(normals, c->normals + (size_t)scene * c->ndof, sizeof(double) * c->ndof, cudaMemcpyDeviceToHost, c->stream));
  CG_CUDA(cudaStreamSynchronize(c->stream));
  return CLOTHGPU_OK;
}

}  // extern "C"
