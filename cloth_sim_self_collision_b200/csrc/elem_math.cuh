// csrc/elem_math.cuh -- element energies shared by the assembly (assembly.cu) and the per-element entry points
// (elemops.cu): WdWddW_CST and MakePositiveDefinite_Sim22 of cloth_internal_physics.h:15-176 with the double[3] helpers
// of utility.h they use.  Same operand order as the reference, compiled with -fmad=false: bit-identical results.
#pragma once

namespace cg {

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// utility.h:21-25
__device__ __forceinline__ void cross3(double* r, const double* a, const double* b) {
  r[0] = a[1] * b[2] - b[1] * a[2];
  r[1] = a[2] * b[0] - b[2] * a[0];
  r[2] = a[0] * b[1] - b[0] * a[1];
}
// unnormalised triangle normal as written in utility.h:48-50 / 66-68
__device__ __forceinline__ void tri_normal(double* n, const double* v1, const double* v2, const double* v3p) {
  n[0] = (v2[1] - v1[1]) * (v3p[2] - v1[2]) - (v3p[1] - v1[1]) * (v2[2] - v1[2]);
  n[1] = (v2[2] - v1[2]) * (v3p[0] - v1[0]) - (v3p[2] - v1[2]) * (v2[0] - v1[0]);
  n[2] = (v2[0] - v1[0]) * (v3p[1] - v1[1]) - (v3p[0] - v1[0]) * (v2[1] - v1[1]);
}
// TriArea3D, utility.h:41-52
__device__ __forceinline__ double tri_area(const double* v1, const double* v2, const double* v3p) {
  double n[3];
  tri_normal(n, v1, v2, v3p);
  return sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) * 0.5;
}
// Distance3D, utility.h:33-37
__device__ __forceinline__ double dist3(const double* a, const double* b) {
  const double t = (b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]) + (b[2] - a[2]) * (b[2] - a[2]);
  return sqrt(t);
}

// MakePositiveDefinite_Sim22, cloth_internal_physics.h:15-67: clamp the symmetric 2x2 stress (s00,s11,s01) to PSD
static __device__ void psd_clamp_sym22(const double* s2, double* s3) {
  const double b = (s2[0] + s2[1]) * 0.5;
  const double d = (s2[0] - s2[1]) * (s2[0] - s2[1]) * 0.25 + s2[2] * s2[2];
  const double e = sqrt(d);
  if (b - e > 1.0e-20) { s3[0] = s2[0]; s3[1] = s2[1]; s3[2] = s2[2]; return; }
  if (b + e < 0) { s3[0] = s3[1] = s3[2] = 0; return; }
  const double l = b + e;
  double t0[2] = {s2[0] - l, s2[2]};
  double t1[2] = {s2[2], s2[1] - l};
  const double q0 = t0[0] * t0[0] + t0[1] * t0[1];
  const double q1 = t1[0] * t1[0] + t1[1] * t1[1];
  if (q0 > q1) {
    if (q0 < 1.0e-20) { s3[0] = s3[1] = s3[2] = 0; return; }
    const double inv = 1.0 / sqrt(q0);
    t0[0] *= inv; t0[1] *= inv;
    s3[0] = l * t0[0] * t0[0]; s3[1] = l * t0[1] * t0[1]; s3[2] = l * t0[0] * t0[1];
  } else {
    if (q1 < 1.0e-20) { s3[0] = s3[1] = s3[2] = 0; return; }
    const double inv = 1.0 / sqrt(q1);
    t1[0] *= inv; t1[1] *= inv;
    s3[0] = l * t1[0] * t1[0]; s3[1] = l * t1[1] * t1[1]; s3[2] = l * t1[0] * t1[1];
  }
}

// WdWddW_CST, cloth_internal_physics.h:70-176.  C, c: rest / current vertex positions [3][3];
// dW: [ino][idim]; ddW: [ino*3+jno][idim*3+jdim] (nullptr: gradient only); W: energy (nullptr: not wanted)
__device__ __forceinline__ void cst_element(const double* C, const double* c, double lambda, double myu, double* W,
                                            double* dW, double* ddW) {
  const double dNdr[3][2] = {{-1.0, -1.0}, {+1.0, +0.0}, {+0.0, +1.0}};
  double D[3][3], G[2][3], d[2][3], Area, E2[3], GG[3], Cons[3][3], S2[3], S3[3];
#pragma unroll
  for (int i = 0; i < 3; i++) { D[0][i] = C[3 + i] - C[i]; D[1][i] = C[6 + i] - C[i]; }
  {  // UnitNormalAreaTri3D, utility.h:58-73
    tri_normal(D[2], C, C + 3, C + 6);
    Area = sqrt(D[2][0] * D[2][0] + D[2][1] * D[2][1] + D[2][2] * D[2][2]) * 0.5;
    const double invlen = 0.5 / Area;
    D[2][0] *= invlen; D[2][1] *= invlen; D[2][2] *= invlen;
  }
  {  // dual basis, :87-96
    cross3(G[0], D[1], D[2]);
    double inv = 1.0 / dot3(G[0], D[0]);
    G[0][0] *= inv; G[0][1] *= inv; G[0][2] *= inv;
    cross3(G[1], D[2], D[0]);
    inv = 1.0 / dot3(G[1], D[1]);
    G[1][0] *= inv; G[1][1] *= inv; G[1][2] *= inv;
  }
#pragma unroll
  for (int i = 0; i < 3; i++) { d[0][i] = c[3 + i] - c[i]; d[1][i] = c[6 + i] - c[i]; }
  E2[0] = 0.5 * (dot3(d[0], d[0]) - dot3(D[0], D[0]));
  E2[1] = 0.5 * (dot3(d[1], d[1]) - dot3(D[1], D[1]));
  E2[2] = 1.0 * (dot3(d[0], d[1]) - dot3(D[0], D[1]));
  GG[0] = dot3(G[0], G[0]); GG[1] = dot3(G[1], G[1]); GG[2] = dot3(G[1], G[0]);
  Cons[0][0] = lambda * GG[0] * GG[0] + 2 * myu * (GG[0] * GG[0]);
  Cons[0][1] = lambda * GG[0] * GG[1] + 2 * myu * (GG[2] * GG[2]);
  Cons[0][2] = lambda * GG[0] * GG[2] + 2 * myu * (GG[0] * GG[2]);
  Cons[1][0] = lambda * GG[1] * GG[0] + 2 * myu * (GG[2] * GG[2]);
  Cons[1][1] = lambda * GG[1] * GG[1] + 2 * myu * (GG[1] * GG[1]);
  Cons[1][2] = lambda * GG[1] * GG[2] + 2 * myu * (GG[2] * GG[1]);
  Cons[2][0] = lambda * GG[2] * GG[0] + 2 * myu * (GG[0] * GG[2]);
  Cons[2][1] = lambda * GG[2] * GG[1] + 2 * myu * (GG[2] * GG[1]);
  Cons[2][2] = lambda * GG[2] * GG[2] + 1 * myu * (GG[0] * GG[1] + GG[2] * GG[2]);
#pragma unroll
  for (int i = 0; i < 3; i++) S2[i] = Cons[i][0] * E2[0] + Cons[i][1] * E2[1] + Cons[i][2] * E2[2];
#pragma unroll
  for (int ino = 0; ino < 3; ino++)
#pragma unroll
    for (int idim = 0; idim < 3; idim++)
      dW[ino * 3 + idim] = Area * (+S2[0] * d[0][idim] * dNdr[ino][0] + S2[2] * d[0][idim] * dNdr[ino][1] +
                                   S2[2] * d[1][idim] * dNdr[ino][0] + S2[1] * d[1][idim] * dNdr[ino][1]);
  if (W) *W = 0.5 * Area * (E2[0] * S2[0] + E2[1] * S2[1] + E2[2] * S2[2]);  // :131-136
  if (ddW == nullptr) return;
  S3[0] = S2[0]; S3[1] = S2[1]; S3[2] = S2[2];
  psd_clamp_sym22(S2, S3);
  // the 16 material terms in the order of :147-162: (a, r) -> d[a][idim]*dNdr[ino][r] on the left,
  // (b, q) -> d[b][jdim]*dNdr[jno][q] on the right, Cons index in the middle
  const int TA[16] = {0, 0, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1, 1, 1};
  const int TR[16] = {0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0};
  const int TC0[16] = {0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2};
  const int TC1[16] = {0, 1, 2, 2, 0, 1, 2, 2, 0, 1, 2, 2, 0, 1, 2, 2};
  const int TB[16] = {0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1};
  const int TQ[16] = {0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0, 0, 1, 1, 0};
#pragma unroll
  for (int ino = 0; ino < 3; ino++) {
#pragma unroll
    for (int jno = 0; jno < 3; jno++) {
      double blk[9];
#pragma unroll
      for (int idim = 0; idim < 3; idim++) {
#pragma unroll
        for (int jdim = 0; jdim < 3; jdim++) {
          double acc = 0;
#pragma unroll
          for (int k = 0; k < 16; k++)
            acc += d[TA[k]][idim] * dNdr[ino][TR[k]] * Cons[TC0[k]][TC1[k]] * d[TB[k]][jdim] * dNdr[jno][TQ[k]];
          blk[idim * 3 + jdim] = acc * Area;
        }
      }
      const double g = Area * (+S3[0] * dNdr[ino][0] * dNdr[jno][0] + S3[2] * dNdr[ino][0] * dNdr[jno][1] +
                               S3[2] * dNdr[ino][1] * dNdr[jno][0] + S3[1] * dNdr[ino][1] * dNdr[jno][1]);
      blk[0] += g; blk[4] += g; blk[8] += g;
#pragma unroll
      for (int k = 0; k < 9; k++) ddW[(ino * 3 + jno) * 9 + k] = blk[k];
    }
  }
}

}  // namespace cg
