/* oracle/cloth_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain C99 restatement of the per-step cloth solve + self-collision pipeline of
 * nobuyuki83/cloth_sim_self_collision (reference file:line cited on each function), written against flat arrays.
 * It is the checker the CUDA path is compared with; it is never linked into, imported by or called from the
 * product library.  It is pinned against the unmodified reference (oracle/_ref/libclothref.so, see
 * oracle/ref_shim.cpp) by tests/test_oracle_vs_ref.py and against the committed vectors in tests/golden/.
 *
 * Floating point: every expression keeps the reference's operand order and association, and the file must be
 * compiled with -ffp-contract=off (oracle/Makefile) -- FMA contraction forks the trajectories.
 * Known reference quirks are reproduced on purpose and marked QUIRK.
 */
#include "cloth_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------ 3-vectors */
typedef struct { double x, y, z; } v3;

static v3 v3_make(double x, double y, double z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
static v3 v3_ld(const double* a, int i) { return v3_make(a[i * 3 + 0], a[i * 3 + 1], a[i * 3 + 2]); }
static v3 v3_add(v3 a, v3 b) { return v3_make(a.x + b.x, a.y + b.y, a.z + b.z); }
static v3 v3_sub(v3 a, v3 b) { return v3_make(a.x - b.x, a.y - b.y, a.z - b.z); }
static v3 v3_scale(double s, v3 a) { return v3_make(a.x * s, a.y * s, a.z * s); }
static v3 v3_neg(v3 a) { return v3_make(-a.x, -a.y, -a.z); }
/* vector3d.h:214-217 */
static double v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* vector3d.h:219-226 */
static v3 v3_cross(v3 a, v3 b) {
  return v3_make(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* vector3d.h:94 */
static double v3_len(v3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
/* vector3d.h:97-103 */
static v3 v3_normalized(v3 a) {
  double invmag = 1.0 / v3_len(a);
  return v3_make(a.x * invmag, a.y * invmag, a.z * invmag);
}
/* vector3d.h:228-230 */
static double v3_triple(v3 a, v3 b, v3 c) {
  return a.x * (b.y * c.z - b.z * c.y) + a.y * (b.z * c.x - b.x * c.z) + a.z * (b.x * c.y - b.y * c.x);
}
/* vector3d.h:63-67: division silently skipped for |d| < 1e-16 */
static v3 v3_div(v3 a, double d) {
  if (fabs(d) < 1.0e-16) return a;
  return v3_make(a.x / d, a.y / d, a.z / d);
}

/* utility.h:15-17 */
static double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
/* utility.h:21-25 */
static void cross3(double* r, const double* a, const double* b) {
  r[0] = a[1] * b[2] - b[1] * a[2];
  r[1] = a[2] * b[0] - b[2] * a[0];
  r[2] = a[0] * b[1] - b[0] * a[1];
}
/* normal (unnormalised) of triangle v1 v2 v3 as written in utility.h:48-50 / 66-68 */
static void tri_normal(double* n, const double* v1, const double* v2, const double* v3p) {
  n[0] = (v2[1] - v1[1]) * (v3p[2] - v1[2]) - (v3p[1] - v1[1]) * (v2[2] - v1[2]);
  n[1] = (v2[2] - v1[2]) * (v3p[0] - v1[0]) - (v3p[2] - v1[2]) * (v2[0] - v1[0]);
  n[2] = (v2[0] - v1[0]) * (v3p[1] - v1[1]) - (v3p[0] - v1[0]) * (v2[1] - v1[1]);
}
/* utility.h:41-52 */
static double tri_area(const double* v1, const double* v2, const double* v3p) {
  double n[3];
  tri_normal(n, v1, v2, v3p);
  return sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) * 0.5;
}
/* utility.h:33-37 */
static double dist3(const double* a, const double* b) {
  const double t = (b[0] - a[0]) * (b[0] - a[0]) + (b[1] - a[1]) * (b[1] - a[1]) + (b[2] - a[2]) * (b[2] - a[2]);
  return sqrt(t);
}

/* ------------------------------------------------------------------------------------------------ colliders */
/* kind 0: none (internal_cloth_sparse/main.cpp:107-110); 1: plane z = par[0] (self_contact_sparse/main.cpp:110-114);
 * 2: sphere obstacle centre par[0..2] radius par[3] (internal_cloth_sparse/main.cpp:120-132);
 * 3: sphere container, kind 2 with normal and depth negated (self_contact_sparse/main.cpp:117-133) */
static void collider_eval(int kind, const double* par, const double* p, double* pd, double* n) {
  if (kind == 1) {
    n[0] = 0.0; n[1] = 0.0; n[2] = 1.0;
    *pd = par[0] - p[2];
  } else if (kind == 2 || kind == 3) {
    double len;
    n[0] = p[0] - par[0]; n[1] = p[1] - par[1]; n[2] = p[2] - par[2];
    len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    n[0] /= len; n[1] /= len; n[2] /= len;
    *pd = par[3] - len;
    if (kind == 3) {
      n[0] *= -1; n[1] *= -1; n[2] *= -1;
      *pd *= -1;
    }
  } else {
    n[0] = n[1] = n[2] = 0.0;
    *pd = -100;
  }
}

/* ------------------------------------------------------------------------------------------------ energies */
/* MakePositiveDefinite_Sim22, cloth_internal_physics.h:15-67: clamp the symmetric 2x2 stress (s00,s11,s01) to PSD */
static void psd_clamp_sym22(const double* s2, double* s3) {
  const double b = (s2[0] + s2[1]) * 0.5;
  const double d = (s2[0] - s2[1]) * (s2[0] - s2[1]) * 0.25 + s2[2] * s2[2];
  const double e = sqrt(d);
  double l, t0[2], t1[2], q0, q1;
  if (b - e > 1.0e-20) { s3[0] = s2[0]; s3[1] = s2[1]; s3[2] = s2[2]; return; }
  if (b + e < 0) { s3[0] = s3[1] = s3[2] = 0; return; }
  l = b + e;
  t0[0] = s2[0] - l; t0[1] = s2[2];
  t1[0] = s2[2];     t1[1] = s2[1] - l;
  q0 = t0[0] * t0[0] + t0[1] * t0[1];
  q1 = t1[0] * t1[0] + t1[1] * t1[1];
  if (q0 > q1) {
    double inv;
    if (q0 < 1.0e-20) { s3[0] = s3[1] = s3[2] = 0; return; }
    inv = 1.0 / sqrt(q0);
    t0[0] *= inv; t0[1] *= inv;
    s3[0] = l * t0[0] * t0[0]; s3[1] = l * t0[1] * t0[1]; s3[2] = l * t0[0] * t0[1];
  } else {
    double inv;
    if (q1 < 1.0e-20) { s3[0] = s3[1] = s3[2] = 0; return; }
    inv = 1.0 / sqrt(q1);
    t1[0] *= inv; t1[1] *= inv;
    s3[0] = l * t1[0] * t1[0]; s3[1] = l * t1[1] * t1[1]; s3[2] = l * t1[0] * t1[1];
  }
}

/* WdWddW_CST, cloth_internal_physics.h:70-176.  C, c: [3][3]; dW: [3][3]; ddW: [ino][jno][idim][jdim] */
void orc_cst(const double* C, const double* c, double lambda, double myu, double* W, double* dW, double* ddW) {
  static const double dNdr[3][2] = {{-1.0, -1.0}, {+1.0, +0.0}, {+0.0, +1.0}};
  double D[3][3], G[2][3], d[2][3], Area, E2[3], GG[3], Cons[3][3], S2[3], S3[3];
  int i, ino, jno, idim, jdim;
  for (i = 0; i < 3; i++) { D[0][i] = C[3 + i] - C[i]; D[1][i] = C[6 + i] - C[i]; }
  { /* UnitNormalAreaTri3D, utility.h:58-73 */
    double invlen;
    tri_normal(D[2], C, C + 3, C + 6);
    Area = sqrt(D[2][0] * D[2][0] + D[2][1] * D[2][1] + D[2][2] * D[2][2]) * 0.5;
    invlen = 0.5 / Area;
    D[2][0] *= invlen; D[2][1] *= invlen; D[2][2] *= invlen;
  }
  { /* dual basis, :87-96 */
    double inv;
    cross3(G[0], D[1], D[2]);
    inv = 1.0 / dot3(G[0], D[0]);
    G[0][0] *= inv; G[0][1] *= inv; G[0][2] *= inv;
    cross3(G[1], D[2], D[0]);
    inv = 1.0 / dot3(G[1], D[1]);
    G[1][0] *= inv; G[1][1] *= inv; G[1][2] *= inv;
  }
  for (i = 0; i < 3; i++) { d[0][i] = c[3 + i] - c[i]; d[1][i] = c[6 + i] - c[i]; }
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
  for (i = 0; i < 3; i++) S2[i] = Cons[i][0] * E2[0] + Cons[i][1] * E2[1] + Cons[i][2] * E2[2];
  *W = 0.5 * Area * (E2[0] * S2[0] + E2[1] * S2[1] + E2[2] * S2[2]);
  for (ino = 0; ino < 3; ino++)
    for (idim = 0; idim < 3; idim++)
      dW[ino * 3 + idim] = Area * (+S2[0] * d[0][idim] * dNdr[ino][0] + S2[2] * d[0][idim] * dNdr[ino][1] +
                                   S2[2] * d[1][idim] * dNdr[ino][0] + S2[1] * d[1][idim] * dNdr[ino][1]);
  S3[0] = S2[0]; S3[1] = S2[1]; S3[2] = S2[2];
  psd_clamp_sym22(S2, S3);
  for (ino = 0; ino < 3; ino++) {
    for (jno = 0; jno < 3; jno++) {
      double* blk = ddW + (ino * 3 + jno) * 9;
      double g;
      for (idim = 0; idim < 3; idim++) {
        for (jdim = 0; jdim < 3; jdim++) {
          /* the 16 material terms in the order of :147-162; (a, r) -> d[a][idim]*dNdr[ino][r] on the left,
           * (b, s) -> d[b][jdim]*dNdr[jno][s] on the right, Cons index in the middle */
          static const int T[16][6] = {
              {0, 0, 0, 0, 0, 0}, {0, 0, 0, 1, 1, 1}, {0, 0, 0, 2, 0, 1}, {0, 0, 0, 2, 1, 0},
              {1, 1, 1, 0, 0, 0}, {1, 1, 1, 1, 1, 1}, {1, 1, 1, 2, 0, 1}, {1, 1, 1, 2, 1, 0},
              {0, 1, 2, 0, 0, 0}, {0, 1, 2, 1, 1, 1}, {0, 1, 2, 2, 0, 1}, {0, 1, 2, 2, 1, 0},
              {1, 0, 2, 0, 0, 0}, {1, 0, 2, 1, 1, 1}, {1, 0, 2, 2, 0, 1}, {1, 0, 2, 2, 1, 0}};
          double acc = 0;
          int t;
          for (t = 0; t < 16; t++)
            acc += d[T[t][0]][idim] * dNdr[ino][T[t][1]] * Cons[T[t][2]][T[t][3]] * d[T[t][4]][jdim] * dNdr[jno][T[t][5]];
          blk[idim * 3 + jdim] = acc * Area;
        }
      }
      g = Area * (+S3[0] * dNdr[ino][0] * dNdr[jno][0] + S3[2] * dNdr[ino][0] * dNdr[jno][1] +
                  S3[2] * dNdr[ino][1] * dNdr[jno][0] + S3[1] * dNdr[ino][1] * dNdr[jno][1]);
      blk[0] += g; blk[4] += g; blk[8] += g;
    }
  }
}

/* WdWddW_Bend, cloth_internal_physics.h:180-238.  C, c: [4][3]; dW: [4][3]; ddW: [4][4][3][3].
 * QUIRK: the reference accumulates the energy into the caller's variable; here *W is the accumulation from 0. */
void orc_bend(const double* C, const double* c, double stiff, double* W, double* dW, double* ddW) {
  const double A0 = tri_area(C + 0, C + 6, C + 9);
  const double A1 = tri_area(C + 3, C + 9, C + 6);
  const double L0 = dist3(C + 6, C + 9);
  const double H0 = A0 * 2.0 / L0;
  const double H1 = A1 * 2.0 / L0;
  double e23[3], e02[3], e03[3], e12[3], e13[3], cot023, cot032, cot123, cot132, tmp0, K[4], w = 0;
  int i, ino, jno, idim, jdim;
  for (i = 0; i < 3; i++) {
    e23[i] = C[9 + i] - C[6 + i];
    e02[i] = C[6 + i] - C[0 + i];
    e03[i] = C[9 + i] - C[0 + i];
    e12[i] = C[6 + i] - C[3 + i];
    e13[i] = C[9 + i] - C[3 + i];
  }
  { const double r2 = -dot3(e02, e23), r3 = +dot3(e03, e23); cot023 = r2 / H0; cot032 = r3 / H0; }
  { const double r2 = -dot3(e12, e23), r3 = +dot3(e13, e23); cot123 = r2 / H1; cot132 = r3 / H1; }
  tmp0 = stiff / ((A0 + A1) * L0 * L0);
  K[0] = -cot023 - cot032; K[1] = -cot123 - cot132; K[2] = cot032 + cot132; K[3] = cot023 + cot123;
  for (i = 0; i < 144; i++) ddW[i] = 0;
  for (ino = 0; ino < 4; ino++)
    for (jno = 0; jno < 4; jno++) {
      const double tmp = K[ino] * K[jno] * tmp0;
      double* blk = ddW + (ino * 4 + jno) * 9;
      blk[0] = tmp; blk[4] = tmp; blk[8] = tmp;
    }
  for (ino = 0; ino < 4; ino++)
    for (idim = 0; idim < 3; idim++) {
      double g = 0;
      for (jno = 0; jno < 4; jno++)
        for (jdim = 0; jdim < 3; jdim++) g += ddW[(ino * 4 + jno) * 9 + idim * 3 + jdim] * c[jno * 3 + jdim];
      dW[ino * 3 + idim] = g;
      w += g * c[ino * 3 + idim];
    }
  *W = w;
}

/* WdWddW_Contact, cloth_internal_physics.h:243-280 */
void orc_contact(const double* c, double stiff, double clearance, int col_kind, const double* col_par, double* W,
                 double* dW, double* ddW) {
  double pd, n[3];
  int i, j;
  collider_eval(col_kind, col_par, c, &pd, n);
  pd += clearance;
  if (pd < 0) {
    *W = 0;
    for (i = 0; i < 3; i++) dW[i] = 0;
    for (i = 0; i < 9; i++) ddW[i] = 0;
    return;
  }
  *W = 0.5 * stiff * pd * pd;
  for (i = 0; i < 3; i++) dW[i] = -stiff * pd * n[i];
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) ddW[i * 3 + j] = stiff * n[i] * n[j];
}

/* ------------------------------------------------------------------------------------------------ block CRS */
/* CMatrixSquareSparse::Mearge, matrix_square_sparse.cpp:88-138.  emat: [n][n] blocks of 9; local-diagonal blocks go
 * to valDia, the others to the crs slot of (row, col) and are dropped when the pattern has no such slot (:122). */
static void merge_blocks(int n, const int* ids, const double* emat, const int* rowptr, const int* colidx,
                         double* valDia, double* valCrs, int* scratch) {
  int a, b, k, j;
  for (a = 0; a < n; a++) {
    const int row = ids[a];
    for (j = rowptr[row]; j < rowptr[row + 1]; j++) scratch[colidx[j]] = j;
    for (b = 0; b < n; b++) {
      const double* src = emat + (a * n + b) * 9;
      if (a == b) {
        for (k = 0; k < 9; k++) valDia[row * 9 + k] += src[k];
      } else {
        const int slot = scratch[ids[b]];
        if (slot == -1) continue;
        for (k = 0; k < 9; k++) valCrs[slot * 9 + k] += src[k];
      }
    }
    for (j = rowptr[row]; j < rowptr[row + 1]; j++) scratch[colidx[j]] = -1;
  }
}

/* AddWdWddW_Cloth / AddWdW_Cloth, solve_internal_sparse.h:19-76 / 81-132 (with_hessian = 0 for the latter) */
static void add_cloth(int np, int ntri, int nquad, const int* tri, const int* quad, const double* xyz0,
                      const double* xyz, double lambda, double myu, double stiff_bend, double* dW, int with_hessian,
                      const int* rowptr, const int* colidx, double* valDia, double* valCrs) {
  int* scratch = NULL;
  int it, iq, a, i;
  if (with_hessian) {
    scratch = (int*)malloc(sizeof(int) * (size_t)np);
    for (i = 0; i < np; i++) scratch[i] = -1;
  }
  for (it = 0; it < ntri; it++) {
    const int* ids = tri + it * 3;
    double C[9], c[9], e, de[9], dde[81];
    for (a = 0; a < 3; a++)
      for (i = 0; i < 3; i++) { C[a * 3 + i] = xyz0[ids[a] * 3 + i]; c[a * 3 + i] = xyz[ids[a] * 3 + i]; }
    orc_cst(C, c, lambda, myu, &e, de, dde);
    for (a = 0; a < 3; a++)
      for (i = 0; i < 3; i++) dW[ids[a] * 3 + i] += de[a * 3 + i];
    if (with_hessian) merge_blocks(3, ids, dde, rowptr, colidx, valDia, valCrs, scratch);
  }
  for (iq = 0; iq < nquad; iq++) {
    const int* ids = quad + iq * 4;
    double C[12], c[12], e, de[12], dde[144];
    for (a = 0; a < 4; a++)
      for (i = 0; i < 3; i++) { C[a * 3 + i] = xyz0[ids[a] * 3 + i]; c[a * 3 + i] = xyz[ids[a] * 3 + i]; }
    orc_bend(C, c, stiff_bend, &e, de, dde);
    for (a = 0; a < 4; a++)
      for (i = 0; i < 3; i++) dW[ids[a] * 3 + i] += de[a * 3 + i];
    if (with_hessian) merge_blocks(4, ids, dde, rowptr, colidx, valDia, valCrs, scratch);
  }
  free(scratch);
}

/* AddWdW_Gravity, solve_internal_sparse.h:165-183 */
static void add_gravity(int np, const double* gravity, double mass_point, double* dW) {
  int ip;
  for (ip = 0; ip < np; ip++) {
    dW[ip * 3 + 0] -= mass_point * gravity[0];
    dW[ip * 3 + 1] -= mass_point * gravity[1];
    dW[ip * 3 + 2] -= mass_point * gravity[2];
  }
}

/* SetBoundaryCondition, matrix_square_sparse.cpp:272-302: fixed rows -> identity, fixed columns -> 0 */
static void set_boundary_condition(int np, const int* rowptr, const int* colidx, const int* bc, double* valDia,
                                   double* valCrs) {
  int i, j, k;
  const int ncrs = rowptr[np];
  for (i = 0; i < np; i++) {
    if (bc[i] == 0) continue;
    for (k = 0; k < 9; k++) valDia[i * 9 + k] = 0.0;
    valDia[i * 9 + 0] = 1.0; valDia[i * 9 + 4] = 1.0; valDia[i * 9 + 8] = 1.0;
    for (j = rowptr[i]; j < rowptr[i + 1]; j++)
      for (k = 0; k < 9; k++) valCrs[j * 9 + k] = 0.0;
  }
  for (j = 0; j < ncrs; j++) {
    if (bc[colidx[j]] == 0) continue;
    for (k = 0; k < 9; k++) valCrs[j * 9 + k] = 0.0;
  }
}

/* left/right-hand side of the backward-Euler step, solve_internal_sparse.h:281-313 */
void orc_assemble(int np, int ntri, int nquad, const int* tri, const int* quad, const int* bc, const double* xyz0,
                  const double* xyz, const double* uvw, const int* rowptr, const int* colidx, double dt, double lambda,
                  double myu, double stiff_bend, const double* gravity, double mass_point, double stiff_contact,
                  double clearance, int col_kind, const double* col_par, double* valDia, double* valCrs, double* b) {
  const int ndof = np * 3, ncrs = rowptr[np];
  int i, ip, k;
  for (i = 0; i < np * 9; i++) valDia[i] = 0.0;
  for (i = 0; i < ncrs * 9; i++) valCrs[i] = 0.0;
  for (i = 0; i < ndof; i++) b[i] = 0;
  add_cloth(np, ntri, nquad, tri, quad, xyz0, xyz, lambda, myu, stiff_bend, b, 1, rowptr, colidx, valDia, valCrs);
  for (ip = 0; ip < np; ip++) { /* AddWdWddW_Contact, :137-161 */
    double e, de[3], dde[9];
    orc_contact(xyz + ip * 3, stiff_contact, clearance, col_kind, col_par, &e, de, dde);
    for (i = 0; i < 3; i++) b[ip * 3 + i] += de[i];
    for (k = 0; k < 9; k++) valDia[ip * 9 + k] += dde[k];
  }
  add_gravity(np, gravity, mass_point, b);
  for (i = 0; i < ndof; i++) b[i] = -b[i] + mass_point * uvw[i] / dt;
  for (ip = 0; ip < np; ip++) {
    valDia[ip * 9 + 0] += mass_point / (dt * dt);
    valDia[ip * 9 + 4] += mass_point / (dt * dt);
    valDia[ip * 9 + 8] += mass_point / (dt * dt);
  }
  set_boundary_condition(np, rowptr, colidx, bc, valDia, valCrs);
  for (ip = 0; ip < np; ip++) {
    if (bc[ip] == 0) continue;
    b[ip * 3 + 0] = 0; b[ip * 3 + 1] = 0; b[ip * 3 + 2] = 0;
  }
}

/* CMatrixSquareSparse::MatVec, len == 3, matrix_square_sparse.cpp:218-242: off-diagonals in crs order, diagonal last */
void orc_matvec(int np, const int* rowptr, const int* colidx, const double* valDia, const double* valCrs, double alpha,
                const double* x, double beta, double* y) {
  int i, j;
  for (i = 0; i < np; i++) {
    const double* a;
    const double* xj;
    y[i * 3 + 0] *= beta; y[i * 3 + 1] *= beta; y[i * 3 + 2] *= beta;
    for (j = rowptr[i]; j < rowptr[i + 1]; j++) {
      a = valCrs + j * 9;
      xj = x + colidx[j] * 3;
      y[i * 3 + 0] += alpha * (a[0] * xj[0] + a[1] * xj[1] + a[2] * xj[2]);
      y[i * 3 + 1] += alpha * (a[3] * xj[0] + a[4] * xj[1] + a[5] * xj[2]);
      y[i * 3 + 2] += alpha * (a[6] * xj[0] + a[7] * xj[1] + a[8] * xj[2]);
    }
    a = valDia + i * 9;
    xj = x + i * 3;
    y[i * 3 + 0] += alpha * (a[0] * xj[0] + a[1] * xj[1] + a[2] * xj[2]);
    y[i * 3 + 1] += alpha * (a[3] * xj[0] + a[4] * xj[1] + a[5] * xj[2]);
    y[i * 3 + 2] += alpha * (a[6] * xj[0] + a[7] * xj[1] + a[8] * xj[2]);
  }
}

/* InnerProduct, matrix_square_sparse.cpp:313-324: serial left-to-right sum */
double orc_dot(int n, const double* a, const double* b) {
  double r = 0.0;
  int i;
  for (i = 0; i < n; i++) r += a[i] * b[i];
  return r;
}
/* AXPY, matrix_square_sparse.cpp:327-336 */
static void axpy(int n, double a, const double* x, double* y) {
  int i;
  for (i = 0; i < n; i++) y[i] += a * x[i];
}

/* Initialize_ILU0, ilu_sparse.cpp:114-134: first crs slot of each row whose column is above the diagonal */
void orc_ilu_diaind(int np, const int* rowptr, const int* colidx, int* diaind) {
  int i, j;
  for (i = 0; i < np; i++) {
    diaind[i] = rowptr[i + 1];
    for (j = rowptr[i]; j < rowptr[i + 1]; j++)
      if (colidx[j] > i) { diaind[i] = j; break; }
  }
}

/* 3x3 inverse by cofactors, ilu_sparse.cpp:77-97 */
static void inv33(double* a) {
  const double det = +a[0] * a[4] * a[8] + a[3] * a[7] * a[2] + a[6] * a[1] * a[5] - a[0] * a[7] * a[5] -
                     a[6] * a[4] * a[2] - a[3] * a[1] * a[8];
  const double inv_det = 1.0 / det;
  double t[9];
  memcpy(t, a, sizeof(t));
  a[0] = inv_det * (t[4] * t[8] - t[5] * t[7]);
  a[1] = inv_det * (t[2] * t[7] - t[1] * t[8]);
  a[2] = inv_det * (t[1] * t[5] - t[2] * t[4]);
  a[3] = inv_det * (t[5] * t[6] - t[3] * t[8]);
  a[4] = inv_det * (t[0] * t[8] - t[2] * t[6]);
  a[5] = inv_det * (t[2] * t[3] - t[0] * t[5]);
  a[6] = inv_det * (t[3] * t[7] - t[4] * t[6]);
  a[7] = inv_det * (t[1] * t[6] - t[0] * t[7]);
  a[8] = inv_det * (t[0] * t[4] - t[1] * t[3]);
}

/* SetValueILU + DoILUDecomp (len == 3), ilu_sparse.cpp:352-384, 507-569.  In place: on return valDia holds D^-1,
 * the strictly-upper blocks hold D^-1 U and the strictly-lower blocks hold L. */
void orc_ilu_factor(int np, const int* rowptr, const int* colidx, double* valDia, double* valCrs) {
  int* dia = (int*)malloc(sizeof(int) * (size_t)np);
  int* slot = (int*)malloc(sizeof(int) * (size_t)np);
  int i, j, r, nsing = 0;
  orc_ilu_diaind(np, rowptr, colidx, dia);
  for (i = 0; i < np; i++) slot[i] = -1;
  for (i = 0; i < np; i++) {
    int ik;
    double* dii = valDia + i * 9;
    double det;
    for (j = rowptr[i]; j < rowptr[i + 1]; j++) slot[colidx[j]] = j;
    for (ik = rowptr[i]; ik < dia[i]; ik++) { /* lower blocks of row i, ascending column k */
      const int k = colidx[ik];
      const double* vik = valCrs + ik * 9;
      int kj;
      for (kj = dia[k]; kj < rowptr[k + 1]; kj++) { /* upper blocks of row k (already scaled by D_k^-1) */
        const int jc = colidx[kj];
        const double* vkj = valCrs + kj * 9;
        double* vij;
        if (jc != i) {
          if (slot[jc] == -1) continue;
          vij = valCrs + slot[jc] * 9;
        } else {
          vij = dii;
        }
        for (r = 0; r < 3; r++) {
          vij[r * 3 + 0] -= vik[r * 3 + 0] * vkj[0] + vik[r * 3 + 1] * vkj[3] + vik[r * 3 + 2] * vkj[6];
          vij[r * 3 + 1] -= vik[r * 3 + 0] * vkj[1] + vik[r * 3 + 1] * vkj[4] + vik[r * 3 + 2] * vkj[7];
          vij[r * 3 + 2] -= vik[r * 3 + 0] * vkj[2] + vik[r * 3 + 1] * vkj[5] + vik[r * 3 + 2] * vkj[8];
        }
      }
    }
    det = +dii[0] * dii[4] * dii[8] + dii[3] * dii[7] * dii[2] + dii[6] * dii[1] * dii[5] - dii[0] * dii[7] * dii[5] -
          dii[6] * dii[4] * dii[2] - dii[3] * dii[1] * dii[8];
    if (fabs(det) > 1.0e-30) {
      inv33(dii);
    } else { /* :544-551: singular pivot is left as is; give up after more than 10 of them */
      nsing++;
      if (nsing > 10) break;
    }
    for (j = dia[i]; j < rowptr[i + 1]; j++) { /* U_ij <- D_i^-1 U_ij */
      double* vij = valCrs + j * 9;
      double t[9];
      memcpy(t, vij, sizeof(t));
      for (r = 0; r < 3; r++) {
        vij[r * 3 + 0] = dii[r * 3 + 0] * t[0] + dii[r * 3 + 1] * t[3] + dii[r * 3 + 2] * t[6];
        vij[r * 3 + 1] = dii[r * 3 + 0] * t[1] + dii[r * 3 + 1] * t[4] + dii[r * 3 + 2] * t[7];
        vij[r * 3 + 2] = dii[r * 3 + 0] * t[2] + dii[r * 3 + 1] * t[5] + dii[r * 3 + 2] * t[8];
      }
    }
    for (j = rowptr[i]; j < rowptr[i + 1]; j++) slot[colidx[j]] = -1;
  }
  free(dia);
  free(slot);
}

/* ForwardSubstitution + BackwardSubstitution (len == 3), ilu_sparse.cpp:185-215, 295-324 */
static void ilu_solve_dia(int np, const int* rowptr, const int* colidx, const int* dia, const double* iluDia,
                          const double* iluCrs, double* vec) {
  int i, j;
  for (i = 0; i < np; i++) {
    double t0 = vec[i * 3 + 0], t1 = vec[i * 3 + 1], t2 = vec[i * 3 + 2];
    const double* d = iluDia + i * 9;
    for (j = rowptr[i]; j < dia[i]; j++) {
      const double* a = iluCrs + j * 9;
      const double x0 = vec[colidx[j] * 3 + 0], x1 = vec[colidx[j] * 3 + 1], x2 = vec[colidx[j] * 3 + 2];
      t0 -= a[0] * x0 + a[1] * x1 + a[2] * x2;
      t1 -= a[3] * x0 + a[4] * x1 + a[5] * x2;
      t2 -= a[6] * x0 + a[7] * x1 + a[8] * x2;
    }
    vec[i * 3 + 0] = d[0] * t0 + d[1] * t1 + d[2] * t2;
    vec[i * 3 + 1] = d[3] * t0 + d[4] * t1 + d[5] * t2;
    vec[i * 3 + 2] = d[6] * t0 + d[7] * t1 + d[8] * t2;
  }
  for (i = np - 1; i >= 0; i--) {
    double t0 = vec[i * 3 + 0], t1 = vec[i * 3 + 1], t2 = vec[i * 3 + 2];
    for (j = dia[i]; j < rowptr[i + 1]; j++) {
      const double* a = iluCrs + j * 9;
      const double x0 = vec[colidx[j] * 3 + 0], x1 = vec[colidx[j] * 3 + 1], x2 = vec[colidx[j] * 3 + 2];
      t0 -= a[0] * x0 + a[1] * x1 + a[2] * x2;
      t1 -= a[3] * x0 + a[4] * x1 + a[5] * x2;
      t2 -= a[6] * x0 + a[7] * x1 + a[8] * x2;
    }
    vec[i * 3 + 0] = t0; vec[i * 3 + 1] = t1; vec[i * 3 + 2] = t2;
  }
}
void orc_ilu_solve(int np, const int* rowptr, const int* colidx, const double* iluDia, const double* iluCrs, double* vec) {
  int* dia = (int*)malloc(sizeof(int) * (size_t)np);
  orc_ilu_diaind(np, rowptr, colidx, dia);
  ilu_solve_dia(np, rowptr, colidx, dia, iluDia, iluCrs, vec);
  free(dia);
}

/* Solve_PCG, ilu_sparse.cpp:628-709.  r is destroyed (holds the final residual), x is the solution. */
void orc_pcg(int np, const int* rowptr, const int* colidx, const double* valDia, const double* valCrs,
             const double* iluDia, const double* iluCrs, double* r, double* x, double* conv_ratio, int* iteration) {
  const double tol = *conv_ratio;
  const int mx_iter = *iteration;
  const int ndof = np * 3;
  double inv_sq0, rPr, sq;
  double *Pr, *p;
  int* dia;
  int i, it;
  for (i = 0; i < ndof; i++) x[i] = 0;
  {
    const double sq0 = orc_dot(ndof, r, r);
    if (sq0 < 1.0e-30) { *conv_ratio = 0.0; *iteration = 0; return; }
    inv_sq0 = 1.0 / sq0;
  }
  Pr = (double*)calloc((size_t)ndof, sizeof(double));
  p = (double*)calloc((size_t)ndof, sizeof(double));
  dia = (int*)malloc(sizeof(int) * (size_t)np);
  orc_ilu_diaind(np, rowptr, colidx, dia);
  memcpy(Pr, r, sizeof(double) * (size_t)ndof);
  ilu_solve_dia(np, rowptr, colidx, dia, iluDia, iluCrs, Pr);
  memcpy(p, Pr, sizeof(double) * (size_t)ndof);
  rPr = orc_dot(ndof, r, Pr);
  for (it = 0; it < mx_iter; it++) {
    {
      double* Ap = Pr; /* the reference reuses the Pr storage for A p */
      double pAp, alpha;
      orc_matvec(np, rowptr, colidx, valDia, valCrs, 1.0, p, 0.0, Ap);
      pAp = orc_dot(ndof, p, Ap);
      alpha = rPr / pAp;
      axpy(ndof, -alpha, Ap, r);
      axpy(ndof, +alpha, p, x);
    }
    sq = orc_dot(ndof, r, r);
    if (sq * inv_sq0 < tol * tol) {
      *conv_ratio = sqrt(sq * inv_sq0);
      *iteration = it;
      free(Pr); free(p); free(dia);
      return;
    }
    {
      double rPr1, beta;
      for (i = 0; i < ndof; i++) Pr[i] = r[i];
      ilu_solve_dia(np, rowptr, colidx, dia, iluDia, iluCrs, Pr);
      rPr1 = orc_dot(ndof, r, Pr);
      beta = rPr1 / rPr;
      rPr = rPr1;
      for (i = 0; i < ndof; i++) p[i] = Pr[i] + beta * p[i];
    }
  }
  sq = orc_dot(ndof, r, r);
  *conv_ratio = sqrt(sq * inv_sq0);
  free(Pr); free(p); free(dia);
}

/* Solve_CG, matrix_square_sparse.cpp:338-406.  QUIRKS: the loop starts at 1 and on convergence reports the
 * *previous* residual ratio. */
void orc_cg(int np, const int* rowptr, const int* colidx, const double* valDia, const double* valCrs, double* r,
            double* x, double* conv_ratio, int* iteration) {
  const double tol = *conv_ratio;
  const int mx_iter = *iteration;
  const int ndof = np * 3;
  double sq, inv_sq0, *Ap, *p;
  int i, it;
  for (i = 0; i < ndof; i++) x[i] = 0;
  sq = orc_dot(ndof, r, r);
  if (sq < 1.0e-30) { *conv_ratio = 0.0; *iteration = 0; return; }
  inv_sq0 = 1.0 / sq;
  /* the reference's work vectors are zero-initialised std::vectors; MatVec scales its output by beta = 0 first
   * (matrix_square_sparse.cpp:218-242), so garbage here (NaN * 0) would poison the solve */
  Ap = (double*)calloc((size_t)ndof, sizeof(double));
  p = (double*)calloc((size_t)ndof, sizeof(double));
  memcpy(p, r, sizeof(double) * (size_t)ndof);
  *iteration = mx_iter;
  for (it = 1; it < mx_iter; it++) {
    double pAp, alpha, sq_new, beta;
    orc_matvec(np, rowptr, colidx, valDia, valCrs, 1.0, p, 0.0, Ap);
    pAp = orc_dot(ndof, p, Ap);
    alpha = sq / pAp;
    axpy(ndof, alpha, p, x);
    axpy(ndof, -alpha, Ap, r);
    sq_new = orc_dot(ndof, r, r);
    if (sq_new * inv_sq0 < tol * tol) {
      *conv_ratio = sqrt(sq * inv_sq0);
      *iteration = it;
      break;
    }
    beta = sq_new / sq;
    sq = sq_new;
    for (i = 0; i < ndof; i++) p[i] = r[i] + beta * p[i];
  }
  free(Ap); free(p);
}

/* MakeNormal, self_contact_sparse/main.cpp:183-211 (visualisation hand-off, not part of the timed step): unit triangle
 * normals (UnitNormalAreaTri3D, utility.h:58-73) scattered to the three corners in triangle order, then normalised */
void orc_make_normal(int np, int ntri, const int* tri, const double* xyz, double* normal) {
  int t, a, k, ip;
  for (k = 0; k < np * 3; k++) normal[k] = 0;
  for (t = 0; t < ntri; t++) {
    double n[3], area, invlen;
    tri_normal(n, xyz + tri[t * 3] * 3, xyz + tri[t * 3 + 1] * 3, xyz + tri[t * 3 + 2] * 3);
    area = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]) * 0.5;
    invlen = 0.5 / area;
    n[0] *= invlen; n[1] *= invlen; n[2] *= invlen;
    for (a = 0; a < 3; a++)
      for (k = 0; k < 3; k++) normal[tri[t * 3 + a] * 3 + k] += n[k];
  }
  for (ip = 0; ip < np; ip++) {
    const double sq = normal[ip * 3] * normal[ip * 3] + normal[ip * 3 + 1] * normal[ip * 3 + 1] + normal[ip * 3 + 2] * normal[ip * 3 + 2];
    const double inv = 1.0 / sqrt(sq);
    normal[ip * 3] *= inv; normal[ip * 3 + 1] *= inv; normal[ip * 3 + 2] *= inv;
  }
}

/* UpdateIntermidiateVelocity, solve_internal_sparse.h:329-369 */
void orc_update_intermediate_velocity(int np, int ntri, int nquad, const int* tri, const int* quad, const int* bc,
                                      const double* xyz0, const double* xyz, double* uvw, double dt, double lambda,
                                      double myu, double stiff_bend, const double* gravity, double mass_point) {
  double* dW = (double*)calloc((size_t)np * 3, sizeof(double));
  int ip;
  add_cloth(np, ntri, nquad, tri, quad, xyz0, xyz, lambda, myu, stiff_bend, dW, 0, NULL, NULL, NULL, NULL);
  add_gravity(np, gravity, mass_point, dW);
  for (ip = 0; ip < np; ip++) {
    if (bc[ip] == 0) continue;
    uvw[ip * 3 + 0] = 0; uvw[ip * 3 + 1] = 0; uvw[ip * 3 + 2] = 0;
  }
  for (ip = 0; ip < np; ip++) {
    uvw[ip * 3 + 0] += -(0.5 * dt / mass_point) * dW[ip * 3 + 0];
    uvw[ip * 3 + 1] += -(0.5 * dt / mass_point) * dW[ip * 3 + 1];
    uvw[ip * 3 + 2] += -(0.5 * dt / mass_point) * dW[ip * 3 + 2];
  }
  free(dW);
}

/* ------------------------------------------------------------------------------------------------ AABB */
typedef struct { double lo[3], hi[3]; int nonempty; } aabb;

static aabb aabb_empty(void) { aabb b; memset(&b, 0, sizeof(b)); return b; }
/* CAABB3D::AddPoint, aabb.h:102-122 */
static void aabb_add(aabb* b, v3 p, double eps) {
  const double c[3] = {p.x, p.y, p.z};
  int a;
  if (eps <= 0) return;
  if (b->nonempty) {
    for (a = 0; a < 3; a++) {
      b->lo[a] = (b->lo[a] < c[a] - eps) ? b->lo[a] : c[a] - eps;
      b->hi[a] = (b->hi[a] > c[a] + eps) ? b->hi[a] : c[a] + eps;
    }
  } else {
    b->nonempty = 1;
    for (a = 0; a < 3; a++) { b->lo[a] = c[a] - eps; b->hi[a] = c[a] + eps; }
  }
}
/* CAABB3D::IsIntersect, aabb.h:73-84: closed intervals */
static int aabb_intersect(const aabb* a, const aabb* b) {
  int k;
  if (!a->nonempty || !b->nonempty) return 0;
  for (k = 0; k < 3; k++) {
    if (a->hi[k] < b->lo[k]) return 0;
    if (a->lo[k] > b->hi[k]) return 0;
  }
  return 1;
}
/* CAABB3D::IsInside, aabb.h:137-144 */
static int aabb_inside(const aabb* b, v3 p) {
  if (!b->nonempty) return 0;
  return p.x >= b->lo[0] && p.x <= b->hi[0] && p.y >= b->lo[1] && p.y <= b->hi[1] && p.z >= b->lo[2] && p.z <= b->hi[2];
}
static aabb aabb_ld(const double* boxes, int i) {
  aabb b;
  b.lo[0] = boxes[i * 6 + 0]; b.hi[0] = boxes[i * 6 + 1];
  b.lo[1] = boxes[i * 6 + 2]; b.hi[1] = boxes[i * 6 + 3];
  b.lo[2] = boxes[i * 6 + 4]; b.hi[2] = boxes[i * 6 + 5];
  b.nonempty = 1;
  return b;
}
static void aabb_st(double* boxes, int i, const aabb* b) {
  boxes[i * 6 + 0] = b->lo[0]; boxes[i * 6 + 1] = b->hi[0];
  boxes[i * 6 + 2] = b->lo[1]; boxes[i * 6 + 3] = b->hi[1];
  boxes[i * 6 + 4] = b->lo[2]; boxes[i * 6 + 5] = b->hi[2];
}

/* BuildBoundingBoxChild_Prx / _CCD, bvh_aabb.cpp:23-61 / 64-106.  mode 0: leaves = 3 vertices padded by delta/2;
 * mode 1: leaves = 3 vertices at x and at x + dt*v, padded by 1e-10.  Internal node = union of its children
 * (CAABB3D::operator+=, aabb.h:55-72). */
static void refit_node(int i, const int* nodes, const int* tri, const double* xyz, const double* uvw, double par,
                       int mode, aabb* bb) {
  const int c0 = nodes[i * 3 + 1], c1 = nodes[i * 3 + 2];
  int a;
  if (c1 == -1) {
    const int* t = tri + c0 * 3;
    aabb b = aabb_empty();
    if (mode == 0) {
      for (a = 0; a < 3; a++) aabb_add(&b, v3_ld(xyz, t[a]), par * 0.5);
    } else {
      const double eps = 1.0e-10;
      for (a = 0; a < 3; a++) aabb_add(&b, v3_ld(xyz, t[a]), eps);
      for (a = 0; a < 3; a++) {
        const int n = t[a];
        aabb_add(&b, v3_make(xyz[n * 3 + 0] + par * uvw[n * 3 + 0], xyz[n * 3 + 1] + par * uvw[n * 3 + 1],
                             xyz[n * 3 + 2] + par * uvw[n * 3 + 2]), eps);
      }
    }
    bb[i] = b;
    return;
  }
  refit_node(c0, nodes, tri, xyz, uvw, par, mode, bb);
  refit_node(c1, nodes, tri, xyz, uvw, par, mode, bb);
  {
    aabb b = bb[c0];
    const aabb* o = &bb[c1];
    if (o->nonempty) {
      if (!b.nonempty) {
        b = *o;
      } else {
        for (a = 0; a < 3; a++) {
          b.hi[a] = (b.hi[a] > o->hi[a]) ? b.hi[a] : o->hi[a];
          b.lo[a] = (b.lo[a] < o->lo[a]) ? b.lo[a] : o->lo[a];
        }
      }
    }
    bb[i] = b;
  }
}
static void refit(int nnode, const int* nodes, const int* tri, const double* xyz, const double* uvw, double par,
                  int mode, double* boxes) {
  aabb* bb = (aabb*)calloc((size_t)nnode, sizeof(aabb));
  int i;
  refit_node(0, nodes, tri, xyz, uvw, par, mode, bb);
  for (i = 0; i < nnode; i++) aabb_st(boxes, i, &bb[i]);
  free(bb);
}
void orc_refit_prx(int nnode, const int* nodes, int ntri, const int* tri, int np, const double* xyz, double delta,
                   double* boxes) {
  (void)ntri; (void)np;
  refit(nnode, nodes, tri, xyz, NULL, delta, 0, boxes);
}
void orc_refit_ccd(int nnode, const int* nodes, int ntri, const int* tri, int np, const double* xyz, const double* uvw,
                   double dt, double* boxes) {
  (void)ntri; (void)np;
  refit(nnode, nodes, tri, xyz, uvw, dt, 1, boxes);
}

/* ------------------------------------------------------------------------------------------------ predicates */
/* DistanceFaceVertex, bvh_aabb.cpp:321-340 */
static double dist_face_vertex(v3 p0, v3 p1, v3 p2, v3 p3, double* w0, double* w1) {
  const v3 v20 = v3_sub(p0, p2), v21 = v3_sub(p1, p2);
  const double t0 = v3_dot(v20, v20), t1 = v3_dot(v21, v21), t2 = v3_dot(v20, v21);
  const double t3 = v3_dot(v20, v3_sub(p3, p2)), t4 = v3_dot(v21, v3_sub(p3, p2));
  const double det = t0 * t1 - t2 * t2;
  const double invdet = 1.0 / det;
  double w2;
  v3 pw;
  *w0 = (+t1 * t3 - t2 * t4) * invdet;
  *w1 = (-t2 * t3 + t0 * t4) * invdet;
  w2 = 1 - *w0 - *w1;
  pw = v3_add(v3_add(v3_scale(*w0, p0), v3_scale(*w1, p1)), v3_scale(w2, p2));
  return v3_len(v3_sub(pw, p3));
}

/* DistanceEdgeEdge, bvh_aabb.cpp:343-384 (parallel-edge branch :350-371) */
static double dist_edge_edge(v3 p0, v3 p1, v3 q0, v3 q1, double* rp, double* rq) {
  const v3 vp = v3_sub(p1, p0), vq = v3_sub(q1, q0);
  if (v3_len(v3_cross(vp, vq)) < 1.0e-10) {
    const v3 pq0 = v3_sub(p0, q0);
    const v3 nvp = v3_normalized(vp);
    const v3 vert = v3_sub(pq0, v3_scale(v3_dot(pq0, nvp), nvp));
    const double dist = v3_len(vert);
    const double lp0 = v3_dot(p0, nvp), lp1 = v3_dot(p1, nvp), lq0 = v3_dot(q0, nvp), lq1 = v3_dot(q1, nvp);
    const double p_min = (lp0 < lp1) ? lp0 : lp1, p_max = (lp0 > lp1) ? lp0 : lp1;
    const double q_min = (lq0 < lq1) ? lq0 : lq1, q_max = (lq0 > lq1) ? lq0 : lq1;
    double lm;
    if (p_max < q_min) lm = (p_max + q_min) * 0.5;
    else if (q_max < p_min) lm = (q_max + p_min) * 0.5;
    else if (p_max < q_max) lm = (p_max + q_min) * 0.5;
    else lm = (q_max + p_min) * 0.5;
    *rp = (lm - lp0) / (lp1 - lp0);
    *rq = (lm - lq0) / (lq1 - lq0);
    return dist;
  }
  {
    const double t0 = v3_dot(vp, vp), t1 = v3_dot(vq, vq), t2 = v3_dot(vp, vq);
    const double t3 = v3_dot(vp, v3_sub(q0, p0)), t4 = v3_dot(vq, v3_sub(q0, p0));
    const double det = t0 * t1 - t2 * t2;
    const double invdet = 1.0 / det;
    v3 pc, qc;
    *rp = (+t1 * t3 - t2 * t4) * invdet;
    *rq = (+t2 * t3 - t0 * t4) * invdet;
    pc = v3_add(p0, v3_scale(*rp, vp));
    qc = v3_add(q0, v3_scale(*rq, vq));
    return v3_len(v3_sub(pc, qc));
  }
}

/* Height, vector3d.h:254-267: signed distance of v4 from the plane of (v1,v2,v3) */
static double tet_height(v3 v1, v3 v2, v3 v3p, v3 v4) {
  double nx = (v2.y - v1.y) * (v3p.z - v1.z) - (v2.z - v1.z) * (v3p.y - v1.y);
  double ny = (v2.z - v1.z) * (v3p.x - v1.x) - (v2.x - v1.x) * (v3p.z - v1.z);
  double nz = (v2.x - v1.x) * (v3p.y - v1.y) - (v2.y - v1.y) * (v3p.x - v1.x);
  const double s = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
  nx *= s; ny *= s; nz *= s;
  return (v4.x - v1.x) * nx + (v4.y - v1.y) * ny + (v4.z - v1.z) * nz;
}

/* IsContact_FV_Proximity, bvh_aabb.cpp:392-412 (mgn is forced to 0 at :407) */
static int is_fv_prx(int i0, int i1, int i2, int i3, v3 p0, v3 p1, v3 p2, v3 p3, const aabb* bb, double delta) {
  double w0, w1, w2, dist;
  if (i3 == i0 || i3 == i1 || i3 == i2) return 0;
  if (!aabb_inside(bb, p3)) return 0;
  if (fabs(tet_height(p0, p1, p2, p3)) > delta) return 0;
  dist = dist_face_vertex(p0, p1, p2, p3, &w0, &w1);
  w2 = 1 - w0 - w1;
  if (dist > delta) return 0;
  if (w0 < -0.0 || w0 > 1 + 0.0) return 0;
  if (w1 < -0.0 || w1 > 1 + 0.0) return 0;
  if (w2 < -0.0 || w2 > 1 + 0.0) return 0;
  return 1;
}

/* IsContact_EE_Proximity, bvh_aabb.cpp:415-438 */
static int is_ee_prx(int i0, int i1, int j0, int j1, v3 p0, v3 p1, v3 q0, v3 q1, double delta) {
  double rp, rq, dist;
  v3 pm, qm;
  if (i0 == j0 || i0 == j1 || i1 == j0 || i1 == j1) return 0;
  if (q0.x + delta < p0.x && q0.x + delta < p1.x && q1.x + delta < p0.x && q1.x + delta < p1.x) return 0;
  if (q0.x - delta > p0.x && q0.x - delta > p1.x && q1.x - delta > p0.x && q1.x - delta > p1.x) return 0;
  if (q0.y + delta < p0.y && q0.y + delta < p1.y && q1.y + delta < p0.y && q1.y + delta < p1.y) return 0;
  if (q0.y - delta > p0.y && q0.y - delta > p1.y && q1.y - delta > p0.y && q1.y - delta > p1.y) return 0;
  if (q0.z + delta < p0.z && q0.z + delta < p1.z && q1.z + delta < p0.z && q1.z + delta < p1.z) return 0;
  if (q0.z - delta > p0.z && q0.z - delta > p1.z && q1.z - delta > p0.z && q1.z - delta > p1.z) return 0;
  dist = dist_edge_edge(p0, p1, q0, q1, &rp, &rq);
  if (dist > delta) return 0;
  if (rp < 0) return 0;
  if (rp > 1) return 0;
  if (rq < 0) return 0;
  if (rq > 1) return 0;
  pm = v3_add(v3_scale(1 - rp, p0), v3_scale(rp, p1));
  qm = v3_add(v3_scale(1 - rq, q0), v3_scale(rq, q1));
  if (v3_len(v3_sub(pm, qm)) > delta) return 0;
  return 1;
}

/* EvaluateCubic, bvh_aabb.cpp:563-568 */
static double cubic(double r, double k0, double k1, double k2, double k3) {
  return k0 + k1 * r + k2 * r * r + k3 * r * r * r;
}
/* FindRootCubic + BisectRangeCubicRoot, bvh_aabb.cpp:571-595.  QUIRKS: the counter is pre-decremented, so 15 means
 * 14 halvings; the sign test always uses the value at the ORIGINAL left end (v0 is never refreshed). */
static double cubic_root(double r0, double r1, double v0, double k0, double k1, double k2, double k3) {
  int n;
  for (n = 0; n < 14; n++) {
    const double r2 = 0.5 * (r0 + r1);
    const double v2 = cubic(r2, k0, k1, k2, k3);
    if (v0 * v2 < 0) r1 = r2; else r0 = r2;
  }
  return 0.5 * (r0 + r1);
}
/* FindCoplanerInterp, bvh_aabb.cpp:598-662: time in [0,1] at which 4 moving points become coplanar, or -1 */
static double coplanar_time(v3 s0, v3 s1, v3 s2, v3 s3, v3 e0, v3 e1, v3 e2, v3 e3) {
  const v3 x1 = v3_sub(s1, s0), x2 = v3_sub(s2, s0), x3 = v3_sub(s3, s0);
  const v3 v1 = v3_sub(v3_sub(e1, e0), x1), v2 = v3_sub(v3_sub(e2, e0), x2), v3v = v3_sub(v3_sub(e3, e0), x3);
  const double k0 = v3_triple(x3, x1, x2);
  const double k1 = v3_triple(v3v, x1, x2) + v3_triple(x3, v1, x2) + v3_triple(x3, x1, v2);
  const double k2 = v3_triple(v3v, v1, x2) + v3_triple(v3v, x1, v2) + v3_triple(x3, v1, v2);
  const double k3 = v3_triple(v3v, v1, v2);
  const double r0 = -0.0, r1 = +1.0;
  const double f0 = cubic(r0, k0, k1, k2, k3), f1 = cubic(r1, k0, k1, k2, k3);
  const double det = k2 * k2 - 3 * k1 * k3;
  if (fabs(k3) < 1.0e-10 && fabs(k2) > 1.0e-10) { /* quadratic */
    const double r2 = -k1 / (2 * k2);
    const double f2 = cubic(r2, k0, k1, k2, k3);
    if (r2 > 0 && r2 < 1) {
      if (f0 * f2 < 0) return cubic_root(r0, r2, f0, k0, k1, k2, k3);
      else if (f2 * f1 < 0) return cubic_root(r2, r1, f2, k0, k1, k2, k3);
    }
  }
  if (det > 0 && fabs(k3) > 1.0e-10) { /* cubic with two extrema */
    const double r3 = (-k2 - sqrt(det)) / (3 * k3);
    const double f3 = cubic(r3, k0, k1, k2, k3);
    double r4, f4;
    if (r3 > 0 && r3 < 1) {
      if (f0 * f3 < 0) return cubic_root(r0, r3, f0, k0, k1, k2, k3);
      else if (f3 * f1 < 0) return cubic_root(r3, r1, f3, k0, k1, k2, k3);
    }
    r4 = (-k2 + sqrt(det)) / (3 * k3);
    f4 = cubic(r4, k0, k1, k2, k3);
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

/* GetMinDist_LineSegPoint, vector3d.h:282-295 */
static v3 closest_on_segment(v3 p, v3 s, v3 e) {
  const v3 d = v3_sub(e, s);
  const double a = v3_dot(d, d);
  const double b = v3_dot(d, v3_sub(s, p));
  double t = -b / a;
  if (t < 0) t = 0;
  if (t > 1) t = 1;
  return v3_add(s, v3_scale(t, d));
}

/* IsContact_FV_CCD, bvh_aabb.cpp:666-719 */
static int is_fv_ccd(int i0, int i1, int i2, int i3, v3 p0, v3 p1, v3 p2, v3 p3, v3 q0, v3 q1, v3 q2, v3 q3,
                     const aabb* bb) {
  const double eps = 1.0e-10;
  double r0, r1, dist, t, w0, w1, w2;
  aabb bbp = aabb_empty();
  if (i3 == i0 || i3 == i1 || i3 == i2) return 0;
  aabb_add(&bbp, p3, eps);
  aabb_add(&bbp, q3, eps);
  if (!aabb_intersect(bb, &bbp)) return 0;
  { /* the vertex stays on one side of the start-time face plane */
    const v3 n = v3_cross(v3_sub(p1, p0), v3_sub(p2, p0));
    const double t0 = v3_dot(v3_sub(p0, p3), n), t1 = v3_dot(v3_sub(q0, q3), n);
    const double t2 = v3_dot(v3_sub(q1, q3), n), t3 = v3_dot(v3_sub(q2, q3), n);
    if (t0 * t1 > 0 && t0 * t2 > 0 && t0 * t3 > 0) return 0;
  }
  dist = dist_face_vertex(p0, p1, p2, p3, &r0, &r1);
  {
    const double vn0 = v3_len(v3_sub(p0, q0)), vn1 = v3_len(v3_sub(p1, q1)), vn2 = v3_len(v3_sub(p2, q2));
    const double vn3 = v3_len(v3_sub(p3, q3));
    double vnt = (vn0 > vn1) ? vn0 : vn1;
    double max_app, r2;
    vnt = (vn2 > vnt) ? vn2 : vnt;
    max_app = vnt + vn3;
    r2 = 1 - r0 - r1;
    if (dist > max_app) return 0;
    if (r0 < 0 || r0 > 1 || r1 < 0 || r1 > 1 || r2 < 0 || r2 > 1) {
      const double d01 = v3_len(v3_sub(closest_on_segment(p3, p0, p1), p3));
      const double d12 = v3_len(v3_sub(closest_on_segment(p3, p1, p2), p3));
      const double d20 = v3_len(v3_sub(closest_on_segment(p3, p2, p0), p3));
      if (d01 > max_app && d12 > max_app && d20 > max_app) return 0;
    }
  }
  t = coplanar_time(p0, p1, p2, p3, q0, q1, q2, q3);
  if (t < 0 || t > 1) return 0;
  {
    const v3 p0m = v3_add(v3_scale(1 - t, p0), v3_scale(t, q0));
    const v3 p1m = v3_add(v3_scale(1 - t, p1), v3_scale(t, q1));
    const v3 p2m = v3_add(v3_scale(1 - t, p2), v3_scale(t, q2));
    const v3 p3m = v3_add(v3_scale(1 - t, p3), v3_scale(t, q3));
    dist_face_vertex(p0m, p1m, p2m, p3m, &w0, &w1);
  }
  w2 = 1 - w0 - w1;
  if (w0 < 0 || w0 > 1) return 0;
  if (w1 < 0 || w1 > 1) return 0;
  if (w2 < 0 || w2 > 1) return 0;
  return 1;
}

/* IsContact_EE_CCD, bvh_aabb.cpp:723-793 (the CSAT / max-approach culls are commented out in the reference).
 * QUIRK: the distance gate is the literal 1e-2, not the clearance. */
static int is_ee_ccd(int i0, int i1, int j0, int j1, v3 p0s, v3 p1s, v3 q0s, v3 q1s, v3 p0e, v3 p1e, v3 q0e, v3 q1e) {
  const double eps = 1.0e-10;
  aabb bbq = aabb_empty(), bbp = aabb_empty();
  double t, w0, w1, dist;
  if (i0 == j0 || i0 == j1 || i1 == j0 || i1 == j1) return 0;
  aabb_add(&bbq, q0s, eps); aabb_add(&bbq, q1s, eps); aabb_add(&bbq, q0e, eps); aabb_add(&bbq, q1e, eps);
  aabb_add(&bbp, p0s, eps); aabb_add(&bbp, p1s, eps); aabb_add(&bbp, p0e, eps); aabb_add(&bbp, p1e, eps);
  if (!aabb_intersect(&bbp, &bbq)) return 0;
  t = coplanar_time(p0s, p1s, q0s, q1s, p0e, p1e, q0e, q1e);
  if (t < 0 || t > 1) return 0;
  {
    const v3 p0m = v3_add(v3_scale(1 - t, p0s), v3_scale(t, p0e));
    const v3 p1m = v3_add(v3_scale(1 - t, p1s), v3_scale(t, p1e));
    const v3 q0m = v3_add(v3_scale(1 - t, q0s), v3_scale(t, q0e));
    const v3 q1m = v3_add(v3_scale(1 - t, q1s), v3_scale(t, q1e));
    dist = dist_edge_edge(p0m, p1m, q0m, q1m, &w0, &w1);
  }
  if (w0 < 0 || w0 > 1) return 0;
  if (w1 < 0 || w1 > 1) return 0;
  if (dist > 1.0e-2) return 0;
  return 1;
}

static v3 ld4(const double* p, int k) { return v3_make(p[k * 3], p[k * 3 + 1], p[k * 3 + 2]); }
double orc_dist_fv(const double* p, double* w0, double* w1) {
  return dist_face_vertex(ld4(p, 0), ld4(p, 1), ld4(p, 2), ld4(p, 3), w0, w1);
}
double orc_dist_ee(const double* p, double* rp, double* rq) {
  return dist_edge_edge(ld4(p, 0), ld4(p, 1), ld4(p, 2), ld4(p, 3), rp, rq);
}
double orc_coplanar(const double* s, const double* e) {
  return coplanar_time(ld4(s, 0), ld4(s, 1), ld4(s, 2), ld4(s, 3), ld4(e, 0), ld4(e, 1), ld4(e, 2), ld4(e, 3));
}
int orc_is_fv_prx(const int* id, const double* p, const double* box, double delta) {
  aabb bb = aabb_ld(box, 0);
  return is_fv_prx(id[0], id[1], id[2], id[3], ld4(p, 0), ld4(p, 1), ld4(p, 2), ld4(p, 3), &bb, delta);
}
int orc_is_ee_prx(const int* id, const double* p, double delta) {
  return is_ee_prx(id[0], id[1], id[2], id[3], ld4(p, 0), ld4(p, 1), ld4(p, 2), ld4(p, 3), delta);
}
int orc_is_fv_ccd(const int* id, const double* p, const double* q, const double* box) {
  aabb bb = aabb_ld(box, 0);
  return is_fv_ccd(id[0], id[1], id[2], id[3], ld4(p, 0), ld4(p, 1), ld4(p, 2), ld4(p, 3), ld4(q, 0), ld4(q, 1),
                   ld4(q, 2), ld4(q, 3), &bb);
}
int orc_is_ee_ccd(const int* id, const double* p, const double* q) {
  return is_ee_ccd(id[0], id[1], id[2], id[3], ld4(p, 0), ld4(p, 1), ld4(p, 2), ld4(p, 3), ld4(q, 0), ld4(q, 1),
                   ld4(q, 2), ld4(q, 3));
}

/* ------------------------------------------------------------------------------------------------ contact sets */
/* CContactElement, bvh_aabb.h:65-106: canonical 4-tuple.  FV: triangle sorted ascending, vertex last.
 * EE: the edge holding the global minimum first (minimum first), the other edge ascending. */
typedef struct { int n[4]; int is_fv; int seq; } celem;

static celem celem_make(int is_fv, int j0, int j1, int j2, int j3) {
  celem e;
  e.is_fv = is_fv;
  e.seq = 0;
  if (is_fv) {
    int a = j0, b = j1, c = j2, t;
    if (a > b) { t = a; a = b; b = t; }
    if (b > c) { t = b; b = c; c = t; }
    if (a > b) { t = a; a = b; b = t; }
    e.n[0] = a; e.n[1] = b; e.n[2] = c; e.n[3] = j3;
  } else {
    int a0 = j0, a1 = j1, b0 = j2, b1 = j3, t;
    const int mA = a0 < a1 ? a0 : a1, mB = b0 < b1 ? b0 : b1;
    if (mB < mA) { t = a0; a0 = b0; b0 = t; t = a1; a1 = b1; b1 = t; } /* edge with the global minimum first */
    if (a1 < a0) { t = a0; a0 = a1; a1 = t; }
    if (b1 < b0) { t = b0; b0 = b1; b1 = t; }
    e.n[0] = a0; e.n[1] = a1; e.n[2] = b0; e.n[3] = b1;
  }
  return e;
}
/* operator<, bvh_aabb.h:96-102: lexicographic on the 4 indices, is_fv ignored */
static int celem_cmp(const void* pa, const void* pb) {
  const celem* a = (const celem*)pa;
  const celem* b = (const celem*)pb;
  int k;
  for (k = 0; k < 4; k++)
    if (a->n[k] != b->n[k]) return a->n[k] < b->n[k] ? -1 : 1;
  return a->seq < b->seq ? -1 : (a->seq > b->seq ? 1 : 0); /* insertion order: std::set keeps the first */
}

typedef struct { celem* e; int n, cap; } celem_list;
static void celem_push(celem_list* l, celem e) {
  if (l->n == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 1024;
    l->e = (celem*)realloc(l->e, sizeof(celem) * (size_t)l->cap);
  }
  e.seq = l->n;
  l->e[l->n++] = e;
}
/* emulate std::set<CContactElement>: sort by key, keep the first inserted of each key */
static int celem_finish(celem_list* l, int* out, int cap) {
  int i, n = 0;
  qsort(l->e, (size_t)l->n, sizeof(celem), celem_cmp);
  for (i = 0; i < l->n; i++) {
    if (i > 0 && memcmp(l->e[i].n, l->e[i - 1].n, sizeof(int) * 4) == 0) continue;
    if (n < cap) {
      out[n * 5 + 0] = l->e[i].n[0]; out[n * 5 + 1] = l->e[i].n[1]; out[n * 5 + 2] = l->e[i].n[2];
      out[n * 5 + 3] = l->e[i].n[3]; out[n * 5 + 4] = l->e[i].is_fv;
    }
    n++;
  }
  free(l->e);
  l->e = NULL; l->n = l->cap = 0;
  return n;
}

typedef struct {
  const int* nodes; const double* boxes; const int* tri; const double* xyz; const double* uvw;
  double dt, delta; int ccd;
  celem_list* out;      /* contact elements, or */
  int* pairs; int npairs, pair_cap; /* leaf pairs when out == NULL */
} trav;

/* the 6 face-vertex + 9 edge-edge tests of one leaf pair, bvh_aabb.cpp:476-535 (proximity) / 834-900 (CCD) */
static void leaf_pair_tests(trav* tv, int ib0, int ib1) {
  static const int EA[9] = {0, 0, 0, 1, 1, 1, 2, 2, 2}, EB[9] = {0, 1, 2, 0, 1, 2, 0, 1, 2};
  const int it = tv->nodes[ib0 * 3 + 1], jt = tv->nodes[ib1 * 3 + 1];
  const int* in = tv->tri + it * 3;
  const int* jn = tv->tri + jt * 3;
  v3 p[3], q[3], pe[3], qe[3];
  int a, k;
  for (a = 0; a < 3; a++) { p[a] = v3_ld(tv->xyz, in[a]); q[a] = v3_ld(tv->xyz, jn[a]); }
  if (tv->ccd) {
    for (a = 0; a < 3; a++) {
      const int i = in[a], j = jn[a];
      pe[a] = v3_make(tv->xyz[i * 3] + tv->dt * tv->uvw[i * 3], tv->xyz[i * 3 + 1] + tv->dt * tv->uvw[i * 3 + 1],
                      tv->xyz[i * 3 + 2] + tv->dt * tv->uvw[i * 3 + 2]);
      qe[a] = v3_make(tv->xyz[j * 3] + tv->dt * tv->uvw[j * 3], tv->xyz[j * 3 + 1] + tv->dt * tv->uvw[j * 3 + 1],
                      tv->xyz[j * 3 + 2] + tv->dt * tv->uvw[j * 3 + 2]);
    }
    {
      const aabb b0 = aabb_ld(tv->boxes, ib0), b1 = aabb_ld(tv->boxes, ib1);
      for (a = 0; a < 3; a++)
        if (is_fv_ccd(in[0], in[1], in[2], jn[a], p[0], p[1], p[2], q[a], pe[0], pe[1], pe[2], qe[a], &b0))
          celem_push(tv->out, celem_make(1, in[0], in[1], in[2], jn[a]));
      for (a = 0; a < 3; a++)
        if (is_fv_ccd(jn[0], jn[1], jn[2], in[a], q[0], q[1], q[2], p[a], qe[0], qe[1], qe[2], pe[a], &b1))
          celem_push(tv->out, celem_make(1, jn[0], jn[1], jn[2], in[a]));
    }
    for (k = 0; k < 9; k++) {
      const int a0 = EA[k], a1 = (EA[k] + 1) % 3, b0 = EB[k], b1 = (EB[k] + 1) % 3;
      if (is_ee_ccd(in[a0], in[a1], jn[b0], jn[b1], p[a0], p[a1], q[b0], q[b1], pe[a0], pe[a1], qe[b0], qe[b1]))
        celem_push(tv->out, celem_make(0, in[a0], in[a1], jn[b0], jn[b1]));
    }
  } else {
    /* QUIRK (bvh_aabb.cpp:490-505): the vertex is tested against the box of BVH node number <triangle index>,
     * not against the leaf's own box */
    const aabb b0 = aabb_ld(tv->boxes, it), b1 = aabb_ld(tv->boxes, jt);
    for (a = 0; a < 3; a++)
      if (is_fv_prx(in[0], in[1], in[2], jn[a], p[0], p[1], p[2], q[a], &b0, tv->delta))
        celem_push(tv->out, celem_make(1, in[0], in[1], in[2], jn[a]));
    for (a = 0; a < 3; a++)
      if (is_fv_prx(jn[0], jn[1], jn[2], in[a], q[0], q[1], q[2], p[a], &b1, tv->delta))
        celem_push(tv->out, celem_make(1, jn[0], jn[1], jn[2], in[a]));
    for (k = 0; k < 9; k++) {
      const int a0 = EA[k], a1 = (EA[k] + 1) % 3, b0e = EB[k], b1e = (EB[k] + 1) % 3;
      if (is_ee_prx(in[a0], in[a1], jn[b0e], jn[b1e], p[a0], p[a1], q[b0e], q[b1e], tv->delta))
        celem_push(tv->out, celem_make(0, in[a0], in[a1], jn[b0e], jn[b1e]));
    }
  }
}

/* dual-tree descent, bvh_aabb.cpp:441-475 / 797-833: child order (0,0),(1,0),(0,1),(1,1) */
static void trav_pair(trav* tv, int i0, int i1) {
  const aabb b0 = aabb_ld(tv->boxes, i0), b1 = aabb_ld(tv->boxes, i1);
  const int c00 = tv->nodes[i0 * 3 + 1], c01 = tv->nodes[i0 * 3 + 2];
  const int c10 = tv->nodes[i1 * 3 + 1], c11 = tv->nodes[i1 * 3 + 2];
  const int leaf0 = c01 == -1, leaf1 = c11 == -1;
  if (!aabb_intersect(&b0, &b1)) return;
  if (!leaf0 && !leaf1) {
    trav_pair(tv, c00, c10); trav_pair(tv, c01, c10); trav_pair(tv, c00, c11); trav_pair(tv, c01, c11);
  } else if (!leaf0 && leaf1) {
    trav_pair(tv, c00, i1); trav_pair(tv, c01, i1);
  } else if (leaf0 && !leaf1) {
    trav_pair(tv, i0, c10); trav_pair(tv, i0, c11);
  } else if (tv->out) {
    leaf_pair_tests(tv, i0, i1);
  } else {
    if (tv->npairs < tv->pair_cap) { tv->pairs[tv->npairs * 2] = i0; tv->pairs[tv->npairs * 2 + 1] = i1; }
    tv->npairs++;
  }
}
/* self traversal, bvh_aabb.cpp:540-557 / 904-923 */
static void trav_self(trav* tv, int i) {
  const int c0 = tv->nodes[i * 3 + 1], c1 = tv->nodes[i * 3 + 2];
  if (c1 == -1) return;
  trav_pair(tv, c0, c1);
  trav_self(tv, c0);
  trav_self(tv, c1);
}

int orc_leaf_pairs(int nnode, const int* nodes, const double* boxes, int* pairs, int cap) {
  trav tv;
  (void)nnode;
  memset(&tv, 0, sizeof(tv));
  tv.nodes = nodes; tv.boxes = boxes; tv.pairs = pairs; tv.pair_cap = cap;
  trav_self(&tv, 0);
  return tv.npairs;
}
int orc_contacts_prx(int nnode, const int* nodes, const double* boxes, int ntri, const int* tri, int np,
                     const double* xyz, double delta, int* elems, int cap) {
  trav tv;
  celem_list l = {NULL, 0, 0};
  (void)nnode; (void)ntri; (void)np;
  memset(&tv, 0, sizeof(tv));
  tv.nodes = nodes; tv.boxes = boxes; tv.tri = tri; tv.xyz = xyz; tv.delta = delta; tv.ccd = 0; tv.out = &l;
  trav_self(&tv, 0);
  return celem_finish(&l, elems, cap);
}
int orc_contacts_ccd(int nnode, const int* nodes, const double* boxes, int ntri, const int* tri, int np,
                     const double* xyz, const double* uvw, double dt, double delta, int* elems, int cap) {
  trav tv;
  celem_list l = {NULL, 0, 0};
  (void)nnode; (void)ntri; (void)np;
  memset(&tv, 0, sizeof(tv));
  tv.nodes = nodes; tv.boxes = boxes; tv.tri = tri; tv.xyz = xyz; tv.uvw = uvw; tv.dt = dt; tv.delta = delta;
  tv.ccd = 1; tv.out = &l;
  trav_self(&tv, 0);
  return celem_finish(&l, elems, cap);
}

/* ------------------------------------------------------------------------------------------------ response */
/* v[i] += n * a * b, evaluated as (n*a)*b like `-norm.x*imp_mod*w0` in the reference */
static void add3(double* v, int i, v3 n, double a, double b) {
  v[i * 3 + 0] += n.x * a * b;
  v[i * 3 + 1] += n.y * a * b;
  v[i * 3 + 2] += n.z * a * b;
}
/* v[i] += n * a */
static void add3s(double* v, int i, v3 n, double a) {
  v[i * 3 + 0] += n.x * a;
  v[i * 3 + 1] += n.y * a;
  v[i * 3 + 2] += n.z * a;
}

/* SelfCollisionImpulse_Proximity, self_collision_cloth.cpp:20-112: sequential (Gauss-Seidel) over the elements */
void orc_impulse_prx(int np, double* uvw, double delta, double stiffness, double dt, double mass, const double* xyz,
                     int nelem, const int* elems) {
  int ie;
  (void)np;
  for (ie = 0; ie < nelem; ie++) {
    const int* n = elems + ie * 5;
    const v3 p0 = v3_ld(xyz, n[0]), p1 = v3_ld(xyz, n[1]), p2 = v3_ld(xyz, n[2]), p3 = v3_ld(xyz, n[3]);
    const v3 v0 = v3_ld(uvw, n[0]), v1 = v3_ld(uvw, n[1]), v2 = v3_ld(uvw, n[2]), v3v = v3_ld(uvw, n[3]);
    if (n[4]) {
      double w0, w1, w2, p_depth, rel_v, imp_el, imp_ie, imp_min, imp_mod;
      v3 pc, norm;
      {
        const double dist = dist_face_vertex(p0, p1, p2, p3, &w0, &w1);
        if (w0 < 0 || w0 > 1) continue;
        if (w1 < 0 || w1 > 1) continue;
        if (dist > delta) continue;
      }
      w2 = 1.0 - w0 - w1;
      pc = v3_add(v3_add(v3_scale(w0, p0), v3_scale(w1, p1)), v3_scale(w2, p2));
      norm = v3_normalized(v3_sub(p3, pc));
      p_depth = delta - v3_dot(v3_sub(p3, pc), norm);
      rel_v = v3_dot(v3_sub(v3_sub(v3_sub(v3v, v3_scale(w0, v0)), v3_scale(w1, v1)), v3_scale(w2, v2)), norm);
      if (rel_v > 0.1 * p_depth / dt) continue;
      imp_el = dt * stiffness * p_depth;
      imp_ie = mass * (0.1 * p_depth / dt - rel_v);
      imp_min = (imp_el < imp_ie) ? imp_el : imp_ie;
      imp_mod = 2 * imp_min / (1 + w0 * w0 + w1 * w1 + w2 * w2);
      imp_mod /= mass;
      imp_mod *= 0.25;
      add3(uvw, n[0], v3_neg(norm), imp_mod, w0);
      add3(uvw, n[1], v3_neg(norm), imp_mod, w1);
      add3(uvw, n[2], v3_neg(norm), imp_mod, w2);
      add3s(uvw, n[3], norm, imp_mod);
    } else {
      double w01, w23, p_depth, rel_v, imp_el, imp_ie, imp_min, imp_mod;
      v3 c01, c23, norm;
      {
        const double dist = dist_edge_edge(p0, p1, p2, p3, &w01, &w23);
        if (w01 < 0 || w01 > 1) continue;
        if (w23 < 0 || w23 > 1) continue;
        if (dist > delta) continue;
      }
      c01 = v3_add(v3_scale(1 - w01, p0), v3_scale(w01, p1));
      c23 = v3_add(v3_scale(1 - w23, p2), v3_scale(w23, p3));
      norm = v3_normalized(v3_sub(c23, c01));
      p_depth = delta - v3_len(v3_sub(c23, c01));
      rel_v = v3_dot(v3_sub(v3_sub(v3_add(v3_scale(1 - w23, v2), v3_scale(w23, v3v)), v3_scale(1 - w01, v0)),
                            v3_scale(w01, v1)), norm);
      if (rel_v > 0.1 * p_depth / dt) continue;
      imp_el = dt * stiffness * p_depth;
      imp_ie = mass * (0.1 * p_depth / dt - rel_v);
      imp_min = (imp_el < imp_ie) ? imp_el : imp_ie;
      imp_mod = 2 * imp_min / (w01 * w01 + (1 - w01) * (1 - w01) + w23 * w23 + (1 - w23) * (1 - w23));
      imp_mod /= mass;
      imp_mod *= 0.25;
      add3(uvw, n[0], v3_neg(norm), imp_mod, 1 - w01);
      add3(uvw, n[1], v3_neg(norm), imp_mod, w01);
      add3(uvw, n[2], norm, imp_mod, 1 - w23);
      add3(uvw, n[3], norm, imp_mod, w23);
    }
  }
}

/* SelfCollisionImpulse_CCD, self_collision_cloth.cpp:116-212.
 * QUIRK (:141): the impact time is re-solved with the UN-scaled velocity (end points p + v, not p + dt*v). */
void orc_impulse_ccd(int np, double* uvw, double delta, double stiffness, double dt, double mass, const double* xyz,
                     int nelem, const int* elems) {
  int ie;
  (void)np; (void)stiffness;
  for (ie = 0; ie < nelem; ie++) {
    const int* n = elems + ie * 5;
    const v3 p0 = v3_ld(xyz, n[0]), p1 = v3_ld(xyz, n[1]), p2 = v3_ld(xyz, n[2]), p3 = v3_ld(xyz, n[3]);
    const v3 v0 = v3_ld(uvw, n[0]), v1 = v3_ld(uvw, n[1]), v2 = v3_ld(uvw, n[2]), v3v = v3_ld(uvw, n[3]);
    const double t = coplanar_time(p0, p1, p2, p3, v3_add(p0, v0), v3_add(p1, v1), v3_add(p2, v2), v3_add(p3, v3v));
    v3 p0m, p1m, p2m, p3m;
    if (t < 0 || t > 1) continue;
    p0m = v3_add(p0, v3_scale(t, v0));
    p1m = v3_add(p1, v3_scale(t, v1));
    p2m = v3_add(p2, v3_scale(t, v2));
    p3m = v3_add(p3, v3_scale(t, v3v));
    if (n[4]) {
      double w0, w1, w2, rel_v, imp, imp_mod;
      v3 pc, norm;
      {
        const double dist = dist_face_vertex(p0m, p1m, p2m, p3m, &w0, &w1);
        if (w0 < 0 || w0 > 1) continue;
        if (w1 < 0 || w1 > 1) continue;
        if (dist > delta) continue;
      }
      w2 = 1.0 - w0 - w1;
      pc = v3_add(v3_add(v3_scale(w0, p0), v3_scale(w1, p1)), v3_scale(w2, p2));
      norm = v3_normalized(v3_sub(p3, pc));
      rel_v = v3_dot(v3_sub(v3_sub(v3_sub(v3v, v3_scale(w0, v0)), v3_scale(w1, v1)), v3_scale(w2, v2)), norm);
      if (rel_v > 0.1 * delta / dt) continue;
      imp = mass * (0.1 * delta / dt - rel_v);
      imp_mod = 2 * imp / (1.0 + w0 * w0 + w1 * w1 + w2 * w2);
      imp_mod /= mass;
      imp_mod *= 0.1;
      add3(uvw, n[0], v3_neg(norm), imp_mod, w0);
      add3(uvw, n[1], v3_neg(norm), imp_mod, w1);
      add3(uvw, n[2], v3_neg(norm), imp_mod, w2);
      add3s(uvw, n[3], norm, imp_mod);
    } else {
      double w01, w23, rel_v, imp, imp_mod;
      v3 c01, c23, norm;
      {
        const double dist = dist_edge_edge(p0m, p1m, p2m, p3m, &w01, &w23);
        if (w01 < 0 || w01 > 1) continue;
        if (w23 < 0 || w23 > 1) continue;
        if (dist > delta) continue;
      }
      c01 = v3_add(v3_scale(1 - w01, p0), v3_scale(w01, p1));
      c23 = v3_add(v3_scale(1 - w23, p2), v3_scale(w23, p3));
      norm = v3_normalized(v3_sub(c23, c01));
      rel_v = v3_dot(v3_sub(v3_sub(v3_add(v3_scale(1 - w23, v2), v3_scale(w23, v3v)), v3_scale(1 - w01, v0)),
                            v3_scale(w01, v1)), norm);
      if (rel_v > 0.1 * delta / dt) continue;
      imp = mass * (0.1 * delta / dt - rel_v);
      imp_mod = 2 * imp / (w01 * w01 + (1 - w01) * (1 - w01) + w23 * w23 + (1 - w23) * (1 - w23));
      imp_mod /= mass;
      imp_mod *= 0.1;
      add3(uvw, n[0], v3_neg(norm), imp_mod, 1 - w01);
      add3(uvw, n[1], v3_neg(norm), imp_mod, w01);
      add3(uvw, n[2], norm, imp_mod, 1 - w23);
      add3(uvw, n[3], norm, imp_mod, w23);
    }
  }
}

/* ------------------------------------------------------------------------------------------------ impact zones */
/* Zones are disjoint vertex sets (a vertex already in a zone pulls every later element that touches it into that
 * zone), so membership is kept as zone_of[vertex]; zone numbering follows the reference: a new zone is appended,
 * a merge keeps the untouched zones in order and appends the merged one (self_collision_cloth.cpp:220-274). */
typedef struct { int np; int nzone; int* zone_of; } riz_state;

void* orc_riz_new(void) { return calloc(1, sizeof(riz_state)); }
void orc_riz_free(void* h) {
  riz_state* r = (riz_state*)h;
  free(r->zone_of);
  free(r);
}
void orc_riz_make(void* h, int nelem, const int* elems, int np, const int* edge_index, const int* edge_array) {
  riz_state* r = (riz_state*)h;
  int ie, i, k;
  if (r->zone_of == NULL) {
    r->np = np;
    r->zone_of = (int*)malloc(sizeof(int) * (size_t)np);
    for (i = 0; i < np; i++) r->zone_of[i] = -1;
  }
  for (ie = 0; ie < nelem; ie++) {
    const int* n = elems + ie * 5;
    int touched[64], ntouched = 0; /* <= 4 * (1 + valence) candidates, deduplicated below */
    int* dyn = NULL;
    int* inc = touched;
    int cap = 64, target;
    for (i = 0; i < 4; i++) {
      const int ino = n[i];
      for (k = edge_index[ino] - 1; k < edge_index[ino + 1]; k++) {
        const int z = r->zone_of[k < edge_index[ino] ? ino : edge_array[k]];
        int seen = 0, j;
        if (z < 0) continue;
        for (j = 0; j < ntouched; j++) if (inc[j] == z) { seen = 1; break; }
        if (seen) continue;
        if (ntouched == cap) {
          int* nd = (int*)malloc(sizeof(int) * (size_t)cap * 2);
          memcpy(nd, inc, sizeof(int) * (size_t)cap);
          free(dyn);
          dyn = nd; inc = nd; cap *= 2;
        }
        inc[ntouched++] = z;
      }
    }
    if (ntouched == 0) {
      target = r->nzone++;
    } else if (ntouched == 1) {
      target = inc[0];
    } else {
      /* renumber: untouched zones keep their relative order, the merged zone becomes the last one */
      int* remap = (int*)malloc(sizeof(int) * (size_t)r->nzone);
      int next = 0, z;
      for (z = 0; z < r->nzone; z++) {
        int hit = 0, j;
        for (j = 0; j < ntouched; j++) if (inc[j] == z) { hit = 1; break; }
        remap[z] = hit ? -2 : next++;
      }
      target = next;
      for (z = 0; z < r->nzone; z++) if (remap[z] == -2) remap[z] = target;
      for (i = 0; i < r->np; i++) if (r->zone_of[i] >= 0) r->zone_of[i] = remap[r->zone_of[i]];
      r->nzone = next + 1;
      free(remap);
    }
    for (i = 0; i < 4; i++) r->zone_of[n[i]] = target;
    free(dyn);
  }
}
int orc_riz_get(void* h, int np, int* zone_of_node) {
  riz_state* r = (riz_state*)h;
  int i;
  for (i = 0; i < np; i++) zone_of_node[i] = (r->zone_of && i < r->np) ? r->zone_of[i] : -1;
  return r->nzone;
}

/* ApplyRigidImpactZone, self_collision_cloth.cpp:298-348: each zone moves rigidly (mean velocity + rotation);
 * sums run over the zone's vertices in ascending index.
 * QUIRK (:327-329): the "inertia tensor" diagonal uses v.v (velocity) where q.q (position) is meant. */
void orc_riz_apply(void* h, int np, double* uvw, const double* xyz, const double* uvw0) {
  riz_state* r = (riz_state*)h;
  int z, i;
  if (!r->zone_of) return;
  for (z = 0; z < r->nzone; z++) {
    v3 gc = v3_make(0, 0, 0), av = v3_make(0, 0, 0), L = v3_make(0, 0, 0), omg;
    double I[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, Iinv[9];
    int cnt = 0;
    for (i = 0; i < np; i++) {
      if (r->zone_of[i] != z) continue;
      gc = v3_add(gc, v3_ld(xyz, i));
      av = v3_add(av, v3_ld(uvw0, i));
      cnt++;
    }
    if (cnt == 0) continue;
    gc = v3_div(gc, (double)cnt);
    av = v3_div(av, (double)cnt);
    for (i = 0; i < np; i++) {
      v3 p, v, q;
      double vv;
      if (r->zone_of[i] != z) continue;
      p = v3_ld(xyz, i);
      v = v3_ld(uvw0, i);
      L = v3_add(L, v3_cross(v3_sub(p, gc), v3_sub(v, av)));
      q = v3_sub(p, gc);
      vv = v3_dot(v, v);
      I[0] += vv - q.x * q.x; I[1] += -q.x * q.y;     I[2] += -q.x * q.z;
      I[3] += -q.y * q.x;     I[4] += vv - q.y * q.y; I[5] += -q.y * q.z;
      I[6] += -q.z * q.x;     I[7] += -q.z * q.y;     I[8] += vv - q.z * q.z;
    }
    { /* CalcInvMat3, self_collision_cloth.cpp:278-296 */
      const double det = +I[0] * I[4] * I[8] + I[3] * I[7] * I[2] + I[6] * I[1] * I[5] - I[0] * I[7] * I[5] -
                         I[6] * I[4] * I[2] - I[3] * I[1] * I[8];
      const double inv_det = 1.0 / det;
      Iinv[0] = inv_det * (I[4] * I[8] - I[5] * I[7]);
      Iinv[1] = inv_det * (I[2] * I[7] - I[1] * I[8]);
      Iinv[2] = inv_det * (I[1] * I[5] - I[2] * I[4]);
      Iinv[3] = inv_det * (I[5] * I[6] - I[3] * I[8]);
      Iinv[4] = inv_det * (I[0] * I[8] - I[2] * I[6]);
      Iinv[5] = inv_det * (I[2] * I[3] - I[0] * I[5]);
      Iinv[6] = inv_det * (I[3] * I[7] - I[4] * I[6]);
      Iinv[7] = inv_det * (I[1] * I[6] - I[0] * I[7]);
      Iinv[8] = inv_det * (I[0] * I[4] - I[1] * I[3]);
    }
    omg.x = Iinv[0] * L.x + Iinv[1] * L.y + Iinv[2] * L.z;
    omg.y = Iinv[3] * L.x + Iinv[4] * L.y + Iinv[5] * L.z;
    omg.z = Iinv[6] * L.x + Iinv[7] * L.y + Iinv[8] * L.z;
    for (i = 0; i < np; i++) {
      v3 rot;
      if (r->zone_of[i] != z) continue;
      rot = v3_neg(v3_cross(v3_sub(v3_ld(xyz, i), gc), omg));
      uvw[i * 3 + 0] = av.x + rot.x;
      uvw[i * 3 + 1] = av.y + rot.y;
      uvw[i * 3 + 2] = av.z + rot.z;
    }
  }
}

/* ------------------------------------------------------------------------------------------------ setup helpers */
/* CJaggedArray::SetEdgeOfElem (is_dia = false) + optional Sort, jagged_array.h:158-207, 71-90 */
int orc_edge_of_elem(const int* elem, int nelem, int nnoel, int nnode, int do_sort, int* index, int* array, int cap) {
  int* n2e_idx = (int*)calloc((size_t)nnode + 1, sizeof(int));
  int* n2e = (int*)malloc(sizeof(int) * (size_t)nelem * (size_t)nnoel);
  int* fill = (int*)malloc(sizeof(int) * (size_t)nnode);
  int* flag = (int*)malloc(sizeof(int) * (size_t)nnode);
  int e, a, i, k, pass, nnz = 0;
  for (e = 0; e < nelem * nnoel; e++) n2e_idx[elem[e] + 1]++;
  for (i = 0; i < nnode; i++) n2e_idx[i + 1] += n2e_idx[i];
  for (i = 0; i < nnode; i++) fill[i] = n2e_idx[i];
  for (e = 0; e < nelem; e++)
    for (a = 0; a < nnoel; a++) n2e[fill[elem[e * nnoel + a]]++] = e;
  for (pass = 0; pass < 2; pass++) {
    for (i = 0; i < nnode; i++) flag[i] = -1;
    nnz = 0;
    for (i = 0; i < nnode; i++) {
      if (pass == 0) index[i] = nnz;
      flag[i] = i;
      for (k = n2e_idx[i]; k < n2e_idx[i + 1]; k++)
        for (a = 0; a < nnoel; a++) {
          const int j = elem[n2e[k] * nnoel + a];
          if (flag[j] == i) continue;
          flag[j] = i;
          if (pass == 1 && nnz < cap) array[nnz] = j;
          nnz++;
        }
    }
    if (pass == 0) index[nnode] = nnz;
  }
  if (do_sort) {
    for (i = 0; i < nnode; i++) { /* insertion sort per row (the reference bubble-sorts; same result) */
      const int s = index[i], t = index[i + 1] < cap ? index[i + 1] : cap;
      for (a = s + 1; a < t; a++) {
        const int v = array[a];
        k = a - 1;
        while (k >= s && array[k] > v) { array[k + 1] = array[k]; k--; }
        array[k + 1] = v;
      }
    }
  }
  free(n2e_idx); free(n2e); free(fill); free(flag);
  return nnz;
}

/* ------------------------------------------------------------------------------------------------ whole scene */
typedef struct {
  int np, ntri, nquad, nnode, ncrs;
  double *xyz0, *xyz, *uvw;
  int *bc, *tri, *quad, *nodes, *rowptr, *colidx, *edge_index, *edge_array;
  double *valDia, *valCrs, *iluDia, *iluCrs, *boxes;
} sim_state;

static void* dup_mem(const void* p, size_t n) {
  void* q = malloc(n ? n : 1);
  if (p && n) memcpy(q, p, n);
  return q;
}

/* init sequence of self_contact_sparse/main.cpp:561-583 (BVH topology is an input) */
void* orc_sim_new(int np, int ntri, int nquad, const int* tri, const int* quad, const int* bc, const double* xyz0,
                  int nnode, const int* nodes) {
  sim_state* s = (sim_state*)calloc(1, sizeof(sim_state));
  int cap;
  s->np = np; s->ntri = ntri; s->nquad = nquad; s->nnode = nnode;
  s->xyz0 = (double*)dup_mem(xyz0, sizeof(double) * (size_t)np * 3);
  s->xyz = (double*)dup_mem(xyz0, sizeof(double) * (size_t)np * 3);
  s->uvw = (double*)calloc((size_t)np * 3, sizeof(double));
  s->bc = (int*)dup_mem(bc, sizeof(int) * (size_t)np);
  s->tri = (int*)dup_mem(tri, sizeof(int) * (size_t)ntri * 3);
  s->quad = (int*)dup_mem(quad, sizeof(int) * (size_t)nquad * 4);
  s->nodes = (int*)dup_mem(nodes, sizeof(int) * (size_t)nnode * 3);
  cap = ntri * 9;
  s->edge_index = (int*)malloc(sizeof(int) * ((size_t)np + 1));
  s->edge_array = (int*)malloc(sizeof(int) * (size_t)cap);
  orc_edge_of_elem(tri, ntri, 3, np, 0, s->edge_index, s->edge_array, cap);
  cap = nquad * 16;
  s->rowptr = (int*)malloc(sizeof(int) * ((size_t)np + 1));
  s->colidx = (int*)malloc(sizeof(int) * (size_t)cap);
  s->ncrs = orc_edge_of_elem(quad, nquad, 4, np, 1, s->rowptr, s->colidx, cap);
  s->valDia = (double*)calloc((size_t)np * 9, sizeof(double));
  s->valCrs = (double*)calloc((size_t)s->ncrs * 9 + 1, sizeof(double));
  s->iluDia = (double*)calloc((size_t)np * 9, sizeof(double));
  s->iluCrs = (double*)calloc((size_t)s->ncrs * 9 + 1, sizeof(double));
  s->boxes = (double*)calloc((size_t)nnode * 6 + 1, sizeof(double));
  return s;
}
void orc_sim_free(void* h) {
  sim_state* s = (sim_state*)h;
  free(s->xyz0); free(s->xyz); free(s->uvw); free(s->bc); free(s->tri); free(s->quad); free(s->nodes);
  free(s->rowptr); free(s->colidx); free(s->edge_index); free(s->edge_array);
  free(s->valDia); free(s->valCrs); free(s->iluDia); free(s->iluCrs); free(s->boxes);
  free(s);
}
void orc_sim_set_state(void* h, const double* xyz, const double* uvw) {
  sim_state* s = (sim_state*)h;
  memcpy(s->xyz, xyz, sizeof(double) * (size_t)s->np * 3);
  memcpy(s->uvw, uvw, sizeof(double) * (size_t)s->np * 3);
}
void orc_sim_get_state(void* h, double* xyz, double* uvw) {
  sim_state* s = (sim_state*)h;
  memcpy(xyz, s->xyz, sizeof(double) * (size_t)s->np * 3);
  memcpy(uvw, s->uvw, sizeof(double) * (size_t)s->np * 3);
}

/* StepTime_InternalDynamicsILU, solve_internal_sparse.h:255-326 */
void orc_sim_step_internal(void* h, double dt, double lambda, double myu, double stiff_bend, const double* gravity,
                           double mass_point, double stiff_contact, double clearance, int col_kind,
                           const double* col_par, double* conv_ratio, int* iteration) {
  sim_state* s = (sim_state*)h;
  const int ndof = s->np * 3;
  double* b = (double*)malloc(sizeof(double) * (size_t)ndof);
  double* x = (double*)malloc(sizeof(double) * (size_t)ndof);
  double cr = 1.0e-4;
  int it = 100, i;
  orc_assemble(s->np, s->ntri, s->nquad, s->tri, s->quad, s->bc, s->xyz0, s->xyz, s->uvw, s->rowptr, s->colidx, dt,
               lambda, myu, stiff_bend, gravity, mass_point, stiff_contact, clearance, col_kind, col_par, s->valDia,
               s->valCrs, b);
  memcpy(s->iluDia, s->valDia, sizeof(double) * (size_t)s->np * 9);
  memcpy(s->iluCrs, s->valCrs, sizeof(double) * (size_t)s->ncrs * 9);
  orc_ilu_factor(s->np, s->rowptr, s->colidx, s->iluDia, s->iluCrs);
  orc_pcg(s->np, s->rowptr, s->colidx, s->valDia, s->valCrs, s->iluDia, s->iluCrs, b, x, &cr, &it);
  for (i = 0; i < ndof; i++) s->xyz[i] += x[i];
  for (i = 0; i < ndof; i++) s->uvw[i] = x[i] / dt;
  if (conv_ratio) *conv_ratio = cr;
  if (iteration) *iteration = it;
  free(b); free(x);
}

/* StepTime_InternalDynamics, solve_internal_sparse.h:185-253: the same assembly, solved by Solve_CG (1e-4 / 1000) */
void orc_sim_step_internal_cg(void* h, double dt, double lambda, double myu, double stiff_bend, const double* gravity,
                              double mass_point, double stiff_contact, double clearance, int col_kind,
                              const double* col_par, double* conv_ratio, int* iteration) {
  sim_state* s = (sim_state*)h;
  const int ndof = s->np * 3;
  double* b = (double*)malloc(sizeof(double) * (size_t)ndof);
  double* x = (double*)malloc(sizeof(double) * (size_t)ndof);
  double cr = 1.0e-4;
  int it = 1000, i;
  orc_assemble(s->np, s->ntri, s->nquad, s->tri, s->quad, s->bc, s->xyz0, s->xyz, s->uvw, s->rowptr, s->colidx, dt,
               lambda, myu, stiff_bend, gravity, mass_point, stiff_contact, clearance, col_kind, col_par, s->valDia,
               s->valCrs, b);
  orc_cg(s->np, s->rowptr, s->colidx, s->valDia, s->valCrs, b, x, &cr, &it);
  for (i = 0; i < ndof; i++) s->xyz[i] += x[i];
  for (i = 0; i < ndof; i++) s->uvw[i] = x[i] / dt;
  if (conv_ratio) *conv_ratio = cr;
  if (iteration) *iteration = it;
  free(b); free(x);
}

/* GetIntermidiateVelocityContactResolved, self_collision_cloth.cpp:355-456.  stats as in oracle/ref_shim.cpp */
int orc_sim_resolve_contacts(void* h, double* uvw, const double* xyz_prev, double dt, double clearance,
                             double mass_point, double stiffness, int* stats) {
  sim_state* s = (sim_state*)h;
  const int cap = 1 << 20;
  int* elems = (int*)malloc(sizeof(int) * (size_t)cap * 5);
  int applied, n, itr, i, nccd = 0, nriz = 0, last_nn = 0;
  if (stats) for (i = 0; i < ORC_NSTATS; i++) stats[i] = 0;
  orc_refit_prx(s->nnode, s->nodes, s->ntri, s->tri, s->np, xyz_prev, clearance, s->boxes);
  n = orc_contacts_prx(s->nnode, s->nodes, s->boxes, s->ntri, s->tri, s->np, xyz_prev, clearance, elems, cap);
  if (stats) stats[1] = n;
  applied = n > 0;
  orc_impulse_prx(s->np, uvw, clearance, stiffness, dt, mass_point, xyz_prev, n, elems);
  for (itr = 0; itr < 5; itr++) {
    orc_refit_ccd(s->nnode, s->nodes, s->ntri, s->tri, s->np, xyz_prev, uvw, dt, s->boxes);
    n = orc_contacts_ccd(s->nnode, s->nodes, s->boxes, s->ntri, s->tri, s->np, xyz_prev, uvw, dt, clearance, elems, cap);
    if (stats) { stats[8 + itr] = n; stats[2] = ++nccd; }
    if (n == 0) { /* :416 early return skips the impact-zone loop */
      if (stats) stats[4] = applied;
      free(elems);
      return applied;
    }
    applied = 1;
    orc_impulse_ccd(s->np, uvw, clearance, stiffness, dt, mass_point, xyz_prev, n, elems);
  }
  {
    double* uvw0 = (double*)dup_mem(uvw, sizeof(double) * (size_t)s->np * 3);
    void* riz = orc_riz_new();
    int* zone = (int*)malloc(sizeof(int) * (size_t)s->np);
    for (itr = 0; itr < 100; itr++) {
      orc_refit_ccd(s->nnode, s->nodes, s->ntri, s->tri, s->np, xyz_prev, uvw, dt, s->boxes);
      n = orc_contacts_ccd(s->nnode, s->nodes, s->boxes, s->ntri, s->tri, s->np, xyz_prev, uvw, dt, clearance, elems, cap);
      orc_riz_get(riz, s->np, zone);
      last_nn = 0;
      for (i = 0; i < s->np; i++) last_nn += zone[i] >= 0;
      if (stats) { stats[16 + itr] = n; stats[3] = ++nriz; stats[5] = last_nn; }
      if (n == 0) break;
      orc_riz_make(riz, n, elems, s->np, s->edge_index, s->edge_array);
      orc_riz_apply(riz, s->np, uvw, xyz_prev, uvw0);
    }
    orc_riz_free(riz);
    free(zone);
    free(uvw0);
  }
  if (stats) stats[4] = applied;
  free(elems);
  return applied;
}

/* StepTime, self_contact_sparse/main.cpp:213-263 without MakeNormal */
void orc_sim_step_self_contact(void* h, double dt, double lambda, double myu, double stiff_bend, const double* gravity,
                               double mass_point, double stiff_contact, double clearance, int col_kind,
                               const double* col_par, int* stats) {
  sim_state* s = (sim_state*)h;
  double* xyz1 = (double*)dup_mem(s->xyz, sizeof(double) * (size_t)s->np * 3);
  int iters = -1, applied, ip;
  orc_sim_step_internal(h, dt, lambda, myu, stiff_bend, gravity, mass_point, stiff_contact, clearance, col_kind,
                        col_par, NULL, &iters);
  applied = orc_sim_resolve_contacts(h, s->uvw, xyz1, dt, clearance, mass_point, stiff_contact, stats);
  if (stats) stats[0] = iters;
  if (applied) {
    for (ip = 0; ip < s->np; ip++) {
      s->xyz[ip * 3 + 0] = xyz1[ip * 3 + 0] + s->uvw[ip * 3 + 0] * dt;
      s->xyz[ip * 3 + 1] = xyz1[ip * 3 + 1] + s->uvw[ip * 3 + 1] * dt;
      s->xyz[ip * 3 + 2] = xyz1[ip * 3 + 2] + s->uvw[ip * 3 + 2] * dt;
    }
    orc_update_intermediate_velocity(s->np, s->ntri, s->nquad, s->tri, s->quad, s->bc, s->xyz0, s->xyz, s->uvw, dt,
                                     lambda, myu, stiff_bend, gravity, mass_point);
  }
  free(xyz1);
}
