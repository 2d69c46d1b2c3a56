// csrc/linsolve.cu -- block-CRS (3x3) linear algebra: SpMV, BLAS-1, ILU(0) factor / solve, PCG and CG drivers.
//
// Replaces CMatrixSquareSparse::MatVec (matrix_square_sparse.cpp:167-270, len == 3 at :218-242), InnerProduct
// (:313-324), AXPY (:327-336), Solve_CG (:338-406), CPreconditionerILU::SetValueILU / DoILUDecomp /
// ForwardSubstitution / BackwardSubstitution (ilu_sparse.cpp:352-384, 507-569, 185-215, 295-324) and Solve_PCG
// (:628-709).
//
// Parity: SpMV, the factorisation and the triangular solves keep the reference's per-row operation order and the
// dot products return the value of the reference's serial left-to-right loop (k_xdot_*, below), so PCG / CG iterates
// and iteration counts are bit-identical to it (compiled with -fmad=false).  The fast build (CLOTHGPU_FAST_MODE) uses
// a fixed-shape tree for the dot products instead.
//
// The triangular solves and the factorisation are dependency chains over the rows (1021 levels on a 256^2 grid in
// the reference's natural ordering, which parity forces us to keep).  The production path is the strip-pipelined
// wavefront of trisolve.cu / ilufactor.cu; the level-scheduled kernels in this file (one CTA per scene, a block
// barrier per level) are the fallback for patterns whose plans do not fit shared memory and the CLOTHGPU_TRISOLVE=levels
// measurement switch.
#include <climits>
#include <cstdio>
#include <cstdlib>

#include "ctx.h"

namespace cg {


// ------------------------------------------------------------------------------------------------ SpMV
// one thread per (row, component): y_r = beta*y_r + alpha * sum_j (a_r0 x0 + a_r1 x1 + a_r2 x2), off-diagonal blocks in
// crs order, diagonal last.  If partials != nullptr the block also reduces x.y (fused p.Ap of PCG).
__global__ void __launch_bounds__(256) k_matvec(int np, int ncrs, const int* __restrict__ rowptr,
                                                 const int* __restrict__ colidx, const double* __restrict__ valDia,
                                                 const double* __restrict__ valCrs, double alpha,
                                                 const double* __restrict__ x, double beta, double* __restrict__ y,
                                                 const int* __restrict__ cnt) {
  const int s = blockIdx.y;
  if (cnt && cnt[s * NCNT + CNT_PCG_DONE]) return;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= np * 3) return;
  const int i = tid / 3, r = tid - i * 3;
  const double* xs = x + (size_t)s * np * 3;
  const double* crs = valCrs + (size_t)s * ncrs * 9;
  double acc = (beta == 0.0) ? 0.0 : y[(size_t)s * np * 3 + tid] * beta;
  const int j0 = rowptr[i], j1 = rowptr[i + 1];
  for (int j = j0; j < j1; j++) {
    const double* a = crs + (size_t)j * 9 + r * 3;
    const double* xj = xs + (size_t)colidx[j] * 3;
    acc += alpha * (a[0] * xj[0] + a[1] * xj[1] + a[2] * xj[2]);
  }
  {
    const double* a = valDia + ((size_t)s * np + i) * 9 + r * 3;
    const double* xj = xs + (size_t)i * 3;
    acc += alpha * (a[0] * xj[0] + a[1] * xj[1] + a[2] * xj[2]);
  }
  y[(size_t)s * np * 3 + tid] = acc;
}

// ------------------------------------------------------------------------------------------------ reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
// block-wide sum in a fixed order; result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* sm) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sm[w] = v;
  __syncthreads();
  double r = 0;
  if (threadIdx.x == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    for (int k = 0; k < nw; k++) r += sm[k];
  }
  return r;
}

// what the last block does with the finished sum (Solve_PCG / Solve_CG scalar bookkeeping, ilu_sparse.cpp:640-705)
enum { POST_STORE = 0, POST_PCG_INIT = 1, POST_PCG_CHECK = 2, POST_FINAL_CONV = 3, POST_CG_CHECK = 4,
       POST_SKIP = 5 };  // POST_SKIP: k_update_xr without its fused reduction (the exact dot follows)

struct DotArgs {
  int n;
  int post;
  int slot;      // scal[] slot that receives the sum
  int iter;      // current iteration (POST_*_CHECK)
  double tol2;   // squared tolerance
};

__device__ void dot_post(const DotArgs& A, double sum, double* scal, int* cnt) {
  scal[A.slot] = sum;
  if (A.post == POST_PCG_INIT) {  // ilu_sparse.cpp:648-655: early out on a zero right-hand side
    if (sum < 1.0e-30) {
      cnt[CNT_PCG_DONE] = 1;
      cnt[CNT_PCG_ITER] = 0;
      scal[SC_CONV] = 0.0;
    } else {
      scal[SC_INV_SQ0] = 1.0 / sum;
    }
  } else if (A.post == POST_PCG_CHECK) {  // :684-690
    if (sum * scal[SC_INV_SQ0] < A.tol2) {
      scal[SC_CONV] = sqrt(sum * scal[SC_INV_SQ0]);
      cnt[CNT_PCG_ITER] = A.iter;
      cnt[CNT_PCG_DONE] = 1;
    }
  } else if (A.post == POST_FINAL_CONV) {  // :706-708
    scal[SC_CONV] = sqrt(sum * scal[SC_INV_SQ0]);
  } else if (A.post == POST_CG_CHECK) {  // matrix_square_sparse.cpp:389-394: reports the PREVIOUS residual
    if (sum * scal[SC_INV_SQ0] < A.tol2) {
      scal[SC_CONV] = sqrt(scal[SC_SQ] * scal[SC_INV_SQ0]);
      cnt[CNT_PCG_ITER] = A.iter;
      cnt[CNT_PCG_DONE] = 1;
    }
  }
}

__device__ __forceinline__ void reduce_finish(double v, const DotArgs& A, double* partials, double* scal, int* cnt,
                                              double* sm) {
  const double bs = block_sum(v, sm);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = bs;
    __threadfence();
    const int t = atomicAdd(&cnt[CNT_TICKET], 1);
    last = (t == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last block to arrive sums the per-block partials in a fixed order (lane-strided, then a shuffle tree)
  if (threadIdx.x < 32) {
    double acc = 0;
    for (int k = threadIdx.x; k < (int)gridDim.x; k += 32) acc += __ldcg(partials + k);
    acc = warp_sum(acc);
    if (threadIdx.x == 0) {
      cnt[CNT_TICKET] = 0;
      dot_post(A, acc, scal, cnt);
    }
  }
}

// a.b per scene; skip_done: scenes whose solver has converged keep their scalars
__global__ void __launch_bounds__(256) k_dot(DotArgs A, const double* __restrict__ a, const double* __restrict__ b,
                                              double* __restrict__ partials, double* __restrict__ scal,
                                              int* __restrict__ cnt, int npartial, int skip_done) {
  __shared__ double sm[8];
  const int s = blockIdx.y;
  int* c = cnt + s * NCNT;
  if (skip_done && c[CNT_PCG_DONE]) return;
  const double* as = a + (size_t)s * A.n;
  const double* bs = b + (size_t)s * A.n;
  double v = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) v += as[i] * bs[i];
  reduce_finish(v, A, partials + (size_t)s * npartial, scal + s * NSCAL, c, sm);
}

// ------------------------------------------------------------------------------------------------ exact serial dot
// InnerProduct (matrix_square_sparse.cpp:313-324) is a serial loop, r += a[i]*b[i], and every PCG scalar derives from
// it; the self-collision response downstream is chaotic (a 1e-16 relative change of the velocities entering
// GetIntermidiateVelocityContactResolved moves its output by 1e-2, measured on the reference itself), so whole steps
// only match the reference if the dot products do, bit for bit.  The three kernels below return exactly the value of
// that serial loop, in parallel:
//   while the running sum s stays inside one binade [2^e, 2^(e+1)) with ulp u, fl(s + p) = s + u*rint(p/u) (unless p/u
//   is an exact tie), i.e. the rounded additions are EXACT integer additions in units of u and therefore associative.
//   k_xdot_partial : approximate (tree) sums per chunk of 256 products -> predicted binade of s at each chunk start
//   k_xdot_stats   : per chunk, in units of the predicted u: integer total K, min / max of the integer prefix sums,
//                    a "bad" flag (tie, overflow)
//   k_xdot_walk    : one warp per scene walks the chunks 32 at a time (integer scan); a chunk is accepted when the
//                    exact s at its start is in the predicted binade and S + prefix stays strictly inside
//                    (2^52, 2^53) throughout; any other chunk (a binade crossing, a tie, the first chunk: ~15 of 768
//                    at 256^2) is redone 32 elements at a time with the true s, element by element where needed.
// No approximation anywhere: a chunk is either provably exact or recomputed.
// (integers are carried as doubles: every quantity that decides about or enters an ACCEPTED chunk is an integer of
// magnitude < 2^53 and therefore exact; a sum that got rounded has passed 2^53 on the way, which the min / max of the
// prefix sums record, so such a chunk can only be rejected)
struct XChunk {
  double K, pmin, pmax;
  int e, bad;
};
static_assert(sizeof(XChunk) == 32, "XChunk layout");
enum { XD_CHUNK = 256 };

__device__ __forceinline__ bool xd_binade(double s, int& e) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(s) & 0x7fffffffffffffffULL;
  const int be = (int)(bits >> 52);
  if (be == 0 || be == 0x7ff) return false;  // zero, subnormal, inf, nan: no binade arithmetic
  e = be - 1023;
  return e >= -960;  // 2^(e-52) and its inverse are normal powers of two
}
__device__ __forceinline__ double xd_pow2(int k) { return __longlong_as_double((long long)(1023 + k) << 52); }
__device__ __forceinline__ double xd_scan_incl(double v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}
// S + [lo, hi] strictly inside the binade of sign(s)
__device__ __forceinline__ bool xd_inside(bool positive, double S, double lo, double hi) {
  const double L = 4503599627370497.0, H = 9007199254740991.0;  // 2^52 + 1, 2^53 - 1
  return positive ? (S + lo >= L && S + hi <= H) : (S + lo >= -H && S + hi <= -L);
}

// lane-local statistics of 8 consecutive products in units of u = 2^(e-52): integer total, min / max of the prefix
// sums, and whether any element is not exactly representable (tie or too large)
__device__ __forceinline__ void xd_lane_stats(const double (&p)[8], double scale, double& run, double& lmin,
                                              double& lmax, bool& bad) {
  run = 0; lmin = 0; lmax = 0; bad = false;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const double q = p[j] * scale;
    const double kf = rint(q);
    if (!(fabs(q) < 9007199254740992.0) || fabs(q - kf) == 0.5) bad = true;
    else run += kf;
    lmin = fmin(lmin, run);
    lmax = fmax(lmax, run);
  }
}

__global__ void __launch_bounds__(256) k_xdot_partial(int n, int nch, const double* __restrict__ a,
                                                       const double* __restrict__ b, double* __restrict__ tsum,
                                                       const int* __restrict__ cnt, int skip_done) {
  const int s = blockIdx.y;
  if (skip_done && cnt[s * NCNT + CNT_PCG_DONE]) return;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (c >= nch) return;
  const double* as = a + (size_t)s * n;
  const double* bs = b + (size_t)s * n;
  double v = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int i = c * XD_CHUNK + j * 32 + lane;
    if (i < n) v += as[i] * bs[i];
  }
  v = warp_sum(v);
  if (lane == 0) tsum[(size_t)s * nch + c] = v;
}

__global__ void __launch_bounds__(256) k_xdot_stats(int n, int nch, const double* __restrict__ a,
                                                     const double* __restrict__ b, const double* __restrict__ tsum,
                                                     XChunk* __restrict__ st, const int* __restrict__ cnt, int skip_done) {
  __shared__ double sm[8];
  __shared__ double base;
  const int s = blockIdx.y;
  if (skip_done && cnt[s * NCNT + CNT_PCG_DONE]) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 8, c = c0 + w;
  const double* ts = tsum + (size_t)s * nch;
  // approximate value of the running sum at the start of this CTA's first chunk
  double acc = 0;
  for (int k = threadIdx.x; k < c0; k += 256) acc += ts[k];
  const double tot = block_sum(acc, sm);
  if (threadIdx.x == 0) base = tot;
  __syncthreads();
  if (c >= nch) return;
  double start = base;
  {
    const double mine = (lane < 8 && c0 + lane < nch) ? ts[c0 + lane] : 0.0;
    for (int k = 0; k < w; k++) start += __shfl_sync(0xffffffffu, mine, k);
  }
  XChunk out;
  out.K = 0; out.pmin = 0; out.pmax = 0; out.e = 0; out.bad = 1;
  int e;
  if (xd_binade(start, e)) {
    out.e = e;
    const double scale = xd_pow2(-(e - 52));
    const double* as = a + (size_t)s * n;
    const double* bs = b + (size_t)s * n;
    // lane l owns the 8 consecutive elements c*256 + l*8 .. +7: local integer prefix, then a scan over the lanes
    double pr[8];
    const int i0 = c * XD_CHUNK + lane * 8;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int i = i0 + j;
      pr[j] = i < n ? as[i] * bs[i] : 0.0;
    }
    double run, lmin, lmax;
    bool bad;
    xd_lane_stats(pr, scale, run, lmin, lmax, bad);
    const double inc = xd_scan_incl(run, lane);
    const double exc = inc - run;
    double mn = exc + lmin, mx = exc + lmax;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const bool anybad = __any_sync(0xffffffffu, bad);
    out.K = __shfl_sync(0xffffffffu, inc, 31);
    out.pmin = mn;
    out.pmax = mx;
    out.bad = (anybad || !(mx - mn < 9007199254740992.0)) ? 1 : 0;
    if (out.bad) out.K = 0;
  }
  if (lane == 0) st[(size_t)s * nch + c] = out;
}

enum { XD_WALK_THREADS = 256, XD_CACHE = 24 };  // cached "suspect" chunks: those whose slow path is foreseeable

// ---- pieces of the walk, executed by one warp
// chunks [beg, end) 32 at a time with their precomputed statistics; a chunk that cannot be accepted is redone with the
// true running sum (products from global memory, lane statistics on the fly)
struct XVariant {  // lane statistics of one cached chunk for one binade, with the scan over the lanes done
  int e;                   // the binade (INT_MIN: not available)
  const double4* st;       // per lane {exclusive prefix of the lane totals, lmin, lmax, bad}
  const double* inc;       // per lane inclusive prefix
};
__device__ __forceinline__ void xd_slow_chunk(double& sum, const double (&pr)[8], const double* __restrict__ prs, int lane,
                                              const XVariant& VA, const XVariant& VB);
__device__ void xd_group_walk(double& sum, const XChunk* __restrict__ S, int beg, int end, const double* __restrict__ as,
                              const double* __restrict__ bs, int n, int lane) {
  int pos = beg;
  while (pos < end) {
    const int c = pos + lane;
    XChunk ch;
    if (c < end) ch = S[c];
    else { ch.K = 0; ch.pmin = 0; ch.pmax = 0; ch.e = 0; ch.bad = 1; }
    int e = 0;
    const bool okbin = xd_binade(sum, e);
    const double Kl = ch.bad ? 0.0 : ch.K;
    const double inc = xd_scan_incl(Kl, lane);
    bool ok = false;
    double S0 = 0;
    if (okbin) {
      S0 = sum * xd_pow2(-(e - 52));
      if (!ch.bad && ch.e == e) ok = xd_inside(sum > 0, S0 + (inc - Kl), ch.pmin, ch.pmax);
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    const int f = (m == 0xffffffffu) ? 32 : (__ffs(~m) - 1);
    if (f > 0) {
      const double tot = __shfl_sync(0xffffffffu, inc, f - 1);
      sum = (S0 + tot) * xd_pow2(e - 52);
    }
    pos += f;
    if (f < 32 && pos < end) {
      double pr[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int i = pos * XD_CHUNK + lane * 8 + j;
        pr[j] = i < n ? as[i] * bs[i] : 0.0;
      }
      const XVariant none = {INT_MIN, nullptr, nullptr};
      xd_slow_chunk(sum, pr, nullptr, lane, none, none);
      pos++;
    }
  }
}

// one chunk with the true running sum: lane l owns its elements l*8 .. l*8+7 (pr; prs = the same 256 products in
// shared memory, or null).  VA / VB: lane statistics precomputed for two foreseeable binades, already scanned over the
// lanes, so that a step costs a few loads, a ballot and no shuffle; any other binade is evaluated on the fly.
// Lanes are accepted up to the first one whose prefix leaves the binade (or holds a tie); that lane's 8 products are
// added by true floating-point additions, exactly as the serial loop does, and the rest continues in the new binade.
// (products beyond n are exact zeros: adding them changes nothing, in either path)
__device__ __forceinline__ void xd_slow_chunk(double& sum, const double (&pr)[8], const double* __restrict__ prs, int lane,
                                              const XVariant& VA, const XVariant& VB) {
  int L0 = 0;
  while (L0 < 32) {
    int e2 = 0;
    int f2 = L0;
    if (xd_binade(sum, e2)) {
      const double S2 = sum * xd_pow2(-(e2 - 52));
      if (e2 == VA.e || e2 == VB.e) {
        const XVariant& V = e2 == VA.e ? VA : VB;
        const double base = L0 > 0 ? V.inc[L0 - 1] : 0.0;
        const double4 q = V.st[lane];
        const bool okl = lane < L0 || (q.w == 0.0 && xd_inside(sum > 0, S2 + (q.x - base), q.y, q.z));
        const unsigned m2 = __ballot_sync(0xffffffffu, okl);
        f2 = (m2 == 0xffffffffu) ? 32 : (__ffs(~m2) - 1);
        if (f2 > L0) sum = (S2 + (V.inc[f2 - 1] - base)) * xd_pow2(e2 - 52);
      } else {
        double run, lmin, lmax;
        bool bad;
        xd_lane_stats(pr, xd_pow2(-(e2 - 52)), run, lmin, lmax, bad);
        if (lane < L0) { run = 0; lmin = 0; lmax = 0; bad = false; }
        const double inc2 = xd_scan_incl(run, lane);
        const bool okl = lane < L0 || (!bad && xd_inside(sum > 0, S2 + (inc2 - run), lmin, lmax));
        const unsigned m2 = __ballot_sync(0xffffffffu, okl);
        f2 = (m2 == 0xffffffffu) ? 32 : (__ffs(~m2) - 1);
        if (f2 > L0) {
          const double tot2 = __shfl_sync(0xffffffffu, inc2, f2 - 1);
          sum = (S2 + tot2) * xd_pow2(e2 - 52);
        }
      }
    }
    L0 = f2;
    if (L0 < 32) {
      double t = sum;
      if (prs) {  // every lane adds the same eight products (shared-memory broadcast): no shuffle on the chain
#pragma unroll
        for (int j = 0; j < 8; j++) t = t + prs[L0 * 8 + j];
        sum = t;
      } else {
#pragma unroll
        for (int j = 0; j < 8; j++) t = t + pr[j];
        sum = __shfl_sync(0xffffffffu, t, L0);
      }
      L0++;
    }
  }
}

// One CTA per scene.  All warps prepare, warp 0 walks:
//   1. chunk statistics -> shared memory; "suspect" chunks (flagged, the first, and those where the predicted binade
//      changes) are listed in order; the runs of chunks between two suspects are "segments" (one binade, no flags)
//   2. for the first XD_CACHE suspects: products and lane statistics for the two foreseeable binades (the predicted one
//      at the chunk's start and at the next chunk's start) -> shared memory; for every segment: its combined
//      statistics (integer total, min / max of the prefix sums) -> one O(1) check replaces the chunk-by-chunk walk
//   3. the walk: segment, suspect, segment, ...: ~2 x (number of binade crossings) short steps instead of nch / 32
template <bool SMEM_STATS>
__global__ void __launch_bounds__(XD_WALK_THREADS) k_xdot_walk(DotArgs A, int nch, const double* __restrict__ a,
                                                                const double* __restrict__ b,
                                                                const XChunk* __restrict__ st, double* __restrict__ scal,
                                                                int* __restrict__ cnt, int skip_done) {
  extern __shared__ __align__(16) unsigned char xsm[];
  __shared__ int nsus_total;
  __shared__ int sus[XD_CACHE];
  __shared__ int sus_e[XD_CACHE][2];
  __shared__ double seg_K[XD_CACHE + 1], seg_min[XD_CACHE + 1], seg_max[XD_CACHE + 1];
  __shared__ int seg_e[XD_CACHE + 1], seg_ok[XD_CACHE + 1];
  __shared__ int wcount[XD_WALK_THREADS / 32];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int* cn = cnt + s * NCNT;
  if ((skip_done & 1) && cn[CNT_PCG_DONE]) return;
  const bool dbg = (skip_done & 2) != 0;  // CLOTHGPU_XDOT_DEBUG: phase cycle counts of scene 0 (measurement aid)
  long long tk[6] = {0, 0, 0, 0, 0, 0};
  if (dbg) tk[0] = clock64();
  const int n = A.n;
  const double* as = a + (size_t)s * n;
  const double* bs = b + (size_t)s * n;
  const XChunk* cs = st + (size_t)s * nch;
  double* pcache = reinterpret_cast<double*>(xsm);                                        // XD_CACHE x 256 products
  double4* lstat = reinterpret_cast<double4*>(xsm + XD_CACHE * XD_CHUNK * 8);             // XD_CACHE x 2 x 32 lanes
  double* linc = reinterpret_cast<double*>(xsm + XD_CACHE * XD_CHUNK * 8 + XD_CACHE * 2 * 32 * 32);
  XChunk* sst = reinterpret_cast<XChunk*>(xsm + XD_CACHE * XD_CHUNK * 8 + XD_CACHE * 2 * 32 * 40);  // nch (SMEM_STATS)
  if (SMEM_STATS) {
    const int4* src = reinterpret_cast<const int4*>(cs);
    int4* dst = reinterpret_cast<int4*>(sst);
    for (int k = tid; k < nch * 2; k += XD_WALK_THREADS) dst[k] = src[k];
  }
  __syncthreads();
  const XChunk* S = SMEM_STATS ? sst : cs;
  // ---- 1. ordered list of the suspect chunks: thread t looks at a contiguous range of chunks
  const int per = (nch + XD_WALK_THREADS - 1) / XD_WALK_THREADS;
  const int c0 = tid * per, c1 = min(nch, c0 + per);
  int mine = 0;
  for (int c = c0; c < c1; c++) mine += (c == 0 || S[c].bad || (c + 1 < nch && S[c + 1].e != S[c].e)) ? 1 : 0;
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wcount[warp] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < warp; w++) base += wcount[w];
  if (tid == XD_WALK_THREADS - 1) nsus_total = base + incl;
  int posn = base + incl - mine;
  for (int c = c0; c < c1; c++) {
    if (c == 0 || S[c].bad || (c + 1 < nch && S[c + 1].e != S[c].e)) {
      if (posn < XD_CACHE) {
        sus[posn] = c;
        sus_e[posn][0] = S[c].bad && S[c].e == 0 ? INT_MIN : S[c].e;
        sus_e[posn][1] = (c + 1 < nch && !(S[c + 1].bad && S[c + 1].e == 0) && S[c + 1].e != S[c].e) ? S[c + 1].e : INT_MIN;
      }
      posn++;
    }
  }
  __syncthreads();
  const int nsus = min(nsus_total, XD_CACHE);
  if (dbg) tk[1] = clock64();
  // ---- 2a. products of the cached suspects
  {  // (all loads of a thread issued before the first use: one global-memory latency, not one per suspect)
    double pa[XD_CACHE], pb[XD_CACHE];
#pragma unroll
    for (int q = 0; q < XD_CACHE; q++) {
      const int i = q < nsus ? sus[q] * XD_CHUNK + tid : n;
      pa[q] = i < n ? as[i] : 0.0;
      pb[q] = i < n ? bs[i] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < XD_CACHE; q++)
      if (q < nsus) pcache[q * XD_CHUNK + tid] = pa[q] * pb[q];
  }
  // ---- 2b. segment statistics: segment g = chunks between suspect g-1 and suspect g (the last one runs to nch)
  for (int g = warp; g <= nsus; g += XD_WALK_THREADS / 32) {
    const int beg = g == 0 ? 0 : sus[g - 1] + 1, end = g < nsus ? sus[g] : nch;
    const int len = max(end - beg, 0), lp = (len + 31) / 32;
    const int b0 = beg + lane * lp, b1 = min(end, b0 + lp);
    double run = 0, mn = 0, mx = 0;
    bool bad = false;
    for (int c = b0; c < b1; c++) {
      const XChunk ch = S[c];
      bad = bad || ch.bad || ch.e != S[beg].e;
      mn = fmin(mn, run + ch.pmin);
      mx = fmax(mx, run + ch.pmax);
      run += ch.K;
    }
    const double inc = xd_scan_incl(run, lane);
    mn += inc - run;
    mx += inc - run;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const bool anybad = __any_sync(0xffffffffu, bad);
    const double tot = __shfl_sync(0xffffffffu, inc, 31);
    if (lane == 0) {
      seg_K[g] = tot; seg_min[g] = fmin(mn, 0.0); seg_max[g] = fmax(mx, 0.0);
      seg_e[g] = len > 0 ? S[beg].e : 0;
      // the tail segment is only trustworthy when every suspect is cached (otherwise it contains uncached suspects)
      seg_ok[g] = (len > 0 && !anybad && (g < nsus || nsus_total <= XD_CACHE)) ? 1 : 0;
    }
  }
  __syncthreads();
  if (dbg) tk[2] = clock64();
  // ---- 2c. lane statistics of the cached suspects for their two foreseeable binades
  for (int k = warp; k < nsus * 2; k += XD_WALK_THREADS / 32) {
    const int e = sus_e[k >> 1][k & 1];
    if (e == INT_MIN) continue;
    double pr[8];
#pragma unroll
    for (int j = 0; j < 8; j++) pr[j] = pcache[(k >> 1) * XD_CHUNK + lane * 8 + j];
    double run, lmin, lmax;
    bool bad;
    xd_lane_stats(pr, xd_pow2(-(e - 52)), run, lmin, lmax, bad);
    const double inc = xd_scan_incl(run, lane);
    const double exc = inc - run;
    // the scanned form is only used when every prefix is an exactly represented integer (|.| < 2^52)
    const bool big = !(fabs(exc + lmin) < 4503599627370496.0) || !(fabs(exc + lmax) < 4503599627370496.0);
    if (__any_sync(0xffffffffu, big)) {
      if (lane == 0) sus_e[k >> 1][k & 1] = INT_MIN;
      continue;
    }
    lstat[k * 32 + lane] = make_double4(exc, lmin, lmax, bad ? 1.0 : 0.0);
    linc[k * 32 + lane] = inc;
  }
  __syncthreads();
  if (tid >= 32) return;
  if (dbg) tk[3] = clock64();

  // ---- 3. the walk
  double sum = 0.0;
  int nslow_seg = 0;
  for (int g = 0; g <= nsus; g++) {
    const int beg = g == 0 ? 0 : sus[g - 1] + 1, end = g < nsus ? sus[g] : nch;
    if (end > beg) {
      int e = 0;
      bool fast = false;
      if (seg_ok[g] && xd_binade(sum, e) && e == seg_e[g]) {
        const double S0 = sum * xd_pow2(-(e - 52));
        if (xd_inside(sum > 0, S0, seg_min[g], seg_max[g])) {
          sum = (S0 + seg_K[g]) * xd_pow2(e - 52);
          fast = true;
        }
      }
      if (!fast) { xd_group_walk(sum, S, beg, end, as, bs, n, lane); nslow_seg++; }
    }
    if (g < nsus) {
      double pr[8];
#pragma unroll
      for (int j = 0; j < 8; j++) pr[j] = pcache[g * XD_CHUNK + lane * 8 + j];
      const XVariant VA = {sus_e[g][0], lstat + (g * 2) * 32, linc + (g * 2) * 32};
      const XVariant VB = {sus_e[g][1], lstat + (g * 2 + 1) * 32, linc + (g * 2 + 1) * 32};
      xd_slow_chunk(sum, pr, pcache + g * XD_CHUNK, lane, VA, VB);
    }
  }
  if (lane == 0) dot_post(A, sum, scal + s * NSCAL, cn);
  if (dbg && lane == 0 && s == 0) {
    tk[4] = clock64();
    printf("xdot walk: suspects %d (cached %d), segments walked slowly %d; cycles: stats+list %lld, products+segments %lld, "
           "lane stats %lld, walk %lld\n", nsus_total, nsus, nslow_seg, tk[1] - tk[0], tk[2] - tk[1], tk[3] - tk[2], tk[4] - tk[3]);
  }
}

// PCG update (ilu_sparse.cpp:674-682): alpha = rPr / pAp; r += (-alpha) Ap; x += alpha p; then r.r
// CG update (matrix_square_sparse.cpp:378-386): alpha = sq / pAp; x += alpha p; r += (-alpha) Ap
__global__ void __launch_bounds__(256) k_update_xr(DotArgs A, int rpr_slot, const double* __restrict__ Ap,
                                                    const double* __restrict__ p, double* __restrict__ r,
                                                    double* __restrict__ x, double* __restrict__ partials,
                                                    double* __restrict__ scal, int* __restrict__ cnt, int npartial) {
  __shared__ double sm[8];
  const int s = blockIdx.y;
  int* c = cnt + s * NCNT;
  if (c[CNT_PCG_DONE]) return;
  double* sc = scal + s * NSCAL;
  const double alpha = sc[rpr_slot] / sc[SC_PAP];
  const size_t off = (size_t)s * A.n;
  double v = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < A.n; i += gridDim.x * blockDim.x) {
    const double ri = r[off + i] + (-alpha) * Ap[off + i];
    r[off + i] = ri;
    x[off + i] += alpha * p[off + i];
    v += ri * ri;
  }
  if (A.post == POST_SKIP) return;
  reduce_finish(v, A, partials + (size_t)s * npartial, sc, c, sm);
}

// p = z + beta p with beta = scal[num] / scal[den]  (ilu_sparse.cpp:699-703, matrix_square_sparse.cpp:396-402)
__global__ void k_xpay(int n, int num_slot, int den_slot, const double* __restrict__ z, double* __restrict__ p,
                       const double* __restrict__ scal, const int* __restrict__ cnt) {
  const int s = blockIdx.y;
  if (cnt[s * NCNT + CNT_PCG_DONE]) return;
  const double beta = scal[s * NSCAL + num_slot] / scal[s * NSCAL + den_slot];
  const size_t off = (size_t)s * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    p[off + i] = z[off + i] + beta * p[off + i];
}

__global__ void k_axpy(int n, double a, const double* __restrict__ x, double* __restrict__ y) {
  const size_t off = (size_t)blockIdx.y * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[off + i] += a * x[off + i];
}
__global__ void k_copy(int n, const double* __restrict__ x, double* __restrict__ y, const int* __restrict__ cnt) {
  if (cnt && cnt[blockIdx.y * NCNT + CNT_PCG_DONE]) return;
  const size_t off = (size_t)blockIdx.y * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[off + i] = x[off + i];
}
__global__ void k_set_scalar_slot(double* scal, int dst, int src) { scal[blockIdx.x * NSCAL + dst] = scal[blockIdx.x * NSCAL + src]; }
__global__ void k_pcg_reset(int* cnt, double* scal, int max_iter, double conv_init) {
  int* c = cnt + blockIdx.x * NCNT;
  scal[blockIdx.x * NSCAL + SC_CONV] = conv_init;  // Solve_CG leaves conv_ratio untouched when it does not converge
  c[CNT_PCG_DONE] = 0;
  c[CNT_PCG_ITER] = max_iter;
  c[CNT_TICKET] = 0;
}
// X += x ; V = x / dt   (solve_internal_sparse.h:323-325)
__global__ void k_advance(int n, double dt, const double* __restrict__ x, double* __restrict__ xyz, double* __restrict__ uvw) {
  const size_t off = (size_t)blockIdx.y * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    xyz[off + i] += x[off + i];
    uvw[off + i] = x[off + i] / dt;
  }
}

// ------------------------------------------------------------------------------------------------ ILU(0)
// DoILUDecomp, ilu_sparse.cpp:507-569.  One CTA per scene; rows of one dependency level are independent.  A warp
// serves three rows at a time, nine lanes per row = the nine entries of a 3x3 block; the row's update list
// (precomputed: ik, kj, ij slot triples in the reference's order) is applied sequentially, entry-parallel.
__global__ void __launch_bounds__(1024) k_ilu_factor(int np, int ncrs, int nlev, const int* __restrict__ lev_ptr,
                                                      const int* __restrict__ lev_rows, const int* __restrict__ rowptr,
                                                      const int* __restrict__ diaind, const int* __restrict__ op_ptr,
                                                      const int4* __restrict__ ops, double* iluDia, double* iluCrs,
                                                      int* __restrict__ cnt) {
  const int s = blockIdx.x;
  double* dia = iluDia + (size_t)s * np * 9;
  double* crs = iluCrs + (size_t)s * ncrs * 9;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int sub = lane / 9, e = lane - sub * 9;
  const int r = e / 3, q = e - r * 3;
  for (int lev = 0; lev < nlev; lev++) {
    const int b0 = lev_ptr[lev], nrows = lev_ptr[lev + 1] - b0;
    for (int base = warp * 3; base < nrows; base += nwarp * 3) {
      const int row = (sub < 3 && base + sub < nrows) ? lev_rows[b0 + base + sub] : -1;
      const int o0 = row >= 0 ? op_ptr[row] : 0;
      const int nop = row >= 0 ? op_ptr[row + 1] - o0 : 0;
      int maxop = nop;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) maxop = max(maxop, __shfl_xor_sync(0xffffffffu, maxop, o));
      double* dii = dia + (size_t)(row >= 0 ? row : 0) * 9;
      for (int o = 0; o < maxop; o++) {
        if (o < nop) {
          const int4 op = ops[o0 + o];
          const double* vik = crs + (size_t)op.x * 9;
          const double* vkj = crs + (size_t)op.y * 9;
          double* vij = op.z < 0 ? dii : crs + (size_t)op.z * 9;
          const double v = vij[e] - (vik[r * 3 + 0] * vkj[q] + vik[r * 3 + 1] * vkj[3 + q] + vik[r * 3 + 2] * vkj[6 + q]);
          vij[e] = v;
        }
        __syncwarp();
      }
      // D_i <- D_i^-1 by cofactors (CalcInvMat3, ilu_sparse.cpp:77-97); a singular pivot is left as is (:541-551)
      double a[9];
      bool invert = false;
      if (row >= 0) {
#pragma unroll
        for (int k = 0; k < 9; k++) a[k] = dii[k];
        const double det = +a[0] * a[4] * a[8] + a[3] * a[7] * a[2] + a[6] * a[1] * a[5] - a[0] * a[7] * a[5] -
                           a[6] * a[4] * a[2] - a[3] * a[1] * a[8];
        invert = fabs(det) > 1.0e-30;
        if (invert) {
          const double inv_det = 1.0 / det;
          double t[9];
#pragma unroll
          for (int k = 0; k < 9; k++) t[k] = a[k];
          a[0] = inv_det * (t[4] * t[8] - t[5] * t[7]);
          a[1] = inv_det * (t[2] * t[7] - t[1] * t[8]);
          a[2] = inv_det * (t[1] * t[5] - t[2] * t[4]);
          a[3] = inv_det * (t[5] * t[6] - t[3] * t[8]);
          a[4] = inv_det * (t[0] * t[8] - t[2] * t[6]);
          a[5] = inv_det * (t[2] * t[3] - t[0] * t[5]);
          a[6] = inv_det * (t[3] * t[7] - t[4] * t[6]);
          a[7] = inv_det * (t[1] * t[6] - t[0] * t[7]);
          a[8] = inv_det * (t[0] * t[4] - t[1] * t[3]);
        } else if (e == 0) {
          atomicAdd(&cnt[s * NCNT + CNT_NSING], 1);
        }
      }
      __syncwarp();
      if (row >= 0 && invert) dii[e] = a[e];
      // U_ij <- D_i^-1 U_ij for the blocks above the diagonal (:556-567)
      {
        const int j0 = row >= 0 ? diaind[row] : 0;
        const int nup = row >= 0 ? rowptr[row + 1] - j0 : 0;
        int maxup = nup;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) maxup = max(maxup, __shfl_xor_sync(0xffffffffu, maxup, o));
        for (int jj = 0; jj < maxup; jj++) {
          double* vij = crs + (size_t)(j0 + (jj < nup ? jj : 0)) * 9;
          double v = 0;
          if (jj < nup) v = a[r * 3 + 0] * vij[q] + a[r * 3 + 1] * vij[3 + q] + a[r * 3 + 2] * vij[6 + q];
          __syncwarp();
          if (jj < nup) vij[e] = v;
        }
      }
    }
    __syncthreads();
  }
}

// ForwardSubstitution + BackwardSubstitution (len == 3), ilu_sparse.cpp:185-215, 295-324.  One CTA per scene;
// three lanes per row (one per component), ten rows per warp; a block barrier separates dependency levels.
__global__ void __launch_bounds__(320) k_ilu_solve(int np, int ncrs, int nlev_f, const int* __restrict__ lev_f_ptr,
                                                    const int* __restrict__ lev_f_rows, int nlev_b,
                                                    const int* __restrict__ lev_b_ptr, const int* __restrict__ lev_b_rows,
                                                    const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                    const int* __restrict__ diaind, const double* __restrict__ iluDia,
                                                    const double* __restrict__ iluCrs, double* vec,
                                                    const double* __restrict__ src, const int* __restrict__ cnt) {
  const int s = blockIdx.x;
  if (cnt && cnt[s * NCNT + CNT_PCG_DONE]) return;
  const double* dia = iluDia + (size_t)s * np * 9;
  const double* crs = iluCrs + (size_t)s * ncrs * 9;
  double* v = vec + (size_t)s * np * 3;
  if (src) {  // fused "Pr = r" copy of Solve_PCG (ilu_sparse.cpp:692-694)
    const double* sv = src + (size_t)s * np * 3;
    for (int i = threadIdx.x; i < np * 3; i += blockDim.x) v[i] = sv[i];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int g = lane / 3, r = lane - g * 3;  // g == 10 for the two idle lanes
  const int src_lane = g * 3;
  for (int lev = 0; lev < nlev_f; lev++) {
    const int b0 = lev_f_ptr[lev], nrows = lev_f_ptr[lev + 1] - b0;
    for (int base = warp * 10; base < nrows; base += nwarp * 10) {
      const bool on = (g < 10) && (base + g < nrows);
      const int row = on ? lev_f_rows[b0 + base + g] : 0;
      double t = 0;
      if (on) {
        t = v[row * 3 + r];
        const int j0 = rowptr[row], j1 = diaind[row];
        for (int j = j0; j < j1; j++) {
          const double* a = crs + (size_t)j * 9 + r * 3;
          const double* xj = v + (size_t)colidx[j] * 3;
          t -= a[0] * xj[0] + a[1] * xj[1] + a[2] * xj[2];
        }
      }
      const double t0 = __shfl_sync(0xffffffffu, t, src_lane & 31);
      const double t1 = __shfl_sync(0xffffffffu, t, (src_lane + 1) & 31);
      const double t2 = __shfl_sync(0xffffffffu, t, (src_lane + 2) & 31);
      if (on) {
        const double* d = dia + (size_t)row * 9 + r * 3;
        v[row * 3 + r] = d[0] * t0 + d[1] * t1 + d[2] * t2;
      }
    }
    __syncthreads();
  }
  for (int lev = 0; lev < nlev_b; lev++) {
    const int b0 = lev_b_ptr[lev], nrows = lev_b_ptr[lev + 1] - b0;
    for (int base = warp * 10; base < nrows; base += nwarp * 10) {
      const bool on = (g < 10) && (base + g < nrows);
      if (on) {
        const int row = lev_b_rows[b0 + base + g];
        double t = v[row * 3 + r];
        const int j0 = diaind[row], j1 = rowptr[row + 1];
        for (int j = j0; j < j1; j++) {
          const double* a = crs + (size_t)j * 9 + r * 3;
          const double* xj = v + (size_t)colidx[j] * 3;
          t -= a[0] * xj[0] + a[1] * xj[1] + a[2] * xj[2];
        }
        v[row * 3 + r] = t;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ launchers
static dim3 vec_grid(Ctx* c, int n) { return dim3(std::min(div_up(n, 256), c->npartial), c->B); }

int launch_matvec(Ctx* c, double alpha, const double* x, double beta, double* y) {
  dim3 g(div_up(c->ndof, 256), c->B);
  k_matvec<<<g, 256, 0, KS(c, "k_matvec")>>>(c->np, c->ncrs, c->rowptr, c->colidx, c->valDia, c->valCrs, alpha, x, beta, y, nullptr);
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}

static int dot_launch(Ctx* c, const double* a, const double* b, int slot, int post, int iter, double tol2, int skip_done) {
  DotArgs A;
  A.n = c->ndof; A.post = post; A.slot = slot; A.iter = iter; A.tol2 = tol2;
  if (c->exact_dot) {  // the reference's serial sum, bit for bit (three kernels, see k_xdot_*)
    const int nch = c->xd_nch;
    dim3 g(div_up(nch, 8), c->B);
    XChunk* st = reinterpret_cast<XChunk*>(c->xd_stat);
    static const int xdbg = getenv("CLOTHGPU_XDOT_DEBUG") ? 2 : 0;
    k_xdot_partial<<<g, 256, 0, KS(c, "k_xdot_partial")>>>(c->ndof, nch, a, b, c->xd_tsum, c->cnt, skip_done);
    CG_LAUNCH_CHECK(c);
    k_xdot_stats<<<g, 256, 0, KS(c, "k_xdot_stats")>>>(c->ndof, nch, a, b, c->xd_tsum, st, c->cnt, skip_done);
    CG_LAUNCH_CHECK(c);
    {
      const size_t cache_bytes = (size_t)XD_CACHE * XD_CHUNK * 8 + (size_t)XD_CACHE * 2 * 32 * 40, stat_bytes = (size_t)nch * sizeof(XChunk);
      if (cache_bytes + stat_bytes <= 200 * 1024) {
        if (!c->xd_attr_set) {
          CG_CUDA(cudaFuncSetAttribute(k_xdot_walk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
          c->xd_attr_set = true;
        }
        k_xdot_walk<true><<<c->B, XD_WALK_THREADS, cache_bytes + stat_bytes, KS(c, "k_xdot_walk")>>>(A, nch, a, b, st, c->scal,
                                                                                                c->cnt, skip_done | xdbg);
      } else {
        if (!c->xd_attr_set) {
          CG_CUDA(cudaFuncSetAttribute(k_xdot_walk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
          c->xd_attr_set = true;
        }
        k_xdot_walk<false><<<c->B, XD_WALK_THREADS, cache_bytes, KS(c, "k_xdot_walk")>>>(A, nch, a, b, st, c->scal, c->cnt,
                                                                                    skip_done | xdbg);
      }
      CG_LAUNCH_CHECK(c);
    }
    return CLOTHGPU_OK;
  }
  dim3 g(std::min(div_up(c->ndof, 256), c->npartial), c->B);
  k_dot<<<g, 256, 0, KS(c, "k_dot")>>>(A, a, b, c->partials, c->scal, c->cnt, c->npartial, skip_done);
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}
int launch_dot(Ctx* c, const double* a, const double* b, int scal_slot) {
  return dot_launch(c, a, b, scal_slot, POST_STORE, 0, 0.0, 0);
}
int launch_axpy(Ctx* c, double a, const double* x, double* y) {
  k_axpy<<<vec_grid(c, c->ndof), 256, 0, KS(c, "k_axpy")>>>(c->ndof, a, x, y);
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}
int launch_ilu_setvalue(Ctx* c) {
  c->ilu_packed_valid = false;
  CG_CUDA(cudaMemcpyAsync(c->iluDia, c->valDia, sizeof(double) * (size_t)c->B * c->np * 9, cudaMemcpyDeviceToDevice, c->stream));
  CG_CUDA(cudaMemcpyAsync(c->iluCrs, c->valCrs, sizeof(double) * (size_t)c->B * c->ncrs * 9, cudaMemcpyDeviceToDevice, c->stream));
  return CLOTHGPU_OK;
}
int launch_ilu_factor(Ctx* c) {
  c->ilu_packed_valid = false;
  if (c->use_strips_fac) return launch_ilu_factor_strips(c);
  k_ilu_factor<<<c->B, 1024, 0, KS(c, "k_ilu_factor")>>>(c->np, c->ncrs, c->nlev_f, c->lev_f_ptr, c->lev_f_rows, c->rowptr, c->diaind,
                                             c->op_ptr, c->ops, c->iluDia, c->iluCrs, c->cnt);
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}
static int ilu_solve_launch(Ctx* c, double* vec, const double* src, bool skip_done) {
  if (c->use_strips) return launch_ilu_solve_strips(c, vec, src, skip_done);
  k_ilu_solve<<<c->B, 320, 0, KS(c, "k_ilu_solve")>>>(c->np, c->ncrs, c->nlev_f, c->lev_f_ptr, c->lev_f_rows, c->nlev_b,
                                           c->lev_b_ptr, c->lev_b_rows, c->rowptr, c->colidx, c->diaind, c->iluDia,
                                           c->iluCrs, vec, src, skip_done ? c->cnt : nullptr);
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}
int launch_ilu_solve(Ctx* c, double* vec) { return ilu_solve_launch(c, vec, nullptr, false); }

static int all_done(Ctx* c, bool* done) {
  CG_CHECK(fetch_counters(c));
  *done = true;
  for (int s = 0; s < c->B; s++)
    if (!c->h_cnt[s * NCNT + CNT_PCG_DONE]) *done = false;
  return CLOTHGPU_OK;
}
static int fetch_solver_result(Ctx* c, double* conv_out, int* iter_out) {
  CG_CUDA(cudaMemcpyAsync(c->h_scal, c->scal, sizeof(double) * NSCAL * c->B, cudaMemcpyDeviceToHost, c->stream));
  CG_CHECK(fetch_counters(c));
  for (int s = 0; s < c->B; s++) {
    if (conv_out) conv_out[s] = c->h_scal[s * NSCAL + SC_CONV];
    if (iter_out) iter_out[s] = c->h_cnt[s * NCNT + CNT_PCG_ITER];
  }
  return CLOTHGPU_OK;
}

// Solve_PCG, ilu_sparse.cpp:628-709.  rPr lives in scal[SC_RPR] / scal[SC_DOT] alternately (old / new).
int run_pcg(Ctx* c, double tol, int max_iter, double* conv_out, int* iter_out) {
  const int n = c->ndof;
  const double tol2 = tol * tol;
  k_pcg_reset<<<c->B, 1, 0, KS(c, "k_pcg_reset")>>>(c->cnt, c->scal, max_iter, tol);
  CG_LAUNCH_CHECK(c);
  CG_CUDA(cudaMemsetAsync(c->vec_x, 0, sizeof(double) * (size_t)c->B * n, c->stream));
  CG_CHECK(dot_launch(c, c->vec_r, c->vec_r, SC_SQ, POST_PCG_INIT, 0, tol2, 0));
  CG_CHECK(ilu_solve_launch(c, c->vec_z, c->vec_r, true));
  k_copy<<<vec_grid(c, n), 256, 0, KS(c, "k_copy")>>>(n, c->vec_z, c->vec_p, c->cnt);
  CG_LAUNCH_CHECK(c);
  int cur = SC_RPR, nxt = SC_DOT;
  CG_CHECK(dot_launch(c, c->vec_r, c->vec_z, cur, POST_STORE, 0, tol2, 1));
  bool done = false;
  for (int it = 0; it < max_iter; it++) {
    {
      dim3 g(div_up(n, 256), c->B);
      k_matvec<<<g, 256, 0, KS(c, "k_matvec")>>>(c->np, c->ncrs, c->rowptr, c->colidx, c->valDia, c->valCrs, 1.0, c->vec_p, 0.0,
                                         c->vec_z, c->cnt);
      CG_LAUNCH_CHECK(c);
    }
    CG_CHECK(dot_launch(c, c->vec_p, c->vec_z, SC_PAP, POST_STORE, 0, tol2, 1));
    {
      DotArgs A;
      A.n = n; A.post = c->exact_dot ? POST_SKIP : POST_PCG_CHECK; A.slot = SC_SQ; A.iter = it; A.tol2 = tol2;
      k_update_xr<<<vec_grid(c, n), 256, 0, KS(c, "k_update_xr")>>>(A, cur, c->vec_z, c->vec_p, c->vec_r, c->vec_x, c->partials,
                                                         c->scal, c->cnt, c->npartial);
      CG_LAUNCH_CHECK(c);
      if (c->exact_dot) CG_CHECK(dot_launch(c, c->vec_r, c->vec_r, SC_SQ, POST_PCG_CHECK, it, tol2, 1));
    }
    // every kernel of the loop skips scenes whose convergence flag is set, so the host only has to look at the flags
    // when the solve is likely over: not before the iteration count of the previous solve minus one
    if (it + 2 >= c->pcg_expect) {
      CG_CHECK(all_done(c, &done));
      if (done) break;
    }
    CG_CHECK(ilu_solve_launch(c, c->vec_z, c->vec_r, true));
    CG_CHECK(dot_launch(c, c->vec_r, c->vec_z, nxt, POST_STORE, 0, tol2, 1));
    k_xpay<<<vec_grid(c, n), 256, 0, KS(c, "k_xpay")>>>(n, nxt, cur, c->vec_z, c->vec_p, c->scal, c->cnt);
    CG_LAUNCH_CHECK(c);
    std::swap(cur, nxt);
  }
  if (!done) CG_CHECK(dot_launch(c, c->vec_r, c->vec_r, SC_SQ, POST_FINAL_CONV, 0, tol2, 1));
  CG_CHECK(fetch_solver_result(c, conv_out, iter_out));
  c->pcg_expect = 0;
  for (int s = 0; s < c->B; s++) c->pcg_expect = std::max(c->pcg_expect, c->h_cnt[s * NCNT + CNT_PCG_ITER] + 1);
  return CLOTHGPU_OK;
}

// Solve_CG, matrix_square_sparse.cpp:338-406: the loop starts at 1 and on convergence reports the PREVIOUS ratio
int run_cg(Ctx* c, double tol, int max_iter, double* conv_out, int* iter_out) {
  const int n = c->ndof;
  const double tol2 = tol * tol;
  k_pcg_reset<<<c->B, 1, 0, KS(c, "k_pcg_reset")>>>(c->cnt, c->scal, max_iter, tol);
  CG_LAUNCH_CHECK(c);
  CG_CUDA(cudaMemsetAsync(c->vec_x, 0, sizeof(double) * (size_t)c->B * n, c->stream));
  CG_CHECK(dot_launch(c, c->vec_r, c->vec_r, SC_SQ, POST_PCG_INIT, 0, tol2, 0));
  k_copy<<<vec_grid(c, n), 256, 0, KS(c, "k_copy")>>>(n, c->vec_r, c->vec_p, c->cnt);
  CG_LAUNCH_CHECK(c);
  bool done = false;
  for (int it = 1; it < max_iter; it++) {
    {
      dim3 g(div_up(n, 256), c->B);
      k_matvec<<<g, 256, 0, KS(c, "k_matvec")>>>(c->np, c->ncrs, c->rowptr, c->colidx, c->valDia, c->valCrs, 1.0, c->vec_p, 0.0,
                                         c->vec_z, c->cnt);
      CG_LAUNCH_CHECK(c);
    }
    CG_CHECK(dot_launch(c, c->vec_p, c->vec_z, SC_PAP, POST_STORE, 0, tol2, 1));
    {
      DotArgs A;
      A.n = n; A.post = c->exact_dot ? POST_SKIP : POST_CG_CHECK; A.slot = SC_DOT; A.iter = it; A.tol2 = tol2;
      k_update_xr<<<vec_grid(c, n), 256, 0, KS(c, "k_update_xr")>>>(A, SC_SQ, c->vec_z, c->vec_p, c->vec_r, c->vec_x, c->partials,
                                                         c->scal, c->cnt, c->npartial);
      CG_LAUNCH_CHECK(c);
      if (c->exact_dot) CG_CHECK(dot_launch(c, c->vec_r, c->vec_r, SC_DOT, POST_CG_CHECK, it, tol2, 1));
    }
    if ((it & 7) == 0 || it == max_iter - 1) {  // unpreconditioned CG runs hundreds of cheap iterations: poll sparsely
      CG_CHECK(all_done(c, &done));
      if (done) break;
    }
    // beta = sq_new / sq ; sq = sq_new ; p = r + beta p
    k_xpay<<<vec_grid(c, n), 256, 0, KS(c, "k_xpay")>>>(n, SC_DOT, SC_SQ, c->vec_r, c->vec_p, c->scal, c->cnt);
    CG_LAUNCH_CHECK(c);
    k_set_scalar_slot<<<c->B, 1, 0, KS(c, "k_set_scalar_slot")>>>(c->scal, SC_SQ, SC_DOT);
    CG_LAUNCH_CHECK(c);
  }
  return fetch_solver_result(c, conv_out, iter_out);
}

// StepTime_InternalDynamicsILU (solve_internal_sparse.h:255-326) / StepTime_InternalDynamics (:185-253)
int run_step_internal(Ctx* c, const clothgpu_params& p, bool use_ilu, double* conv_out, int* iter_out) {
  phase_begin(c, 0);
  CG_CHECK(launch_assemble(c, p, true));
  phase_end(c, 0);
  if (use_ilu) {
    phase_begin(c, 1);
    CG_CHECK(launch_ilu_setvalue(c));
    CG_CHECK(launch_ilu_factor(c));
    phase_end(c, 1);
    phase_begin(c, 2);
    CG_CHECK(run_pcg(c, 1.0e-4, 100, conv_out, iter_out));
    phase_end(c, 2);
  } else {
    phase_begin(c, 2);
    CG_CHECK(run_cg(c, 1.0e-4, 1000, conv_out, iter_out));
    phase_end(c, 2);
  }
  k_advance<<<vec_grid(c, c->ndof), 256, 0, KS(c, "k_advance")>>>(c->ndof, p.dt, c->vec_x, c->xyz, c->uvw);
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}

}  // namespace cg
