// csrc/trisolve.cu -- ILU(0) triangular solves as a strip-pipelined, flag-synchronised wavefront.
//
// Replaces CPreconditionerILU::ForwardSubstitution / BackwardSubstitution (ilu_sparse.cpp:185-215, 295-324).
//
// The sweeps are dependency chains (1021 levels on a 256^2 grid in the reference's ordering, which parity forces
// us to keep), so a single scene is latency bound, not bandwidth bound.  The design removes every global-memory
// round trip and every block barrier from the per-level critical path:
//
//   * rows are cut into NS contiguous "strips" (row ranges; reversed for the backward sweep), one CTA per strip.
//     A row only depends on rows of the same or of an EARLIER strip, so strips form a pipeline: each runs the
//     levels of its own rows in order and lags its predecessor by a few levels.  Strips are handed out through an
//     atomic ticket, so a CTA only ever waits for CTAs that are already running (no co-residency requirement).
//   * inside a strip the rows of one level are grouped into "micro-steps" of <= 10 rows = one warp (3 lanes per
//     row, one per vector component).  A single compute warp walks the micro-steps; consecutive micro-steps are
//     ordered by __syncwarp only.  The strip's slice of the solution vector lives in shared memory.
//   * the matrix blocks a micro-step needs are packed once per factorisation into the exact order the lanes read
//     them (k_pack) and streamed into a shared-memory ring (TS_NST stages of TS_K micro-steps, fixed strides so the
//     lanes address them with immediates) by 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) issued by
//     a loader warp that runs ahead of the compute warp.
//   * the compute warp is software pipelined: a single warp retires ~0.25 instructions per cycle, so the cost of
//     a micro-step is its dynamic instruction count.  Operand slots are stored as byte offsets, padding operands
//     point at a zero row with zero coefficients (no per-lane predicates), and the products of operands that were
//     produced two or more micro-steps ago are issued in the shadow of the previous step's dependent chain.
//   * values produced by other strips ("ghosts") travel through a global exchange vector that is filled with a
//     sentinel bit pattern before every sweep: a publisher warp stores each finished row (plain 8-byte stores, no
//     fence, no flag), an import warp polls the VALUES themselves (one L2 round trip per poll, every leading ghost
//     whose three components have arrived is imported at once) and hands them to the compute warp in shared memory.
//     One round trip instead of flag-then-data and no release fence: 3.7 -> ~1.5 us of lag per strip.
//
// Arithmetic per row is exactly the reference's (same operand order, -fmad=false): results are bit identical.
// The strip count is chosen per topology by simulating the pipeline on the host (build_triplan).
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "ctx.h"

namespace cg {

// ------------------------------------------------------------------------------------------------ host: plan
namespace {

struct Sched {                 // micro-step schedule of one candidate strip size
  int R = 0, ns = 0, nms = 0;
  std::vector<int> strip_ms;   // ns+1
  std::vector<int> ms_rows;    // nms*TS_ROWS (global row ids, -1 = unused)
  std::vector<int> ms_of_row;  // global micro-step index per row
};

}  // namespace

static inline int plan_pos(bool fwd, int np, int row) { return fwd ? row : np - 1 - row; }

// ---- per-micro-step layouts shared by the planner, the verifier and the kernels
// operand offsets: the first TS_MAXD operands of row group q are eight consecutive ints (two LDS.128 per lane), the
// rest (rows with more operands) follow operand-major
__host__ __device__ static inline int ts_slot_index(int d, int q) {
  return d < TS_MAXD ? TS_META_HDR + q * TS_MAXD + d : TS_META_HDR + TS_ROWS * TS_MAXD + (d - TS_MAXD) * TS_ROWS + q;
}
// coefficients: element e of a lane (forward: 0..2 = its D^-1 row, then 3 per operand; backward: 3 per operand) is
// half of a 16-byte unit, units are lane-major: one conflict-free LDS.128 fetches two elements
__host__ __device__ static inline int ts_val_index(int e, int lane) { return ((e >> 1) * 32 + lane) * 2 + (e & 1); }
static inline int ts_meta_ints(int gmax) { return (TS_META_HDR + TS_ROWS * TS_MAXD + std::max(0, gmax - TS_MAXD) * TS_ROWS + 3) / 4 * 4; }
static inline int ts_val_doubles(int gmax, bool fwd) { return ((3 * (gmax + (fwd ? 1 : 0)) + 1) / 2) * 64; }

static void make_sched(int np, const std::vector<int>& lev, bool fwd, int R, Sched& S) {
  S.R = R;
  S.ns = (np + R - 1) / R;
  S.strip_ms.assign(S.ns + 1, 0);
  S.ms_rows.clear();
  S.ms_of_row.assign(np, -1);
  std::vector<int> rows;
  for (int k = 0; k < S.ns; k++) {
    const int p0 = k * R, p1 = std::min(np, p0 + R);
    rows.resize(p1 - p0);
    for (int p = p0; p < p1; p++) rows[p - p0] = plan_pos(fwd, np, p);  // pos -> row is the same map
    std::stable_sort(rows.begin(), rows.end(), [&](int a, int b) { return lev[a] < lev[b]; });
    size_t i = 0;
    while (i < rows.size()) {
      size_t j = i;
      while (j < rows.size() && j - i < (size_t)TS_ROWS && lev[rows[j]] == lev[rows[i]]) j++;
      const int m = (int)(S.ms_rows.size() / TS_ROWS);
      for (size_t q = 0; q < (size_t)TS_ROWS; q++) {
        const int row = (i + q < j) ? rows[i + q] : -1;
        S.ms_rows.push_back(row);
        if (row >= 0) S.ms_of_row[row] = m;
      }
      i = j;
    }
    while ((S.ms_rows.size() / TS_ROWS - S.strip_ms[k]) % TS_K) S.ms_rows.insert(S.ms_rows.end(), TS_ROWS, -1);
    S.strip_ms[k + 1] = (int)(S.ms_rows.size() / TS_ROWS);
  }
  S.nms = S.strip_ms[S.ns];
}

// pipeline makespan in units of one micro-step, with a fixed cross-strip latency `lag`
static double simulate(int np, const int* rowptr, const int* colidx, const int* dia, bool fwd, const Sched& S,
                       double lag, int* max_ghost_out) {
  std::vector<double> fin(S.nms, 0.0);
  std::vector<int> seen(np, -1);
  double makespan = 0;
  int max_ghost = 0;
  for (int k = 0; k < S.ns; k++) {
    double t = 0;
    int nghost = 0;
    for (int m = S.strip_ms[k]; m < S.strip_ms[k + 1]; m++) {
      double start = t;
      for (int q = 0; q < TS_ROWS; q++) {
        const int row = S.ms_rows[(size_t)m * TS_ROWS + q];
        if (row < 0) continue;
        const int j0 = fwd ? rowptr[row] : dia[row], j1 = fwd ? dia[row] : rowptr[row + 1];
        for (int j = j0; j < j1; j++) {
          const int c = colidx[j];
          if (plan_pos(fwd, np, c) / S.R == k) continue;
          start = std::max(start, fin[S.ms_of_row[c]] + lag);
          if (seen[c] != k) { seen[c] = k; nghost++; }
        }
      }
      t = start + 1.0;
      fin[m] = t;
    }
    makespan = std::max(makespan, t);
    max_ghost = std::max(max_ghost, nghost);
  }
  *max_ghost_out = max_ghost;
  return makespan;
}

int build_triplan_host(int np, const int* rowptr, const int* colidx, const int* dia, bool fwd, int force_ns,
                       TriPlanHost& P) {
  const int ncrs = rowptr[np];
  // dependency levels (as in the level-scheduled formulation)
  std::vector<int> lev(np, 0);
  if (fwd) {
    for (int i = 0; i < np; i++) {
      int l = 0;
      for (int j = rowptr[i]; j < dia[i]; j++) l = std::max(l, lev[colidx[j]] + 1);
      lev[i] = l;
    }
  } else {
    for (int i = np - 1; i >= 0; i--) {
      int l = 0;
      for (int j = dia[i]; j < rowptr[i + 1]; j++) l = std::max(l, lev[colidx[j]] + 1);
      lev[i] = l;
    }
  }
  // choose the strip count: simulate the pipeline for a list of candidates, keep the shortest makespan
  int gmax0 = 0;
  for (int i = 0; i < np; i++) gmax0 = std::max(gmax0, fwd ? dia[i] - rowptr[i] : rowptr[i + 1] - dia[i]);
  const long long ring_bytes = (long long)TS_NST * TS_K * (ts_meta_ints(gmax0) * 4 + std::max(64, ts_val_doubles(gmax0, fwd)) * 8) + TS_SMEM_HDR;
  const long long smem_budget = 224 * 1024;
  static const int cand[] = {1, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28, 32, 36, 40, 44, 48, 52, 56,
                             64, 72, 80, 88, 96, 104, 112, 128, 144};
  Sched best;
  double best_t = -1;
  int best_ghost = 0;
  for (int ci = 0; ci < (int)(sizeof(cand) / sizeof(cand[0])); ci++) {
    int ns = force_ns > 0 ? force_ns : cand[ci];
    if (ns > np) ns = np;
    const int R = (np + ns - 1) / ns;
    Sched S;
    make_sched(np, lev, fwd, R, S);
    int mg = 0;
    const double t = simulate(np, rowptr, colidx, dia, fwd, S, 8.0, &mg);
    int longest = 0;
    for (int k2 = 0; k2 < S.ns; k2++) longest = std::max(longest, S.strip_ms[k2 + 1] - S.strip_ms[k2]);
    const long long smem = (long long)(R + mg + 2) * 32 + ring_bytes + (long long)longest * TS_ROWS * 4 + (long long)mg * 12 + 16;
    const bool fits = smem <= smem_budget;
    if (fits && (best_t < 0 || t < best_t)) {
      best_t = t;
      best = std::move(S);
      best_ghost = mg;
    }
    if (force_ns > 0) {
      if (!fits) { set_error("forced strip count does not fit shared memory"); return CLOTHGPU_ERR_ARG; }
      break;
    }
  }
  if (best_t < 0) {  // huge meshes: more strips than any candidate; grow until the strip slice fits
    int ns = 144;
    for (;;) {
      ns *= 2;
      if (ns > np) { set_error("triangular-solve plan: no strip size fits shared memory"); return CLOTHGPU_ERR_CAPACITY; }
      const int R = (np + ns - 1) / ns;
      Sched S;
      make_sched(np, lev, fwd, R, S);
      int mg = 0;
      const double t = simulate(np, rowptr, colidx, dia, fwd, S, 8.0, &mg);
      int longest = 0;
      for (int k2 = 0; k2 < S.ns; k2++) longest = std::max(longest, S.strip_ms[k2 + 1] - S.strip_ms[k2]);
      if ((long long)(R + mg + 2) * 32 + ring_bytes + (long long)longest * TS_ROWS * 4 + (long long)mg * 12 + 16 <= smem_budget) { best_t = t; best = std::move(S); best_ghost = mg; break; }
    }
  }
  const Sched& S = best;
  P.fwd = fwd;
  P.np = np;
  P.ns = S.ns;
  P.R = S.R;
  P.nms = S.nms;
  P.sim_makespan = best_t;
  P.max_ghost = best_ghost;
  P.strip_ms = S.strip_ms;
  P.ms_rows = S.ms_rows;
  P.max_steps = 0;
  for (int k = 0; k < S.ns; k++) P.max_steps = std::max(P.max_steps, S.strip_ms[k + 1] - S.strip_ms[k]);

  // materialise: per micro-step meta block and packed-value source map (fixed strides), ghost lists
  int gmax = 0;  // largest operand count of any row
  for (int i = 0; i < np; i++) gmax = std::max(gmax, fwd ? dia[i] - rowptr[i] : rowptr[i + 1] - dia[i]);
  P.ndmax = gmax;
  P.meta_stride = ts_meta_ints(gmax);                                   // ints per micro-step
  P.val_stride = std::max(64, ts_val_doubles(gmax, fwd));               // doubles per micro-step
  P.meta.assign((size_t)S.nms * P.meta_stride, 0);
  P.pk_src.assign((size_t)S.nms * P.val_stride, -1);
  P.strip_gh.assign(S.ns + 1, 0); P.gh_row.clear(); P.gh_src.clear(); P.gh_need.clear();
  P.nfast = 0;
  std::vector<int> ghost_slot(np, -1), ghost_owner(np, -1);
  // first pass: ghost lists (the row count of the shared-memory vector depends on max_ghost)
  for (int k = 0; k < S.ns; k++) {
    int nghost = 0;
    for (int m = S.strip_ms[k]; m < S.strip_ms[k + 1]; m++)
      for (int q = 0; q < TS_ROWS; q++) {
        const int row = S.ms_rows[(size_t)m * TS_ROWS + q];
        if (row < 0) continue;
        const int j0 = fwd ? rowptr[row] : dia[row], j1 = fwd ? dia[row] : rowptr[row + 1];
        for (int j = j0; j < j1; j++) {
          const int c = colidx[j], cp = plan_pos(fwd, np, c);
          if (cp / S.R == k || ghost_owner[c] == k) continue;
          ghost_owner[c] = k;
          ghost_slot[c] = nghost++;
          P.gh_row.push_back(c);
          P.gh_src.push_back(cp / S.R);
          P.gh_need.push_back(S.ms_of_row[c] - S.strip_ms[cp / S.R] + 1);
        }
      }
    P.strip_gh[k + 1] = (int)P.gh_row.size();
    P.max_ghost = std::max(P.max_ghost, nghost);
  }
  const int zero_off = (S.R + P.max_ghost) * 32, dummy_off = zero_off + 32;  // byte offsets of the two spare rows
  std::fill(ghost_owner.begin(), ghost_owner.end(), -1);
  for (int k = 0; k < S.ns; k++) {
    const int p0 = k * S.R;
    int nghost = 0;
    for (int m = S.strip_ms[k]; m < S.strip_ms[k + 1]; m++) {
      const int* rows = &S.ms_rows[(size_t)m * TS_ROWS];
      int* M = &P.meta[(size_t)m * P.meta_stride];
      int* V = &P.pk_src[(size_t)m * P.val_stride];
      const int lay0 = fwd ? 1 : 0;  // layer 0 of a forward step holds the D^-1 rows
      int ndmax = 0, nearmask = 0;
      for (int q = 0; q <= TS_ROWS; q++) M[4 + q] = dummy_off;
      for (int d = 0; d < std::max(gmax, (int)TS_MAXD); d++)
        for (int q = 0; q < TS_ROWS; q++) M[ts_slot_index(d, q)] = zero_off;
      for (int q = 0; q < TS_ROWS; q++) {
        const int row = rows[q];
        if (row < 0) continue;
        M[4 + q] = (plan_pos(fwd, np, row) - p0) * 32;
        const int j0 = fwd ? rowptr[row] : dia[row], j1 = fwd ? dia[row] : rowptr[row + 1];
        ndmax = std::max(ndmax, j1 - j0);
        for (int d = 0; d < j1 - j0; d++) {
          const int j = j0 + d, c = colidx[j];
          const int cp = plan_pos(fwd, np, c);
          int slot;
          if (cp / S.R == k) {
            slot = cp - p0;
            // an operand produced by the immediately preceding micro-step must be read after that step's barrier
            if (S.ms_of_row[c] == m - 1) nearmask |= 1 << std::min(d, 30);
          } else {
            if (ghost_owner[c] != k) { ghost_owner[c] = k; ghost_slot[c] = nghost++; }  // same order as the first pass
            slot = S.R + ghost_slot[c];
          }
          M[ts_slot_index(d, q)] = slot * 32;
          for (int kk = 0; kk < 3; kk++)
            for (int r = 0; r < 3; r++) V[ts_val_index((lay0 + d) * 3 + kk, q * 3 + r)] = j * 9 + r * 3 + kk;
        }
        if (fwd)
          for (int kk = 0; kk < 3; kk++)
            for (int r = 0; r < 3; r++) V[ts_val_index(kk, q * 3 + r)] = ncrs * 9 + row * 9 + r * 3 + kk;
      }
      // operands beyond TS_MAXD are handled by the un-pipelined tail loop, i.e. always after the barrier
      int fmin = std::min(ndmax, (int)TS_MAXD);
      for (int d = 0; d < std::min(ndmax, (int)TS_MAXD); d++)
        if (nearmask >> d & 1) { fmin = d; break; }
      M[0] = ndmax;
      M[1] = nghost;  // ghosts that must have been imported before this micro-step runs
      M[2] = nearmask & ((1 << TS_MAXD) - 1);
      M[3] = fmin;
      P.nfast += ndmax <= TS_MAXD;
    }
  }
  P.max_meta_bytes = P.meta_stride * 4;
  P.max_val_bytes = P.val_stride * 8;
  P.nval = (long long)P.pk_src.size();
  if (P.pk_src.size() > (size_t)0x7fffffff) { set_error("triangular-solve plan exceeds 2^31 packed values"); return CLOTHGPU_ERR_CAPACITY; }
  return CLOTHGPU_OK;
}

// Structural self-check of a plan (host logic, no GPU): every row is scheduled exactly once, every dependency is
// either an own row of an earlier micro-step (flagged "near" when it is the immediately preceding one) or a ghost
// whose producer strip is earlier and whose import is required in time, padding points at the zero row with no
// matrix value, and every packed value maps to the right matrix entry.
int verify_triplan_host(int np, const int* rowptr, const int* colidx, const int* dia, const TriPlanHost& P) {
  const bool fwd = P.fwd;
  const int ncrs = rowptr[np];
#define PLAN_FAIL(msg) do { set_error(std::string("plan: ") + msg); return CLOTHGPU_ERR_STATE; } while (0)
  std::vector<int> ms_of_row(np, -1);
  if ((long long)P.ms_rows.size() != (long long)P.nms * TS_ROWS) PLAN_FAIL("ms_rows size");
  for (int m = 0; m < P.nms; m++)
    for (int q = 0; q < TS_ROWS; q++) {
      const int row = P.ms_rows[(size_t)m * TS_ROWS + q];
      if (row < 0) continue;
      if (row >= np || ms_of_row[row] != -1) PLAN_FAIL("row scheduled twice or out of range");
      ms_of_row[row] = m;
    }
  for (int i = 0; i < np; i++)
    if (ms_of_row[i] < 0) PLAN_FAIL("row not scheduled");
  if ((long long)P.meta.size() != (long long)P.nms * P.meta_stride || (long long)P.pk_src.size() != (long long)P.nms * P.val_stride ||
      P.meta_stride % 4 || P.val_stride % 2)
    PLAN_FAIL("array sizes / strides");
  const int zero_off = (P.R + P.max_ghost) * 32, dummy_off = zero_off + 32, lay0 = fwd ? 1 : 0;
  for (int k = 0; k < P.ns; k++) {
    const int p0 = k * P.R, g0 = P.strip_gh[k], ng = P.strip_gh[k + 1] - g0;
    if ((P.strip_ms[k + 1] - P.strip_ms[k]) % TS_K) PLAN_FAIL("strip length not a multiple of the stage size");
    int prev_need = 0;
    for (int m = P.strip_ms[k]; m < P.strip_ms[k + 1]; m++) {
      const int* M = &P.meta[(size_t)m * P.meta_stride];
      const int* V = &P.pk_src[(size_t)m * P.val_stride];
      const int ndmax = M[0], gneed = M[1], nearmask = M[2], fmin = M[3];
      if (gneed < prev_need || gneed > ng) PLAN_FAIL("ghost requirement not monotone");
      prev_need = gneed;
      if (ndmax > P.ndmax || fmin < 0 || fmin > std::min(ndmax, (int)TS_MAXD)) PLAN_FAIL("operand count / chain split");
      for (int d = 0; d < fmin; d++)
        if (nearmask >> d & 1) PLAN_FAIL("near operand inside the early chain prefix");
      std::vector<char> used((size_t)P.val_stride, 0);
      for (int q = 0; q < TS_ROWS; q++) {
        const int row = P.ms_rows[(size_t)m * TS_ROWS + q];
        const int nd = row < 0 ? 0 : (fwd ? dia[row] - rowptr[row] : rowptr[row + 1] - dia[row]);
        if (row < 0) {
          if (M[4 + q] != dummy_off) PLAN_FAIL("unused lane must write the dummy row");
        } else {
          const int pos = plan_pos(fwd, np, row);
          if (pos / P.R != k || M[4 + q] != (pos - p0) * 32) PLAN_FAIL("own slot");
          if (nd > ndmax) PLAN_FAIL("dependency count");
        }
        const int j0 = row < 0 ? 0 : (fwd ? rowptr[row] : dia[row]);
        for (int d = 0; d < P.ndmax; d++) {
          const int off = M[ts_slot_index(d, q)];
          if (d >= nd) {
            if (off != zero_off) PLAN_FAIL("padding operand must read the zero row");
            continue;
          }
          const int c = colidx[j0 + d], cp = plan_pos(fwd, np, c);
          if (cp / P.R == k) {
            if (off != (cp - p0) * 32 || ms_of_row[c] >= m) PLAN_FAIL("own dependency not earlier");
            if (ms_of_row[c] == m - 1 && d < TS_MAXD && !(nearmask >> d & 1)) PLAN_FAIL("near operand not flagged");
          } else {
            const int gi = off / 32 - P.R;
            if (cp / P.R > k || gi < 0 || gi >= gneed || P.gh_row[g0 + gi] != c || P.gh_src[g0 + gi] != cp / P.R ||
                P.gh_need[g0 + gi] != ms_of_row[c] - P.strip_ms[cp / P.R] + 1)
              PLAN_FAIL("ghost dependency");
          }
          for (int kk = 0; kk < 3; kk++)
            for (int r = 0; r < 3; r++) {
              const int idx = ts_val_index((lay0 + d) * 3 + kk, q * 3 + r);
              if (V[idx] != (j0 + d) * 9 + r * 3 + kk) PLAN_FAIL("packed value map");
              used[idx] = 1;
            }
        }
        if (fwd && row >= 0)
          for (int kk = 0; kk < 3; kk++)
            for (int r = 0; r < 3; r++) {
              const int idx = ts_val_index(kk, q * 3 + r);
              if (V[idx] != ncrs * 9 + row * 9 + r * 3 + kk) PLAN_FAIL("packed diagonal map");
              used[idx] = 1;
            }
      }
      if (M[4 + TS_ROWS] != dummy_off) PLAN_FAIL("idle lanes must write the dummy row");
      for (int i = 0; i < P.val_stride; i++)
        if (!used[i] && V[i] != -1) PLAN_FAIL("padding value must be empty");
    }
  }
#undef PLAN_FAIL
  return CLOTHGPU_OK;
}

template <class T>
static int up(T** p, const std::vector<T>& h) {
  *p = nullptr;
  CG_CUDA(cudaMalloc((void**)p, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) CG_CUDA(cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return CLOTHGPU_OK;
}

int upload_triplan(Ctx* c, const TriPlanHost& H, TriPlan& D) {
  D.fwd = H.fwd; D.ns = H.ns; D.R = H.R; D.nms = H.nms; D.max_meta_bytes = H.max_meta_bytes;
  D.max_val_bytes = H.max_val_bytes; D.max_ghost = H.max_ghost; D.max_steps = H.max_steps; D.nval = H.nval;
  D.sim_makespan = H.sim_makespan; D.meta_stride = H.meta_stride; D.val_stride = H.val_stride; D.ndmax = H.ndmax;
  CG_CHECK(up(&D.strip_ms, H.strip_ms)); CG_CHECK(up(&D.strip_gh, H.strip_gh)); CG_CHECK(up(&D.ms_rows, H.ms_rows));
  CG_CHECK(up(&D.meta, H.meta));
  CG_CHECK(up(&D.pk_src, H.pk_src)); CG_CHECK(up(&D.gh_row, H.gh_row)); CG_CHECK(up(&D.gh_src, H.gh_src));
  CG_CHECK(up(&D.gh_need, H.gh_need));
  CG_CUDA(cudaMalloc((void**)&D.packed, sizeof(double) * (size_t)std::max<long long>(H.nval, 1) * c->B));
  CG_CUDA(cudaMalloc((void**)&D.progress, sizeof(unsigned long long) * (size_t)H.ns * c->B));
  CG_CUDA(cudaMemset(D.progress, 0, sizeof(unsigned long long) * (size_t)H.ns * c->B));
  CG_CUDA(cudaMalloc((void**)&D.ticket, sizeof(unsigned long long)));
  CG_CUDA(cudaMemset(D.ticket, 0, sizeof(unsigned long long)));
  D.pbase = 0;
  D.tbase = 0;
  D.smem_bytes = TS_SMEM_HDR + (D.R + D.max_ghost + 2) * 32 + TS_NST * TS_K * (D.meta_stride * 4 + D.val_stride * 8) +
                 D.max_steps * TS_ROWS * 4 + D.max_ghost * 12 + 16;
  return CLOTHGPU_OK;
}

void free_triplan(TriPlan& D) {
  void* ptrs[] = {D.strip_ms, D.strip_gh, D.ms_rows, D.meta, D.pk_src, D.gh_row, D.gh_src,
                  D.gh_need, D.packed, D.progress, D.ticket};
  for (void* p : ptrs) cudaFree(p);
  D = TriPlan();
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t a = smem_u32(bar);
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cta(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// "not yet produced" marker of the exchange vector: a NaN payload no arithmetic generates; a solution component
// that happens to carry it (propagated from such an input) is exchanged as the canonical NaN instead
#define TS_SENTINEL 0xFFFFFFFFFFFFFFFFull
#define TS_CANON_NAN 0x7FF8000000000000ull

struct TriArgs {
  int np, ns, R, B;
  int meta_stride, val_stride;  // ints / doubles per micro-step
  int xs_rows;                  // rows of the shared-memory vector (own + ghosts + zero row + dummy row)
  int max_steps;                // longest strip, in micro-steps
  long long nval;
  const int *strip_ms, *strip_gh, *ms_rows, *gh_row, *gh_src, *gh_need;
  const int* meta;
  const double* packed;
  unsigned long long *progress, *ticket;
  unsigned long long pbase, tbase;
  unsigned long long* xch;      // B x np x 3: exchange vector (sentinel-filled before the launch)
  const double* src;
  double* vec;
  const int* cnt;
  unsigned long long* dbg;  // optional: 8 counters per CTA (clothgpu_debug_trisolve)
};

__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ------------------------------------------------------------------------------------------------ kernels
// gather the factor's blocks into the order the solve lanes read them (once per factorisation)
__global__ void __launch_bounds__(256) k_pack(long long nval, const int* __restrict__ pk_src, int ncrs9, int np9,
                                              const double* __restrict__ dia, const double* __restrict__ crs,
                                              double* __restrict__ packed) {
  const int s = blockIdx.y;
  const double* d = dia + (size_t)s * np9;
  const double* c = crs + (size_t)s * ncrs9;
  double* o = packed + (size_t)s * nval;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nval; i += (long long)gridDim.x * blockDim.x) {
    const int src = pk_src[i];
    o[i] = src < 0 ? 0.0 : (src < ncrs9 ? c[src] : d[src - ncrs9]);
  }
}

// one CTA per (strip, scene): warp 0 computes, warp 1 streams matrix chunks, warp 2 publishes, warp 3 imports
template <bool FWD, int ND, bool DBG, int NCW>
__global__ void __launch_bounds__((NCW + 3) * 32, 1) k_trisolve(const TriArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + TS_NST;
  uint32_t* flags = reinterpret_cast<uint32_t*>(smem + 128);  // [0] micro-steps done, [1] ghosts imported, [2] ticket
  unsigned char* xsc = smem + TS_SMEM_HDR;                    // the strip's slice of the vector, 32 bytes per row
  double* xs = reinterpret_cast<double*>(xsc);
  const int meta_bytes = A.meta_stride * 4, val_bytes = A.val_stride * 8;
  const int stage_bytes = TS_K * (meta_bytes + val_bytes);
  unsigned char* stages = xsc + (size_t)A.xs_rows * 32;
  int* s_rows = reinterpret_cast<int*>(stages + (size_t)TS_NST * stage_bytes);  // row ids of the strip's micro-steps
  int* s_gh = s_rows + A.max_steps * TS_ROWS;                                    // ghost descriptors {row, src, need}
  constexpr int LAY0 = FWD ? 1 : 0;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    const unsigned long long t = atomicAdd(A.ticket, 1ULL) - A.tbase;
    flags[2] = (uint32_t)t;
  }
  __syncthreads();
  const uint32_t ticket = flags[2];
  const int strip = (int)(ticket / (uint32_t)A.B), scene = (int)(ticket % (uint32_t)A.B);
  if (A.cnt && A.cnt[scene * NCNT + CNT_PCG_DONE]) return;

  const int ms0 = A.strip_ms[strip], nsteps = A.strip_ms[strip + 1] - ms0;  // a multiple of TS_K
  const int p0 = strip * A.R, nown = min(A.R, A.np - p0);
  double* out = A.vec + (size_t)scene * A.np * 3;
  const double* in = (FWD && A.src) ? A.src + (size_t)scene * A.np * 3 : out;

  if (tid < 8) xs[(A.xs_rows - 2) * 4 + tid] = 0.0;  // zero row (padding operands) and dummy row (idle lanes)
  if (tid == 0) {
    for (int i = 0; i < TS_NST; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    flags[0] = 0;
    flags[1] = 0;
    flags[3] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // Start-up, off the loader's path: the loader warp starts streaming matrix chunks at once; the other warps load
  // the strip's slice of the vector (right-hand side now, solution as the sweep proceeds) with eight loads in flight
  // per thread and meet at a named barrier; the publisher / import warps then stage their own topology lists.
  if (warp != NCW) {
    const int nth = (NCW + 2) * 32, t2 = warp < NCW ? tid : tid - 32;
    const int n3 = nown * 3;
    for (int i0 = t2; i0 < n3; i0 += nth * 8) {
      double v[8];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int idx = i0 + j * nth;
        const int p = p0 + idx / 3, comp = idx - (idx / 3) * 3;
        const int row = FWD ? p : A.np - 1 - p;
        v[j] = idx < n3 ? in[(size_t)row * 3 + comp] : 0.0;
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int idx = i0 + j * nth;
        if (idx < n3) xs[(idx / 3) * 4 + (idx - (idx / 3) * 3)] = v[j];
      }
    }
    asm volatile("bar.sync 3, %0;" ::"r"(nth) : "memory");
  }

  if (NCW == 2 && warp < 2) {
    // ---------------------------------------------------------------- two compute warps, alternating micro-steps
    // Warp (s & 1) owns micro-step s.  While the other warp runs the dependent tail of step s-1, this warp already
    // fetches its coefficients, waits for its ghosts and computes the products of the operands that are at least two
    // micro-steps old ("phase A"); after the hand-off (named barrier: arrive by the finisher of s-1, sync here) it
    // only loads the NB operands that may come from step s-1, finishes the chain, applies D^-1 and stores
    // ("phase B").  Forward sweeps have their near operands at the end of the chain (highest columns), backward
    // sweeps at the beginning; a step whose near operands do not fit that shape runs entirely after the hand-off.
    const int cw = warp;
    constexpr int NB = FWD ? (ND < 3 ? ND : 3) : (ND < 2 ? ND : 2);
    const int g = lane / 3, r = lane - g * 3;
    const int gq = g < TS_ROWS ? g : TS_ROWS - 1;
    const int l0 = (gq * 3) & 31;
    uint32_t seen = 0;
    constexpr int NE = 3 * (ND + LAY0);
    int myoff = 0, off[ND], ndmax = 0, nearmask = 0, fmin = 0;
    uint32_t gneed = 0;
    double av[ND][3], dd0 = 0.0, dd1 = 0.0, dd2 = 0.0;
#define TS_LOAD_STEP2(MP_, VP_)                                                                  \
    {                                                                                            \
      const int* Mp = (MP_);                                                                     \
      const double2* Vp = reinterpret_cast<const double2*>(VP_) + lane;                          \
      const int4 h0 = *reinterpret_cast<const int4*>(Mp);                                        \
      ndmax = h0.x; gneed = (uint32_t)h0.y; nearmask = h0.z; fmin = h0.w;                        \
      myoff = Mp[4 + g] + r * 8;                                                                 \
      {                                                                                          \
        const int4 s0 = *reinterpret_cast<const int4*>(Mp + TS_META_HDR + gq * TS_MAXD);         \
        const int so[4] = {s0.x, s0.y, s0.z, s0.w};                                              \
        _Pragma("unroll") for (int d = 0; d < 4 && d < ND; d++) off[d] = so[d];                  \
        if (ND > 4) {                                                                            \
          const int4 s1 = *reinterpret_cast<const int4*>(Mp + TS_META_HDR + gq * TS_MAXD + 4);   \
          const int so1[4] = {s1.x, s1.y, s1.z, s1.w};                                           \
          _Pragma("unroll") for (int d = 4; d < ND; d++) off[d] = so1[d - 4];                    \
        }                                                                                        \
      }                                                                                          \
      double cv[NE + 1];                                                                         \
      _Pragma("unroll") for (int u = 0; u < (NE + 1) / 2; u++) {                                 \
        const double2 c2 = Vp[u * 32];                                                           \
        cv[2 * u] = c2.x; cv[2 * u + 1] = c2.y;                                                  \
      }                                                                                          \
      _Pragma("unroll") for (int d = 0; d < ND; d++) {                                           \
        av[d][0] = cv[(LAY0 + d) * 3 + 0]; av[d][1] = cv[(LAY0 + d) * 3 + 1]; av[d][2] = cv[(LAY0 + d) * 3 + 2]; \
      }                                                                                          \
      if (FWD) { dd0 = cv[0]; dd1 = cv[1]; dd2 = cv[2]; }                                        \
    }
#define TS_PRODUCT(D_)                                                                           \
    {                                                                                            \
      const double2 x01 = *reinterpret_cast<const double2*>(xsc + off[D_]);                      \
      const double x2 = *reinterpret_cast<const double*>(xsc + off[D_] + 16);                    \
      pr[D_] = av[D_][0] * x01.x + av[D_][1] * x01.y + av[D_][2] * x2;                           \
    }
    const bool dbg2 = DBG && cw == 0;
    unsigned long long t_begin2 = 0, w_a = 0, w_wait = 0, w_b = 0, w_pre = 0;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (dbg2) t_begin2 = gtime_ns();
    if (cw < nsteps) {
      mbar_wait(&full[0], 0);
      TS_LOAD_STEP2(reinterpret_cast<const int*>(stages + cw * meta_bytes),
                    reinterpret_cast<const double*>(stages + TS_K * meta_bytes + cw * val_bytes))
    }
    for (int s = cw; s < nsteps; s += 2) {
      const int sg = s / TS_K, k = s - sg * TS_K, st = sg % TS_NST;
      const unsigned char* sb = stages + (size_t)st * stage_bytes;
      if (gneed > seen) {
        do { seen = ld_acquire_cta(&flags[1]); } while (seen < gneed);
      }
      const bool conform = ndmax <= ND && (FWD ? fmin >= ND - NB : (nearmask >> NB) == 0);
      if (dbg2) c0 = clock64();
      double t = *reinterpret_cast<const double*>(xsc + myoff);
      double pr[ND];
      // ---- phase A: everything that cannot depend on micro-step s-1
      if (conform) {
        if (FWD) {
#pragma unroll
          for (int d = 0; d < ND - NB; d++) TS_PRODUCT(d)
#pragma unroll
          for (int d = 0; d < ND - NB; d++) t -= pr[d];
        } else {
#pragma unroll
          for (int d = NB; d < ND; d++) TS_PRODUCT(d)
        }
      }
      // ---- hand-off: micro-step s-1 (other warp) has stored its rows
      if (dbg2) c1 = clock64();
      if (s > 0) asm volatile("bar.sync %0, 64;" ::"r"(1 + ((s - 1) & 1)) : "memory");
      if (dbg2) c2 = clock64();
      // ---- phase B
      if (conform) {
        if (FWD) {
#pragma unroll
          for (int d = ND - NB; d < ND; d++) TS_PRODUCT(d)
#pragma unroll
          for (int d = ND - NB; d < ND; d++) t -= pr[d];
        } else {
#pragma unroll
          for (int d = 0; d < NB; d++) TS_PRODUCT(d)
#pragma unroll
          for (int d = 0; d < ND; d++) t -= pr[d];
        }
      } else {
#pragma unroll
        for (int d = 0; d < ND; d++) TS_PRODUCT(d)
#pragma unroll
        for (int d = 0; d < ND; d++) t -= pr[d];
        if (ndmax > ND) {  // rows with more operands than the unrolled body holds (ND == TS_MAXD only)
          const int* Mk = reinterpret_cast<const int*>(sb + k * meta_bytes);
          const double* Vk = reinterpret_cast<const double*>(sb + TS_K * meta_bytes + k * val_bytes);
          for (int d = ND; d < ndmax; d++) {
            const int o = Mk[ts_slot_index(d, gq)];
            const double a0 = Vk[ts_val_index((LAY0 + d) * 3 + 0, lane)], a1 = Vk[ts_val_index((LAY0 + d) * 3 + 1, lane)],
                         a2 = Vk[ts_val_index((LAY0 + d) * 3 + 2, lane)];
            const double x0 = *reinterpret_cast<const double*>(xsc + o), x1 = *reinterpret_cast<const double*>(xsc + o + 8),
                         x2 = *reinterpret_cast<const double*>(xsc + o + 16);
            t -= a0 * x0 + a1 * x1 + a2 * x2;
          }
        }
      }
      if (FWD) {  // v_i = D_i^-1 t
        const double t0 = __shfl_sync(0xffffffffu, t, l0);
        const double t1 = __shfl_sync(0xffffffffu, t, (l0 + 1) & 31);
        const double t2 = __shfl_sync(0xffffffffu, t, (l0 + 2) & 31);
        t = dd0 * t0 + dd1 * t1 + dd2 * t2;
      }
      *reinterpret_cast<double*>(xsc + myoff) = t;
      __syncwarp();
      if (s + 1 < nsteps) asm volatile("bar.arrive %0, 64;" ::"r"(1 + (s & 1)) : "memory");
      if (lane == 0) {
        // each compute warp has its own monotone "done" word (the publisher takes the maximum): a shared word could
        // be overwritten with an older value by the slower of the two writers
        st_release_cta(&flags[cw == 0 ? 0 : 3], (uint32_t)(s + 1));
        if (k == TS_K - 1) mbar_arrive(&empty[st]);
      }
      if (dbg2) c3 = clock64();
      // ---- fetch the image of this warp's next micro-step (s + 2), possibly from the next stage
      if (s + 2 < nsteps) {
        const int s2 = s + 2, sg2 = s2 / TS_K, k2 = s2 - sg2 * TS_K;
        if (sg2 != sg) mbar_wait(&full[sg2 % TS_NST], (uint32_t)(sg2 / TS_NST) & 1u);
        const unsigned char* nb2 = stages + (size_t)(sg2 % TS_NST) * stage_bytes;
        TS_LOAD_STEP2(reinterpret_cast<const int*>(nb2 + k2 * meta_bytes),
                      reinterpret_cast<const double*>(nb2 + TS_K * meta_bytes + k2 * val_bytes))
      }
      if (dbg2) {
        w_a += (unsigned long long)(c1 - c0); w_wait += (unsigned long long)(c2 - c1);
        w_b += (unsigned long long)(c3 - c2); w_pre += (unsigned long long)(clock64() - c3);
      }
    }
    if (dbg2 && lane == 0) {  // per step of THIS warp (it owns every other micro-step): phase A, hand-off wait, phase B, prefetch
      unsigned long long* o = A.dbg + (size_t)blockIdx.x * 8;
      o[0] = t_begin2; o[1] = gtime_ns(); o[2] = w_a; o[3] = w_wait; o[4] = w_b; o[5] = w_pre;
      o[6] = (unsigned long long)((nsteps + 1) / 2); o[7] = (unsigned long long)strip;
    }
#undef TS_PRODUCT
#undef TS_LOAD_STEP2
  } else if (NCW == 1 && warp == 0) {
    // ---------------------------------------------------------------- compute warp
    // A single warp retires ~0.25 dependent instructions per cycle, so a micro-step costs what its longest
    // dependent chain costs -- provided the independent work is in ONE basic block for the scheduler to overlap.
    // Hence: the operand count ND is a template parameter, every row executes all ND products (padding operands
    // read the zero row with zero coefficients: exact), and there is no branch between the barrier of the
    // previous micro-step and the store of this one except the (rarely taken) ghost wait.
    const int g = lane / 3, r = lane - g * 3;
    const int gq = g < TS_ROWS ? g : TS_ROWS - 1;
    const int l0 = (gq * 3) & 31;
    uint32_t seen = 0;
    constexpr bool dbg = DBG;  // cycle counters are compiled out of the production instantiation
    unsigned long long t_begin = 0, w_mbar = 0, w_ghost = 0, w_comp = 0, w_sig = 0;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (dbg) t_begin = gtime_ns();
    // register image of the micro-step about to run: operand offsets and coefficients do not depend on the
    // solution, so they are fetched one micro-step ahead, in the shadow of the previous step's subtraction chain
    constexpr int NE = 3 * (ND + LAY0);  // coefficients per lane and micro-step
    int myoff = 0, off[ND];
    uint32_t gneed = 0;
    int ndmax = 0;
    double av[ND][3], dd0 = 0.0, dd1 = 0.0, dd2 = 0.0;
#define TS_LOAD_STEP(MP_, VP_)                                                                   \
    {                                                                                            \
      const int* Mp = (MP_);                                                                     \
      const double2* Vp = reinterpret_cast<const double2*>(VP_) + lane;                          \
      gneed = (uint32_t)Mp[1];                                                                   \
      ndmax = Mp[0];                                                                             \
      myoff = Mp[4 + g] + r * 8;                                                                 \
      {                                                                                          \
        const int4 s0 = *reinterpret_cast<const int4*>(Mp + TS_META_HDR + gq * TS_MAXD);         \
        const int so[4] = {s0.x, s0.y, s0.z, s0.w};                                              \
        _Pragma("unroll") for (int d = 0; d < 4 && d < ND; d++) off[d] = so[d];                  \
        if (ND > 4) {                                                                            \
          const int4 s1 = *reinterpret_cast<const int4*>(Mp + TS_META_HDR + gq * TS_MAXD + 4);   \
          const int so1[4] = {s1.x, s1.y, s1.z, s1.w};                                           \
          _Pragma("unroll") for (int d = 4; d < ND; d++) off[d] = so1[d - 4];                    \
        }                                                                                        \
      }                                                                                          \
      double cv[NE + 1];                                                                         \
      _Pragma("unroll") for (int u = 0; u < (NE + 1) / 2; u++) {                                 \
        const double2 c2 = Vp[u * 32];                                                           \
        cv[2 * u] = c2.x; cv[2 * u + 1] = c2.y;                                                  \
      }                                                                                          \
      _Pragma("unroll") for (int d = 0; d < ND; d++) {                                           \
        av[d][0] = cv[(LAY0 + d) * 3 + 0]; av[d][1] = cv[(LAY0 + d) * 3 + 1]; av[d][2] = cv[(LAY0 + d) * 3 + 2]; \
      }                                                                                          \
      if (FWD) { dd0 = cv[0]; dd1 = cv[1]; dd2 = cv[2]; }                                        \
    }
    const int nstage = nsteps / TS_K;
    if (nstage > 0) {
      if (dbg) c0 = clock64();
      mbar_wait(&full[0], 0);
      if (dbg) w_mbar += (unsigned long long)(clock64() - c0);
      TS_LOAD_STEP(reinterpret_cast<const int*>(stages), reinterpret_cast<const double*>(stages + TS_K * meta_bytes))
    }
    for (int sg = 0; sg < nstage; sg++) {
      const int st = sg % TS_NST;
      const unsigned char* sb = stages + (size_t)st * stage_bytes;
      const unsigned char* nb = stages + (size_t)((sg + 1) % TS_NST) * stage_bytes;
      const bool more_stages = sg + 1 < nstage;
#pragma unroll
      for (int k = 0; k < TS_K; k++) {
        const int* Mk = reinterpret_cast<const int*>(sb + k * meta_bytes);
        const double* Vk = reinterpret_cast<const double*>(sb + TS_K * meta_bytes + k * val_bytes);
        const bool last = (k == TS_K - 1);
        if (dbg) c0 = clock64();
        if (last && more_stages) mbar_wait(&full[(sg + 1) % TS_NST], (uint32_t)((sg + 1) / TS_NST) & 1u);
        if (dbg) c1 = clock64();
        if (gneed > seen) {
          do { seen = ld_acquire_cta(&flags[1]); } while (seen < gneed);
        }
        if (dbg) { c2 = clock64(); w_mbar += (unsigned long long)(c1 - c0); w_ghost += (unsigned long long)(c2 - c1); }
        const int cur_off = myoff, cur_ndmax = ndmax;
        const double e0 = dd0, e1 = dd1, e2 = dd2;
        double t = *reinterpret_cast<const double*>(xsc + cur_off);
        double pr[ND];
#pragma unroll
        for (int d = 0; d < ND; d++) {
          const double2 x01 = *reinterpret_cast<const double2*>(xsc + off[d]);
          const double x2 = *reinterpret_cast<const double*>(xsc + off[d] + 16);
          pr[d] = av[d][0] * x01.x + av[d][1] * x01.y + av[d][2] * x2;
        }
        // fetch the next micro-step's image (same stage, or the first step of the next stage)
        if (!last) {
          TS_LOAD_STEP(reinterpret_cast<const int*>(sb + (k + 1) * meta_bytes),
                       reinterpret_cast<const double*>(sb + TS_K * meta_bytes + (k + 1) * val_bytes))
        } else if (more_stages) {
          TS_LOAD_STEP(reinterpret_cast<const int*>(nb), reinterpret_cast<const double*>(nb + TS_K * meta_bytes))
        }
#pragma unroll
        for (int d = 0; d < ND; d++) t -= pr[d];
        if (cur_ndmax > ND) {  // rows with more operands than the unrolled body holds (ND == TS_MAXD only)
          for (int d = ND; d < cur_ndmax; d++) {
            const int o = Mk[ts_slot_index(d, gq)];
            const double a0 = Vk[ts_val_index((LAY0 + d) * 3 + 0, lane)], a1 = Vk[ts_val_index((LAY0 + d) * 3 + 1, lane)],
                         a2 = Vk[ts_val_index((LAY0 + d) * 3 + 2, lane)];
            const double x0 = *reinterpret_cast<const double*>(xsc + o), x1 = *reinterpret_cast<const double*>(xsc + o + 8),
                         x2 = *reinterpret_cast<const double*>(xsc + o + 16);
            t -= a0 * x0 + a1 * x1 + a2 * x2;
          }
        }
        if (FWD) {  // v_i = D_i^-1 t
          const double t0 = __shfl_sync(0xffffffffu, t, l0);
          const double t1 = __shfl_sync(0xffffffffu, t, (l0 + 1) & 31);
          const double t2 = __shfl_sync(0xffffffffu, t, (l0 + 2) & 31);
          t = e0 * t0 + e1 * t1 + e2 * t2;
        }
        *reinterpret_cast<double*>(xsc + cur_off) = t;
        __syncwarp();
        if (dbg) c3 = clock64();
        if (lane == 0) st_release_cta(&flags[0], (uint32_t)(sg * TS_K + k + 1));
        if (last && lane == 0) mbar_arrive(&empty[st]);
        if (dbg) {
          __syncwarp();
          w_comp += (unsigned long long)(c3 - c2);
          w_sig += (unsigned long long)(clock64() - c3);
        }
      }
    }
#undef TS_LOAD_STEP
    if (dbg && lane == 0) {
      unsigned long long* o = A.dbg + (size_t)blockIdx.x * 8;
      o[0] = t_begin; o[1] = gtime_ns(); o[2] = w_mbar; o[3] = w_ghost; o[4] = w_comp; o[5] = w_sig;
      o[6] = (unsigned long long)nsteps; o[7] = (unsigned long long)strip;
    }
  } else if (warp == NCW) {
    // ---------------------------------------------------------------- loader warp: TMA bulk copies, TS_NST stages ahead
    if (lane == 0) {
      const unsigned char* meta = reinterpret_cast<const unsigned char*>(A.meta) + (size_t)ms0 * meta_bytes;
      const unsigned char* packed = reinterpret_cast<const unsigned char*>(A.packed + (size_t)scene * A.nval) + (size_t)ms0 * val_bytes;
      const int nstage = nsteps / TS_K;
      for (int sg = 0; sg < nstage; sg++) {
        const int st = sg % TS_NST;
        if (sg >= TS_NST) mbar_wait(&empty[st], (uint32_t)(sg / TS_NST - 1) & 1u);
        unsigned char* dst = stages + (size_t)st * stage_bytes;
        mbar_expect_tx(&full[st], (uint32_t)stage_bytes);
        bulk_g2s(dst, meta + (size_t)sg * TS_K * meta_bytes, (uint32_t)(TS_K * meta_bytes), &full[st]);
        bulk_g2s(dst + TS_K * meta_bytes, packed + (size_t)sg * TS_K * val_bytes, (uint32_t)(TS_K * val_bytes), &full[st]);
      }
    }
  } else if (warp == NCW + 1) {
    // ---------------------------------------------------------------- publisher warp
    // copies finished rows to the result vector and to the exchange vector (consumed by the import warps of later
    // strips); the row ids of the next four micro-steps are prefetched while it waits
    unsigned long long* xch = A.xch + (size_t)scene * A.np * 3;
    const int q3 = lane / 3, qc = lane - q3 * 3;
    const bool on = lane < 30;
    for (int i = lane; i < nsteps * TS_ROWS; i += 32) s_rows[i] = A.ms_rows[(size_t)ms0 * TS_ROWS + i];
    __syncwarp();
    uint32_t pub = 0;
    while (pub < (uint32_t)nsteps) {
      int rows[4];
#pragma unroll
      for (int j = 0; j < 4; j++)
        rows[j] = (on && pub + j < (uint32_t)nsteps) ? s_rows[(pub + j) * TS_ROWS + q3] : -1;
      uint32_t d;
      while ((d = max(ld_acquire_cta(&flags[0]), ld_acquire_cta(&flags[3]))) <= pub) __nanosleep(20);
      d = min(d, pub + 4u);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        if (pub + j < d && rows[j] >= 0) {
          const int p = FWD ? rows[j] : A.np - 1 - rows[j];
          const double v = xs[(p - p0) * 4 + qc];
          unsigned long long bits = (unsigned long long)__double_as_longlong(v);
          if (bits == TS_SENTINEL) bits = TS_CANON_NAN;
          st_relaxed_gpu(xch + (size_t)rows[j] * 3 + qc, bits);
          out[(size_t)rows[j] * 3 + qc] = v;
        }
      }
      pub = d;
    }
  } else {
    // ---------------------------------------------------------------- import warp
    // sliding window over the strip's ghost list (ordered by first use): every poll reads the three components of
    // up to 32 ghosts from the exchange vector; all leading ghosts that have arrived are imported at once
    const int g0 = A.strip_gh[strip], ngh = A.strip_gh[strip + 1] - g0;
    const unsigned long long* xch = A.xch + (size_t)scene * A.np * 3;
    for (int i = lane; i < ngh; i += 32) s_gh[i * 3] = A.gh_row[g0 + i];  // only the row ids are needed here
    __syncwarp();
    int w0 = 0;
    while (w0 < ngh) {
      const int gi = w0 + lane;
      const bool in = gi < ngh;
      const unsigned long long* px = xch + (size_t)(in ? s_gh[gi * 3] : 0) * 3;
      unsigned long long v0 = 0, v1 = 0, v2 = 0;
      int nlead = 0;
      do {
        if (in) { v0 = ld_relaxed_gpu(px); v1 = ld_relaxed_gpu(px + 1); v2 = ld_relaxed_gpu(px + 2); }
        const bool ready = in && v0 != TS_SENTINEL && v1 != TS_SENTINEL && v2 != TS_SENTINEL;
        const unsigned m = __ballot_sync(0xffffffffu, ready);
        nlead = __ffs(~m) - 1;  // number of leading ready ghosts (32 if all)
        if (nlead < 0) nlead = 32;
      } while (nlead == 0);
      if (lane < nlead) {
        double* dst = xs + (size_t)(A.R + gi) * 4;
        dst[0] = __longlong_as_double((long long)v0); dst[1] = __longlong_as_double((long long)v1);
        dst[2] = __longlong_as_double((long long)v2);
      }
      __syncwarp();
      w0 += nlead;
      if (lane == 0) st_release_cta(&flags[1], (uint32_t)w0);
    }
  }
}

// ------------------------------------------------------------------------------------------------ launchers
static int pack_plan(Ctx* c, TriPlan& P) {
  if (P.nval <= 0) return CLOTHGPU_OK;
  dim3 g(std::min<long long>((P.nval + 255) / 256, 148 * 16), c->B);
  k_pack<<<g, 256, 0, KS(c, "k_pack")>>>(P.nval, P.pk_src, c->ncrs * 9, c->np * 9, c->iluDia, c->iluCrs, P.packed);
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}

template <bool FWD, int ND>
static int launch_tri_nd(Ctx* c, const TriPlan& P, const TriArgs& A, int grid) {
  if (A.dbg && c->tri_two_warps) {
    CG_CUDA(cudaFuncSetAttribute(k_trisolve<FWD, ND, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes));
    k_trisolve<FWD, ND, true, 2><<<grid, 160, P.smem_bytes, KS(c, FWD ? "k_trisolve_fwd" : "k_trisolve_bwd")>>>(A);
  } else if (A.dbg) {
    CG_CUDA(cudaFuncSetAttribute(k_trisolve<FWD, ND, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes));
    k_trisolve<FWD, ND, true, 1><<<grid, 128, P.smem_bytes, KS(c, FWD ? "k_trisolve_fwd" : "k_trisolve_bwd")>>>(A);
  } else if (c->tri_two_warps) {
    CG_CUDA(cudaFuncSetAttribute(k_trisolve<FWD, ND, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes));
    k_trisolve<FWD, ND, false, 2><<<grid, 160, P.smem_bytes, KS(c, FWD ? "k_trisolve_fwd" : "k_trisolve_bwd")>>>(A);
  } else {
    CG_CUDA(cudaFuncSetAttribute(k_trisolve<FWD, ND, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes));
    k_trisolve<FWD, ND, false, 1><<<grid, 128, P.smem_bytes, KS(c, FWD ? "k_trisolve_fwd" : "k_trisolve_bwd")>>>(A);
  }
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}

template <bool FWD>
static int launch_tri(Ctx* c, TriPlan& P, double* vec, const double* src, bool skip_done) {
  TriArgs A;
  A.np = c->np; A.ns = P.ns; A.R = P.R; A.B = c->B;
  A.meta_stride = P.meta_stride; A.val_stride = P.val_stride; A.xs_rows = P.R + P.max_ghost + 2;
  A.max_steps = P.max_steps;
  A.nval = P.nval;
  A.strip_ms = P.strip_ms; A.strip_gh = P.strip_gh; A.ms_rows = P.ms_rows;
  A.gh_row = P.gh_row; A.gh_src = P.gh_src; A.gh_need = P.gh_need; A.meta = P.meta; A.packed = P.packed;
  A.progress = P.progress; A.ticket = P.ticket;
  A.pbase = P.pbase; A.tbase = P.tbase;
  A.src = src; A.vec = vec; A.cnt = skip_done ? c->cnt : nullptr;
  const int grid = P.ns * c->B;
  // exchange vector: every component starts as "not yet produced"
  if (!c->tri_xch) CG_CUDA(cudaMalloc((void**)&c->tri_xch, sizeof(unsigned long long) * (size_t)c->B * c->ndof));
  CG_CUDA(cudaMemsetAsync(c->tri_xch, 0xFF, sizeof(unsigned long long) * (size_t)c->B * c->ndof, c->stream));
  A.xch = c->tri_xch;
  A.dbg = nullptr;
  if (c->tri_debug) {
    if (!c->tri_dbg_buf) CG_CUDA(cudaMalloc((void**)&c->tri_dbg_buf, sizeof(unsigned long long) * 8 * 2 * 4096));
    if (grid <= 4096) A.dbg = c->tri_dbg_buf + (FWD ? 0 : 8 * 4096);
    c->tri_dbg_grid[FWD ? 0 : 1] = grid;
  }
  switch (std::min(std::max(P.ndmax, 1), (int)TS_MAXD)) {  // operand layers per micro-step
    case 1: CG_CHECK((launch_tri_nd<FWD, 1>(c, P, A, grid))); break;
    case 2: CG_CHECK((launch_tri_nd<FWD, 2>(c, P, A, grid))); break;
    case 3: CG_CHECK((launch_tri_nd<FWD, 3>(c, P, A, grid))); break;
    case 4: CG_CHECK((launch_tri_nd<FWD, 4>(c, P, A, grid))); break;
    case 5: CG_CHECK((launch_tri_nd<FWD, 5>(c, P, A, grid))); break;
    case 6: CG_CHECK((launch_tri_nd<FWD, 6>(c, P, A, grid))); break;
    case 7: CG_CHECK((launch_tri_nd<FWD, 7>(c, P, A, grid))); break;
    default: CG_CHECK((launch_tri_nd<FWD, 8>(c, P, A, grid))); break;
  }
  P.pbase += (unsigned long long)P.max_steps + 2;
  P.tbase += (unsigned long long)grid;
  return CLOTHGPU_OK;
}

int launch_ilu_solve_strips(Ctx* c, double* vec, const double* src, bool skip_done) {
  if (!c->ilu_packed_valid) {
    CG_CHECK(pack_plan(c, c->plan_f));
    CG_CHECK(pack_plan(c, c->plan_b));
    c->ilu_packed_valid = true;
  }
  CG_CHECK(launch_tri<true>(c, c->plan_f, vec, src, skip_done));
  CG_CHECK(launch_tri<false>(c, c->plan_b, vec, nullptr, skip_done));
  return CLOTHGPU_OK;
}

}  // namespace cg

using namespace cg;

extern "C" int clothgpu_debug_trisolve(clothgpu_ctx* c, int on, unsigned long long* out, int cap_ctas, int* nctas) {
  if (!c) { set_error("null context"); return CLOTHGPU_ERR_ARG; }
  CG_CUDA(cudaStreamSynchronize(c->stream));
  if (out && nctas && c->tri_dbg_buf) {
    for (int dir = 0; dir < 2; dir++) {
      const int n = std::min(c->tri_dbg_grid[dir], cap_ctas);
      nctas[dir] = n;
      CG_CUDA(cudaMemcpy(out + (size_t)dir * cap_ctas * 8, c->tri_dbg_buf + (size_t)dir * 8 * 4096,
                         sizeof(unsigned long long) * 8 * n, cudaMemcpyDeviceToHost));
    }
  }
  c->tri_debug = on != 0;
  return CLOTHGPU_OK;
}

extern "C" int clothgpu_selftest_triplan(int np, const int* crs_rowptr, const int* crs_colidx, int force_ns, int* stats) {
  if (np < 1 || !crs_rowptr || !crs_colidx) { set_error("bad argument"); return CLOTHGPU_ERR_ARG; }
  std::vector<int> dia(np);
  for (int i = 0; i < np; i++) {
    dia[i] = crs_rowptr[i + 1];
    for (int j = crs_rowptr[i]; j < crs_rowptr[i + 1]; j++)
      if (crs_colidx[j] > i) { dia[i] = j; break; }
  }
  for (int dir = 0; dir < 2; dir++) {
    TriPlanHost P;
    CG_CHECK(build_triplan_host(np, crs_rowptr, crs_colidx, dia.data(), dir == 0, force_ns, P));
    CG_CHECK(verify_triplan_host(np, crs_rowptr, crs_colidx, dia.data(), P));
    if (dir == 0) {  // the factorisation follows the forward sweep's schedule
      FacPlanHost F;
      CG_CHECK(build_facplan_host(np, crs_rowptr, crs_colidx, dia.data(), P, F));
      CG_CHECK(verify_facplan_host(np, crs_rowptr, crs_colidx, dia.data(), P, F));
    }
    if (stats) {
      int* o = stats + dir * 8;
      o[0] = P.ns; o[1] = P.R; o[2] = P.nms; o[3] = P.max_steps; o[4] = P.max_ghost; o[5] = (int)P.sim_makespan;
      o[6] = P.nfast; o[7] = P.max_val_bytes;
    }
  }
  return CLOTHGPU_OK;
}
