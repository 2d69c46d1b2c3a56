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
//     them (k_pack) and streamed into a 4-stage shared-memory ring by 1-D TMA bulk copies (cp.async.bulk +
//     mbarrier complete_tx) issued by a loader warp that runs ahead of the compute warp.
//   * values produced by other strips ("ghosts") are imported into shared memory by an import warp that polls the
//     producer strip's progress counter (ld.acquire.gpu) ahead of need; finished rows are written back to global
//     memory and the strip's own progress counter is advanced (fence + st.release.gpu) by a publisher warp.
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

inline int round_up(int v, int a) { return (v + a - 1) / a * a; }

struct Sched {                 // micro-step schedule of one candidate strip size
  int R = 0, ns = 0, nms = 0;
  std::vector<int> strip_ms;   // ns+1
  std::vector<int> ms_rows;    // nms*TS_ROWS (global row ids, -1 = unused)
  std::vector<int> ms_of_row;  // global micro-step index per row
};

}  // namespace

static inline int plan_pos(bool fwd, int np, int row) { return fwd ? row : np - 1 - row; }

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
  const int smem_budget = 200 * 1024;
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
    const long long smem = (long long)(R + mg) * 32 + 56 * 1024;
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
      if ((long long)(R + mg) * 32 + 56 * 1024 <= smem_budget) { best_t = t; best = std::move(S); best_ghost = mg; break; }
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

  // materialise: per micro-step meta chunk, packed-value source map, ghost lists
  P.meta.clear(); P.meta_off.assign(1, 0); P.val_off.assign(1, 0); P.pk_src.clear();
  P.strip_gh.assign(S.ns + 1, 0); P.gh_row.clear(); P.gh_src.clear(); P.gh_need.clear();
  P.max_meta_bytes = 0; P.max_val_bytes = 0; P.nfast = 0;
  std::vector<int> ghost_slot(np, -1), ghost_owner(np, -1);
  for (int k = 0; k < S.ns; k++) {
    const int p0 = k * S.R;
    int nghost = 0;
    for (int m = S.strip_ms[k]; m < S.strip_ms[k + 1]; m++) {
      const int* rows = &S.ms_rows[(size_t)m * TS_ROWS];
      int ndep[TS_ROWS], ndmax = 0, nrows = 0;
      for (int q = 0; q < TS_ROWS; q++) {
        ndep[q] = 0;
        if (rows[q] < 0) continue;
        nrows++;
        ndep[q] = fwd ? dia[rows[q]] - rowptr[rows[q]] : rowptr[rows[q] + 1] - dia[rows[q]];
        ndmax = std::max(ndmax, ndep[q]);
      }
      const int nint = round_up(TS_META_HDR + ndmax * TS_ROWS, 4);
      const size_t mbase = P.meta.size();
      P.meta.resize(mbase + nint, 0);
      int* M = &P.meta[mbase];
      const int nlayer = ndmax + (fwd ? 1 : 0);
      const size_t vbase = P.pk_src.size();
      P.pk_src.resize(vbase + (size_t)nlayer * 96, -1);
      int* V = &P.pk_src[vbase];
      int fmin = ndmax, fast = ndmax <= TS_MAXD ? 1 : 0;
      for (int q = 0; q < TS_ROWS; q++) {
        const int row = rows[q];
        M[q] = row < 0 ? -1 : plan_pos(fwd, np, row) - p0;
        M[TS_ROWS + q] = ndep[q];
        M[28 + q] = M[38 + q] = -1;
        if (row < 0) continue;
        const int j0 = fwd ? rowptr[row] : dia[row];
        int nnear = 0;
        for (int d = 0; d < ndep[q]; d++) {
          const int j = j0 + d, c = colidx[j];
          const int cp = plan_pos(fwd, np, c);
          int slot;
          if (cp / S.R == k) {
            slot = cp - p0;
            if (S.ms_of_row[c] == m - 1) {  // produced by the immediately preceding micro-step: a "near" operand
              if (nnear == 0) M[28 + q] = d;
              else if (nnear == 1) M[38 + q] = d;
              else fast = 0;
              if (nnear == 0) fmin = std::min(fmin, d);
              nnear++;
            }
          } else {
            if (ghost_owner[c] != k) {
              ghost_owner[c] = k;
              ghost_slot[c] = nghost++;
              P.gh_row.push_back(c);
              P.gh_src.push_back(cp / S.R);
              P.gh_need.push_back(S.ms_of_row[c] - S.strip_ms[cp / S.R] + 1);
            }
            slot = S.R + ghost_slot[c];
          }
          M[TS_META_HDR + d * TS_ROWS + q] = slot;
          for (int kk = 0; kk < 3; kk++)
            for (int r = 0; r < 3; r++) V[(d * 3 + kk) * 32 + q * 3 + r] = j * 9 + r * 3 + kk;
        }
        if (fwd)
          for (int kk = 0; kk < 3; kk++)
            for (int r = 0; r < 3; r++) V[(ndmax * 3 + kk) * 32 + q * 3 + r] = ncrs * 9 + row * 9 + r * 3 + kk;
      }
      M[2 * TS_ROWS + 0] = ndmax;
      M[2 * TS_ROWS + 1] = nghost;  // ghosts that must have been imported before this micro-step runs
      M[2 * TS_ROWS + 2] = nrows;
      M[2 * TS_ROWS + 3] = fast;
      M[2 * TS_ROWS + 4] = fmin;
      P.nfast += fast;
      P.meta_off.push_back((int)(P.meta.size() * sizeof(int)));
      P.val_off.push_back((int)P.pk_src.size());
      P.max_meta_bytes = std::max(P.max_meta_bytes, nint * (int)sizeof(int));
      P.max_val_bytes = std::max(P.max_val_bytes, nlayer * 96 * (int)sizeof(double));
    }
    P.strip_gh[k + 1] = (int)P.gh_row.size();
    P.max_ghost = std::max(P.max_ghost, nghost);
  }
  P.nval = (long long)P.pk_src.size();
  if (P.pk_src.size() > (size_t)0x7fffffff) { set_error("triangular-solve plan exceeds 2^31 packed values"); return CLOTHGPU_ERR_CAPACITY; }
  return CLOTHGPU_OK;
}

// Structural self-check of a plan (host logic, no GPU): every row is scheduled exactly once, every dependency is
// either an own row of an earlier micro-step or a ghost whose producer strip is earlier and whose import is
// required in time, and every packed value maps to the right matrix entry.
int verify_triplan_host(int np, const int* rowptr, const int* colidx, const int* dia, const TriPlanHost& P) {
  const bool fwd = P.fwd;
  const int ncrs = rowptr[np];
  std::vector<int> ms_of_row(np, -1);
  for (int m = 0; m < P.nms; m++)
    for (int q = 0; q < TS_ROWS; q++) {
      const int row = P.ms_rows[(size_t)m * TS_ROWS + q];
      if (row < 0) continue;
      if (row >= np || ms_of_row[row] != -1) { set_error("plan: row scheduled twice or out of range"); return CLOTHGPU_ERR_STATE; }
      ms_of_row[row] = m;
    }
  for (int i = 0; i < np; i++)
    if (ms_of_row[i] < 0) { set_error("plan: row not scheduled"); return CLOTHGPU_ERR_STATE; }
  if ((int)P.meta_off.size() != P.nms + 1 || (int)P.val_off.size() != P.nms + 1) { set_error("plan: offset tables"); return CLOTHGPU_ERR_STATE; }
  for (int k = 0; k < P.ns; k++) {
    const int p0 = k * P.R, g0 = P.strip_gh[k], ng = P.strip_gh[k + 1] - g0;
    int prev_need = 0;
    for (int m = P.strip_ms[k]; m < P.strip_ms[k + 1]; m++) {
      if (P.meta_off[m] % 16 || P.val_off[m] % 2) { set_error("plan: chunk alignment"); return CLOTHGPU_ERR_STATE; }
      const int* M = &P.meta[P.meta_off[m] / 4];
      const int* V = &P.pk_src[P.val_off[m]];
      const int ndmax = M[2 * TS_ROWS], gneed = M[2 * TS_ROWS + 1];
      if (gneed < prev_need || gneed > ng) { set_error("plan: ghost requirement not monotone"); return CLOTHGPU_ERR_STATE; }
      prev_need = gneed;
      const int nlayer = ndmax + (fwd ? 1 : 0);
      if (P.val_off[m + 1] - P.val_off[m] != nlayer * 96) { set_error("plan: value chunk size"); return CLOTHGPU_ERR_STATE; }
      if (P.meta_off[m + 1] - P.meta_off[m] != round_up(TS_META_HDR + ndmax * TS_ROWS, 4) * 4) { set_error("plan: meta chunk size"); return CLOTHGPU_ERR_STATE; }
      for (int q = 0; q < TS_ROWS; q++) {
        const int row = P.ms_rows[(size_t)m * TS_ROWS + q];
        if (row < 0) {
          if (M[q] != -1 || M[TS_ROWS + q] != 0) { set_error("plan: unused lane not blank"); return CLOTHGPU_ERR_STATE; }
          continue;
        }
        const int pos = plan_pos(fwd, np, row);
        if (pos / P.R != k || M[q] != pos - p0) { set_error("plan: own slot"); return CLOTHGPU_ERR_STATE; }
        const int j0 = fwd ? rowptr[row] : dia[row], j1 = fwd ? dia[row] : rowptr[row + 1];
        if (M[TS_ROWS + q] != j1 - j0 || j1 - j0 > ndmax) { set_error("plan: dependency count"); return CLOTHGPU_ERR_STATE; }
        const int fast = M[2 * TS_ROWS + 3], fmin = M[2 * TS_ROWS + 4];
        if (fast && ndmax > TS_MAXD) { set_error("plan: fast step with too many operands"); return CLOTHGPU_ERR_STATE; }
        int nnear = 0;
        for (int d = 0; d < j1 - j0; d++) {
          const int c = colidx[j0 + d], cp = plan_pos(fwd, np, c);
          const bool is_near = cp / P.R == k && ms_of_row[c] == m - 1;
          if (!is_near) continue;
          nnear++;
          if (fast && (d < fmin || (M[28 + q] != d && M[38 + q] != d))) { set_error("plan: near operand not marked"); return CLOTHGPU_ERR_STATE; }
        }
        if (fast && nnear > 2) { set_error("plan: fast step with more than two near operands"); return CLOTHGPU_ERR_STATE; }
        if (fast && (fmin < 0 || fmin > ndmax)) { set_error("plan: chain split"); return CLOTHGPU_ERR_STATE; }
        for (int d = 0; d < j1 - j0; d++) {
          const int c = colidx[j0 + d], cp = plan_pos(fwd, np, c), slot = M[TS_META_HDR + d * TS_ROWS + q];
          if (cp / P.R == k) {
            if (slot != cp - p0 || ms_of_row[c] >= m) { set_error("plan: own dependency not earlier"); return CLOTHGPU_ERR_STATE; }
          } else {
            const int gi = slot - P.R;
            if (cp / P.R > k || gi < 0 || gi >= gneed || P.gh_row[g0 + gi] != c || P.gh_src[g0 + gi] != cp / P.R ||
                P.gh_need[g0 + gi] != ms_of_row[c] - P.strip_ms[cp / P.R] + 1) {
              set_error("plan: ghost dependency");
              return CLOTHGPU_ERR_STATE;
            }
          }
          for (int kk = 0; kk < 3; kk++)
            for (int r = 0; r < 3; r++)
              if (V[(d * 3 + kk) * 32 + q * 3 + r] != (j0 + d) * 9 + r * 3 + kk) { set_error("plan: packed value map"); return CLOTHGPU_ERR_STATE; }
        }
        if (fwd)
          for (int kk = 0; kk < 3; kk++)
            for (int r = 0; r < 3; r++)
              if (V[(ndmax * 3 + kk) * 32 + q * 3 + r] != ncrs * 9 + row * 9 + r * 3 + kk) { set_error("plan: packed diagonal map"); return CLOTHGPU_ERR_STATE; }
      }
    }
  }
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
  D.sim_makespan = H.sim_makespan;
  CG_CHECK(up(&D.strip_ms, H.strip_ms)); CG_CHECK(up(&D.strip_gh, H.strip_gh)); CG_CHECK(up(&D.ms_rows, H.ms_rows));
  CG_CHECK(up(&D.meta_off, H.meta_off)); CG_CHECK(up(&D.val_off, H.val_off)); CG_CHECK(up(&D.meta, H.meta));
  CG_CHECK(up(&D.pk_src, H.pk_src)); CG_CHECK(up(&D.gh_row, H.gh_row)); CG_CHECK(up(&D.gh_src, H.gh_src));
  CG_CHECK(up(&D.gh_need, H.gh_need));
  CG_CUDA(cudaMalloc((void**)&D.packed, sizeof(double) * (size_t)std::max<long long>(H.nval, 1) * c->B));
  CG_CUDA(cudaMalloc((void**)&D.progress, sizeof(unsigned long long) * (size_t)H.ns * c->B));
  CG_CUDA(cudaMemset(D.progress, 0, sizeof(unsigned long long) * (size_t)H.ns * c->B));
  CG_CUDA(cudaMalloc((void**)&D.ticket, sizeof(unsigned long long)));
  CG_CUDA(cudaMemset(D.ticket, 0, sizeof(unsigned long long)));
  D.pbase = 0;
  D.tbase = 0;
  const int stage_meta = round_up(D.max_meta_bytes, 128), stage_val = round_up(std::max(D.max_val_bytes, 16), 128);
  D.smem_bytes = TS_SMEM_HDR + TS_NST * (stage_meta + stage_val) + (D.R + D.max_ghost) * 32 + 64;
  return CLOTHGPU_OK;
}

void free_triplan(TriPlan& D) {
  void* ptrs[] = {D.strip_ms, D.strip_gh, D.ms_rows, D.meta_off, D.val_off, D.meta, D.pk_src, D.gh_row, D.gh_src,
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

struct TriArgs {
  int np, ns, R, B;
  int stage_meta, stage_bytes;
  long long nval;
  const int *strip_ms, *strip_gh, *ms_rows, *meta_off, *val_off, *gh_row, *gh_src, *gh_need;
  const int* meta;
  const double* packed;
  unsigned long long *progress, *ticket;
  unsigned long long pbase, tbase;
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
template <bool FWD>
__global__ void __launch_bounds__(128) k_trisolve(const TriArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + TS_NST;
  uint32_t* flags = reinterpret_cast<uint32_t*>(smem + 16 * TS_NST);  // [0] micro-steps done, [1] ghosts imported, [2] ticket
  unsigned char* stages = smem + TS_SMEM_HDR;
  double* xs = reinterpret_cast<double*>(stages + TS_NST * A.stage_bytes);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    const unsigned long long t = atomicAdd(A.ticket, 1ULL) - A.tbase;
    flags[2] = (uint32_t)t;
  }
  __syncthreads();
  const uint32_t ticket = flags[2];
  const int strip = (int)(ticket / (uint32_t)A.B), scene = (int)(ticket % (uint32_t)A.B);
  if (A.cnt && A.cnt[scene * NCNT + CNT_PCG_DONE]) return;

  const int ms0 = A.strip_ms[strip], nsteps = A.strip_ms[strip + 1] - ms0;
  const int p0 = strip * A.R, nown = min(A.R, A.np - p0);
  double* out = A.vec + (size_t)scene * A.np * 3;
  const double* in = (FWD && A.src) ? A.src + (size_t)scene * A.np * 3 : out;

  // the strip's slice of the vector: right-hand side now, solution as the sweep proceeds
  for (int idx = tid; idx < nown * 3; idx += blockDim.x) {
    const int p = p0 + idx / 3, comp = idx - (idx / 3) * 3;
    const int row = FWD ? p : A.np - 1 - p;
    xs[(idx / 3) * 4 + comp] = in[(size_t)row * 3 + comp];
  }
  if (tid == 0) {
    for (int i = 0; i < TS_NST; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    flags[0] = 0;
    flags[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    // ---------------------------------------------------------------- compute warp
    // Software pipeline over the micro-steps.  Everything of step s+1 that does not depend on step s ("phase A":
    // matrix values, far operands, their products, the prefix of the subtraction chain) is issued in the shadow of
    // the dependent chain of step s ("phase B": operands produced by step s-1..s, the rest of the chain, the D^-1
    // multiply).  The plan marks per row which (at most two) operands are "near" = produced by the immediately
    // preceding micro-step of this strip; steps that do not fit (more than TS_MAXD operands or more than two near
    // ones in a row) take the generic path.
    const int g = lane / 3, r = lane - g * 3;
    const bool act = g < TS_ROWS;
    const int gq = act ? g : TS_ROWS - 1;
    const int l0 = (gq * 3) & 31;
    uint32_t seen = 0;
    const bool dbg = A.dbg != nullptr;
    unsigned long long t_begin = 0, w_mbar = 0, w_ghost = 0, w_comp = 0, w_sig = 0;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (dbg) t_begin = gtime_ns();

    // state of the step in flight (filled by phase A)
    int fast = 0, myslot = -1, ndmax = 0, fmin = 0, n0 = -1, n1 = -1, ns0 = 0, ns1 = 0;
    double tp = 0.0, p[TS_MAXD], an0[3], an1[3], dd[3];
#pragma unroll
    for (int d = 0; d < TS_MAXD; d++) p[d] = 0.0;
    an0[0] = an0[1] = an0[2] = an1[0] = an1[1] = an1[2] = dd[0] = dd[1] = dd[2] = 0.0;

    // waits of a step: its matrix chunk and its imported operands
    auto wait_step = [&](int s) {
      if (dbg) c0 = clock64();
      mbar_wait(&full[s % TS_NST], (uint32_t)(s / TS_NST) & 1u);
      if (dbg) c1 = clock64();
      const int* M = reinterpret_cast<const int*>(stages + (size_t)(s % TS_NST) * A.stage_bytes);
      const uint32_t gneed = (uint32_t)M[2 * TS_ROWS + 1];
      if (gneed > seen) {
        do { seen = ld_acquire_cta(&flags[1]); } while (seen < gneed);
      }
      if (dbg) { c2 = clock64(); w_mbar += (unsigned long long)(c1 - c0); w_ghost += (unsigned long long)(c2 - c1); }
    };
    // phase A of a fast step: loads and products of the far operands, chain prefix
#define TS_PHASE_A(S_)                                                                                          \
    {                                                                                                           \
      const int* M = reinterpret_cast<const int*>(stages + (size_t)((S_) % TS_NST) * A.stage_bytes);           \
      const double* V = reinterpret_cast<const double*>(stages + (size_t)((S_) % TS_NST) * A.stage_bytes + A.stage_meta); \
      fast = M[2 * TS_ROWS + 3];                                                                                \
      if (fast) {                                                                                               \
        myslot = act ? M[gq] : -1;                                                                              \
        const int nd = act ? M[TS_ROWS + gq] : 0;                                                               \
        ndmax = M[2 * TS_ROWS];                                                                                 \
        fmin = M[2 * TS_ROWS + 4];                                                                              \
        n0 = act ? M[28 + gq] : -1;                                                                             \
        n1 = act ? M[38 + gq] : -1;                                                                             \
        tp = myslot >= 0 ? xs[myslot * 4 + r] : 0.0;                                                            \
        _Pragma("unroll") for (int d = 0; d < TS_MAXD; d++) {                                                   \
          if (d < ndmax) {                                                                                      \
            const int sl = M[TS_META_HDR + d * TS_ROWS + gq];                                                   \
            const double a0 = V[(d * 3 + 0) * 32 + lane], a1 = V[(d * 3 + 1) * 32 + lane], a2 = V[(d * 3 + 2) * 32 + lane]; \
            const double2 x01 = *reinterpret_cast<const double2*>(xs + sl * 4);                                 \
            const double x2 = xs[sl * 4 + 2];                                                                   \
            const double pd = a0 * x01.x + a1 * x01.y + a2 * x2;                                                \
            const bool is_near = (d == n0) || (d == n1);                                                           \
            p[d] = (d < nd && !is_near) ? pd : 0.0;                                                                \
            if (d == n0) { an0[0] = a0; an0[1] = a1; an0[2] = a2; ns0 = sl; }                                   \
            if (d == n1) { an1[0] = a0; an1[1] = a1; an1[2] = a2; ns1 = sl; }                                   \
          }                                                                                                     \
        }                                                                                                       \
        _Pragma("unroll") for (int d = 0; d < TS_MAXD; d++) {                                                   \
          if (d < fmin) tp -= p[d];                                                                             \
        }                                                                                                       \
        if (FWD) {                                                                                              \
          dd[0] = V[(ndmax * 3 + 0) * 32 + lane]; dd[1] = V[(ndmax * 3 + 1) * 32 + lane]; dd[2] = V[(ndmax * 3 + 2) * 32 + lane]; \
        }                                                                                                       \
      }                                                                                                         \
    }

    if (nsteps > 0) {
      wait_step(0);
      TS_PHASE_A(0)
    }
    for (int s = 0; s < nsteps; s++) {
      const int st = s % TS_NST;
      const bool more = s + 1 < nsteps;
      if (more) wait_step(s + 1);
      if (dbg) c2 = clock64();
      double t;
      int slot_out;
      if (fast) {
        // ---- phase B of step s: near operands, rest of the chain
        const int b_slot = myslot, b_n0 = n0, b_n1 = n1, b_fmin = fmin, b_ndmax = ndmax;
        const double2 y01 = *reinterpret_cast<const double2*>(xs + ns0 * 4);
        const double y2 = xs[ns0 * 4 + 2];
        const double2 z01 = *reinterpret_cast<const double2*>(xs + ns1 * 4);
        const double z2 = xs[ns1 * 4 + 2];
        const double pn0 = an0[0] * y01.x + an0[1] * y01.y + an0[2] * y2;
        const double pn1 = an1[0] * z01.x + an1[1] * z01.y + an1[2] * z2;
        double q[TS_MAXD];
#pragma unroll
        for (int d = 0; d < TS_MAXD; d++) q[d] = (d == b_n0) ? pn0 : ((d == b_n1) ? pn1 : p[d]);
        t = tp;
        const double b_dd0 = dd[0], b_dd1 = dd[1], b_dd2 = dd[2];
        // ---- phase A of step s+1 (independent of the chain below: scheduled into its stalls)
        if (more) TS_PHASE_A(s + 1)
#pragma unroll
        for (int d = 0; d < TS_MAXD; d++) {
          if (d >= b_fmin && d < b_ndmax) t -= q[d];
        }
        if (FWD) {  // v_i = D_i^-1 t
          const double t0 = __shfl_sync(0xffffffffu, t, l0);
          const double t1 = __shfl_sync(0xffffffffu, t, (l0 + 1) & 31);
          const double t2 = __shfl_sync(0xffffffffu, t, (l0 + 2) & 31);
          t = b_dd0 * t0 + b_dd1 * t1 + b_dd2 * t2;
        }
        slot_out = b_slot;
      } else {
        // ---- generic path: everything after the previous step has completed
        const int* M = reinterpret_cast<const int*>(stages + (size_t)st * A.stage_bytes);
        const double* V = reinterpret_cast<const double*>(stages + (size_t)st * A.stage_bytes + A.stage_meta);
        slot_out = act ? M[gq] : -1;
        const int nd = act ? M[TS_ROWS + gq] : 0;
        const int ndm = M[2 * TS_ROWS];
        t = slot_out >= 0 ? xs[slot_out * 4 + r] : 0.0;
        for (int d = 0; d < ndm; d++) {
          const int sl = M[TS_META_HDR + d * TS_ROWS + gq];
          const double a0 = V[(d * 3 + 0) * 32 + lane], a1 = V[(d * 3 + 1) * 32 + lane], a2 = V[(d * 3 + 2) * 32 + lane];
          const double x0 = xs[sl * 4 + 0], x1 = xs[sl * 4 + 1], x2 = xs[sl * 4 + 2];
          const double pd = a0 * x0 + a1 * x1 + a2 * x2;
          t -= (d < nd) ? pd : 0.0;
        }
        if (FWD) {
          const double t0 = __shfl_sync(0xffffffffu, t, l0);
          const double t1 = __shfl_sync(0xffffffffu, t, (l0 + 1) & 31);
          const double t2 = __shfl_sync(0xffffffffu, t, (l0 + 2) & 31);
          const double d0 = V[(ndm * 3 + 0) * 32 + lane], d1 = V[(ndm * 3 + 1) * 32 + lane], d2 = V[(ndm * 3 + 2) * 32 + lane];
          t = d0 * t0 + d1 * t1 + d2 * t2;
        }
        if (more) TS_PHASE_A(s + 1)
      }
      if (slot_out >= 0) xs[slot_out * 4 + r] = t;
      __syncwarp();
      if (dbg) c3 = clock64();
      if (lane == 0) {
        st_release_cta(&flags[0], (uint32_t)(s + 1));
        mbar_arrive(&empty[st]);
      }
      if (dbg) {
        __syncwarp();
        w_comp += (unsigned long long)(c3 - c2);
        w_sig += (unsigned long long)(clock64() - c3);
      }
    }
    if (dbg && lane == 0) {
      unsigned long long* o = A.dbg + (size_t)blockIdx.x * 8;
      o[0] = t_begin; o[1] = gtime_ns(); o[2] = w_mbar; o[3] = w_ghost; o[4] = w_comp; o[5] = w_sig;
      o[6] = (unsigned long long)nsteps; o[7] = (unsigned long long)strip;
    }
#undef TS_PHASE_A
  } else if (warp == 1) {
    // ---------------------------------------------------------------- loader warp: TMA bulk copies, NST ahead
    const unsigned char* meta = reinterpret_cast<const unsigned char*>(A.meta);
    const double* packed = A.packed + (size_t)scene * A.nval;
    for (int s0 = 0; s0 < nsteps; s0 += 32) {
      const int n = min(32, nsteps - s0);
      const int m = ms0 + s0 + min(lane, n - 1);
      const int mo = A.meta_off[m], mo1 = A.meta_off[m + 1], vo = A.val_off[m], vo1 = A.val_off[m + 1];
      for (int j = 0; j < n; j++) {
        const int mo_j = __shfl_sync(0xffffffffu, mo, j), mb = __shfl_sync(0xffffffffu, mo1, j) - mo_j;
        const int vo_j = __shfl_sync(0xffffffffu, vo, j), vb = (__shfl_sync(0xffffffffu, vo1, j) - vo_j) * 8;
        if (lane == 0) {
          const int s = s0 + j, st = s % TS_NST;
          if (s >= TS_NST) mbar_wait(&empty[st], (uint32_t)(s / TS_NST - 1) & 1u);
          unsigned char* dst = stages + (size_t)st * A.stage_bytes;
          mbar_expect_tx(&full[st], (uint32_t)(mb + vb));
          bulk_g2s(dst, meta + mo_j, (uint32_t)mb, &full[st]);
          if (vb > 0) bulk_g2s(dst + A.stage_meta, packed + vo_j, (uint32_t)vb, &full[st]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ---------------------------------------------------------------- publisher warp
    unsigned long long* prog = A.progress + (size_t)scene * A.ns + strip;
    uint32_t pub = 0;
    while (pub < (uint32_t)nsteps) {
      uint32_t d;
      do { d = ld_acquire_cta(&flags[0]); } while (d <= pub);
      const int items = (int)(d - pub) * 30;
#pragma unroll 4
      for (int it = lane; it < items; it += 32) {
        const int s = (int)pub + it / 30, q = it - (it / 30) * 30;
        const int row = A.ms_rows[(size_t)(ms0 + s) * TS_ROWS + q / 3];
        if (row >= 0) {
          const int p = FWD ? row : A.np - 1 - row;
          out[(size_t)row * 3 + (q % 3)] = xs[(p - p0) * 4 + (q % 3)];
        }
      }
      __syncwarp();
      if (lane == 0) {
        __threadfence();
        st_release_gpu(prog, A.pbase + d);
      }
      pub = d;
    }
  } else {
    // ---------------------------------------------------------------- import warp
    const int g0 = A.strip_gh[strip], ngh = A.strip_gh[strip + 1] - g0;
    const int gl = lane / 3, comp = lane - gl * 3;
    const unsigned long long* prog = A.progress + (size_t)scene * A.ns;
    const bool on = gl < TS_ROWS;
    int row = 0, srcs = 0, need = 0;
    if (on && gl < ngh) { row = A.gh_row[g0 + gl]; srcs = A.gh_src[g0 + gl]; need = A.gh_need[g0 + gl]; }
    int cached_src = -1;
    unsigned long long cached = 0;
    for (int b = 0; b < ngh; b += TS_ROWS) {
      const int gi = b + gl;
      const bool mine = on && gi < ngh;
      const int crow = row, csrc = srcs, cneed = need;
      // prefetch the next batch's description before waiting on this one
      const int gn = gi + TS_ROWS;
      if (on && gn < ngh) { row = A.gh_row[g0 + gn]; srcs = A.gh_src[g0 + gn]; need = A.gh_need[g0 + gn]; }
      if (mine) {
        const unsigned long long want = A.pbase + (unsigned long long)cneed;
        if (cached_src != csrc) { cached_src = csrc; cached = 0; }
        while (cached < want) cached = ld_acquire_gpu(prog + csrc);
        xs[(A.R + gi) * 4 + comp] = __ldcg(out + (size_t)crow * 3 + comp);
      }
      __syncwarp();
      if (lane == 0) st_release_cta(&flags[1], (uint32_t)min(b + TS_ROWS, ngh));
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

template <bool FWD>
static int launch_tri(Ctx* c, TriPlan& P, double* vec, const double* src, bool skip_done) {
  TriArgs A;
  A.np = c->np; A.ns = P.ns; A.R = P.R; A.B = c->B;
  A.stage_meta = round_up(P.max_meta_bytes, 128);
  A.stage_bytes = A.stage_meta + round_up(std::max(P.max_val_bytes, 16), 128);
  A.nval = P.nval;
  A.strip_ms = P.strip_ms; A.strip_gh = P.strip_gh; A.ms_rows = P.ms_rows; A.meta_off = P.meta_off; A.val_off = P.val_off;
  A.gh_row = P.gh_row; A.gh_src = P.gh_src; A.gh_need = P.gh_need; A.meta = P.meta; A.packed = P.packed;
  A.progress = P.progress; A.ticket = P.ticket;
  A.pbase = P.pbase; A.tbase = P.tbase;
  A.src = src; A.vec = vec; A.cnt = skip_done ? c->cnt : nullptr;
  const int grid = P.ns * c->B;
  A.dbg = nullptr;
  if (c->tri_debug) {
    if (!c->tri_dbg_buf) CG_CUDA(cudaMalloc((void**)&c->tri_dbg_buf, sizeof(unsigned long long) * 8 * 2 * 4096));
    if (grid <= 4096) A.dbg = c->tri_dbg_buf + (FWD ? 0 : 8 * 4096);
    c->tri_dbg_grid[FWD ? 0 : 1] = grid;
  }
  CG_CUDA(cudaFuncSetAttribute(k_trisolve<FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes));
  k_trisolve<FWD><<<grid, 128, P.smem_bytes, KS(c, FWD ? "k_trisolve_fwd" : "k_trisolve_bwd")>>>(A);
  CG_LAUNCH_CHECK(c);
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
    if (stats) {
      int* o = stats + dir * 8;
      o[0] = P.ns; o[1] = P.R; o[2] = P.nms; o[3] = P.max_steps; o[4] = P.max_ghost; o[5] = (int)P.sim_makespan;
      o[6] = P.nfast; o[7] = P.max_val_bytes;
    }
  }
  return CLOTHGPU_OK;
}
