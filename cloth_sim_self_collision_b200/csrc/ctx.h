// csrc/ctx.h -- internal context of libclothgpu (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/clothgpu.h"

namespace cg {

// frontier entry of the dual-tree traversal: a node pair plus its position in the reference's depth-first order
// (pre = preorder rank of the self() node whose children started this pair recursion, path = child choices so
// far, 2 bits per descent, left aligned) -- see bvh.cu
struct PairEnt {
  int i0, i1;
  uint32_t pre;
  uint32_t depth;
  uint64_t path_hi;  // first 32 descents
  uint64_t path_lo;  // next 32 descents
};
static_assert(sizeof(PairEnt) == 32, "PairEnt layout");

// raw contact element emitted by the narrow phase: canonical key + DFS rank (pre, path, slot)
struct RawElem {
  int n[4];
  uint32_t pre;
  uint32_t slot;  // 0..5 face-vertex tests, 6..14 edge-edge tests, in the reference's order
  uint64_t path_hi;
  uint64_t path_lo;
};
static_assert(sizeof(RawElem) == 40, "RawElem layout");

enum { NSCAL = 16, NCNT = 96 };
// strip-pipelined triangular solves / factorisation (trisolve.cu, ilufactor.cu)
enum { TS_ROWS = 10, TS_NST = 3, TS_K = 4, TS_META_HDR = 16, TS_MAXD = 8, TS_SMEM_HDR = 256 };

// host-side plan of one sweep direction (pure host data: built and verified without a GPU)
struct TriPlanHost {
  bool fwd = true;
  int np = 0, ns = 0, R = 0, nms = 0, max_meta_bytes = 0, max_val_bytes = 0, max_ghost = 0, max_steps = 0;
  int nfast = 0;  // micro-steps whose operands all fit the software pipeline (<= TS_MAXD)
  int ndmax = 0, meta_stride = 0, val_stride = 0;  // largest operand count; ints / doubles per micro-step
  long long nval = 0;
  double sim_makespan = 0;
  std::vector<int> strip_ms, strip_gh, ms_rows, pk_src, gh_row, gh_src, gh_need, meta;
};
// strip-pipelined factorisation (ilufactor.cu): follows the forward sweep's strips and micro-steps
#ifndef FC_SPLIT
#define FC_SPLIT 1  // compute warps per row triplet (2: every bundle of independent updates is split between two warps)
#endif
enum { FC_H = 8, FC_NST = 3, FC_NIB = 3, FC_META_HDR = 48, FC_SMEM_HDR = 256, FC_BUNDLE = 6,
       FC_RWARPS = 4,              // row triplets per micro-step (10 rows -> 4 triplets)
       FC_CWARPS = FC_RWARPS * FC_SPLIT,  // compute warps
       FC_HB = FC_BUNDLE / FC_SPLIT,      // updates of a bundle one warp issues
       FC_NDI = 4,                 // micro-steps of D^-1 hand-off slots (compute -> drain warps)
       FC_DWARPS = 3,              // drain warps: one thread per (row, block entry) copies finished blocks to global
       FC_THREADS = (FC_CWARPS + 2 + FC_DWARPS) * 32 };  // compute, loader, import, drain
struct FacPlanHost {
  int ns = 0, R = 0, nms = 0, H = 0, gmaxB = 0, gmaxU = 0, gmaxOps = 0, max_imp = 0, meta_stride = 0, row_stride = 0;
  int max_steps = 0, gmaxLo = 0, gT = 0, kg_stride = 0, imp_bytes = 0;
  std::vector<int> strip_ms, ms_rows, meta, pk_src, imp_ptr, imp_row, imp_src, imp_need;
};
struct FacPlan {
  int ns = 0, R = 0, nms = 0, gmaxB = 0, gmaxU = 0, gmaxOps = 0, max_imp = 0, meta_stride = 0, row_stride = 0;
  int max_steps = 0, smem_bytes = 0, kg_stride = 0, gT = 0;
  long long nval = 0;
  int *strip_ms = nullptr, *meta = nullptr, *pk_src = nullptr, *imp_ptr = nullptr, *imp_row = nullptr, *imp_src = nullptr,
      *imp_need = nullptr;
  double *packed = nullptr, *uexp = nullptr;  // B x nval packed row blocks; B x np x gmaxU x 12 exported upper parts
  unsigned long long *progress = nullptr, *ticket = nullptr;
  unsigned long long pbase = 0, tbase = 0;
};
// its device image plus the per-context pipeline state
struct TriPlan {
  bool fwd = true;
  int ns = 0, R = 0, nms = 0, max_meta_bytes = 0, max_val_bytes = 0, max_ghost = 0, max_steps = 0, smem_bytes = 0;
  int meta_stride = 0, val_stride = 0, ndmax = 0;
  long long nval = 0;
  double sim_makespan = 0;
  int *strip_ms = nullptr, *strip_gh = nullptr, *ms_rows = nullptr;
  int *meta = nullptr, *pk_src = nullptr, *gh_row = nullptr, *gh_src = nullptr, *gh_need = nullptr;
  double* packed = nullptr;                // B x nval: factor blocks in lane order
  unsigned long long* progress = nullptr;  // B x ns: micro-steps published per strip (monotone across launches)
  unsigned long long* ticket = nullptr;    // strip hand-out counter (monotone across launches)
  unsigned long long pbase = 0, tbase = 0;
};
// per-scene device counters (ints): layout of cnt[]
enum {
  CNT_LEAF = 0,      // number of leaf pairs
  CNT_RAW = 1,       // number of raw contact elements
  CNT_ELEM = 2,      // number of unique contact elements
  CNT_OVERFLOW = 3,  // bit 0: frontier, bit 1: leaf pairs, bit 2: raw elements, bit 3: elems
  CNT_ACTIVE = 4,    // scene takes part in the current contact pass
  CNT_NSING = 5,     // singular pivots seen by the ILU factorisation
  CNT_PCG_DONE = 6,
  CNT_PCG_ITER = 7,
  CNT_RIZ_NODES = 8,
  CNT_NLEVEL = 9,    // impulse dependency levels
  CNT_TICKET = 10,   // arrival counter of the "last block reduces" pattern
  CNT_CAND = 11,     // narrow-phase candidates (leaf pair, test slot) that survived the cheap culls
  CNT_RAWPREV = 12,  // raw contact elements of the previous detection pass (incremental CCD passes)
  CNT_GBAR = 13,     // arrival counter of the level barrier of the fused traversal kernel
  CNT_NDIRTY = 15,   // dirty leaves of the current incremental pass (entries of dirty_list)
  CNT_IMP_SLOW = 14, // impulse elements whose fast arithmetic could not be certified (cumulative, diagnostic)
  CNT_FRONT = 16,    // frontier sizes per traversal depth: CNT_FRONT + depth  (depth < 72)
  CNT_RANK_BAD = 90  // CLOTHGPU_RANK_CHECK: [0] narrow-phase hits whose rebuilt DFS rank differs from the carried one, [1] hits checked (cumulative)
};
// per-scene device scalars (doubles): layout of scal[]
enum { SC_INV_SQ0 = 0, SC_RPR = 1, SC_PAP = 2, SC_SQ = 3, SC_ALPHA = 4, SC_BETA = 5, SC_CONV = 6, SC_TOL2 = 7,
       SC_DOT = 8 };

struct Ctx {
  int device = 0;
  int B = 1, np = 0, ntri = 0, nquad = 0, ncrs = 0, nnode = 0, ndof = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;  // proximity detection of a step, concurrent with the implicit solve
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp0 = nullptr, evp1 = nullptr;
  long long launches = 0;
  bool profiling = false;
  double phase_ms[CLOTHGPU_NPHASE] = {0};
  // per-kernel timing (clothgpu_set_kernel_timing): every launch is bracketed by events and synchronised
  bool ktiming = false;
  const char* kt_name = nullptr;
  cudaEvent_t kt0 = nullptr, kt1 = nullptr;
  std::map<std::string, std::pair<double, long long>> ktimes;  // name -> (total ms, launches)

  // ---- shared topology
  int *tri = nullptr, *quad = nullptr, *bc = nullptr, *rowptr = nullptr, *colidx = nullptr, *diaind = nullptr;
  double* xyz0 = nullptr;
  // assembly gather lists (CSR): per vertex the incident (triangle, corner) / (hinge, corner) entries in
  // element order; per off-diagonal slot the (triangle, block) / (hinge, block) entries
  int *vt_ptr = nullptr, *vt_ent = nullptr, *vq_ptr = nullptr, *vq_ent = nullptr;
  int *st_ptr = nullptr, *st_ent = nullptr, *sq_ptr = nullptr, *sq_ent = nullptr;
  int* slot_row = nullptr;  // row index of every off-diagonal slot
  // hinge constants (rest shape only): K[4] per hinge and stiff/((A0+A1) L0^2)
  double *quadK = nullptr, *quadT = nullptr;
  double quad_stiff = -1.0;  // stiff_bend the cached quadT was computed for
  // ILU level schedules
  int nlev_f = 0, nlev_b = 0, maxrows_f = 0, maxrows_b = 0;
  int *lev_f_ptr = nullptr, *lev_f_rows = nullptr, *lev_b_ptr = nullptr, *lev_b_rows = nullptr;
  // strip-pipelined triangular solves
  TriPlan plan_f, plan_b;
  FacPlan plan_fac;
  unsigned long long* tri_xch = nullptr;  // exchange vector of the triangular sweeps (trisolve.cu)
  bool tri_two_warps = false;  // two alternating compute warps per strip (CLOTHGPU_TRI_WARPS=2; measured: no gain)
  int pcg_expect = 0;  // iterations the previous PCG solve needed (+1): when to start polling the convergence flag
  bool ilu_packed_valid = false;  // plan_*.packed mirrors iluDia / iluCrs
  bool tri_debug = false;         // clothgpu_debug_trisolve: per-CTA wait / compute cycle counters
  unsigned long long* tri_dbg_buf = nullptr;
  int tri_dbg_grid[2] = {0, 0};
  bool use_strips_fac = true;     // false: the factor plan did not fit shared memory (irregular patterns)
  bool use_strips = true;         // CLOTHGPU_TRISOLVE=levels selects the level-scheduled single-CTA kernels
  // ILU update lists: per row the (ik, kj, ij) slot triples in the reference's order (ij < 0: diagonal block)
  int *op_ptr = nullptr;
  int4* ops = nullptr;
  long long nops = 0;
  // BVH
  int* nodes = nullptr;     // nnode x 3
  int* node_tri = nullptr;  // nnode x 4: leaf -> {v0, v1, v2, triangle}
  uint32_t* pre = nullptr;  // preorder rank per node
  int bvh_height = 0, ninternal = 0;
  int* internal_list = nullptr;  // ids of the internal nodes
  int *edge_index = nullptr, *edge_array = nullptr;

  // ---- per-scene state (scene-major)
  double *xyz = nullptr, *uvw = nullptr, *xyz_prev = nullptr, *uvw0 = nullptr;
  double *valDia = nullptr, *valCrs = nullptr, *iluDia = nullptr, *iluCrs = nullptr;
  double *vec_r = nullptr, *vec_x = nullptr, *vec_p = nullptr, *vec_z = nullptr;
  double *tri_dW = nullptr, *tri_ddW = nullptr, *quad_dW = nullptr;
  double* col_table = nullptr;
  double* normals = nullptr;  // vertex normals of clothgpu_make_normal (allocated on first use)
  double* gravity = nullptr;  // B x 3
  std::vector<double> gravity_override;  // B x 3 on the host, NaN = no override
  double* boxes = nullptr;
  int* refit_flag = nullptr;
  // incremental CCD detection: velocities the last CCD boxes were built from, changed vertices / BVH nodes
  double* uvw_snap = nullptr;
  int *vdirty = nullptr, *ndirty = nullptr;
  int* dirty_list = nullptr;            // incremental refit: the dirty leaf nodes (CNT_NDIRTY entries per scene)
  int* refit_nchild = nullptr;          // incremental refit: number of dirty children per node
  unsigned long long* npath = nullptr;  // per node: child choices from the root, left aligned (bit 63 - d = choice at depth d)
  int* ndepth = nullptr;                // per node: depth
  bool inc_refit = true;                // CLOTHGPU_INC_REFIT=0: incremental passes refit the whole tree
  bool slim_frontier = true;            // CLOTHGPU_SLIM_FRONTIER=0: frontier / leaf entries always carry the depth-first rank (32 bytes)
  bool leaf_slim = false;               // layout of the leaf list of the last traversal (8-byte pairs or 32-byte entries)
  bool riz_fused = true;                // CLOTHGPU_RIZ_FUSED=0: impact zones applied scene by scene through a library sort of all vertices
  bool rank_check = false;              // CLOTHGPU_RANK_CHECK=1: k_narrow cross-checks pair_rank against the ranks the frontier traversal carries
  int query_cap = 256;                  // CLOTHGPU_QUERY_CAP: per-leaf frontier capacity of k_query (tests force the overflow path)
  int canon_limit = -1;                 // raw-list length up to which the last pass used k_canon_small
  bool canon_small = true;              // CLOTHGPU_CANON_SMALL=0: every raw list goes through the library sort
  bool query_trav = true;               // CLOTHGPU_QUERY_TRAVERSAL=0: incremental passes use the frontier traversal
  bool incremental_ccd = true;  // CLOTHGPU_INCREMENTAL=0 re-runs every pass in full
  int imp_threads = 256;        // CLOTHGPU_IMP_THREADS: threads of the data-flow impulse kernel
  int imp_sleep_ns = 100;       // CLOTHGPU_IMP_SLEEP: nanoseconds between two polls of a predecessor's flag
  int imp_lanes = 32;           // CLOTHGPU_IMP_LANES: working lanes per warp of the data-flow impulse kernel
  bool imp_fast = true;         // CLOTHGPU_IMP_FAST=0: impulses with the plain IEEE divisions / square roots only (imp_fast.cuh)
  int trav_coop = 0;            // max scenes for which the fused (cooperative) traversal is co-resident; 0 = off
  int trav_sms = 148;
  double *scal = nullptr, *partials = nullptr;
  // exact (serial-order) dot products, linsolve.cu k_xdot_*: chunk sums and per-chunk integer statistics
  bool exact_dot = true;  // CLOTHGPU_DOT=tree selects the tree reduction (deterministic, not the reference's order)
  int xd_nch = 0;
  bool xd_attr_set = false;
  double* xd_tsum = nullptr;
  void* xd_stat = nullptr;
  int* cnt = nullptr;
  int* zone = nullptr;    // union-find parent per vertex, -1 = not a member
  int* zone_label = nullptr;
  // traversal queues (capacities are per scene)
  long long cap_front = 0, cap_leaf = 0, cap_raw = 0, cap_elem = 0;
  PairEnt *frontA = nullptr, *frontB = nullptr, *leaf = nullptr;
  int4* cand = nullptr;  // 2 x cap_leaf candidates per scene: {leaf node i0, leaf node i1, test slot, leaf pair index}
  RawElem *raw = nullptr, *raw_sorted = nullptr;
  int* elems = nullptr;  // cap_elem x 5
  // impulse scheduling scratch
  int *imp_level = nullptr, *imp_order = nullptr, *imp_lptr = nullptr, *imp_pred = nullptr;
  // generic scratch for library sorts/scans
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  unsigned long long *sort_keys = nullptr, *sort_keys2 = nullptr;
  int *sort_vals = nullptr, *sort_vals2 = nullptr;
  long long cap_sort = 0;

  // pinned host mirrors
  int* h_cnt = nullptr;      // B x NCNT
  double* h_scal = nullptr;  // B x NSCAL
  int npartial = 0;          // blocks used by the reduction kernels
};

void set_error(const std::string& s);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define CG_CUDA(call)                                                         \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) return cg::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)
#define CG_CHECK(call)              \
  do {                              \
    int r__ = (call);               \
    if (r__ != CLOTHGPU_OK) return r__; \
  } while (0)
// stream argument of a kernel launch; in kernel-timing mode it also records the start event
#define KS(c, name) (cg::ktime_begin((c), (name)), (c)->stream)
#define CG_LAUNCH_CHECK(c)                                                    \
  do {                                                                        \
    (c)->launches++;                                                          \
    if ((c)->ktiming) cg::ktime_end(c);                                       \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) return cg::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
  } while (0)

struct Ctx;
void ktime_begin(Ctx* c, const char* name);
void ktime_end(Ctx* c);
inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// buffer lookup: device pointer of one scene's buffer, element size and count
struct BufInfo {
  void* ptr;
  size_t elem_bytes;
  long long count;
};
int buf_info(Ctx* c, int buf, int scene, BufInfo* out);
double* vec_buf(Ctx* c, int buf);  // scene 0 pointer of a np*3 vector buffer, or null

// phase timing helpers (no-ops unless profiling is on)
void phase_begin(Ctx* c, int phase);
void phase_end(Ctx* c, int phase);

// ---- kernels' host launchers (one translation unit each)
int launch_assemble(Ctx* c, const clothgpu_params& p, bool with_hessian);
int launch_update_velocity(Ctx* c, const clothgpu_params& p);
int launch_matvec(Ctx* c, double alpha, const double* x, double beta, double* y);
int launch_dot(Ctx* c, const double* a, const double* b, int scal_slot);
int launch_axpy(Ctx* c, double a, const double* x, double* y);
int launch_ilu_setvalue(Ctx* c);
int launch_ilu_factor(Ctx* c);
int launch_ilu_solve(Ctx* c, double* vec);
int launch_ilu_solve_strips(Ctx* c, double* vec, const double* src, bool skip_done);
int build_triplan_host(int np, const int* rowptr, const int* colidx, const int* dia, bool fwd, int force_ns,
                       TriPlanHost& P);
int verify_triplan_host(int np, const int* rowptr, const int* colidx, const int* dia, const TriPlanHost& P);
int upload_triplan(Ctx* c, const TriPlanHost& H, TriPlan& D);
void free_triplan(TriPlan& D);
int build_facplan_host(int np, const int* rowptr, const int* colidx, const int* dia, const TriPlanHost& T, FacPlanHost& F);
int verify_facplan_host(int np, const int* rowptr, const int* colidx, const int* dia, const TriPlanHost& T, const FacPlanHost& F);
int upload_facplan(Ctx* c, const FacPlanHost& H, FacPlan& D);
void free_facplan(FacPlan& D);
int launch_ilu_factor_strips(Ctx* c);
int run_pcg(Ctx* c, double tol, int max_iter, double* conv_out, int* iter_out);
int run_cg(Ctx* c, double tol, int max_iter, double* conv_out, int* iter_out);
int run_step_internal(Ctx* c, const clothgpu_params& p, bool use_ilu, double* conv_out, int* iter_out);
int launch_refit(Ctx* c, const double* xyz, const double* uvw, double par, int mode, bool incremental = false);
int trav_all_max_scenes();
int run_contacts(Ctx* c, const double* xyz, const double* uvw, double dt, double delta, int ccd, int* counts,
                 bool prelaunched = false, bool incremental = false);
int launch_contacts_async(Ctx* c, const double* xyz, const double* uvw, double dt, double delta, int ccd);
int set_all_active(Ctx* c, int v);
int launch_impulse(Ctx* c, const double* xyz, double* uvw, double delta, double stiffness, double dt, double mass,
                   int ccd);
int launch_riz_reset(Ctx* c);
int launch_riz_make(Ctx* c);
int launch_riz_apply(Ctx* c, const double* xyz, double* uvw);
int run_resolve(Ctx* c, const clothgpu_params& p, int* applied, int* stats, bool prox_prelaunched = false);
int fetch_counters(Ctx* c);  // cnt -> h_cnt, synchronous
int ensure_scratch(Ctx* c, size_t bytes);
int ensure_sort_capacity(Ctx* c, long long n);
int grow_queues(Ctx* c, int overflow_bits);
int set_active(Ctx* c, const int* active);
int build_topology_device(Ctx* c);  // setup.cu: diaind, slot rows, assembly gather lists
int bvh_build_topdown_host(int ntri, const int* tri, int np, const double* xyz, std::vector<int>& nodes, int* nfallback);  // per-scene participation flags (null: all scenes)

}  // namespace cg

struct clothgpu_ctx : public cg::Ctx {};
