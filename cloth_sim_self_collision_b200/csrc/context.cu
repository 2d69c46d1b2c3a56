// csrc/context.cu -- context life cycle, topology-only precomputation, buffer access, timers.
//
// Replaces the init sequence of self_contact_sparse/main.cpp:561-583: mat_A.Initialize/SetPattern
// (matrix_square_sparse.cpp:46-58,140-163), ilu_A.Initialize_ILU0 (ilu_sparse.cpp:114-134) and the topology
// inputs of the self-collision (aEdge, aNodeBVH).  Everything that depends on topology only is derived once here
// on the host and uploaded: the ILU dependency levels and update lists, the assembly gather lists, the BVH
// preorder ranks.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

#include <nvtx3/nvToolsExt.h>

#include "ctx.h"

namespace cg {

static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  g_err = buf;
  return CLOTHGPU_ERR_CUDA;
}

template <class T>
static int dev_alloc(T** p, long long n) {
  *p = nullptr;
  if (n <= 0) n = 1;
  CG_CUDA(cudaMalloc((void**)p, (size_t)n * sizeof(T)));
  return CLOTHGPU_OK;
}
template <class T>
static int dev_upload(T** p, const std::vector<T>& h) {
  CG_CHECK(dev_alloc(p, (long long)h.size()));
  if (!h.empty()) CG_CUDA(cudaMemcpy(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return CLOTHGPU_OK;
}
template <class T>
static int dev_upload(T** p, const T* h, long long n) {
  CG_CHECK(dev_alloc(p, n));
  if (n > 0) CG_CUDA(cudaMemcpy(*p, h, (size_t)n * sizeof(T), cudaMemcpyHostToDevice));
  return CLOTHGPU_OK;
}
template <class T>
static int dev_zeros(T** p, long long n) {
  CG_CHECK(dev_alloc(p, n));
  CG_CUDA(cudaMemset(*p, 0, (size_t)std::max<long long>(n, 1) * sizeof(T)));
  return CLOTHGPU_OK;
}

static void free_queues(Ctx* c) {
  cudaFree(c->frontA); cudaFree(c->frontB); cudaFree(c->leaf); cudaFree(c->cand); cudaFree(c->raw); cudaFree(c->raw_sorted);
  c->cand = nullptr;
  cudaFree(c->elems); cudaFree(c->imp_level); cudaFree(c->imp_order); cudaFree(c->imp_pred); cudaFree(c->imp_lptr);
  c->frontA = c->frontB = c->leaf = nullptr;
  c->raw = c->raw_sorted = nullptr;
  c->elems = nullptr;
  c->imp_level = c->imp_order = c->imp_pred = c->imp_lptr = nullptr;
}
static int alloc_queues(Ctx* c) {
  CG_CHECK(dev_alloc(&c->frontA, c->cap_front * c->B));
  CG_CHECK(dev_alloc(&c->frontB, c->cap_front * c->B));
  CG_CHECK(dev_alloc(&c->leaf, c->cap_leaf * c->B));
  CG_CHECK(dev_alloc(&c->cand, c->cap_leaf * 2 * c->B));  // int4 records
  CG_CHECK(dev_alloc(&c->raw, c->cap_raw * c->B));
  CG_CHECK(dev_alloc(&c->raw_sorted, c->cap_raw * c->B));
  CG_CHECK(dev_alloc(&c->elems, c->cap_elem * 5 * c->B));
  CG_CHECK(dev_alloc(&c->imp_level, c->cap_elem * c->B));
  CG_CHECK(dev_alloc(&c->imp_order, c->cap_elem * c->B));
  CG_CHECK(dev_alloc(&c->imp_pred, c->cap_elem * 4 * c->B));
  CG_CHECK(dev_alloc(&c->imp_lptr, (c->cap_elem + 2) * 2 * c->B));
  return CLOTHGPU_OK;
}
// grow the traversal queues whose overflow bit is set (bit 0 frontier, 1 leaf pairs, 2 raw, 3 elems)
int grow_queues(Ctx* c, int bits) {
  CG_CUDA(cudaStreamSynchronize(c->stream));
  const long long hard = 1LL << 28;
  if (!(bits & 15)) return CLOTHGPU_OK;  // bit 4 (k_query's per-leaf frontier): nothing to grow, the retry is a full pass
  if (bits & 1) c->cap_front *= 2;
  if (bits & 2) c->cap_leaf *= 2;
  if (bits & 4) c->cap_raw *= 2;
  if (bits & (4 | 8)) c->cap_elem *= 2;
  if (c->cap_front > hard || c->cap_leaf > hard || c->cap_raw > hard || c->cap_elem > hard) {
    set_error("contact work queue exceeded its hard limit");
    return CLOTHGPU_ERR_CAPACITY;
  }
  free_queues(c);
  return alloc_queues(c);
}

int ensure_scratch(Ctx* c, size_t bytes) {
  if (bytes <= c->scratch_bytes) return CLOTHGPU_OK;
  CG_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(c->scratch);
  c->scratch = nullptr;
  c->scratch_bytes = 0;
  bytes = bytes + bytes / 4 + 4096;
  CG_CUDA(cudaMalloc(&c->scratch, bytes));
  c->scratch_bytes = bytes;
  return CLOTHGPU_OK;
}
int ensure_sort_capacity(Ctx* c, long long n) {
  if (n <= c->cap_sort) return CLOTHGPU_OK;
  CG_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(c->sort_keys); cudaFree(c->sort_keys2); cudaFree(c->sort_vals); cudaFree(c->sort_vals2);
  c->cap_sort = n + n / 4 + 1024;
  CG_CHECK(dev_alloc(&c->sort_keys, c->cap_sort));
  CG_CHECK(dev_alloc(&c->sort_keys2, c->cap_sort));
  CG_CHECK(dev_alloc(&c->sort_vals, c->cap_sort));
  CG_CHECK(dev_alloc(&c->sort_vals2, c->cap_sort));
  return CLOTHGPU_OK;
}

int fetch_counters(Ctx* c) {
  CG_CUDA(cudaMemcpyAsync(c->h_cnt, c->cnt, sizeof(int) * NCNT * c->B, cudaMemcpyDeviceToHost, c->stream));
  CG_CUDA(cudaStreamSynchronize(c->stream));
  return CLOTHGPU_OK;
}

void ktime_begin(Ctx* c, const char* name) {
  if (!c->ktiming) return;
  c->kt_name = name;
  cudaEventRecord(c->kt0, c->stream);
}
void ktime_end(Ctx* c) {
  if (!c->ktiming || !c->kt_name) return;
  cudaEventRecord(c->kt1, c->stream);
  cudaEventSynchronize(c->kt1);
  float ms = 0;
  cudaEventElapsedTime(&ms, c->kt0, c->kt1);
  auto& e = c->ktimes[c->kt_name];
  e.first += ms;
  e.second += 1;
  static const bool trace = getenv("CLOTHGPU_KTRACE") != nullptr;  // one line per launch (measurement aid)
  if (trace) fprintf(stderr, "ktrace %s %.3f us\n", c->kt_name, ms * 1000.0);
  c->kt_name = nullptr;
}

// phases of a step: NVTX ranges always (free without a tool attached; nsys / ncu --nvtx show them), device timers
// only in profiling mode (they synchronise)
static const char* const kPhaseName[CLOTHGPU_NPHASE] = {"assembly", "ilu_factor", "pcg", "bvh_refit", "traverse_narrow",
                                                        "sort_unique", "impulse", "riz", "velocity_fixup"};
void phase_begin(Ctx* c, int phase) {
  nvtxRangePushA((phase >= 0 && phase < CLOTHGPU_NPHASE && kPhaseName[phase]) ? kPhaseName[phase] : "clothgpu");
  if (c->profiling) cudaEventRecord(c->evp0, c->stream);
}
void phase_end(Ctx* c, int phase) {
  nvtxRangePop();
  if (!c->profiling) return;
  cudaEventRecord(c->evp1, c->stream);
  cudaEventSynchronize(c->evp1);
  float ms = 0;
  cudaEventElapsedTime(&ms, c->evp0, c->evp1);
  c->phase_ms[phase] += ms;
}

double* vec_buf(Ctx* c, int buf) {
  switch (buf) {
    case CLOTHGPU_BUF_XYZ: return c->xyz;
    case CLOTHGPU_BUF_UVW: return c->uvw;
    case CLOTHGPU_BUF_XYZ_PREV: return c->xyz_prev;
    case CLOTHGPU_BUF_UVW0: return c->uvw0;
    case CLOTHGPU_BUF_VEC_R: return c->vec_r;
    case CLOTHGPU_BUF_VEC_X: return c->vec_x;
    case CLOTHGPU_BUF_VEC_P: return c->vec_p;
    case CLOTHGPU_BUF_VEC_Z: return c->vec_z;
    default: return nullptr;
  }
}

int buf_info(Ctx* c, int buf, int scene, BufInfo* o) {
  if (scene < 0 || scene >= c->B) {
    set_error("scene index out of range");
    return CLOTHGPU_ERR_ARG;
  }
  const long long s = scene;
  double* v = vec_buf(c, buf);
  if (v) {
    *o = {v + s * c->ndof, sizeof(double), c->ndof};
    return CLOTHGPU_OK;
  }
  switch (buf) {
    case CLOTHGPU_BUF_VAL_DIA: *o = {c->valDia + s * c->np * 9, sizeof(double), (long long)c->np * 9}; break;
    case CLOTHGPU_BUF_VAL_CRS: *o = {c->valCrs + s * c->ncrs * 9, sizeof(double), (long long)c->ncrs * 9}; break;
    case CLOTHGPU_BUF_ILU_DIA: *o = {c->iluDia + s * c->np * 9, sizeof(double), (long long)c->np * 9}; break;
    case CLOTHGPU_BUF_ILU_CRS: *o = {c->iluCrs + s * c->ncrs * 9, sizeof(double), (long long)c->ncrs * 9}; break;
    case CLOTHGPU_BUF_BOXES: *o = {c->boxes + s * c->nnode * 6, sizeof(double), (long long)c->nnode * 6}; break;
    case CLOTHGPU_BUF_ELEMS: *o = {c->elems + s * c->cap_elem * 5, sizeof(int), c->cap_elem * 5}; break;
    case CLOTHGPU_BUF_ZONE: *o = {c->zone_label + s * c->np, sizeof(int), (long long)c->np}; break;
    case CLOTHGPU_BUF_COLLIDER_TABLE: *o = {c->col_table + s * c->np * 4, sizeof(double), (long long)c->np * 4}; break;
    case CLOTHGPU_BUF_LEAF_PAIRS: *o = {c->leaf + s * c->cap_leaf, sizeof(int), c->cap_leaf * 2}; break;
    default:
      set_error("unknown buffer id");
      return CLOTHGPU_ERR_ARG;
  }
  return CLOTHGPU_OK;
}

// ---------------------------------------------------------------------------------------------- precomputation
static int build_topology(Ctx* c, const int* tri, const int* quad, const int* rowptr, const int* colidx) {
  const int np = c->np;
  (void)tri; (void)quad;
  // Initialize_ILU0's m_diaInd (ilu_sparse.cpp:114-134): first slot of the row whose column is above the diagonal
  std::vector<int> dia(np);
  for (int i = 0; i < np; i++) {
    dia[i] = rowptr[i + 1];
    for (int j = rowptr[i]; j < rowptr[i + 1]; j++)
      if (colidx[j] > i) { dia[i] = j; break; }
  }
  // (the device copy of diaind, the slot -> row map and the assembly gather lists are built on the device: setup.cu)
  CG_CHECK(build_topology_device(c));

  // dependency levels of the forward (L) and backward (U) sweeps; rows of one level are mutually independent
  auto levelize = [&](bool fwd, int& nlev, int& maxrows, int** d_ptr, int** d_rows) -> int {
    std::vector<int> lev(np, 0);
    nlev = 0;
    if (fwd) {
      for (int i = 0; i < np; i++) {
        int l = 0;
        for (int j = rowptr[i]; j < dia[i]; j++) l = std::max(l, lev[colidx[j]] + 1);
        lev[i] = l;
        nlev = std::max(nlev, l + 1);
      }
    } else {
      for (int i = np - 1; i >= 0; i--) {
        int l = 0;
        for (int j = dia[i]; j < rowptr[i + 1]; j++) l = std::max(l, lev[colidx[j]] + 1);
        lev[i] = l;
        nlev = std::max(nlev, l + 1);
      }
    }
    std::vector<int> ptr(nlev + 1, 0), rows(np);
    for (int i = 0; i < np; i++) ptr[lev[i] + 1]++;
    maxrows = 0;
    for (int l = 0; l < nlev; l++) {
      maxrows = std::max(maxrows, ptr[l + 1]);
      ptr[l + 1] += ptr[l];
    }
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (int i = 0; i < np; i++) rows[fill[lev[i]]++] = i;
    CG_CHECK(dev_upload(d_ptr, ptr));
    CG_CHECK(dev_upload(d_rows, rows));
    return CLOTHGPU_OK;
  };
  CG_CHECK(levelize(true, c->nlev_f, c->maxrows_f, &c->lev_f_ptr, &c->lev_f_rows));
  CG_CHECK(levelize(false, c->nlev_b, c->maxrows_b, &c->lev_b_ptr, &c->lev_b_rows));
  {
    const char* mode = getenv("CLOTHGPU_TRISOLVE");
    c->use_strips = !(mode && strcmp(mode, "levels") == 0);
    c->use_strips_fac = c->use_strips;
    {
      const char* tw = getenv("CLOTHGPU_TRI_WARPS");
      c->tri_two_warps = tw && atoi(tw) == 2;
    }
    const char* fs = getenv("CLOTHGPU_STRIPS");
    const int force_ns = fs ? atoi(fs) : 0;
    for (int dir = 0; dir < 2 && c->use_strips; dir++) {
      TriPlanHost H;
      CG_CHECK(build_triplan_host(np, rowptr, colidx, dia.data(), dir == 0, force_ns, H));
      CG_CHECK(upload_triplan(c, H, dir == 0 ? c->plan_f : c->plan_b));
      if (dir == 0) {
        // the strip-pipelined factorisation keeps the imports of FC_NIB micro-steps in shared memory; patterns
        // without band structure (hundreds of far operands per micro-step) do not fit -- those contexts factorise
        // with the level-scheduled kernel (k_ilu_factor) instead, same arithmetic
        FacPlanHost F;
        // the factorisation may use narrower strips than the forward sweep (CLOTHGPU_FAC_STRIPS; default below): its
        // micro-step is bound by shared-memory traffic per row, so fewer rows per step and CTA shorten the step while
        // the number of steps on the critical path (the dependency levels) stays the same
        // (a single scene leaves most SMs idle, so it takes twice the strips; batches are throughput bound and keep
        // the forward sweep's strips: every CTA pays the fixed per-step cost)
        const char* ffs = getenv("CLOTHGPU_FAC_STRIPS");
        int fac_ns = ffs ? atoi(ffs) : (c->B == 1 && 2 * H.ns <= 148 ? 2 * H.ns : 0);
        TriPlanHost HF;
        if (fac_ns > 0 && fac_ns != H.ns && build_triplan_host(np, rowptr, colidx, dia.data(), true, fac_ns, HF) != CLOTHGPU_OK) {
          if (ffs) return CLOTHGPU_ERR_ARG;  // a forced count must fit
          fac_ns = 0;
        }
        CG_CHECK(build_facplan_host(np, rowptr, colidx, dia.data(), (fac_ns > 0 && fac_ns != H.ns) ? HF : H, F));
        const int rc = upload_facplan(c, F, c->plan_fac);
        if (rc == CLOTHGPU_ERR_CAPACITY) {
          free_facplan(c->plan_fac);
          c->use_strips_fac = false;
        } else if (rc != CLOTHGPU_OK) {
          return rc;
        }
      }
    }
  }

  // update list of DoILUDecomp (ilu_sparse.cpp:507-540): for row i, lower blocks k ascending, upper blocks j of
  // row k ascending, target (i,j) if present in the pattern (or the diagonal when j == i)
  {
    std::vector<int> op_ptr(np + 1, 0);
    std::vector<int4> ops;
    std::vector<int> slot(np, -1);
    for (int i = 0; i < np; i++) {
      for (int j = rowptr[i]; j < rowptr[i + 1]; j++) slot[colidx[j]] = j;
      for (int ik = rowptr[i]; ik < dia[i]; ik++) {
        const int k = colidx[ik];
        for (int kj = dia[k]; kj < rowptr[k + 1]; kj++) {
          const int jc = colidx[kj];
          int ij;
          if (jc == i) ij = -1;
          else if (slot[jc] >= 0) ij = slot[jc];
          else continue;
          ops.push_back(make_int4(ik, kj, ij, ik - rowptr[i]));
        }
      }
      for (int j = rowptr[i]; j < rowptr[i + 1]; j++) slot[colidx[j]] = -1;
      op_ptr[i + 1] = (int)ops.size();
    }
    c->nops = (long long)ops.size();
    CG_CHECK(dev_upload(&c->op_ptr, op_ptr));
    CG_CHECK(dev_upload(&c->ops, ops));
  }

  return CLOTHGPU_OK;
}

static int build_bvh(Ctx* c, const int* nodes, const int* tri) {
  const int nn = c->nnode;
  // preorder rank in the order of the self traversal: self(n) visits n, then self(child0), self(child1)
  // (bvh_aabb.cpp:540-557); explicit stack, the trees are ~20 deep but the recursion is not needed
  std::vector<uint32_t> pre(nn, 0);
  std::vector<unsigned long long> npath(nn, 0ULL);
  std::vector<int> depth(nn, 0), internal;
  std::vector<int> stack;
  stack.push_back(0);
  uint32_t r = 0;
  int height = 0, visited = 0;
  while (!stack.empty()) {
    const int i = stack.back();
    stack.pop_back();
    if (i < 0 || i >= nn) { set_error("BVH child index out of range"); return CLOTHGPU_ERR_ARG; }
    pre[i] = r++;
    visited++;
    if (visited > nn) { set_error("BVH topology is not a tree"); return CLOTHGPU_ERR_ARG; }
    height = std::max(height, depth[i]);
    const int c0 = nodes[i * 3 + 1], c1 = nodes[i * 3 + 2];
    if (c1 == -1) {
      if (c0 < 0 || c0 >= c->ntri) { set_error("BVH leaf triangle index out of range"); return CLOTHGPU_ERR_ARG; }
      continue;
    }
    if (c0 < 0 || c0 >= nn || c1 < 0 || c1 >= nn) { set_error("BVH child index out of range"); return CLOTHGPU_ERR_ARG; }
    if (nodes[c0 * 3] != i || nodes[c1 * 3] != i) { set_error("BVH parent links are inconsistent"); return CLOTHGPU_ERR_ARG; }
    internal.push_back(i);
    depth[c0] = depth[c1] = depth[i] + 1;
    npath[c0] = npath[i];
    npath[c1] = npath[i] | (1ULL << (63 - depth[i]));
    stack.push_back(c1);
    stack.push_back(c0);
  }
  if (visited != nn) { set_error("BVH has unreachable nodes"); return CLOTHGPU_ERR_ARG; }
  if (height > 60) { set_error("BVH deeper than 60 levels is not supported"); return CLOTHGPU_ERR_ARG; }
  c->bvh_height = height;
  c->ninternal = (int)internal.size();
  CG_CHECK(dev_upload(&c->nodes, nodes, (long long)nn * 3));
  {
    // leaf node -> {vertex 0, 1, 2, triangle}: one 16-byte load in the narrow phase instead of node -> triangle -> vertices
    std::vector<int> nt((size_t)nn * 4, -1);
    for (int i = 0; i < nn; i++) {
      if (nodes[i * 3 + 2] != -1) continue;
      const int t = nodes[i * 3 + 1];
      if (t < 0 || t >= c->ntri) { set_error("BVH leaf refers to a triangle out of range"); return CLOTHGPU_ERR_ARG; }
      nt[(size_t)i * 4 + 0] = tri[t * 3 + 0]; nt[(size_t)i * 4 + 1] = tri[t * 3 + 1]; nt[(size_t)i * 4 + 2] = tri[t * 3 + 2];
      nt[(size_t)i * 4 + 3] = t;
    }
    CG_CHECK(dev_upload(&c->node_tri, nt));
  }
  CG_CHECK(dev_upload(&c->pre, pre));
  CG_CHECK(dev_upload(&c->npath, npath));
  CG_CHECK(dev_upload(&c->ndepth, depth));
  CG_CHECK(dev_upload(&c->internal_list, internal));
  return CLOTHGPU_OK;
}

static int create_impl(Ctx* c, int nscene, int np, int ntri, int nquad, const int* tri, const int* quad,
                       const int* bcflag, const double* xyz0, const int* rowptr, const int* colidx, int nnode,
                       const int* bvh_nodes, const int* edge_index, const int* edge_array) {
  c->B = nscene; c->np = np; c->ntri = ntri; c->nquad = nquad; c->ncrs = rowptr[np]; c->nnode = nnode;
  c->ndof = np * 3;
  const long long B = nscene;
  CG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CG_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  CG_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CG_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  CG_CUDA(cudaEventCreate(&c->ev0)); CG_CUDA(cudaEventCreate(&c->ev1));
  CG_CUDA(cudaEventCreate(&c->evp0)); CG_CUDA(cudaEventCreate(&c->evp1));
  CG_CUDA(cudaEventCreate(&c->kt0)); CG_CUDA(cudaEventCreate(&c->kt1));
  CG_CHECK(dev_upload(&c->tri, tri, (long long)ntri * 3));
  CG_CHECK(dev_upload(&c->quad, quad, (long long)nquad * 4));
  CG_CHECK(dev_upload(&c->bc, bcflag, np));
  CG_CHECK(dev_upload(&c->xyz0, xyz0, (long long)np * 3));
  CG_CHECK(dev_upload(&c->rowptr, rowptr, np + 1));
  CG_CHECK(dev_upload(&c->colidx, colidx, c->ncrs));
  CG_CHECK(build_topology(c, tri, quad, rowptr, colidx));
  CG_CHECK(dev_alloc(&c->quadK, (long long)nquad * 4));
  CG_CHECK(dev_alloc(&c->quadT, nquad));

  CG_CHECK(dev_alloc(&c->xyz, B * c->ndof)); CG_CHECK(dev_zeros(&c->uvw, B * c->ndof));
  CG_CHECK(dev_alloc(&c->xyz_prev, B * c->ndof)); CG_CHECK(dev_zeros(&c->uvw0, B * c->ndof));
  for (int s = 0; s < nscene; s++) {
    CG_CUDA(cudaMemcpy(c->xyz + (long long)s * c->ndof, xyz0, sizeof(double) * c->ndof, cudaMemcpyHostToDevice));
    CG_CUDA(cudaMemcpy(c->xyz_prev + (long long)s * c->ndof, xyz0, sizeof(double) * c->ndof, cudaMemcpyHostToDevice));
  }
  CG_CHECK(dev_zeros(&c->valDia, B * np * 9)); CG_CHECK(dev_zeros(&c->valCrs, B * c->ncrs * 9));
  CG_CHECK(dev_zeros(&c->iluDia, B * np * 9)); CG_CHECK(dev_zeros(&c->iluCrs, B * c->ncrs * 9));
  CG_CHECK(dev_zeros(&c->vec_r, B * c->ndof)); CG_CHECK(dev_zeros(&c->vec_x, B * c->ndof));
  CG_CHECK(dev_zeros(&c->vec_p, B * c->ndof)); CG_CHECK(dev_zeros(&c->vec_z, B * c->ndof));
  CG_CHECK(dev_alloc(&c->tri_dW, B * ntri * 9)); CG_CHECK(dev_alloc(&c->tri_ddW, B * ntri * 81));
  CG_CHECK(dev_alloc(&c->quad_dW, B * nquad * 12));
  CG_CHECK(dev_zeros(&c->col_table, B * np * 4));
  CG_CHECK(dev_zeros(&c->gravity, B * 3));
  c->gravity_override.assign((size_t)B * 3, std::numeric_limits<double>::quiet_NaN());
  c->npartial = 296;
  CG_CHECK(dev_zeros(&c->scal, B * NSCAL)); CG_CHECK(dev_zeros(&c->partials, B * c->npartial * 4));
  c->xd_nch = div_up(c->ndof, 256);
  CG_CHECK(dev_zeros(&c->xd_tsum, (long long)B * c->xd_nch));
  {
    unsigned char* st = nullptr;
    CG_CHECK(dev_zeros(&st, (long long)B * c->xd_nch * 32));
    c->xd_stat = st;
    const char* dm = getenv("CLOTHGPU_DOT");
#ifdef CLOTHGPU_FAST_MODE
    c->exact_dot = dm && strcmp(dm, "exact") == 0;  // the fast build reduces with trees unless asked otherwise
#else
    c->exact_dot = !(dm && strcmp(dm, "tree") == 0);
#endif
  }
  CG_CHECK(dev_zeros(&c->cnt, B * NCNT));
  CG_CUDA(cudaMallocHost((void**)&c->h_cnt, sizeof(int) * NCNT * B));
  CG_CUDA(cudaMallocHost((void**)&c->h_scal, sizeof(double) * NSCAL * B));
  memset(c->h_cnt, 0, sizeof(int) * NCNT * B);
  CG_CHECK(dev_alloc(&c->zone, B * np)); CG_CHECK(dev_alloc(&c->zone_label, B * np));
  CG_CUDA(cudaMemset(c->zone, 0xff, sizeof(int) * B * np));
  CG_CUDA(cudaMemset(c->zone_label, 0xff, sizeof(int) * B * np));

  if (nnode > 0) {
    if (nnode != 2 * ntri - 1) { set_error("nnode must be 2*ntri-1"); return CLOTHGPU_ERR_ARG; }
    if (!bvh_nodes || !edge_index || !edge_array) { set_error("bvh_nodes / edge arrays missing"); return CLOTHGPU_ERR_ARG; }
    CG_CHECK(build_bvh(c, bvh_nodes, tri));
    CG_CHECK(dev_upload(&c->edge_index, edge_index, np + 1));
    CG_CHECK(dev_upload(&c->edge_array, edge_array, edge_index[np]));
    CG_CHECK(dev_zeros(&c->boxes, B * nnode * 6));
    CG_CHECK(dev_zeros(&c->refit_flag, B * nnode));
    CG_CHECK(dev_zeros(&c->uvw_snap, B * c->ndof));
    CG_CHECK(dev_zeros(&c->vdirty, B * np));
    CG_CHECK(dev_zeros(&c->ndirty, B * nnode));
    CG_CHECK(dev_zeros(&c->refit_nchild, B * nnode));
    CG_CHECK(dev_zeros(&c->dirty_list, B * nnode));
    {
      const char* e1 = getenv("CLOTHGPU_INC_REFIT");
      c->inc_refit = !(e1 && atoi(e1) == 0);
      const char* e7 = getenv("CLOTHGPU_SLIM_FRONTIER");
      c->slim_frontier = !(e7 && atoi(e7) == 0);
      const char* e6 = getenv("CLOTHGPU_RIZ_FUSED");
      c->riz_fused = !(e6 && atoi(e6) == 0);
      const char* e5 = getenv("CLOTHGPU_RANK_CHECK");
      c->rank_check = e5 && atoi(e5) != 0;
      const char* e4 = getenv("CLOTHGPU_QUERY_CAP");
      if (e4) c->query_cap = std::max(1, atoi(e4));
      const char* e3 = getenv("CLOTHGPU_CANON_SMALL");
      c->canon_small = !(e3 && atoi(e3) == 0);
      const char* e2 = getenv("CLOTHGPU_QUERY_TRAVERSAL");
      c->query_trav = !(e2 && atoi(e2) == 0);
    }
    {
      const char* inc = getenv("CLOTHGPU_INCREMENTAL");
      c->incremental_ccd = !(inc && atoi(inc) == 0);
    }
    {
      // fused traversal: one CTA per SM and scene, all co-resident (cooperative launch); leave room for the strip
      // kernels that run concurrently on the other stream
      cudaDeviceProp pr;
      CG_CUDA(cudaGetDeviceProperties(&pr, c->device));
      c->trav_sms = pr.multiProcessorCount;
      const char* it = getenv("CLOTHGPU_IMP_THREADS");
      if (it) c->imp_threads = std::min(1024, std::max(32, atoi(it) / 32 * 32));
      const char* isl = getenv("CLOTHGPU_IMP_SLEEP");
      if (isl) c->imp_sleep_ns = std::min(100000, std::max(0, atoi(isl)));
      const char* il = getenv("CLOTHGPU_IMP_LANES");
      if (il) c->imp_lanes = std::min(32, std::max(1, atoi(il)));
      const char* ifa = getenv("CLOTHGPU_IMP_FAST");
      c->imp_fast = !(ifa && atoi(ifa) == 0);
      const char* tc = getenv("CLOTHGPU_FUSED_TRAVERSAL");
      c->trav_coop = (pr.cooperativeLaunch && !(tc && atoi(tc) == 0)) ? trav_all_max_scenes() : 0;
    }
    c->cap_front = std::max<long long>(16LL * ntri, 1 << 14);
    c->cap_leaf = std::max<long long>(16LL * ntri, 1 << 14);
    c->cap_raw = std::max<long long>(4LL * ntri, 1 << 14);
    c->cap_elem = c->cap_raw;
    CG_CHECK(alloc_queues(c));
  }
  CG_CUDA(cudaDeviceSynchronize());
  return CLOTHGPU_OK;
}

}  // namespace cg

using namespace cg;

extern "C" {

const char* clothgpu_last_error(void) { return g_err.c_str(); }
int clothgpu_version(void) { return CLOTHGPU_VERSION; }
const char* clothgpu_build_mode(void) {
#ifdef CLOTHGPU_FAST_MODE
  return "fast: -fmad=true, tree reductions (NOT bit-compatible with the reference)";
#else
  return "parity: -fmad=false, serial-order dot products (bit-compatible with the reference)";
#endif
}
int clothgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int clothgpu_create(clothgpu_ctx** out, int device, int nscene, int np, int ntri, int nquad, const int* tri,
                    const int* quad, const int* bcflag, const double* xyz0, const int* crs_rowptr,
                    const int* crs_colidx, int nnode, const int* bvh_nodes, const int* edge_index,
                    const int* edge_array) {
  if (!out) { set_error("out is null"); return CLOTHGPU_ERR_ARG; }
  *out = nullptr;
  if (nscene < 1 || np < 1 || ntri < 0 || nquad < 0 || !tri || !quad || !bcflag || !xyz0 || !crs_rowptr || !crs_colidx) {
    set_error("clothgpu_create: bad sizes or null arrays");
    return CLOTHGPU_ERR_ARG;
  }
  for (long long i = 0; i < (long long)ntri * 3; i++)
    if (tri[i] < 0 || tri[i] >= np) { set_error("triangle vertex index out of range"); return CLOTHGPU_ERR_ARG; }
  for (long long i = 0; i < (long long)nquad * 4; i++)
    if (quad[i] < 0 || quad[i] >= np) { set_error("hinge vertex index out of range"); return CLOTHGPU_ERR_ARG; }
  if (crs_rowptr[0] != 0) { set_error("crs_rowptr[0] must be 0"); return CLOTHGPU_ERR_ARG; }
  for (int i = 0; i < np; i++) {
    if (crs_rowptr[i + 1] < crs_rowptr[i]) { set_error("crs_rowptr must be non-decreasing"); return CLOTHGPU_ERR_ARG; }
    for (int j = crs_rowptr[i]; j < crs_rowptr[i + 1]; j++) {
      const int col = crs_colidx[j];
      if (col < 0 || col >= np || col == i || (j > crs_rowptr[i] && crs_colidx[j - 1] >= col)) {
        set_error("crs_colidx must be ascending per row, in range and without the diagonal");
        return CLOTHGPU_ERR_ARG;
      }
    }
  }
  if (clothgpu_device_count() == 0) {
    set_error("no CUDA device visible: libclothgpu has no CPU fallback");
    return CLOTHGPU_ERR_CUDA;
  }
  if (device >= 0) CG_CUDA(cudaSetDevice(device));
  clothgpu_ctx* c = new clothgpu_ctx();
  {
    const cudaError_t e = cudaGetDevice(&c->device);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__); }
  }
  const int r = create_impl(c, nscene, np, ntri, nquad, tri, quad, bcflag, xyz0, crs_rowptr, crs_colidx, nnode,
                            bvh_nodes, edge_index, edge_array);
  if (r != CLOTHGPU_OK) {
    clothgpu_destroy(c);
    return r;
  }
  *out = c;
  return CLOTHGPU_OK;
}

void clothgpu_destroy(clothgpu_ctx* c) {
  if (!c) return;
  if (c->stream) cudaStreamSynchronize(c->stream);
  void* ptrs[] = {c->tri, c->quad, c->bc, c->rowptr, c->colidx, c->diaind, c->xyz0, c->vt_ptr, c->vt_ent, c->vq_ptr,
                  c->vq_ent, c->slot_row, c->st_ptr, c->st_ent, c->sq_ptr, c->sq_ent, c->quadK, c->quadT, c->lev_f_ptr,
                  c->lev_f_rows, c->lev_b_ptr, c->lev_b_rows, c->op_ptr, c->ops, c->nodes, c->node_tri, c->pre, c->internal_list,
                  c->edge_index, c->edge_array, c->xyz, c->uvw, c->xyz_prev, c->uvw0, c->valDia, c->valCrs, c->iluDia,
                  c->iluCrs, c->vec_r, c->vec_x, c->vec_p, c->vec_z, c->tri_dW, c->tri_ddW, c->quad_dW, c->col_table,
                  c->gravity, c->boxes, c->refit_flag, c->uvw_snap, c->vdirty, c->ndirty, c->refit_nchild, c->dirty_list, c->npath, c->ndepth, c->scal, c->partials, c->xd_tsum, c->xd_stat, c->cnt, c->zone, c->zone_label,
                  c->scratch, c->sort_keys, c->sort_keys2, c->sort_vals, c->sort_vals2};
  for (void* p : ptrs) cudaFree(p);
  cudaFree(c->tri_dbg_buf);
  cudaFree(c->normals);
  cudaFree(c->tri_xch);
  free_facplan(c->plan_fac);
  free_triplan(c->plan_f);
  free_triplan(c->plan_b);
  free_queues(c);
  if (c->h_cnt) cudaFreeHost(c->h_cnt);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->evp0) cudaEventDestroy(c->evp0);
  if (c->evp1) cudaEventDestroy(c->evp1);
  if (c->kt0) cudaEventDestroy(c->kt0);
  if (c->kt1) cudaEventDestroy(c->kt1);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int clothgpu_nscene(const clothgpu_ctx* c) { return c ? c->B : 0; }

long long clothgpu_buf_len(const clothgpu_ctx* c, int buf) {
  BufInfo bi;
  if (!c || buf_info(const_cast<clothgpu_ctx*>(c), buf, 0, &bi) != CLOTHGPU_OK) return -1;
  return bi.count;
}

int clothgpu_write(clothgpu_ctx* c, int buf, int scene, const void* host, long long count) {
  if (!c || !host) { set_error("null argument"); return CLOTHGPU_ERR_ARG; }
  cudaSetDevice(c->device);
  BufInfo bi;
  CG_CHECK(buf_info(c, buf, scene, &bi));
  if (buf == CLOTHGPU_BUF_LEAF_PAIRS) { set_error("LEAF_PAIRS is read-only"); return CLOTHGPU_ERR_ARG; }
  if (count < 0 || count > bi.count) { set_error("count exceeds buffer length"); return CLOTHGPU_ERR_ARG; }
  if (!bi.ptr) { set_error("buffer not allocated (context created without a BVH?)"); return CLOTHGPU_ERR_STATE; }
  CG_CUDA(cudaMemcpyAsync(bi.ptr, host, (size_t)count * bi.elem_bytes, cudaMemcpyHostToDevice, c->stream));
  CG_CUDA(cudaStreamSynchronize(c->stream));
  if (buf == CLOTHGPU_BUF_ILU_DIA || buf == CLOTHGPU_BUF_ILU_CRS) c->ilu_packed_valid = false;
  if (buf == CLOTHGPU_BUF_ELEMS) {  // the caller supplies the contact elements of the next response op
    if (count % 5) { set_error("ELEMS count must be a multiple of 5"); return CLOTHGPU_ERR_ARG; }
    c->h_cnt[scene * NCNT + CNT_ELEM] = (int)(count / 5);
    CG_CUDA(cudaMemcpyAsync(c->cnt + scene * NCNT + CNT_ELEM, c->h_cnt + scene * NCNT + CNT_ELEM, sizeof(int),
                            cudaMemcpyHostToDevice, c->stream));
    CG_CUDA(cudaStreamSynchronize(c->stream));
  }
  return CLOTHGPU_OK;
}

int clothgpu_read(clothgpu_ctx* c, int buf, int scene, void* host, long long count) {
  if (!c || !host) { set_error("null argument"); return CLOTHGPU_ERR_ARG; }
  cudaSetDevice(c->device);
  BufInfo bi;
  CG_CHECK(buf_info(c, buf, scene, &bi));
  if (count < 0 || count > bi.count) { set_error("count exceeds buffer length"); return CLOTHGPU_ERR_ARG; }
  if (!bi.ptr) { set_error("buffer not allocated (context created without a BVH?)"); return CLOTHGPU_ERR_STATE; }
  if (buf == CLOTHGPU_BUF_LEAF_PAIRS) {  // strip the traversal bookkeeping: keep the two node ids per entry
    const long long n = count / 2;
    if (c->leaf_slim) {  // the last traversal wrote 8-byte pairs, scene stride cap_leaf pairs
      const int2* src = reinterpret_cast<const int2*>(c->leaf) + (long long)scene * c->cap_leaf;
      CG_CUDA(cudaMemcpyAsync(host, src, (size_t)n * sizeof(int2), cudaMemcpyDeviceToHost, c->stream));
      CG_CUDA(cudaStreamSynchronize(c->stream));
      return CLOTHGPU_OK;
    }
    std::vector<PairEnt> tmp((size_t)n);
    CG_CUDA(cudaMemcpyAsync(tmp.data(), bi.ptr, (size_t)n * sizeof(PairEnt), cudaMemcpyDeviceToHost, c->stream));
    CG_CUDA(cudaStreamSynchronize(c->stream));
    int* o = (int*)host;
    for (long long i = 0; i < n; i++) { o[i * 2] = tmp[i].i0; o[i * 2 + 1] = tmp[i].i1; }
    return CLOTHGPU_OK;
  }
  CG_CUDA(cudaMemcpyAsync(host, bi.ptr, (size_t)count * bi.elem_bytes, cudaMemcpyDeviceToHost, c->stream));
  CG_CUDA(cudaStreamSynchronize(c->stream));
  return CLOTHGPU_OK;
}

int clothgpu_upload_state(clothgpu_ctx* c, int scene, const double* xyz, const double* uvw) {
  if (!c || !xyz || !uvw || scene < 0 || scene >= c->B) { set_error("bad argument"); return CLOTHGPU_ERR_ARG; }
  cudaSetDevice(c->device);
  CG_CUDA(cudaMemcpyAsync(c->xyz + (long long)scene * c->ndof, xyz, sizeof(double) * c->ndof, cudaMemcpyHostToDevice, c->stream));
  CG_CUDA(cudaMemcpyAsync(c->uvw + (long long)scene * c->ndof, uvw, sizeof(double) * c->ndof, cudaMemcpyHostToDevice, c->stream));
  CG_CUDA(cudaStreamSynchronize(c->stream));
  return CLOTHGPU_OK;
}
int clothgpu_download_state(clothgpu_ctx* c, int scene, double* xyz, double* uvw) {
  if (!c || !xyz || !uvw || scene < 0 || scene >= c->B) { set_error("bad argument"); return CLOTHGPU_ERR_ARG; }
  cudaSetDevice(c->device);
  CG_CUDA(cudaMemcpyAsync(xyz, c->xyz + (long long)scene * c->ndof, sizeof(double) * c->ndof, cudaMemcpyDeviceToHost, c->stream));
  CG_CUDA(cudaMemcpyAsync(uvw, c->uvw + (long long)scene * c->ndof, sizeof(double) * c->ndof, cudaMemcpyDeviceToHost, c->stream));
  CG_CUDA(cudaStreamSynchronize(c->stream));
  return CLOTHGPU_OK;
}

int clothgpu_alloc_host(void** ptr, long long bytes) {
  if (!ptr || bytes <= 0) { set_error("bad argument"); return CLOTHGPU_ERR_ARG; }
  *ptr = nullptr;
  CG_CUDA(cudaMallocHost(ptr, (size_t)bytes));
  return CLOTHGPU_OK;
}
int clothgpu_free_host(void* ptr) {
  if (ptr) CG_CUDA(cudaFreeHost(ptr));
  return CLOTHGPU_OK;
}

int clothgpu_set_scene_gravity(clothgpu_ctx* c, int scene, const double* g) {
  if (!c) { set_error("null context"); return CLOTHGPU_ERR_ARG; }
  if (scene < 0) {
    c->gravity_override.assign((size_t)c->B * 3, std::numeric_limits<double>::quiet_NaN());
    return CLOTHGPU_OK;
  }
  if (scene >= c->B || !g) { set_error("bad argument"); return CLOTHGPU_ERR_ARG; }
  for (int k = 0; k < 3; k++) c->gravity_override[(size_t)scene * 3 + k] = g[k];
  return CLOTHGPU_OK;
}

int clothgpu_last_solver(clothgpu_ctx* c, double* conv_ratio, int* iteration) {
  if (!c) { set_error("null context"); return CLOTHGPU_ERR_ARG; }
  for (int s = 0; s < c->B; s++) {
    if (conv_ratio) conv_ratio[s] = c->h_scal[s * NSCAL + SC_CONV];
    if (iteration) iteration[s] = c->h_cnt[s * NCNT + CNT_PCG_ITER];
  }
  return CLOTHGPU_OK;
}

int clothgpu_synchronize(clothgpu_ctx* c) {
  if (!c) { set_error("null context"); return CLOTHGPU_ERR_ARG; }
  cudaSetDevice(c->device);
  CG_CUDA(cudaStreamSynchronize(c->stream));
  return CLOTHGPU_OK;
}

int clothgpu_timer_start(clothgpu_ctx* c) {
  if (!c) { set_error("null context"); return CLOTHGPU_ERR_ARG; }
  cudaSetDevice(c->device);
  CG_CUDA(cudaStreamSynchronize(c->stream));
  CG_CUDA(cudaEventRecord(c->ev0, c->stream));
  return CLOTHGPU_OK;
}
int clothgpu_timer_stop(clothgpu_ctx* c, float* ms) {
  if (!c || !ms) { set_error("null argument"); return CLOTHGPU_ERR_ARG; }
  cudaSetDevice(c->device);
  CG_CUDA(cudaEventRecord(c->ev1, c->stream));
  CG_CUDA(cudaEventSynchronize(c->ev1));
  CG_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return CLOTHGPU_OK;
}
long long clothgpu_launch_count(const clothgpu_ctx* c) { return c ? c->launches : 0; }

int clothgpu_set_profiling(clothgpu_ctx* c, int on) {
  if (!c) { set_error("null context"); return CLOTHGPU_ERR_ARG; }
  c->profiling = on != 0;
  for (double& v : c->phase_ms) v = 0;
  return CLOTHGPU_OK;
}
int clothgpu_set_kernel_timing(clothgpu_ctx* c, int on) {
  if (!c) { set_error("null context"); return CLOTHGPU_ERR_ARG; }
  cudaSetDevice(c->device);
  CG_CUDA(cudaStreamSynchronize(c->stream));
  c->ktiming = on != 0;
  c->ktimes.clear();
  return CLOTHGPU_OK;
}
int clothgpu_kernel_times(clothgpu_ctx* c, char* out, int cap) {
  if (!c || !out || cap < 1) { set_error("bad argument"); return CLOTHGPU_ERR_ARG; }
  std::string s;
  char line[256];
  for (const auto& kv : c->ktimes) {
    snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.second, kv.second.first);
    s += line;
  }
  if ((int)s.size() + 1 > cap) { set_error("kernel_times: buffer too small"); return CLOTHGPU_ERR_ARG; }
  memcpy(out, s.c_str(), s.size() + 1);
  return CLOTHGPU_OK;
}
int clothgpu_phase_ms(clothgpu_ctx* c, double* ms) {
  if (!c || !ms) { set_error("null argument"); return CLOTHGPU_ERR_ARG; }
  for (int i = 0; i < CLOTHGPU_NPHASE; i++) { ms[i] = c->phase_ms[i]; c->phase_ms[i] = 0; }
  return CLOTHGPU_OK;
}

}  // extern "C"
