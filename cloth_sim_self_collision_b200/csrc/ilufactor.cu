// csrc/ilufactor.cu -- block ILU(0) factorisation as a strip-pipelined wavefront.
//
// Replaces CPreconditionerILU::DoILUDecomp (ilu_sparse.cpp:387-624, len == 3 at :507-569; CalcInvMat3 :77-97).
//
// Row i of the factorisation needs the finished upper parts U'_k = D_k^-1 U_k of its lower neighbours k, so the
// rows form the same dependency wavefront as the forward substitution (1021 levels on a 256^2 grid).  The kernel
// reuses the forward sweep's strips and micro-steps (trisolve.cu): one CTA per strip, <= 10 rows per micro-step,
// three lanes per row -- lane r owns block-row r of every 3x3 block of its matrix row, so the update
// A_ij -= A_ik U'_kj needs no cross-lane traffic.  Per micro-step
//   * the row's blocks (packed by k_pack into 32-byte padded block rows) arrive by one TMA bulk copy per stage;
//   * the update list (ik, kj, ij) in the reference's order is applied in shared memory;
//   * operands U'_k produced by this strip within the last FC_H micro-steps are read from a shared-memory ring;
//     older ones and those of other strips are fetched by an import lane with TMA bulk copies from a padded
//     global copy of the upper parts ("uexp"), gated by the producer strip's progress counter;
//   * D_i is inverted by cofactors, U'_i = D_i^-1 U_i is written to the ring, to the ILU arrays and to uexp.
// Arithmetic per entry is the reference's (same operand order, -fmad=false): results are bit identical.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "ctx.h"

namespace cg {

namespace {
inline int round_up(int v, int a) { return (v + a - 1) / a * a; }
}  // namespace

// ------------------------------------------------------------------------------------------------ host: plan
int build_facplan_host(int np, const int* rowptr, const int* colidx, const int* dia, const TriPlanHost& T,
                       FacPlanHost& F) {
  const int ncrs = rowptr[np];
  F.ns = T.ns; F.R = T.R; F.nms = T.nms; F.strip_ms = T.strip_ms; F.ms_rows = T.ms_rows;
  F.H = FC_H;
  int gB = 0, gU = 0;
  for (int i = 0; i < np; i++) {
    gB = std::max(gB, rowptr[i + 1] - rowptr[i] + 1);
    gU = std::max(gU, rowptr[i + 1] - dia[i]);
  }
  F.gmaxB = gB;
  F.gmaxU = std::max(gU, 1);
  const int gstride = (gB + 1) * 96;  // bytes of one row group in the scratch: its blocks + one dummy block
  const int urow = F.gmaxU * 96;       // bytes of one exported upper part
  // update lists in the reference's order (DoILUDecomp, ilu_sparse.cpp:507-540)
  std::vector<int> op_ptr(np + 1, 0);
  std::vector<int> op_ik, op_kj, op_ij;  // crs slots; ij == -1: diagonal
  {
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
          op_ik.push_back(ik); op_kj.push_back(kj); op_ij.push_back(ij);
        }
      }
      for (int j = rowptr[i]; j < rowptr[i + 1]; j++) slot[colidx[j]] = -1;
      op_ptr[i + 1] = (int)op_ik.size();
    }
  }
  int gOps = 0;
  for (int i = 0; i < np; i++) gOps = std::max(gOps, op_ptr[i + 1] - op_ptr[i]);
  F.gmaxOps = gOps;
  F.meta_stride = round_up(FC_META_HDR + gOps * TS_ROWS * 4, 4);
  F.row_stride = TS_ROWS * (gB + 1) * 12;  // doubles per micro-step of packed row data
  F.meta.assign((size_t)F.nms * F.meta_stride, 0);
  F.pk_src.assign((size_t)F.nms * F.row_stride, -1);
  F.imp_ptr.assign(F.nms + 1, 0);
  F.imp_row.clear(); F.imp_src.clear(); F.imp_need.clear();
  F.max_imp = 0;
  // where each row is produced
  std::vector<int> ms_of_row(np, -1), q_of_row(np, -1);
  for (int m = 0; m < F.nms; m++)
    for (int q = 0; q < TS_ROWS; q++) {
      const int row = F.ms_rows[(size_t)m * TS_ROWS + q];
      if (row >= 0) { ms_of_row[row] = m; q_of_row[row] = q; }
    }
  const int ring_bytes = FC_H * TS_ROWS * urow;
  std::vector<int> imp_slot(np, -1), imp_stamp(np, -1);
  for (int k = 0; k < F.ns; k++) {
    for (int m = F.strip_ms[k]; m < F.strip_ms[k + 1]; m++) {
      int* M = &F.meta[(size_t)m * F.meta_stride];
      int* V = &F.pk_src[(size_t)m * F.row_stride];
      const int ml = m - F.strip_ms[k];
      int nops = 0, nimp = 0, nupmax = 0, lomax = 0;
      for (int q = 0; q < TS_ROWS; q++) {
        const int row = F.ms_rows[(size_t)m * TS_ROWS + q];
        M[4 + q] = row;
        M[14 + q] = row < 0 ? 0 : rowptr[row];
        M[24 + q] = row < 0 ? 0 : dia[row] - rowptr[row];
        M[34 + q] = row < 0 ? 0 : rowptr[row + 1] - dia[row];
        if (row < 0) continue;
        nops = std::max(nops, op_ptr[row + 1] - op_ptr[row]);
        nupmax = std::max(nupmax, M[34 + q]);
        lomax = std::max(lomax, M[24 + q]);
        // packed row data: blocks in crs order, the diagonal at block index gB-1, block rows padded to 4 doubles
        const int nb = rowptr[row + 1] - rowptr[row];
        for (int b = 0; b < nb; b++)
          for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) V[(q * (gB + 1) + b) * 12 + r * 4 + c] = (rowptr[row] + b) * 9 + r * 3 + c;
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) V[(q * (gB + 1) + gB - 1) * 12 + r * 4 + c] = ncrs * 9 + row * 9 + r * 3 + c;
      }
      for (int o = 0; o < gOps; o++)
        for (int q = 0; q < TS_ROWS; q++) {
          int* E = &M[FC_META_HDR + (o * TS_ROWS + q) * 4];
          const int row = F.ms_rows[(size_t)m * TS_ROWS + q];
          const int dummy = q * gstride + gB * 96;
          E[0] = dummy; E[1] = ring_bytes; E[2] = dummy; E[3] = 0;  // padding: dummy -= dummy * 0
          if (row < 0 || o >= op_ptr[row + 1] - op_ptr[row]) continue;
          const int ik = op_ik[op_ptr[row] + o], kj = op_kj[op_ptr[row] + o], ij = op_ij[op_ptr[row] + o];
          const int kr = colidx[ik], u = kj - dia[kr];
          E[0] = q * gstride + (ik - rowptr[row]) * 96;
          E[2] = q * gstride + (ij < 0 ? gB - 1 : ij - rowptr[row]) * 96;
          const int mk = ms_of_row[kr];
          const bool same = kr / T.R == k;  // the factorisation follows the forward sweep: position == row index
          if (same && mk < m && m - mk <= FC_H - 1) {
            E[1] = (((mk - F.strip_ms[k]) % FC_H) * TS_ROWS + q_of_row[kr]) * urow + u * 96;
          } else {
            if (imp_stamp[kr] != m) {
              imp_stamp[kr] = m;
              imp_slot[kr] = nimp++;
              F.imp_row.push_back(kr);
              F.imp_src.push_back(kr / T.R);
              F.imp_need.push_back(mk - F.strip_ms[kr / T.R] + 1);
            }
            E[1] = -1 - (imp_slot[kr] * urow + u * 96);
          }
        }
      M[0] = nops; M[1] = nimp; M[2] = nupmax; M[3] = lomax;
      (void)ml;
      F.imp_ptr[m + 1] = (int)F.imp_row.size();
      F.max_imp = std::max(F.max_imp, nimp);
    }
  }
  F.max_steps = T.max_steps;
  if (F.pk_src.size() > (size_t)0x7fffffff) { set_error("factor plan exceeds 2^31 packed values"); return CLOTHGPU_ERR_CAPACITY; }
  return CLOTHGPU_OK;
}

// structural self check (host only): operands resolve to rows that are finished and still in the ring, or to imports
// of earlier strips / old enough steps; update lists follow the reference order; packed data maps to the right entries
int verify_facplan_host(int np, const int* rowptr, const int* colidx, const int* dia, const TriPlanHost& T,
                        const FacPlanHost& F) {
#define PLAN_FAIL(msg) do { set_error(std::string("factor plan: ") + msg); return CLOTHGPU_ERR_STATE; } while (0)
  const int ncrs = rowptr[np];
  const int gB = F.gmaxB, gstride = (gB + 1) * 96, urow = F.gmaxU * 96, ring_bytes = FC_H * TS_ROWS * urow;
  std::vector<int> ms_of_row(np, -1), q_of_row(np, -1);
  for (int m = 0; m < F.nms; m++)
    for (int q = 0; q < TS_ROWS; q++) {
      const int row = F.ms_rows[(size_t)m * TS_ROWS + q];
      if (row >= 0) { ms_of_row[row] = m; q_of_row[row] = q; }
    }
  std::vector<int> slot(np, -1);
  for (int k = 0; k < F.ns; k++)
    for (int m = F.strip_ms[k]; m < F.strip_ms[k + 1]; m++) {
      const int* M = &F.meta[(size_t)m * F.meta_stride];
      const int* V = &F.pk_src[(size_t)m * F.row_stride];
      const int i0 = F.imp_ptr[m], nimp = F.imp_ptr[m + 1] - i0;
      if (M[1] != nimp || nimp > F.max_imp) PLAN_FAIL("import count");
      for (int e = 0; e < nimp; e++) {
        const int kr = F.imp_row[i0 + e], src = F.imp_src[i0 + e];
        if (src != kr / T.R || src > k || F.imp_need[i0 + e] != ms_of_row[kr] - F.strip_ms[src] + 1) PLAN_FAIL("import entry");
        if (src == k && ms_of_row[kr] >= m) PLAN_FAIL("import of a row that is not finished");
      }
      for (int q = 0; q < TS_ROWS; q++) {
        const int row = F.ms_rows[(size_t)m * TS_ROWS + q];
        if (M[4 + q] != row) PLAN_FAIL("row id");
        if (row < 0) {
          for (int o = 0; o < F.gmaxOps; o++) {
            const int* E = &M[FC_META_HDR + (o * TS_ROWS + q) * 4];
            if (E[0] != q * gstride + gB * 96 || E[1] != ring_bytes || E[2] != E[0]) PLAN_FAIL("padding op");
          }
          continue;
        }
        if (M[14 + q] != rowptr[row] || M[24 + q] != dia[row] - rowptr[row] || M[34 + q] != rowptr[row + 1] - dia[row]) PLAN_FAIL("row header");
        for (int b = 0; b < rowptr[row + 1] - rowptr[row]; b++)
          for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++)
              if (V[(q * (gB + 1) + b) * 12 + r * 4 + c] != (c < 3 ? (rowptr[row] + b) * 9 + r * 3 + c : -1)) PLAN_FAIL("packed block");
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++)
            if (V[(q * (gB + 1) + gB - 1) * 12 + r * 4 + c] != ncrs * 9 + row * 9 + r * 3 + c) PLAN_FAIL("packed diagonal");
        // replay the reference's update order
        for (int j = rowptr[row]; j < rowptr[row + 1]; j++) slot[colidx[j]] = j;
        int o = 0;
        for (int ik = rowptr[row]; ik < dia[row]; ik++) {
          const int kr = colidx[ik];
          for (int kj = dia[kr]; kj < rowptr[kr + 1]; kj++) {
            const int jc = colidx[kj];
            int ij;
            if (jc == row) ij = -1;
            else if (slot[jc] >= 0) ij = slot[jc];
            else continue;
            if (o >= M[0]) PLAN_FAIL("update list too short");
            const int* E = &M[FC_META_HDR + (o * TS_ROWS + q) * 4];
            if (E[0] != q * gstride + (ik - rowptr[row]) * 96) PLAN_FAIL("op ik");
            if (E[2] != q * gstride + (ij < 0 ? gB - 1 : ij - rowptr[row]) * 96) PLAN_FAIL("op ij");
            const int u = kj - dia[kr], mk = ms_of_row[kr];
            if (E[1] >= 0) {
              if (kr / T.R != k || mk >= m || m - mk > FC_H - 1) PLAN_FAIL("ring operand out of reach");
              if (E[1] != (((mk - F.strip_ms[k]) % FC_H) * TS_ROWS + q_of_row[kr]) * urow + u * 96) PLAN_FAIL("ring operand offset");
            } else {
              const int off = -1 - E[1], e = off / urow;
              if (e >= nimp || F.imp_row[i0 + e] != kr || off % urow != u * 96) PLAN_FAIL("import operand");
            }
            o++;
          }
        }
        for (int j = rowptr[row]; j < rowptr[row + 1]; j++) slot[colidx[j]] = -1;
        for (; o < F.gmaxOps; o++) {
          const int* E = &M[FC_META_HDR + (o * TS_ROWS + q) * 4];
          if (E[0] != q * gstride + gB * 96 || E[1] != ring_bytes || E[2] != E[0]) PLAN_FAIL("padding op after the list");
        }
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

int upload_facplan(Ctx* c, const FacPlanHost& H, FacPlan& D) {
  D.ns = H.ns; D.R = H.R; D.nms = H.nms; D.gmaxB = H.gmaxB; D.gmaxU = H.gmaxU; D.gmaxOps = H.gmaxOps;
  D.max_imp = H.max_imp; D.meta_stride = H.meta_stride; D.row_stride = H.row_stride; D.max_steps = H.max_steps;
  D.nval = (long long)H.pk_src.size();
  CG_CHECK(up(&D.strip_ms, H.strip_ms)); CG_CHECK(up(&D.meta, H.meta)); CG_CHECK(up(&D.pk_src, H.pk_src));
  CG_CHECK(up(&D.imp_ptr, H.imp_ptr)); CG_CHECK(up(&D.imp_row, H.imp_row)); CG_CHECK(up(&D.imp_src, H.imp_src));
  CG_CHECK(up(&D.imp_need, H.imp_need));
  CG_CUDA(cudaMalloc((void**)&D.packed, sizeof(double) * (size_t)std::max<long long>(D.nval, 1) * c->B));
  CG_CUDA(cudaMalloc((void**)&D.uexp, sizeof(double) * (size_t)c->np * D.gmaxU * 12 * c->B));
  CG_CUDA(cudaMemset(D.uexp, 0, sizeof(double) * (size_t)c->np * D.gmaxU * 12 * c->B));
  CG_CUDA(cudaMalloc((void**)&D.progress, sizeof(unsigned long long) * (size_t)H.ns * c->B));
  CG_CUDA(cudaMemset(D.progress, 0, sizeof(unsigned long long) * (size_t)H.ns * c->B));
  CG_CUDA(cudaMalloc((void**)&D.ticket, sizeof(unsigned long long)));
  CG_CUDA(cudaMemset(D.ticket, 0, sizeof(unsigned long long)));
  const int urow = D.gmaxU * 96;
  D.smem_bytes = FC_SMEM_HDR + FC_H * TS_ROWS * urow + 96 + FC_NIB * std::max(D.max_imp, 1) * urow +
                 FC_NST * (D.meta_stride * 4 + D.row_stride * 8);
  if (D.smem_bytes > 227 * 1024) { set_error("factor plan does not fit shared memory"); return CLOTHGPU_ERR_CAPACITY; }
  return CLOTHGPU_OK;
}

void free_facplan(FacPlan& D) {
  void* ptrs[] = {D.strip_ms, D.meta, D.pk_src, D.imp_ptr, D.imp_row, D.imp_src, D.imp_need, D.packed, D.uexp,
                  D.progress, D.ticket};
  for (void* p : ptrs) cudaFree(p);
  D = FacPlan();
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t f_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void f_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(f_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void f_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(f_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(f_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void f_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t a = f_smem_u32(bar);
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
__device__ __forceinline__ void f_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(f_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(f_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t f_ld_acquire_cta(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(f_smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void f_st_release_cta(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(f_smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long f_ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void f_st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct FacArgs {
  int np, ncrs, ns, B;
  int gmaxB, gmaxU, meta_stride, row_stride, max_imp;
  long long nval;
  const int *strip_ms, *meta, *imp_ptr, *imp_row, *imp_src, *imp_need;
  const double* packed;
  double *uexp, *iluDia, *iluCrs;
  unsigned long long *progress, *ticket;
  unsigned long long pbase, tbase;
  int* cnt;
  unsigned long long* dbg;  // optional: 8 counters per CTA (clothgpu_debug_trisolve, slot of the forward sweep)
};

// one CTA per (strip, scene): warp 0 computes, warp 1 lane 0 streams row data, warp 2 publishes, warp 3 lane 0 imports
__global__ void __launch_bounds__(128, 1) k_ilu_factor_strips(const FacArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);        // FC_NST
  uint64_t* empty = full + FC_NST;                            // FC_NST
  uint64_t* ifull = empty + FC_NST;                           // FC_NIB
  uint32_t* flags = reinterpret_cast<uint32_t*>(smem + 128);  // [0] micro-steps done, [2] ticket
  const int urow = A.gmaxU * 96;
  const int ring_bytes = FC_H * TS_ROWS * urow;
  unsigned char* ring = smem + FC_SMEM_HDR;          // + zero block right behind it
  unsigned char* imp = ring + ring_bytes + 96;
  const int imp_bytes = max(A.max_imp, 1) * urow;
  unsigned char* stages = imp + (size_t)FC_NIB * imp_bytes;
  const int meta_bytes = A.meta_stride * 4, row_bytes = A.row_stride * 8, stage_bytes = meta_bytes + row_bytes;
  const int gstride = (A.gmaxB + 1) * 96;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    const unsigned long long t = atomicAdd(A.ticket, 1ULL) - A.tbase;
    flags[2] = (uint32_t)t;
  }
  __syncthreads();
  const uint32_t ticket = flags[2];
  const int strip = (int)(ticket / (uint32_t)A.B), scene = (int)(ticket % (uint32_t)A.B);
  const int ms0 = A.strip_ms[strip], nsteps = A.strip_ms[strip + 1] - ms0;
  double* dia = A.iluDia + (size_t)scene * A.np * 9;
  double* crs = A.iluCrs + (size_t)scene * A.ncrs * 9;
  double* uexp = A.uexp + (size_t)scene * A.np * A.gmaxU * 12;
  if (tid < 12) reinterpret_cast<double*>(ring + ring_bytes)[tid] = 0.0;  // zero operand of the padding updates
  if (tid == 0) {
    for (int i = 0; i < FC_NST; i++) { f_mbar_init(&full[i], 1); f_mbar_init(&empty[i], 1); }
    for (int i = 0; i < FC_NIB; i++) f_mbar_init(&ifull[i], 1);
    flags[0] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    // ---------------------------------------------------------------- compute warp
    const int g = lane / 3, r = lane - g * 3;
    const bool act = g < TS_ROWS;
    const int gq = act ? g : TS_ROWS - 1;
    const bool dbg = A.dbg != nullptr;
    unsigned long long t_begin = 0, w_full = 0, w_imp = 0, w_ops = 0, w_tail = 0;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (dbg) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); t_begin = t; }
    for (int m = 0; m < nsteps; m++) {
      const int st = m % FC_NST;
      if (dbg) c0 = clock64();
      f_mbar_wait(&full[st], (uint32_t)(m / FC_NST) & 1u);
      if (dbg) c1 = clock64();
      f_mbar_wait(&ifull[m % FC_NIB], (uint32_t)(m / FC_NIB) & 1u);
      if (dbg) c2 = clock64();
      const int* M = reinterpret_cast<const int*>(stages + (size_t)st * stage_bytes);
      unsigned char* sc = stages + (size_t)st * stage_bytes + meta_bytes;  // this step's row blocks
      const unsigned char* ib = imp + (size_t)(m % FC_NIB) * imp_bytes;
      const int nops = M[0], nupmax = M[2], lomax = M[3];
      const int row = act ? M[4 + gq] : -1;
      // ---- updates A_ij -= A_ik U'_kj in the reference's order; lane r owns block-row r
      int4 e_next = *reinterpret_cast<const int4*>(M + FC_META_HDR + gq * 4);
      for (int o = 0; o < nops; o++) {
        const int4 e = e_next;
        // the next update's addresses do not depend on this update's stores: fetch them first
        e_next = *reinterpret_cast<const int4*>(M + FC_META_HDR + ((o + 1 < nops ? o + 1 : o) * TS_ROWS + gq) * 4);
        const unsigned char* pk = sc + e.x + r * 32;
        const unsigned char* pb = e.y >= 0 ? ring + e.y : ib + (-1 - e.y);
        unsigned char* pt = sc + e.z + r * 32;
        const double2 a01 = *reinterpret_cast<const double2*>(pk);
        const double a2 = *reinterpret_cast<const double*>(pk + 16);
        const double2 b00 = *reinterpret_cast<const double2*>(pb), b10 = *reinterpret_cast<const double2*>(pb + 32),
                      b20 = *reinterpret_cast<const double2*>(pb + 64);
        const double b02 = *reinterpret_cast<const double*>(pb + 16), b12 = *reinterpret_cast<const double*>(pb + 48),
                     b22 = *reinterpret_cast<const double*>(pb + 80);
        double2 t01 = *reinterpret_cast<const double2*>(pt);
        double t2 = *reinterpret_cast<const double*>(pt + 16);
        t01.x = t01.x - (a01.x * b00.x + a01.y * b10.x + a2 * b20.x);
        t01.y = t01.y - (a01.x * b00.y + a01.y * b10.y + a2 * b20.y);
        t2 = t2 - (a01.x * b02 + a01.y * b12 + a2 * b22);
        if (act) {
          *reinterpret_cast<double2*>(pt) = t01;
          *reinterpret_cast<double*>(pt + 16) = t2;
        }
      }
      __syncwarp();
      if (dbg) c3 = clock64();
      // ---- D_i <- D_i^-1 by cofactors (CalcInvMat3, ilu_sparse.cpp:77-97); a singular pivot is left as is (:541-551)
      double a[9];
      {
        const unsigned char* pd = sc + gq * gstride + (A.gmaxB - 1) * 96;
#pragma unroll
        for (int rr = 0; rr < 3; rr++) {
          const double2 v01 = *reinterpret_cast<const double2*>(pd + rr * 32);
          a[rr * 3 + 0] = v01.x; a[rr * 3 + 1] = v01.y;
          a[rr * 3 + 2] = *reinterpret_cast<const double*>(pd + rr * 32 + 16);
        }
        const double det = +a[0] * a[4] * a[8] + a[3] * a[7] * a[2] + a[6] * a[1] * a[5] - a[0] * a[7] * a[5] -
                           a[6] * a[4] * a[2] - a[3] * a[1] * a[8];
        if (fabs(det) > 1.0e-30) {
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
        } else if (row >= 0 && r == 0) {
          atomicAdd(&A.cnt[scene * NCNT + CNT_NSING], 1);
        }
      }
      const double ar0 = r == 0 ? a[0] : (r == 1 ? a[3] : a[6]);
      const double ar1 = r == 0 ? a[1] : (r == 1 ? a[4] : a[7]);
      const double ar2 = r == 0 ? a[2] : (r == 1 ? a[5] : a[8]);
      const int crs0 = act ? M[14 + gq] : 0, nlo = act ? M[24 + gq] : 0, nup = act ? M[34 + gq] : 0;
      if (row >= 0) {
        double* o = dia + (size_t)row * 9 + r * 3;
        o[0] = ar0; o[1] = ar1; o[2] = ar2;
      }
      // ---- U'_ij = D_i^-1 U_ij (:556-567): to the ring (next steps of this strip), the ILU array and the export copy
      unsigned char* myring = ring + (size_t)((m % FC_H) * TS_ROWS + gq) * urow + r * 32;
      for (int u = 0; u < nupmax; u++) {
        const unsigned char* pu = sc + gq * gstride + (nlo + (u < nup ? u : 0)) * 96;
        const double2 c0 = *reinterpret_cast<const double2*>(pu), c1 = *reinterpret_cast<const double2*>(pu + 32),
                      c2 = *reinterpret_cast<const double2*>(pu + 64);
        const double c02 = *reinterpret_cast<const double*>(pu + 16), c12 = *reinterpret_cast<const double*>(pu + 48),
                     c22 = *reinterpret_cast<const double*>(pu + 80);
        const double v0 = ar0 * c0.x + ar1 * c1.x + ar2 * c2.x;
        const double v1 = ar0 * c0.y + ar1 * c1.y + ar2 * c2.y;
        const double v2 = ar0 * c02 + ar1 * c12 + ar2 * c22;
        if (row >= 0 && u < nup) {
          *reinterpret_cast<double2*>(myring + u * 96) = make_double2(v0, v1);
          *reinterpret_cast<double*>(myring + u * 96 + 16) = v2;
          double* oc = crs + (size_t)(crs0 + nlo + u) * 9 + r * 3;
          oc[0] = v0; oc[1] = v1; oc[2] = v2;
          double* ou = uexp + ((size_t)row * A.gmaxU + u) * 12 + r * 4;
          *reinterpret_cast<double2*>(ou) = make_double2(v0, v1);
          ou[2] = v2;
        }
      }
      // ---- the finished lower part L_ik
      for (int b = 0; b < lomax; b++) {
        if (row >= 0 && b < nlo) {
          const unsigned char* pl = sc + gq * gstride + b * 96 + r * 32;
          const double2 l01 = *reinterpret_cast<const double2*>(pl);
          const double l2 = *reinterpret_cast<const double*>(pl + 16);
          double* oc = crs + (size_t)(crs0 + b) * 9 + r * 3;
          oc[0] = l01.x; oc[1] = l01.y; oc[2] = l2;
        }
      }
      __syncwarp();
      if (lane == 0) {
        f_st_release_cta(&flags[0], (uint32_t)(m + 1));
        f_mbar_arrive(&empty[st]);
      }
      if (dbg) {
        w_full += (unsigned long long)(c1 - c0); w_imp += (unsigned long long)(c2 - c1);
        w_ops += (unsigned long long)(c3 - c2); w_tail += (unsigned long long)(clock64() - c3);
      }
    }
    if (dbg && lane == 0) {
      unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      unsigned long long* o = A.dbg + (size_t)blockIdx.x * 8;
      o[0] = t_begin; o[1] = t; o[2] = w_full; o[3] = w_imp; o[4] = w_ops; o[5] = w_tail;
      o[6] = (unsigned long long)nsteps; o[7] = (unsigned long long)strip;
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- row-data loader: one TMA pair per micro-step
    if (lane == 0) {
      const unsigned char* meta = reinterpret_cast<const unsigned char*>(A.meta) + (size_t)ms0 * meta_bytes;
      const unsigned char* packed = reinterpret_cast<const unsigned char*>(A.packed + (size_t)scene * A.nval) + (size_t)ms0 * row_bytes;
      for (int m = 0; m < nsteps; m++) {
        const int st = m % FC_NST;
        if (m >= FC_NST) f_mbar_wait(&empty[st], (uint32_t)(m / FC_NST - 1) & 1u);
        unsigned char* dst = stages + (size_t)st * stage_bytes;
        f_mbar_expect_tx(&full[st], (uint32_t)stage_bytes);
        f_bulk_g2s(dst, meta + (size_t)m * meta_bytes, (uint32_t)meta_bytes, &full[st]);
        f_bulk_g2s(dst + meta_bytes, packed + (size_t)m * row_bytes, (uint32_t)row_bytes, &full[st]);
      }
    }
  } else if (warp == 2) {
    // ---------------------------------------------------------------- publisher: the compute warp stores its results itself
    // (fire and forget); this warp turns "step done" into a device-wide release of the strip's progress counter
    if (lane == 0) {
      unsigned long long* prog = A.progress + (size_t)scene * A.ns + strip;
      uint32_t pub = 0;
      while (pub < (uint32_t)nsteps) {
        uint32_t d;
        while ((d = f_ld_acquire_cta(&flags[0])) <= pub) __nanosleep(20);
        __threadfence();
        f_st_release_gpu(prog, A.pbase + d);
        pub = d;
      }
    }
  } else {
    // ---------------------------------------------------------------- import lane: operands that are not in the ring
    if (lane == 0) {
      const unsigned long long* prog = A.progress + (size_t)scene * A.ns;
      int cached_src = -1;
      unsigned long long cached = 0;
      for (int m = 0; m < nsteps; m++) {
        const int ibuf = m % FC_NIB;
        // the buffer was last used by step m - FC_NIB
        if (m >= FC_NIB) {
          while (f_ld_acquire_cta(&flags[0]) < (uint32_t)(m - FC_NIB + 1)) __nanosleep(20);
        }
        const int i0 = A.imp_ptr[ms0 + m], n = A.imp_ptr[ms0 + m + 1] - i0;
        if (n == 0) {
          f_mbar_arrive(&ifull[ibuf]);
          continue;
        }
        f_mbar_expect_tx(&ifull[ibuf], (uint32_t)(n * urow));
        for (int e = 0; e < n; e++) {
          const int kr = A.imp_row[i0 + e], src = A.imp_src[i0 + e];
          const unsigned long long want = A.pbase + (unsigned long long)A.imp_need[i0 + e];
          if (src != cached_src) { cached_src = src; cached = 0; }
          while (cached < want) cached = f_ld_acquire_gpu(prog + src);
          // the producer's stores went through the generic proxy; order them before the async-proxy read
          asm volatile("fence.proxy.async.global;" ::: "memory");
          f_bulk_g2s(imp + (size_t)ibuf * imp_bytes + (size_t)e * urow, uexp + (size_t)kr * A.gmaxU * 12, (uint32_t)urow, &ifull[ibuf]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ launcher
__global__ void __launch_bounds__(256) k_pack_rows(long long nval, const int* __restrict__ pk_src, int ncrs9, int np9,
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

int launch_ilu_factor_strips(Ctx* c) {
  FacPlan& P = c->plan_fac;
  {
    dim3 g(std::min<long long>((P.nval + 255) / 256, 148 * 16), c->B);
    k_pack_rows<<<g, 256, 0, KS(c, "k_pack_rows")>>>(P.nval, P.pk_src, c->ncrs * 9, c->np * 9, c->iluDia, c->iluCrs, P.packed);
    CG_LAUNCH_CHECK(c);
  }
  FacArgs A;
  A.np = c->np; A.ncrs = c->ncrs; A.ns = P.ns; A.B = c->B;
  A.gmaxB = P.gmaxB; A.gmaxU = P.gmaxU; A.meta_stride = P.meta_stride; A.row_stride = P.row_stride; A.max_imp = P.max_imp;
  A.nval = P.nval;
  A.strip_ms = P.strip_ms; A.meta = P.meta; A.imp_ptr = P.imp_ptr; A.imp_row = P.imp_row; A.imp_src = P.imp_src;
  A.imp_need = P.imp_need; A.packed = P.packed; A.uexp = P.uexp; A.iluDia = c->iluDia; A.iluCrs = c->iluCrs;
  A.progress = P.progress; A.ticket = P.ticket; A.pbase = P.pbase; A.tbase = P.tbase; A.cnt = c->cnt;
  const int grid = P.ns * c->B;
  A.dbg = nullptr;
  if (c->tri_debug) {
    if (!c->tri_dbg_buf) CG_CUDA(cudaMalloc((void**)&c->tri_dbg_buf, sizeof(unsigned long long) * 8 * 2 * 4096));
    if (grid <= 4096) A.dbg = c->tri_dbg_buf;
    c->tri_dbg_grid[0] = grid;
  }
  CG_CUDA(cudaFuncSetAttribute(k_ilu_factor_strips, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes));
  k_ilu_factor_strips<<<grid, 128, P.smem_bytes, KS(c, "k_ilu_factor_strips")>>>(A);
  CG_LAUNCH_CHECK(c);
  P.pbase += (unsigned long long)P.max_steps + 2;
  P.tbase += (unsigned long long)grid;
  return CLOTHGPU_OK;
}

}  // namespace cg
