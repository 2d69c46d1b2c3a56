// csrc/ilufactor.cu -- block ILU(0) factorisation as a strip-pipelined wavefront.
//
// Replaces CPreconditionerILU::DoILUDecomp (ilu_sparse.cpp:387-624, len == 3 at :507-569; CalcInvMat3 :77-97).
//
// Row i of the factorisation needs the finished upper parts U'_k = D_k^-1 U_k of its lower neighbours k, so the
// rows form the same dependency wavefront as the forward substitution (1021 levels on a 256^2 grid).  The kernel
// uses the forward sweep's strip / micro-step scheme (trisolve.cu; a single scene takes narrower strips): one CTA per
// strip, <= 10 rows per micro-step, nine lanes per row -- lane (r, c) owns entry (r, c) of every 3x3 block of its
// matrix row, so the update A_ij -= A_ik U'_kj needs no cross-lane traffic.  Per micro-step
//   * the row's blocks arrive by one TMA bulk copy per stage, packed by k_pack_rows into a lane-major layout
//     ([block][96 lane slots], 8-byte lane stride): whatever block each row group addresses, the 16 lanes of a
//     half-warp hit 16 distinct bank pairs, so the dynamic per-row block indices cost no bank conflicts;
//   * the update list (ik, kj, ij) in the reference's order is applied in shared memory, one lower block k at a time;
//   * operands U'_k produced by this strip within the last FC_H micro-steps are read from a shared-memory ring;
//     older ones and those of other strips are fetched by an import lane with TMA bulk copies from a padded
//     global copy of the upper parts ("uexp"), gated by the producer strip's progress counter;
//   * D_i is inverted by cofactors (every lane its own entry), U'_i = D_i^-1 U_i goes to the ring;
//   * drain warps copy the finished micro-step (L, D^-1, U') from shared memory to the ILU arrays and to uexp and
//     publish the strip's progress: the compute warps never touch global memory.
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
  const int blk = 768;                 // bytes of one block of the scratch: [96 lane slots] doubles
  const int urow = round_up(F.gmaxU * 72, 16);  // bytes of one exported upper part (dense 3x3 blocks, TMA granule)
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
  // the updates of one row grouped by their lower block k ("k-group"): all updates of a group read the same A_ik
  // and hit distinct targets, so they are mutually independent and can be issued as one bundle
  int gOps = 0, gLo = 0, gT = 0;
  for (int i = 0; i < np; i++) {
    gOps = std::max(gOps, op_ptr[i + 1] - op_ptr[i]);
    gLo = std::max(gLo, dia[i] - rowptr[i]);
    int run = 0;
    for (int o = op_ptr[i]; o < op_ptr[i + 1]; o++) {
      run = (o > op_ptr[i] && op_ik[o] == op_ik[o - 1]) ? run + 1 : 1;
      gT = std::max(gT, run);
    }
  }
  gT = round_up(std::max(gT, 1), FC_BUNDLE);
  F.gmaxOps = gOps; F.gmaxLo = gLo; F.gT = gT;
  F.kg_stride = 12 + gT * TS_ROWS * 2;  // ints per k-group: A_ik offsets (10), bundle count, pad, then (kj, ij) pairs
  F.meta_stride = round_up(FC_META_HDR + gLo * F.kg_stride, 4);
  F.row_stride = (gB + 1) * 96;  // doubles per micro-step of packed row data: gB blocks + one dummy block
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
  // ring of finished upper parts: [step mod FC_H][upper block][16 row slots][9 entries] doubles, then a zero block
  const int ring_bytes = FC_H * F.gmaxU * 9 * 128;
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
        nupmax = std::max(nupmax, M[34 + q]);
        lomax = std::max(lomax, M[24 + q]);
        // packed row data: blocks in crs order, the diagonal at block index gB-1; entry (r, c) of block b of row
        // group q sits at [b*96 + q*9 + r*3 + c] (one lane per entry, 9 lanes per row)
        const int nb = rowptr[row + 1] - rowptr[row];
        for (int b = 0; b < nb; b++)
          for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) V[b * 96 + q * 9 + r * 3 + c] = (rowptr[row] + b) * 9 + r * 3 + c;
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) V[(gB - 1) * 96 + q * 9 + r * 3 + c] = ncrs * 9 + row * 9 + r * 3 + c;
      }
      for (int kg = 0; kg < gLo; kg++) {
        int* K = &M[FC_META_HDR + kg * F.kg_stride];
        int ntk = 0;
        for (int q = 0; q < TS_ROWS; q++) {
          const int row = F.ms_rows[(size_t)m * TS_ROWS + q];
          const int dummy = gB * blk;
          K[q] = dummy;
          for (int t = 0; t < gT; t++) { K[12 + (t * TS_ROWS + q) * 2] = ring_bytes; K[12 + (t * TS_ROWS + q) * 2 + 1] = dummy; }
          if (row < 0 || kg >= dia[row] - rowptr[row]) continue;
          const int ik = rowptr[row] + kg, kr = colidx[ik];
          K[q] = kg * blk;
          int t = 0;
          for (int o = op_ptr[row]; o < op_ptr[row + 1]; o++) {
            if (op_ik[o] != ik) continue;
            const int kj = op_kj[o], ij = op_ij[o], u = kj - dia[kr];
            int* E = &K[12 + (t * TS_ROWS + q) * 2];
            E[1] = (ij < 0 ? gB - 1 : ij - rowptr[row]) * blk;
            const int mk = ms_of_row[kr];
            const bool same = kr / T.R == k;  // the factorisation follows the forward sweep: position == row index
            if (same && mk < m && m - mk <= FC_H - 1) {
              E[0] = (((((mk - F.strip_ms[k]) % FC_H) * F.gmaxU + u) * 16 + q_of_row[kr]) * 9) * 8;
            } else {
              if (imp_stamp[kr] != m) {
                imp_stamp[kr] = m;
                imp_slot[kr] = nimp++;
                F.imp_row.push_back(kr);
                F.imp_src.push_back(kr / T.R);
                F.imp_need.push_back(mk - F.strip_ms[kr / T.R] + 1);
              }
              E[0] = -1 - (imp_slot[kr] * urow + u * 72);
            }
            t++;
          }
          ntk = std::max(ntk, t);
        }
        K[10] = round_up(ntk, FC_BUNDLE);
      }
      nops = lomax;  // number of k-groups this micro-step runs
      M[0] = nops; M[1] = nimp; M[2] = nupmax; M[3] = lomax;
      (void)ml;
      F.imp_ptr[m + 1] = (int)F.imp_row.size();
      F.max_imp = std::max(F.max_imp, nimp);
    }
  }
  // operand offsets become absolute byte offsets into the CTA's shared memory (the layout is a function of the plan):
  // ring operands sit at FC_SMEM_HDR + offset, imported ones in import buffer (local step % FC_NIB)
  {
    const int imp_bytes = round_up(std::max(F.max_imp, 1) * urow, 128);
    F.imp_bytes = imp_bytes;
    for (int k = 0; k < F.ns; k++)
      for (int m = F.strip_ms[k]; m < F.strip_ms[k + 1]; m++) {
        int* M = &F.meta[(size_t)m * F.meta_stride];
        const int ibuf = (m - F.strip_ms[k]) % FC_NIB;
        for (int kg = 0; kg < gLo; kg++) {
          int* K = &M[FC_META_HDR + kg * F.kg_stride];
          for (int t = 0; t < gT; t++)
            for (int q = 0; q < TS_ROWS; q++) {
              int& e = K[12 + (t * TS_ROWS + q) * 2];
              e = e >= 0 ? FC_SMEM_HDR + e : FC_SMEM_HDR + ring_bytes + 128 + ibuf * imp_bytes + (-1 - e);
            }
        }
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
  const int gB = F.gmaxB, blk = 768, urow = round_up(F.gmaxU * 72, 16), ring_bytes = FC_H * F.gmaxU * 9 * 128;
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
          for (int kg = 0; kg < F.gmaxLo; kg++) {
            const int* K = &M[FC_META_HDR + kg * F.kg_stride];
            if (K[q] != gB * blk) PLAN_FAIL("padding k-group of an unused lane");
            for (int t = 0; t < F.gT; t++)
              if (K[12 + (t * TS_ROWS + q) * 2] != FC_SMEM_HDR + ring_bytes || K[12 + (t * TS_ROWS + q) * 2 + 1] != K[q]) PLAN_FAIL("padding update of an unused lane");
          }
          continue;
        }
        if (M[14 + q] != rowptr[row] || M[24 + q] != dia[row] - rowptr[row] || M[34 + q] != rowptr[row + 1] - dia[row]) PLAN_FAIL("row header");
        for (int b = 0; b < rowptr[row + 1] - rowptr[row]; b++)
          for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
              if (V[b * 96 + q * 9 + r * 3 + c] != (rowptr[row] + b) * 9 + r * 3 + c) PLAN_FAIL("packed block");
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) {
            if (V[(gB - 1) * 96 + q * 9 + r * 3 + c] != ncrs * 9 + row * 9 + r * 3 + c) PLAN_FAIL("packed diagonal");
            if (V[gB * 96 + q * 9 + r * 3 + c] != -1) PLAN_FAIL("dummy block must be empty");
          }
        // replay the reference's update order: k ascending, within k the upper blocks of row k ascending
        for (int j = rowptr[row]; j < rowptr[row + 1]; j++) slot[colidx[j]] = j;
        for (int ik = rowptr[row]; ik < dia[row]; ik++) {
          const int kr = colidx[ik], kg = ik - rowptr[row];
          if (kg >= M[0]) PLAN_FAIL("k-group count");
          const int* K = &M[FC_META_HDR + kg * F.kg_stride];
          if (K[q] != kg * blk) PLAN_FAIL("A_ik offset");
          int t = 0;
          for (int kj = dia[kr]; kj < rowptr[kr + 1]; kj++) {
            const int jc = colidx[kj];
            int ij;
            if (jc == row) ij = -1;
            else if (slot[jc] >= 0) ij = slot[jc];
            else continue;
            if (t >= K[10] || t >= F.gT) PLAN_FAIL("bundle too short");
            const int* E = &K[12 + (t * TS_ROWS + q) * 2];
            if (E[1] != (ij < 0 ? gB - 1 : ij - rowptr[row]) * blk) PLAN_FAIL("op ij");
            const int u = kj - dia[kr], mk = ms_of_row[kr];
            const int imp_base = FC_SMEM_HDR + ring_bytes + 128;
            if (E[0] < imp_base) {
              if (kr / T.R != k || mk >= m || m - mk > FC_H - 1) PLAN_FAIL("ring operand out of reach");
              if (E[0] - FC_SMEM_HDR != (((((mk - F.strip_ms[k]) % FC_H) * F.gmaxU + u) * 16 + q_of_row[kr]) * 9) * 8) PLAN_FAIL("ring operand offset");
            } else {
              const int off = E[0] - imp_base - ((m - F.strip_ms[k]) % FC_NIB) * F.imp_bytes, e = off / urow;
              if (off < 0 || e >= nimp || F.imp_row[i0 + e] != kr || off % urow != u * 72) PLAN_FAIL("import operand");
            }
            t++;
          }
          for (; t < F.gT; t++) {
            const int* E = &K[12 + (t * TS_ROWS + q) * 2];
            if (E[0] != FC_SMEM_HDR + ring_bytes || E[1] != gB * blk) PLAN_FAIL("padding update");
          }
        }
        for (int j = rowptr[row]; j < rowptr[row + 1]; j++) slot[colidx[j]] = -1;
        for (int kg = dia[row] - rowptr[row]; kg < F.gmaxLo; kg++) {
          const int* K = &M[FC_META_HDR + kg * F.kg_stride];
          if (K[q] != gB * blk) PLAN_FAIL("padding k-group");
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
  {
    const int urow0 = round_up(std::max(H.gmaxU, 1) * 72, 16);
    const long long need = (long long)FC_SMEM_HDR + (long long)FC_H * std::max(H.gmaxU, 1) * 9 * 128 + 128 +
                           (long long)FC_NIB * round_up(std::max(H.max_imp, 1) * urow0, 128) +
                           (long long)FC_NST * ((long long)H.meta_stride * 4 + (long long)H.row_stride * 8) + FC_NDI * 768;
    if (need > 227 * 1024) { set_error("factor plan does not fit shared memory"); return CLOTHGPU_ERR_CAPACITY; }
  }
  D.ns = H.ns; D.R = H.R; D.nms = H.nms; D.gmaxB = H.gmaxB; D.gmaxU = H.gmaxU; D.gmaxOps = H.gmaxOps;
  D.max_imp = H.max_imp; D.meta_stride = H.meta_stride; D.row_stride = H.row_stride; D.max_steps = H.max_steps;
  D.kg_stride = H.kg_stride; D.gT = H.gT;
  D.nval = (long long)H.pk_src.size();
  CG_CHECK(up(&D.strip_ms, H.strip_ms)); CG_CHECK(up(&D.meta, H.meta)); CG_CHECK(up(&D.pk_src, H.pk_src));
  CG_CHECK(up(&D.imp_ptr, H.imp_ptr)); CG_CHECK(up(&D.imp_row, H.imp_row)); CG_CHECK(up(&D.imp_src, H.imp_src));
  CG_CHECK(up(&D.imp_need, H.imp_need));
  CG_CUDA(cudaMalloc((void**)&D.packed, sizeof(double) * (size_t)std::max<long long>(D.nval, 1) * c->B));
  const size_t uexp_bytes = (size_t)c->np * round_up(D.gmaxU * 72, 16) * c->B;
  CG_CUDA(cudaMalloc((void**)&D.uexp, uexp_bytes));
  CG_CUDA(cudaMemset(D.uexp, 0, uexp_bytes));
  CG_CUDA(cudaMalloc((void**)&D.progress, sizeof(unsigned long long) * (size_t)H.ns * c->B));
  CG_CUDA(cudaMemset(D.progress, 0, sizeof(unsigned long long) * (size_t)H.ns * c->B));
  CG_CUDA(cudaMalloc((void**)&D.ticket, sizeof(unsigned long long)));
  CG_CUDA(cudaMemset(D.ticket, 0, sizeof(unsigned long long)));
  const int urow = round_up(D.gmaxU * 72, 16);
  D.smem_bytes = FC_SMEM_HDR + FC_H * D.gmaxU * 9 * 128 + 128 + FC_NIB * round_up(std::max(D.max_imp, 1) * urow, 128) +
                 FC_NST * (D.meta_stride * 4 + D.row_stride * 8) + FC_NDI * 768;
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
  int gmaxB, gmaxU, meta_stride, row_stride, max_imp, kg_stride;
  long long nval;
  const int *strip_ms, *meta, *imp_ptr, *imp_row, *imp_src, *imp_need;
  const double* packed;
  double *uexp, *iluDia, *iluCrs;
  unsigned long long *progress, *ticket;
  unsigned long long pbase, tbase;
  int* cnt;
  unsigned long long* dbg;  // optional: 8 counters per CTA (clothgpu_debug_trisolve, slot of the forward sweep)
};

// one CTA per (strip, scene).  Compute warps: one per row triplet (FC_SPLIT = 1; nine lanes per row = one lane per
// block entry) -- with FC_SPLIT = 2, warp w and warp w + 4 share a triplet and split every bundle of independent
// updates / upper blocks.  Then one loader warp (lane 0 streams row data), one import warp (lane 0) and three drain
// warps: the compute warps leave their results in shared memory (the row blocks of the stage, the ring of upper parts,
// a D^-1 slot) and never touch global memory; one drain thread per (row, block entry) copies the finished micro-step
// to iluCrs / iluDia / uexp, then one of them publishes the strip's progress device-wide and hands the stage back to
// the loader.
template <bool DBG>
__global__ void __launch_bounds__(FC_THREADS, 1) k_ilu_factor_strips(const FacArgs A) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);        // FC_NST
  uint64_t* empty = full + FC_NST;                            // FC_NST
  uint64_t* ifull = empty + FC_NST;                           // FC_NIB
  uint32_t* flags = reinterpret_cast<uint32_t*>(smem + 128);  // [0] micro-steps computed, [2] ticket
  const int urow = (A.gmaxU * 72 + 15) / 16 * 16;     // bytes of one exported upper part (dense 3x3 blocks)
  const int ring_bytes = FC_H * A.gmaxU * 9 * 128;    // [step mod FC_H][upper block][16 row slots][9 entries] doubles
  unsigned char* ring = smem + FC_SMEM_HDR;          // + a zero block right behind it
  unsigned char* imp = ring + ring_bytes + 128;
  const int imp_bytes = (max(A.max_imp, 1) * urow + 127) / 128 * 128;
  unsigned char* stages = imp + (size_t)FC_NIB * imp_bytes;
  const int meta_bytes = A.meta_stride * 4, row_bytes = A.row_stride * 8, stage_bytes = meta_bytes + row_bytes;
  unsigned char* dinv = stages + (size_t)FC_NST * stage_bytes;  // [step mod FC_NDI][96 lane slots] doubles: D_i^-1

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
  double* uexp = A.uexp + (size_t)scene * A.np * (urow / 8);
  if (tid < 16) reinterpret_cast<double*>(ring + ring_bytes)[tid] = 0.0;  // zero operand of the padding updates
  if (tid == 0) {
    for (int i = 0; i < FC_NST; i++) { f_mbar_init(&full[i], 1); f_mbar_init(&empty[i], 1); }
    for (int i = 0; i < FC_NIB; i++) f_mbar_init(&ifull[i], 1);
    flags[0] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp < FC_CWARPS) {
    // ---------------------------------------------------------------- compute warps
    // lane -> (row group q, block entry e = r*3 + c); lanes 27..31 idle on the slots of a spare group
    const int wq = warp % FC_RWARPS, half = warp / FC_RWARPS;  // row triplet, share of every bundle
    const int gl = lane / 9, e9 = lane - gl * 9, r = e9 / 3, cc = e9 - r * 3;
    const bool act = gl < 3 && wq * 3 + gl < TS_ROWS;
    const int gq = act ? wq * 3 + gl : TS_ROWS;        // group TS_ROWS (slots 90..95 of every block) is scratch for idle lanes
    const int gm = act ? gq : TS_ROWS - 1;               // metadata column the lane reads
    const int slot8 = act ? (gq * 9 + e9) * 8 : (90 + (lane % 6)) * 8;  // byte offset of the lane's entry inside a block
    const int rowbase8 = gq * 72;                         // byte offset of the group's entries inside a block
    // cofactor of the lane's own entry of D^-1 (CalcInvMat3): inv[e] = t[i1] * t[i2] - t[i3] * t[i4], byte offsets
    int o1, o2, o3, o4;
    {
      // entries e = 0..8: (4,8,5,7) (2,7,1,8) (1,5,2,4) (5,6,3,8) (0,8,2,6) (2,3,0,5) (3,7,4,6) (1,6,0,7) (0,4,1,3)
      const unsigned long long t1 = 0x013205124ULL, t2 = 0x467386578ULL, t3 = 0x104023215ULL, t4 = 0x376568487ULL;  // nibble e9
      o1 = (int)((t1 >> (4 * e9)) & 15ULL) * 8;
      o2 = (int)((t2 >> (4 * e9)) & 15ULL) * 8;
      o3 = (int)((t3 >> (4 * e9)) & 15ULL) * 8;
      o4 = (int)((t4 >> (4 * e9)) & 15ULL) * 8;
    }
    const bool dbg = DBG && warp == 0;  // cycle counters are compiled out of the production instantiation
    unsigned long long t_begin = 0, w_full = 0, w_imp = 0, w_ops = 0, w_tail = 0, w_inv = 0, w_bar = 0;
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
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
      const int row = act ? M[4 + gm] : -1;
      // a warp whose triplet holds no row in this micro-step (narrow wavefronts, narrow strips) skips the arithmetic
      const bool wact = __any_sync(0xffffffffu, row >= 0);
      const int nops = wact ? M[0] : 0, nupmax = wact ? M[2] : 0;
      // ---- updates A_ij -= A_ik U'_kj in the reference's order.  Lane (r, c) owns entry (r, c) of every block of
      // its row.  The updates of one lower block k hit distinct targets: they are issued FC_BUNDLE at a time (loads,
      // math, stores); consecutive k-groups are ordered (a later A_ik may be an earlier target, written by the other
      // lanes of the row)
      for (int kg = 0; kg < nops; kg++) {
        const int* K = M + FC_META_HDR + kg * A.kg_stride;
        const unsigned char* pk = sc + K[gm] + (act ? rowbase8 : 720) + r * 24;  // A_ik[r][0..2]
        const int ntk = K[10];
        // idle lanes run the same code on scratch slots with zero coefficients
        const double a0 = act ? *reinterpret_cast<const double*>(pk) : 0.0, a1 = act ? *reinterpret_cast<const double*>(pk + 8) : 0.0,
                     a2 = act ? *reinterpret_cast<const double*>(pk + 16) : 0.0;
        for (int t0 = half * FC_HB; t0 < ntk; t0 += FC_BUNDLE) {
          unsigned char* pt[FC_HB];
          double b0[FC_HB], b1[FC_HB], b2[FC_HB], t[FC_HB];
#pragma unroll
          for (int j = 0; j < FC_HB; j++) {
            const int2 e = *reinterpret_cast<const int2*>(K + 12 + ((t0 + j) * TS_ROWS + gm) * 2);
            const unsigned char* pb = smem + e.x + cc * 8;  // U'_kj[0..2][c]: absolute offset, ring or import buffer
            pt[j] = sc + e.y + slot8;
            b0[j] = *reinterpret_cast<const double*>(pb);
            b1[j] = *reinterpret_cast<const double*>(pb + 24);
            b2[j] = *reinterpret_cast<const double*>(pb + 48);
            t[j] = *reinterpret_cast<const double*>(pt[j]);
          }
#pragma unroll
          for (int j = 0; j < FC_HB; j++) t[j] = t[j] - (a0 * b0[j] + a1 * b1[j] + a2 * b2[j]);
#pragma unroll
          for (int j = 0; j < FC_HB; j++) *reinterpret_cast<double*>(pt[j]) = t[j];
        }
        // with one warp per triplet the ordering is warp-local; two warps have written disjoint targets and both must
        // see all of them before the next A_ik
        if (FC_SPLIT == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(2 + wq), "n"(FC_SPLIT * 32) : "memory");
      }
      if (dbg) c3 = clock64();
      // ---- D_i <- D_i^-1 by cofactors (CalcInvMat3, ilu_sparse.cpp:77-97); a singular pivot is left as is (:541-551).
      // Every lane evaluates the determinant (all nine entries, the reference's expression) but only the cofactor of
      // its own entry; the row of D^-1 it needs for U' comes back through the D^-1 slot in shared memory, which is
      // also where the drain warps pick the block up.  (Both warps of a split triplet write the same values.)
      unsigned char* dslot = dinv + (size_t)(m % FC_NDI) * 768;
      if (wact) {
        const unsigned char* pd = sc + (A.gmaxB - 1) * 768 + (act ? rowbase8 : 720);
        double a[9];
#pragma unroll
        for (int k = 0; k < 9; k++) a[k] = *reinterpret_cast<const double*>(pd + k * 8);
        const double p1 = *reinterpret_cast<const double*>(pd + o1), p2 = *reinterpret_cast<const double*>(pd + o2);
        const double p3 = *reinterpret_cast<const double*>(pd + o3), p4 = *reinterpret_cast<const double*>(pd + o4);
        double mine = *reinterpret_cast<const double*>(pd + e9 * 8);
        const double det = +a[0] * a[4] * a[8] + a[3] * a[7] * a[2] + a[6] * a[1] * a[5] - a[0] * a[7] * a[5] -
                           a[6] * a[4] * a[2] - a[3] * a[1] * a[8];
        const double cof = p1 * p2 - p3 * p4;
        if (fabs(det) > 1.0e-30) {
          const double inv_det = 1.0 / det;
          mine = inv_det * cof;
        } else if (row >= 0 && e9 == 0 && half == 0) {
          atomicAdd(&A.cnt[scene * NCNT + CNT_NSING], 1);
        }
        *reinterpret_cast<double*>(dslot + slot8) = mine;
      }
      __syncwarp();
      const unsigned char* prow = dslot + (act ? rowbase8 : 720) + (act ? r * 24 : 0);
      const double ar0 = *reinterpret_cast<const double*>(prow), ar1 = *reinterpret_cast<const double*>(prow + 8),
                   ar2 = *reinterpret_cast<const double*>(prow + 16);
      const int nlo = act ? M[24 + gm] : 0, nup = act ? M[34 + gm] : 0;
      if (dbg) c4 = clock64();
      // ---- U'_ij = D_i^-1 U_ij (:556-567) into the ring: operand of the next steps of this strip, source of the drain
      unsigned char* myring = ring + (size_t)(((m % FC_H) * A.gmaxU * 16 + (act ? gq : 15)) * 9 + e9) * 8;
      for (int u0 = half * FC_HB; u0 < nupmax; u0 += FC_BUNDLE) {
        double v[FC_HB];
#pragma unroll
        for (int j = 0; j < FC_HB; j++) {
          const int u = u0 + j;
          const unsigned char* pu = sc + (nlo + (u < nup ? u : 0)) * 768 + (act ? rowbase8 : 720) + cc * 8;  // U_ij[0..2][c]
          v[j] = ar0 * *reinterpret_cast<const double*>(pu) + ar1 * *reinterpret_cast<const double*>(pu + 24) +
                 ar2 * *reinterpret_cast<const double*>(pu + 48);
        }
#pragma unroll
        for (int j = 0; j < FC_HB; j++) {
          const int u = u0 + j;
          if (row >= 0 && u < nup) *reinterpret_cast<double*>(myring + (size_t)u * 16 * 72) = v[j];
        }
      }
      // ---- all compute warps have finished the step: ring and D^-1 slot visible to the next steps and to the drain
      if (dbg) c5 = clock64();
      asm volatile("bar.sync 1, %0;" ::"n"(FC_CWARPS * 32) : "memory");
      if (warp == 0 && lane == 0) f_st_release_cta(&flags[0], (uint32_t)(m + 1));
      if (dbg) {
        w_full += (unsigned long long)(c1 - c0); w_imp += (unsigned long long)(c2 - c1);
        w_ops += (unsigned long long)(c3 - c2); w_inv += (unsigned long long)(c4 - c3); w_tail += (unsigned long long)(c5 - c4);
        w_bar += (unsigned long long)(clock64() - c5);
      }
    }
    if (dbg && lane == 0) {
      unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      unsigned long long* o = A.dbg + (size_t)blockIdx.x * 8;
      o[0] = t_begin; o[1] = t; o[2] = w_inv; o[3] = w_full + w_imp; o[4] = w_ops; o[5] = w_tail + (w_bar << 32);
      o[6] = (unsigned long long)nsteps; o[7] = (unsigned long long)strip;
    }
  } else if (warp == FC_CWARPS) {
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
  } else if (warp == FC_CWARPS + 1) {
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
          f_bulk_g2s(imp + (size_t)ibuf * imp_bytes + (size_t)e * urow, uexp + (size_t)kr * (urow / 8), (uint32_t)urow, &ifull[ibuf]);
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- drain warps: shared memory -> iluCrs / iluDia / uexp
    // thread d < 90 owns entry e9 = d % 9 of row group q = d / 9.  After the copies of a micro-step the three warps
    // meet at a named barrier; one thread then releases the strip's progress counter device-wide (the importers of
    // other strips, and of this strip for rows that left the ring, acquire it) and returns the stage to the loader.
    const int d = tid - (FC_CWARPS + 2) * 32;
    const int q = d / 9, e9 = d - q * 9;
    const bool mine = d < TS_ROWS * 9;
    unsigned long long* prog = A.progress + (size_t)scene * A.ns + strip;
    for (int m = 0; m < nsteps; m++) {
      const int st = m % FC_NST;
      while (f_ld_acquire_cta(&flags[0]) < (uint32_t)(m + 1)) __nanosleep(20);
      const int* M = reinterpret_cast<const int*>(stages + (size_t)st * stage_bytes);
      const unsigned char* sc = stages + (size_t)st * stage_bytes + meta_bytes;
      const int row = mine ? M[4 + q] : -1;
      if (row >= 0) {
        const int crs0 = M[14 + q], nlo = M[24 + q], nup = M[34 + q];
        const unsigned char* src = sc + (q * 9 + e9) * 8;
        double* out = crs + (size_t)crs0 * 9 + e9;
        for (int b = 0; b < nlo; b++) out[(size_t)b * 9] = *reinterpret_cast<const double*>(src + b * 768);  // L_ik
        out += (size_t)nlo * 9;
        const unsigned char* rg = ring + (size_t)(((m % FC_H) * A.gmaxU * 16 + q) * 9 + e9) * 8;
        double* ue = uexp + (size_t)row * (urow / 8) + e9;
        for (int u = 0; u < nup; u++) {
          const double v = *reinterpret_cast<const double*>(rg + (size_t)u * 16 * 72);  // U'_ij
          out[(size_t)u * 9] = v;
          ue[u * 9] = v;
        }
        dia[(size_t)row * 9 + e9] = *reinterpret_cast<const double*>(dinv + (size_t)(m % FC_NDI) * 768 + (q * 9 + e9) * 8);
      }
      asm volatile("bar.sync 7, %0;" ::"n"(FC_DWARPS * 32) : "memory");
      if (d == 0) {
        __threadfence();
        f_st_release_gpu(prog, A.pbase + (unsigned long long)(m + 1));
        f_mbar_arrive(&empty[st]);
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
  A.gmaxB = P.gmaxB; A.gmaxU = P.gmaxU; A.meta_stride = P.meta_stride; A.row_stride = P.row_stride; A.max_imp = P.max_imp; A.kg_stride = P.kg_stride;
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
  if (A.dbg) {
    CG_CUDA(cudaFuncSetAttribute(k_ilu_factor_strips<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes));
    k_ilu_factor_strips<true><<<grid, FC_THREADS, P.smem_bytes, KS(c, "k_ilu_factor_strips")>>>(A);
  } else {
    CG_CUDA(cudaFuncSetAttribute(k_ilu_factor_strips<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes));
    k_ilu_factor_strips<false><<<grid, FC_THREADS, P.smem_bytes, KS(c, "k_ilu_factor_strips")>>>(A);
  }
  CG_LAUNCH_CHECK(c);
  P.pbase += (unsigned long long)P.max_steps + 2;
  P.tbase += (unsigned long long)grid;
  return CLOTHGPU_OK;
}

}  // namespace cg
