// csrc/setup.cu -- topology set-up on the device (runs once per mesh / context, never per step).
//
// Replaces, as data-parallel sort / search passes on the GPU:
//   * CJaggedArray::SetEdgeOfElem + Sort (jagged_array.h:158-207, 71-90) and the pattern CMatrixSquareSparse::SetPattern
//     takes (matrix_square_sparse.cpp:140-163): vertex -> vertex adjacency of an element list as CSR with ascending
//     columns, from one 64-bit radix sort + unique of the (row, col) pairs           -> clothgpu_build_adjacency
//   * CPreconditionerILU::Initialize_ILU0's m_diaInd (ilu_sparse.cpp:114-134): first slot above the diagonal per row
//   * the gather lists of the assembly (per vertex / per off-diagonal slot the element contributions in the
//     reference's element order, solve_internal_sparse.h:36-75): stable radix sorts of (destination, entry) pairs
// What stays on the host: the strip / micro-step plans of the ILU sweeps (trisolve.cu, ilufactor.cu) -- a search over
// strip counts by simulating the pipeline, serial by nature and independent of the scene count -- and the BVH
// topology (bvhbuild.cu), whose flood-fill order is part of the reference's node numbering.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include <vector>

#include "ctx.h"

namespace cg {

namespace {
template <class T>
int dalloc(T** p, size_t n, std::vector<void*>& keep) {
  *p = nullptr;
  CG_CUDA(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
  keep.push_back(*p);
  return CLOTHGPU_OK;
}
struct Scope {
  std::vector<void*> ptrs;
  ~Scope() { for (void* p : ptrs) cudaFree(p); }
};
}  // namespace

// ---- adjacency
__global__ void k_adj_pairs(int nelem, int nnoel, const int* __restrict__ elem, int with_dia, unsigned long long* __restrict__ keys) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)nelem * nnoel * nnoel;
  if (i >= n) return;
  const int e = (int)(i / (nnoel * nnoel)), ab = (int)(i % (nnoel * nnoel)), a = ab / nnoel, b = ab % nnoel;
  const int r = elem[e * nnoel + a], c = elem[e * nnoel + b];
  keys[i] = (r == c && !with_dia) ? ~0ULL : ((unsigned long long)(unsigned)r << 32 | (unsigned)c);
}
// index[r] = first position whose row is >= r (keys sorted, `n` valid entries); array = columns
__global__ void k_adj_finish(int nnode, int n, const unsigned long long* __restrict__ keys, int* __restrict__ index, int* __restrict__ array) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) array[i] = (int)(keys[i] & 0xffffffffu);
  if (i <= nnode) {
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((int)(keys[mid] >> 32) < i) lo = mid + 1; else hi = mid;
    }
    index[i] = lo;
  }
}

// ---- first upper slot and row of every slot
__global__ void k_diaind(int np, const int* __restrict__ rowptr, const int* __restrict__ colidx, int* __restrict__ diaind,
                         int* __restrict__ slot_row) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const int j0 = rowptr[i], j1 = rowptr[i + 1];
  int d = j1;
  for (int j = j0; j < j1; j++) {
    slot_row[j] = i;
    if (d == j1 && colidx[j] > i) d = j;
  }
  diaind[i] = d;
}

// ---- gather lists: (destination key, entry) pairs in entry order; a stable sort by key keeps the element order
// mode 0: per vertex, entries (element, corner): key = vertex
// mode 1: per off-diagonal slot, entries (element, corner a, corner b), a != b: key = slot of (n[a], n[b]) in the
//         pattern, or `nkey` (sorted behind every real slot) when the pattern has no such block (Mearge drops those,
//         matrix_square_sparse.cpp:122)
__global__ void k_gather_keys(int mode, int nelem, int nnoel, const int* __restrict__ elem, const int* __restrict__ rowptr,
                              const int* __restrict__ colidx, int nkey, unsigned* __restrict__ keys, int* __restrict__ vals) {
  const long long per = mode == 0 ? nnoel : nnoel * nnoel;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)nelem * per) return;
  const int e = (int)(i / per), k = (int)(i % per);
  vals[i] = (int)i;
  if (mode == 0) {
    keys[i] = (unsigned)elem[e * nnoel + k];
    return;
  }
  const int a = k / nnoel, b = k % nnoel;
  unsigned key = (unsigned)nkey;
  if (a != b) {
    const int row = elem[e * nnoel + a], col = elem[e * nnoel + b];
    int lo = rowptr[row], hi = rowptr[row + 1];
    const int end = hi;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (colidx[mid] < col) lo = mid + 1; else hi = mid;
    }
    if (lo < end && colidx[lo] == col) key = (unsigned)lo;
  }
  keys[i] = key;
}
__global__ void k_key_ptr(int nkey, int n, const unsigned* __restrict__ keys, int* __restrict__ ptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nkey) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < (unsigned)i) lo = mid + 1; else hi = mid;
  }
  ptr[i] = lo;
}

static int bits_for(long long n) {
  int b = 1;
  while ((1LL << b) <= n) b++;
  return b;
}

// one gather list: ptr[nkey+1], ent[ptr[nkey]] (device arrays owned by the caller afterwards)
static int build_gather_list(int mode, int nelem, int nnoel, const int* d_elem, const int* d_rowptr, const int* d_colidx,
                             int nkey, int** out_ptr, int** out_ent) {
  Scope sc;
  const long long n = (long long)nelem * (mode == 0 ? nnoel : nnoel * nnoel);
  if (n > 0x7fffffffLL) { set_error("set-up: gather list exceeds 2^31 entries"); return CLOTHGPU_ERR_CAPACITY; }
  unsigned *k0, *k1;
  int *v0, *v1;
  CG_CHECK(dalloc(&k0, (size_t)n, sc.ptrs)); CG_CHECK(dalloc(&k1, (size_t)n, sc.ptrs));
  CG_CHECK(dalloc(&v0, (size_t)n, sc.ptrs)); CG_CHECK(dalloc(&v1, (size_t)n, sc.ptrs));
  CG_CUDA(cudaMalloc((void**)out_ptr, sizeof(int) * ((size_t)nkey + 1)));
  if (n > 0) {
    k_gather_keys<<<div_up(n, 256), 256>>>(mode, nelem, nnoel, d_elem, d_rowptr, d_colidx, nkey, k0, v0);
    size_t tb = 0;
    const int end_bit = bits_for(nkey);
    CG_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, k0, k1, v0, v1, (int)n, 0, end_bit));
    void* tmp = nullptr;
    CG_CUDA(cudaMalloc(&tmp, std::max<size_t>(tb, 1)));
    sc.ptrs.push_back(tmp);
    CG_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k0, k1, v0, v1, (int)n, 0, end_bit));
  }
  k_key_ptr<<<div_up(nkey + 1, 256), 256>>>(nkey, (int)n, k1, *out_ptr);
  int nvalid = 0;
  CG_CUDA(cudaMemcpy(&nvalid, *out_ptr + nkey, sizeof(int), cudaMemcpyDeviceToHost));
  CG_CUDA(cudaMalloc((void**)out_ent, sizeof(int) * (size_t)std::max(nvalid, 1)));
  if (nvalid > 0) CG_CUDA(cudaMemcpy(*out_ent, v1, sizeof(int) * (size_t)nvalid, cudaMemcpyDeviceToDevice));
  return CLOTHGPU_OK;
}

// diaind, slot_row and the four gather lists of a context whose tri / quad / rowptr / colidx are on the device already
int build_topology_device(Ctx* c) {
  CG_CUDA(cudaMalloc((void**)&c->diaind, sizeof(int) * (size_t)std::max(c->np, 1)));
  CG_CUDA(cudaMalloc((void**)&c->slot_row, sizeof(int) * (size_t)std::max(c->ncrs, 1)));
  k_diaind<<<div_up(c->np, 128), 128>>>(c->np, c->rowptr, c->colidx, c->diaind, c->slot_row);
  CG_CHECK(build_gather_list(0, c->ntri, 3, c->tri, c->rowptr, c->colidx, c->np, &c->vt_ptr, &c->vt_ent));
  CG_CHECK(build_gather_list(0, c->nquad, 4, c->quad, c->rowptr, c->colidx, c->np, &c->vq_ptr, &c->vq_ent));
  CG_CHECK(build_gather_list(1, c->ntri, 3, c->tri, c->rowptr, c->colidx, c->ncrs, &c->st_ptr, &c->st_ent));
  CG_CHECK(build_gather_list(1, c->nquad, 4, c->quad, c->rowptr, c->colidx, c->ncrs, &c->sq_ptr, &c->sq_ent));
  CG_CUDA(cudaDeviceSynchronize());
  return CLOTHGPU_OK;
}

}  // namespace cg

using namespace cg;

extern "C" int clothgpu_build_adjacency(int nelem, int nnoel, const int* elem, int nnode, int with_diagonal, int* index,
                                        int* array, int cap) {
  if (nelem < 0 || nnoel < 1 || nnoel > 8 || nnode < 1 || !elem || !index) { set_error("build_adjacency: bad argument"); return -CLOTHGPU_ERR_ARG; }
  for (long long i = 0; i < (long long)nelem * nnoel; i++)
    if (elem[i] < 0 || elem[i] >= nnode) { set_error("build_adjacency: element index out of range"); return -CLOTHGPU_ERR_ARG; }
  const long long n = (long long)nelem * nnoel * nnoel;
  if (n > 0x7fffffffLL) { set_error("build_adjacency: too many pairs"); return -CLOTHGPU_ERR_CAPACITY; }
  Scope sc;
  int* d_elem;
  unsigned long long *k0, *k1, *k2;
  int *d_index, *d_array, *d_num;
#define ADJ(call) do { const int r__ = (call); if (r__ != CLOTHGPU_OK) return -r__; } while (0)
#define ADJ_CUDA(call) do { const cudaError_t e__ = (call); if (e__ != cudaSuccess) return -cuda_fail(e__, #call, __FILE__, __LINE__); } while (0)
  ADJ(dalloc(&d_elem, (size_t)nelem * nnoel, sc.ptrs));
  ADJ(dalloc(&k0, (size_t)n, sc.ptrs)); ADJ(dalloc(&k1, (size_t)n, sc.ptrs)); ADJ(dalloc(&k2, (size_t)n, sc.ptrs));
  ADJ(dalloc(&d_index, (size_t)nnode + 1, sc.ptrs)); ADJ(dalloc(&d_array, (size_t)n, sc.ptrs)); ADJ(dalloc(&d_num, 1, sc.ptrs));
  ADJ_CUDA(cudaMemcpy(d_elem, elem, sizeof(int) * (size_t)nelem * nnoel, cudaMemcpyHostToDevice));
  int nuniq = 0;
  if (n > 0) {
    k_adj_pairs<<<div_up(n, 256), 256>>>(nelem, nnoel, d_elem, with_diagonal, k0);
    size_t tb = 0, tb2 = 0;
    ADJ_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, k0, k1, (int)n, 0, 64));
    ADJ_CUDA(cub::DeviceSelect::Unique(nullptr, tb2, k1, k2, d_num, (int)n));
    void* tmp = nullptr;
    ADJ_CUDA(cudaMalloc(&tmp, std::max(tb, tb2) + 16));
    sc.ptrs.push_back(tmp);
    ADJ_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tb, k0, k1, (int)n, 0, 64));
    ADJ_CUDA(cub::DeviceSelect::Unique(tmp, tb2, k1, k2, d_num, (int)n));
    ADJ_CUDA(cudaMemcpy(&nuniq, d_num, sizeof(int), cudaMemcpyDeviceToHost));
    // the dropped (row == col) pairs share one sentinel key, sorted last
    unsigned long long last = 0;
    if (nuniq > 0) ADJ_CUDA(cudaMemcpy(&last, k2 + nuniq - 1, sizeof(last), cudaMemcpyDeviceToHost));
    if (nuniq > 0 && last == ~0ULL) nuniq--;
  }
  k_adj_finish<<<div_up(std::max(nuniq, nnode + 1), 256), 256>>>(nnode, nuniq, k2, d_index, d_array);
  ADJ_CUDA(cudaGetLastError());
  ADJ_CUDA(cudaMemcpy(index, d_index, sizeof(int) * ((size_t)nnode + 1), cudaMemcpyDeviceToHost));
  if (array && nuniq > 0) {
    if (cap < nuniq) { set_error("build_adjacency: array too small"); return -CLOTHGPU_ERR_ARG; }
    ADJ_CUDA(cudaMemcpy(array, d_array, sizeof(int) * (size_t)nuniq, cudaMemcpyDeviceToHost));
  }
#undef ADJ
#undef ADJ_CUDA
  return nuniq;
}
