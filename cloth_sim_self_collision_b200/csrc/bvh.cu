// csrc/bvh.cu -- BVH refit, dual-tree self traversal, narrow phase (proximity and CCD), canonical contact sets.
//
// Replaces BuildBoundingBoxChild_Prx / _CCD (bvh_aabb.cpp:23-61, 64-106), GetContactElement_Proximity / _CCD
// (:441-557, 797-923) with their predicates (IsContact_*; dev_math.cuh) and the std::set<CContactElement>
// container (bvh_aabb.h:65-106).
//
// Design (the reference recurses depth first on one core):
//  * refit: one thread per leaf computes its box and climbs; the second arrival at a parent merges the two
//    children (atomic ticket).  min/max only, so the result is order independent and bit-exact.  Incremental CCD
//    passes (positions fixed, few velocities changed) rebuild only the dirty leaves and their ancestors
//    (k_refit_mark / k_refit_inc).
//  * traversal, full passes: breadth first over a frontier of node pairs.  It starts from the (child0, child1) pair of
//    EVERY internal node (the self() calls of bvh_aabb.cpp:540-557) and expands one tree level per launch; surviving
//    pairs are pushed with warp-aggregated atomics.  The set of leaf pairs reached is the same as the depth-first
//    walk's.  Incremental passes: one warp per dirty leaf queries the tree for the leaves its box meets (k_query).
//  * std::set keeps the FIRST inserted of two elements with equal keys and its comparison ignores is_fv
//    (bvh_aabb.h:96-102), so when a face-vertex and an edge-edge element share a 4-tuple the winner depends on the
//    visiting order of the reference's depth-first walk.  That order is a function of the two leaves alone: preorder
//    rank of their lowest common ancestor (the self() node) + the child choices of the pair recursion, 2 bits per
//    descent.  It is rebuilt from the per-node root paths for narrow-phase hits (pair_rank) -- frontier and leaf
//    entries are just the two node ids -- or carried along in 32-byte entries (CLOTHGPU_SLIM_FRONTIER=0, the
//    cooperative kernel, the rank cross-check); sorting raw elements by (key, DFS rank) and keeping the first of each
//    key reproduces the set bit for bit.
#include <cub/device/device_merge_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "ctx.h"
#include "dev_math.cuh"

namespace cg {

// ------------------------------------------------------------------------------------------------ refit
__device__ __forceinline__ void st_box(double* boxes, int i, const Box& b) {
  double* p = boxes + (size_t)i * 6;
  p[0] = b.lo[0]; p[1] = b.hi[0]; p[2] = b.lo[1]; p[3] = b.hi[1]; p[4] = b.lo[2]; p[5] = b.hi[2];
}
__device__ __forceinline__ Box ld_box_cg(const double* boxes, int i) {  // bypass L1: written by another SM
  Box b;
  const double* p = boxes + (size_t)i * 6;
  b.lo[0] = __ldcg(p + 0); b.hi[0] = __ldcg(p + 1);
  b.lo[1] = __ldcg(p + 2); b.hi[1] = __ldcg(p + 3);
  b.lo[2] = __ldcg(p + 4); b.hi[2] = __ldcg(p + 5);
  return b;
}

// mode 0: BuildBoundingBoxChild_Prx (leaf = 3 vertices padded by par/2); mode 1: _CCD (leaf = 3 vertices at x and
// at x + par*v, padded by 1e-10).  Internal node = CAABB3D::operator+= of its children (aabb.h:55-72).
__global__ void __launch_bounds__(256) k_refit(int nnode, int np, const int* __restrict__ nodes,
                                                const int* __restrict__ tri, const double* __restrict__ xyz,
                                                const double* __restrict__ uvw, double par, int mode, double* boxes,
                                                int* flags, const int* __restrict__ cnt,
                                                const int* __restrict__ vdirty, int* ndirty) {
  const int s = blockIdx.y;
  if (!cnt[s * NCNT + CNT_ACTIVE]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnode) return;
  if (nodes[i * 3 + 2] != -1) return;
  const double* x = xyz + (size_t)s * np * 3;
  double* bx = boxes + (size_t)s * nnode * 6;
  int* fl = flags + (size_t)s * nnode;
  const int t = nodes[i * 3 + 1];
  const int n0 = tri[t * 3 + 0], n1 = tri[t * 3 + 1], n2 = tri[t * 3 + 2];
  Box b;
  if (mode == 0) {
    const double eps = par * 0.5;
    box_init(b, ld3(x, n0), eps);
    box_add(b, ld3(x, n1), eps);
    box_add(b, ld3(x, n2), eps);
  } else {
    const double* v = uvw + (size_t)s * np * 3;
    const double eps = 1.0e-10;
    box_init(b, ld3(x, n0), eps);
    box_add(b, ld3(x, n1), eps);
    box_add(b, ld3(x, n2), eps);
    box_add(b, v3(x[n0 * 3] + par * v[n0 * 3], x[n0 * 3 + 1] + par * v[n0 * 3 + 1], x[n0 * 3 + 2] + par * v[n0 * 3 + 2]), eps);
    box_add(b, v3(x[n1 * 3] + par * v[n1 * 3], x[n1 * 3 + 1] + par * v[n1 * 3 + 1], x[n1 * 3 + 2] + par * v[n1 * 3 + 2]), eps);
    box_add(b, v3(x[n2 * 3] + par * v[n2 * 3], x[n2 * 3 + 1] + par * v[n2 * 3 + 1], x[n2 * 3 + 2] + par * v[n2 * 3 + 2]), eps);
  }
  st_box(bx, i, b);
  // incremental detection: a node is dirty when a vertex of a triangle below it changed its velocity
  int* nd = ndirty ? ndirty + (size_t)s * nnode : nullptr;
  if (nd) {
    const int* vd = vdirty + (size_t)s * np;
    nd[i] = vd[n0] | vd[n1] | vd[n2];
  }
  int cur = nodes[i * 3 + 0];
  while (cur >= 0) {
    // release our subtree's boxes / acquire the sibling's in one fenced atomic
    int seen;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(seen) : "l"(fl + cur) : "memory");
    if (seen == 0) return;  // the sibling subtree is not finished: its last thread continues
    if (nd) nd[cur] = __ldcg(nd + nodes[cur * 3 + 1]) | __ldcg(nd + nodes[cur * 3 + 2]);
    Box a = ld_box_cg(bx, nodes[cur * 3 + 1]);
    const Box o = ld_box_cg(bx, nodes[cur * 3 + 2]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      a.hi[k] = (a.hi[k] > o.hi[k]) ? a.hi[k] : o.hi[k];
      a.lo[k] = (a.lo[k] < o.lo[k]) ? a.lo[k] : o.lo[k];
    }
    st_box(bx, cur, a);
    cur = nodes[cur * 3 + 0];
  }
}

// Incremental refit (second and later CCD passes of a step: positions fixed, velocities changed on few vertices).  The
// box of a leaf whose three vertices kept their velocity is unchanged, and so is every subtree without such a leaf:
// only the dirty leaves and their ancestors are recomputed.
//   k_refit_mark: every dirty leaf climbs and marks its ancestors dirty until it meets one that is marked already;
//                 a node that turns dirty counts itself at its parent (nchild = number of dirty children, 1 or 2).
//   k_refit_inc : every dirty leaf rebuilds its box and climbs; a parent with one dirty child is merged at once with
//                 the sibling's box of the previous pass, a parent with two dirty children by the second arrival.
__device__ __forceinline__ bool leaf_dirty(int i, const int* __restrict__ nodes, const int* __restrict__ tri,
                                           const int* __restrict__ vd, int& n0, int& n1, int& n2) {
  const int t = nodes[i * 3 + 1];
  n0 = tri[t * 3 + 0]; n1 = tri[t * 3 + 1]; n2 = tri[t * 3 + 2];
  return (vd[n0] | vd[n1] | vd[n2]) != 0;
}
__global__ void __launch_bounds__(256) k_refit_mark(int nnode, int np, const int* __restrict__ nodes,
                                                     const int* __restrict__ tri, const int* __restrict__ vdirty,
                                                     int* ndirty, int* nchild, const int* __restrict__ cnt) {
  const int s = blockIdx.y;
  if (!cnt[s * NCNT + CNT_ACTIVE]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnode) return;
  if (nodes[i * 3 + 2] != -1) return;
  int n0, n1, n2;
  if (!leaf_dirty(i, nodes, tri, vdirty + (size_t)s * np, n0, n1, n2)) return;
  int* nd = ndirty + (size_t)s * nnode;
  int* nc = nchild + (size_t)s * nnode;
  nd[i] = 1;
  int x = i;
  for (;;) {
    const int p = nodes[x * 3 + 0];
    if (p < 0) break;
    atomicAdd(&nc[p], 1);
    if (atomicExch(&nd[p], 1) != 0) break;  // an earlier climber continues from here
    x = p;
  }
}
__global__ void __launch_bounds__(256) k_refit_inc(int nnode, int np, const int* __restrict__ nodes,
                                                    const int* __restrict__ tri, const double* __restrict__ xyz,
                                                    const double* __restrict__ uvw, double par, double* boxes, int* flags,
                                                    int* cnt, const int* __restrict__ vdirty,
                                                    const int* __restrict__ nchild, int* __restrict__ dirty_list) {
  const int s = blockIdx.y;
  if (!cnt[s * NCNT + CNT_ACTIVE]) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nnode) return;
  if (nodes[i * 3 + 2] != -1) return;
  int n0, n1, n2;
  if (!leaf_dirty(i, nodes, tri, vdirty + (size_t)s * np, n0, n1, n2)) return;
  dirty_list[(size_t)s * nnode + atomicAdd(&cnt[s * NCNT + CNT_NDIRTY], 1)] = i;  // the leaves k_query starts from
  const double* x = xyz + (size_t)s * np * 3;
  const double* v = uvw + (size_t)s * np * 3;
  double* bx = boxes + (size_t)s * nnode * 6;
  int* fl = flags + (size_t)s * nnode;
  const int* nc = nchild + (size_t)s * nnode;
  Box b;
  const double eps = 1.0e-10;  // BuildBoundingBoxChild_CCD, as in k_refit mode 1
  box_init(b, ld3(x, n0), eps);
  box_add(b, ld3(x, n1), eps);
  box_add(b, ld3(x, n2), eps);
  box_add(b, v3(x[n0 * 3] + par * v[n0 * 3], x[n0 * 3 + 1] + par * v[n0 * 3 + 1], x[n0 * 3 + 2] + par * v[n0 * 3 + 2]), eps);
  box_add(b, v3(x[n1 * 3] + par * v[n1 * 3], x[n1 * 3 + 1] + par * v[n1 * 3 + 1], x[n1 * 3 + 2] + par * v[n1 * 3 + 2]), eps);
  box_add(b, v3(x[n2 * 3] + par * v[n2 * 3], x[n2 * 3 + 1] + par * v[n2 * 3 + 1], x[n2 * 3 + 2] + par * v[n2 * 3 + 2]), eps);
  st_box(bx, i, b);
  int cur = nodes[i * 3 + 0];
  while (cur >= 0) {
    if (nc[cur] == 2) {
      // both subtrees change: release ours / acquire the sibling's in one fenced atomic, the second arrival merges
      int seen;
      asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(seen) : "l"(fl + cur) : "memory");
      if (seen == 0) return;
    }
    Box a = ld_box_cg(bx, nodes[cur * 3 + 1]);
    const Box o = ld_box_cg(bx, nodes[cur * 3 + 2]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      a.hi[k] = (a.hi[k] > o.hi[k]) ? a.hi[k] : o.hi[k];
      a.lo[k] = (a.lo[k] < o.lo[k]) ? a.lo[k] : o.lo[k];
    }
    st_box(bx, cur, a);
    cur = nodes[cur * 3 + 0];
  }
}

// ------------------------------------------------------------------------------------------------ traversal
__global__ void k_trav_reset(int* cnt) {
  int* c = cnt + blockIdx.x * NCNT;
  for (int k = threadIdx.x; k < CNT_RANK_BAD - CNT_FRONT; k += blockDim.x) c[CNT_FRONT + k] = 0;
  if (threadIdx.x == 0) {
    c[CNT_RAWPREV] = c[CNT_RAW];  // raw elements of the previous detection (now in the other raw buffer)
    c[CNT_LEAF] = 0; c[CNT_RAW] = 0; c[CNT_ELEM] = 0; c[CNT_OVERFLOW] = 0; c[CNT_CAND] = 0; c[CNT_GBAR] = 0;
  }
}

// the self() calls: every internal node starts one pair recursion on its two children
// SLIM: frontier / leaf entries are just the two node ids (8 bytes instead of 32); the depth-first rank of a leaf
// pair is then rebuilt from the root paths for narrow-phase hits (pair_rank), as for the leaf query
template <bool SLIM>
__global__ void k_trav_init(int ninternal, const int* __restrict__ internal_list, const int* __restrict__ nodes,
                            const uint32_t* __restrict__ pre, PairEnt* front, long long cap, int* cnt,
                            int nnode, const int* __restrict__ ndirty) {
  const int s = blockIdx.y;
  if (!cnt[s * NCNT + CNT_ACTIVE]) return;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 0) cnt[s * NCNT + CNT_FRONT] = ninternal;
  if (k >= ninternal) return;
  const int n = internal_list[k];
  PairEnt e;
  e.i0 = nodes[n * 3 + 1];
  e.i1 = nodes[n * 3 + 2];
  if (ndirty && !ndirty[(size_t)s * nnode + n]) e.i0 = -1;  // nothing below this node changed: a dead entry
  if (SLIM) {
    reinterpret_cast<int2*>(front)[(size_t)s * cap + k] = make_int2(e.i0, e.i1);
    return;
  }
  e.pre = pre[n];
  e.depth = 0;
  e.path_hi = 0;
  e.path_lo = 0;
  front[(size_t)s * cap + k] = e;
}

// reserve n (0..4) consecutive output slots for every lane of a fully converged warp with one atomic per warp
__device__ __forceinline__ int warp_reserve(int n, int* counter) {
  const unsigned lane = threadIdx.x & 31;
  int incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (unsigned)o) incl += v;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  int base = 0;
  if (lane == 31 && total > 0) base = atomicAdd(counter, total);
  base = __shfl_sync(0xffffffffu, base, 31);
  return base + incl - n;
}

__device__ __forceinline__ void path_push(PairEnt& e, unsigned code) {
  if (e.depth < 32) e.path_hi |= (uint64_t)code << (62 - 2 * e.depth);
  else e.path_lo |= (uint64_t)code << (62 - 2 * (e.depth - 32));
  e.depth++;
}

// one level of the pair recursion (bvh_aabb.cpp:441-475 / 797-833): prune on CAABB3D::IsIntersect, descend in the
// child order (0,0),(1,0),(0,1),(1,1), collect leaf pairs
template <bool SLIM>
__device__ __forceinline__ void trav_step_body(int d, int nnode, const int* __restrict__ nodes,
                                               const double* __restrict__ boxes, const PairEnt* __restrict__ fin,
                                               PairEnt* __restrict__ fout, long long cap_front,
                                               PairEnt* __restrict__ leaf, long long cap_leaf, int* cnt,
                                               const int* __restrict__ ndirty) {
  const int s = blockIdx.y;
  int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int* nd = ndirty ? ndirty + (size_t)s * nnode : nullptr;
  const int nin = min(c[CNT_FRONT + d], (int)min(cap_front, (long long)INT_MAX));
  if (nin == 0) return;
  const double* bx = boxes + (size_t)s * nnode * 6;
  const PairEnt* in = fin + (size_t)s * cap_front;
  PairEnt* out = fout + (size_t)s * cap_front;
  PairEnt* lf = leaf + (size_t)s * cap_leaf;
  const int stride = gridDim.x * blockDim.x;
  const int nround = (nin + stride - 1) / stride;
  for (int rnd = 0; rnd < nround; rnd++) {
    const int idx = rnd * stride + blockIdx.x * blockDim.x + threadIdx.x;
    // The child pairs are pruned BEFORE they are queued (box test and, for incremental passes, the dirty test of the
    // children), and intersecting leaf pairs go straight to the leaf list: dead pairs never travel through the
    // frontier (32 bytes written and read back each).  An entry of depth > 0 has therefore passed its own test
    // already; the seeds of k_trav_init (depth 0) are tested here.  Same pairs, same paths as the recursion.
    PairEnt e;
    int na = 0, nb = 0;            // children on either side (1 = the node itself is a leaf)
    int a[2] = {0, 0}, b[2] = {0, 0};
    bool la[2] = {false, false}, lb[2] = {false, false};
    unsigned okmask = 0;           // bit (ib * 2 + ia): child pair (a[ia], b[ib]) survives
    bool self_leaf = false;        // a seed whose two sides are leaves
    if (idx < nin) {
      if (SLIM) {
        const int2 p = reinterpret_cast<const int2*>(fin)[(size_t)s * cap_front + idx];
        e.i0 = p.x; e.i1 = p.y; e.depth = (uint32_t)d; e.pre = 0; e.path_hi = 0; e.path_lo = 0;
      } else {
        e = in[idx];
      }
      bool live = e.i0 >= 0;
      if (live && e.depth == 0) {  // (the seeds are exactly the entries of level 0)
        live = (!nd || nd[e.i0] || nd[e.i1]) && box_intersect(ld_box(bx, e.i0), ld_box(bx, e.i1));
      }
      if (live) {
        const int c00 = nodes[e.i0 * 3 + 1], c01 = nodes[e.i0 * 3 + 2];
        const int c10 = nodes[e.i1 * 3 + 1], c11 = nodes[e.i1 * 3 + 2];
        const bool l0 = c01 == -1, l1 = c11 == -1;
        if (l0 && l1) {
          self_leaf = true;
        } else {
          if (l0) { na = 1; a[0] = e.i0; la[0] = true; }
          else { na = 2; a[0] = c00; a[1] = c01; }
          if (l1) { nb = 1; b[0] = e.i1; lb[0] = true; }
          else { nb = 2; b[0] = c10; b[1] = c11; }
          Box ba[2], bb[2];
          bool da[2] = {true, true}, db[2] = {true, true};
#pragma unroll
          for (int k = 0; k < 2; k++) {
            if (k < na) {
              ba[k] = ld_box(bx, a[k]);
              if (!l0) la[k] = nodes[a[k] * 3 + 2] == -1;
              if (nd) da[k] = nd[a[k]] != 0;
            }
            if (k < nb) {
              bb[k] = ld_box(bx, b[k]);
              if (!l1) lb[k] = nodes[b[k] * 3 + 2] == -1;
              if (nd) db[k] = nd[b[k]] != 0;
            }
          }
#pragma unroll
          for (int ib = 0; ib < 2; ib++)
#pragma unroll
            for (int ia = 0; ia < 2; ia++)
              if (ia < na && ib < nb && (da[ia] || db[ib]) && box_intersect(ba[ia], bb[ib])) okmask |= 1u << (ib * 2 + ia);
        }
      }
    }
    // survivors: leaf pairs to the leaf list, the others to the next frontier
    unsigned leafmask = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
      if ((okmask >> k & 1) && la[k & 1] && lb[k >> 1]) leafmask |= 1u << k;
    const unsigned frontmask = okmask & ~leafmask;
    const int nfront = __popc(frontmask), nleaf = __popc(leafmask) + (self_leaf ? 1 : 0);
    int fpos = warp_reserve(nfront, &c[CNT_FRONT + d + 1]);
    int lpos = warp_reserve(nleaf, &c[CNT_LEAF]);
    if (nfront > 0 && (long long)fpos + nfront > cap_front) { atomicOr(&c[CNT_OVERFLOW], 1); okmask &= leafmask; }
    if (nleaf > 0 && (long long)lpos + nleaf > cap_leaf) { atomicOr(&c[CNT_OVERFLOW], 2); okmask &= ~leafmask; self_leaf = false; }
    if (SLIM) {
      int2* lf2 = reinterpret_cast<int2*>(leaf) + (size_t)s * cap_leaf;
      int2* out2 = reinterpret_cast<int2*>(fout) + (size_t)s * cap_front;
      if (self_leaf) lf2[lpos++] = make_int2(e.i0, e.i1);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (!(okmask >> k & 1)) continue;
        const int2 ch = make_int2(a[k & 1], b[k >> 1]);
        if (leafmask >> k & 1) lf2[lpos++] = ch;
        else out2[fpos++] = ch;
      }
      continue;
    }
    if (self_leaf) lf[lpos++] = e;
    if (okmask) {
      PairEnt ch = e;
      path_push(ch, 0);  // depth + 1, code bits filled below
      const int sh = (e.depth < 32) ? 62 - 2 * e.depth : 62 - 2 * (e.depth - 32);
      uint64_t* pw = (e.depth < 32) ? &ch.path_hi : &ch.path_lo;
      const uint64_t p0 = *pw;
      // child order of the recursion: (0,0),(1,0),(0,1),(1,1); with a leaf on one side the two pairs get codes 0, 1
      const bool both = na == 2 && nb == 2;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (!(okmask >> k & 1)) continue;
        const int ia = k & 1, ib = k >> 1;
        const unsigned code = both ? (unsigned)k : (unsigned)(na == 2 ? ia : ib);
        ch.i0 = a[ia]; ch.i1 = b[ib];
        *pw = p0 | ((uint64_t)code << sh);
        if (leafmask >> k & 1) lf[lpos++] = ch;
        else out[fpos++] = ch;
      }
    }
  }
}

template <bool SLIM>
__global__ void __launch_bounds__(256) k_trav_step(int d, int nnode, const int* __restrict__ nodes,
                                                    const double* __restrict__ boxes, const PairEnt* __restrict__ fin,
                                                    PairEnt* __restrict__ fout, long long cap_front,
                                                    PairEnt* __restrict__ leaf, long long cap_leaf, int* cnt,
                                                    const int* __restrict__ ndirty) {
  trav_step_body<SLIM>(d, nnode, nodes, boxes, fin, fout, cap_front, leaf, cap_leaf, cnt, ndirty);
}

// All levels of the pair recursion in ONE cooperative launch (co-residency guaranteed by the launch API): the CTAs
// of a scene meet at a counter barrier between two levels instead of paying a kernel launch per BVH depth.
__global__ void __launch_bounds__(256) k_trav_all(int ndepth, int nnode, const int* __restrict__ nodes,
                                                   const double* __restrict__ boxes, PairEnt* frontA, PairEnt* frontB,
                                                   long long cap_front, PairEnt* __restrict__ leaf, long long cap_leaf,
                                                   int* cnt, const int* __restrict__ ndirty) {
  const int s = blockIdx.y;
  int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  for (int d = 0; d < ndepth; d++) {
    // level d is complete (barrier below / previous kernel): an empty frontier ends the recursion for every CTA
    if (((volatile int*)c)[CNT_FRONT + d] == 0) break;
    trav_step_body<false>(d, nnode, nodes, boxes, (d & 1) ? frontB : frontA, (d & 1) ? frontA : frontB, cap_front, leaf,
                          cap_leaf, cnt, ndirty);
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(&c[CNT_GBAR], 1);
      const int want = (d + 1) * (int)gridDim.x;
      while (((volatile int*)c)[CNT_GBAR] < want) { }
      __threadfence();
    }
    __syncthreads();
  }
}

// Incremental passes without a frontier: a leaf pair survives the pair recursion iff the two leaf boxes intersect
// (the boxes of the ancestors enclose them) and, in a dirty-restricted pass, iff one of the two leaves is dirty (the
// ancestors of a dirty leaf are dirty).  So every dirty leaf queries the tree for the leaves its box meets -- one
// launch, no level barriers.  A pair of two dirty leaves is reported by the one with the smaller node index.  The
// pair is stored in the recursion's orientation (first = the leaf under child 0 of the lowest common ancestor); its
// depth-first rank is not carried along but rebuilt from the two root-to-leaf paths for the (rare) narrow-phase hits
// (pair_rank below): entries with pre == 0xffffffff are "lazy".
enum { QRY_WARPS = 4, QRY_CAP = 256 };
__global__ void __launch_bounds__(QRY_WARPS * 32) k_query(int nnode, const int* __restrict__ nodes,
                                                           const double* __restrict__ boxes,
                                                           const unsigned long long* __restrict__ npath,
                                                           const int* __restrict__ ndirty,
                                                           const int* __restrict__ dirty_list, PairEnt* __restrict__ leaf,
                                                           long long cap_leaf, int* cnt, int qcap, int slim) {
  // one warp per dirty leaf, breadth first: the lanes test the children of up to 32 frontier nodes at once, so a query
  // costs one round of dependent loads per tree level instead of one per visited node
  __shared__ int q[QRY_WARPS][2][QRY_CAP];
  const int s = blockIdx.y;
  int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int ndl = c[CNT_NDIRTY];
  const int* nd = ndirty + (size_t)s * nnode;
  const int* dl = dirty_list + (size_t)s * nnode;
  const double* bx = boxes + (size_t)s * nnode * 6;
  PairEnt* lf = leaf + (size_t)s * cap_leaf;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int w = blockIdx.x * QRY_WARPS + warp; w < ndl; w += gridDim.x * QRY_WARPS) {
    const int i = dl[w];
    const Box bl = ld_box(bx, i);
    const unsigned long long pi = npath[i];
    int cur = 0, ncur = 1;
    if (lane == 0) q[warp][0][0] = 0;
    __syncwarp();
    while (ncur > 0) {
      int nnext = 0;
      for (int base = 0; base < ncur; base += 32) {
        const int n = (base + lane < ncur) ? q[warp][cur][base + lane] : -1;
        int ch[2] = {-1, -1};
        bool hit[2] = {false, false}, isleaf[2] = {false, false};
        if (n >= 0) {
          ch[0] = nodes[n * 3 + 1];
          ch[1] = nodes[n * 3 + 2];
          if (ch[1] != -1) {  // (a single-leaf tree has no pairs)
#pragma unroll
            for (int k = 0; k < 2; k++) {
              hit[k] = box_intersect(ld_box(bx, ch[k]), bl);
              isleaf[k] = nodes[ch[k] * 3 + 2] == -1;
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
          if (hit[k] && isleaf[k]) {
            const int m = ch[k];
            if (m != i && !(nd[m] && m < i)) {  // two dirty leaves: the smaller node index reports the pair
              const int pos = atomicAdd(&c[CNT_LEAF], 1);
              if (pos >= cap_leaf) {
                atomicOr(&c[CNT_OVERFLOW], 2);
              } else {
                PairEnt e;
                const bool first = pi < npath[m];  // 0 at the first differing choice = under child 0 of the common ancestor
                e.i0 = first ? i : m;
                e.i1 = first ? m : i;
                e.pre = 0xffffffffu;
                e.depth = 0;
                e.path_hi = 0;
                e.path_lo = 0;
                if (slim) reinterpret_cast<int2*>(leaf)[(size_t)s * cap_leaf + pos] = make_int2(e.i0, e.i1);
                else lf[pos] = e;
              }
            }
          }
        }
        const int npush = (hit[0] && !isleaf[0] ? 1 : 0) + (hit[1] && !isleaf[1] ? 1 : 0);
        int incl = npush;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        int pos = nnext + incl - npush;
        nnext += __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll
        for (int k = 0; k < 2; k++) {
          if (hit[k] && !isleaf[k]) {
            if (pos < qcap) q[warp][cur ^ 1][pos] = ch[k];
            pos++;
          }
        }
      }
      if (nnext > qcap) {  // a frontier wider than the queue: the host repeats the pass with the frontier traversal
        if (lane == 0) atomicOr(&c[CNT_OVERFLOW], 16);
        nnext = 0;
      }
      __syncwarp();
      cur ^= 1;
      ncur = nnext;
    }
  }
}

// depth-first rank of the leaf pair (i0, i1) of the pair recursion: preorder rank of the lowest common ancestor (the
// node whose self() call starts the recursion on its two children) and the child codes of the descents
__device__ void pair_rank(int i0, int i1, const int* __restrict__ nodes, const uint32_t* __restrict__ pre,
                          const unsigned long long* __restrict__ npath, const int* __restrict__ ndepth, PairEnt& e) {
  const unsigned long long pa = npath[i0], pb = npath[i1];
  const int da = ndepth[i0], db = ndepth[i1];
  const int dl = __clzll((long long)(pa ^ pb));  // depth of the lowest common ancestor
  int lca = i0;
  for (int k = da; k > dl; k--) lca = nodes[lca * 3 + 0];
  e.pre = pre[lca];
  e.depth = 0;
  e.path_hi = 0;
  e.path_lo = 0;
  int la = da - dl - 1, lb = db - dl - 1;       // descents left on either side below the ancestor's children
  unsigned long long ra = pa << (dl + 1), rb = pb << (dl + 1);  // (dl + 1 <= 61)
  while (la > 0 || lb > 0) {
    unsigned code;
    if (la > 0 && lb > 0) {
      code = (unsigned)(ra >> 63) | ((unsigned)(rb >> 63) << 1);
      ra <<= 1; rb <<= 1; la--; lb--;
    } else if (la > 0) {
      code = (unsigned)(ra >> 63);
      ra <<= 1; la--;
    } else {
      code = (unsigned)(rb >> 63);
      rb <<= 1; lb--;
    }
    path_push(e, code);
  }
}

// ------------------------------------------------------------------------------------------------ narrow phase
// CContactElement's canonical ordering (bvh_aabb.h:67-95)
__device__ __forceinline__ void canon_fv(int j0, int j1, int j2, int j3, int* n) {
  int a = j0, b = j1, c = j2, t;
  if (a > b) { t = a; a = b; b = t; }
  if (b > c) { t = b; b = c; c = t; }
  if (a > b) { t = a; a = b; b = t; }
  n[0] = a; n[1] = b; n[2] = c; n[3] = j3;
}
__device__ __forceinline__ void canon_ee(int j0, int j1, int j2, int j3, int* n) {
  int a0 = j0, a1 = j1, b0 = j2, b1 = j3, t;
  const int mA = a0 < a1 ? a0 : a1, mB = b0 < b1 ? b0 : b1;
  if (mB < mA) { t = a0; a0 = b0; b0 = t; t = a1; a1 = b1; b1 = t; }
  if (a1 < a0) { t = a0; a0 = a1; a1 = t; }
  if (b1 < b0) { t = b0; b0 = b1; b1 = t; }
  n[0] = a0; n[1] = a1; n[2] = b0; n[3] = b1;
}

__device__ __forceinline__ void emit(const PairEnt& e, int slot, const int* key, RawElem* raw, long long cap_raw, int* c) {
  const int k = atomicAdd(&c[CNT_RAW], 1);
  if (k >= cap_raw) { atomicOr(&c[CNT_OVERFLOW], 4); return; }
  RawElem r;
  r.n[0] = key[0]; r.n[1] = key[1]; r.n[2] = key[2]; r.n[3] = key[3];
  r.pre = e.pre; r.slot = (uint32_t)slot; r.path_hi = e.path_hi; r.path_lo = e.path_lo;
  raw[k] = r;
}

// The 6 face-vertex + 9 edge-edge tests of every leaf pair (bvh_aabb.cpp:476-535 proximity, 834-900 CCD), in two
// passes.  k_cull: one thread per (leaf pair, test slot) runs the cheap prefix of the predicate (shared vertex, box /
// slab / sign culls) and pushes the survivors into a candidate queue (warp-aggregated atomics).  k_narrow: one thread
// per candidate runs the full predicate -- the cubic coplanarity solve of the CCD tests, which the reference pays
// for every vertex-disjoint edge pair, is then executed by converged warps on a ~50x shorter list.
struct SlotData {
  int n[4];     // FV: face i0,i1,i2 + vertex i3 ; EE: edge i0,i1 + edge j0,j1
  V3 p[4];      // positions at the start of the step
  V3 q[4];      // positions at the end of the step (CCD only)
  Box bb;       // box the FV tests cull against
};

__device__ __forceinline__ void load_slot(int ccd, int slot, int i0, int i1, const int4* __restrict__ node_tri,
                                          const double* __restrict__ bx, const double* __restrict__ x,
                                          const double* __restrict__ v, double dt, SlotData& S) {
  const int4 ta = node_tri[i0], tb = node_tri[i1];  // {v0, v1, v2, triangle} of the two leaves
  const int ia[3] = {ta.x, ta.y, ta.z}, ib[3] = {tb.x, tb.y, tb.z};
  if (slot < 6) {
    const bool first = slot < 3;
    const int a = first ? slot : slot - 3;
    const int* f = first ? ia : ib;
    const int* o = first ? ib : ia;
    S.n[0] = f[0]; S.n[1] = f[1]; S.n[2] = f[2]; S.n[3] = o[a];
    // reference quirk (bvh_aabb.cpp:490-505): the proximity FV test culls against the box of BVH node number
    // <triangle index>, not against the box of the face's own leaf; the CCD test uses the leaf (:855)
    S.bb = ld_box(bx, ccd ? (first ? i0 : i1) : (first ? ta.w : tb.w));
  } else {
    const int k = slot - 6, a0 = k / 3, a1 = (a0 + 1) % 3, b0 = k - a0 * 3, b1 = (b0 + 1) % 3;
    S.n[0] = ia[a0]; S.n[1] = ia[a1]; S.n[2] = ib[b0]; S.n[3] = ib[b1];
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int i = S.n[k];
    S.p[k] = ld3(x, i);
    if (ccd) S.q[k] = v3(x[i * 3] + dt * v[i * 3], x[i * 3 + 1] + dt * v[i * 3 + 1], x[i * 3 + 2] + dt * v[i * 3 + 2]);
  }
}

// one thread per leaf pair: the six vertices are loaded once, the 15 cheap culls run on registers, the survivors'
// slot mask is compacted with one atomic per warp
// (register caps: both kernels are bound by load latency, two / four resident CTAs per SM hide it better than the
// 144 registers the compiler takes when left alone -- cull 0.34 -> 0.24 ms, narrow 0.41 -> 0.27 ms per step)
#ifndef CULL_T
#define CULL_T 256
#define CULL_B 2
#endif
#ifndef NARROW_B
#define NARROW_B 4
#endif
__global__ void __launch_bounds__(CULL_T, CULL_B) k_cull(int ccd, int nnode, int np, const int4* __restrict__ node_tri,
                                               const double* __restrict__ boxes,
                                               const double* __restrict__ xyz, const double* __restrict__ uvw,
                                               double dt, double delta, const PairEnt* __restrict__ leaf,
                                               long long cap_leaf, int4* __restrict__ cand, long long cap_cand, int* cnt,
                                               int slim) {
  const int s = blockIdx.y;
  int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int nleaf = min(c[CNT_LEAF], (int)min(cap_leaf, (long long)INT_MAX));
  const double* bx = boxes + (size_t)s * nnode * 6;
  const double* x = xyz + (size_t)s * np * 3;
  const double* v = ccd ? uvw + (size_t)s * np * 3 : nullptr;
  const PairEnt* lf = leaf + (size_t)s * cap_leaf;
  int4* cd = cand + (size_t)s * cap_cand;
  const int lane = threadIdx.x & 31;
  for (int i0 = blockIdx.x * blockDim.x; i0 < nleaf; i0 += gridDim.x * blockDim.x) {
    const int idx = i0 + threadIdx.x;
    unsigned mask = 0;
    int ei0 = 0, ei1 = 0;
    if (idx < nleaf) {
      int2 e;
      if (slim) e = reinterpret_cast<const int2*>(leaf)[(size_t)s * cap_leaf + idx];
      else e = make_int2(lf[idx].i0, lf[idx].i1);
      ei0 = e.x; ei1 = e.y;
      const int4 ta = node_tri[e.x], tb = node_tri[e.y];  // {v0, v1, v2, triangle} of the two leaves
      const int it = ta.w, jt = tb.w;
      const int in[3] = {ta.x, ta.y, ta.z}, jn[3] = {tb.x, tb.y, tb.z};
      V3 p[3], q[3], pe[3], qe[3];
#pragma unroll
      for (int a = 0; a < 3; a++) { p[a] = ld3(x, in[a]); q[a] = ld3(x, jn[a]); }
      if (ccd) {
#pragma unroll
        for (int a = 0; a < 3; a++) {
          const int i = in[a], j = jn[a];
          pe[a] = v3(x[i * 3] + dt * v[i * 3], x[i * 3 + 1] + dt * v[i * 3 + 1], x[i * 3 + 2] + dt * v[i * 3 + 2]);
          qe[a] = v3(x[j * 3] + dt * v[j * 3], x[j * 3 + 1] + dt * v[j * 3 + 1], x[j * 3 + 2] + dt * v[j * 3 + 2]);
        }
        const Box b0 = ld_box(bx, e.x), b1 = ld_box(bx, e.y);
#pragma unroll
        for (int a = 0; a < 3; a++) {
          if (maybe_fv_ccd(in[0], in[1], in[2], jn[a], p[0], p[1], p[2], q[a], pe[0], pe[1], pe[2], qe[a], b0)) mask |= 1u << a;
          if (maybe_fv_ccd(jn[0], jn[1], jn[2], in[a], q[0], q[1], q[2], p[a], qe[0], qe[1], qe[2], pe[a], b1)) mask |= 1u << (3 + a);
        }
#pragma unroll
        for (int k = 0; k < 9; k++) {
          const int a0 = k / 3, a1 = (a0 + 1) % 3, b0e = k - a0 * 3, b1e = (b0e + 1) % 3;
          if (maybe_ee_ccd(in[a0], in[a1], jn[b0e], jn[b1e], p[a0], p[a1], q[b0e], q[b1e], pe[a0], pe[a1], qe[b0e], qe[b1e]))
            mask |= 1u << (6 + k);
        }
      } else {
        // reference quirk (bvh_aabb.cpp:490-505): box of BVH node number <triangle index>
        const Box b0 = ld_box(bx, it), b1 = ld_box(bx, jt);
#pragma unroll
        for (int a = 0; a < 3; a++) {
          if (maybe_fv_prx(in[0], in[1], in[2], jn[a], q[a], b0)) mask |= 1u << a;
          if (maybe_fv_prx(jn[0], jn[1], jn[2], in[a], p[a], b1)) mask |= 1u << (3 + a);
        }
#pragma unroll
        for (int k = 0; k < 9; k++) {
          const int a0 = k / 3, a1 = (a0 + 1) % 3, b0e = k - a0 * 3, b1e = (b0e + 1) % 3;
          if (maybe_ee_prx(in[a0], in[a1], jn[b0e], jn[b1e], p[a0], p[a1], q[b0e], q[b1e], delta)) mask |= 1u << (6 + k);
        }
      }
    }
    // warp-level compaction: exclusive prefix of the survivor counts, one atomic per warp
    const int n = __popc(mask);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;
    int base = 0;
    if (lane == 31) base = atomicAdd(&c[CNT_CAND], total);
    base = __shfl_sync(0xffffffffu, base, 31);
    long long pos = (long long)base + incl - n;
    while (mask) {
      const int slot = __ffs(mask) - 1;
      mask &= mask - 1;
      if (pos < cap_cand) cd[pos] = make_int4(ei0, ei1, slot, idx);
      else atomicOr(&c[CNT_OVERFLOW], 2);
      pos++;
    }
  }
}

__global__ void __launch_bounds__(128, NARROW_B) k_narrow(int ccd, int nnode, int np, const int4* __restrict__ node_tri,
                                                 const double* __restrict__ boxes,
                                                 const double* __restrict__ xyz, const double* __restrict__ uvw,
                                                 double dt, double delta, const PairEnt* __restrict__ leaf,
                                                 long long cap_leaf, const int4* __restrict__ cand, long long cap_cand,
                                                 RawElem* __restrict__ raw, long long cap_raw, int* cnt,
                                                 const int* __restrict__ nodes, const uint32_t* __restrict__ pre,
                                                 const unsigned long long* __restrict__ npath,
                                                 const int* __restrict__ ndepth, int rank_check, int slim) {
  const int s = blockIdx.y;
  int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int ncand = (int)min((long long)c[CNT_CAND], cap_cand);
  const double* bx = boxes + (size_t)s * nnode * 6;
  const double* x = xyz + (size_t)s * np * 3;
  const double* v = ccd ? uvw + (size_t)s * np * 3 : nullptr;
  const PairEnt* lf = leaf + (size_t)s * cap_leaf;
  const int4* cd = cand + (size_t)s * cap_cand;
  RawElem* rw = raw + (size_t)s * cap_raw;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ncand; k += gridDim.x * blockDim.x) {
    const int4 cs = cd[k];  // {leaf node i0, leaf node i1, slot, leaf pair index}
    const int slot = cs.z;
    SlotData S;
    load_slot(ccd, slot, cs.x, cs.y, node_tri, bx, x, v, dt, S);
    bool hit;
    if (ccd)
      hit = slot < 6 ? is_fv_ccd(S.n[0], S.n[1], S.n[2], S.n[3], S.p[0], S.p[1], S.p[2], S.p[3], S.q[0], S.q[1], S.q[2], S.q[3], S.bb)
                     : is_ee_ccd(S.n[0], S.n[1], S.n[2], S.n[3], S.p[0], S.p[1], S.p[2], S.p[3], S.q[0], S.q[1], S.q[2], S.q[3]);
    else
      hit = slot < 6 ? is_fv_prx(S.n[0], S.n[1], S.n[2], S.n[3], S.p[0], S.p[1], S.p[2], S.p[3], S.bb, delta)
                     : is_ee_prx(S.n[0], S.n[1], S.n[2], S.n[3], S.p[0], S.p[1], S.p[2], S.p[3], delta);
    if (hit) {
      int key[4];
      if (slot < 6) canon_fv(S.n[0], S.n[1], S.n[2], S.n[3], key);
      else canon_ee(S.n[0], S.n[1], S.n[2], S.n[3], key);
      PairEnt e;  // the pair's DFS rank is only needed for the (rare) hits
      if (slim) { e.i0 = cs.x; e.i1 = cs.y; e.pre = 0xffffffffu; }
      else e = lf[cs.w];
      if (e.pre == 0xffffffffu) {
        pair_rank(e.i0, e.i1, nodes, pre, npath, ndepth, e);
      } else if (rank_check) {
        // CLOTHGPU_RANK_CHECK=1 (tests): a pair of the frontier traversal carries its rank; the reconstruction from the
        // root paths and the orientation rule of k_query must give the same
        PairEnt e2 = e;
        pair_rank(e.i0, e.i1, nodes, pre, npath, ndepth, e2);
        if (e2.pre != e.pre || e2.path_hi != e.path_hi || e2.path_lo != e.path_lo || !(npath[e.i0] < npath[e.i1]))
          atomicAdd(&c[CNT_RANK_BAD], 1);
        atomicAdd(&c[CNT_RANK_BAD + 1], 1);
      }
      emit(e, slot, key, rw, cap_raw, c);
    }
  }
}

// Incremental detection (second and later CCD passes of a step: positions fixed, only velocities change): a contact
// element depends on its four vertices only, so the elements of the previous pass whose vertices all kept their
// velocity are carried over; everything else is re-tested by the dirty-restricted traversal.  Carried-over and
// re-found elements may coincide (same key, same DFS rank): the canonical set removes the duplicates.
__global__ void k_mark_dirty(int np, const double* __restrict__ uvw, double* __restrict__ snap, int* __restrict__ vdirty,
                             int* cnt) {
  const int s = blockIdx.y;
  if (!cnt[s * NCNT + CNT_ACTIVE]) return;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= np) return;
  if (v == 0) cnt[s * NCNT + CNT_NDIRTY] = 0;  // the incremental refit that follows rebuilds the dirty-leaf list
  const size_t o = ((size_t)s * np + v) * 3;
  int d = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double a = uvw[o + k];
    d |= __double_as_longlong(a) != __double_as_longlong(snap[o + k]);
    snap[o + k] = a;
  }
  vdirty[(size_t)s * np + v] = d;
}
__global__ void k_keep_clean(int np, const RawElem* __restrict__ prev, long long cap_raw, RawElem* __restrict__ raw,
                             const int* __restrict__ vdirty, int* cnt) {
  const int s = blockIdx.y;
  int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int n = min((long long)c[CNT_RAWPREV], cap_raw);
  const int* vd = vdirty + (size_t)s * np;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const RawElem r = prev[(size_t)s * cap_raw + i];
    if (vd[r.n[0]] | vd[r.n[1]] | vd[r.n[2]] | vd[r.n[3]]) continue;
    if (i > 0) {
      // prev is sorted by (key, DFS rank): an exact duplicate (a carried-over element that the dirty-restricted
      // traversal emitted again through a third dirty vertex of its leaf pair) sits right behind its twin; carrying
      // both would let the list grow pass after pass
      const RawElem q = prev[(size_t)s * cap_raw + i - 1];
      if (q.n[0] == r.n[0] && q.n[1] == r.n[1] && q.n[2] == r.n[2] && q.n[3] == r.n[3] && q.pre == r.pre &&
          q.path_hi == r.path_hi && q.path_lo == r.path_lo && q.slot == r.slot)
        continue;
    }
    const int k = atomicAdd(&c[CNT_RAW], 1);
    if (k >= cap_raw) { atomicOr(&c[CNT_OVERFLOW], 4); continue; }
    raw[(size_t)s * cap_raw + k] = r;
  }
}

// ------------------------------------------------------------------------------------------------ canonical set
struct RawLess {
  __device__ __forceinline__ bool operator()(const RawElem& a, const RawElem& b) const {
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (a.n[k] != b.n[k]) return a.n[k] < b.n[k];
    if (a.pre != b.pre) return a.pre < b.pre;
    if (a.path_hi != b.path_hi) return a.path_hi < b.path_hi;
    if (a.path_lo != b.path_lo) return a.path_lo < b.path_lo;
    return a.slot < b.slot;
  }
};

__global__ void k_mark_unique(int n, const RawElem* __restrict__ raw, int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int f = 1;
  if (i > 0) {
    const RawElem& a = raw[i - 1];
    const RawElem& b = raw[i];
    f = (a.n[0] != b.n[0] || a.n[1] != b.n[1] || a.n[2] != b.n[2] || a.n[3] != b.n[3]) ? 1 : 0;
  }
  flag[i] = f;
}
__global__ void k_compact_elems(int n, const RawElem* __restrict__ raw, const int* __restrict__ flag,
                                const int* __restrict__ pos, int* __restrict__ elems, long long cap_elem, int* cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag[i]) {
    const int k = pos[i];
    if (k < cap_elem) {
      elems[k * 5 + 0] = raw[i].n[0]; elems[k * 5 + 1] = raw[i].n[1];
      elems[k * 5 + 2] = raw[i].n[2]; elems[k * 5 + 3] = raw[i].n[3];
      elems[k * 5 + 4] = raw[i].slot < 6 ? 1 : 0;
    } else {
      atomicOr(&cnt[CNT_OVERFLOW], 8);
    }
  }
  if (i == n - 1) cnt[CNT_ELEM] = pos[i] + flag[i];
}

// The canonical set of a SMALL raw list (most incremental passes: a few hundred to a few thousand raw elements) in one
// launch for all scenes: one CTA per scene sorts its list in shared memory (bitonic network on the 40-byte records,
// same comparison as the library sort), marks the first record of every key, scans and compacts -- instead of a
// library merge sort, a flag kernel, a library scan and a compaction kernel per scene.  The sorted list is written
// back (it is the "previous" list of the next incremental pass).  Scenes with more than CANON_SMALL raw elements are
// left to the library path.
enum { CANON_SMALL = 2048, CANON_BIG = 32768, CANON_T = 1024 };  // (a 4096-record list takes 176 us here: no faster than the library path)
// In batches (limit > CANON_SMALL) lists of up to `limit` elements are sorted by the same CTA in place in global
// memory: slower per scene than the library sort, but all scenes run at once instead of eight launches per scene.
__global__ void __launch_bounds__(CANON_T) k_canon_small(int limit, long long cap_raw, long long cap_elem, RawElem* __restrict__ raw_all,
                                                          int* __restrict__ elems_all, int* cnt) {
  extern __shared__ __align__(16) unsigned char canon_sm[];
  __shared__ int wsum[CANON_T / 32];
  __shared__ int carry;
  const int s = blockIdx.x;
  int* c = cnt + s * NCNT;
  if (!c[CNT_ACTIVE]) return;
  const int n = c[CNT_RAW];
  if (n > limit) return;  // the library path handles this scene
  if (n == 0) { if (threadIdx.x == 0) c[CNT_ELEM] = 0; return; }
  RawElem* raw = raw_all + (size_t)s * cap_raw;
  int* elems = elems_all + (size_t)s * cap_elem * 5;
  int N = 1;
  while (N < n) N <<= 1;
  const bool in_smem = N <= CANON_SMALL;
  if (!in_smem && N > cap_raw) return;  // (no room for the padding: library path; the host applies the same rule)
  RawElem* sm = in_smem ? reinterpret_cast<RawElem*>(canon_sm) : raw;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    if (i < n) { if (in_smem) sm[i] = raw[i]; }
    else { RawElem r; r.n[0] = r.n[1] = r.n[2] = r.n[3] = INT_MAX; r.pre = 0xffffffffu; r.slot = 0xffffffffu; r.path_hi = ~0ULL; r.path_lo = ~0ULL; sm[i] = r; }
  }
  __syncthreads();
  const RawLess less;
  for (int k = 2; k <= N; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const RawElem a = sm[i], b = sm[l];
          const bool up = (i & k) == 0;
          if (up ? less(b, a) : less(a, b)) { sm[i] = b; sm[l] = a; }
        }
      }
      __syncthreads();
    }
  }
  if (in_smem)
    for (int i = threadIdx.x; i < n; i += blockDim.x) raw[i] = sm[i];
  // first record of every key -> compact (std::set keeps the first inserted of equal keys: smallest DFS rank)
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int f = 0;
    if (i < n) {
      f = 1;
      if (i > 0) {
        const RawElem& a = sm[i - 1];
        const RawElem& b = sm[i];
        f = (a.n[0] != b.n[0] || a.n[1] != b.n[1] || a.n[2] != b.n[2] || a.n[3] != b.n[3]) ? 1 : 0;
      }
    }
    int incl = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; w++) woff += wsum[w];
    const int pos = carry + woff + incl - f;
    if (f) {
      if (pos < cap_elem) {
        const RawElem& r = sm[i];
        elems[pos * 5 + 0] = r.n[0]; elems[pos * 5 + 1] = r.n[1]; elems[pos * 5 + 2] = r.n[2]; elems[pos * 5 + 3] = r.n[3];
        elems[pos * 5 + 4] = r.slot < 6 ? 1 : 0;
      } else {
        atomicOr(&c[CNT_OVERFLOW], 8);
      }
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = pos + f;
    __syncthreads();
  }
  if (threadIdx.x == 0) c[CNT_ELEM] = carry;
}

__global__ void k_set_active(int* cnt, int v) { cnt[blockIdx.x * NCNT + CNT_ACTIVE] = v; }

// scenes whose fused traversal (one CTA per SM and scene) is co-resident: resident CTAs of k_trav_all per SM
int trav_all_max_scenes() {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trav_all, 256, 0) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return nb;
}

int launch_refit(Ctx* c, const double* xyz, const double* uvw, double par, int mode, bool incremental) {
  // mode 1 keeps a snapshot of the velocities the boxes were built from; an incremental refit first marks the
  // vertices whose velocity changed since then and propagates "dirty" up the tree together with the boxes
  if (mode == 1 && incremental) {
    dim3 gd(div_up(c->np, 256), c->B);
    k_mark_dirty<<<gd, 256, 0, KS(c, "k_mark_dirty")>>>(c->np, uvw, c->uvw_snap, c->vdirty, c->cnt);
    CG_LAUNCH_CHECK(c);
  } else if (mode == 1) {
    CG_CUDA(cudaMemcpyAsync(c->uvw_snap, uvw, sizeof(double) * (size_t)c->B * c->ndof, cudaMemcpyDeviceToDevice, c->stream));
  }
  CG_CUDA(cudaMemsetAsync(c->refit_flag, 0, sizeof(int) * (size_t)c->B * c->nnode, c->stream));
  dim3 g(div_up(c->nnode, 256), c->B);
  const bool inc = mode == 1 && incremental;
  if (inc && c->inc_refit) {
    // only the dirty leaves and their ancestors are rebuilt; every other box is the one of the previous pass
    CG_CUDA(cudaMemsetAsync(c->ndirty, 0, sizeof(int) * (size_t)c->B * c->nnode, c->stream));
    CG_CUDA(cudaMemsetAsync(c->refit_nchild, 0, sizeof(int) * (size_t)c->B * c->nnode, c->stream));
    k_refit_mark<<<g, 256, 0, KS(c, "k_refit_mark")>>>(c->nnode, c->np, c->nodes, c->tri, c->vdirty, c->ndirty, c->refit_nchild, c->cnt);
    CG_LAUNCH_CHECK(c);
    k_refit_inc<<<g, 256, 0, KS(c, "k_refit_inc")>>>(c->nnode, c->np, c->nodes, c->tri, xyz, uvw, par, c->boxes, c->refit_flag, c->cnt,
                                             c->vdirty, c->refit_nchild, c->dirty_list);
    CG_LAUNCH_CHECK(c);
    return CLOTHGPU_OK;
  }
  k_refit<<<g, 256, 0, KS(c, "k_refit")>>>(c->nnode, c->np, c->nodes, c->tri, xyz, uvw, par, mode, c->boxes, c->refit_flag, c->cnt,
                                   inc ? c->vdirty : nullptr, inc ? c->ndirty : nullptr);
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}

static int traverse_and_test(Ctx* c, const double* xyz, const double* uvw, double dt, double delta, int ccd,
                             bool incremental = false) {
  const int* nd = incremental ? c->ndirty : nullptr;
  k_trav_reset<<<c->B, 64, 0, KS(c, "k_trav_reset")>>>(c->cnt);
  CG_LAUNCH_CHECK(c);
  if (incremental) {  // carry over the previous pass's elements whose vertices did not change
    dim3 gk(148, c->B);
    k_keep_clean<<<gk, 256, 0, KS(c, "k_keep_clean")>>>(c->np, c->raw_sorted, c->cap_raw, c->raw, c->vdirty, c->cnt);
    CG_LAUNCH_CHECK(c);
  }
  const bool query = incremental && c->query_trav && c->inc_refit;  // (the incremental refit lists the dirty leaves)
  const bool coop = !query && incremental && c->trav_coop > 0 && c->B <= c->trav_coop && !c->ktiming;
  // 8-byte frontier / leaf entries wherever the per-level kernels or the leaf query produce them (the cooperative
  // kernel and the rank cross-check keep the 32-byte entries that carry the depth-first rank)
  const bool slim = c->slim_frontier && !c->rank_check && !coop;
  c->leaf_slim = slim;
  if (query) {
    dim3 g(std::max(8, 148 * 8 / c->B), c->B);
    k_query<<<g, QRY_WARPS * 32, 0, KS(c, "k_query")>>>(c->nnode, c->nodes, c->boxes, c->npath, nd, c->dirty_list, c->leaf, c->cap_leaf,
                                                c->cnt, std::min((int)QRY_CAP, c->query_cap), slim ? 1 : 0);
    CG_LAUNCH_CHECK(c);
  } else {
    dim3 g(div_up(c->ninternal, 256), c->B);
    if (slim) k_trav_init<true><<<g, 256, 0, KS(c, "k_trav_init")>>>(c->ninternal, c->internal_list, c->nodes, c->pre, c->frontA, c->cap_front, c->cnt,
                                                             c->nnode, nd);
    else k_trav_init<false><<<g, 256, 0, KS(c, "k_trav_init")>>>(c->ninternal, c->internal_list, c->nodes, c->pre, c->frontA, c->cap_front, c->cnt,
                                                          c->nnode, nd);
    CG_LAUNCH_CHECK(c);
  }
  // every descent moves at least one side of the pair one level down: bvh_height + 1 steps reach all leaf pairs
  bool fused = query;
  // (used for the incremental passes only: their frontiers are tiny and nothing else runs beside them, while the
  // full passes overlap the solve on the second stream, where a cooperative launch would have to wait for room)
  if (coop) {
    int ndepth = c->bvh_height + 1, nnode = c->nnode;
    long long capf = c->cap_front, capl = c->cap_leaf;
    const int* nodes = c->nodes; const double* boxes = c->boxes; PairEnt* fa = c->frontA; PairEnt* fb = c->frontB;
    PairEnt* lf = c->leaf; int* cnt = c->cnt;
    void* args[] = {&ndepth, &nnode, &nodes, &boxes, &fa, &fb, &capf, &lf, &capl, &cnt, &nd};
    dim3 g(c->trav_sms, c->B);
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_trav_all, g, dim3(256), args, 0, c->stream);
    if (e == cudaSuccess) {
      c->launches++;
      fused = true;
    } else {
      cudaGetLastError();  // e.g. too many CTAs for co-residency: fall back to one launch per level
      c->trav_coop = 0;
    }
  }
  for (int d = 0; d <= c->bvh_height && !fused; d++) {
    dim3 g(incremental ? 74 : 148 * 4, c->B);  // incremental frontiers are small: a small grid launches faster
    PairEnt* fin = (d & 1) ? c->frontB : c->frontA;
    PairEnt* fout = (d & 1) ? c->frontA : c->frontB;
    if (slim) k_trav_step<true><<<g, 256, 0, KS(c, "k_trav_step")>>>(d, c->nnode, c->nodes, c->boxes, fin, fout, c->cap_front, c->leaf, c->cap_leaf, c->cnt, nd);
    else k_trav_step<false><<<g, 256, 0, KS(c, "k_trav_step")>>>(d, c->nnode, c->nodes, c->boxes, fin, fout, c->cap_front, c->leaf, c->cap_leaf, c->cnt, nd);
    CG_LAUNCH_CHECK(c);
  }
  {
    dim3 g(148 * 8 * (256 / CULL_T), c->B);
    k_cull<<<g, CULL_T, 0, KS(c, "k_cull")>>>(ccd, c->nnode, c->np, (const int4*)c->node_tri, c->boxes, xyz, uvw, dt, delta, c->leaf,
                                   c->cap_leaf, c->cand, c->cap_leaf * 2, c->cnt, slim ? 1 : 0);
    CG_LAUNCH_CHECK(c);
    dim3 g2(148 * 4, c->B);
    k_narrow<<<g2, 128, 0, KS(c, "k_narrow")>>>(ccd, c->nnode, c->np, (const int4*)c->node_tri, c->boxes, xyz, uvw, dt, delta, c->leaf,
                                        c->cap_leaf, c->cand, c->cap_leaf * 2, c->raw, c->cap_raw, c->cnt, c->nodes, c->pre, c->npath,
                                        c->ndepth, c->rank_check ? 1 : 0, slim ? 1 : 0);
    CG_LAUNCH_CHECK(c);
  }
  return CLOTHGPU_OK;
}

// raw elements -> std::set order with first-insert-wins duplicates (per scene; counts come back to the host, which
// needs them anyway for the loop control of self_collision_cloth.cpp:416,449)
static int canonical_set(Ctx* c, int s, int nraw) {
  int* cnt = c->cnt + s * NCNT;
  if (nraw == 0) {
    CG_CUDA(cudaMemsetAsync(cnt + CNT_ELEM, 0, sizeof(int), c->stream));
    return CLOTHGPU_OK;
  }
  RawElem* raw = c->raw + (size_t)s * c->cap_raw;
  size_t tb = 0, tb2 = 0;
  CG_CUDA(cub::DeviceMergeSort::SortKeys(nullptr, tb, raw, nraw, RawLess(), c->stream));
  CG_CHECK(ensure_sort_capacity(c, nraw));
  int* flag = c->sort_vals;
  int* pos = c->sort_vals2;
  CG_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb2, flag, pos, nraw, c->stream));
  CG_CHECK(ensure_scratch(c, std::max(tb, tb2)));
  CG_CUDA(cub::DeviceMergeSort::SortKeys(c->scratch, tb, raw, nraw, RawLess(), c->stream));
  c->launches += 2;
  k_mark_unique<<<div_up(nraw, 256), 256, 0, KS(c, "k_mark_unique")>>>(nraw, raw, flag);
  CG_LAUNCH_CHECK(c);
  CG_CUDA(cub::DeviceScan::ExclusiveSum(c->scratch, tb2, flag, pos, nraw, c->stream));
  c->launches += 2;
  k_compact_elems<<<div_up(nraw, 256), 256, 0, KS(c, "k_compact_elems")>>>(nraw, raw, flag, pos, c->elems + (size_t)s * c->cap_elem * 5,
                                                            c->cap_elem, cnt);
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}

// GetContactElement_Proximity / _CCD for the scenes whose CNT_ACTIVE flag is set; counts[B] out (0 for inactive)
// the device side of a detection pass without any host synchronisation (used to run the proximity detection of a
// step on a second stream, concurrently with the implicit solve)
int launch_contacts_async(Ctx* c, const double* xyz, const double* uvw, double dt, double delta, int ccd) {
  return traverse_and_test(c, xyz, uvw, dt, delta, ccd);
}

int run_contacts(Ctx* c, const double* xyz, const double* uvw, double dt, double delta, int ccd, int* counts,
                 bool prelaunched, bool incremental) {
  if (c->nnode == 0) { set_error("context has no BVH"); return CLOTHGPU_ERR_STATE; }
  for (int attempt = 0;; attempt++) {
    phase_begin(c, 4);
    // a retry after a queue overflow is always a full pass (the previous raw list does not survive the regrowth)
    if (!(prelaunched && attempt == 0)) CG_CHECK(traverse_and_test(c, xyz, uvw, dt, delta, ccd, incremental && attempt == 0));
    // small raw lists (most incremental passes) are sorted / compacted right away, one launch for all scenes, before
    // the counters come back: such a pass costs one host round trip, not two.  The kernel skips scenes whose list is
    // longer than CANON_SMALL; those go through the library sort below.
    const int canon_limit = !c->canon_small ? -1 : (c->B >= 4 ? (int)std::min<long long>(CANON_BIG, c->cap_raw / 2) : CANON_SMALL);
    if (c->canon_small) {
      const int smem = CANON_SMALL * (int)sizeof(RawElem);
      CG_CUDA(cudaFuncSetAttribute(k_canon_small, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(smem, 48 * 1024)));
      k_canon_small<<<c->B, CANON_T, smem, KS(c, "k_canon_small")>>>(canon_limit, c->cap_raw, c->cap_elem, c->raw, c->elems, c->cnt);
      CG_LAUNCH_CHECK(c);
    }
    c->canon_limit = canon_limit;
    CG_CHECK(fetch_counters(c));
    phase_end(c, 4);
    int bits = 0;
    for (int s = 0; s < c->B; s++) bits |= c->h_cnt[s * NCNT + CNT_OVERFLOW];
    if ((bits & ~8) == 0) break;  // (bit 3, the element buffer, is reported below)
    if (attempt >= 12) { set_error("contact work queues keep overflowing"); return CLOTHGPU_ERR_CAPACITY; }
    CG_CHECK(grow_queues(c, bits));
  }
  phase_begin(c, 5);
  {
    bool any_large = false;
    for (int s = 0; s < c->B; s++) {
      if (!c->h_cnt[s * NCNT + CNT_ACTIVE]) continue;
      const int nraw = c->h_cnt[s * NCNT + CNT_RAW];
      if (nraw <= c->canon_limit) continue;
      CG_CHECK(canonical_set(c, s, nraw));
      any_large = true;
    }
    if (any_large) CG_CHECK(fetch_counters(c));
  }
  phase_end(c, 5);
  for (int s = 0; s < c->B; s++) {
    if (c->h_cnt[s * NCNT + CNT_OVERFLOW] & 8) { set_error("contact element buffer overflow"); return CLOTHGPU_ERR_CAPACITY; }
    if (counts) counts[s] = c->h_cnt[s * NCNT + CNT_ACTIVE] ? c->h_cnt[s * NCNT + CNT_ELEM] : 0;
  }
  // the (sorted, not yet de-duplicated) raw list of this pass is the "previous" list of the next incremental pass
  std::swap(c->raw, c->raw_sorted);
  return CLOTHGPU_OK;
}

int set_all_active(Ctx* c, int v) {
  k_set_active<<<c->B, 1, 0, KS(c, "k_set_active")>>>(c->cnt, v);
  CG_LAUNCH_CHECK(c);
  return CLOTHGPU_OK;
}

}  // namespace cg
