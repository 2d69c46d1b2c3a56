// csrc/bvhbuild.cu -- BVH *topology* builder (host code, setup only: runs once per mesh, never per step).
//
// Replaces MakeBVHTopology_TopDown (bvh_aabb.cpp:294-313) = MakeTriSurTri (:109-138) + DevideTriAryConnex (:156-290).
// The per-step path depends on the node numbering (the face-vertex proximity test reads the box of node number
// *triangle id*, bvh_aabb.cpp:490-505), so a drop-in has to hand out the very same CNodeBVH array: the splitting rule,
// the order of every list, the node numbering and the floating-point expressions below reproduce the reference's
// (host fp64, no contraction), which tests/test_bvh_builder.py checks node for node against the compiled reference.
//
// Algorithm: a node's triangles are split into (child 0) the edge-connected set grown from a "kernel" triangle
// through triangles on the positive side of a plane (centre of the centroid box, normal = its longest axis; a
// covariance power iteration when no axis is strictly longest) and (child 1) the rest.  Children get the next two
// node ids at split time; child 0 is split completely before child 1 (explicit stack instead of recursion, so deep
// trees of 512^2 meshes need no large call stack).
//
// Where the reference crashes (bvh_aabb.cpp:238 writes aTri2Node[-1]: on exactly square sub-grids the extents tie,
// the power iteration started from (1,1,1) can die on anti-diagonal centroid pairs, and one child comes out empty --
// SURVEY.md section 8c) this builder keeps going with a deterministic rule: longest axis with ties resolved x, y, z;
// if that still leaves a child empty the first triangle is split off.  The reference has no result there, so no
// parity is lost; on every mesh the reference survives the two trees are identical.
#include <cmath>
#include <cstring>
#include <vector>

#include "ctx.h"

namespace cg {

namespace {

struct V3 {
  double x, y, z;
};
inline V3 sub(const V3& a, const V3& b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct Builder {
  int ntri, np;
  const int* tri;
  const double* xyz;
  std::vector<int> sur;       // ntri x 3: triangle across each edge, -1 on the boundary
  std::vector<int> tri2node;  // node each triangle currently belongs to
  std::vector<int> nodes;     // n x {iroot, ichild0, ichild1}
  int nfallback = 0;

  V3 centroid(int t) const {
    const int a = tri[t * 3], b = tri[t * 3 + 1], c = tri[t * 3 + 2];
    return V3{(xyz[a * 3] + xyz[b * 3] + xyz[c * 3]) / 3.0, (xyz[a * 3 + 1] + xyz[b * 3 + 1] + xyz[c * 3 + 1]) / 3.0,
              (xyz[a * 3 + 2] + xyz[b * 3 + 2] + xyz[c * 3 + 2]) / 3.0};
  }

  // triangle across every edge (edge k is opposite to corner k): the first triangle around the edge's first vertex, in
  // ascending triangle order, that holds the edge reversed
  void adjacency() {
    std::vector<int> ptr(np + 1, 0), ent((size_t)ntri * 3);
    for (int i = 0; i < ntri * 3; i++) ptr[tri[i] + 1]++;
    for (int i = 0; i < np; i++) ptr[i + 1] += ptr[i];
    {
      std::vector<int> fill(ptr.begin(), ptr.end() - 1);
      for (int t = 0; t < ntri; t++)
        for (int k = 0; k < 3; k++) ent[fill[tri[t * 3 + k]]++] = t;
    }
    sur.assign((size_t)ntri * 3, -1);
    for (int t = 0; t < ntri; t++)
      for (int e = 0; e < 3; e++) {
        const int a = tri[t * 3 + (e + 1) % 3], b = tri[t * 3 + (e + 2) % 3];
        for (int q = ptr[a]; q < ptr[a + 1] && sur[t * 3 + e] < 0; q++) {
          const int u = ent[q];
          if (u == t) continue;
          for (int f = 0; f < 3; f++)
            if (tri[u * 3 + (f + 1) % 3] == b && tri[u * 3 + (f + 2) % 3] == a) {
              sur[t * 3 + e] = u;
              break;
            }
        }
      }
  }

  // first triangle of the list that does not lie on the plane; the normal is flipped so that it lies on the
  // positive side
  int find_kernel(const std::vector<int>& list, const V3& org, V3& dir) const {
    for (int t : list) {
      const double d = dot(sub(centroid(t), org), dir);
      if (std::fabs(d) < 1.0e-10) continue;
      if (d < 0) { dir.x *= -1; dir.y *= -1; dir.z *= -1; }
      return t;
    }
    return -1;
  }

  // grow child 0 from the kernel; returns its triangles in discovery order
  void flood(int ker, int node, int ch0, const V3& org, const V3& dir, std::vector<int>& out) {
    out.clear();
    std::vector<int> stack;
    tri2node[ker] = ch0;
    out.push_back(ker);
    stack.push_back(ker);
    while (!stack.empty()) {
      const int t = stack.back();
      stack.pop_back();
      for (int e = 0; e < 3; e++) {
        const int u = sur[t * 3 + e];
        if (u < 0 || tri2node[u] != node) continue;
        if (dot(sub(centroid(u), org), dir) < 0) continue;
        stack.push_back(u);
        tri2node[u] = ch0;
        out.push_back(u);
      }
    }
  }

  void split(int node, const std::vector<int>& list, std::vector<int>& l0, std::vector<int>& l1) {
    const double eps = 1.0e-10;
    double mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    bool first = true;
    for (int t : list) {
      const V3 g = centroid(t);
      const double p[3] = {g.x, g.y, g.z};
      for (int d = 0; d < 3; d++) {
        const double lo = p[d] - eps, hi = p[d] + eps;
        if (first) { mn[d] = lo; mx[d] = hi; }
        else { mn[d] = mn[d] < lo ? mn[d] : lo; mx[d] = mx[d] > hi ? mx[d] : hi; }
      }
      first = false;
    }
    const double lx = mx[0] - mn[0], ly = mx[1] - mn[1], lz = mx[2] - mn[2];
    V3 dir{0, 0, 0};
    if (lx > ly && lx > lz) dir = V3{1, 0, 0};
    if (ly > lz && ly > lx) dir = V3{0, 1, 0};
    if (lz > lx && lz > ly) dir = V3{0, 0, 1};
    const V3 centre{(mn[0] + mx[0]) * 0.5, (mn[1] + mx[1]) * 0.5, (mn[2] + mx[2]) * 0.5};
    V3 org = centre;
    int ker = find_kernel(list, org, dir);
    if (ker < 0) {  // no strictly longest axis: dominant direction of the centroid covariance, ten power iterations
      V3 s{0, 0, 0};
      for (int t : list) { const V3 g = centroid(t); s.x += g.x; s.y += g.y; s.z += g.z; }
      const double n = (double)list.size();
      org = V3{s.x / n, s.y / n, s.z / n};
      double m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (int t : list) {
        const V3 w = sub(centroid(t), org);
        const double a[3] = {w.x, w.y, w.z};
        for (int i = 0; i < 3; i++)
          for (int j = 0; j < 3; j++) m[i][j] += a[i] * a[j];
      }
      double d[3] = {1, 1, 1};
      for (int it = 0; it < 10; it++) {
        const double t0 = m[0][0] * d[0] + m[0][1] * d[1] + m[0][2] * d[2];
        const double t1 = m[1][0] * d[0] + m[1][1] * d[1] + m[1][2] * d[2];
        const double t2 = m[2][0] * d[0] + m[2][1] * d[1] + m[2][2] * d[2];
        const double len = std::sqrt(t0 * t0 + t1 * t1 + t2 * t2);
        d[0] = t0 / len; d[1] = t1 / len; d[2] = t2 / len;
      }
      dir = V3{d[0], d[1], d[2]};
      ker = find_kernel(list, org, dir);
    }
    const int ch0 = (int)nodes.size() / 3, ch1 = ch0 + 1;
    auto rest = [&]() {
      l1.clear();
      for (int t : list)
        if (tri2node[t] != ch0) l1.push_back(t);
    };
    if (ker >= 0) {
      flood(ker, node, ch0, org, dir, l0);
      rest();
    }
    if (ker < 0 || l1.empty()) {
      // the reference has no result here (out-of-bounds write / empty child): deterministic continuation
      nfallback++;
      for (int t : list) tri2node[t] = node;
      int ax = 0;
      const double len[3] = {lx, ly, lz};
      for (int d = 1; d < 3; d++)
        if (len[d] > len[ax]) ax = d;
      dir = V3{ax == 0 ? 1.0 : 0.0, ax == 1 ? 1.0 : 0.0, ax == 2 ? 1.0 : 0.0};
      org = centre;
      ker = find_kernel(list, org, dir);
      if (ker >= 0) {
        flood(ker, node, ch0, org, dir, l0);
        rest();
      }
      if (ker < 0 || l1.empty()) {  // coincident centroids: split off the first triangle
        for (int t : list) tri2node[t] = node;
        l0.assign(1, list[0]);
        tri2node[list[0]] = ch0;
        rest();
      }
    }
    for (int t : l1) tri2node[t] = ch1;
    nodes.resize(nodes.size() + 6, -1);
    nodes[ch0 * 3] = node;
    nodes[ch1 * 3] = node;
    nodes[node * 3 + 1] = ch0;
    nodes[node * 3 + 2] = ch1;
  }

  void run() {
    adjacency();
    tri2node.assign(ntri, 0);
    nodes.assign(3, -1);
    std::vector<int> all(ntri);
    for (int t = 0; t < ntri; t++) all[t] = t;
    if (ntri == 1) {  // (the reference asserts list.size() > 1; a single triangle is a single leaf)
      nodes[1] = 0;
      return;
    }
    struct Item {
      int node;
      std::vector<int> list;
    };
    std::vector<Item> todo;
    todo.push_back(Item{0, std::move(all)});
    std::vector<int> l0, l1;
    while (!todo.empty()) {
      Item it = std::move(todo.back());
      todo.pop_back();
      if (it.list.size() == 1) {
        nodes[it.node * 3 + 1] = it.list[0];
        nodes[it.node * 3 + 2] = -1;
        continue;
      }
      split(it.node, it.list, l0, l1);
      const int ch0 = nodes[it.node * 3 + 1], ch1 = nodes[it.node * 3 + 2];
      todo.push_back(Item{ch1, l1});  // child 0 is split completely first
      todo.push_back(Item{ch0, l0});
    }
  }
};

}  // namespace

int bvh_build_topdown_host(int ntri, const int* tri, int np, const double* xyz, std::vector<int>& nodes, int* nfallback) {
  if (ntri < 1 || np < 1 || !tri || !xyz) { set_error("bvh build: bad argument"); return CLOTHGPU_ERR_ARG; }
  for (int i = 0; i < ntri * 3; i++)
    if (tri[i] < 0 || tri[i] >= np) { set_error("bvh build: triangle index out of range"); return CLOTHGPU_ERR_ARG; }
  Builder b;
  b.ntri = ntri; b.np = np; b.tri = tri; b.xyz = xyz;
  b.run();
  nodes.swap(b.nodes);
  if (nfallback) *nfallback = b.nfallback;
  return CLOTHGPU_OK;
}

}  // namespace cg

extern "C" int clothgpu_bvh_build_topdown(int ntri, const int* tri, int np, const double* xyz, int* nodes, int cap_nodes,
                                          int* nfallback) {
  std::vector<int> out;
  const int rc = cg::bvh_build_topdown_host(ntri, tri, np, xyz, out, nfallback);
  if (rc != CLOTHGPU_OK) return -rc;
  const int n = (int)out.size() / 3;
  if (!nodes || cap_nodes < n) { cg::set_error("bvh build: node buffer too small (needs 2*ntri-1 nodes)"); return -CLOTHGPU_ERR_ARG; }
  memcpy(nodes, out.data(), sizeof(int) * out.size());
  return n;
}
