"""Synthetic grid-cloth scenes: the inputs the reference demos build in their main() functions.

These mirror the *caller contract* of the hot path (mesh, hinge quads, sparsity pattern, vertex adjacency,
constants), not the hot path itself:

* ``square_cloth_xz``  -- SetClothShape_Square of self_contact_sparse/main.cpp:31-106 (cloth in the y=0 plane)
* ``square_cloth_xy``  -- SetClothShape_Square of internal_cloth_sparse/main.cpp:31-103 (z=0 plane, row ix=0 pinned)
* ``pattern_from_elems`` -- CJaggedArray::SetEdgeOfElem(+Sort) as used at self_contact_sparse/main.cpp:575-580
* ``SELF_CONTACT`` / ``INTERNAL_CLOTH`` -- the compile-time constants of the two demos (main.cpp:138-156 / 137-154)

A scene is a plain dict of contiguous numpy arrays (int32 / float64), consumed by the ctypes bindings.
"""
import numpy as np


def _grid_topology(ndiv_x, ndiv_z):
    """Triangles and hinge quads of an (ndiv_x+1) x (ndiv_z+1) vertex grid, vertex id = ix*(ndiv_z+1)+iz."""
    nz1 = ndiv_z + 1
    ix, iz = np.meshgrid(np.arange(ndiv_x), np.arange(ndiv_z), indexing="ij")
    ix = ix.ravel()
    iz = iz.ravel()
    v00 = ix * nz1 + iz
    v10 = (ix + 1) * nz1 + iz
    v01 = ix * nz1 + iz + 1
    v11 = (ix + 1) * nz1 + iz + 1
    # two triangles per cell, in cell order
    tri = np.stack([v00, v10, v01, v11, v01, v10], axis=1).reshape(-1, 3)
    # hinge across the cell diagonal
    q0 = np.stack([v00, v11, v10, v01], axis=1)
    # hinge across the edge shared with the next cell in z
    ix, iz = np.meshgrid(np.arange(ndiv_x), np.arange(ndiv_z - 1), indexing="ij")
    ix = ix.ravel()
    iz = iz.ravel()
    q1 = np.stack([(ix + 1) * nz1 + iz, ix * nz1 + iz + 2, (ix + 1) * nz1 + iz + 1, ix * nz1 + iz + 1], axis=1)
    # hinge across the edge shared with the next cell in x
    ix, iz = np.meshgrid(np.arange(ndiv_x - 1), np.arange(ndiv_z), indexing="ij")
    ix = ix.ravel()
    iz = iz.ravel()
    q2 = np.stack([ix * nz1 + iz + 1, (ix + 2) * nz1 + iz, (ix + 1) * nz1 + iz, (ix + 1) * nz1 + iz + 1], axis=1)
    quad = np.concatenate([q0, q1, q2], axis=0)
    return np.ascontiguousarray(tri, dtype=np.int32), np.ascontiguousarray(quad, dtype=np.int32)


def pattern_from_elems(elem, nnode):
    """Vertex->vertex adjacency (no self) from elements, ascending per row: (index[nnode+1], array[nnz])."""
    elem = np.asarray(elem)
    n = elem.shape[1]
    rows = np.repeat(elem, n, axis=1).ravel()
    cols = np.tile(elem, (1, n)).ravel()
    keep = rows != cols
    key = np.unique(rows[keep].astype(np.int64) * nnode + cols[keep].astype(np.int64))
    r = (key // nnode).astype(np.int32)
    c = (key % nnode).astype(np.int32)
    index = np.zeros(nnode + 1, dtype=np.int32)
    np.add.at(index, r + 1, 1)
    index = np.cumsum(index, dtype=np.int64).astype(np.int32)
    return index, np.ascontiguousarray(c)


def _finish(xyz0, bc, tri, quad, total_area, areal_density):
    np_ = xyz0.shape[0]
    rowptr, colidx = pattern_from_elems(quad, np_)
    eidx, earr = pattern_from_elems(tri, np_)
    return {
        "np": int(np_), "ntri": int(tri.shape[0]), "nquad": int(quad.shape[0]),
        "xyz0": np.ascontiguousarray(xyz0, dtype=np.float64).ravel(),
        "bc": np.ascontiguousarray(bc, dtype=np.int32),
        "tri": tri.ravel(), "quad": quad.ravel(),
        "rowptr": rowptr, "colidx": colidx,
        "edge_index": eidx, "edge_array": earr,
        "mass_point": total_area * areal_density / float(np_),
        "bvh": None,
    }


def square_cloth_xz(ndiv_x, ndiv_z, h, areal_density=0.04):
    """self_contact_sparse scene: vertices (ix*h, 0, iz*h), nothing pinned."""
    tri, quad = _grid_topology(ndiv_x, ndiv_z)
    ix, iz = np.meshgrid(np.arange(ndiv_x + 1), np.arange(ndiv_z + 1), indexing="ij")
    xyz0 = np.stack([ix.ravel() * h, np.zeros(ix.size), iz.ravel() * h], axis=1)
    bc = np.zeros(xyz0.shape[0], dtype=np.int32)
    sc = _finish(xyz0, bc, tri, quad, (ndiv_x * h) * (ndiv_z * h), areal_density)
    sc["ndiv"] = (ndiv_x, ndiv_z)
    sc["h"] = h
    return sc


def square_cloth_xy(ndiv, cloth_size, areal_density=1.0):
    """internal_cloth_sparse scene: vertices (ix*h, iy*h, 0) with h = cloth_size/ndiv, row ix=0 pinned."""
    h = cloth_size / ndiv
    tri, quad = _grid_topology(ndiv, ndiv)
    ix, iy = np.meshgrid(np.arange(ndiv + 1), np.arange(ndiv + 1), indexing="ij")
    xyz0 = np.stack([ix.ravel() * h, iy.ravel() * h, np.zeros(ix.size)], axis=1)
    bc = np.zeros(xyz0.shape[0], dtype=np.int32)
    bc[: ndiv + 1] = 1
    sc = _finish(xyz0, bc, tri, quad, cloth_size * cloth_size, areal_density)
    sc["ndiv"] = (ndiv, ndiv)
    sc["h"] = h
    return sc


def jittered_rest(sc, amplitude_frac=0.2, seed=88172645463325252):
    """Rest positions with a deterministic in-plane jitter (xorshift64), used only to feed a BVH *topology*
    builder: the reference's own builder crashes on exactly-square grids."""
    x = np.array(sc["xyz0"], dtype=np.float64).reshape(-1, 3).copy()
    s = np.uint64(seed)
    amp = amplitude_frac * sc["h"]
    planar = [a for a in range(3) if np.ptp(x[:, a]) > 0]
    m64 = (1 << 64) - 1
    state = int(s)
    out = np.empty((x.shape[0], len(planar)))
    for i in range(x.shape[0]):
        for k in range(len(planar)):
            state ^= (state << 13) & m64
            state ^= state >> 7
            state ^= (state << 17) & m64
            out[i, k] = (state >> 11) / float(1 << 53) * 2.0 - 1.0
    for k, a in enumerate(planar):
        x[:, a] += amp * out[:, k]
    return x.ravel()


def reference_bvh(sc, jitter=True):
    """The BVH topology the reference's own builder (MakeBVHTopology_TopDown, bvh_aabb.cpp:294-313) produces, from
    libclothgpu's host builder (csrc/bvhbuild.cu; node-for-node identical, tests/test_bvh_builder.py).  jitter=True
    feeds it the seeded-jitter copy of the rest positions: on that input the
    reference's builder itself survives square grids, so the tree is exactly the one the reference would use."""
    from . import clothgpu as G
    return G.bvh_build_topdown(sc["tri"], jittered_rest(sc) if jitter else sc["xyz0"])


def median_split_bvh(sc, xyz=None):
    """A balanced BVH *topology* over the triangles (median split of the centroids along the longest axis).

    Output layout = the reference's std::vector<CNodeBVH>: rows {iroot, ichild0, ichild1}, node 0 is the root,
    a leaf has ichild1 == -1 and ichild0 == triangle index, children are numbered after their parent.  The
    reference's own builder (MakeBVHTopology_TopDown, bvh_aabb.cpp:294-313) crashes on exactly-square grids
    (SURVEY.md section 8c); any valid binary tree is a legal input of the per-step path, and this one needs
    nothing but numpy, so benches and tests on the GPU box use it for the large grids."""
    x = np.asarray(sc["xyz0"] if xyz is None else xyz, dtype=np.float64).reshape(-1, 3)
    tri = np.asarray(sc["tri"]).reshape(-1, 3)
    cent = x[tri].mean(axis=1)
    ntri = tri.shape[0]
    nodes = np.full((2 * ntri - 1, 3), -1, dtype=np.int32)
    nnode = 1
    stack = [(0, np.arange(ntri))]
    while stack:
        inode, idx = stack.pop()
        if idx.size == 1:
            nodes[inode, 1] = idx[0]
            nodes[inode, 2] = -1
            continue
        c = cent[idx]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        order = np.argsort(c[:, axis], kind="stable")
        half = idx.size // 2
        c0, c1 = nnode, nnode + 1
        nnode += 2
        nodes[inode, 1] = c0
        nodes[inode, 2] = c1
        nodes[c0, 0] = inode
        nodes[c1, 0] = inode
        stack.append((c1, idx[order[half:]]))
        stack.append((c0, idx[order[:half]]))
    assert nnode == 2 * ntri - 1
    return nodes


# demo constants -------------------------------------------------------------------------------------------
SELF_CONTACT = dict(dt=0.01, lam=0.0, myu=30.0, stiff_bend=1.0e-2, gravity=(0.0, 0.05, -10.0),
                    stiff_contact=300.0, clearance=0.01, areal_density=0.04,
                    col_kind=1, col_par=(-0.5, 0.0, 0.0, 0.0))
INTERNAL_CLOTH = dict(dt=0.02, lam=1.0, myu=4.0, stiff_bend=1.0e-3, gravity=(0.0, 0.0, -10.0),
                      stiff_contact=1.0e3, clearance=0.02, areal_density=1.0,
                      col_kind=0, col_par=(0.0, 0.0, 0.0, 0.0))

# collider kinds of the C-ABI (include/clothgpu.h)
COLLIDER_NONE, COLLIDER_PLANE_Z, COLLIDER_SPHERE, COLLIDER_SPHERE_INSIDE = 0, 1, 2, 3


def config1_default():
    """self_contact_sparse default demo: 5 x 51 vertices (elem 0.1, 0.4 x 5.0)."""
    return square_cloth_xz(4, 50, 0.1, SELF_CONTACT["areal_density"])


def config3_256():
    """256 x 256 vertices, h = 0.1, self_contact_sparse constants."""
    return square_cloth_xz(255, 255, 0.1, SELF_CONTACT["areal_density"])


def config4_512():
    """512 x 512 vertices."""
    return square_cloth_xz(511, 511, 0.1, SELF_CONTACT["areal_density"])


def crumpled_state(sc, seed=7, amp_frac=3.0):
    """Config 4's synthetic crumpled start: rest grid + a smooth out-of-plane displacement (sum of 8 seeded
    sinusoids, amplitude 3h, wavelengths 8h..64h), zero velocity -- contact candidates from step 1."""
    rng = np.random.default_rng(seed)
    x = np.array(sc["xyz0"], dtype=np.float64).reshape(-1, 3).copy()
    h = sc["h"]
    normal_axis = int(np.argmin(np.ptp(x, axis=0)))
    planar = [a for a in range(3) if a != normal_axis]
    d = np.zeros(x.shape[0])
    for _ in range(8):
        lam = h * rng.uniform(8.0, 64.0)
        ang = rng.uniform(0.0, 2.0 * np.pi)
        ph = rng.uniform(0.0, 2.0 * np.pi)
        k = 2.0 * np.pi / lam
        d += np.sin(k * (np.cos(ang) * x[:, planar[0]] + np.sin(ang) * x[:, planar[1]]) + ph)
    x[:, normal_axis] += amp_frac * h * d / 8.0
    return x.ravel().copy(), np.zeros(x.size)


def relabelled(sc, seed=11):
    """The same mesh with its vertices renumbered by a random permutation: an irregular sparsity pattern (no band
    structure, rows with scattered lower / upper neighbours) for the strip / micro-step plans."""
    rng = np.random.default_rng(seed)
    n = sc["np"]
    perm = rng.permutation(n).astype(np.int32)  # new id of old vertex i
    xyz0 = np.empty((n, 3))
    xyz0[perm] = np.asarray(sc["xyz0"]).reshape(-1, 3)
    bc = np.empty(n, dtype=np.int32)
    bc[perm] = sc["bc"]
    tri = perm[np.asarray(sc["tri"]).reshape(-1, 3)]
    quad = perm[np.asarray(sc["quad"]).reshape(-1, 4)]
    out = _finish(xyz0, bc, np.ascontiguousarray(tri, dtype=np.int32), np.ascontiguousarray(quad, dtype=np.int32), 1.0, 1.0)
    out["mass_point"] = sc["mass_point"]
    out["h"] = sc.get("h", 0.1)
    out["perm"] = perm
    return out
