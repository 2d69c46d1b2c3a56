"""The BVH topology builder of the product (csrc/bvhbuild.cu through clothgpu_bvh_build_topdown, host code) against
the reference's MakeBVHTopology_TopDown (bvh_aabb.cpp:294-313): the same CNodeBVH array, node for node, on every mesh
the reference survives; a valid tree where the reference crashes (bvh_aabb.cpp:238, exactly square sub-grids)."""
import hashlib
import json
import os

import numpy as np
import pytest

from cloth_sim_self_collision_b200 import clothgpu as G
from cloth_sim_self_collision_b200 import scenes as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bvh_hashes.json")


def _hash(nodes):
    return hashlib.sha256(np.ascontiguousarray(nodes, dtype=np.int32).tobytes()).hexdigest()


def _check_tree(nodes, ntri):
    nodes = np.asarray(nodes)
    assert nodes.shape == (2 * ntri - 1, 3)
    assert nodes[0, 0] == -1
    leaf = nodes[:, 2] == -1
    assert leaf.sum() == ntri and np.array_equal(np.sort(nodes[leaf, 1]), np.arange(ntri))
    inner = np.nonzero(~leaf)[0]
    for k in (1, 2):
        ch = nodes[inner, k]
        assert np.all(ch > inner), "children are numbered after their parent"
        assert np.array_equal(nodes[ch, 0], inner)
    assert np.array_equal(np.sort(np.concatenate([nodes[inner, 1], nodes[inner, 2]])), np.arange(1, 2 * ntri - 1))


def _mesh(case):
    nx, nz, kind = case
    sc = S.square_cloth_xz(nx, nz, 0.1)
    if kind == "jitter":
        return sc["tri"], S.jittered_rest(sc)
    if kind == "relabel":
        r = S.relabelled(sc)
        return r["tri"], r["xyz0"]
    return sc["tri"], sc["xyz0"]


GOLDEN_CASES = [(4, 50, "rest"), (15, 15, "rest"), (7, 9, "relabel"), (31, 31, "jitter"), (127, 127, "jitter"),
                (255, 255, "jitter")]


def test_golden_hashes():
    """Hashes of the reference's trees, written by oracle/make_golden.py from the compiled reference."""
    gold = json.load(open(GOLD))
    for case in GOLDEN_CASES:
        tri, xyz = _mesh(case)
        nodes, nfb = G.bvh_build_topdown(tri, xyz, return_fallbacks=True)
        assert nfb == 0
        assert _hash(nodes) == gold["%d_%d_%s" % case], case


def test_against_reference_sweep(ref):
    """Every small grid (plain, jittered, randomly renumbered): identical to the reference wherever it survives.  The
    builder reports the nodes where the reference would have crashed; those meshes are checked for validity only."""
    same = crashed = 0
    for nx in range(1, 12):
        for nz in range(1, 12):
            for kind in ("rest", "jitter", "relabel"):
                tri, xyz = _mesh((nx, nz, kind))
                nodes, nfb = G.bvh_build_topdown(tri, xyz, return_fallbacks=True)
                _check_tree(nodes, len(tri) // 3)
                if nfb:
                    crashed += 1
                    continue
                assert np.array_equal(nodes, ref.bvh_build(tri, xyz)), (nx, nz, kind)
                same += 1
    assert same > 300 and crashed > 0


def test_config3_tree_is_the_reference_tree(ref):
    """The 256 x 256 bench topology (seeded-jitter copy of the rest shape)."""
    sc = S.config3_256()
    nodes = S.reference_bvh(sc)
    assert np.array_equal(nodes, ref.bvh_build(sc["tri"], S.jittered_rest(sc)))


def test_square_grids_where_the_reference_crashes():
    for n in (2, 32, 64):
        sc = S.square_cloth_xz(n - 1, n - 1, 0.1)
        nodes, nfb = G.bvh_build_topdown(sc["tri"], sc["xyz0"], return_fallbacks=True)
        _check_tree(nodes, sc["ntri"])
    # a single triangle is a single leaf
    nodes = G.bvh_build_topdown(np.array([0, 1, 2], dtype=np.int32), np.array([0, 0, 0, 1, 0, 0, 0, 0, 1.0]))
    assert nodes.tolist() == [[-1, 0, -1]]
    with pytest.raises(G.ClothGPUError):
        G.bvh_build_topdown(np.array([0, 1, 5], dtype=np.int32), np.zeros(9))
