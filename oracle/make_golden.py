"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libclothref.so).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python -m oracle.make_golden
The vectors are small (a 5 x 21 and the default 5 x 51 scene) and are committed; tests/test_oracle_golden.py checks
the C restatement against them, tests/test_gpu_*.py check the CUDA path against them on the GPU box.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloth_sim_self_collision_b200 import scenes as S  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from tests import common as T  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def phases(ref):
    """Per-phase dumps of one implicit step + one contact resolution on a folded 5 x 21 sheet."""
    sc = S.square_cloth_xz(4, 20, 0.1, S.SELF_CONTACT["areal_density"])
    sc["bvh"] = ref.bvh_build(sc["tri"], S.jittered_rest(sc))
    c = S.SELF_CONTACT
    p = T.params_for(sc, c, col_par=(0.2, 0, 0, 0))
    x, v = T.wavy_state(sc)
    dia, crs, b = ref.assemble(sc, x, v, p)
    idia, icrs = ref.ilu_factor(sc, dia, crs)
    rng = np.random.default_rng(5)
    vec = rng.standard_normal(sc["np"] * 3)
    yv = rng.standard_normal(sc["np"] * 3)
    pcg_x, pcg_r, conv, it = ref.pcg(sc, dia, crs, idia, icrs, b)
    g = dict(bvh=sc["bvh"], x=x, v=v, col_par=p.col_par, dia=dia, crs=crs, b=b, idia=idia, icrs=icrs, vec=vec, yv=yv,
             matvec=ref.matvec(sc, dia, crs, 1.0, vec, 0.0, yv), matvec_ab=ref.matvec(sc, dia, crs, -0.7, vec, 1.3, yv),
             ilu_solve=ref.ilu_solve(sc, idia, icrs, vec), pcg_x=pcg_x, pcg_r=pcg_r, pcg_conv=conv, pcg_iter=it,
             upd_v=ref.update_intermediate_velocity(sc, x, v, p))
    # contact phases
    fx, fv = T.folded_state(sc, gap=0.004, vel=10.0)
    bp = ref.refit_prx(sc, fx, c["clearance"])
    ep = ref.contacts_prx(sc, bp, fx, c["clearance"])
    v1 = ref.impulse_prx(fv, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"], fx, ep)
    bc = ref.refit_ccd(sc, fx, v1, c["dt"])
    ec = ref.contacts_ccd(sc, bc, fx, v1, c["dt"], c["clearance"])
    v2 = ref.impulse_ccd(v1, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"], fx, ec)
    riz = ref.riz_new()
    riz.make(ec, sc)
    nz, zones = riz.zones(sc["np"])
    v3 = riz.apply(v2, fx, v2)
    pp = T.params_for(sc, c)
    vr, applied, stats = ref.sim(sc).resolve_contacts(fv, fx, pp)
    g.update(fx=fx, fv=fv, box_prx=bp, leaf_pairs_prx=ref.leaf_pairs(sc, bp), elems_prx=ep, v_after_prx=v1, box_ccd=bc,
             leaf_pairs_ccd=ref.leaf_pairs(sc, bc), elems_ccd=ec, v_after_ccd=v2, riz_nzone=nz, riz_zone=zones,
             v_after_riz=v3, resolve_v=vr, resolve_applied=int(applied), resolve_stats=stats)
    np.savez_compressed(os.path.join(OUT, "phases_5x21.npz"), **g)
    print("phases_5x21: prx", ep.shape[0], "ccd", ec.shape[0], "zones", nz, "stats", stats[:6])


def trajectory(ref):
    """Config 1: the default self_contact_sparse scene, 200 free-running steps."""
    sc = S.config1_default()
    sc["bvh"] = ref.bvh_build(sc["tri"], sc["xyz0"])
    p = T.params_for(sc, S.SELF_CONTACT)
    sim = ref.sim(sc)
    keep = [10, 35, 36, 40, 45, 50, 60, 80, 100, 150, 200]
    stats = np.zeros((200, 40), dtype=np.int32)
    xs, vs = [], []
    for step in range(1, 201):
        st = sim.step_self_contact(p)
        stats[step - 1] = st[:40]
        if step in keep:
            x, v = sim.get_state()
            xs.append(x)
            vs.append(v)
    x, v = sim.get_state()
    checksum = float(sum(x[i] * (1 + i % 7) for i in range(x.size)))
    np.savez_compressed(os.path.join(OUT, "traj_config1.npz"), bvh=sc["bvh"], steps=np.array(keep), x=np.array(xs),
                        v=np.array(vs), stats=stats, checksum=checksum)
    print("traj_config1: checksum %.15f" % checksum, "max prx", stats[:, 1].max(), "riz steps", (stats[:, 3] > 0).sum())


def internal(ref):
    """Config 2 family: internal_cloth_sparse scene (17 x 17, pinned row, sphere obstacle), 40 steps."""
    sc = S.square_cloth_xy(16, 1.0)
    p = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=2, col_par=(0.1, 0.5, -0.8, 0.3))
    sim = ref.sim(sc)
    its = []
    xs, vs = [], []
    for step in range(1, 41):
        cr, it = sim.step_internal(p)
        its.append(it)
        if step in (1, 10, 40):
            x, v = sim.get_state()
            xs.append(x)
            vs.append(v)
    np.savez_compressed(os.path.join(OUT, "traj_internal17.npz"), steps=np.array([1, 10, 40]), x=np.array(xs),
                        v=np.array(vs), iters=np.array(its, dtype=np.int32))
    print("traj_internal17: iters", its[:10])


def internal26(ref):
    """The internal_cloth_sparse demo exactly as shipped (internal_cloth_sparse/main.cpp:137-154: ndiv 25, cloth 1 x 1)
    with the sphere collider (two presses of the space key, :481-488), 40 steps."""
    sc = S.square_cloth_xy(25, 1.0)
    p = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=2, col_par=(0.1, 0.5, -0.8, 0.3))
    sim = ref.sim(sc)
    xs, vs, its = [], [], []
    for step in range(1, 41):
        cr, it = sim.step_internal(p)
        its.append(it)
        if step in (10, 40):
            x, v = sim.get_state()
            xs.append(x)
            vs.append(v)
    np.savez_compressed(os.path.join(OUT, "traj_internal26.npz"), steps=np.array([10, 40]), x=np.array(xs),
                        v=np.array(vs), iters=np.array(its, dtype=np.int32))
    print("traj_internal26: iters", its[:10])


def bvh_hashes(ref):
    """sha256 of the CNodeBVH arrays MakeBVHTopology_TopDown builds (tests/test_bvh_builder.py::GOLDEN_CASES)."""
    import hashlib
    import json
    from tests.test_bvh_builder import GOLDEN_CASES, _mesh
    out = {}
    for case in GOLDEN_CASES:
        tri, xyz = _mesh(case)
        nodes = ref.bvh_build(tri, xyz)
        out["%d_%d_%s" % case] = hashlib.sha256(np.ascontiguousarray(nodes, dtype=np.int32).tobytes()).hexdigest()
    json.dump(out, open(os.path.join(OUT, "bvh_hashes.json"), "w"), indent=1, sort_keys=True)
    print("bvh_hashes:", len(out), "trees")


if __name__ == "__main__":
    ref = O.load_ref()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    os.makedirs(OUT, exist_ok=True)
    phases(ref)
    trajectory(ref)
    internal(ref)
    internal26(ref)
    bvh_hashes(ref)
