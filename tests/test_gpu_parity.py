"""Parity of the CUDA path (through the C ABI) against the oracle and the committed golden vectors.

Bars:
  * integer / index work (leaf-pair sets, contact elements incl. is_fv, impact-zone partitions): exact
  * every floating-point result -- assembly, SpMV, dot products, ILU factor, triangular solves, PCG / CG iterates,
    boxes, impulses, impact-zone velocities, whole steps and whole trajectories: BIT-EXACT (np.array_equal).  The
    kernels keep the reference's per-entry operation order, are compiled without FMA contraction, and the dot products
    reproduce the reference's serial sum exactly (csrc/linsolve.cu k_xdot_*).  The north star asks for 1e-5; bit
    equality is needed in practice because the reference's contact response amplifies a 1e-16 perturbation of its
    input velocities to 1e-2 (tools/reference_sensitivity.py, profiles/r02_reference_sensitivity.txt).
  * CLOTHGPU_DOT=tree (the non-parity fast mode: tree reductions) is held to 1e-9 relative on contact-free steps
"""
import os

import numpy as np
import pytest

from cloth_sim_self_collision_b200 import clothgpu as G
from cloth_sim_self_collision_b200 import scenes as S
from tests import common as T

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1.0e-9


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1.0e-300))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "phases_5x21.npz"))


@pytest.fixture(scope="module")
def sc_contact(gold):
    sc = S.square_cloth_xz(4, 20, 0.1, S.SELF_CONTACT["areal_density"])
    sc["bvh"] = np.ascontiguousarray(gold["bvh"])
    return sc


@pytest.fixture(scope="module")
def sc_internal():
    return S.square_cloth_xy(12, 1.0)


def _cases(sc_contact, sc_internal):
    p1 = T.params_for(sc_contact, S.SELF_CONTACT, col_par=(0.2, 0, 0, 0))
    p2 = T.params_for(sc_internal, S.INTERNAL_CLOTH, col_kind=2, col_par=(0.3, 0.5, -0.1, 0.3))
    return [(sc_contact, p1), (sc_internal, p2)]


def test_linear_system_pipeline(oracle, sc_contact, sc_internal):
    for sc, p in _cases(sc_contact, sc_internal):
        g = G.ClothGPU(sc, with_bvh=False)
        x, v = T.wavy_state(sc)
        g.upload_state(x, v)
        g.assemble(p)
        dia, crs, b = oracle.assemble(sc, x, v, p)
        assert np.array_equal(g.read(G.BUF_VAL_DIA), dia)
        assert np.array_equal(g.read(G.BUF_VAL_CRS), crs)
        assert np.array_equal(g.read(G.BUF_VEC_R), b)
        rng = np.random.default_rng(5)
        xx = rng.standard_normal(sc["np"] * 3)
        yy = rng.standard_normal(sc["np"] * 3)
        for alpha, beta in [(1.0, 0.0), (-0.7, 1.3)]:
            g.write(G.BUF_VEC_P, xx)
            g.write(G.BUF_VEC_Z, yy)
            g.matvec(alpha, G.BUF_VEC_P, beta, G.BUF_VEC_Z)
            assert np.array_equal(g.read(G.BUF_VEC_Z), oracle.matvec(sc, dia, crs, alpha, xx, beta, yy))
        g.write(G.BUF_VEC_Z, yy)
        d = g.dot(G.BUF_VEC_P, G.BUF_VEC_Z)[0]
        assert d == oracle.dot(xx, yy)
        g.axpy(0.37, G.BUF_VEC_P, G.BUF_VEC_Z)
        assert np.array_equal(g.read(G.BUF_VEC_Z), yy + 0.37 * xx)
        # ILU(0)
        g.ilu_setvalue()
        g.ilu_factor()
        idia, icrs = oracle.ilu_factor(sc, dia, crs)
        assert np.array_equal(g.read(G.BUF_ILU_DIA), idia)
        assert np.array_equal(g.read(G.BUF_ILU_CRS), icrs)
        g.write(G.BUF_VEC_P, xx)
        g.ilu_solve(G.BUF_VEC_P)
        assert np.array_equal(g.read(G.BUF_VEC_P), oracle.ilu_solve(sc, idia, icrs, xx))
        # PCG / CG
        ox, orr, ocr, oit = oracle.pcg(sc, dia, crs, idia, icrs, b)
        cr, it = g.pcg()
        assert it[0] == oit
        assert np.array_equal(g.read(G.BUF_VEC_X), ox)
        assert np.array_equal(g.read(G.BUF_VEC_R), orr)
        assert cr[0] == ocr
        g.write(G.BUF_VEC_R, b)
        ox, orr, ocr, oit = oracle.cg(sc, dia, crs, b, 1.0e-4, 1000)
        cr, it = g.cg(1.0e-4, 1000)
        assert it[0] == oit
        # hundreds of unpreconditioned iterations, every dot product the reference's serial sum: still bit for bit
        assert np.array_equal(g.read(G.BUF_VEC_X), ox)
        assert cr[0] == ocr
        # zero right-hand side: early out (ilu_sparse.cpp:651)
        g.write(G.BUF_VEC_R, np.zeros(sc["np"] * 3))
        cr, it = g.pcg()
        assert (cr[0], it[0]) == (0.0, 0)
        assert not g.read(G.BUF_VEC_X).any()
        # gradient-only fix-up
        g.upload_state(x, v)
        g.update_intermediate_velocity(p)
        assert np.array_equal(g.download_state()[1], oracle.update_intermediate_velocity(sc, x, v, p))
        g.close()


def test_golden_linear_phases(gold, sc_contact):
    """The committed vectors were produced by the UNMODIFIED reference (oracle/make_golden.py)."""
    sc = sc_contact
    p = T.params_for(sc, S.SELF_CONTACT, col_par=tuple(gold["col_par"]))
    g = G.ClothGPU(sc, with_bvh=False)
    g.upload_state(gold["x"], gold["v"])
    g.assemble(p)
    assert np.array_equal(g.read(G.BUF_VAL_DIA), gold["dia"])
    assert np.array_equal(g.read(G.BUF_VAL_CRS), gold["crs"])
    assert np.array_equal(g.read(G.BUF_VEC_R), gold["b"])
    g.ilu_setvalue()
    g.ilu_factor()
    assert np.array_equal(g.read(G.BUF_ILU_DIA), gold["idia"])
    assert np.array_equal(g.read(G.BUF_ILU_CRS), gold["icrs"])
    g.write(G.BUF_VEC_P, gold["vec"])
    g.write(G.BUF_VEC_Z, gold["yv"])
    g.matvec(1.0, G.BUF_VEC_P, 0.0, G.BUF_VEC_Z)
    assert np.array_equal(g.read(G.BUF_VEC_Z), gold["matvec"])
    g.write(G.BUF_VEC_Z, gold["yv"])
    g.matvec(-0.7, G.BUF_VEC_P, 1.3, G.BUF_VEC_Z)
    assert np.array_equal(g.read(G.BUF_VEC_Z), gold["matvec_ab"])
    g.ilu_solve(G.BUF_VEC_P)
    assert np.array_equal(g.read(G.BUF_VEC_P), gold["ilu_solve"])
    cr, it = g.pcg()
    assert it[0] == int(gold["pcg_iter"])
    assert np.array_equal(g.read(G.BUF_VEC_X), gold["pcg_x"])
    g.upload_state(gold["x"], gold["v"])
    g.update_intermediate_velocity(p)
    assert np.array_equal(g.download_state()[1], gold["upd_v"])
    g.close()


def _sorted_pairs(p):
    p = np.asarray(p).reshape(-1, 2)
    return p[np.lexsort((p[:, 1], p[:, 0]))]


def test_golden_contact_phases(gold, sc_contact):
    sc = sc_contact
    c = S.SELF_CONTACT
    g = G.ClothGPU(sc)
    fx, fv = gold["fx"], gold["fv"]
    g.write(G.BUF_XYZ_PREV, fx)
    g.write(G.BUF_UVW, fv)
    # proximity
    g.bvh_refit_prx(G.BUF_XYZ_PREV, c["clearance"])
    assert np.array_equal(g.read(G.BUF_BOXES).reshape(-1, 6), gold["box_prx"])
    n = g.contacts_prx(G.BUF_XYZ_PREV, c["clearance"])[0]
    assert np.array_equal(_sorted_pairs(g.leaf_pairs()), _sorted_pairs(gold["leaf_pairs_prx"]))
    assert np.array_equal(g.elems(n), gold["elems_prx"])
    g.impulse_prx(G.BUF_XYZ_PREV, G.BUF_UVW, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"])
    assert np.array_equal(g.read(G.BUF_UVW), gold["v_after_prx"])
    # CCD
    g.bvh_refit_ccd(G.BUF_XYZ_PREV, G.BUF_UVW, c["dt"])
    assert np.array_equal(g.read(G.BUF_BOXES).reshape(-1, 6), gold["box_ccd"])
    n = g.contacts_ccd(G.BUF_XYZ_PREV, G.BUF_UVW, c["dt"], c["clearance"])[0]
    assert np.array_equal(_sorted_pairs(g.leaf_pairs()), _sorted_pairs(gold["leaf_pairs_ccd"]))
    assert np.array_equal(g.elems(n), gold["elems_ccd"])
    assert n > 0
    g.impulse_ccd(G.BUF_XYZ_PREV, G.BUF_UVW, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"])
    assert np.array_equal(g.read(G.BUF_UVW), gold["v_after_ccd"])
    # impact zones on the same elements
    g.write(G.BUF_UVW0, gold["v_after_ccd"])
    g.riz_reset()
    g.riz_make()
    _check_partition(g.read(G.BUF_ZONE), gold["riz_zone"])
    g.riz_apply(G.BUF_XYZ_PREV, G.BUF_UVW)
    assert np.array_equal(g.read(G.BUF_UVW), gold["v_after_riz"])
    # the whole resolution from the same start
    p = T.params_for(sc, c)
    g.write(G.BUF_XYZ_PREV, fx)
    g.write(G.BUF_UVW, fv)
    applied, stats = g.resolve_self_contact(p)
    assert np.array_equal(g.read(G.BUF_UVW), gold["resolve_v"])
    assert int(applied[0]) == int(gold["resolve_applied"])
    assert np.array_equal(stats[0][:5], gold["resolve_stats"][:5])
    assert np.array_equal(stats[0][8:116], gold["resolve_stats"][8:116])
    g.close()


def _check_partition(label_gpu, zone_ref):
    """Same partition of the member vertices (zone numbering differs: smallest vertex id vs creation order)."""
    label_gpu = np.asarray(label_gpu)
    zone_ref = np.asarray(zone_ref)
    assert np.array_equal(label_gpu >= 0, zone_ref >= 0)
    m = label_gpu >= 0
    pairs = set(zip(label_gpu[m].tolist(), zone_ref[m].tolist()))
    assert len(pairs) == len(set(label_gpu[m].tolist())) == len(set(zone_ref[m].tolist()))


def _contact_pipeline_vs_oracle(oracle, sc, x, v, c, rounds=3):
    """Teacher-forced comparison of every contact phase against the oracle on one state."""
    g = G.ClothGPU(sc)
    g.write(G.BUF_XYZ_PREV, x)
    g.write(G.BUF_UVW, v)
    found = {}
    g.bvh_refit_prx(G.BUF_XYZ_PREV, c["clearance"])
    bp = oracle.refit_prx(sc, x, c["clearance"])
    assert np.array_equal(g.read(G.BUF_BOXES).reshape(-1, 6), bp)
    n = g.contacts_prx(G.BUF_XYZ_PREV, c["clearance"])[0]
    assert np.array_equal(_sorted_pairs(g.leaf_pairs()), _sorted_pairs(oracle.leaf_pairs(sc, bp)))
    ep = oracle.contacts_prx(sc, bp, x, c["clearance"])
    assert np.array_equal(g.elems(n), ep)
    found["prx"] = n
    g.impulse_prx(G.BUF_XYZ_PREV, G.BUF_UVW, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"])
    v1 = oracle.impulse_prx(v, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"], x, ep)
    assert np.array_equal(g.read(G.BUF_UVW), v1)
    riz = oracle.riz_new()
    g.riz_reset()
    v0 = None
    for r in range(rounds):
        g.bvh_refit_ccd(G.BUF_XYZ_PREV, G.BUF_UVW, c["dt"])
        bc = oracle.refit_ccd(sc, x, v1, c["dt"])
        assert np.array_equal(g.read(G.BUF_BOXES).reshape(-1, 6), bc)
        n = g.contacts_ccd(G.BUF_XYZ_PREV, G.BUF_UVW, c["dt"], c["clearance"])[0]
        assert np.array_equal(_sorted_pairs(g.leaf_pairs()), _sorted_pairs(oracle.leaf_pairs(sc, bc)))
        ec = oracle.contacts_ccd(sc, bc, x, v1, c["dt"], c["clearance"])
        assert np.array_equal(g.elems(n), ec)
        found["ccd%d" % r] = n
        if n == 0:
            break
        if r == 0:
            g.impulse_ccd(G.BUF_XYZ_PREV, G.BUF_UVW, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"])
            v1 = oracle.impulse_ccd(v1, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"], x, ec)
            assert np.array_equal(g.read(G.BUF_UVW), v1)
            v0 = v1.copy()
            g.write(G.BUF_UVW0, v0)
        else:
            g.riz_make()
            riz.make(ec, sc)
            _check_partition(g.read(G.BUF_ZONE), riz.zones(sc["np"])[1])
            g.riz_apply(G.BUF_XYZ_PREV, G.BUF_UVW)
            v1 = riz.apply(v1, x, v0)
            assert np.array_equal(g.read(G.BUF_UVW), v1)
    g.close()
    return found


def test_contact_pipeline_small_folded(oracle, sc_contact):
    x, v = T.folded_state(sc_contact, gap=0.004, vel=10.0)
    found = _contact_pipeline_vs_oracle(oracle, sc_contact, x, v, S.SELF_CONTACT)
    assert found["prx"] > 0 and found["ccd0"] > 0


def _scene_with_bvh(oracle, nx, nz):
    """BVH topology: the reference builder's tree (MakeBVHTopology_TopDown on the seeded-jitter rest copy) from the
    product's host builder, which tests/test_bvh_builder.py pins node for node to the reference."""
    sc = S.square_cloth_xz(nx, nz, 0.1, S.SELF_CONTACT["areal_density"])
    sc["bvh"] = S.reference_bvh(sc)
    return sc


def test_contact_pipeline_48x48_folded(oracle):
    sc = _scene_with_bvh(oracle, 47, 47)
    x, v = T.folded_state(sc, gap=0.006, vel=6.0)
    found = _contact_pipeline_vs_oracle(oracle, sc, x, v, S.SELF_CONTACT, rounds=4)
    assert found["prx"] > 100 and found["ccd0"] > 10 and found["ccd1"] > 0


def test_resolve_and_step_vs_oracle(oracle):
    sc = _scene_with_bvh(oracle, 23, 40)
    c = S.SELF_CONTACT
    p = T.params_for(sc, c)
    x, v = T.folded_state(sc, gap=0.006, vel=4.0)
    g = G.ClothGPU(sc)
    # GetIntermidiateVelocityContactResolved
    g.write(G.BUF_XYZ_PREV, x)
    g.write(G.BUF_UVW, v)
    applied, stats = g.resolve_self_contact(p)
    sim = oracle.sim(sc)
    ov, oapplied, ostats = sim.resolve_contacts(v, x, p)
    assert bool(applied[0]) == oapplied
    assert np.array_equal(stats[0][1:5], ostats[1:5])
    assert np.array_equal(stats[0][8:116], ostats[8:116])
    assert np.array_equal(g.read(G.BUF_UVW), ov)
    assert ostats[1] > 0 and ostats[2] > 0
    # StepTime, teacher forced over a few steps
    for step in range(4):
        g.upload_state(x, v)
        sim.set_state(x, v)
        st = g.step_self_contact(p)[0]
        ost = sim.step_self_contact(p)
        gx, gv = g.download_state()
        x, v = sim.get_state()
        assert st[0] == ost[0], "PCG iterations"
        assert np.array_equal(st[1:5], ost[1:5])
        assert np.array_equal(gx, x) and np.array_equal(gv, v)
    g.close()


def test_internal_trajectory_golden():
    """Config-2 family: 17 x 17 pinned sheet falling onto a sphere, 40 free-running steps (reference vectors)."""
    gold = np.load(os.path.join(GOLD, "traj_internal17.npz"))
    sc = S.square_cloth_xy(16, 1.0)
    p = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=2, col_par=(0.1, 0.5, -0.8, 0.3))
    g = G.ClothGPU(sc, with_bvh=False)
    its = []
    k = 0
    for step in range(1, 41):
        its.append(int(g.step_internal(p)[0]))
        if step == int(gold["steps"][k]):
            x, v = g.download_state()
            assert np.array_equal(x, gold["x"][k]) and np.array_equal(v, gold["v"][k])
            k += 1
    assert its == list(gold["iters"])
    g.close()


def test_config1_trajectory_golden():
    """Config 1: the default self_contact_sparse scene, 200 FREE-RUNNING steps (proximity + CCD impulses and impact
    zones from step ~36 on): statistics and states bit-identical to the reference's own trajectory throughout, and the
    survey's known-answer checksum of the final positions."""
    gold = np.load(os.path.join(GOLD, "traj_config1.npz"))
    sc = S.config1_default()
    sc["bvh"] = np.ascontiguousarray(gold["bvh"])
    p = T.params_for(sc, S.SELF_CONTACT)
    g = G.ClothGPU(sc)
    keep = list(gold["steps"])
    errs = {}
    for step in range(1, 201):
        st = g.step_self_contact(p)[0]
        if step <= len(gold["stats"]):
            assert np.array_equal(st[:5], gold["stats"][step - 1][:5]), step
        if step in keep:
            k = keep.index(step)
            x, v = g.download_state()
            errs[step] = (rel_err(x, gold["x"][k]), rel_err(v, gold["v"][k]))
            assert np.array_equal(x, gold["x"][k]) and np.array_equal(v, gold["v"][k]), (step, errs[step])
    assert len(errs) == len(keep)
    x, _ = g.download_state()
    assert float(np.dot(x, 1.0 + (np.arange(x.size) % 7))) == -89.699184471192837
    g.close()


def test_batch_equals_single(oracle, sc_contact):
    """Scenes of a batch are independent: each one matches the same scene run alone (bit for bit)."""
    sc = sc_contact
    p = T.params_for(sc, S.SELF_CONTACT)
    states = [T.folded_state(sc, seed=3 + k, gap=0.004 + 0.001 * k, vel=10.0 - 3.0 * k) for k in range(3)]
    states.append(T.wavy_state(sc))  # a scene without self contact: leaves the pipeline early
    singles = []
    for x, v in states:
        g = G.ClothGPU(sc)
        g.upload_state(x, v)
        st = g.step_self_contact(p, 2)
        singles.append((g.download_state(), st[0].copy()))
        g.close()
    gb = G.ClothGPU(sc, nscene=len(states))
    for s, (x, v) in enumerate(states):
        gb.upload_state(x, v, scene=s)
    st = gb.step_self_contact(p, 2)
    for s in range(len(states)):
        x, v = gb.download_state(scene=s)
        assert np.array_equal(x, singles[s][0][0]) and np.array_equal(v, singles[s][0][1])
        assert np.array_equal(st[s], singles[s][1])
    gb.close()


def test_large_scene_properties(oracle):
    """256 x 256 (BASELINE.json config 3) at full size through size-independent properties: the matrix the GPU
    assembles is symmetric block-wise, A x matches a float128-free re-evaluation on sampled rows via the oracle's
    SpMV on the GPU's own matrix, ILU solve inverts the factors, and contact sets are sorted and unique."""
    sc = S.config3_256()
    sc["bvh"] = S.reference_bvh(sc)
    p = T.params_for(sc, S.SELF_CONTACT)
    x, v = T.folded_state(sc, gap=0.006, vel=3.0)
    g = G.ClothGPU(sc)
    g.upload_state(x, v)
    g.assemble(p)
    dia = g.read(G.BUF_VAL_DIA)
    crs = g.read(G.BUF_VAL_CRS)
    # block symmetry: A_ij = A_ji^T
    rp, ci = sc["rowptr"], sc["colidx"]
    rows = np.repeat(np.arange(sc["np"]), np.diff(rp))
    key = rows.astype(np.int64) * sc["np"] + ci
    tkey = ci.astype(np.int64) * sc["np"] + rows
    pos = np.searchsorted(key, tkey)
    assert np.array_equal(key[pos], tkey)
    blk = crs.reshape(-1, 3, 3)
    assert np.abs(blk - blk[pos].transpose(0, 2, 1)).max() <= 1.0e-9 * np.abs(blk).max()
    # SpMV and triangular solves against the oracle on the GPU's own matrices (bit-exact paths)
    rng = np.random.default_rng(1)
    xx = rng.standard_normal(sc["np"] * 3)
    g.write(G.BUF_VEC_P, xx)
    g.matvec(1.0, G.BUF_VEC_P, 0.0, G.BUF_VEC_Z)
    assert np.array_equal(g.read(G.BUF_VEC_Z), oracle.matvec(sc, dia, crs, 1.0, xx, 0.0, np.zeros_like(xx)))
    g.ilu_setvalue()
    g.ilu_factor()
    idia, icrs = oracle.ilu_factor(sc, dia, crs)
    assert np.array_equal(g.read(G.BUF_ILU_DIA), idia) and np.array_equal(g.read(G.BUF_ILU_CRS), icrs)
    g.ilu_solve(G.BUF_VEC_P)
    assert np.array_equal(g.read(G.BUF_VEC_P), oracle.ilu_solve(sc, idia, icrs, xx))
    # contact sets: sorted, unique, canonical
    g.write(G.BUF_XYZ_PREV, x)
    g.bvh_refit_ccd(G.BUF_XYZ_PREV, G.BUF_UVW, p.dt)
    n = g.contacts_ccd(G.BUF_XYZ_PREV, G.BUF_UVW, p.dt, p.clearance)[0]
    e = g.elems(n)
    assert n > 1000
    k = e[:, :4].astype(np.int64)
    order = np.lexsort((k[:, 3], k[:, 2], k[:, 1], k[:, 0]))
    assert np.array_equal(order, np.arange(n))
    assert len(np.unique(k, axis=0)) == n
    fv = e[:, 4] == 1
    assert np.all(k[fv, 0] < k[fv, 1]) and np.all(k[fv, 1] < k[fv, 2])
    assert np.all(k[~fv, 0] < k[~fv, 1]) and np.all(k[~fv, 2] < k[~fv, 3]) and np.all(k[~fv, 0] < k[~fv, 2])
    # and the oracle agrees on the set (about 10 s of CPU)
    bc = oracle.refit_ccd(sc, x, v, p.dt)
    assert np.array_equal(g.read(G.BUF_BOXES).reshape(-1, 6), bc)
    assert np.array_equal(e, oracle.contacts_ccd(sc, bc, x, v, p.dt, p.clearance))
    g.close()


def _linear_chain_vs_oracle(oracle, sc, p, x, v):
    """assemble -> ILU factor -> triangular solve -> PCG on one state, against the oracle (bit-exact paths)."""
    g = G.ClothGPU(sc, with_bvh=False)
    g.upload_state(x, v)
    g.assemble(p)
    dia, crs = g.read(G.BUF_VAL_DIA), g.read(G.BUF_VAL_CRS)
    g.ilu_setvalue()
    g.ilu_factor()
    idia, icrs = oracle.ilu_factor(sc, dia, crs)
    assert np.array_equal(g.read(G.BUF_ILU_DIA), idia) and np.array_equal(g.read(G.BUF_ILU_CRS), icrs)
    rng = np.random.default_rng(3)
    xx = rng.standard_normal(sc["np"] * 3)
    g.write(G.BUF_VEC_P, xx)
    g.ilu_solve(G.BUF_VEC_P)
    assert np.array_equal(g.read(G.BUF_VEC_P), oracle.ilu_solve(sc, idia, icrs, xx))
    g.close()


def test_irregular_numbering_linear_chain(oracle):
    """A randomly renumbered 24 x 31 mesh: no band structure, many more ghosts / imports per strip, operands far
    beyond the shared-memory ring of the factorisation -- and a forced strip count that is not the planner's."""
    base = S.square_cloth_xz(23, 30, 0.1, S.SELF_CONTACT["areal_density"])
    sc = S.relabelled(base)
    p = T.params_for(sc, S.SELF_CONTACT)
    xb, vb = T.wavy_state(base)
    x = np.empty_like(xb).reshape(-1, 3)
    v = np.empty_like(vb).reshape(-1, 3)
    x[sc["perm"]] = xb.reshape(-1, 3)
    v[sc["perm"]] = vb.reshape(-1, 3)
    for strips in (None, "5"):
        if strips:
            os.environ["CLOTHGPU_STRIPS"] = strips
        try:
            _linear_chain_vs_oracle(oracle, sc, p, x.ravel().copy(), v.ravel().copy())
        finally:
            os.environ.pop("CLOTHGPU_STRIPS", None)


def test_config2_internal_128(oracle):
    """BASELINE.json configs[1]: internal_cloth_sparse 128 x 128, both variants of SURVEY.md 8d (2a: cloth_size 1,
    PCG runs into the 100-iteration cap; 2b: element size 0.04), three free-running steps against the oracle."""
    for size in (1.0, 127 * 0.04):
        sc = S.square_cloth_xy(127, size)
        p = T.params_for(sc, S.INTERNAL_CLOTH)
        g = G.ClothGPU(sc, with_bvh=False)
        sim = oracle.sim(sc)
        for step in range(3):
            it = int(g.step_internal(p)[0])
            oit = int(sim.step_internal(p)[1])
            assert it == oit, (size, step, it, oit)
        gx, gv = g.download_state()
        ox, ov = sim.get_state()
        if size == 1.0:
            assert it == 100  # 2a never converges: 100 iterations of rounding-dominated iterates, still bit for bit
        assert np.array_equal(gx, ox) and np.array_equal(gv, ov), (size, rel_err(gx, ox), rel_err(gv, ov))
        g.close()


def test_config4_crumpled_512(oracle):
    """BASELINE.json configs[3]: 512 x 512 crumpled cloth.  Factor and triangular solves bit-exact on the GPU's own
    matrix (72 / 64 strips instead of 26), and one full self-contact step against the oracle."""
    sc = S.config4_512()
    sc["bvh"] = S.reference_bvh(sc)
    p = T.params_for(sc, S.SELF_CONTACT)
    x, v = S.crumpled_state(sc)
    _linear_chain_vs_oracle(oracle, sc, p, x, v)
    g = G.ClothGPU(sc)
    g.upload_state(x, v)
    st = g.step_self_contact(p)[0]
    gx, gv = g.download_state()
    sim = oracle.sim(sc)
    sim.set_state(x, v)
    ost = sim.step_self_contact(p)
    ox, ov = sim.get_state()
    assert np.array_equal(st[:5], ost[:5]), (st[:5], ost[:5])
    assert np.array_equal(gx, ox) and np.array_equal(gv, ov)
    g.close()


def test_tiny_and_ragged_scenes(oracle):
    """The smallest meshes the reference's scene generator can make (2 x 2 and 2 x 3 vertices: one or two cells, a
    single strip, rows without any dependency) and a 3 x 9 strip, three self-contact steps each."""
    for nx, nz in ((1, 1), (1, 2), (2, 8)):
        sc = S.square_cloth_xz(nx, nz, 0.1, S.SELF_CONTACT["areal_density"])
        # any valid binary tree is a legal input of the per-step path: a median-split tree here, the reference
        # builder's tree (S.reference_bvh) everywhere else
        sc["bvh"] = S.median_split_bvh(sc)
        p = T.params_for(sc, S.SELF_CONTACT)
        g = G.ClothGPU(sc)
        sim = oracle.sim(sc)
        for _ in range(3):
            st = g.step_self_contact(p)[0]
            ost = sim.step_self_contact(p)
            assert np.array_equal(st[:5], ost[:5]), (nx, nz, st[:5], ost[:5])
        gx, gv = g.download_state()
        ox, ov = sim.get_state()
        assert np.array_equal(gx, ox) and np.array_equal(gv, ov), (nx, nz)
        g.close()


def test_contact_ops_need_a_bvh_and_zero_clearance_finds_nothing(oracle, sc_contact):
    p = T.params_for(sc_contact, S.SELF_CONTACT)
    g = G.ClothGPU(sc_contact, with_bvh=False)
    with pytest.raises(G.ClothGPUError):
        g.step_self_contact(p)  # CLOTHGPU_ERR_STATE: context created without a BVH
    g.close()
    # clearance 0: CAABB3D::AddPoint ignores eps <= 0, every proximity box stays empty (aabb.h:102-122)
    x, v = T.folded_state(sc_contact)
    g = G.ClothGPU(sc_contact)
    g.upload_state(x, v)
    g.write(G.BUF_XYZ_PREV, x)
    p0 = T.params_for(sc_contact, S.SELF_CONTACT, clearance=0.0)
    applied, stats = g.resolve_self_contact(p0)
    assert stats[0][1] == 0  # no proximity elements
    g.close()


@pytest.mark.gpu
def test_pinned_host_buffers_round_trip():
    """clothgpu_alloc_host buffers (page-locked) through upload_state / download_state(out=...): same bytes as the
    pageable path, and the step run from them gives the same state."""
    sc = S.square_cloth_xz(9, 7, 0.1)
    sc["bvh"] = S.reference_bvh(sc)
    rng = np.random.default_rng(3)
    x0 = sc["xyz0"].reshape(-1) + 1e-3 * rng.standard_normal(sc["np"] * 3)
    v0 = 1e-2 * rng.standard_normal(sc["np"] * 3)
    g = G.ClothGPU(sc)
    hx, hv = g.alloc_host(x0.size), g.alloc_host(v0.size)
    hx[:] = x0
    hv[:] = v0
    g.upload_state(hx, hv)
    x1, v1 = g.download_state()
    assert np.array_equal(x1, x0) and np.array_equal(v1, v0)
    ox, ov = g.alloc_host(x0.size), g.alloc_host(v0.size)
    rx, rv = g.download_state(out=(ox, ov))
    assert rx is ox and np.array_equal(ox, x0) and np.array_equal(ov, v0)
    g.close()


@pytest.mark.gpu
def test_measurement_switches_do_not_change_results(oracle, monkeypatch):
    """The execution variants behind the CLOTHGPU_* switches (strip-pipelined vs level-scheduled sweeps, incremental
    vs full CCD passes, fused vs per-level traversal, data-flow vs level-scheduled impulses) are the same arithmetic:
    a folded cloth stepped through several CCD passes and impact-zone rounds ends bit-identical in all of them."""
    sc = _scene_with_bvh(oracle, 31, 40)
    c = S.SELF_CONTACT
    p = T.params_for(sc, c)
    x0, v0 = T.folded_state(sc, gap=0.006, vel=5.0)
    variants = [{}, {"CLOTHGPU_INCREMENTAL": "0"}, {"CLOTHGPU_FUSED_TRAVERSAL": "0"}, {"CLOTHGPU_TRISOLVE": "levels"},
                {"CLOTHGPU_IMP_THREADS": "64"}, {"CLOTHGPU_IMP_FAST": "0"}, {"CLOTHGPU_QUERY_TRAVERSAL": "0"},
                {"CLOTHGPU_INC_REFIT": "0"}, {"CLOTHGPU_IMP_LANES": "1", "CLOTHGPU_IMP_SLEEP": "0"},
                {"CLOTHGPU_CANON_SMALL": "0"}, {"CLOTHGPU_QUERY_CAP": "2"}, {"CLOTHGPU_RIZ_FUSED": "0"},
                {"CLOTHGPU_SLIM_FRONTIER": "0"}]
    results = []
    for env in variants:
        for k in ("CLOTHGPU_INCREMENTAL", "CLOTHGPU_FUSED_TRAVERSAL", "CLOTHGPU_TRISOLVE", "CLOTHGPU_IMP_THREADS",
                  "CLOTHGPU_IMP_FAST", "CLOTHGPU_QUERY_TRAVERSAL", "CLOTHGPU_INC_REFIT", "CLOTHGPU_IMP_LANES",
                  "CLOTHGPU_IMP_SLEEP", "CLOTHGPU_CANON_SMALL", "CLOTHGPU_QUERY_CAP",
                  "CLOTHGPU_RIZ_FUSED", "CLOTHGPU_SLIM_FRONTIER"):
            monkeypatch.delenv(k, raising=False)
        for k, val in env.items():
            monkeypatch.setenv(k, val)
        g = G.ClothGPU(sc)
        g.upload_state(x0, v0)
        stats = np.array([g.step_self_contact(p, 1)[0][:5] for _ in range(5)])
        results.append((g.download_state(), stats))
        g.close()
    (bx, bv), bst = results[0]
    assert bst[:, 1].max() > 0 and bst[:, 2].max() > 1, "the scene must exercise proximity contacts and several CCD passes"
    for env, ((x, v), st) in zip(variants[1:], results[1:]):
        assert np.array_equal(st, bst), env
        assert np.array_equal(x, bx) and np.array_equal(v, bv), env


def test_step_internal_cg_vs_oracle(oracle):
    """clothgpu_step_internal_cg = StepTime_InternalDynamics (solve_internal_sparse.h:185-253, Solve_CG 1e-4 / 1000),
    teacher forced AND free running (the GPU state is never corrected): bit-identical states, iteration counts and
    convergence ratios."""
    sc = S.square_cloth_xy(12, 1.0)
    p = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=2, col_par=(0.1, 0.5, -0.8, 0.3))
    g = G.ClothGPU(sc, with_bvh=False)
    sim = oracle.sim(sc)
    x, v = sim.get_state()
    for step in range(6):
        if step < 3:
            g.upload_state(x, v)
        cr, it = g.step_internal_cg(p)
        ocr, oit = sim.step_internal_cg(p)
        gx, gv = g.download_state()
        x, v = sim.get_state()
        assert int(it[0]) == oit, (step, it, oit)
        assert cr[0] == ocr
        assert np.array_equal(gx, x) and np.array_equal(gv, v), (step, rel_err(gx, x), rel_err(gv, v))
    assert oit > 3
    g.close()


def test_collider_sphere_inside(oracle):
    """CLOTHGPU_COLLIDER_SPHERE_INSIDE (the container sphere of self_contact_sparse/main.cpp:117-133): assembled
    matrix / right-hand side bit-exact on a state that touches the wall, then free-running implicit steps."""
    sc = S.square_cloth_xy(12, 1.0)
    p = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=3, col_par=(0.5, 0.5, -0.2, 0.62))
    x, v = T.wavy_state(sc)
    g = G.ClothGPU(sc, with_bvh=False)
    g.upload_state(x, v)
    g.assemble(p)
    dia, crs, b = oracle.assemble(sc, x, v, p)
    # the wall must be active for some vertices, or the test says nothing about the collider
    p_none = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=0)
    assert not np.array_equal(oracle.assemble(sc, x, v, p_none)[0], dia)
    assert np.array_equal(g.read(G.BUF_VAL_DIA), dia)
    assert np.array_equal(g.read(G.BUF_VAL_CRS), crs)
    assert np.array_equal(g.read(G.BUF_VEC_R), b)
    sim = oracle.sim(sc)
    sim.set_state(x, v)
    g.upload_state(x, v)
    for step in range(5):
        it = int(g.step_internal(p)[0])
        assert it == int(sim.step_internal(p)[1])
    gx, gv = g.download_state()
    ox, ov = sim.get_state()
    assert np.array_equal(gx, ox) and np.array_equal(gv, ov)
    g.close()


def test_config3_contact_regime_vs_oracle(oracle):
    """The benchmarked workload itself (BASELINE.json configs[2]: 256 x 256, reference BVH tree, from rest): the GPU
    trajectory is taken into the self-contact regime and single steps there (58: first CCD passes; 66: thousands of
    contact elements, five CCD passes, impact zones) are teacher forced on the C restatement AND on the compiled
    reference (oracle/_ref, when it travelled): statistics, per-pass element counts, X and V bit-identical."""
    from oracle import oracle_py as O
    sc = S.config3_256()
    sc["bvh"] = S.reference_bvh(sc)
    p = T.params_for(sc, S.SELF_CONTACT)
    g = G.ClothGPU(sc)
    libs = [("oracle", oracle)]
    ref = O.load_ref()
    if ref is not None:
        libs.append(("reference", ref))
    done = 0
    seen = []
    for target in (58, 66):
        g.step_self_contact(p, target - 1 - done)
        done = target - 1
        x, v = g.download_state()
        st = g.step_self_contact(p)[0]
        done += 1
        gx, gv = g.download_state()
        seen.append(st[:5].copy())
        for name, lib in libs:
            sim = lib.sim(sc)
            sim.set_state(x, v)
            ost = sim.step_self_contact(p)
            ox, ov = sim.get_state()
            assert np.array_equal(st[:5], ost[:5]), (name, target, st[:5], ost[:5])
            assert np.array_equal(st[8:116], ost[8:116]), (name, target)
            assert np.array_equal(gx, ox) and np.array_equal(gv, ov), (name, target, rel_err(gx, ox), rel_err(gv, ov))
    print("config 3 contact regime stats (pcg, prox, ccd passes, riz rounds, applied):", seen)
    assert seen[1][1] > 1000 and seen[1][2] >= 3, "step 66 must be deep in the contact regime"
    g.close()


def _serial_dot_cases(n, rng):
    """Vectors that stress every path of the exact dot: binade crossings up and down, cancellation, exact ties,
    zeros, huge dynamic range, subnormal products, a sum that returns to zero."""
    one = np.ones(n)
    cases = {
        "normal": (rng.standard_normal(n), rng.standard_normal(n)),
        "squares": ((lambda a: (a, a))(rng.standard_normal(n))),
        "wide_range": (rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 12, n), rng.standard_normal(n)),
        "growing": (np.exp(np.linspace(-30, 30, n)), one),
        "shrinking": (np.exp(np.linspace(30, -30, n)), one),
        "alternating_cancel": (np.where(np.arange(n) % 2 == 0, 1.0, -1.0) * (1.0 + 1e-9 * rng.standard_normal(n)), 1e6 * one),
        "ties": (np.where(rng.random(n) < 0.3, 0.5, 1.0) * 2.0 ** rng.integers(-3, 3, n), one),
        "integers": (rng.integers(-1000, 1000, n).astype(np.float64), rng.integers(-1000, 1000, n).astype(np.float64)),
        "mostly_zero": (np.where(rng.random(n) < 0.01, rng.standard_normal(n), 0.0), one),
        "all_zero": (np.zeros(n), one),
        "tiny": (1e-160 * rng.standard_normal(n), 1e-160 * rng.standard_normal(n)),
        "huge": (1e150 * rng.standard_normal(n), 1e150 * rng.random(n)),
    }
    big_then_small = rng.standard_normal(n)
    big_then_small[: n // 3] *= 1e15
    big_then_small[n // 3: 2 * n // 3] *= -1e15
    cases["return_to_small"] = (big_then_small, one)
    sym = rng.standard_normal(n)
    sym[n // 2:] = -sym[: n - n // 2][::-1]
    cases["sum_to_zero"] = (sym, one)
    return cases


def test_exact_dot_is_the_serial_sum(oracle):
    """clothgpu_dot == InnerProduct (matrix_square_sparse.cpp:313-324), the serial left-to-right sum, bit for bit:
    lengths that are / are not multiples of the 256-element chunks, one chunk, 256 x 256 scale."""
    rng = np.random.default_rng(17)
    for nx, nz in ((3, 5), (12, 12), (40, 40), (255, 255)):
        sc = S.square_cloth_xz(nx, nz, 0.1)
        g = G.ClothGPU(sc, with_bvh=False)
        n = sc["np"] * 3
        for name, (a, b) in _serial_dot_cases(n, rng).items():
            g.write(G.BUF_VEC_P, a)
            g.write(G.BUF_VEC_Z, b)
            d = g.dot(G.BUF_VEC_P, G.BUF_VEC_Z)[0]
            assert d == oracle.dot(a, b), (n, name, d, oracle.dot(a, b))
        # non-finite input: whatever the serial loop gives (inf / nan), compared as bit patterns
        a = rng.standard_normal(n)
        a[n // 2] = np.inf
        g.write(G.BUF_VEC_P, a)
        g.write(G.BUF_VEC_Z, np.ones(n))
        d = g.dot(G.BUF_VEC_P, G.BUF_VEC_Z)[0]
        assert np.array([d]).tobytes() == np.array([oracle.dot(a, np.ones(n))]).tobytes()
        g.close()


def test_tree_dot_fast_mode(oracle, monkeypatch):
    """CLOTHGPU_DOT=tree (non-parity fast mode): fixed-shape tree reductions instead of the serial
    sums.  Deterministic run to run; PCG iterates agree with the reference to rounding (1e-9), not bit for bit."""
    monkeypatch.setenv("CLOTHGPU_DOT", "tree")
    sc = S.square_cloth_xy(24, 24 * 0.04)
    p = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=2, col_par=(0.4, 0.5, -0.6, 0.3))
    runs = []
    for rep in range(2):
        g = G.ClothGPU(sc, with_bvh=False)
        its = [int(g.step_internal(p)[0]) for _ in range(5)]
        runs.append((g.download_state(), its))
        g.close()
    assert np.array_equal(runs[0][0][0], runs[1][0][0]) and runs[0][1] == runs[1][1]
    sim = oracle.sim(sc)
    oits = [int(sim.step_internal(p)[1]) for _ in range(5)]
    ox, ov = sim.get_state()
    assert runs[0][1] == oits
    assert rel_err(runs[0][0][0], ox) < RTOL and rel_err(runs[0][0][1], ov) < 1.0e-7
    assert not np.array_equal(runs[0][0][1], ov), "the tree mode is not expected to be bit-identical"


def test_element_level_entry_points(oracle):
    """The reference's per-element free functions (cloth_internal_physics.h, bvh_aabb.h:49-61) as GPU batches through
    the C ABI: bit-identical to the oracle, element by element."""
    rng = np.random.default_rng(23)
    n = 300
    C3 = rng.standard_normal((n, 3, 3))
    c3 = C3 + 0.3 * rng.standard_normal((n, 3, 3))
    for lam, myu in ((0.0, 30.0), (1.0, 4.0)):
        W, dW, ddW = G.elem_cst(C3, c3, lam, myu)
        for k in range(0, n, 7):
            ow, odw, oddw = oracle.cst(C3[k], c3[k], lam, myu)
            assert W[k] == ow and np.array_equal(dW[k].ravel(), np.ravel(odw)) and np.array_equal(ddW[k].ravel(), np.ravel(oddw))
    C4 = rng.standard_normal((n, 4, 3))
    c4 = C4 + 0.2 * rng.standard_normal((n, 4, 3))
    W, dW, ddW = G.elem_bend(C4, c4, 1.0e-2)
    for k in range(0, n, 7):
        ow, odw, oddw = oracle.bend(C4[k], c4[k], 1.0e-2)
        assert W[k] == ow and np.array_equal(dW[k].ravel(), np.ravel(odw)) and np.array_equal(ddW[k].ravel(), np.ravel(oddw))
    # contact: the plane collider's (pd, n) evaluated here like the host callback would
    pts = rng.standard_normal((n, 3))
    pdn = np.stack([-0.5 - pts[:, 2], np.zeros(n), np.zeros(n), np.ones(n)], axis=1)
    W, dW, ddW = G.elem_contact(pdn, 300.0, 0.01)
    for k in range(0, n, 7):
        ow, odw, oddw = oracle.contact(pts[k], 300.0, 0.01, 1, (-0.5, 0, 0, 0))
        assert W[k] == ow and np.array_equal(dW[k], np.ravel(odw)) and np.array_equal(ddW[k].ravel(), np.ravel(oddw))
    assert (W > 0).any() and (W == 0).any()
    # stress clamp
    s2 = rng.standard_normal((n, 3)) * np.array([1.0, 1.0, 0.5])
    s3 = G.elem_psd_sym22(s2)
    lam_min = 0.5 * (s3[:, 0] + s3[:, 1]) - np.sqrt(0.25 * (s3[:, 0] - s3[:, 1]) ** 2 + s3[:, 2] ** 2)
    assert np.all(lam_min > -1e-12)
    psd = 0.5 * (s2[:, 0] + s2[:, 1]) - np.sqrt(0.25 * (s2[:, 0] - s2[:, 1]) ** 2 + s2[:, 2] ** 2) > 1e-20
    assert np.array_equal(s3[psd], s2[psd]) and psd.any() and (~psd).any()
    # geometric predicates
    P = rng.standard_normal((n, 4, 3))
    P[::5, 1] = P[::5, 0] + 1e-13 * rng.standard_normal((len(P[::5]), 3))  # nearly degenerate / parallel cases
    d, w0, w1 = G.dist_face_vertex(P)
    e, rp, rq = G.dist_edge_edge(P)
    Q = P + 0.3 * rng.standard_normal((n, 4, 3))
    t = G.coplanar_interp(P, Q)
    def bits(*vals):  # bit patterns: the nearly degenerate inputs give inf / nan on both sides
        return np.array(vals, dtype=np.float64).tobytes()

    for k in range(n):
        assert bits(d[k], w0[k], w1[k]) == bits(*oracle.dist_fv(P[k])), k
        assert bits(e[k], rp[k], rq[k]) == bits(*oracle.dist_ee(P[k])), k
        assert bits(t[k]) == bits(oracle.coplanar(P[k], Q[k])), k
    assert (t >= 0).any() and (t < 0).any()


def test_make_normal(oracle, sc_contact):
    """clothgpu_make_normal == MakeNormal of the demos (self_contact_sparse/main.cpp:183-211), bit for bit."""
    sc = S.square_cloth_xz(30, 17, 0.1)
    x, v = T.wavy_state(sc)
    g = G.ClothGPU(sc, with_bvh=False)
    g.upload_state(x, v)
    nrm = g.make_normal(G.BUF_XYZ)
    assert np.array_equal(nrm, oracle.make_normal(sc, x))
    assert np.allclose(np.linalg.norm(nrm.reshape(-1, 3), axis=1), 1.0)
    g.close()


def test_device_side_setup(oracle):
    """Topology set-up on the device (csrc/setup.cu): clothgpu_build_adjacency gives the pattern SetPattern takes
    (SetEdgeOfElem on the hinge quads + Sort) and aEdge (on the triangles); the gather lists / diaind built on the device
    at context creation are exercised by every assembly test (bit-exact matrices) -- here once more on a renumbered
    mesh whose rows have scattered neighbours."""
    for sc in (S.square_cloth_xz(7, 5, 0.1), S.relabelled(S.square_cloth_xz(23, 30, 0.1)), S.config3_256()):
        for elem, nnoel, key_i, key_a in ((sc["quad"], 4, "rowptr", "colidx"), (sc["tri"], 3, "edge_index", "edge_array")):
            index, array = G.build_adjacency(elem, nnoel, sc["np"])
            assert np.array_equal(index, sc[key_i]) and np.array_equal(array, sc[key_a])
        if sc["np"] < 2000:
            oi, oa = oracle.edge_of_elem(sc["quad"], 4, sc["np"], True)
            index, array = G.build_adjacency(sc["quad"], 4, sc["np"])
            assert np.array_equal(index, oi) and np.array_equal(array, oa)
            di, da = G.build_adjacency(sc["tri"], 3, sc["np"], with_diagonal=True)
            assert len(da) == len(sc["edge_array"]) + sc["np"]
    with pytest.raises(G.ClothGPUError):
        G.build_adjacency(np.array([0, 1, 99], dtype=np.int32), 3, 5)
    sc = S.relabelled(S.square_cloth_xz(23, 30, 0.1))
    p = T.params_for(sc, S.SELF_CONTACT)
    x = sc["xyz0"] + 0.01 * np.random.default_rng(2).standard_normal(sc["np"] * 3)
    v = np.random.default_rng(3).standard_normal(sc["np"] * 3)
    g = G.ClothGPU(sc, with_bvh=False)
    g.upload_state(x, v)
    g.assemble(p)
    dia, crs, b = oracle.assemble(sc, x, v, p)
    assert np.array_equal(g.read(G.BUF_VAL_DIA), dia) and np.array_equal(g.read(G.BUF_VAL_CRS), crs)
    assert np.array_equal(g.read(G.BUF_VEC_R), b)
    g.close()


def test_fast_library_variant(oracle, tmp_path):
    """libclothgpu_fast.so (the same sources with multiply-add contraction on and tree reductions) is a
    separate, explicitly non-parity build: selected with CLOTHGPU_FAST=1, it reports its mode, and a contact-free
    implicit step agrees with the reference to rounding (1e-9) but not bit for bit."""
    import subprocess
    import sys
    code = ("import numpy as np, sys\n"
            "from cloth_sim_self_collision_b200 import clothgpu as G, scenes as S\n"
            "from tests import common as T\n"
            "assert G.load().clothgpu_build_mode().startswith(b'fast')\n"
            "sc = S.square_cloth_xy(24, 24 * 0.04)\n"
            "p = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=2, col_par=(0.4, 0.5, -0.6, 0.3))\n"
            "g = G.ClothGPU(sc, with_bvh=False)\n"
            "its = [int(g.step_internal(p)[0]) for _ in range(5)]\n"
            "x, v = g.download_state()\n"
            "np.savez(sys.argv[1], x=x, v=v, its=np.array(its))\n")
    out = tmp_path / "fast.npz"
    r = subprocess.run([sys.executable, "-c", code, str(out)], env=dict(os.environ, CLOTHGPU_FAST="1"),
                       cwd=os.path.dirname(GOLD[:-len("/golden")]), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = np.load(out)
    sc = S.square_cloth_xy(24, 24 * 0.04)
    p = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=2, col_par=(0.4, 0.5, -0.6, 0.3))
    sim = oracle.sim(sc)
    oits = [int(sim.step_internal(p)[1]) for _ in range(5)]
    ox, ov = sim.get_state()
    assert list(got["its"]) == oits
    assert rel_err(got["x"], ox) < RTOL and rel_err(got["v"], ov) < 1.0e-7
    assert not np.array_equal(got["v"], ov)
    assert G.load().clothgpu_build_mode().startswith(b"parity")  # this process runs the parity build


def test_certified_fast_arithmetic_selftest():
    """csrc/imp_fast.cuh: quotients / reciprocals / square roots of the impulse chain taken from early reciprocal
    estimates and Newton iterations, each with an exact-residual certificate.  A certified result must be the IEEE
    result bit for bit (0 mismatches over 2e8 operand pairs, half of them placed next to a rounding boundary); what
    cannot be certified falls back to the plain operation and must be rare."""
    import ctypes as C
    lib = G.load()
    out = (C.c_longlong * 10)()
    n = 200_000_000
    lib.clothgpu_selftest_certified_math.argtypes = [C.c_int, C.c_longlong, C.c_ulonglong, C.POINTER(C.c_longlong)]
    assert lib.clothgpu_selftest_certified_math(0, n, 20261018, out) == 0, lib.clothgpu_last_error()
    o = np.array(list(out), dtype=np.int64).reshape(5, 2)
    print("certified math: mismatches / uncertified per test:", o.tolist())
    assert (o[:, 0] == 0).all(), o
    assert (o[:, 1] < 0.01 * n).all(), o


def test_pair_rank_reconstruction_matches_the_traversal(oracle, monkeypatch):
    """The leaf-query traversal of the incremental CCD passes (k_query) does not carry the depth-first rank of a leaf
    pair; the narrow phase rebuilds it from the two root-to-leaf paths (pair_rank).  The rank only decides FV / EE key
    collisions, which trajectories rarely contain, so it is checked directly: with CLOTHGPU_RANK_CHECK=1 every
    narrow-phase hit of the full (frontier) passes is re-ranked and compared with the rank the traversal carried,
    together with the orientation rule of the query."""
    monkeypatch.setenv("CLOTHGPU_RANK_CHECK", "1")
    sc = _scene_with_bvh(oracle, 31, 40)
    p = T.params_for(sc, S.SELF_CONTACT)
    x0, v0 = T.folded_state(sc, gap=0.006, vel=5.0)
    g = G.ClothGPU(sc)
    g.upload_state(x0, v0)
    st = None
    for _ in range(4):
        st = g.step_self_contact(p, 1)[0]
    g.close()
    assert st[15] > 1000, "the scene must produce narrow-phase hits to check (%d)" % st[15]
    assert st[14] == 0, "pair_rank disagrees with the traversal on %d of %d hits" % (st[14], st[15])
