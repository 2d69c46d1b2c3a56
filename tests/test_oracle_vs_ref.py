"""Pins the C restatement (oracle/cloth_oracle.c) to the UNMODIFIED reference (oracle/_ref/libclothref.so).

Every comparison is bit-exact: both sides are fp64, single-threaded and compiled with -ffp-contract=off.
Skipped when the reference library has not been built (it needs /root/reference at build time; the prebuilt
.so travels with the repo snapshot).
"""
import numpy as np
import pytest

from cloth_sim_self_collision_b200 import scenes as S
from tests import common as T


def _bvh(ref, sc):
    if sc["bvh"] is None:
        sc["bvh"] = ref.bvh_build(sc["tri"], S.jittered_rest(sc))
    return sc


@pytest.fixture(scope="module")
def sc_contact(ref):
    return _bvh(ref, S.square_cloth_xz(4, 20, 0.1, S.SELF_CONTACT["areal_density"]))


@pytest.fixture(scope="module")
def sc_internal(ref):
    return _bvh(ref, S.square_cloth_xy(12, 1.0))


def test_element_energies_bit_exact(ref, oracle):
    rng = np.random.default_rng(0)
    for k in range(200):
        C = rng.standard_normal((3, 3))
        c = C + 0.3 * rng.standard_normal((3, 3)) * (k % 3)
        lam, myu = (0.0, 30.0) if k % 2 else (1.0, 4.0)
        a = ref.cst(C, c, lam, myu)
        b = oracle.cst(C, c, lam, myu)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        C4 = rng.standard_normal((4, 3))
        c4 = C4 + 0.2 * rng.standard_normal((4, 3))
        a = ref.bend(C4, c4, 1.0e-2)
        b = oracle.bend(C4, c4, 1.0e-2)
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for kind, par in [(0, (0, 0, 0, 0)), (1, (-0.5, 0, 0, 0)), (2, (0.1, 0.5, -0.8, 0.3)), (3, (0.5, 0.0, 2.0, 3.0))]:
        for _ in range(50):
            p = rng.standard_normal(3) * 2
            a = ref.contact(p, 300.0, 0.01, kind, par)
            b = oracle.contact(p, 300.0, 0.01, kind, par)
            assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_setup_helpers(ref, oracle, sc_contact):
    sc = sc_contact
    for elem, nnoel, sort in [(sc["quad"], 4, True), (sc["tri"], 3, False)]:
        a = ref.edge_of_elem(elem, nnoel, sc["np"], sort)
        b = oracle.edge_of_elem(elem, nnoel, sc["np"], sort)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(ref.ilu_diaind(sc["rowptr"], sc["colidx"]), oracle.ilu_diaind(sc["rowptr"], sc["colidx"]))
    # the numpy scene generator builds the same pattern / adjacency (as sets per row)
    a = ref.edge_of_elem(sc["quad"], 4, sc["np"], True)
    assert np.array_equal(a[0], sc["rowptr"]) and np.array_equal(a[1], sc["colidx"])
    a = ref.edge_of_elem(sc["tri"], 3, sc["np"], False)
    assert np.array_equal(a[0], sc["edge_index"])
    for i in range(sc["np"]):
        assert set(a[1][a[0][i]:a[0][i + 1]]) == set(sc["edge_array"][sc["edge_index"][i]:sc["edge_index"][i + 1]])


@pytest.mark.parametrize("which", ["contact", "internal_sphere"])
def test_linear_system_pipeline_bit_exact(ref, oracle, sc_contact, sc_internal, which):
    if which == "contact":
        sc = sc_contact
        p = T.params_for(sc, S.SELF_CONTACT, col_par=(0.2, 0, 0, 0))  # plane lifted so that some vertices touch
    else:
        sc = sc_internal
        p = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=2, col_par=(0.3, 0.5, -0.1, 0.3))
    x, v = T.wavy_state(sc)
    A = ref.assemble(sc, x, v, p)
    B = oracle.assemble(sc, x, v, p)
    for a, b in zip(A, B):
        assert np.array_equal(a, b)
    dia, crs, rhs = A
    assert np.abs(crs).max() > 0 and np.abs(rhs).max() > 0
    rng = np.random.default_rng(5)
    xx = rng.standard_normal(sc["np"] * 3)
    yy = rng.standard_normal(sc["np"] * 3)
    assert np.array_equal(ref.matvec(sc, dia, crs, 1.0, xx, 0.0, yy), oracle.matvec(sc, dia, crs, 1.0, xx, 0.0, yy))
    assert np.array_equal(ref.matvec(sc, dia, crs, -0.7, xx, 1.3, yy), oracle.matvec(sc, dia, crs, -0.7, xx, 1.3, yy))
    assert ref.dot(xx, yy) == oracle.dot(xx, yy)
    fa = ref.ilu_factor(sc, dia, crs)
    fb = oracle.ilu_factor(sc, dia, crs)
    assert np.array_equal(fa[0], fb[0]) and np.array_equal(fa[1], fb[1])
    assert np.array_equal(ref.ilu_solve(sc, fa[0], fa[1], xx), oracle.ilu_solve(sc, fa[0], fa[1], xx))
    pa = ref.pcg(sc, dia, crs, fa[0], fa[1], rhs)
    pb = oracle.pcg(sc, dia, crs, fa[0], fa[1], rhs)
    assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1]) and pa[2:] == pb[2:]
    assert 0 <= pa[3] < 100 and pa[2] < 1.0e-4
    ca = ref.cg(sc, dia, crs, rhs, 1.0e-4, 1000)
    cb = oracle.cg(sc, dia, crs, rhs, 1.0e-4, 1000)
    assert np.array_equal(ca[0], cb[0]) and np.array_equal(ca[1], cb[1]) and ca[2:] == cb[2:]
    assert np.array_equal(ref.update_intermediate_velocity(sc, x, v, p), oracle.update_intermediate_velocity(sc, x, v, p))


def test_pcg_zero_rhs_early_out(ref, oracle, sc_contact):
    sc = sc_contact
    p = T.params_for(sc, S.SELF_CONTACT)
    x, v = T.wavy_state(sc)
    dia, crs, _ = ref.assemble(sc, x, v, p)
    f = ref.ilu_factor(sc, dia, crs)
    z = np.zeros(sc["np"] * 3)
    a = ref.pcg(sc, dia, crs, f[0], f[1], z)
    b = oracle.pcg(sc, dia, crs, f[0], f[1], z)
    assert a[2:] == b[2:] == (0.0, 0)


def test_predicates_bit_exact(ref, oracle):
    rng = np.random.default_rng(11)
    n_true = [0, 0, 0, 0]
    for k in range(3000):
        scale = [1.0, 0.1, 0.01][k % 3]
        tri = rng.standard_normal((3, 3)) * 0.1
        bary = rng.dirichlet((1, 1, 1))
        nrm = np.cross(tri[1] - tri[0], tri[2] - tri[0])
        nrm /= np.linalg.norm(nrm)
        p3 = bary @ tri + nrm * 0.008 * rng.standard_normal() + 0.01 * scale * rng.standard_normal(3)
        P = np.vstack([tri, p3])
        Q = P + 0.02 * rng.standard_normal((4, 3)) + nrm * 0.02 * np.array([[1], [1], [1], [-1]]) * np.sign(nrm @ (tri[0] - p3))
        box = np.array([P[:3, 0].min() - .005, P[:3, 0].max() + .005, P[:3, 1].min() - .005, P[:3, 1].max() + .005,
                        P[:3, 2].min() - .005, P[:3, 2].max() + .005])
        ids = [0, 1, 2, 3]
        assert ref.dist_fv(P) == oracle.dist_fv(P)
        assert ref.dist_ee(P) == oracle.dist_ee(P)
        assert ref.coplanar(P, Q) == oracle.coplanar(P, Q)
        r = [ref.is_fv_prx(ids, P, box, 0.01), ref.is_ee_prx(ids, P, 0.01), ref.is_fv_ccd(ids, P, Q, box),
             ref.is_ee_ccd(ids, P, Q)]
        o = [oracle.is_fv_prx(ids, P, box, 0.01), oracle.is_ee_prx(ids, P, 0.01), oracle.is_fv_ccd(ids, P, Q, box),
             oracle.is_ee_ccd(ids, P, Q)]
        assert r == o
        for i in range(4):
            n_true[i] += r[i]
    assert min(n_true) > 20, n_true  # every predicate answered "contact" often enough to mean something
    # parallel edges take the special branch of DistanceEdgeEdge
    for k in range(100):
        p0 = rng.standard_normal(3); d = rng.standard_normal(3)
        off = rng.standard_normal(3) * 0.01
        P = np.vstack([p0, p0 + d, p0 + off + 0.3 * d, p0 + off + (0.3 + rng.uniform(0.1, 2)) * d])
        assert ref.dist_ee(P) == oracle.dist_ee(P)


def test_contact_detection_and_response_bit_exact(ref, oracle, sc_contact):
    sc = sc_contact
    c = S.SELF_CONTACT
    x, v = T.folded_state(sc, gap=0.004, vel=10.0)
    bp_a = ref.refit_prx(sc, x, c["clearance"])
    bp_b = oracle.refit_prx(sc, x, c["clearance"])
    assert np.array_equal(bp_a, bp_b)
    lp_a = ref.leaf_pairs(sc, bp_a)
    lp_b = oracle.leaf_pairs(sc, bp_a)
    assert lp_a.shape[0] > sc["ntri"] and np.array_equal(lp_a, lp_b)
    ea = ref.contacts_prx(sc, bp_a, x, c["clearance"])
    eb = oracle.contacts_prx(sc, bp_a, x, c["clearance"])
    assert ea.shape[0] > 10 and np.array_equal(ea, eb)
    assert 0 < ea[:, 4].sum() < ea.shape[0]  # both FV and EE elements present
    va = ref.impulse_prx(v, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"], x, ea)
    vb = oracle.impulse_prx(v, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"], x, ea)
    assert np.array_equal(va, vb) and not np.array_equal(va, v)
    bc_a = ref.refit_ccd(sc, x, va, c["dt"])
    bc_b = oracle.refit_ccd(sc, x, va, c["dt"])
    assert np.array_equal(bc_a, bc_b)
    ca = ref.contacts_ccd(sc, bc_a, x, va, c["dt"], c["clearance"])
    cb = oracle.contacts_ccd(sc, bc_a, x, va, c["dt"], c["clearance"])
    assert ca.shape[0] > 5 and np.array_equal(ca, cb)
    wa = ref.impulse_ccd(va, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"], x, ca)
    wb = oracle.impulse_ccd(va, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"], x, ca)
    assert np.array_equal(wa, wb) and not np.array_equal(wa, va)
    # impact zones: three rounds on growing element sets, membership + velocities identical
    ra, rb = ref.riz_new(), oracle.riz_new()
    cur_a = wa.copy()
    cur_b = wa.copy()
    for rnd in range(3):
        elems = ca[rnd::3] if rnd < 2 else ca
        ra.make(elems, sc)
        rb.make(elems, sc)
        na, za = ra.zones(sc["np"])
        nb, zb = rb.zones(sc["np"])
        assert na == nb and np.array_equal(za, zb) and na > 0
        cur_a = ra.apply(cur_a, x, wa)
        cur_b = rb.apply(cur_b, x, wa)
        assert np.array_equal(cur_a, cur_b)
    assert not np.array_equal(cur_a, wa)


def test_resolve_contacts_orchestration_bit_exact(ref, oracle, sc_contact):
    sc = sc_contact
    p = T.params_for(sc, S.SELF_CONTACT)
    x, v = T.folded_state(sc, gap=0.004, vel=10.0)
    sa, so = ref.sim(sc), oracle.sim(sc)
    va, aa, sta = sa.resolve_contacts(v, x, p)
    vb, ab, stb = so.resolve_contacts(v, x, p)
    assert aa == ab and np.array_equal(sta, stb) and np.array_equal(va, vb)
    assert sta[1] > 0 and sta[2] == 5 and sta[3] >= 2  # proximity hits, all CCD passes used, impact zones entered


def test_default_scene_trajectory_bit_exact(ref, oracle):
    """Config 1 (self_contact_sparse default scene), free-running from rest, incl. the survey's known answers."""
    sc = S.config1_default()
    sc["bvh"] = ref.bvh_build(sc["tri"], sc["xyz0"])  # the reference builder works on this 5 x 51 grid as is
    p = T.params_for(sc, S.SELF_CONTACT)
    sa, so = ref.sim(sc), oracle.sim(sc)
    hist = {}
    riz_steps = 0
    for step in range(200):
        sta = sa.step_self_contact(p)
        stb = so.step_self_contact(p)
        assert np.array_equal(sta, stb), step
        hist[int(sta[0])] = hist.get(int(sta[0]), 0) + 1
        riz_steps += sta[3] > 0
        if step % 20 == 19 or sta[3] > 0:
            xa, va = sa.get_state()
            xb, vb = so.get_state()
            assert np.array_equal(xa, xb) and np.array_equal(va, vb), step
    xa, va = sa.get_state()
    checksum = float(sum(xa[i] * (1 + i % 7) for i in range(xa.size)))
    assert checksum == -89.699184471192837      # SURVEY.md section 8(c) known answer
    assert hist == {6: 94, 8: 68, 7: 21, 5: 17}
    assert riz_steps == 73


def test_internal_cloth_trajectory_bit_exact(ref, oracle):
    """Config 2 family (internal_cloth_sparse): pinned row, sphere obstacle."""
    sc = S.square_cloth_xy(16, 1.0)
    p = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=2, col_par=(0.1, 0.5, -0.8, 0.3))
    sa, so = ref.sim(sc), oracle.sim(sc)
    for step in range(40):
        ra = sa.step_internal(p)
        rb = so.step_internal(p)
        assert ra[1] == rb[1] and abs(ra[0] - rb[0]) <= 1e-5 * rb[0]  # the reference prints 6 digits only
    xa, va = sa.get_state()
    xb, vb = so.get_state()
    assert np.array_equal(xa, xb) and np.array_equal(va, vb)
    assert np.abs(xa - sc["xyz0"]).max() > 0.1


def test_internal_cloth_cg_trajectory_bit_exact(ref, oracle):
    """StepTime_InternalDynamics (solve_internal_sparse.h:185-253, the unpreconditioned Solve_CG variant) and the
    SPHERE_INSIDE collider of self_contact_sparse/main.cpp:117-133 inside a step."""
    sc = S.square_cloth_xy(12, 1.0)
    for kind, par in ((2, (0.1, 0.5, -0.8, 0.3)), (3, (0.5, 0.5, -0.2, 0.9))):
        p = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=kind, col_par=par)
        sa, so = ref.sim(sc), oracle.sim(sc)
        for step in range(6):
            ra = sa.step_internal_cg(p)
            rb = so.step_internal_cg(p)
            assert ra[1] == rb[1] and abs(ra[0] - rb[0]) <= 1e-5 * rb[0]
        xa, va = sa.get_state()
        xb, vb = so.get_state()
        assert np.array_equal(xa, xb) and np.array_equal(va, vb)
        assert ra[1] > 3
