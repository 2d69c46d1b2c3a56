"""The C restatement (oracle/cloth_oracle.c) against the committed golden vectors.

tests/golden/*.npz were produced from the unmodified reference by oracle/make_golden.py; this file needs neither
/root/reference nor the compiled reference, so it pins the oracle on any box.  All comparisons are bit-exact.
"""
import os

import numpy as np
import pytest

from cloth_sim_self_collision_b200 import scenes as S
from tests import common as T

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def phases():
    return np.load(os.path.join(GOLD, "phases_5x21.npz"))


def _scene_5x21(g):
    sc = S.square_cloth_xz(4, 20, 0.1, S.SELF_CONTACT["areal_density"])
    sc["bvh"] = np.ascontiguousarray(g["bvh"])
    return sc


def test_linear_system_phases(oracle, phases):
    g = phases
    sc = _scene_5x21(g)
    p = T.params_for(sc, S.SELF_CONTACT, col_par=tuple(g["col_par"]))
    x, v = g["x"], g["v"]
    gx, gv = T.wavy_state(sc)
    assert np.array_equal(gx, x) and np.array_equal(gv, v)  # the state generator itself is reproducible
    dia, crs, b = oracle.assemble(sc, x, v, p)
    assert np.array_equal(dia, g["dia"]) and np.array_equal(crs, g["crs"]) and np.array_equal(b, g["b"])
    idia, icrs = oracle.ilu_factor(sc, dia, crs)
    assert np.array_equal(idia, g["idia"]) and np.array_equal(icrs, g["icrs"])
    assert np.array_equal(oracle.matvec(sc, dia, crs, 1.0, g["vec"], 0.0, g["yv"]), g["matvec"])
    assert np.array_equal(oracle.matvec(sc, dia, crs, -0.7, g["vec"], 1.3, g["yv"]), g["matvec_ab"])
    assert np.array_equal(oracle.ilu_solve(sc, idia, icrs, g["vec"]), g["ilu_solve"])
    px, pr, conv, it = oracle.pcg(sc, dia, crs, idia, icrs, b)
    assert np.array_equal(px, g["pcg_x"]) and np.array_equal(pr, g["pcg_r"])
    assert conv == float(g["pcg_conv"]) and it == int(g["pcg_iter"])
    assert np.array_equal(oracle.update_intermediate_velocity(sc, x, v, p), g["upd_v"])


def test_contact_phases(oracle, phases):
    g = phases
    sc = _scene_5x21(g)
    c = S.SELF_CONTACT
    fx, fv = g["fx"], g["fv"]
    bp = oracle.refit_prx(sc, fx, c["clearance"])
    assert np.array_equal(bp, g["box_prx"])
    assert np.array_equal(oracle.leaf_pairs(sc, bp), g["leaf_pairs_prx"])
    ep = oracle.contacts_prx(sc, bp, fx, c["clearance"])
    assert np.array_equal(ep, g["elems_prx"])
    v1 = oracle.impulse_prx(fv, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"], fx, ep)
    assert np.array_equal(v1, g["v_after_prx"])
    bc = oracle.refit_ccd(sc, fx, v1, c["dt"])
    assert np.array_equal(bc, g["box_ccd"])
    assert np.array_equal(oracle.leaf_pairs(sc, bc), g["leaf_pairs_ccd"])
    ec = oracle.contacts_ccd(sc, bc, fx, v1, c["dt"], c["clearance"])
    assert np.array_equal(ec, g["elems_ccd"])
    v2 = oracle.impulse_ccd(v1, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"], fx, ec)
    assert np.array_equal(v2, g["v_after_ccd"])
    riz = oracle.riz_new()
    riz.make(ec, sc)
    nz, zones = riz.zones(sc["np"])
    assert nz == int(g["riz_nzone"]) and np.array_equal(zones, g["riz_zone"])
    assert np.array_equal(riz.apply(v2, fx, v2), g["v_after_riz"])
    vr, applied, stats = oracle.sim(sc).resolve_contacts(fv, fx, T.params_for(sc, c))
    assert np.array_equal(vr, g["resolve_v"]) and int(applied) == int(g["resolve_applied"])
    assert np.array_equal(stats, g["resolve_stats"])


def test_config1_trajectory(oracle):
    g = np.load(os.path.join(GOLD, "traj_config1.npz"))
    sc = S.config1_default()
    sc["bvh"] = np.ascontiguousarray(g["bvh"])
    p = T.params_for(sc, S.SELF_CONTACT)
    sim = oracle.sim(sc)
    keep = list(g["steps"])
    for step in range(1, 201):
        st = sim.step_self_contact(p)
        assert np.array_equal(st[:40], g["stats"][step - 1]), step
        if step in keep:
            x, v = sim.get_state()
            k = keep.index(step)
            assert np.array_equal(x, g["x"][k]) and np.array_equal(v, g["v"][k]), step
    x, _ = sim.get_state()
    assert float(sum(x[i] * (1 + i % 7) for i in range(x.size))) == float(g["checksum"]) == -89.699184471192837


def test_internal_cloth_trajectory(oracle):
    g = np.load(os.path.join(GOLD, "traj_internal17.npz"))
    sc = S.square_cloth_xy(16, 1.0)
    p = T.params_for(sc, S.INTERNAL_CLOTH, col_kind=2, col_par=(0.1, 0.5, -0.8, 0.3))
    sim = oracle.sim(sc)
    keep = list(g["steps"])
    for step in range(1, 41):
        _, it = sim.step_internal(p)
        assert it == g["iters"][step - 1]
        if step in keep:
            x, v = sim.get_state()
            k = keep.index(step)
            assert np.array_equal(x, g["x"][k]) and np.array_equal(v, g["v"][k])
