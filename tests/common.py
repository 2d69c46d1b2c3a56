"""Shared helpers for the test-suite: scenes, parameters, deformed states."""
import numpy as np

from cloth_sim_self_collision_b200 import scenes as S
from oracle import oracle_py as O


def params_for(sc, consts, **over):
    c = dict(consts)
    c.update(over)
    return O.Params(c["dt"], c["lam"], c["myu"], c["stiff_bend"], c["gravity"], sc["mass_point"],
                    c["stiff_contact"], c["clearance"], c["col_kind"], c["col_par"])


def wavy_state(sc, amp=0.3, seed=1, vel=0.5):
    """A smooth out-of-plane deformation + random velocities: exercises every energy term."""
    rng = np.random.default_rng(seed)
    x = np.array(sc["xyz0"]).reshape(-1, 3).copy()
    h = sc["h"]
    u = x / (h * 7.0)
    normal_axis = int(np.argmin(np.ptp(x, axis=0)))
    planar = [a for a in range(3) if a != normal_axis]
    x[:, normal_axis] += amp * h * (np.sin(u[:, planar[0]] * 1.3) + np.cos(u[:, planar[1]] * 0.7 + 0.3))
    x[:, planar[0]] *= 0.97
    x += 0.05 * h * rng.standard_normal(x.shape)
    v = vel * rng.standard_normal(x.shape)
    return x.ravel().copy(), v.ravel().copy()


def folded_state(sc, seed=3, gap=0.006, vel=0.3):
    """Fold the sheet over itself along its long axis so that layers sit within the contact clearance and move
    towards each other: produces proximity contacts, CCD contacts and impact zones on small scenes."""
    rng = np.random.default_rng(seed)
    x0 = np.array(sc["xyz0"]).reshape(-1, 3)
    h = sc["h"]
    ext = np.ptp(x0, axis=0)
    normal_axis = int(np.argmin(ext))
    long_axis = int(np.argmax(ext))
    x = x0.copy()
    L = ext[long_axis]
    s = x0[:, long_axis] - x0[:, long_axis].min()
    half = 0.5 * L
    over = s > half
    # layer 2 lies on top of layer 1, mirrored, at height 2r (slightly staggered so vertices do not coincide)
    x[over, long_axis] = x0[:, long_axis].min() + half - (s[over] - half) + 0.37 * h
    x[over, normal_axis] += gap
    x += 0.002 * h * rng.standard_normal(x.shape)
    v = vel * 0.1 * rng.standard_normal(x.shape)
    v[over, normal_axis] -= vel
    v[~over, normal_axis] += vel
    return x.ravel().copy(), v.ravel().copy()


def small_self_contact_scene(ref_or_none=None, nx=4, nz=20):
    sc = S.square_cloth_xz(nx, nz, 0.1, S.SELF_CONTACT["areal_density"])
    return sc
