"""How sensitive is the REFERENCE's own self-collision response to rounding of its input?  (CPU only.)

Runs the unmodified reference (oracle/_ref; the C restatement when it is absent) from rest to the start of a step in
the self-contact regime of the 256 x 256 bench scene, performs the implicit solve, and then calls
GetIntermidiateVelocityContactResolved (self_collision_cloth.cpp:355-456) on the resulting velocities and on copies
perturbed by a relative 1e-16 .. 1e-11.  The output moves by ~1e-2 of max|v| for every perturbation size, with
identical contact statistics: the response is discontinuous in its input (coplanarity cubics of nearly flat regions
are rounding noise), so a whole step can only be compared with the reference if the velocities entering the response
are bit-identical -- which is why libclothgpu reproduces the reference's serial dot products exactly.

usage: python tools/reference_sensitivity.py [step=57] [grid=256] > profiles/r02_reference_sensitivity.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloth_sim_self_collision_b200 import scenes as S  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from tests import common as T  # noqa: E402

step = int(sys.argv[1]) if len(sys.argv) > 1 else 57
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 256
lib = O.load_ref()
kind = "reference (oracle/_ref)"
if lib is None:
    lib, kind = O.load_oracle(), "C restatement (oracle/cloth_oracle.c)"
sc = S.square_cloth_xz(grid - 1, grid - 1, 0.1, S.SELF_CONTACT["areal_density"])
sc["bvh"] = S.reference_bvh(sc)
p = T.params_for(sc, S.SELF_CONTACT)
sim = lib.sim(sc)
for _ in range(step - 1):
    sim.step_self_contact(p)
x0, v0 = sim.get_state()
sim.step_internal(p)
x1, v1 = sim.get_state()


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


base, app, st = sim.resolve_contacts(v1, x0, p)
print("checker: %s; %dx%d cloth, trajectory step %d" % (kind, grid, grid, step))
print("stats (pcg, prox elems, ccd passes, riz rounds, applied):", st[:5].tolist(), "ccd elems per pass:", st[8:13].tolist())
print("response changed the velocities by (rel. to max|v|):", rel(base, v1))
rng = np.random.default_rng(0)
for eps in (1e-16, 1e-15, 1e-13, 1e-11):
    vp = v1 * (1 + eps * rng.standard_normal(v1.size))
    out, app, st2 = sim.resolve_contacts(vp, x0, p)
    dv = np.abs(out - base)
    print("input perturbed by rel %.1e (max rel %.2e) -> output differs by rel %.3e, %d components above 1e-9, "
          "statistics equal: %s" % (eps, rel(vp, v1), rel(out, base), int((dv > 1e-9 * np.abs(base).max()).sum()),
                                    bool(np.array_equal(st2[:16], st[:16]))))
