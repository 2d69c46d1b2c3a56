#!/usr/bin/env python
"""CPU probe (oracle only): structure of the Gauss-Seidel impulse sweeps along the bench trajectory -- elements per
pass, how many of them write, depth of the full dependency chain and of the chain through writers only.
python tools/imp_dep_probe.py [extra_steps_after_57]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloth_sim_self_collision_b200 import scenes as S  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from tests import common as T  # noqa: E402

extra = int(sys.argv[1]) if len(sys.argv) > 1 else 9
d = np.load(os.path.join(ROOT, "gpurun_out", "parity_probe_57.npz"))
sc = S.config3_256()
sc["bvh"] = S.reference_bvh(sc)
p = T.params_for(sc, S.SELF_CONTACT)
lib = O.load_oracle()
sim = lib.sim(sc)
sim.set_state(d["x0"], d["v0"])
for k in range(extra):
    st = sim.step_self_contact(p)
    print("step", 57 + k, st[:6], st[8:14], flush=True)
x_prev, v = sim.get_state()
sim.step_internal(p)
x1, uvw = sim.get_state()


def analyse(name, elems, uvw, fn):
    ne = len(elems)
    last = {}
    lastw = {}
    depth = np.zeros(ne, dtype=np.int64)
    depthw = np.zeros(ne, dtype=np.int64)
    nw = 0
    cur = uvw.copy()
    for k, row in enumerate(elems):
        new = fn(cur, row[None, :])
        wrote = not np.array_equal(new, cur)
        cur = new
        dd = 0; dw = 0
        for vv in row[:4]:
            vv = int(vv)
            if vv in last: dd = max(dd, depth[last[vv]])
            if vv in lastw: dw = max(dw, depthw[lastw[vv]])
        depth[k] = dd + 1
        depthw[k] = dw + 1
        for vv in row[:4]:
            last[int(vv)] = k
            if wrote: lastw[int(vv)] = k
        nw += wrote
    print("%s: E %d  fv %d  writers %d  depth(all) %d  depth(through writers) %d" %
          (name, ne, int(elems[:, 4].sum()) if ne else 0, nw, depth.max() if ne else 0, depthw.max() if ne else 0), flush=True)
    return cur


dt, cl, m, k = p.dt, p.clearance, p.mass_point, p.stiff_contact
boxes = lib.refit_prx(sc, x_prev, cl)
e = lib.contacts_prx(sc, boxes, x_prev, cl)
uvw = analyse("prox", e, uvw, lambda u, el: lib.impulse_prx(u, cl, k, dt, m, x_prev, el))
for it in range(5):
    boxes = lib.refit_ccd(sc, x_prev, uvw, dt)
    e = lib.contacts_ccd(sc, boxes, x_prev, uvw, dt, cl)
    if len(e) == 0:
        break
    uvw = analyse("ccd %d" % it, e, uvw, lambda u, el: lib.impulse_ccd(u, cl, k, dt, m, x_prev, el))


def distances(name, elems):
    """distribution of the distance (in elements) to the predecessor on each vertex"""
    last = {}
    far = {256: 0, 1024: 0, 4096: 0, 16384: 0}
    anyfar = {256: 0, 1024: 0, 4096: 0, 16384: 0}
    for k, row in enumerate(elems):
        dmax = 0
        for vv in row[:4]:
            vv = int(vv)
            if vv in last:
                dmax = max(dmax, k - last[vv])
            last[vv] = k
        for t in far:
            if dmax >= t:
                anyfar[t] += 1
    print("%s: %d elements; elements whose farthest predecessor is >= 256 / 1024 / 4096 / 16384 back: %s" %
          (name, len(elems), [anyfar[t] for t in (256, 1024, 4096, 16384)]), flush=True)


if os.environ.get("PROBE_DIST"):
    boxes = lib.refit_prx(sc, x_prev, cl)
    distances("prox", lib.contacts_prx(sc, boxes, x_prev, cl))
