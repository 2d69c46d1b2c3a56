#!/usr/bin/env python
"""GPU probe: per-step device time and workload statistics along the bench trajectory (config 3)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloth_sim_self_collision_b200 import clothgpu as G, scenes as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 70
sc = S.square_cloth_xz(n - 1, n - 1, 0.1, S.SELF_CONTACT["areal_density"])
sc["bvh"] = S.reference_bvh(sc)
c = dict(S.SELF_CONTACT)
p = G.make_params(c["dt"], c["lam"], c["myu"], c["stiff_bend"], c["gravity"], sc["mass_point"], c["stiff_contact"],
                  c["clearance"], c["col_kind"], c["col_par"])
g = G.ClothGPU(sc)
for step in range(1, nsteps + 1):
    g.timer_start()
    st = g.step_self_contact(p, 1)[0]
    ms = g.timer_stop()
    if step > 28:
        print("step %3d %7.2f ms  pcg %d  prox %6d  ccd passes %d  riz rounds %3d  ccd elems %s" %
              (step, ms, st[0], st[1], st[2], st[3], list(st[8:8 + max(st[2], 1)])))
g.close()
