#!/usr/bin/env python
"""GPU probe: config 4 (512 x 512 vertices, synthetic crumpled start): ms per step, 10 steps."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloth_sim_self_collision_b200 import clothgpu as G, scenes as S  # noqa: E402

sc = S.config4_512()
sc["bvh"] = S.reference_bvh(sc)
c = dict(S.SELF_CONTACT)
p = G.make_params(c["dt"], c["lam"], c["myu"], c["stiff_bend"], c["gravity"], sc["mass_point"], c["stiff_contact"],
                  c["clearance"], c["col_kind"], c["col_par"])
x, v = S.crumpled_state(sc)
g = G.ClothGPU(sc)
for rep in range(2):  # the first run sizes the contact queues
    g.upload_state(x, v)
    ms = []
    for step in range(10):
        g.timer_start()
        st = g.step_self_contact(p, 1)[0]
        ms.append(g.timer_stop())
    if rep == 1:
        print("config 4 (512^2 crumpled): %.1f ms per step (mean of 10), per step %s" % (np.mean(ms), np.round(ms, 1)))
        print("  last step: pcg %d, prox %d, ccd passes %d, riz rounds %d" % (st[0], st[1], st[2], st[3]))
g.close()
