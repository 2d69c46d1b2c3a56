#!/usr/bin/env python
"""GPU probe: ms per scene-step for batches of independent scenes on one GPU (config 5 on a single device)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloth_sim_self_collision_b200 import clothgpu as G, scenes as S, batch as BT  # noqa: E402

n = 256
sc = S.square_cloth_xz(n - 1, n - 1, 0.1, S.SELF_CONTACT["areal_density"])
sc["bvh"] = S.reference_bvh(sc)
c = dict(S.SELF_CONTACT)
p = G.make_params(c["dt"], c["lam"], c["myu"], c["stiff_bend"], c["gravity"], sc["mass_point"], c["stiff_contact"],
                  c["clearance"], c["col_kind"], c["col_par"])
for B in [int(a) for a in sys.argv[1:]] or [1, 4, 8, 16, 32]:
    g = G.ClothGPU(sc, nscene=B)
    for s in range(B):
        g.set_scene_gravity(s, BT.scene_gravity(c["gravity"], s))
    g.step_self_contact(p, 40)
    g.timer_start()
    g.step_self_contact(p, 10)
    ms = g.timer_stop() / 10
    print("B=%3d  %8.2f ms per batch step  %6.2f ms per scene-step" % (B, ms, ms / B), flush=True)
    g.close()
