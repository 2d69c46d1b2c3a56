#!/usr/bin/env python
"""GPU probe: dependency-chain depth of the Gauss-Seidel impulse sweep (elements in std::set order, an element
depends on the previous element touching any of its four vertices) along the bench trajectory."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloth_sim_self_collision_b200 import clothgpu as G, scenes as S  # noqa: E402

n = 256
sc = S.square_cloth_xz(n - 1, n - 1, 0.1, S.SELF_CONTACT["areal_density"])
sc["bvh"] = S.reference_bvh(sc)
c = dict(S.SELF_CONTACT)
p = G.make_params(c["dt"], c["lam"], c["myu"], c["stiff_bend"], c["gravity"], sc["mass_point"], c["stiff_contact"],
                  c["clearance"], c["col_kind"], c["col_par"])
g = G.ClothGPU(sc)
for step in range(1, 70):
    st = g.step_self_contact(p, 1)[0]
    if step in (56, 60, 64, 66, 69):
        # proximity elements of the step just taken are gone; re-detect on the current state
        g.bvh_refit_prx(G.BUF_XYZ, c["clearance"])
        ne = g.contacts_prx(G.BUF_XYZ, c["clearance"])[0]
        e = g.elems(ne)
        last = {}
        depth = np.zeros(len(e), dtype=np.int64)
        for k, row in enumerate(e):
            d = 0
            for v in row[:4]:
                if v in last:
                    d = max(d, depth[last[v]])
            depth[k] = d + 1
            for v in row[:4]:
                last[int(v)] = k
        print("step %d: %d proximity elements, %d distinct vertices, chain depth %d" %
              (step, len(e), len(last), depth.max() if len(e) else 0))
g.close()
