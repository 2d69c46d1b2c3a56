#!/usr/bin/env python
"""GPU probe: does the triangular solve slow down when the L2 is flushed between solves (as inside PCG, where the
SpMV streams the 61 MB matrix)?  Prints per-kernel averages with and without an SpMV between the solves."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloth_sim_self_collision_b200 import clothgpu as G, scenes as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sc = S.square_cloth_xz(n - 1, n - 1, 0.1, S.SELF_CONTACT["areal_density"])
g = G.ClothGPU(sc, with_bvh=False)
rng = np.random.default_rng(0)
g.write(G.BUF_VAL_DIA, np.tile(np.eye(3).ravel() * 20.0, sc["np"]))
g.write(G.BUF_VAL_CRS, rng.standard_normal(g.ncrs * 9) * 0.1)
g.ilu_setvalue()
g.ilu_factor()
g.write(G.BUF_VEC_Z, rng.standard_normal(sc["np"] * 3))
g.write(G.BUF_VEC_P, rng.standard_normal(sc["np"] * 3))
for mode in ("solve only", "matvec + solve"):
    for _ in range(2):
        g.ilu_solve(G.BUF_VEC_Z)
    g.set_kernel_timing(True)
    for _ in range(10):
        if mode != "solve only":
            g.matvec(1.0, G.BUF_VEC_P, 0.0, G.BUF_VEC_X)
        g.ilu_solve(G.BUF_VEC_Z)
    kt = g.kernel_times()
    g.set_kernel_timing(False)
    print(mode, {k: round(v[1] / v[0], 4) for k, v in kt.items()})
g.close()
