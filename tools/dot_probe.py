"""Times clothgpu_dot (the exact serial-order dot product) on 256^2-sized vectors; run under ncu for a profile."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloth_sim_self_collision_b200 import clothgpu as G  # noqa: E402
from cloth_sim_self_collision_b200 import scenes as S  # noqa: E402

sc = S.square_cloth_xz(255, 255, 0.1)
g = G.ClothGPU(sc, with_bvh=False)
rng = np.random.default_rng(0)
n = sc["np"] * 3
a = rng.standard_normal(n)
b = a * (1 + 0.1 * rng.standard_normal(n))
g.write(G.BUF_VEC_P, a)
g.write(G.BUF_VEC_Z, b)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
g.dot(G.BUF_VEC_P, G.BUF_VEC_Z)
g.set_kernel_timing(True)
for _ in range(reps):
    g.dot(G.BUF_VEC_P, G.BUF_VEC_Z)
for k, (cnt, ms) in sorted(g.kernel_times().items()):
    print("%-18s %4d launches  %.2f us each" % (k, cnt, ms * 1e3 / cnt))
g.set_kernel_timing(False)
t0 = time.perf_counter()
for _ in range(reps):
    g.dot(G.BUF_VEC_P, G.BUF_VEC_Z)
print("end to end incl. host sync + readback: %.1f us per dot" % ((time.perf_counter() - t0) * 1e6 / reps))
g.close()
