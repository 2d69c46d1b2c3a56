#!/usr/bin/env python
"""GPU probe: per-phase device time (CUDA events around the phases, kernels NOT individually timed) of the bench scene
over a window of steps: python tools/phase_probe.py [first] [last]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloth_sim_self_collision_b200 import clothgpu as G, scenes as S  # noqa: E402

first = int(sys.argv[1]) if len(sys.argv) > 1 else 41
last = int(sys.argv[2]) if len(sys.argv) > 2 else 50
sc = S.config3_256()
sc["bvh"] = S.reference_bvh(sc)
c = S.SELF_CONTACT
p = G.make_params(c["dt"], c["lam"], c["myu"], c["stiff_bend"], c["gravity"], sc["mass_point"], c["stiff_contact"],
                  c["clearance"], c["col_kind"], c["col_par"])
g = G.ClothGPU(sc)
g.step_self_contact(p, first - 1)
n = last - first + 1
g.timer_start()
g.step_self_contact(p, n)
plain = g.timer_stop() / n
g.close()
g = G.ClothGPU(sc)
g.step_self_contact(p, first - 1)
g.set_profiling(True)
g.timer_start()
g.step_self_contact(p, n)
prof = g.timer_stop() / n
ph = g.phase_ms()
print("steps %d..%d: %.3f ms/step plain, %.3f ms/step with phase events" % (first, last, plain, prof))
print({k: round(v / n, 3) for k, v in ph.items()})
g.close()
