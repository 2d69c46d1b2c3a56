#!/usr/bin/env python
"""ncu target: the bench scene up to step N-1 unprofiled, step N between cudaProfilerStart/Stop.
ncu --profile-from-start off --kernel-name regex:k_imp_flow --launch-count 2 --set full --import-source on ... python tools/imp_profile.py [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cloth_sim_self_collision_b200 import clothgpu as G, scenes as S  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65
sc = S.config3_256()
sc["bvh"] = S.reference_bvh(sc)
c = S.SELF_CONTACT
p = G.make_params(c["dt"], c["lam"], c["myu"], c["stiff_bend"], c["gravity"], sc["mass_point"], c["stiff_contact"],
                  c["clearance"], c["col_kind"], c["col_par"])
g = G.ClothGPU(sc)
g.step_self_contact(p, n - 1)
rt = torch.cuda.cudart()
torch.cuda.synchronize()
rt.cudaProfilerStart()
st = g.step_self_contact(p, 1)[0]
torch.cuda.synchronize()
rt.cudaProfilerStop()
print("step", n, st[:6], st[8:13])
g.close()
