"""Per-launch kernel times of ONE batch step in the self-contact regime (CLOTHGPU_KTRACE): python tools/batch_trace.py [scenes] [setup]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CLOTHGPU_KTRACE"] = "1"
from cloth_sim_self_collision_b200 import batch as BT  # noqa: E402
from cloth_sim_self_collision_b200 import clothgpu as G  # noqa: E402
from cloth_sim_self_collision_b200 import scenes as S  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
setup = int(sys.argv[2]) if len(sys.argv) > 2 else 58
sc = S.config3_256()
sc["bvh"] = S.reference_bvh(sc)
c = S.SELF_CONTACT
p = G.make_params(c["dt"], c["lam"], c["myu"], c["stiff_bend"], c["gravity"], sc["mass_point"], c["stiff_contact"],
                  c["clearance"], c["col_kind"], c["col_par"])
g = G.ClothGPU(sc, nscene=B)
for s in range(B):
    g.set_scene_gravity(s, BT.scene_gravity(c["gravity"], s))
g.step_self_contact(p, setup)
g.step_self_contact(p, 1)
sys.stderr.write("ktrace ==== begin\n")
g.set_kernel_timing(True)
st = g.step_self_contact(p, 1)
g.set_kernel_timing(False)
sys.stderr.write("ktrace ==== end stats %s\n" % st[0][:6].tolist())
g.close()
