"""Non-parity fast mode (libclothgpu_fast.so: FMA contraction on, tree reductions) against the parity build: ms per
step and the deviation from the parity trajectory.  usage: python tools/fast_mode_probe.py [grid=256] [setup=40] [steps=10]
(re-executes itself with CLOTHGPU_FAST=0 / 1; prints one JSON line per build)"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(grid, setup, steps, out):
    from cloth_sim_self_collision_b200 import clothgpu as G
    from cloth_sim_self_collision_b200 import scenes as S
    sc = S.square_cloth_xz(grid - 1, grid - 1, 0.1, S.SELF_CONTACT["areal_density"])
    sc["bvh"] = S.reference_bvh(sc)
    c = S.SELF_CONTACT
    p = G.make_params(c["dt"], c["lam"], c["myu"], c["stiff_bend"], c["gravity"], sc["mass_point"], c["stiff_contact"],
                      c["clearance"], c["col_kind"], c["col_par"])
    g = G.ClothGPU(sc)
    g.step_self_contact(p, setup)
    x0, v0 = g.download_state()
    g.step_self_contact(p, steps)  # rehearsal
    g.upload_state(x0, v0)
    g.synchronize()
    g.timer_start()
    st = g.step_self_contact(p, steps)
    ms = g.timer_stop() / steps
    x, v = g.download_state()
    np.savez(out, x=x, v=v)
    print(json.dumps({"build": G.load().clothgpu_build_mode().decode().split(":")[0], "grid": grid,
                      "trajectory_steps": "%d..%d" % (setup + 1, setup + steps), "ms_per_step": ms,
                      "stats_last": [int(k) for k in st[0][:5]]}), flush=True)
    g.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
        sys.exit(0)
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    setup = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    outs = []
    for fast in ("0", "1"):
        out = "/tmp/fast_mode_probe_%s.npz" % fast
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child", str(grid), str(setup), str(steps), out],
                              env=dict(os.environ, CLOTHGPU_FAST=fast))
        outs.append(np.load(out))
    rx = float(np.abs(outs[1]["x"] - outs[0]["x"]).max() / np.abs(outs[0]["x"]).max())
    rv = float(np.abs(outs[1]["v"] - outs[0]["v"]).max() / np.abs(outs[0]["v"]).max())
    print(json.dumps({"fast_vs_parity_after_%d_steps" % (setup + steps): {"rel_x": rx, "rel_v": rv}}))
