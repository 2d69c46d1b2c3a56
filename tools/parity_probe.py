"""Phase-by-phase parity probe of one self-contact step of the 256^2 bench scene (GPU vs the CPU checker).
usage: python tools/parity_probe.py [step] [grid]   -- writes gpurun_out/parity_probe_<step>.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloth_sim_self_collision_b200 import clothgpu as G  # noqa: E402
from cloth_sim_self_collision_b200 import scenes as S  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from tests import common as T  # noqa: E402

step = int(sys.argv[1]) if len(sys.argv) > 1 else 57
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 256
sc = S.square_cloth_xz(grid - 1, grid - 1, 0.1, S.SELF_CONTACT["areal_density"])
sc["bvh"] = S.reference_bvh(sc)
c = S.SELF_CONTACT
p = T.params_for(sc, c)
lib = O.load_ref() or O.load_oracle()
g = G.ClothGPU(sc)
g.step_self_contact(p, step - 1)
x0, v0 = g.download_state()
out = {"x0": x0, "v0": v0}


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


# implicit step
g.upload_state(x0, v0)
it = g.step_internal(p)[0]
gx1, gv1 = g.download_state()
sim = lib.sim(sc)
sim.set_state(x0, v0)
ocr, oit = sim.step_internal(p)
ox1, ov1 = sim.get_state()
print("implicit: iters", it, oit, "rel x", rel(gx1, ox1), "rel v", rel(gv1, ov1))
# contact phases on the oracle's velocities
x, v = x0, ov1
g.write(G.BUF_XYZ_PREV, x)
g.write(G.BUF_UVW, v)
g.bvh_refit_prx(G.BUF_XYZ_PREV, c["clearance"])
bp = lib.refit_prx(sc, x, c["clearance"])
print("boxes prx equal", np.array_equal(g.read(G.BUF_BOXES).reshape(-1, 6), bp))
n = g.contacts_prx(G.BUF_XYZ_PREV, c["clearance"])[0]
ge = g.elems(n)
ep = lib.contacts_prx(sc, bp, x, c["clearance"])
print("prx elems", n, len(ep), np.array_equal(ge, ep))
out["elems_prx_gpu"] = ge
out["elems_prx_cpu"] = ep
g.impulse_prx(G.BUF_XYZ_PREV, G.BUF_UVW, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"])
v1 = lib.impulse_prx(v, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"], x, ep)
gv = g.read(G.BUF_UVW)
print("v after prx impulse: equal", np.array_equal(gv, v1), "rel", rel(gv, v1), "ndiff", int((gv != v1).sum()))
out["v_prx_gpu"] = gv
out["v_prx_cpu"] = v1
g.write(G.BUF_UVW, v1)
riz = lib.riz_new()
g.riz_reset()
v0r = None
for r in range(8):
    g.bvh_refit_ccd(G.BUF_XYZ_PREV, G.BUF_UVW, c["dt"])
    bc = lib.refit_ccd(sc, x, v1, c["dt"])
    print("pass", r, "boxes ccd equal", np.array_equal(g.read(G.BUF_BOXES).reshape(-1, 6), bc))
    n = g.contacts_ccd(G.BUF_XYZ_PREV, G.BUF_UVW, c["dt"], c["clearance"])[0]
    ge = g.elems(n)
    ec = lib.contacts_ccd(sc, bc, x, v1, c["dt"], c["clearance"])
    print("pass", r, "ccd elems", n, len(ec), np.array_equal(ge, ec))
    out["elems_ccd%d_gpu" % r] = ge
    out["elems_ccd%d_cpu" % r] = ec
    if len(ec) == 0:
        break
    if r < 5:
        g.impulse_ccd(G.BUF_XYZ_PREV, G.BUF_UVW, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"])
        vb = v1
        v1 = lib.impulse_ccd(v1, c["clearance"], c["stiff_contact"], c["dt"], sc["mass_point"], x, ec)
        gv = g.read(G.BUF_UVW)
        print("pass", r, "v after ccd impulse: equal", np.array_equal(gv, v1), "rel", rel(gv, v1), "ndiff", int((gv != v1).sum()))
        out["v_ccd%d_in" % r] = vb
        out["v_ccd%d_gpu" % r] = gv
        out["v_ccd%d_cpu" % r] = v1
        g.write(G.BUF_UVW, v1)
        if r == 4:
            v0r = v1.copy()
            g.write(G.BUF_UVW0, v0r)
    else:
        g.riz_make()
        riz.make(ec, sc)
        g.riz_apply(G.BUF_XYZ_PREV, G.BUF_UVW)
        v1 = riz.apply(v1, x, v0r)
        gv = g.read(G.BUF_UVW)
        print("pass", r, "v after riz: equal", np.array_equal(gv, v1), "rel", rel(gv, v1))
        g.write(G.BUF_UVW, v1)
# whole resolve
g.write(G.BUF_XYZ_PREV, x)
g.write(G.BUF_UVW, ov1)
applied, stats = g.resolve_self_contact(p)
ov, oapp, ostats = sim.resolve_contacts(ov1, x, p)
gv = g.read(G.BUF_UVW)
print("resolve: stats", stats[0][:6], ostats[:6], "equal v", np.array_equal(gv, ov), "rel", rel(gv, ov))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "parity_probe_%d.npz" % step), **out)
g.close()

# whole step, teacher forced, in the execution variants
sim.set_state(x0, v0)
ost = sim.step_self_contact(p)
ox, ov = sim.get_state()
for label, env, prof in (("default", {}, False), ("profiling(no overlap)", {}, True), ("incremental off", {"CLOTHGPU_INCREMENTAL": "0"}, False),
                         ("default again", {}, False)):
    for k, val in env.items():
        os.environ[k] = val
    g2 = G.ClothGPU(sc)
    for k in env:
        os.environ.pop(k)
    for rep in range(2):
        g2.upload_state(x0, v0)
        if prof:
            g2.set_profiling(True)
        st = g2.step_self_contact(p)[0]
        gx, gv = g2.download_state()
        print(label, "rep", rep, "stats", st[:5], ost[:5], "rel x", rel(gx, ox), "rel v", rel(gv, ov), "ndiff v", int((gv != ov).sum()))
    g2.close()
