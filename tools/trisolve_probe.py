#!/usr/bin/env python
"""GPU probe: time of one ILU(0) forward+backward solve for several grids and strip counts (CLOTHGPU_STRIPS)."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cloth_sim_self_collision_b200 import clothgpu as G, scenes as S  # noqa: E402


def plan_stats(lib, sc, force):
    rp = np.ascontiguousarray(sc["rowptr"], dtype=np.int32)
    ci = np.ascontiguousarray(sc["colidx"], dtype=np.int32)
    st = np.zeros(16, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rc = lib.clothgpu_selftest_triplan(sc["np"], ip(rp), ip(ci), force, ip(st))
    return rc, st.reshape(2, 8)


def main():
    lib = G.load()
    cases = [(64, 1), (64, 2), (64, 4), (64, 8), (128, 0), (128, 7), (256, 0), (256, 13), (256, 52), (512, 0)]
    if len(sys.argv) > 1:
        cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]]
    reps = 20
    for n, force in cases:
        sc = S.square_cloth_xz(n - 1, n - 1, 0.1, S.SELF_CONTACT["areal_density"])
        rc, st = plan_stats(lib, sc, force)
        if rc != 0:
            print("grid %d strips %d: plan rejected (%s)" % (n, force, lib.clothgpu_last_error().decode()))
            continue
        if force:
            os.environ["CLOTHGPU_STRIPS"] = str(force)
        else:
            os.environ.pop("CLOTHGPU_STRIPS", None)
        g = G.ClothGPU(sc, with_bvh=False)
        rng = np.random.default_rng(0)
        # a diagonally dominant matrix with the scene's pattern: values do not matter for timing
        g.write(G.BUF_VAL_DIA, np.tile(np.eye(3).ravel() * 20.0, sc["np"]))
        g.write(G.BUF_VAL_CRS, rng.standard_normal(g.ncrs * 9) * 0.1)
        g.ilu_setvalue()
        g.ilu_factor()
        g.write(G.BUF_VEC_Z, rng.standard_normal(sc["np"] * 3))
        for _ in range(3):
            g.ilu_solve(G.BUF_VEC_Z)
        g.timer_start()
        for _ in range(reps):
            g.ilu_solve(G.BUF_VEC_Z)
        ms = g.timer_stop() / reps
        g.timer_start()
        for _ in range(5):
            g.ilu_setvalue()
            g.ilu_factor()
        fms = g.timer_stop() / 5
        print("grid %3d^2: setvalue + factor %.3f ms" % (n, fms))
        cyc = ms * 1e-3 * 1.965e9
        print("grid %3d^2 strips f/b %3d/%3d  R %5d  longest strip %4d ms  sim makespan %5d/%5d : %.3f ms per solve "
              "(fwd+bwd), %.0f cycles per simulated micro-step" %
              (n, st[0, 0], st[1, 0], st[0, 1], st[0, 3], st[0, 5], st[1, 5], ms, cyc / (st[0, 5] + st[1, 5])))
        if os.environ.get("PROBE_DEBUG_FACTOR"):
            cap = 256
            out = np.zeros(2 * cap * 8, dtype=np.uint64)
            nct = np.zeros(2, dtype=np.int32)
            lib.clothgpu_debug_trisolve(g.h, 1, None, 0, None)
            g.ilu_setvalue()
            g.ilu_factor()
            lib.clothgpu_debug_trisolve(g.h, 0, out.ctypes.data_as(C.c_void_p), cap, nct.ctypes.data_as(C.c_void_p))
            rows = out.reshape(2, cap, 8)[0, :nct[0]]
            rows = rows[np.argsort(rows[:, 7])]
            t0 = rows[:, 0].min()
            print("  factor: strip start_us end_us | per-step cycles: waits, updates, inverse, U'+L stores, barrier | steps")
            for r in rows[:: max(1, len(rows) // 8)]:
                n_ = max(int(r[6]), 1)
                tail, bar = int(r[5]) & 0xffffffff, int(r[5]) >> 32
                print("   %3d %8.1f %8.1f | %6.0f %6.0f %6.0f %6.0f %6.0f | %d" %
                      (r[7], (r[0] - t0) / 1e3, (r[1] - t0) / 1e3, r[3] / n_, r[4] / n_, r[2] / n_, tail / n_, bar / n_, n_))
        if os.environ.get("PROBE_DEBUG"):
            cap = 256
            out = np.zeros(2 * cap * 8, dtype=np.uint64)
            nct = np.zeros(2, dtype=np.int32)
            lib.clothgpu_debug_trisolve(g.h, 1, None, 0, None)
            g.ilu_solve(G.BUF_VEC_Z)
            lib.clothgpu_debug_trisolve(g.h, 0, out.ctypes.data_as(C.c_void_p), cap, nct.ctypes.data_as(C.c_void_p))
            o = out.reshape(2, cap, 8)
            for d in range(2):
                rows = o[d, :nct[d]]
                rows = rows[np.argsort(rows[:, 7])]
                t0 = rows[:, 0].min()
                print("  %s: strip start_us end_us | per-step cycles: mbar ghost comp signal | steps" % ("fwd" if d == 0 else "bwd"))
                for r in rows[:: max(1, len(rows) // 8)]:
                    n = max(int(r[6]), 1)
                    print("   %3d %8.1f %8.1f | %6.0f %6.0f %6.0f %6.0f | %d" %
                          (r[7], (r[0] - t0) / 1e3, (r[1] - t0) / 1e3, r[2] / n, r[3] / n, r[4] / n, r[5] / n, n))
        g.close()


if __name__ == "__main__":
    main()
