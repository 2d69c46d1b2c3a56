#!/usr/bin/env python
"""bench.py -- ms per implicit step with self-contact on a 256 x 256 cloth (BASELINE.json config 3).

One "step" = StepTime() of self_contact_sparse/main.cpp:213-263 without MakeNormal: assembly + ILU(0)-PCG + BVH refit /
traversal / proximity + CCD narrow phase / impulses / rigid impact zones + velocity fix-up.

Workload: 256 x 256 vertices, h = 0.1, self_contact_sparse constants, edge-on drop
onto the plane z = -0.5.  The collider is reached at step ~32, the cloth buckles and touches ITSELF from step ~54 on,
so `--setup-steps` (56) untimed steps from rest bring the scene into the self-contact regime (the last W of them are
the warm-up) and the K timed steps (57..76) continue that trajectory: thousands of contact elements, up to five CCD
impulse passes and impact-zone rounds per step.  Synthetic data, no files read.  The BVH topology is the reference
builder's own tree (MakeBVHTopology_TopDown on the seeded-jitter copy of the rest shape) from
the product's host builder (clothgpu_bvh_build_topdown).

  python bench.py --gpus N --steps K --warmup W           our arm (CUDA, one process per GPU; N > 1 under torchrun)
  python bench.py --impl reference ...                    the reference's CPU implementation on the host cores

Extra keys of the JSON line: "parity" (first timed step teacher-forced against the CPU reference; a mismatch makes the
run fail), "contact_free_window" (steps 41..50, before the first self-contact), "batch" (BASELINE.json configs[4]: 64
independent scenes sharded over the N GPUs, ms per scene-step and step-level achieved GB/s), "checksum_step"
(trajectory checksum after step setup+5, printed by both arms).

N > 1: every rank owns `--scenes-per-gpu` independent scenes (per-scene gravity perturbation as in config 5), no
collective on the data path; NCCL only gathers the final state checksums.  value = ms per scene-step of the whole job.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cloth_sim_self_collision_b200 import scenes as S  # noqa: E402

METRIC = "ms per implicit step w/ self-contact (256² cloth); % HBM roofline"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=256, help="vertices per side")
    ap.add_argument("--setup-steps", type=int, default=56,
                    help="untimed steps from rest before the timed region (incl. warm-up): 56 puts the timed window "
                         "(steps 57..76) inside the self-contact regime, which starts at step ~54")
    ap.add_argument("--scenes-per-gpu", type=int, default=1)
    ap.add_argument("--cpu-steps", type=int, default=3, help="steps of the bounded CPU baseline sample (0 = skip)")
    ap.add_argument("--ref-steps", type=int, default=20, help="cap on the timed steps of --impl reference")
    ap.add_argument("--kernel-steps", type=int, default=0, help="steps of the per-kernel timing pass (0 = the whole timed window)")
    ap.add_argument("--free-steps", type=int, default=10,
                    help="extra report: steps of a second window BEFORE the first self-contact (0 = skip)")
    ap.add_argument("--free-start", type=int, default=41, help="first trajectory step of that window")
    ap.add_argument("--batch-scenes", type=int, default=64,
                    help="extra report (BASELINE.json configs[4]): total scenes of the batch, sharded over the GPUs (0 = skip)")
    ap.add_argument("--batch-steps", type=int, default=5, help="timed steps of the batch window")
    ap.add_argument("--batch-setup", type=int, default=56, help="untimed steps of the batch from rest")
    ap.add_argument("--profile-range", default="", choices=["", "timed", "free", "batch"],
                    help="bracket that region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    return ap.parse_args()


def make_scene(grid):
    sc = S.square_cloth_xz(grid - 1, grid - 1, 0.1, S.SELF_CONTACT["areal_density"])
    sc["bvh"] = S.reference_bvh(sc)
    return sc


def consts():
    return dict(S.SELF_CONTACT)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.path = "/tmp/clothgpu_clocks_%d_%d.csv" % (os.getpid(), index)
        self.proc = None

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if parts[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def algorithmic_bytes(sc, nleaf):
    """Compulsory bytes per launch of the hot kernels (SURVEY.md section 8d / DESIGN.md), one scene."""
    P, T, Q = sc["np"], sc["ntri"], sc["nquad"]
    Cc = int(sc["rowptr"][-1])
    N = 2 * T - 1
    return {
        # level-scheduled single-CTA kernels (CLOTHGPU_TRISOLVE=levels): ILU values once + column indices + vector
        "k_ilu_solve": 72 * (P + Cc) + 4 * Cc + 8 * P + 48 * P + 24 * P,
        "k_ilu_factor": 2 * 72 * (P + Cc),
        # strip-pipelined sweeps: forward reads the lower blocks + D^-1, backward the upper blocks; column indices
        # (4 B per block) and the vector read + written once each (24 B per row each way)
        "k_trisolve_fwd": 72 * (Cc // 2 + P) + 4 * (Cc // 2) + 48 * P,
        "k_trisolve_bwd": 72 * (Cc // 2) + 4 * (Cc // 2) + 48 * P,
        # every ILU block read and written once, every finished upper block read once more by its consumers
        "k_ilu_factor_strips": 2 * 72 * (P + Cc) + 72 * (Cc // 2),
        # gather of the factor into lane order: every block read and written once
        "k_pack": 2 * 72 * (P + Cc),
        "k_pack_rows": 2 * 72 * (P + Cc),
        "k_matvec": 72 * (P + Cc) + 4 * Cc + 4 * P + 48 * P,
        # (k_cull / k_narrow / k_trav_step: no bandwidth model -- their re-reads of vertices and boxes are served by
        # L1/L2, ncu measures ~48 MB of DRAM traffic per full pass; they are latency / divergence bound)
        "k_tri_elements": T * (12 + 90 * 8) + 48 * P,
        "k_gather_slot": Cc * 72 + T * 6 * 72,
        "k_gather_vertex": P * (72 + 24 + 48) + T * 3 * (72 + 24),
        "k_refit": 12 * T + 48 * P + 2 * 48 * N,
        # exact dot product: both vectors are read by the chunk-sum pass and again by the statistics pass
        "k_xdot_partial": 48 * P,
        "k_xdot_stats": 48 * P,
    }


# what bounds the kernels that have no meaningful bandwidth figure
LIMITERS = {
    "k_imp_flow": "latency: the Gauss-Seidel impulse sweep is one dependency chain (depth ~0.9 E; ~1.1 us of dependent "
                  "fp64 arithmetic per hop: coplanarity cubic + 14 bisection halvings + 7 divisions, 85-95 % of the "
                  "elements write); algorithmic bytes E*212",
    "k_ilu_factor_strips": "latency: 1021 dependency levels of the reference's row ordering, strip-pipelined wavefront",
    "k_trisolve_fwd": "latency: 1021 dependency levels, strip-pipelined wavefront (one CTA per strip)",
    "k_trisolve_bwd": "latency: 1021 dependency levels, strip-pipelined wavefront (one CTA per strip)",
    "k_xdot_walk": "latency: one warp per scene replays the ~15 binade crossings of the serial sum",
    "k_trav_step": "launch / L2 latency: one launch per BVH level, small frontiers",
    "k_narrow": "fp64 pipe + load latency: full predicates (cubic solves) on the culled candidates",
    "k_cull": "L2 bandwidth / load latency: 15 cheap predicate prefixes per leaf pair, operands re-read from L2",
}


def traffic_table():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            pass
    return {}


def cpu_lib():
    from oracle import oracle_py as O
    ref = O.load_ref()
    if ref is not None:
        return ref, "reference"
    return O.load_oracle(), "port"


def cpu_params(sc, c):
    from oracle import oracle_py as O
    return O.Params(c["dt"], c["lam"], c["myu"], c["stiff_bend"], c["gravity"], sc["mass_point"], c["stiff_contact"],
                    c["clearance"], c["col_kind"], c["col_par"])


def step_bytes(sc, pcg_iters, ccd_passes, nleaf):
    """Algorithmic bytes of one StepTime(): assembly + ILU set/factor + PCG prologue and
    iterations + BVH refit per detection pass + the leaf-pair operands of the narrow phase."""
    P, T, Q = sc["np"], sc["ntri"], sc["nquad"]
    Cc = int(sc["rowptr"][-1])
    N = 2 * T - 1
    assembly = 96 * P + 72 * P + 12 * T + 16 * Q + 72 * Cc
    ilu = 2 * 72 * (P + Cc)
    spmv = 72 * (P + Cc) + 4 * Cc + 4 * P + 48 * P
    precond = 72 * (P + Cc) + 4 * Cc + 8 * P + 48 * P
    blas1 = 17 * 24 * P
    pcg = (precond + 5 * 24 * P) + pcg_iters * (spmv + precond + blas1)
    refit = 12 * T + 48 * P + 2 * 48 * N
    passes = 1 + ccd_passes
    return assembly + ilu + pcg + passes * refit + passes * nleaf * 312


def _ref_batch_worker(args_tuple):
    """One host process of the reference arm's batch sample: scene `s` (config-5 gravity perturbation) stepped from
    the given state; returns the wall time of its steps."""
    grid, s, x, v, nsteps = args_tuple
    from cloth_sim_self_collision_b200 import batch as BT
    sc = make_scene(grid)
    c = consts()
    lib, _ = cpu_lib()
    cp = cpu_params(sc, c)
    g = BT.scene_gravity(c["gravity"], s)
    for k in range(3):
        cp.gravity[k] = g[k]
    sim = lib.sim(sc)
    sim.set_state(x, v)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        sim.step_self_contact(cp)
    return time.perf_counter() - t0


def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref = the unmodified sources behind a C shim; the C port when
    that was not built), single threaded like the reference, from rest through the same trajectory and the same
    timed steps as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cloth_sim_self_collision_b200 import batch as BT
    sc = make_scene(args.grid)
    c = consts()
    lib, kind = cpu_lib()
    if kind == "reference":  # the reference's own builder on the same input (the same tree, tests/test_bvh_builder.py)
        sc["bvh"] = lib.bvh_build(sc["tri"], S.jittered_rest(sc))
    p = cpu_params(sc, c)
    sim = lib.sim(sc)
    nsteps = max(1, min(args.steps, args.ref_steps))
    for _ in range(args.setup_steps):
        sim.step_self_contact(p)
    x56, v56 = sim.get_state()
    t0 = time.perf_counter()
    stats = []
    chk = None
    for k in range(nsteps):
        stats.append(sim.step_self_contact(p))
        if k + 1 == min(5, nsteps):
            chk = BT.state_checksum(sim.get_state()[0])
    ms = (time.perf_counter() - t0) * 1000.0 / nsteps
    sample = ("%d timed steps (of --steps %d; trajectory steps %d..%d) after %d untimed setup steps from rest, %dx%d "
              "cloth, 1 thread (the reference is single threaded)" %
              (nsteps, args.steps, args.setup_steps + 1, args.setup_steps + nsteps, args.setup_steps, args.grid, args.grid))
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms/step", "n_gpus": args.gpus, "steps": nsteps,
        "warmup": 0, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, sc, 1),
        "cpu_baseline": {"value": ms, "unit": "ms/step", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": ms, "unit": "ms/step", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stats_last_step": {"pcg_iters": int(stats[-1][0]), "prox_elems": int(stats[-1][1]),
                            "ccd_passes": int(stats[-1][2]), "riz_rounds": int(stats[-1][3])},
        "checksum_step": {"step": args.setup_steps + min(5, nsteps), "value": chk},
    }
    # ---- batch (BASELINE.json configs[4]) on the host: all cores, one scene per process, a bounded sample
    if args.batch_scenes > 0:
        import multiprocessing as mp
        cores = max(1, min(os.cpu_count() or 1, args.batch_scenes))
        bs = max(1, min(args.batch_steps, 2))
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_ref_batch_worker, [(args.grid, s, x56, v56, bs) for s in range(cores)])
        wall = time.perf_counter() - t0
        line["batch"] = {
            "scenes_total": args.batch_scenes, "ms_per_scene_step": wall * 1000.0 / (cores * bs), "cores": cores,
            "kind": kind,
            "sample": "%d concurrent host processes (one per core), each one scene of the batch (its gravity "
                      "perturbation) for %d steps from scene 0's state after step %d; the batch of %d scenes is %.1f "
                      "such waves, process start-up included" % (cores, bs, args.setup_steps, args.batch_scenes,
                                                                 args.batch_scenes / cores)}
    print(json.dumps(line), flush=True)


def workload_config(args, sc, nscene_total):
    return {
        "workload": "self_contact_sparse %dx%d grid cloth, edge-on drop onto plane z=-0.5, trajectory steps %d..%d: "
                    "inside the self-contact regime (BASELINE.json configs[2])" %
                    (args.grid, args.grid, args.setup_steps + 1, args.setup_steps + args.steps),
        "np": sc["np"], "ntri": sc["ntri"], "nquad": sc["nquad"], "ncrs": int(sc["rowptr"][-1]),
        "scenes_total": nscene_total, "scenes_per_gpu": args.scenes_per_gpu, "setup_steps": args.setup_steps,
        "l2": "per-step working set ~0.35 GB per scene (matrix + ILU + element buffers + queues) exceeds the 126 MB "
              "L2; no explicit flush",
        "bvh": "MakeBVHTopology_TopDown tree of the seeded-jitter rest copy (clothgpu_bvh_build_topdown, == reference)",
        "rehearsal": "the timed trajectory is run once untimed first (sizes the on-demand contact queues), then the "
                     "state is rewound",
    }


def _parity(g, G, sc, c, p, x0, v0, gravity):
    """Teacher-forced check of ONE step of scene 0 from (x0, v0): our step through the C ABI against the CPU reference
    (oracle/_ref, or the C port) on the same state.  Returns the report and the CPU time of that step."""
    lib, kind = cpu_lib()
    g.upload_state(x0, v0, scene=0)
    st = g.step_self_contact(p, 1)[0]
    gx, gv = g.download_state(0)
    sim = lib.sim(sc)
    sim.set_state(x0, v0)
    cp = cpu_params(sc, c)
    for k in range(3):
        cp.gravity[k] = gravity[k]
    t0 = time.perf_counter()
    ost = sim.step_self_contact(cp)
    cms = (time.perf_counter() - t0) * 1000.0
    ox, ov = sim.get_state()
    rx = float(np.abs(gx - ox).max() / max(np.abs(ox).max(), 1e-300))
    rv = float(np.abs(gv - ov).max() / max(np.abs(ov).max(), 1e-300))
    eq = bool(np.array_equal(st[:5], ost[:5]) and np.array_equal(st[8:116], ost[8:116]))
    bit = bool(np.array_equal(gx, ox) and np.array_equal(gv, ov))
    rep = {"checker": kind, "stats_equal": eq, "bit_identical": bit, "rel_x": rx, "rel_v": rv, "tol": 0.0,
           "stats": [int(k) for k in st[:5]], "ok": bool(eq and bit)}
    return rep, cms, sim, cp


def run_ours(args):
    from cloth_sim_self_collision_b200 import clothgpu as G
    from cloth_sim_self_collision_b200 import batch as BT
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if G.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device visible; the product path has no CPU fallback")

    sc = make_scene(args.grid)
    c = consts()
    p = G.make_params(c["dt"], c["lam"], c["myu"], c["stiff_bend"], c["gravity"], sc["mass_point"], c["stiff_contact"],
                      c["clearance"], c["col_kind"], c["col_par"])
    Sn = args.scenes_per_gpu
    g = G.ClothGPU(sc, nscene=Sn, device=local)
    mine = BT.shard_scenes(world * Sn, world, rank)  # contiguous block of global scene indices of this rank
    for s, gs in enumerate(mine):  # config-5 style gravity-y perturbation (scene 0 = config 3 itself)
        g.set_scene_gravity(s, BT.scene_gravity(c["gravity"], gs))
    grav0 = BT.scene_gravity(c["gravity"], mine[0])
    K, W = args.steps, args.warmup
    setup = max(0, args.setup_steps - W)

    def barrier():
        if dist is not None:
            dist.barrier()
        g.synchronize()

    def rewind(states, ctx=None):
        ctx = ctx or g
        for s in range(len(states)):
            ctx.upload_state(states[s][0], states[s][1], scene=s)

    # ---- setup + warm-up (untimed)
    rest = [g.download_state(s) for s in range(Sn)]
    g.step_self_contact(p, setup)
    g.step_self_contact(p, W)
    start_state = [g.download_state(s) for s in range(Sn)]
    # capacity rehearsal (untimed): the device work queues of the contact passes grow on demand (realloc + retry of
    # the pass); running the timed trajectory once sizes them, then the state is rewound, so that the timed region
    # measures the steady state and not a one-off allocation.  The state after 5 steps gives the trajectory checksum
    # that the reference arm prints for the same step.
    k5 = min(5, K)
    g.step_self_contact(p, k5)
    chk5 = BT.state_checksum(g.download_state(0)[0])
    g.step_self_contact(p, K - k5)
    rewind(start_state)

    # ---- timed region, state resident in HBM
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = g.launch_count()
    g.timer_start()
    _prof(args, "timed", True)
    stats = g.step_self_contact(p, K)
    _prof(args, "timed", False)
    ms_dev = g.timer_stop()
    launches = g.launch_count() - l0
    barrier()
    clocks = sampler.stop()
    end_state = [g.download_state(s) for s in range(Sn)]

    # ---- end to end through the public API with HOST buffers: every step uploads X,V and downloads X,V
    # (the caller's X, V live in page-locked host memory, clothgpu_alloc_host: one DMA per copy)
    host = []
    for x, v in start_state:
        hx, hv = g.alloc_host(x.size), g.alloc_host(v.size)
        hx[:] = x
        hv[:] = v
        host.append((hx, hv))
    barrier()
    t0 = time.perf_counter()
    g.timer_start()
    for _ in range(K):
        for s in range(Sn):
            g.upload_state(host[s][0], host[s][1], scene=s)
        g.step_self_contact(p, 1)
        for s in range(Sn):
            g.download_state(s, out=host[s])
    ms_e2e = g.timer_stop()
    wall_e2e = (time.perf_counter() - t0) * 1000.0
    barrier()
    ms_e2e = max(ms_e2e, wall_e2e)  # host-side time (copies, launches) counts in the end-to-end figure
    for s in range(Sn):  # the round trip must not change the trajectory
        assert np.array_equal(host[s][0], end_state[s][0]), "e2e trajectory differs from the device-resident one"

    # ---- batch of independent scenes (BASELINE.json configs[4] / SURVEY.md 8d config 5): --batch-scenes scenes in
    # total, a contiguous block per GPU, per-scene gravity perturbation, no collective on the data path
    batch = None
    if args.batch_scenes > 0:
        bm = BT.shard_scenes(args.batch_scenes, world, rank)
        Sb = len(bm)
        gb = G.ClothGPU(sc, nscene=max(Sb, 1), device=local)
        for s, gs in enumerate(bm):
            gb.set_scene_gravity(s, BT.scene_gravity(c["gravity"], gs))
        gb.step_self_contact(p, args.batch_setup)
        b0 = [gb.download_state(s) for s in range(max(Sb, 1))]
        gb.step_self_contact(p, args.batch_steps)  # rehearsal, then rewind
        rewind(b0, gb)
        gb.synchronize()
        barrier()
        gb.timer_start()
        _prof(args, "batch", True)
        bst = gb.step_self_contact(p, args.batch_steps)
        _prof(args, "batch", False)
        bms = gb.timer_stop()
        barrier()
        nleaf_b = int(len(gb.leaf_pairs(0)))
        bsum = [BT.state_checksum(gb.download_state(s)[0]) for s in range(Sb)]
        bkern = None
        if rank == 0:  # where the batch step spends its time: per-kernel device times of one more step
            gb.set_kernel_timing(True)
            gb.step_self_contact(p, 1)
            bk = gb.kernel_times()
            gb.set_kernel_timing(False)
            btot = sum(v[1] for v in bk.values())
            bkern = {"step_ms_sum_of_kernels": btot,
                     "top": {k: {"launches": n, "ms": ms} for k, (n, ms) in sorted(bk.items(), key=lambda kv: -kv[1][1])[:14]}}
        gb.close()
        if dist is not None:
            bms = BT.max_over_ranks(dist, [bms], device="cuda")[0]
        peak, peak_src = load_peaks()
        its = float(np.mean([st[0] for st in bst[:Sb]])) if Sb else 0.0
        passes = float(np.mean([st[2] + st[3] for st in bst[:Sb]])) if Sb else 0.0
        per_scene_bytes = step_bytes(sc, its, passes, nleaf_b)
        ms_scene = bms / (args.batch_steps * args.batch_scenes)
        gbps = per_scene_bytes * len(bm) / (bms / args.batch_steps * 1.0e-3) / 1.0e9 if bms > 0 else None
        batch = {"scenes_total": args.batch_scenes, "scenes_per_gpu": Sb, "steps": args.batch_steps,
                 "trajectory_steps": "%d..%d" % (args.batch_setup + 1, args.batch_setup + args.batch_steps),
                 "ms_per_batch_step": bms / args.batch_steps, "ms_per_scene_step": ms_scene,
                 "algorithmic_bytes_per_scene_step": per_scene_bytes, "achieved_GBps_per_gpu": gbps,
                 "frac_of_hbm_peak": gbps / peak if gbps else None, "peak": peak,
                 "stats_scene0_last": [int(k) for k in bst[0][:5]],
                 "pcg_iters_mean": its, "detection_passes_mean": passes, "checksums": bsum[:4],
                 "kernels_one_step": bkern}

    # ---- max over ranks
    if dist is not None:
        import torch
        ms_dev, ms_e2e = BT.max_over_ranks(dist, [ms_dev, ms_e2e], device="cuda")
        tl = torch.tensor([float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
        launches_total = int(tl[0])
        # NCCL gathers the results (one checksum per scene; the north star's only collective)
        checksums = [float(v) for v in BT.gather_scene_values(dist, [BT.state_checksum(x) for x, _ in end_state],
                                                              world * Sn, device="cuda")]
    else:
        launches_total = launches
        checksums = [BT.state_checksum(x) for x, _ in end_state]

    nscene_total = world * Sn
    line = None
    ok = True
    if rank == 0:
        # ---- per-kernel timing pass (CUDA events around every launch, same trajectory from the same state)
        rewind(start_state)
        g.set_kernel_timing(True)
        g.set_profiling(True)
        ks = K if args.kernel_steps <= 0 else max(1, min(args.kernel_steps, K))
        imp_elems = 0
        det_passes = 0
        for _ in range(ks):
            st1 = g.step_self_contact(p, 1)[0]
            imp_elems += int(st1[1]) + int(sum(st1[8:8 + 5]))   # elements run through the Gauss-Seidel impulse sweeps
            det_passes += 1 + int(st1[2]) + int(st1[3])
        kt = g.kernel_times()
        phases = g.phase_ms()
        nleaf = int(len(g.leaf_pairs(0)))
        g.set_kernel_timing(False)
        g.set_profiling(False)
        tot = sum(v[1] for v in kt.values())
        top = sorted(kt.items(), key=lambda kv: -kv[1][1])
        peak, peak_src = load_peaks()
        ab = algorithmic_bytes(sc, nleaf)
        if "k_imp_flow" in kt and kt["k_imp_flow"][0] > 0:  # E * (20 + 4 * 48) bytes per sweep
            ab["k_imp_flow"] = 212.0 * imp_elems / kt["k_imp_flow"][0]
        traffic = traffic_table()
        dom_name, (dom_n, dom_ms) = top[0]
        roof = {"bound": "hbm", "kernel": dom_name, "achieved": None, "peak": peak, "unit": "GB/s", "frac": None,
                "traffic": traffic.get(dom_name), "peak_source": peak_src,
                "kernel_share_of_step": dom_ms / tot if tot > 0 else None,
                "avg_launch_ms": dom_ms / dom_n, "launches_per_step": dom_n / ks}
        if dom_name in ab:
            per_launch = ab[dom_name] * Sn
            roof["algorithmic_bytes_per_launch"] = per_launch
            roof["achieved"] = per_launch / (dom_ms / dom_n * 1.0e-3) / 1.0e9
            roof["frac"] = roof["achieved"] / peak
        if dom_name in LIMITERS:
            roof["limiter"] = LIMITERS[dom_name]
        roof["per_kernel"] = {}
        for name, (n, ms) in top[:12]:
            e = {"launches_per_step": n / ks, "ms_per_step": ms / ks, "share": ms / tot}
            if name in ab:
                e["achieved_GBps"] = ab[name] * Sn / (ms / n * 1.0e-3) / 1.0e9
                e["frac_of_peak"] = e["achieved_GBps"] / peak
            if name in LIMITERS:
                e["limiter"] = LIMITERS[name].split(":")[0]
            roof["per_kernel"][name] = e
        roof["phase_ms_per_step"] = {k: v / ks for k, v in phases.items()}
        roof["impulse_elements_per_step"] = imp_elems / ks
        roof["detection_passes_per_step"] = det_passes / ks
        if "k_imp_flow" in kt and imp_elems > 0:
            roof["impulse_ns_per_element"] = kt["k_imp_flow"][1] * 1.0e6 / imp_elems

        # ---- parity of the benchmarked trajectory + CPU baseline: the reference (or its C port) on this box's host,
        # a bounded sample of the same workload from the state at the start of the timed region
        cpu = None
        parity = None
        if args.cpu_steps > 0:
            parity, cms1, sim, cp = _parity(g, G, sc, c, p, start_state[0][0], start_state[0][1], grav0)
            ok = ok and parity["ok"]
            t0 = time.perf_counter()
            for _ in range(args.cpu_steps - 1):
                sim.step_self_contact(cp)
            cms = (cms1 + (time.perf_counter() - t0) * 1000.0) / args.cpu_steps
            cpu = {"value": cms, "unit": "ms/step", "cores": 1, "kind": parity["checker"],
                   "sample": "%d steps of scene 0 from the state at the start of the timed region (trajectory steps "
                             "%d..%d), 1 thread of %d host cores (the reference is single threaded)" %
                             (args.cpu_steps, args.setup_steps + 1, args.setup_steps + args.cpu_steps,
                              os.cpu_count() or 1)}
            parity["step"] = args.setup_steps + 1

        # ---- second window BEFORE the first self-contact (implicit solve + detection passes that find nothing):
        # the figure round 1 reported as its headline, kept as an extra
        free = None
        if args.free_steps > 0 and args.free_start >= 1:
            rewind(rest)
            g.step_self_contact(p, args.free_start - 1)
            f0 = [g.download_state(s) for s in range(Sn)]
            g.step_self_contact(p, args.free_steps)
            rewind(f0)
            g.synchronize()
            g.timer_start()
            _prof(args, "free", True)
            fst = g.step_self_contact(p, args.free_steps)
            _prof(args, "free", False)
            fms = g.timer_stop() / (args.free_steps * Sn)
            free = {"trajectory_steps": "%d..%d" % (args.free_start, args.free_start + args.free_steps - 1),
                    "ms_per_step": fms, "prox_elems_last": int(fst[0][1]), "ccd_passes_last": int(fst[0][2])}
            if args.cpu_steps > 0:
                fpar, fcms, _, _ = _parity(g, G, sc, c, p, f0[0][0], f0[0][1], grav0)
                ok = ok and fpar["ok"]
                fpar["step"] = args.free_start
                free["cpu_ms_per_step"] = fcms
                free["parity"] = fpar

        value = ms_dev / (K * nscene_total)
        line = {
            "metric": METRIC, "value": value, "unit": "ms/step", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args, sc, nscene_total),
            "clocks": clocks,
            "e2e": {"value": ms_e2e / (K * nscene_total), "unit": "ms/step",
                    "h2d_bytes_per_step": 2 * 8 * sc["np"] * 3 * nscene_total,
                    "d2h_bytes_per_step": 2 * 8 * sc["np"] * 3 * nscene_total},  # X and V of every scene of the job
            "gpu_launches": launches_total,
            "roofline": roof,
            "cpu_baseline": cpu,
            "parity": parity,
            "stats_last_step": {"pcg_iters": int(stats[0][0]), "prox_elems": int(stats[0][1]),
                                "ccd_passes": int(stats[0][2]), "riz_rounds": int(stats[0][3])},
            "checksums": checksums[:8],
            "checksum_step": {"step": args.setup_steps + k5, "value": chk5},
            "contact_free_window": free,
            "batch": batch,
        }
        print(json.dumps(line), flush=True)
    g.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        raise SystemExit("bench.py: PARITY MISMATCH against the CPU reference on the benchmarked trajectory (see 'parity')")


def _prof(args, which, start):
    if args.profile_range != which:
        return
    import torch
    rt = torch.cuda.cudart()
    torch.cuda.synchronize()
    (rt.cudaProfilerStart if start else rt.cudaProfilerStop)()


def main():
    args = parse_args()
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: `python bench.py --gpus N` re-launches itself one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
