#!/usr/bin/env python
"""bench.py -- ms per implicit step with self-contact on a 256 x 256 cloth (BASELINE.json config 3).

One "step" = StepTime() of self_contact_sparse/main.cpp:213-263 without MakeNormal: assembly + ILU(0)-PCG + BVH refit /
traversal / proximity + CCD narrow phase / impulses / rigid impact zones + velocity fix-up.

Workload: 256 x 256 vertices, h = 0.1, self_contact_sparse constants, edge-on drop
onto the plane z = -0.5.  Contacts start at step ~32, so `--setup-steps` (40) untimed steps bring the scene into the
contact regime (the last W of them are the warm-up); the K timed steps continue that trajectory.  Synthetic data,
no files read.  The BVH topology is a median-split tree (an input of the path; the reference's own builder crashes
on square grids).

  python bench.py --gpus N --steps K --warmup W           our arm (CUDA, one process per GPU; N > 1 under torchrun)
  python bench.py --impl reference ...                    the reference's CPU implementation on the host cores

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
    ap.add_argument("--setup-steps", type=int, default=40, help="untimed steps before the timed region (incl. warm-up)")
    ap.add_argument("--scenes-per-gpu", type=int, default=1)
    ap.add_argument("--cpu-steps", type=int, default=3, help="steps of the bounded CPU baseline sample (0 = skip)")
    ap.add_argument("--ref-steps", type=int, default=5, help="cap on the timed steps of --impl reference")
    ap.add_argument("--kernel-steps", type=int, default=0, help="steps of the per-kernel timing pass (0 = the whole timed window)")
    ap.add_argument("--dense-steps", type=int, default=6,
                    help="extra report: steps of a second window in the dense self-contact regime (0 = skip)")
    ap.add_argument("--dense-start", type=int, default=64, help="first trajectory step of that window")
    ap.add_argument("--profile-range", default="", choices=["", "timed", "dense"],
                    help="bracket that region with cudaProfilerStart/Stop (ncu --profile-from-start off)")
    return ap.parse_args()


def make_scene(grid):
    sc = S.square_cloth_xz(grid - 1, grid - 1, 0.1, S.SELF_CONTACT["areal_density"])
    sc["bvh"] = S.median_split_bvh(sc)
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


def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref = the unmodified sources behind a C shim; the C port when
    that was not built), single threaded like the reference, from rest through the same trajectory."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc = make_scene(args.grid)
    c = consts()
    lib, kind = cpu_lib()
    p = cpu_params(sc, c)
    sim = lib.sim(sc)
    nsteps = max(1, min(args.steps, args.ref_steps))
    for _ in range(args.setup_steps):
        sim.step_self_contact(p)
    t0 = time.perf_counter()
    stats = []
    for _ in range(nsteps):
        stats.append(sim.step_self_contact(p))
    ms = (time.perf_counter() - t0) * 1000.0 / nsteps
    sample = ("%d timed steps (of --steps %d) after %d untimed setup steps from rest, %dx%d cloth, 1 thread (the "
              "reference is single threaded)" % (nsteps, args.steps, args.setup_steps, args.grid, args.grid))
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms/step", "n_gpus": args.gpus, "steps": nsteps,
        "warmup": 0, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, sc, 1),
        "cpu_baseline": {"value": ms, "unit": "ms/step", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": ms, "unit": "ms/step", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stats_last_step": {"pcg_iters": int(stats[-1][0]), "prox_elems": int(stats[-1][1]),
                            "ccd_passes": int(stats[-1][2]), "riz_rounds": int(stats[-1][3])},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sc, nscene_total):
    return {
        "workload": "self_contact_sparse %dx%d grid cloth, edge-on drop onto plane z=-0.5, trajectory steps %d..%d "
                    "(BASELINE.json configs[2])" % (args.grid, args.grid, args.setup_steps + 1,
                                                    args.setup_steps + args.steps),
        "np": sc["np"], "ntri": sc["ntri"], "nquad": sc["nquad"], "ncrs": int(sc["rowptr"][-1]),
        "scenes_total": nscene_total, "scenes_per_gpu": args.scenes_per_gpu, "setup_steps": args.setup_steps,
        "l2": "per-step working set ~0.35 GB per scene (matrix + ILU + element buffers + queues) exceeds the 126 MB "
              "L2; no explicit flush",
        "bvh": "median-split topology (input)",
        "rehearsal": "the timed trajectory is run once untimed first (sizes the on-demand contact queues), then the "
                     "state is rewound",
    }


def run_ours(args):
    from cloth_sim_self_collision_b200 import clothgpu as G
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
    from cloth_sim_self_collision_b200 import batch as BT
    mine = BT.shard_scenes(world * Sn, world, rank)  # contiguous block of global scene indices of this rank
    for s, gs in enumerate(mine):  # config-5 style gravity-y perturbation (scene 0 = config 3 itself)
        g.set_scene_gravity(s, BT.scene_gravity(c["gravity"], gs))
    K, W = args.steps, args.warmup
    setup = max(0, args.setup_steps - W)

    def barrier():
        if dist is not None:
            dist.barrier()
        g.synchronize()

    # ---- setup + warm-up (untimed)
    g.step_self_contact(p, setup)
    g.step_self_contact(p, W)
    start_state = [g.download_state(s) for s in range(Sn)]
    # capacity rehearsal (untimed): the device work queues of the contact passes grow on demand (realloc + retry of
    # the pass); running the timed trajectory once sizes them, then the state is rewound, so that the timed region
    # measures the steady state and not a one-off allocation
    g.step_self_contact(p, K)
    for s in range(Sn):
        g.upload_state(start_state[s][0], start_state[s][1], scene=s)

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
    if rank == 0:
        # ---- per-kernel timing pass (CUDA events around every launch, same trajectory from the same state)
        for s in range(Sn):
            g.upload_state(start_state[s][0], start_state[s][1], scene=s)
        g.set_kernel_timing(True)
        g.set_profiling(True)
        ks = K if args.kernel_steps <= 0 else max(1, min(args.kernel_steps, K))
        g.step_self_contact(p, ks)
        kt = g.kernel_times()
        phases = g.phase_ms()
        nleaf = int(len(g.leaf_pairs(0)))
        g.set_kernel_timing(False)
        g.set_profiling(False)
        tot = sum(v[1] for v in kt.values())
        top = sorted(kt.items(), key=lambda kv: -kv[1][1])
        peak, peak_src = load_peaks()
        ab = algorithmic_bytes(sc, nleaf)
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
        roof["per_kernel"] = {}
        for name, (n, ms) in top[:10]:
            e = {"launches_per_step": n / ks, "ms_per_step": ms / ks, "share": ms / tot}
            if name in ab:
                e["achieved_GBps"] = ab[name] * Sn / (ms / n * 1.0e-3) / 1.0e9
                e["frac_of_peak"] = e["achieved_GBps"] / peak
            roof["per_kernel"][name] = e
        roof["phase_ms_per_step"] = {k: v / ks for k, v in phases.items()}

        # ---- CPU baseline: the reference (or its C port) on this box's host, bounded sample of the same workload
        cpu = None
        if args.cpu_steps > 0:
            lib, kind = cpu_lib()
            sim = lib.sim(sc)
            sim.set_state(start_state[0][0], start_state[0][1])
            cp = cpu_params(sc, c)
            cp.gravity[1] = 0.05 + 0.001 * 0
            t0 = time.perf_counter()
            for _ in range(args.cpu_steps):
                sim.step_self_contact(cp)
            cms = (time.perf_counter() - t0) * 1000.0 / args.cpu_steps
            cpu = {"value": cms, "unit": "ms/step", "cores": 1, "kind": kind,
                   "sample": "%d steps of scene 0 from the state at the start of the timed region (trajectory steps "
                             "%d..%d), 1 thread of %d host cores (the reference is single threaded)" %
                             (args.cpu_steps, args.setup_steps + 1, args.setup_steps + args.cpu_steps,
                              os.cpu_count() or 1)}
            # parity spot check of the first CPU step against our first timed step is done in tests/, not here

        # ---- second window, deep in the self-contact regime (thousands of contact elements, 5 CCD passes, impact
        # zones): reported next to the headline so that the contact-free part of the trajectory does not flatter it
        dense = None
        if args.dense_steps > 0 and args.dense_start > args.setup_steps + K:
            for s in range(Sn):
                g.upload_state(end_state[s][0], end_state[s][1], scene=s)
            g.step_self_contact(p, args.dense_start - 1 - (args.setup_steps + K))
            d0 = [g.download_state(s) for s in range(Sn)]
            g.step_self_contact(p, args.dense_steps)  # rehearsal (queue sizes), then rewind
            for s in range(Sn):
                g.upload_state(d0[s][0], d0[s][1], scene=s)
            g.synchronize()
            g.timer_start()
            _prof(args, "dense", True)
            dst = g.step_self_contact(p, args.dense_steps)
            _prof(args, "dense", False)
            dms = g.timer_stop() / (args.dense_steps * Sn)
            dense = {"trajectory_steps": "%d..%d" % (args.dense_start, args.dense_start + args.dense_steps - 1),
                     "ms_per_step": dms, "prox_elems_last": int(dst[0][1]), "ccd_passes_last": int(dst[0][2]),
                     "riz_rounds_last": int(dst[0][3])}
            if args.cpu_steps > 0:
                lib, kind = cpu_lib()
                sim = lib.sim(sc)
                sim.set_state(d0[0][0], d0[0][1])
                cp = cpu_params(sc, c)
                cp.gravity[1] = 0.05 + 0.001 * 0
                t0 = time.perf_counter()
                sim.step_self_contact(cp)
                dense["cpu_ms_per_step"] = (time.perf_counter() - t0) * 1000.0
                dense["cpu_kind"] = kind

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
            "stats_last_step": {"pcg_iters": int(stats[0][0]), "prox_elems": int(stats[0][1]),
                                "ccd_passes": int(stats[0][2]), "riz_rounds": int(stats[0][3])},
            "checksums": checksums[:8],
            "dense_contact_window": dense,
        }
        print(json.dumps(line), flush=True)
    g.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


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
