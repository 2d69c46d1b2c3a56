#!/usr/bin/env python
"""Summarise ncu outputs into small text files kept under profiles/.

  python profiles/ncu_summarize.py launches gpurun_out/x_launches.csv  > profiles/x_launches.txt
      per-kernel launch count / total time / share, from `ncu --metrics gpu__time_duration.sum --csv`
  python profiles/ncu_summarize.py full gpurun_out/x_full.ncu-rep       > profiles/x_full.txt
      per captured launch: duration, DRAM bytes, throughput percentages, occupancy, registers, top warp stalls
      (reads the report with `ncu -i ... --page raw --csv`)
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").replace("cg::", "")


def launches(path):
    rows = [l for l in open(path, newline="") if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = OrderedDict()
    total = 0.0
    n = 0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, r["Block Size"], r["Grid Size"]])
        a[0] += 1
        a[1] += us
        total += us
        n += 1
    print("# %d launches, %.3f ms total kernel time (cold-cache, serialised: shares, not absolutes)" % (n, total / 1000))
    print("%-28s %8s %12s %10s %8s  %s" % ("kernel", "launches", "total_us", "avg_us", "share", "block / last grid"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-28s %8d %12.1f %10.2f %7.2f%%  %s %s" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / total, a[2], a[3]))


WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occ_pct"),
    ("sm__inst_executed.avg.per_cycle_active", "ipc_active"),
    ("smsp__inst_executed.sum", "inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [(h, i) for h, i in col.items()
                  if h.startswith("smsp__average_warp_latency_issue_stalled") or
                  h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    if not stall_cols:
        stall_cols = [(h, i) for h, i in col.items() if "issue_stalled" in h and h.endswith(".ratio")]
    for r in data:
        print("== %s  block %s grid %s" % (short(r[col["Kernel Name"]]), r[col["Block Size"]], r[col["Grid Size"]]))
        for m, nm in WANT:
            if m in col:
                print("   %-18s %s %s" % (nm, r[col[m]], units[col[m]]))
        st = []
        for h, i in stall_cols:
            try:
                st.append((float(r[i].replace(",", "")), h))
            except ValueError:
                pass
        st.sort(reverse=True)
        for v, h in st[:6]:
            print("   stall %-60s %.2f" % (re.sub(r"smsp__average_warps?_(latency_)?issue_stalled_", "", h), v))


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def traffic(paths):
    """{bench kernel name: DRAM bytes (read + write) per launch}, averaged over the captured launches."""
    import json
    acc = {}
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(hdr)}
        for r in data:
            name = r[col["Kernel Name"]]
            k = short(name)
            m = re.match(r"k_trisolve<\(bool\)?(\d)", name.replace("void ", "").replace("cg::", ""))
            if k.startswith("k_trisolve"):
                k = "k_trisolve_fwd" if re.search(r"k_trisolve<(\(bool\))?1", name) else "k_trisolve_bwd"
            k = re.sub(r"<.*", "", k)
            b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
                to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            acc.setdefault(k, []).append(b)
    print(json.dumps({k: sum(v) / len(v) for k, v in sorted(acc.items())}, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2:])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
