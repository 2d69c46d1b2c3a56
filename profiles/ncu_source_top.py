"""Top SASS instructions by warp-stall samples from an ncu report: python profiles/ncu_source_top.py rep.ncu-rep [N]
(reads `ncu -i rep --page source --csv`; the first kernel of the report)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
heads = [i for i, r in enumerate(rows) if "Source" in r and len(r) > 5]
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
h0 = heads[kidx]
end = heads[kidx + 1] if len(heads) > kidx + 1 else len(rows)
hdr = rows[h0]
ci = {h: i for i, h in enumerate(hdr)}
s_all = ci["Warp Stall Sampling (All Samples)"]
items = []
for k, r in enumerate(rows[h0 + 1:end]):
    if len(r) <= s_all:
        continue
    try:
        v = float(r[s_all])
    except ValueError:
        continue
    stalls = {h: float(r[i]) for h, i in ci.items() if h.startswith("stall_") and r[i] not in ("", "0")}
    items.append((v, k, r[ci["Source"]], r[ci["Instructions Executed"]], stalls))
tot = sum(i[0] for i in items) or 1.0
print("total samples", tot, "instructions", len(items))
for v, k, src, nex, stalls in sorted(items, reverse=True)[:top]:
    big = sorted(stalls.items(), key=lambda kv: -kv[1])[:2]
    print("%5.1f%% #%-5d x%-6s %-70s %s" % (100 * v / tot, k, nex, src[:70], big))
