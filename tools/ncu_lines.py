"""Per-CUDA-source-line summary of an ncu report's source page (warp-state samples, instructions, top stall reasons).
usage: python tools/ncu_lines.py <report.ncu-rep> [top N] [kernel regex]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 3:
    cmd += ["-k", "regex:" + sys.argv[3]]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
cur, idx, data, tot, stall_tot = None, None, {}, 0, {}
for r in csv.reader(raw.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Line No":
        idx = {}
        for i, n in enumerate(r):
            idx.setdefault(n, i)
        stalls = [n for n in r if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if idx is None or len(r) < len(idx) or not r[0]:
        continue
    try:
        ns, inst = int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]])
    except ValueError:
        continue
    if not ns and not inst:
        continue
    st = {k: int(r[idx[k]] or 0) for k in stalls}
    key = (cur, int(r[0]))
    if key in data:
        o = data[key]
        data[key] = (o[0] + ns, o[1] + inst, o[2], {k: o[3][k] + st[k] for k in st})
    else:
        data[key] = (ns, inst, r[1].strip(), st)
    tot += ns
    for k, v in st.items():
        stall_tot[k] = stall_tot.get(k, 0) + v
print(f"total samples {tot}, warp instructions {sum(v[1] for v in data.values())}")
for k, v in sorted(stall_tot.items(), key=lambda x: -x[1])[:9]:
    print(f"  {k:26s} {100 * v / max(tot, 1):6.2f} %")
byf = {}
for (f, l), v in data.items():
    a = byf.setdefault(f, [0, 0])
    a[0] += v[0]
    a[1] += v[1]
for f, (ns, inst) in sorted(byf.items(), key=lambda x: -x[1][0]):
    print(f"{f:36s} {100 * ns / max(tot, 1):6.2f} %  inst {inst}")
for (f, l), (ns, inst, src, st) in sorted(data.items(), key=lambda x: -x[1][0])[:topn]:
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"{100 * ns / max(tot, 1):5.2f}% {f[:16]:16s}:{l:5d} inst={inst:10d} {top[0][0][6:]}={top[0][1] * 100 // max(ns, 1)}% "
          f"{top[1][0][6:]}={top[1][1] * 100 // max(ns, 1)}% | {src[:80]}")
