"""Turns the scratch ncu captures of tools/ncu_capture.sh (gpurun_out/) into the committed summaries under profiles/.
usage: python tools/ncu_summarize.py <tag>      e.g. r01b"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01b"

# ---- launch list ----
rows = []
with open(os.path.join(OUT, "launches.csv")) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.reader(io.StringIO("".join(lines)))
hdr = next(rd)
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
iu = hdr.index("Metric Unit")
agg = {}
for r in rd:
    if len(r) <= iv or r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    unit = r[iu]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    name = r[ik].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
with open(os.path.join(PROF, f"{tag}_launch_list_summary.txt"), "w") as f:
    f.write(f"# {tag} - ncu launch list (gpu__time_duration.sum, --clock-control none) of\n")
    f.write("#   python bench.py --frames 300 --steps 1 --warmup 1 --cpu-sample 4 --kitti-frames 0 --c5-frames 0\n")
    f.write("# Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  Raw list: "
            f"{tag}_launches.csv\n")
    f.write(f"# total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches\n")
    f.write(f"{'kernel':40s} {'launches':>8s} {'total_us':>12s} {'share_%':>8s} {'avg_us':>10s}\n")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k[:40]:40s} {n:8d} {t:12.1f} {100 * t / tot:8.2f} {t / n:10.1f}\n")
subprocess.check_call(["cp", os.path.join(OUT, "launches.csv"), os.path.join(PROF, f"{tag}_launches.csv")])

# ---- --set full: selected metrics per kernel ----
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "launch__waves_per_multiprocessor", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
           "smsp__inst_executed.sum"]
with open(os.path.join(PROF, f"{tag}_ncu_set_full_selected.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["report", "Kernel Name"] + METRICS)
    for rep in ("prof_line", "prof_orb"):
        p = os.path.join(OUT, rep + ".ncu-rep")
        if not os.path.exists(p):
            continue
        raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(io.StringIO(raw)))
        h, units = rr[0], rr[1]
        idx = {n: i for i, n in enumerate(h)}
        w.writerow([rep + " (units)", ""] + [units[idx[m]] if m in idx else "" for m in METRICS])
        seen = set()
        for r in rr[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].replace("pl::", "")
            if name in seen:
                continue
            seen.add(name)
            w.writerow([rep, name] + [r[idx[m]] if m in idx else "" for m in METRICS])

# ---- stall reasons of k_lsd_grow (source page) ----
p = os.path.join(OUT, "prof_line.ncu-rep")
if os.path.exists(p):
    raw = subprocess.run(["ncu", "-i", p, "--page", "source", "--csv", "-k", "regex:k_lsd_grow"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rr) if "Address" in r and "# Samples" in r)
    h = rr[hi]
    idx = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = {s: 0 for s in stalls}
    ns = inst = 0
    for r in rr[hi + 1:]:
        if len(r) < len(h):
            continue
        try:
            ns += int(r[idx["# Samples"]])
        except ValueError:
            continue
        inst += int(r[idx["Instructions Executed"]])
        for s in stalls:
            tot[s] += int(r[idx[s]] or 0)
    with open(os.path.join(PROF, f"{tag}_lsd_grow_stalls.txt"), "w") as f:
        f.write(f"# {tag} - k_lsd_grow, 300 frames, ncu --set full source page: warp-state samples by stall reason\n")
        f.write(f"# samples {ns}, warp instructions executed {inst}\n")
        for s, v in sorted(tot.items(), key=lambda x: -x[1]):
            if v:
                f.write(f"{s:28s} {v:9d} {100.0 * v / max(ns, 1):6.2f} %\n")
print("written to", PROF)
