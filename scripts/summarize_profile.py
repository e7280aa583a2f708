"""Turns gpurun_out/{launches.csv, prof_main.ncu-rep, bench.json} into the committed
summaries under profiles/ (text + csv; the .ncu-rep itself stays in gpurun_out/)."""
import collections
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
G = ROOT / "gpurun_out"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip() or "?"
out = ROOT / "profiles"
out.mkdir(exist_ok=True)

# 1. launch list of one bench step (ncu --metrics gpu__time_duration.sum --clock-control none)
lines = [l for l in open(G / "launches.csv") if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
durs = [float(r["Metric Value"].replace(",", "")) for r in rows]
idx = [i for i, n in enumerate(names) if n.startswith("k_measure")]
start = (idx[1] if len(idx) > 1 else idx[0]) - 3   # second step (after the warm-up); the three verdict kernels precede the write batch
end = (idx[2] - 3) if len(idx) > 2 else len(names)
agg = collections.OrderedDict()
for n, d in zip(names[start:end], durs[start:end]):
    k = re.sub(r"\(.*", "", n)
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += d
tot = sum(v[1] for v in agg.values())
with open(out / f"{tag}_launches_one_step.csv", "w") as f:
    f.write("kernel,launches,total_us,share_pct\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{v[0]},{v[1] / 1e3:.1f},{100 * v[1] / tot:.1f}\n")
    f.write(f"\"TOTAL\",{end - start},{tot / 1e3:.1f},100.0\n")
(out / f"{tag}_launches_raw.csv").write_text("".join(lines))

# 2. key metrics of the full captures
raw = subprocess.run(["ncu", "-i", str(G / "prof_main.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr = r[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
ki = hdr.index("Kernel Name")
seen = set()
with open(out / f"{tag}_ncu_full_summary.txt", "w") as f:
    f.write("ncu --set full --clock-control none --import-source on, bench.py --steps 1 --warmup 1 (one GPU)\n")
    f.write("units: " + ", ".join(f"{w}={r[1][hdr.index(w)]}" for w in want if w in hdr) + "\n\n")
    for row in r[2:]:
        k = re.sub(r"\(.*", "", row[ki])
        if k in seen:
            continue
        seen.add(k)
        f.write(k + "\n")
        for w in want:
            if w in hdr:
                f.write(f"    {w:72s} {row[hdr.index(w)]}\n")
        d = [(float(row[i].replace(",", "")), hdr[i]) for i in range(len(hdr))
             if "smsp__pcsamp_warps_issue_stalled" in hdr[i] and "not_issued" not in hdr[i] and row[i] not in ("", "n/a")]
        t = sum(x for x, _ in d) or 1
        f.write("    stall samples: " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100 * x / t:.0f}%"
                                                   for x, h in sorted(d, reverse=True)[:8]) + "\n\n")
        if k == "k_fanout":
            traffic = float(row[hdr.index("dram__bytes_read.sum")].replace(",", "")) * \
                {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[r[1][hdr.index("dram__bytes_read.sum")]] + \
                float(row[hdr.index("dram__bytes_write.sum")].replace(",", "")) * \
                {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[r[1][hdr.index("dram__bytes_write.sum")]]
            (out / "fanout_traffic.json").write_text(json.dumps(
                {"dram_bytes_per_launch": traffic, "source": f"profiles/{tag}_ncu_full_summary.txt (dram__bytes_read.sum + dram__bytes_write.sum of one k_fanout "
                                                             f"launch, ncu --set full on the build of commit {head}; a committed figure, not measured by the run that prints it)"}))
if (G / "bench.json").exists():
    (out / f"{tag}_bench.json").write_text((G / "bench.json").read_text())
print("wrote", sorted(p.name for p in out.iterdir()))
