"""Summarise ncu outputs brought back in gpurun_out/ into tracked markdown under profiles/.
usage: python profiles/summarize.py <round tag, e.g. r01> [launches_per_rep]"""
import collections
import csv
import re
import subprocess
import sys

tag = sys.argv[1]
per_rep = int(sys.argv[2]) if len(sys.argv) > 2 else None


def read_launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return [(re.sub(r"\(.*", "", r["Kernel Name"]).replace("unnamed>::", "").replace("void ", ""), r["Grid Size"], float(r["Metric Value"]))
            for r in csv.DictReader(lines)]


rows = read_launches(f"gpurun_out/launches_{tag}.csv")
if per_rep:
    rows = rows[-per_rep:]
agg = collections.OrderedDict()
for k, g, v in rows:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, _, v in rows)
with open(f"profiles/launches_{tag}_summary.md", "w") as f:
    f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` on `python profiles/prof_target.py`\n\n")
    f.write(f"Second repetition only (first = warm-up): {len(rows)} launches, {tot / 1e6:.3f} ms of kernel time (cold-cache, serialised: compare SHARES).\n\n")
    f.write("| kernel | launches | total µs | share | avg µs |\n|---|---:|---:|---:|---:|\n")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {v / 1e3:.1f} | {v / tot:.3f} | {v / n / 1e3:.1f} |\n")
    f.write("\n## launch sequence of the first (largest) batch\n\n| kernel | grid | µs |\n|---|---|---:|\n")
    for k, g, v in rows[:32]:
        f.write(f"| `{k}` | {g} | {v / 1e3:.1f} |\n")

raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{tag}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units, data = r[0], r[1], r[2:]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "launch__grid_size"]
want = [w for w in want if w in col]
with open(f"profiles/ncu_full_{tag}_summary.md", "w") as f:
    f.write(f"# ncu --set full ({tag}): selected metrics per captured launch (`gpurun_out/prof_{tag}.ncu-rep`, not tracked)\n\n")
    f.write("| id | kernel | " + " | ".join(f"{w} [{units[col[w]]}]" for w in want) + " |\n|---|---|" + "---:|" * len(want) + "\n")
    for d in data:
        name = re.sub(r"\(.*", "", d[col["Kernel Name"]]).replace("unnamed>::", "").replace("void ", "")
        f.write(f"| {d[col['ID']]} | `{name}` | " + " | ".join(d[col[w]] for w in want) + " |\n")
print("written")
