"""Per-opcode instruction histogram and stall summary of one kernel from an ncu report's SASS source page.
usage: python profiles/sass_hist.py <report.ncu-rep> <kernel name substring> [top]"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
lines = out.splitlines()
# several launches may match: take the first table
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(lines[start:end]))
ops = collections.Counter()
tot = 0
stalls = collections.Counter()
for r in rows:
    src = r["Source"].strip()
    parts = src.split()
    op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
    op = op.split(".")[0]
    n = int(r["Instructions Executed"] or 0)
    ops[op] += n
    tot += n
    for k, v in r.items():
        if k.startswith("stall_") and "Not Issued" not in k and v:
            stalls[k] += int(v)
print(f"{kern}: {tot} warp instructions, {len(rows)} SASS lines")
for op, n in ops.most_common(top):
    print(f"  {op:12s} {n:12d} {n / tot:6.3f}")
st = sum(stalls.values())
print("stalls:", ", ".join(f"{k[6:]} {v / st:.2f}" for k, v in stalls.most_common(8)))
# hottest SASS lines by samples
hot = sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:12]
for r in hot:
    print(f"  {int(r['# Samples']):6d} smp {int(r['Instructions Executed']):10d} ex thr {r['Avg. Threads Executed']:>5s}  {r['Source'].strip()[:90]}")
