"""Where a kernel's warp instructions go, from an ncu report captured with --import-source on: the SASS listing is cut into
contiguous regions of (nearly) equal execution count -- loop bodies, loop-invariant set-up, rarely taken branches -- and each
region is printed with its share of the executed instructions and of the stall samples.
usage: python profiles/sass_regions.py <report.ncu-rep> <kernel name substring> [regions]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(lines[start:end]))
seq = []
for r in rows:
    try:
        seq.append((int(r["Instructions Executed"]), r["Source"].strip(), r["Avg. Threads Executed"], int(r["# Samples"] or 0)))
    except ValueError:
        pass
tot = sum(s[0] for s in seq)
smp = max(1, sum(s[3] for s in seq))
print(f"{kern}: {tot} warp instructions, {len(seq)} SASS lines, {smp} stall samples")
regions, i = [], 0
while i < len(seq):
    j, n, q = i, 0, 0
    while j < len(seq) and abs(seq[j][0] - seq[i][0]) <= 0.03 * max(seq[i][0], 1):
        n += seq[j][0]
        q += seq[j][3]
        j += 1
    regions.append((n, i, j, seq[i][0], q))
    i = j
print("| share of instructions | share of samples | SASS lines | executions | threads | first instruction |\n|---:|---:|---|---:|---:|---|")
for n, i, j, ex, q in sorted(regions, reverse=True)[:top]:
    print(f"| {n / tot:.3f} | {q / smp:.3f} | {i}-{j - 1} ({j - i}) | {ex} | {seq[i][2]} | `{seq[i][1][:60]}` |")
print("\nhottest instructions by stall samples:")
for s in sorted(seq, key=lambda x: -x[3])[:10]:
    print(f"  {s[3] / smp:.3f}  executed {s[0]:9d}  threads {s[2]:>3s}  {s[1][:80]}")
