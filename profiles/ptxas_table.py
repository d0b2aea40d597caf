"""ptxas -v of every kernel, as a table: python profiles/ptxas_table.py r02 > profiles/ptxas_r02.md  (build container; no GPU needed)."""
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FLAGS = ["-std=c++20", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false", "-Xptxas", "-v", "-c", "-o", "/dev/null", "-I", os.path.join(ROOT, "include")]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"\(.*", "", o.replace("(anonymous namespace)::", "")).replace("chadgpu::", "").replace("void ", "") for o in out]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    rows = []
    for src in sorted(glob.glob(os.path.join(ROOT, "chad_tsdf_b200", "csrc", "*.cu"))):
        err = subprocess.run(["nvcc"] + FLAGS + [src], capture_output=True, text=True).stderr
        cur = None
        for line in err.splitlines():
            m = re.search(r"Compiling entry function '(\w+)'", line)
            if m:
                cur = {"name": m.group(1), "file": os.path.basename(src), "stack": 0, "ss": 0, "sl": 0, "regs": 0, "smem": 0}
                rows.append(cur)
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                cur["stack"], cur["ss"], cur["sl"] = map(int, m.groups())
            m = re.search(r"Used (\d+) registers", line)
            if m:
                cur["regs"] = int(m.group(1))
                s = re.search(r"(\d+) bytes smem", line)
                cur["smem"] = int(s.group(1)) if s else 0
    names = demangle([r["name"] for r in rows])
    for r, n in zip(rows, names):
        r["pretty"] = n
    rows.sort(key=lambda r: (-r["regs"], r["pretty"]))
    print(f"# ptxas -v (sm_100a, the flags of chad_tsdf_b200/build.py): registers, static shared memory, spills per kernel -- round {tag[1:]}\n")
    print("`nvcc " + " ".join(FLAGS[:8]) + " -c csrc/<file>.cu` in the build container, final code of the round (`python profiles/ptxas_table.py`).\n")
    print("| kernel | file | registers | static smem [B] | stack [B] | spill stores / loads [B] |\n|---|---|---:|---:|---:|---:|")
    for r in rows:
        print(f"| `{r['pretty']}` | `{r['file']}` | {r['regs']} | {r['smem']} | {r['stack']} | {r['ss']} / {r['sl']} |")
    spilled = [r["pretty"] for r in rows if r["ss"] or r["sl"]]
    print(f"\n{len(rows)} kernels; spills: {', '.join(spilled) if spilled else 'none'}.")


if __name__ == "__main__":
    main()
