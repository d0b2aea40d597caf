#!/bin/bash
# round 2, call i (1 GPU): levels kernel solo threshold sweep (value leg only), fold at 4 CTAs per SM
TAG=${1:-r02i}
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --quick > gpurun_out/sweep_${TAG}_$name.json 2> gpurun_out/sweep_${TAG}_$name.err; echo "$name rc=$?"; }
run base A=1
run solo512 CHAD_LEVELS_SOLO=512
run solo2k CHAD_LEVELS_SOLO=2048
run solo64 CHAD_LEVELS_SOLO=64
run lv256 CHAD_LEVELS_THREADS=256
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        k = d.get("kernel_ms_per_step", {})
        print(f.split("_")[-1][:-5].ljust(10), round(d["ms_per_step"], 3), "fold", k.get("runs_fold_kernel"), "fin", [v for kk, v in k.items() if kk.startswith("finalize")], "frac", round(d["roofline"]["frac"], 3))
    except Exception as ex:  # noqa: BLE001
        print(f, "no line:", ex)
PY
