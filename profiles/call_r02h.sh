#!/bin/bash
# round 2, call h (1 GPU): knob sweep of the persistent kernels that run beside the main stream (value leg only: --quick)
TAG=${1:-r02h}
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 120 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-parity --quick > gpurun_out/sweep_${TAG}_$name.json 2> gpurun_out/sweep_${TAG}_$name.err; echo "$name rc=$?"; }
run base A=1
run fold4 CHAD_FOLD_CTAS=4
run fold5 CHAD_FOLD_CTAS=5
run fold2 CHAD_FOLD_CTAS=2
run solo32k CHAD_LEVELS_SOLO=32768
run solo128k CHAD_LEVELS_SOLO=131072
run lv1024 CHAD_LEVELS_THREADS=1024 CHAD_LEVELS_CTAS=74
run first24 CHAD_FIRST_BATCH=24
run batch12 A=1 B=1
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
