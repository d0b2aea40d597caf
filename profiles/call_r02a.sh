#!/bin/bash
# round 2, first call: GPU tests (sweep un-gated, cooperative levels kernel), smoke, short bench, A/B of the cooperative launch
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/t_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_$TAG.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
CHAD_LEVELS_COOP=0 timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_nocoop.json 2> gpurun_out/bench_${TAG}_nocoop.err; echo "bench nocoop rc=$?"
python - <<PY
import json
for n in ("bench_$TAG", "bench_${TAG}_nocoop"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
        print(json.dumps(d.get("kernel_ms_per_step")))
    except Exception as ex:  # noqa: BLE001
        print(n, "no line:", ex)
PY
