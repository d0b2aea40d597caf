#!/bin/bash
# round 2, call m (1 GPU): the stash-width fix -- the two new width tests, the urban drive at 1000 scans against the reference pin, cfg1
TAG=${1:-r02m}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "descriptor_keys" > gpurun_out/t_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_${TAG}.log
timeout 330 python bench.py --workload cfg3:1000 --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_cfg3.json 2> gpurun_out/bench_${TAG}_cfg3.err; echo "bench cfg3:1000 rc=$?"
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_cfg1.json 2> gpurun_out/bench_${TAG}_cfg1.err; echo "bench cfg1 rc=$?"
python - <<PY
import json
for n in ("cfg3", "cfg1"):
    try:
        d = json.loads(open(f"gpurun_out/bench_${TAG}_{n}.json").read().strip().splitlines()[-1])
        print(n, "value %.4g" % d.get("value"), "ms %.3f" % d.get("ms_per_step"), "e2e %.4g" % (d.get("e2e") or {}).get("value"), "parity", d.get("parity_checked"), d.get("parity")[-120:])
    except Exception as ex:  # noqa: BLE001
        print(n, "no line:", ex)
PY
