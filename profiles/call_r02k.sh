#!/bin/bash
# round 2, call k (4 GPUs): deferred closes -- world-2 / world-4 parity tests, bench at N = 4 and N = 2
TAG=${1:-r02k}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -q -k "world2 or world4 or two_gpus" > gpurun_out/t_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_${TAG}.log
run() { name=$1; n=$2; shift; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/bench_${TAG}_$name.err | cut -c1-300; }
run n4 4 A=1
run n2 2 A=1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("parity_checked"))
        print(json.dumps(d.get("kernel_ms_per_step")))
    except Exception as ex:  # noqa: BLE001
        print(f, "no line:", ex)
PY
