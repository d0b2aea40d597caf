#!/bin/bash
# round 2, call l (2 GPUs): closes decoupled from the DAG stage -- world-2 parity tests (skewed ranks), bench N = 2
TAG=${1:-r02l}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py -m gpu -q -k "world2 or two_gpus" > gpurun_out/t_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_${TAG}.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err; echo "n2 rc=$?"; tail -2 gpurun_out/bench_${TAG}_n2.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_n2.json").read().strip().splitlines()[-1])
    print(d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("parity_checked"))
    print(json.dumps(d.get("kernel_ms_per_step")))
except Exception as ex:  # noqa: BLE001
    print("no line:", ex)
PY
