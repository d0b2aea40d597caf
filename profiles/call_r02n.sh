#!/bin/bash
# round 2, call n (2 GPUs): sharded inserts that copy 1 / world of a scan over the host link and all-gather the slices -- bench N = 2 (parity
# against the pins, e2e through host buffers), then the world-2 parity tests (pageable host buffers: the staging path of the same code)
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err; rc=$?; echo "n2 rc=$rc"; tail -2 gpurun_out/bench_${TAG}_n2.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_n2.json").read().strip().splitlines()[-1])
    print(d.get("value"), d.get("ms_per_step"), json.dumps(d.get("e2e")), d.get("parity_checked"))
except Exception as ex:  # noqa: BLE001
    print("no line:", ex)
PY
if [ $rc -ne 0 ]; then echo "bench failed: tests skipped"; exit 1; fi
timeout 170 python -m pytest tests/test_gpu_sharded.py -m gpu -q -k "world2 or two_gpus" > gpurun_out/t_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_${TAG}.log
