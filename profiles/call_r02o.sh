#!/bin/bash
# round 2, call o (N GPUs): slices assembled by ONE grouped all-gather per batch -- bench at N ranks; with N = 2 also a world-2 parity test and the C++ class on two GPUs
TAG=${1:-r02o}; N=${2:-2}
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; rc=$?; echo "n$N rc=$rc"; tail -2 gpurun_out/bench_${TAG}_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${TAG}_n$N.json").read().strip().splitlines()[-1])
    print(d.get("value"), d.get("ms_per_step"), json.dumps(d.get("e2e")), d.get("parity_checked"))
    print(json.dumps(d.get("kernel_ms_per_step")))
except Exception as ex:  # noqa: BLE001
    print("no line:", ex)
PY
if [ $rc -ne 0 ]; then echo "bench failed: tests skipped"; exit 1; fi
if [ $N -eq 2 ]; then
timeout 120 python -m pytest tests/test_gpu_sharded.py -m gpu -q -k "(test_sharded_world2 and not fine) or two_gpus" > gpurun_out/t_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_${TAG}.log
fi
