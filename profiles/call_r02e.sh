#!/bin/bash
# round 2, call e (2 GPUs): sharded map with the synchronous gather; reader / persistence / 2-GPU facade tests; bench N = 2 and N = 1
TAG=${1:-r02e}
mkdir -p gpurun_out
CHAD_TRACE=1 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/sharded_worker.py gpurun_out morton: > gpurun_out/worker_${TAG}.log 2>&1; echo "traced worker rc=$?"; grep -c "chad r" gpurun_out/worker_${TAG}.log; cat gpurun_out/result0.json 2>/dev/null | head -c 700; echo
timeout 500 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_readers.py tests/test_facade.py -m gpu -q -k "world2 or readers or facade or two_gpus" > gpurun_out/t_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/t_${TAG}.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/bench_${TAG}_n2.err
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench n1 rc=$?"; tail -3 gpurun_out/bench_${TAG}_n1.err
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_${TAG}_n1c8.json 2> gpurun_out/bench_${TAG}_n1c8.err; echo "bench n1 (8 connections) rc=$?"
python - <<PY
import json
for n in ("bench_${TAG}_n2", "bench_${TAG}_n1", "bench_${TAG}_n1c8"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("parity_checked"), d.get("parity"))
        print(json.dumps(d.get("kernel_ms_per_step")))
        print(json.dumps(d["config"].get("nvlink")), json.dumps(d.get("memory")))
    except Exception as ex:  # noqa: BLE001
        print(n, "no line:", ex)
PY
