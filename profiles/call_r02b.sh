#!/bin/bash
# round 2, call b (2 GPUs): the Morton-range sharded map (C++ / NCCL) against the oracle, single-GPU regressions of the scan-aligned
# ray tiles, bench at N = 1 and N = 2 (ONE map) with the parity check
TAG=${1:-r02b}
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/t_${TAG}_sharded.log 2>&1; echo "sharded pytest rc=$?"; tail -15 gpurun_out/t_${TAG}_sharded.log
timeout 400 python -m pytest tests/test_gpu_stages.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/t_${TAG}_parity.log 2>&1; echo "parity pytest rc=$?"; tail -4 gpurun_out/t_${TAG}_parity.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench n1 rc=$?"; tail -3 gpurun_out/bench_${TAG}_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/bench_${TAG}_n2.err
python - <<PY
import json
for n in ("bench_${TAG}_n1", "bench_${TAG}_n2"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("parity_checked"), d.get("parity"))
        print(json.dumps(d.get("kernel_ms_per_step")))
        print(json.dumps(d["config"].get("nvlink")))
    except Exception as ex:  # noqa: BLE001
        print(n, "no line:", ex)
PY
