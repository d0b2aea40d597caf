#!/bin/bash
# round 2, call f (2 GPUs): sharded map with device-side ordering of gather vs exchange, one-kernel splitters, sparse scatter
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_readers.py -m gpu -q -k "world2 or readers or two_gpus" > gpurun_out/t_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/t_${TAG}.log
for share in 256 176; do
CHAD_SHARD_RANK0_SHARE=$share timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_n2_s$share.json 2> gpurun_out/bench_${TAG}_n2_s$share.err; echo "bench n2 share $share rc=$?"; tail -3 gpurun_out/bench_${TAG}_n2_s$share.err
done
python - <<PY
import json
for n in ("bench_${TAG}_n2_s256", "bench_${TAG}_n2_s176"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("parity_checked"))
        print(json.dumps(d.get("kernel_ms_per_step")))
    except Exception as ex:  # noqa: BLE001
        print(n, "no line:", ex)
PY
