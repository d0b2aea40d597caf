#!/bin/bash
# round 2, call j (8 GPUs): the Morton-range sharded map at N = 8 and N = 4 (ONE map, strong scaling), parity checked inside bench.py
TAG=${1:-r02j}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { name=$1; n=$2; shift; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "$name rc=$?"; tail -2 gpurun_out/bench_${TAG}_$name.err | cut -c1-300; }
run n8 8 A=1
run n8_share64 8 CHAD_SHARD_RANK0_SHARE=64
run n4 4 A=1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("parity_checked"))
        print(json.dumps(d.get("kernel_ms_per_step")))
        print(json.dumps(d["config"].get("nvlink")))
    except Exception as ex:  # noqa: BLE001
        print(f, "no line:", ex)
PY
