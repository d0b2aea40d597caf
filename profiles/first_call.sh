#!/bin/bash
# One gpurun call that re-establishes the whole evidence set on a fresh B200 (about 8 minutes):
#   gpurun --timeout 900 -- 'bash profiles/first_call.sh r02a'
# then, here:  python profiles/summarize.py r02a 78   (tracked summaries under profiles/)
# 1. GPU tests  2. smoke  3. bench line (both arms)  4. ncu launch list  5. ncu --set full of the three hot kernels  6. device timeline
TAG=${1:-rXX}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/t_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/t_$TAG.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
# the random-scene parameter sweep through the CUDA path (written after round 1's last GPU minute; make it unconditional once green)
CHAD_GPU_SWEEP=1 timeout 120 python -m pytest tests/test_zz_saved_map_gpu.py -m gpu -q > gpurun_out/t_${TAG}_sweep.log 2>&1; echo "sweep rc=$?"; tail -2 gpurun_out/t_${TAG}_sweep.log
timeout 120 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err; echo "reference arm rc=$?"
python - <<PY
import json
for n in ("bench_$TAG", "bench_${TAG}_ref"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), d.get("clocks"))
    except Exception as ex:  # noqa: BLE001
        print(n, "no line:", ex)
PY
# a number printed under ncu is never a bench value: these two only produce the launch list and the counters
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv \
    python profiles/prof_target.py > gpurun_out/ncu_${TAG}_1.log 2>&1; echo "ncu launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"runs_emit|runs_fold|dag_levels" --launch-skip 3 -c 3 \
    -o gpurun_out/prof_$TAG -f python profiles/prof_target.py 16 > gpurun_out/ncu_${TAG}_2.log 2>&1; echo "ncu full rc=$?"
timeout 90 python profiles/timeline.py 24 > gpurun_out/timeline_$TAG.txt 2>&1; echo "timeline rc=$?"
