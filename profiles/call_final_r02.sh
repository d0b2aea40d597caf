#!/bin/bash
# round 2, final evidence set on ONE fresh B200 (about 17 minutes):
#   gpurun --timeout 1500 -- 'bash profiles/call_final_r02.sh r02z'
# then, here:  python profiles/summarize.py r02z <launches per repetition>
# 1. GPU tests  2. smoke  3. bench line of the metric's config (both legs, parity, CPU baseline)  4. reference arm (short config)
# 5. bench lines of the other BASELINE configs, the two long ones with memory / growth figures  6. ncu launch list + full capture  7. timeline
TAG=${1:-r02z}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/t_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_$TAG.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_cfg1.json 2> gpurun_out/bench_${TAG}_cfg1.err; echo "bench cfg1 rc=$?"
timeout 100 python bench.py --impl reference --workload cfg0 --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_ref_cfg0.json 2> gpurun_out/bench_${TAG}_ref_cfg0.err; echo "reference arm (cfg0) rc=$?"
timeout 120 python bench.py --workload cfg0 --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_cfg0.json 2> gpurun_out/bench_${TAG}_cfg0.err; echo "bench cfg0 rc=$?"
timeout 150 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_cfg2.json 2> gpurun_out/bench_${TAG}_cfg2.err; echo "bench cfg2 rc=$?"
timeout 330 python bench.py --workload cfg3:1000 --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_cfg3.json 2> gpurun_out/bench_${TAG}_cfg3.err; echo "bench cfg3:1000 rc=$?"
timeout 330 python bench.py --workload cfg4:1000 --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_cfg4.json 2> gpurun_out/bench_${TAG}_cfg4.err; echo "bench cfg4:1000 rc=$?"
python - <<PY
import json
for n in ("cfg1", "ref_cfg0", "cfg0", "cfg2", "cfg3", "cfg4"):
    try:
        d = json.loads(open(f"gpurun_out/bench_${TAG}_{n}.json").read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(n, "value %.4g" % d.get("value"), "ms %.3f" % d.get("ms_per_step"), "e2e %.4g" % (d.get("e2e") or {}).get("value"), "frac", r.get("frac"), (r.get("pipeline") or {}).get("frac"),
              "parity", d.get("parity_checked"), "cpu", (d.get("cpu_baseline") or {}).get("value"), d.get("clocks"))
        if d.get("memory"): print("   memory", json.dumps(d["memory"]))
    except Exception as ex:  # noqa: BLE001
        print(n, "no line:", ex, open(f"gpurun_out/bench_${TAG}_{n}.err").read()[-600:])
PY
# a number printed under ncu is never a bench value: these two only produce the launch list and the counters
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv \
    python profiles/prof_target.py > gpurun_out/ncu_${TAG}_1.log 2>&1; echo "ncu launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"runs_emit|runs_fold|dag_levels" --launch-skip 3 -c 3 \
    -o gpurun_out/prof_$TAG -f python profiles/prof_target.py > gpurun_out/ncu_${TAG}_2.log 2>&1; echo "ncu full rc=$?"
timeout 90 python profiles/timeline.py 24 > gpurun_out/timeline_$TAG.txt 2>&1; echo "timeline rc=$?"
