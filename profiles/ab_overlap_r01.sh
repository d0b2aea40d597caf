#!/bin/bash
# A/B of the walk-overlap pipeline (CHAD_OVERLAP_WALK) on one B200, then the GPU test suite under the
# configuration that won. Run with: gpurun --timeout 480 -- 'bash profiles/ab_overlap_r01.sh'. Everything lands in gpurun_out/.
mkdir -p gpurun_out
NEW="tests/test_gpu_golden.py::test_gpu_matches_reference_golden_at_full_size tests/test_gpu_parity.py::test_scan_beyond_the_tile_run_rank_range_then_ordinary_scans"
# 1. the new tests under the overlapped pipeline (full-size golden pins, giant scan followed by ordinary scans)
CHAD_OVERLAP_WALK=1 timeout 150 python -m pytest $NEW -x -q > gpurun_out/ab_new_ov1.log 2>&1; echo "new tests (overlap=1): rc=$?"
tail -3 gpurun_out/ab_new_ov1.log
# 2. bench: serial walk / overlapped walk / overlapped walk with half-submap batches
CHAD_OVERLAP_WALK=0 timeout 110 python bench.py > gpurun_out/ab_ov0_b24.json 2> gpurun_out/ab_ov0_b24.err
CHAD_OVERLAP_WALK=1 timeout 110 python bench.py > gpurun_out/ab_ov1_b24.json 2> gpurun_out/ab_ov1_b24.err
CHAD_OVERLAP_WALK=1 timeout 100 python bench.py --no-cpu-baseline --batch 12 > gpurun_out/ab_ov1_b12.json 2> gpurun_out/ab_ov1_b12.err
BEST=$(python - <<'PY'
import json
best, best_v = "0", 0.0
for name, ov in (("ab_ov0_b24", "0"), ("ab_ov1_b24", "1"), ("ab_ov1_b12", "1")):
    try:
        d = json.loads(open(f"gpurun_out/{name}.json").read().strip().splitlines()[-1])
        v, e = d["value"], d["e2e"]["value"]
    except Exception as ex:  # noqa: BLE001
        print(name, "no line:", ex, flush=True, file=open("gpurun_out/ab_summary.txt", "a"))
        continue
    print(name, "value %.4g points/s (%.3f ms/step), e2e %.4g (%.3f ms/step), clocks %s" % (v, d["ms_per_step"], e, d["e2e"]["ms_per_step"], d["clocks"]),
          file=open("gpurun_out/ab_summary.txt", "a"), flush=True)
    if name.endswith("b24") and v > best_v * (1.02 if ov == "1" else 1.0):  # the overlap must win by more than noise
        best, best_v = ov, v
print(best)
PY
)
cat gpurun_out/ab_summary.txt
echo "suite runs with CHAD_OVERLAP_WALK=$BEST"
# 3. the whole GPU suite under the winner (the new tests again only if they have not run under it yet)
if [ "$BEST" = "1" ]; then
  DESEL="--deselect tests/test_gpu_golden.py::test_gpu_matches_reference_golden_at_full_size --deselect tests/test_gpu_parity.py::test_scan_beyond_the_tile_run_rank_range_then_ordinary_scans"
else
  DESEL=""
fi
CHAD_OVERLAP_WALK=$BEST timeout 330 python -m pytest tests -m gpu -x -q $DESEL > gpurun_out/ab_suite.log 2>&1; echo "suite (overlap=$BEST): rc=$?"
tail -4 gpurun_out/ab_suite.log
echo "$BEST" > gpurun_out/ab_best.txt
CHAD_OVERLAP_WALK=$BEST timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ab_smoke.log 2>&1; echo "smoke: rc=$?"; tail -1 gpurun_out/ab_smoke.log
