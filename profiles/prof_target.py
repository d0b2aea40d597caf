"""Profiling target for ncu: one submap of the bench workload (21 scans of cfg1 = ONE batch at the library's default of 24 scans per
batch, i.e. the batch shape bench.py times) followed by Submap::finalize, run twice in-process (first run = warm-up).
Usage: python profiles/prof_target.py [scans] [max_batch_scans]"""
import os
import sys

os.environ.setdefault("CHAD_FIRST_BATCH", "64")  # no short first batch: the profile shows full batches

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chad_tsdf_b200 import TSDFMap, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 21
w = synth.WORKLOADS["cfg1_traj100_128beam"]
scans = [w.scan(s) for s in range(n)]
m = TSDFMap(w.sdf_res, w.sdf_trunc, max_batch_scans=int(sys.argv[2]) if len(sys.argv) > 2 else 24)
for rep in range(2):
    m.reset()
    m.reset_stats()
    for pts, pos in scans:
        m.insert(pts, pos)
    m.finalize_active()
    print(rep, m.stats())
m.close()
