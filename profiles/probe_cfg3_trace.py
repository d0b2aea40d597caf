"""Third diagnostic for cfg3:1000: host traces (CHAD_TRACE=1) of the paced and the fast device-resident insert of the same trajectory:
batch composition and the chunk count of every closed submap. python profiles/probe_cfg3_trace.py <paced|fast> [scans]"""
import os
import sys

os.environ["CHAD_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chad_tsdf_b200 import TSDFMap, synth  # noqa: E402

mode = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
w = synth.WORKLOADS["cfg3_urban_5km"].truncated(n)
scans = bench.generate_scans(w)
g = TSDFMap(w.sdf_res, w.sdf_trunc)
ptrs = []
for pts, _ in scans:
    p = g.device_alloc(pts.nbytes)
    g.upload(p, pts)
    ptrs.append((p, len(pts)))
for rep in range(2):  # first repetition grows every buffer
    g.reset()
    print(f"==== {mode} repetition {rep}", file=sys.stderr, flush=True)
    for (p, m), (_, pos) in zip(ptrs, scans):
        g.insert_device(p, m, pos)
        if mode == "paced":
            g.flush()
    g.flush()
    g.finalize_active()
print(mode, "roots", len(g.roots()), g.stats())
