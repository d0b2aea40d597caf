"""Device timeline of one bench step from the library's own per-launch CUDA events (all streams):
python profiles/timeline.py [batch] -> prints launch order with begin/end/duration and the gap to the previous launch on the main stream."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chad_tsdf_b200 import TSDFMap, synth  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 24
host = len(sys.argv) > 2 and sys.argv[2] == "host"  # page-locked host inputs through chad_insert_async instead of device-resident ones
w = synth.WORKLOADS["cfg1_traj100_128beam"]
scans = [w.scan(s) for s in range(w.scans)]
m = TSDFMap(w.sdf_res, w.sdf_trunc, max_batch_scans=batch)
ptrs = []
for pts, _ in scans:
    p = m.device_alloc(pts.nbytes)
    m.upload(p, pts)
    ptrs.append((p, len(pts)))
pinned = []
if host:
    import torch
    for pts, _ in scans:
        t = torch.empty((len(pts), 3), dtype=torch.float32, pin_memory=True)
        t.numpy()[...] = pts
        pinned.append(t)
for rep in range(3):
    m.reset()
    if rep == 2:
        m.profile_enable(True)
    if host:
        for t, (_, pos) in zip(pinned, scans):
            m.insert(t, pos, False)
    else:
        for (p, n), (_, pos) in zip(ptrs, scans):
            m.insert_device(p, n, pos)
    m.flush()
tl = m.profile_timeline()
prev_end = 0.0
for name, a, b in tl:
    side = name.startswith("runs_fold") or name.startswith("finalize")
    gap = "" if side else f"gap {a - prev_end:7.3f}"
    print(f"{a:8.3f} {b:8.3f} {b - a:7.3f}  {'   [side] ' if side else ''}{name[:40]:40s} {gap}")
    if not side:
        prev_end = b
m.close()
