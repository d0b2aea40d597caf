import os, sys, time
sys.path.insert(0, "/root/repo")
import torch, numpy as np
from chad_tsdf_b200 import TSDFMap, synth
w = synth.WORKLOADS["cfg1_traj100_128beam"]
scans = [w.scan(s) for s in range(w.scans)]
m = TSDFMap(w.sdf_res, w.sdf_trunc, max_batch_scans=24)
pinned = []
for pts, _ in scans:
    t = torch.empty((len(pts), 3), dtype=torch.float32, pin_memory=True); t.numpy()[...] = pts; pinned.append(t)
ptrs = []
for pts, _ in scans:
    p = m.device_alloc(pts.nbytes); m.upload(p, pts); ptrs.append((p, len(pts)))
for mode in ("device", "host_async", "host_sync", "host_async", "device"):
    for rep in range(3):
        m.reset()
        t0 = time.perf_counter()
        marks = []
        for i, ((t, (_, pos)), (p, n)) in enumerate(zip(zip(pinned, scans), ptrs)):
            if mode == "device": m.insert_device(p, n, pos)
            else: m.insert(t, pos, mode == "host_sync")
            if i in (3, 20, 41, 62, 83): marks.append(round(1e3 * (time.perf_counter() - t0), 2))
        t1 = time.perf_counter()
        m.flush()
        t2 = time.perf_counter()
    print(mode, "loop ms", round(1e3*(t1-t0),2), "flush ms", round(1e3*(t2-t1),2), "total", round(1e3*(t2-t0),2), "marks", marks)
