"""Second diagnostic for the cfg3:1000 parity failure of the evidence run: the host-paced probe (profiles/probe_cfg3_long.py) found the GPU
map equal to the reference build, so the question is whether the result depends on how fast the scans arrive. The same trajectory is
inserted (a) host-paced, with a flush after every scan, (b) as fast as the host can call chad_insert_device (what bench.py does), twice;
all three maps are compared with the reference's pin and with each other. python profiles/probe_cfg3_timing.py [scans]"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chad_tsdf_b200 import TSDFMap, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
w = synth.WORKLOADS["cfg3_urban_5km"].truncated(n)
scans = bench.generate_scans(w)
pin = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_long.json")))["long_cfg3_urban_first1000"]["stable"]["final"] if n == 1000 else None
print("env:", {k: v for k, v in os.environ.items() if k.startswith("CHAD_")})
g = TSDFMap(w.sdf_res, w.sdf_trunc)
ptrs = []
for pts, _ in scans:
    p = g.device_alloc(pts.nbytes)
    g.upload(p, pts)
    ptrs.append((p, len(pts)))


def digest():
    g.finalize_active()
    out = []
    for lv in range(21):
        a, u, d = g.level(lv)
        out.append({"words": int(len(a)), "uniques": int(u), "dupes": int(d), "sha256": hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()})
    return out, g.roots()


def levels():
    return [g.level(lv)[0].copy() for lv in range(21)]


results = {}
for name in ("paced", "fast1", "fast2", "fast_host"):
    g.reset()
    if name == "paced":
        closes = []  # words per level after every scan that closed a submap
        nsub = 0
        for pts, pos in scans:
            g.insert(pts, pos)
            g.flush()
            k = g.stats()["submaps"]
            if k != nsub:
                nsub = k
                closes.append([g.level(lv)[1] for lv in (19, 20)])
    elif name == "fast_host":
        g.insert_many(scans, True)
    else:
        for (p, m), (_, pos) in zip(ptrs, scans):
            g.insert_device(p, m, pos)
    g.flush()
    d, roots = digest()
    results[name] = (d, roots, levels())
    bad = [lv for lv in range(21) if pin and d[lv] != pin["levels"][lv]]
    print(f"{name}: vs pin: differing levels {bad}; roots equal {pin is None or [list(r) for r in roots] == pin['roots']}")
ref = results["paced"]
for name in ("fast1", "fast2", "fast_host"):
    d, roots, lv_arrays = results[name]
    bad = [lv for lv in range(21) if d[lv] != ref[0][lv]]
    print(f"{name} vs paced: differing levels {bad}")
    for lv in bad[-2:]:
        a, b = lv_arrays[lv], ref[2][lv]
        m = min(len(a), len(b))
        diff = np.nonzero(a[:m] != b[:m])[0]
        print(f"  level {lv}: {len(a)} vs {len(b)} words, uniques {d[lv]['uniques']} vs {ref[0][lv]['uniques']}, dupes {d[lv]['dupes']} vs {ref[0][lv]['dupes']}, {len(diff)} differing words, first at {int(diff[0]) if len(diff) else None}")
        if len(diff):
            i = int(diff[0])
            print(f"    {name} {a[max(0, i - 2):i + 5].tolist()}\n    paced {b[max(0, i - 2):i + 5].tolist()}")
            if lv == 20:
                print("    same multiset:", np.array_equal(np.sort(a[1:m]), np.sort(b[1:m])))
            cum = [c[1 if lv == 20 else 0] for c in closes]
            print("    first difference falls into submap", int(np.searchsorted(np.array(cum), i if lv == 20 else 0)), "(cluster level: by unique count at each close)")
