"""Diagnostic for the one parity failure of the round-2 evidence run: configs[3], first 1000 scans (167 submaps), bench.py's hash check said
levels 13..20 differ from the reference pin. Runs the GPU map and the reference build (oracle/_ref, CPU) side by side, compares the level
counters at every submap close, and at the first difference prints where the two level arrays part. Then the same trajectory again after
chad_reset (bench.py re-uses one map for all its steps). Test infrastructure: python profiles/probe_cfg3_long.py [scans]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chad_tsdf_b200 import TSDFMap, synth  # noqa: E402
from oracle import bindings as ob  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
w = synth.WORKLOADS["cfg3_urban_5km"].truncated(n)
scans = bench.generate_scans(w)
g = TSDFMap(w.sdf_res, w.sdf_trunc)
r = ob.RefMap(w.sdf_res, w.sdf_trunc, "stable")


def g_counters(lv):
    words, u, d = C.c_size_t(), C.c_uint32(), C.c_uint32()
    g._check(g._lib.chad_level_words(g._h, lv, C.byref(words)))
    g._check(g._lib.chad_level_counters(g._h, lv, C.byref(u), C.byref(d)))
    return int(words.value), int(u.value), int(d.value)


def r_counters(lv):
    u, d = C.c_uint32(), C.c_uint32()
    r._f("level_counters")(r._h, lv, C.byref(u), C.byref(d))
    return int(r._f("level_words")(r._h, lv)), int(u.value), int(d.value)


def explain(lv):
    ga, ra = g.level(lv)[0], r.level(lv)[0]
    m = min(len(ga), len(ra))
    diff = np.nonzero(ga[:m] != ra[:m])[0]
    print(f"  level {lv}: gpu {len(ga)} words, reference {len(ra)} words, {len(diff)} differing words in the common prefix")
    if len(diff):
        i = int(diff[0])
        print(f"  first difference at word {i}: gpu {ga[max(0, i - 3):i + 6].tolist()}  reference {ra[max(0, i - 3):i + 6].tolist()}")
        if lv == 20:
            same_set = np.array_equal(np.sort(ga[1:m]), np.sort(ra[1:m]))
            print(f"  same multiset of cluster values in the common prefix: {same_set}")


table = []  # reference counters at every submap close
first_bad = None
for phase in ("fresh map", "after chad_reset"):
    if phase != "fresh map":
        g.reset()
        r.close()
        r = ob.RefMap(w.sdf_res, w.sdf_trunc, "stable")
    nsub = 0
    bad = None
    for s, (pts, pos) in enumerate(scans):
        g.insert(pts, pos)
        closed = r.insert(pts, pos)
        if closed != nsub:
            nsub = closed
            g.flush()
            for lv in range(21):
                gc, rc = g_counters(lv), r_counters(lv)
                if gc != rc:
                    bad = (s, nsub, lv, gc, rc)
                    break
            if bad:
                break
    if bad is None:
        g.finalize_active(); r.finalize_active()
        for lv in range(21):
            gc, rc = g_counters(lv), r_counters(lv)
            if gc != rc:
                bad = (len(scans), len(r.roots()), lv, gc, rc)
                break
        if bad is None:
            same = all(np.array_equal(g.level(lv)[0], r.level(lv)[0]) for lv in range(21)) and g.roots() == r.roots()
            print(f"{phase}: {len(r.roots())} submaps, all counters equal, all words equal: {same}")
            if not same:
                for lv in range(20, -1, -1):
                    if not np.array_equal(g.level(lv)[0], r.level(lv)[0]):
                        explain(lv)
            continue
    s, k, lv, gc, rc = bad
    print(f"{phase}: first difference after scan {s} (submap {k} closed): level {lv} (words, uniques, dupes) gpu {gc} reference {rc}")
    for q in range(20, max(lv - 1, -1), -1):
        print(f"  level {q}: gpu {g_counters(q)} reference {r_counters(q)}")
    explain(20)
    explain(lv)
    print("  memory", g.memory())
