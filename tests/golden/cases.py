"""The golden cases: small, deterministic workloads whose reference outputs are pinned in golden.json.
Shared by make_golden.py (which runs the REFERENCE build here) and by the tests (oracle + GPU)."""
import hashlib

import numpy as np

from chad_tsdf_b200 import synth


def _w(name, scene, beams, scans, start, step, res, trunc):
    return synth.Workload(name, scene, beams, scans, start, step, res, trunc)


# name -> (workload, explicit finalize after the last scan)
CASES = {
    # BASELINE.json configs[0]: one 64-beam scan into TSDFMap(0.05, 0.10)
    "cfg0_single_64beam": _w("cfg0_single_64beam", synth.BOX_ROOM, 64, 1, 0.0, 0.0, 0.05, 0.10),
    # configs[1] shape, reduced to 32 beams: 24 scans at 0.25 m => one submap switch inside insert
    "cfg1_traj24_32beam": _w("cfg1_traj24_32beam", synth.BOX_ROOM, 32, 24, -12.5, 0.25, 0.05, 0.10),
    # configs[2]: fine voxels, trunc/res = 3, dense indoor
    "cfg2_fine_indoor_3": _w("cfg2_fine_indoor_3", synth.INDOOR, 64, 3, -1.0, 0.1, 0.02, 0.06),
    # configs[3] shape: 0.10 m voxels, 1 m/scan => two submap switches in 13 scans
    "cfg3_urban13_32beam": _w("cfg3_urban13_32beam", synth.URBAN, 32, 13, 0.0, 1.0, 0.10, 0.20),
}
SPHERE_POINTS = 100_000  # the reference demo's workload shape (main.cpp:7-38)

# BASELINE.json's configs at FULL size (or, for the two that take the reference minutes to hours, a prefix long enough to
# cross many submap switches), pinned in golden_full.json by make_golden.py --full. The GPU path is checked against
# these pins without running any CPU model on the GPU box; the oracle restatement is checked against them on the CPU.
FULL_CASES = {
    # configs[1] = the bench workload, whole: 100 scans x 262 144 points, four submap switches inside insert
    "full_cfg1_traj100_128beam": synth.WORKLOADS["cfg1_traj100_128beam"],
    # configs[2], whole: 20 dense indoor scans at 0.02 m voxels / 0.06 m truncation
    "full_cfg2_fine_indoor": synth.WORKLOADS["cfg2_fine_indoor"],
    # configs[3], the first 120 m of the 5 km drive: 19 submap switches inside insert, 0.10 m voxels
    "full_cfg3_urban_first120": synth.WORKLOADS["cfg3_urban_5km"].truncated(120),
    # configs[4]'s trajectory (0.025 m per scan), first 240 scans: submaps of 201 scans = nine default batches each
    "full_cfg4_traj1000_first240": synth.WORKLOADS["cfg4_traj1000_128beam"].truncated(240),
}


def case_scans(name):
    """[(points, pose)] of a golden case."""
    if name == "sphere_demo_100k":
        return [(synth.sphere_demo_points(SPHERE_POINTS), np.zeros(3, np.float32))], 0.05, 0.10
    w = CASES[name] if name in CASES else FULL_CASES[name]
    return [w.scan(s) for s in range(w.scans)], w.sdf_res, w.sdf_trunc


ALL_CASES = list(CASES) + ["sphere_demo_100k"]


def input_digest(scans) -> str:
    d = hashlib.sha256()
    for pts, pos in scans:
        d.update(np.ascontiguousarray(pts).tobytes())
        d.update(np.ascontiguousarray(pos).tobytes())
    return d.hexdigest()


def run_case(map_factory, name):
    """Insert the case into a fresh map, finalize the active submap, return (map, per-scan voxel digests)."""
    from oracle.bindings import map_digest
    scans, res, trunc = case_scans(name)
    m = map_factory(res, trunc)
    for pts, pos in scans:
        m.insert(pts, pos)
    before = map_digest(m)  # active submap's voxels before the final finalize
    m.finalize_active()
    after = map_digest(m)
    return m, {"before_finalize": {k: before[k] for k in ("voxels_n", "voxels_keys", "voxels_sd_bits", "voxels_weights", "weight_sum")},
               "final": after}
