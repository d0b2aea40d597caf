"""Tier A at full size, measured on the REFERENCE itself: for every full-size golden case, the reference exactly as written
(std::sort, `verbatim`) against the canonical build (std::stable_sort, `stable`) -- how many voxels of the last active submap differ
in their fp32 distance and by how much (in units of sdf_trunc). Since the GPU path equals the canonical build bit for bit
(tests/test_gpu_golden.py), these are also the deviations of the GPU path from the reference as written. Adds the key
"tier_a" to every entry of golden_full.json. Run in the build container only (needs /root/reference): about four minutes."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import bindings as ob  # noqa: E402
from tests.golden import cases  # noqa: E402


def main():
    ob.build("all")
    assert ob.ref_available("stable") and ob.ref_available("verbatim"), "needs /root/reference"
    path = os.path.join(HERE, "golden_full.json")
    golden = json.load(open(path))
    for name in cases.FULL_CASES:
        scans, res, trunc = cases.case_scans(name)
        maps = {v: ob.RefMap(res, trunc, v) for v in ("stable", "verbatim")}
        for pts, pos in scans:
            for m in maps.values():
                m.insert(pts, pos)
        (ks, ss, ws), (kv, sv, wv) = maps["stable"].voxels(), maps["verbatim"].voxels()
        assert np.array_equal(ks, kv) and np.array_equal(ws, wv), "tie order must not change the voxel set or the weights"
        d = np.abs(ss.view(np.float32).astype(np.float64) - sv.view(np.float32).astype(np.float64))
        for m in maps.values():
            m.finalize_active()
        ds, dv = ob.map_digest(maps["stable"]), ob.map_digest(maps["verbatim"])
        golden[name]["tier_a"] = {
            "voxels": int(len(ks)), "voxels_with_different_sd_bits": int((ss != sv).sum()),
            "max_abs_sd_difference_over_trunc": float(d.max() / trunc),
            "voxels_beyond_1e-5_trunc": int((d > 1e-5 * trunc).sum()),
            "voxels_beyond_1e-2_trunc": int((d > 1e-2 * trunc).sum()),
            "dag_levels_identical": [a == b for a, b in zip(ds["levels"], dv["levels"])],
            "roots_identical": ds["roots"] == dv["roots"],
        }
        print(name, golden[name]["tier_a"])
        for m in maps.values():
            m.close()
    with open(path, "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
