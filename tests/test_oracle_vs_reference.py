"""CPU: the oracle restatement against the reference's own sources, live (oracle/_ref, built from /root/reference
where that exists; the prebuilt libraries travel to the GPU box). Skipped when the libraries are absent."""
import numpy as np
import pytest

from chad_tsdf_b200 import synth
from oracle import bindings as ob

needs_ref = pytest.mark.skipif(not (ob.ref_available("stable") and ob.ref_available("verbatim")), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("beams,scans,res,trunc,scene,step", [
    (32, 6, 0.05, 0.10, synth.BOX_ROOM, 1.1),   # a submap switch at scan 5
    (16, 3, 0.02, 0.06, synth.INDOOR, 0.1),
    (32, 4, 0.10, 0.20, synth.URBAN, 2.0),
])
def test_restatement_is_bit_exact_vs_stable_reference(oracle_lib, beams, scans, res, trunc, scene, step):
    w = synth.Workload("t", scene, beams, scans, -1.0, step, res, trunc, seed=77)
    r, o = ob.RefMap(res, trunc, "stable"), ob.OracleMap(res, trunc)
    for s in range(scans):
        pts, pos = w.scan(s)
        assert r.insert(pts, pos) == o.insert(pts, pos)
        for a, b in zip(r.voxels(), o.voxels()):
            assert np.array_equal(a, b)
    r.finalize_active(); o.finalize_active()
    assert ob.map_digest(r) == ob.map_digest(o)
    r.close(); o.close()


@needs_ref
def test_tier_a_against_verbatim_reference(oracle_lib):
    """Reference exactly as written (unstable std::sort): integers exact, floats within 1e-5 * sdf_trunc."""
    w = synth.WORKLOADS["cfg0_single_64beam"]
    pts, pos = w.scan(0)
    r, o = ob.RefMap(w.sdf_res, w.sdf_trunc, "verbatim"), ob.OracleMap(w.sdf_res, w.sdf_trunc)
    r.insert(pts, pos); o.insert(pts, pos)
    rk, rs, rw = r.voxels()
    ok, os_, ow = o.voxels()
    assert np.array_equal(rk, ok) and np.array_equal(rw, ow)
    assert int(ow.sum()) == o.last_scan_stats()[0]  # weight conservation: sum of weights == emitted updates
    tol = 1e-5 * w.sdf_trunc
    assert np.abs(rs.view(np.float32) - os_.view(np.float32)).max() <= tol
    r.close(); o.close()


@needs_ref
def test_point_stage_vs_reference(oracle_lib):
    w = synth.WORKLOADS["cfg0_single_64beam"]
    pts, pos = w.scan(0)
    rxyz, rkeys, rnrm = ob.ref_stage_points(pts, pos, w.sdf_res, "stable")
    oxyz, okeys, order, onrm = ob.oracle_stage_points(pts, pos, w.sdf_res)
    assert np.array_equal(rkeys, okeys)
    assert np.array_equal(rxyz.view(np.uint32), oxyz.view(np.uint32))
    assert np.array_equal(rnrm.view(np.uint32), onrm.view(np.uint32))
    assert np.array_equal(pts[order], oxyz)
    assert np.all(np.diff(okeys.astype(np.int64)) <= 0)  # descending Morton (morton.hpp:85-89)


def _random_scene(rng, n, centre):
    """A few noisy planar patches and a blob around `centre` (float32 points): dense voxels, grazing rays, negative coordinates."""
    parts = []
    for _ in range(int(rng.integers(2, 5))):
        origin = centre + rng.uniform(-6.0, 6.0, 3)
        u, v = rng.normal(size=3), rng.normal(size=3)
        u /= np.linalg.norm(u)
        v -= u * (u @ v)
        v /= np.linalg.norm(v)
        k = n // 4
        ab = rng.uniform(-1.5, 1.5, (k, 2))
        parts.append(origin + ab[:, :1] * u + ab[:, 1:] * v + rng.normal(0.0, 0.004, (k, 3)))
    parts.append(centre + rng.normal(0.0, 0.3, (n // 8, 3)) + rng.uniform(-4.0, 4.0, 3))
    return np.ascontiguousarray(np.concatenate(parts).astype(np.float32))


@needs_ref
@pytest.mark.parametrize("seed", range(12))
def test_restatement_vs_reference_on_random_scenes(oracle_lib, seed):
    """Parameter sweep the fixed workloads do not cover: voxel sizes 0.03-0.2 m, truncation / voxel ratios 1-4, maps centred up to a
    kilometre from the origin in any octant, poses inside and outside the clouds, submap switches; oracle restatement against the
    reference's own sources (canonical tie-break), bit for bit: voxels after every insert, the whole DAG at the end."""
    rng = np.random.default_rng(1000 + seed)
    res = float(rng.choice([0.03, 0.05, 0.08, 0.2]))
    trunc = res * float(rng.choice([1.0, 1.5, 2.0, 3.0, 4.0]))
    centre = rng.uniform(-1.0, 1.0, 3) * float(rng.choice([0.0, 10.0, 1000.0]))
    r, o = ob.RefMap(res, trunc, "stable"), ob.OracleMap(res, trunc)
    pose = centre + rng.uniform(-2.0, 2.0, 3)
    for s in range(int(rng.integers(2, 5))):
        pts = _random_scene(rng, 6000, centre)
        pos = pose.astype(np.float32)
        assert r.insert(pts, pos) == o.insert(pts, pos)
        for a, b in zip(r.voxels(), o.voxels()):
            assert np.array_equal(a, b), f"seed {seed}: res {res} trunc {trunc} scan {s}"
        pose = pose + rng.uniform(-4.0, 4.0, 3)  # sometimes more than 5 m from the submap's first pose
    r.finalize_active(); o.finalize_active()
    assert ob.map_digest(r) == ob.map_digest(o)
    r.close(); o.close()
