"""GPU: TSDFMap::save -> chad::load_dag round trip: the file the C++ class writes on the GPU box is read back by the library's
host-side reader and queried voxel by voxel (green on B200; the reader itself is also checked on the CPU against the oracle,
tests/test_dag_persistence.py)."""
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_saved_map_round_trip(chad_lib, tmp_path):
    from chad_tsdf_b200 import build, capi
    demo, reader = build.build_facade_demo(), build.build_dag_reader()
    r = subprocess.run([demo, "150000"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lib = capi.load()
    # a 12^3 neighbourhood of the voxel where the 5 m sphere crosses the x axis, and the same neighbourhood far outside the band
    g = np.arange(-6, 6)
    near = np.array([lib.chad_morton_encode(100 + int(x), int(y), int(z)) for x in g for y in g for z in g], np.uint64)
    far = np.array([lib.chad_morton_encode(160 + int(x), int(y), int(z)) for x in g for y in g for z in g], np.uint64)
    out = {}
    for name, keys in (("near", near), ("far", far)):
        kf, of = tmp_path / f"{name}.u64", tmp_path / f"{name}.u8"
        keys.tofile(kf)
        q = subprocess.run([reader, str(tmp_path / "facade_demo.chad"), "0", str(kf), str(of)], capture_output=True, text=True, timeout=60)
        assert q.returncode == 0, q.stdout + q.stderr
        head = q.stdout.split()
        assert (np.float32(head[0]), np.float32(head[1]), int(head[2]), int(head[3]), int(head[4])) == (np.float32(0.05), np.float32(0.1), 1, 1, 10)
        out[name] = np.fromfile(of, np.uint8)
    assert np.all(out["far"] == 0xFF)
    present = out["near"] != 0xFF
    assert present.sum() > 20  # the truncation band is four voxels thick there
    # decoded distances follow the analytic sphere (inward normals: positive inside), cf. tests/cpp/facade_demo.cpp
    xyz = np.array([[100 + x, y, z] for x in g for y in g for z in g], np.float64) * 0.05
    sd = (out["near"][present].astype(np.float64) - 127.0) / 127.0 * 0.1
    expect = np.clip(5.0 - np.linalg.norm(xyz[present], axis=1), -0.1, 0.1)
    assert np.mean(np.abs(sd - expect) > 0.03) < 0.1


@pytest.mark.parametrize("seed", range(12))
def test_gpu_vs_oracle_on_random_scenes(chad_lib, oracle_lib, seed):
    """The parameter sweep of tests/test_oracle_vs_reference.py::test_restatement_vs_reference_on_random_scenes (voxel sizes 0.03-0.2 m,
    truncation / voxel ratios 1-4, maps up to a kilometre from the origin, submap switches) through the CUDA path, default settings."""
    from chad_tsdf_b200 import TSDFMap
    from tests.test_oracle_vs_reference import _random_scene
    from tests.test_gpu_parity import _assert_same_state
    rng = np.random.default_rng(1000 + seed)
    res = float(rng.choice([0.03, 0.05, 0.08, 0.2]))
    trunc = res * float(rng.choice([1.0, 1.5, 2.0, 3.0, 4.0]))
    centre = rng.uniform(-1.0, 1.0, 3) * float(rng.choice([0.0, 10.0, 1000.0]))
    g, o = TSDFMap(res, trunc, max_batch_scans=2), oracle_lib.OracleMap(res, trunc)
    pose = centre + rng.uniform(-2.0, 2.0, 3)
    for s in range(int(rng.integers(2, 5))):
        pts = _random_scene(rng, 6000, centre)
        pos = pose.astype(np.float32)
        g.insert(pts, pos); o.insert(pts, pos)
        pose = pose + rng.uniform(-4.0, 4.0, 3)
    _assert_same_state(g, o, check_levels=False)
    g.finalize_active(); o.finalize_active()
    _assert_same_state(g, o)
    g.close(); o.close()
