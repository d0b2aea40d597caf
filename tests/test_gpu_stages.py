"""GPU parity, kernel by kernel, through the C ABI (chad_stage_*), against the CPU oracle."""
import numpy as np
import pytest

from chad_tsdf_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gmap(chad_lib):
    from chad_tsdf_b200 import TSDFMap
    m = TSDFMap(0.05, 0.10)
    yield m
    m.close()


def test_morton_encode_matches_oracle(gmap, oracle_lib):
    rng = np.random.default_rng(1)
    vox = rng.integers(-(1 << 20) + 1, (1 << 20) - 1, size=(20000, 3), dtype=np.int32)
    vox[:8] = [[0, 0, 0], [-1, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [(1 << 20) - 1] * 3, [-(1 << 20)] * 3]
    keys = gmap.stage_morton(vox)
    want = np.array([oracle_lib.morton_encode(int(x), int(y), int(z)) for x, y, z in vox], dtype=np.uint64)
    assert np.array_equal(keys, want)
    assert keys[0] == 0x7000000000000000  # origin


@pytest.mark.parametrize("n", [1, 2, 31, 255, 256, 4095, 4096, 4097, 100_003, 1_500_000])
@pytest.mark.parametrize("nbits", [1, 8, 13, 36, 40, 64])
def test_radix_sort_stable(gmap, n, nbits):
    rng = np.random.default_rng(n * 131 + nbits)
    mask = np.uint64((1 << nbits) - 1) if nbits < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    keys = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) & mask
    if n > 1000:  # heavy duplicates + Morton-like coherence
        keys[: n // 2] = keys[: n // 2] & np.uint64(0xFFF)
        keys[n // 2:] = np.sort(keys[n // 2:])
    vals = np.arange(n, dtype=np.uint32)
    k, v = gmap.stage_sort(keys, vals, nbits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])  # stability: ties keep input order


def test_radix_sort_ignores_high_bits(gmap):
    rng = np.random.default_rng(5)
    n = 50_000
    keys = rng.integers(0, 1 << 20, size=n, dtype=np.uint64) | np.uint64(0xABC << 40)
    vals = np.arange(n, dtype=np.uint32)
    k, v = gmap.stage_sort(keys, vals, 20)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])


@pytest.mark.parametrize("wname,scan", [("cfg0_single_64beam", 0), ("cfg1_traj100_128beam", 7), ("cfg2_fine_indoor", 3), ("cfg3_urban_5km", 2)])
def test_point_stage_matches_oracle(chad_lib, oracle_lib, wname, scan):
    from chad_tsdf_b200 import TSDFMap
    w = synth.WORKLOADS[wname]
    pts, pos = w.scan(scan)
    with TSDFMap(w.sdf_res, w.sdf_trunc) as m:
        xyz, keys, order, nrm = m.stage_points(pts, pos)
    oxyz, okeys, oorder, onrm = oracle_lib.oracle_stage_points(pts, pos, w.sdf_res)
    assert np.array_equal(keys, okeys)
    assert np.array_equal(order, oorder)  # descending Morton, ties by input index
    assert np.array_equal(xyz.view(np.uint32), oxyz.view(np.uint32))
    assert not np.isnan(onrm).any()
    assert np.array_equal(nrm.view(np.uint32), onrm.view(np.uint32))  # bit-exact FP64 plane fit


@pytest.mark.parametrize("wname,scan", [("cfg0_single_64beam", 0), ("cfg2_fine_indoor", 1), ("cfg3_urban_5km", 4)])
def test_band_stage_matches_oracle(chad_lib, oracle_lib, wname, scan):
    from chad_tsdf_b200 import TSDFMap
    w = synth.WORKLOADS[wname]
    pts, pos = w.scan(scan)
    oxyz, okeys, oorder, onrm = oracle_lib.oracle_stage_points(pts, pos, w.sdf_res)
    want_keys, want_sd, want_counts = oracle_lib.oracle_stage_pairs(oxyz, onrm, pos, w.sdf_res, w.sdf_trunc)
    with TSDFMap(w.sdf_res, w.sdf_trunc) as m:
        keys, sd, counts = m.stage_pairs(oxyz, onrm, pos)
    assert np.array_equal(counts, want_counts)
    assert np.array_equal(keys, want_keys)  # same voxels in the same (point, ray step) order
    assert np.array_equal(sd.view(np.uint32), want_sd.view(np.uint32))


def test_small_and_degenerate_scans(chad_lib, oracle_lib):
    """n = 1 .. 9 points (the 'last point is never absorbed' quirk, normals.hpp:100) and many points in one voxel."""
    from chad_tsdf_b200 import TSDFMap
    rng = np.random.default_rng(3)
    pos = np.array([0.1, -0.2, 0.3], np.float32)
    with TSDFMap(0.05, 0.10) as m:
        for n in [1, 2, 7, 8, 9, 10, 33]:
            pts = (rng.random((n, 3)).astype(np.float32) * 0.04 + np.array([2.0, 1.0, 0.5], np.float32)).astype(np.float32)
            got = m.stage_points(pts, pos)
            want = oracle_lib.oracle_stage_points(pts, pos, 0.05)
            assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
            assert np.array_equal(got[3].view(np.uint32), want[3].view(np.uint32)), n
