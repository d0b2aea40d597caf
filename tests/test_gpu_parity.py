"""GPU parity of the whole path (insert -> fold -> finalize) through the C ABI against the CPU oracle:
bit-exact voxel keys, u32 weights, fp32 sd bits, every DAG level word, counters and roots."""
import numpy as np
import pytest

from chad_tsdf_b200 import synth

pytestmark = pytest.mark.gpu


def _assert_same_state(g, o, check_levels=True):
    gk, gs, gw = g.voxels()
    ok, os_, ow = o.voxels()
    assert len(gk) == len(ok)
    assert np.array_equal(gk, ok), "voxel key set differs"
    assert np.array_equal(gw, ow), "weights differ"
    assert np.array_equal(gs, os_), f"sd bits differ at {(gs != os_).sum()} voxels"
    assert g.roots() == o.roots()
    if check_levels:
        for lv in range(21):
            ga, gu, gd = g.level(lv)
            oa, ou, od = o.level(lv)
            assert (gu, gd) == (ou, od), f"level {lv} counters"
            assert np.array_equal(ga, oa), f"level {lv} words"


def _run_pair(w, scans, batch, finalize_every=None, path=0):
    from chad_tsdf_b200 import TSDFMap
    from oracle import bindings as ob
    g = TSDFMap(w.sdf_res, w.sdf_trunc, max_batch_scans=batch, pair_path=path)
    o = ob.OracleMap(w.sdf_res, w.sdf_trunc)
    U = V = 0
    for s in range(scans):
        pts, pos = w.scan(s)
        g.insert(pts, pos)
        o.insert(pts, pos)
        u, v = o.last_scan_stats()
        U += u
        V += v
        if finalize_every and (s + 1) % finalize_every == 0:
            g.finalize_active()
            o.finalize_active()
    return g, o, U, V


@pytest.mark.parametrize("path", [2, 0, 1])
@pytest.mark.parametrize("batch", [1, 4])
def test_single_scan_cfg0(chad_lib, oracle_lib, batch, path):
    w = synth.WORKLOADS["cfg0_single_64beam"]
    g, o, U, V = _run_pair(w, 1, batch, path=path)
    g.flush()
    _assert_same_state(g, o, check_levels=False)
    st = g.stats()
    assert st["updates"] == U and st["points"] == 131072
    if batch == 1:
        assert st["scan_voxels"] == V
    g.finalize_active(); o.finalize_active()
    _assert_same_state(g, o)
    assert g.roots() == [(1, 10)]
    g.close(); o.close()


@pytest.mark.parametrize("path", [2, 0, 1])
@pytest.mark.parametrize("batch", [1, 3, 16])
def test_trajectory_with_submap_switch(chad_lib, oracle_lib, batch, path):
    """cfg1 prefix: 24 scans at 0.25 m/scan => one submap switch inside insert (> 5 m from the first pose)."""
    w = synth.WORKLOADS["cfg1_traj100_128beam"]
    g, o, U, V = _run_pair(w, 24, batch, path=path)
    _assert_same_state(g, o)
    assert len(g.roots()) == 1
    g.finalize_active(); o.finalize_active()
    _assert_same_state(g, o)
    assert g.stats()["updates"] == U
    g.close(); o.close()


@pytest.mark.parametrize("path", [2, 0, 1])
def test_fine_voxels_cfg2(chad_lib, oracle_lib, path):
    w = synth.WORKLOADS["cfg2_fine_indoor"]
    g, o, _, _ = _run_pair(w, 5, 2, finalize_every=2, path=path)
    g.finalize_active(); o.finalize_active()
    _assert_same_state(g, o)
    g.close(); o.close()


@pytest.mark.parametrize("path", [2, 0, 1])
def test_urban_cfg3_many_submaps(chad_lib, oracle_lib, path):
    w = synth.WORKLOADS["cfg3_urban_5km"]
    g, o, _, _ = _run_pair(w, 14, 4, path=path)  # 1 m/scan: switches at scans 6 and 12
    assert len(g.roots()) == 2
    g.finalize_active(); o.finalize_active()
    _assert_same_state(g, o)
    g.close(); o.close()


@pytest.mark.parametrize("path", [2, 0, 1])
def test_sphere_demo_shape(chad_lib, oracle_lib, path):
    """The reference demo's workload shape (main.cpp:7-38): dense points on a 5 m sphere, many points per voxel."""
    from chad_tsdf_b200 import TSDFMap
    from oracle import bindings as ob
    pts = synth.sphere_demo_points(200_000)
    pos = np.zeros(3, np.float32)
    g, o = TSDFMap(0.05, 0.1, pair_path=path), ob.OracleMap(0.05, 0.1)
    g.insert(pts, pos); o.insert(pts, pos)
    g.finalize_active(); o.finalize_active()
    _assert_same_state(g, o)
    # analytic sanity (lvr2.cpp:81-85): decoded sd ~ +-(5 - |voxel|) within discretisation
    g.close(); o.close()


@pytest.mark.parametrize("path", [2, 0, 1])
def test_empty_and_tiny_inputs(chad_lib, oracle_lib, path):
    from chad_tsdf_b200 import TSDFMap
    from oracle import bindings as ob
    g, o = TSDFMap(0.05, 0.1, max_batch_scans=2, pair_path=path), ob.OracleMap(0.05, 0.1)
    pos = np.zeros(3, np.float32)
    empty = np.zeros((0, 3), np.float32)
    g.insert(empty, pos); o.insert(empty, pos)
    g.finalize_active(); o.finalize_active()  # empty octree: root record added twice (submap.hpp:31-46)
    _assert_same_state(g, o)
    assert g.roots() == [(1, 1)]
    one = np.array([[1.0, 2.0, 0.5]], np.float32)
    g.insert(one, pos); o.insert(one, pos)
    g.insert(empty, pos); o.insert(empty, pos)
    g.insert(one * 1.01, pos); o.insert(one * 1.01, pos)
    _assert_same_state(g, o)
    g.finalize_active(); o.finalize_active()
    _assert_same_state(g, o)
    g.close(); o.close()


def test_errors_are_reported(chad_lib):
    from chad_tsdf_b200 import TSDFMap, ChadError
    g = TSDFMap(0.05, 0.1)
    bad = np.array([[1e9, 0, 0], [1, 1, 1]], np.float32)  # outside the 21-bit Morton range
    g.insert(bad, np.zeros(3, np.float32))
    with pytest.raises(ChadError):
        g.flush()
    g.close()
    g = TSDFMap(0.05, 0.1)
    g.insert(np.array([[np.nan, 0, 0]], np.float32), np.zeros(3, np.float32))
    with pytest.raises(ChadError):
        g.flush()
    g.close()
    with pytest.raises(ChadError):
        TSDFMap(-1.0, 0.1)


def test_dense_voxels_many_updates_per_voxel(chad_lib, oracle_lib):
    """Thousands of points in a handful of voxels: long per-voxel segments (multi-pass block sort, fold look-ahead tail)."""
    from chad_tsdf_b200 import TSDFMap
    from oracle import bindings as ob
    rng = np.random.default_rng(11)
    pos = np.array([0.0, 0.0, 0.0], np.float32)
    for path in (2, 0, 1):
        g, o = TSDFMap(0.05, 0.1, max_batch_scans=4, pair_path=path), ob.OracleMap(0.05, 0.1)
        for s in range(4):
            pts = (rng.random((12000, 3)) * np.array([0.12, 0.12, 0.02]) + np.array([2.0, 1.0, 0.5])).astype(np.float32)
            g.insert(pts, pos); o.insert(pts, pos)
        _assert_same_state(g, o, check_levels=False)
        g.finalize_active(); o.finalize_active()
        _assert_same_state(g, o)
        g.close(); o.close()


def test_far_from_origin_wide_keys(chad_lib, oracle_lib):
    """A map 3.2 km from the origin: 17 significant bits per axis, so the sort keys of the points and of the tile-run descriptors are
    wide (more radix passes, no room to stash the run lengths in the descriptor keys)."""
    from chad_tsdf_b200 import TSDFMap
    from oracle import bindings as ob
    w = synth.WORKLOADS["cfg0_single_64beam"]
    shift = np.array([3200.0, -2900.0, 40.0], np.float32)
    for path in (2, 0):
        g, o = TSDFMap(w.sdf_res, w.sdf_trunc, max_batch_scans=2, pair_path=path), ob.OracleMap(w.sdf_res, w.sdf_trunc)
        for s in range(3):
            pts, pos = w.scan(s)
            pts, pos = (pts + shift).astype(np.float32), (pos + shift).astype(np.float32)
            g.insert(pts, pos); o.insert(pts, pos)
        _assert_same_state(g, o, check_levels=False)
        g.finalize_active(); o.finalize_active()
        _assert_same_state(g, o)
        g.close(); o.close()


def test_wide_truncation_band_falls_back_to_block_path(chad_lib, oracle_lib):
    """sdf_trunc / sdf_res = 4: a ray can cross more than four 8^3 blocks, which the tile-run path does not take; the default
    configuration must fall back to the block-binned path by itself and stay exact."""
    from chad_tsdf_b200 import TSDFMap
    from oracle import bindings as ob
    w = synth.Workload("wide", synth.BOX_ROOM, 32, 3, -1.0, 0.5, 0.05, 0.20, seed=21)
    g, o = TSDFMap(w.sdf_res, w.sdf_trunc, max_batch_scans=2), ob.OracleMap(w.sdf_res, w.sdf_trunc)
    for s in range(w.scans):
        pts, pos = w.scan(s)
        g.insert(pts, pos); o.insert(pts, pos)
    _assert_same_state(g, o, check_levels=False)
    g.finalize_active(); o.finalize_active()
    _assert_same_state(g, o)
    g.close(); o.close()


def test_device_dag_reader_matches_the_quantised_voxels(chad_lib, oracle_lib):
    """chad_query_voxels walks the finalised DAG on the device (get_child_addr / try_get_lc, levels.hpp:147-192): every voxel of the
    closed submap must come back with the byte cluster.hpp:13-26 quantises its distance to, every other key as 0xFF."""
    from chad_tsdf_b200 import TSDFMap
    from oracle import bindings as ob
    w = synth.WORKLOADS["cfg0_single_64beam"]
    g, o = TSDFMap(w.sdf_res, w.sdf_trunc), ob.OracleMap(w.sdf_res, w.sdf_trunc)
    pts, pos = w.scan(0)
    g.insert(pts, pos); o.insert(pts, pos)
    keys, sd_bits, _ = o.voxels()
    g.finalize_active(); o.finalize_active()
    sd = sd_bits.view(np.float32)
    t = np.float32(1.0) / np.float32(w.sdf_trunc)
    q = np.clip(sd * t, np.float32(-1.0), np.float32(1.0)) * np.float32(127.0) + np.float32(127.0)   # cluster.hpp:19-26, fp32, truncation
    expect = q.astype(np.uint64).astype(np.uint8)
    got = g.query_voxels(0, keys)
    assert np.array_equal(got, expect)
    absent = np.setdiff1d(keys + np.uint64(1 << 30), keys)[:10000]  # same region, far-away keys
    assert np.all(g.query_voxels(0, absent) == 0xFF)
    empty_neighbours = np.setdiff1d(keys ^ np.uint64(1), keys)       # the other voxel of a pair: mostly absent leaves of PRESENT clusters
    assert np.all(g.query_voxels(0, empty_neighbours) == 0xFF)
    g.close(); o.close()


def test_scan_beyond_the_tile_run_rank_range_then_ordinary_scans(chad_lib, oracle_lib):
    """Maximum sizes: one scan of more than 2^23 points does not fit the tile-run path's 23-bit point rank and goes through the
    global-sort path by itself; the ordinary scans after it are tile-run batches again. The big batch's sorted updates wait in the
    buffers the tile-run path uses for its records, so its fold must be queued before the next batch's ray walk writes them."""
    from chad_tsdf_b200 import TSDFMap
    from oracle import bindings as ob
    n_big = (1 << 23) + 5000
    pts = synth.sphere_demo_points(n_big + 3 * 40_000)
    pos = np.zeros(3, np.float32)
    g, o = TSDFMap(0.05, 0.1, max_batch_scans=1), ob.OracleMap(0.05, 0.1)
    cuts = [0, n_big, n_big + 40_000, n_big + 80_000, n_big + 120_000]
    for a, b in zip(cuts[:-1], cuts[1:]):
        g.insert(pts[a:b], pos); o.insert(pts[a:b], pos)
    _assert_same_state(g, o, check_levels=False)
    g.finalize_active(); o.finalize_active()
    _assert_same_state(g, o)
    assert g.stats()["points"] == len(pts)
    g.close(); o.close()


@pytest.mark.parametrize("beams,want_bits", [(32, 49), (64, 50)])
def test_run_descriptor_keys_of_49_and_50_bits(chad_lib, oracle_lib, monkeypatch, beams, want_bits):
    """The widths at which a run descriptor's key reaches into the sort digit that used to hold the stashed record count: voxel coordinates
    beyond +-8192 (k = 14: 36 block bits) in ONE batch of 24 scans (5 scan bits) of 65 536 / 131 072 points (8 / 9 tile bits). Round 1's
    `nbits <= 50` let the last radix digit order the runs by six bits of the count: the runs of a block were no longer adjacent, several
    warps folded the same block, chunks were inserted twice -- found by bench.py's hash check on the 1000-scan urban drive (no shorter
    test reached these widths). 24 scans 0.1 m apart 500 m down the street: one submap, one batch."""
    from chad_tsdf_b200 import TSDFMap
    from oracle import bindings as ob
    monkeypatch.setenv("CHAD_FIRST_BATCH", "64")  # the whole burst in one batch
    w = synth.Workload("t", synth.URBAN, beams, 24, 500.0, 0.1, 0.05, 0.10, seed=31)
    g, o = TSDFMap(w.sdf_res, w.sdf_trunc, max_batch_scans=24), ob.OracleMap(w.sdf_res, w.sdf_trunc)
    for s in range(w.scans):
        pts, pos = w.scan(s)
        g.insert(pts, pos)
        o.insert(pts, pos)
    g.flush()
    st = g.stats()
    assert st["batches"] == 1 and st["key_bits_pairs"] == 45, st  # k = 14
    tiles = -(-max(len(w.scan(s)[0]) for s in range(3)) // 256)
    assert 3 * 14 - 6 + 5 + (tiles - 1).bit_length() == want_bits
    _assert_same_state(g, o, check_levels=False)
    g.finalize_active(); o.finalize_active()
    _assert_same_state(g, o)
    g.close(); o.close()
