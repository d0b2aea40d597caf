"""CPU: oracle properties, the host-only helpers of the C ABI, and that the product library exports every
symbol include/chad_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from chad_tsdf_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_all_exported_and_bound(chad_lib):
    from chad_tsdf_b200 import capi
    header = open(os.path.join(ROOT, "include", "chad_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(chad_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(chad_lib, name), f"{name} is declared in include/chad_b200.h but not exported"
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)


def test_create_fails_loudly_without_a_gpu(chad_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from chad_tsdf_b200 import TSDFMap, ChadError
    with pytest.raises(ChadError, match="no CPU fallback"):
        TSDFMap(0.05, 0.1)


def test_morton_helpers_match_oracle(chad_lib, oracle_lib):
    rng = np.random.default_rng(0)
    pts = rng.integers(-(1 << 20), 1 << 20, size=(2000, 3))
    for x, y, z in pts.tolist() + [[0, 0, 0], [-1, -1, -1], [(1 << 20) - 1] * 3, [-(1 << 20)] * 3]:
        k = chad_lib.chad_morton_encode(x, y, z)
        assert k == oracle_lib.morton_encode(x, y, z)
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        chad_lib.chad_morton_decode(k, C.byref(a), C.byref(b), C.byref(c))
        assert (a.value, b.value, c.value) == (x, y, z) == oracle_lib.morton_decode(k)
    assert chad_lib.chad_morton_encode(0, 0, 0) == 0x7000000000000000


@pytest.mark.parametrize("k", [3, 5, 9, 11, 16, 20])
def test_compact_key_is_order_preserving_and_invertible(chad_lib, k):
    rng = np.random.default_rng(k)
    v = rng.integers(-(1 << k), 1 << k, size=(4000, 3))
    v[:4] = [[-(1 << k)] * 3, [(1 << k) - 1] * 3, [0, 0, 0], [-1, -1, -1]]
    full = np.array([chad_lib.chad_morton_encode(*map(int, p)) for p in v], dtype=np.uint64)
    comp = np.array([chad_lib.chad_key_compact(int(f), k) for f in full], dtype=np.uint64)
    assert comp.max() < (1 << (3 * k + 3))
    back = np.array([chad_lib.chad_key_expand(int(c), k) for c in comp], dtype=np.uint64)
    assert np.array_equal(back, full)
    assert np.array_equal(np.argsort(full, kind="stable"), np.argsort(comp, kind="stable"))


def test_band_properties(oracle_lib):
    """DDA coverage / weight conservation: first voxel = voxel of p - dir*trunc, last voxels near
    p + dir*trunc, consecutive voxels are face neighbours, |sd| <= trunc, counts within the analytic bound."""
    w = synth.WORKLOADS["cfg0_single_64beam"]
    pts, pos = w.scan(0)
    xyz, keys, order, nrm = oracle_lib.oracle_stage_points(pts[:20000], pos, w.sdf_res)
    pk, sd, counts = oracle_lib.oracle_stage_pairs(xyz, nrm, pos, w.sdf_res, w.sdf_trunc)
    assert counts.sum() == len(pk) and counts.min() >= 1
    assert counts.max() <= 4 + int(np.ceil(2 * np.sqrt(3) * w.sdf_trunc / w.sdf_res))
    assert np.abs(sd).max() <= np.float32(w.sdf_trunc)
    vox = np.array([oracle_lib.morton_decode(int(k)) for k in pk[:5000]])
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
    n_rays = np.searchsorted(starts, 5000) - 1
    for i in range(n_rays):
        seg = vox[int(starts[i]):int(starts[i]) + int(counts[i])]
        assert np.all(np.abs(np.diff(seg, axis=0)).sum(axis=1) == 1)


def test_quantisation_rule(oracle_lib):
    """cluster.hpp:13-27: +trunc -> 254, -trunc -> 0, 0 -> 127, truncation toward zero, 0xFF = empty."""
    t = 0.1
    sd = np.array([t, -t, 0.0, 0.05, -0.05, 1.0, -1.0, 0.0999], np.float32)
    v = oracle_lib.quantise_cluster(sd, 0xFF, t)
    got = [(v >> (8 * i)) & 0xFF for i in range(8)]
    assert got[:3] == [254, 0, 127] and got[5:7] == [254, 0]
    assert got[3] == int(np.float32(np.float32(0.05) * np.float32(1.0 / np.float32(t))) * np.float32(127) + np.float32(127))
    assert oracle_lib.quantise_cluster(sd, 0, t) == 0xFFFFFFFFFFFFFFFF
    assert oracle_lib.quantise_cluster(sd, 0b101, t) & 0xFFFFFF == (127 << 16) | (0xFF << 8) | 254


def test_submap_switch_rule(oracle_lib):
    """tsdf.cpp:51-58: strictly more than 5 m from the submap's FIRST pose; the triggering scan goes to the new submap."""
    o = oracle_lib.OracleMap(0.05, 0.1)
    p = np.array([[1.0, 1.0, 1.0]], np.float32)
    assert o.insert(p, [0, 0, 0]) == 0
    assert o.insert(p, [5.0, 0, 0]) == 0     # exactly 5 m: no switch
    assert o.insert(p, [3.0, 4.0, 0.5]) == 1  # 5.02 m from the first pose
    assert len(o.voxels()[0]) > 0             # the triggering scan is in the new submap
    assert o.insert(p, [3.0, 0.0, 0.5]) == 1  # 4 m from the NEW first pose (3, 4, 0.5): same submap
    assert o.insert(p, [0, 0, 0]) == 2        # 5.02 m from the new first pose: switch again
    o.close()


def test_workloads_have_the_named_shape():
    w = synth.WORKLOADS
    p, _ = w["cfg0_single_64beam"].scan(0)
    assert p.shape == (131072, 3) and p.dtype == np.float32
    p, _ = w["cfg1_traj100_128beam"].scan(99)
    assert p.shape == (262144, 3)
    assert w["cfg1_traj100_128beam"].scans == 100 and w["cfg4_traj1000_128beam"].scans == 1000
    assert not np.isnan(p).any()
