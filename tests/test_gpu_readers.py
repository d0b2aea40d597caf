"""GPU: the read path and persistence of SURVEY.md section 8f through the C ABI -- the device leaf iterator (chad_iterate_leaves),
the submaps' poses (chad_submap_positions) and the restore of a saved map into a fresh context (chad_import_dag), after which inserts
must continue exactly as in an uninterrupted run."""
import numpy as np
import pytest

from chad_tsdf_b200 import synth

pytestmark = pytest.mark.gpu


def _quantise(sd_bits, trunc):
    sd = sd_bits.view(np.float32)
    t = np.float32(1.0) / np.float32(trunc)
    q = np.clip(sd * t, np.float32(-1.0), np.float32(1.0)) * np.float32(127.0) + np.float32(127.0)  # cluster.hpp:19-26, fp32, truncation
    return q.astype(np.uint64).astype(np.uint8)


def test_device_leaf_iterator_lists_every_voxel_of_every_submap(chad_lib, oracle_lib):
    from chad_tsdf_b200 import TSDFMap
    w = synth.Workload("t", synth.BOX_ROOM, 32, 6, -3.0, 1.3, 0.05, 0.10, seed=21)  # a submap switch at scan 4
    g, o = TSDFMap(w.sdf_res, w.sdf_trunc, max_batch_scans=3), oracle_lib.OracleMap(w.sdf_res, w.sdf_trunc)
    closed, poses = [], [[]]
    for s in range(w.scans):
        pts, pos = w.scan(s)
        before = o.voxels()
        if o.insert(pts, pos) != len(closed):
            closed.append(before)
            poses.append([])
        poses[-1].append(pos)
        g.insert(pts, pos)
    closed.append(o.voxels())
    g.finalize_active(); o.finalize_active()
    assert len(g.roots()) == len(closed) == 2
    for submap, (keys, sd_bits, _) in enumerate(closed):
        k, b = g.iterate_leaves(submap)
        assert np.array_equal(k, keys), f"submap {submap}: voxel keys"
        assert np.array_equal(b, _quantise(sd_bits, w.sdf_trunc)), f"submap {submap}: voxel bytes"
        assert np.array_equal(g.submap_positions(submap), np.array(poses[submap], np.float32))
    assert len(g.submap_positions(2)) == 0  # the fresh active submap has no pose yet
    g.close(); o.close()


def test_a_restored_map_continues_like_an_uninterrupted_one(chad_lib, oracle_lib):
    """Insert, save (= finalize + export), restore into a NEW context, insert more: every word of all 21 levels, the counters and the roots
    must equal the oracle's, which never stopped. The dedup sets rebuilt from the level arrays are what makes the later submaps share
    subtrees with the restored ones (levels.hpp:90-93: the sets are never cleared)."""
    from chad_tsdf_b200 import TSDFMap
    w = synth.Workload("t", synth.BOX_ROOM, 32, 9, -3.0, 1.3, 0.05, 0.10, seed=22)
    a, o = TSDFMap(w.sdf_res, w.sdf_trunc, max_batch_scans=2), oracle_lib.OracleMap(w.sdf_res, w.sdf_trunc)
    for s in range(5):  # two submaps, the second one still active ...
        pts, pos = w.scan(s)
        a.insert(pts, pos); o.insert(pts, pos)
    a.finalize_active(); o.finalize_active()  # ... closed by save() (tsdf.cpp:78-81)
    image = a.export_image()
    a.close()
    b = TSDFMap(w.sdf_res, w.sdf_trunc, max_batch_scans=2)
    b.import_image(image)
    assert b.roots() == o.roots()
    for s in range(5, w.scans):
        pts, pos = w.scan(s)
        b.insert(pts, pos); o.insert(pts, pos)
    for x, y in zip(b.voxels(), o.voxels()):
        assert np.array_equal(x, y)
    b.finalize_active(); o.finalize_active()
    assert b.roots() == o.roots() and len(b.roots()) >= 3
    for lv in range(21):
        ga, gu, gd = b.level(lv)
        oa, ou, od = o.level(lv)
        assert (gu, gd) == (ou, od) and np.array_equal(ga, oa), f"DAG level {lv} differs after the restore"
    k, by = b.iterate_leaves(0)  # a restored submap reads like any other
    assert len(k) > 0 and np.all(by != 0xFF)
    with pytest.raises(Exception):
        b.import_image(image)  # only an empty map can be restored into
    b.close(); o.close()
