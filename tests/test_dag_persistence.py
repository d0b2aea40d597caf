"""CPU: the saved-map reader of the C++ facade (chad::load_dag + HostNodeLevels::query, host-only code in libchad_b200.so).
A CHADDAG1 file is written HERE from the oracle's DAG, following the format as INTEGRATION.md section 4 states it (an independent
writer), read back by the library, and every voxel of the closed submaps must come back with the byte cluster.hpp:13-26
quantises its distance to -- the same check tests/test_gpu_parity.py makes for the device-side reader (chad_query_voxels)."""
import ctypes as C
import struct
import subprocess

import numpy as np
import pytest

from chad_tsdf_b200 import synth


def _write_chaddag1(path, m, res, trunc):
    with open(path, "wb") as f:
        f.write(b"CHADDAG1")
        roots = m.roots()
        f.write(struct.pack("<ffI", res, trunc, len(roots)))
        for r in roots:
            f.write(struct.pack("<II", *r))
        for lv in range(20):
            arr, _, _ = m.level(lv)
            f.write(struct.pack("<Q", len(arr)))
            f.write(np.ascontiguousarray(arr, dtype="<u4").tobytes())
        arr, _, _ = m.level(20)
        f.write(struct.pack("<Q", len(arr)))
        f.write(np.ascontiguousarray(arr, dtype="<u8").tobytes())


def _quantise(sd_bits, trunc):
    sd = sd_bits.view(np.float32)
    t = np.float32(1.0) / np.float32(trunc)
    q = np.clip(sd * t, np.float32(-1.0), np.float32(1.0)) * np.float32(127.0) + np.float32(127.0)  # cluster.hpp:19-26, fp32, truncation
    return q.astype(np.uint64).astype(np.uint8)


def _query(exe, chad_file, submap, keys, tmp_path):
    kf, of = tmp_path / f"keys{submap}.u64", tmp_path / f"out{submap}.u8"
    np.ascontiguousarray(keys, dtype="<u8").tofile(kf)
    r = subprocess.run([exe, str(chad_file), str(submap), str(kf), str(of)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout.split(), np.fromfile(of, np.uint8)


def test_saved_map_reads_back_on_the_host(chad_lib, oracle_lib, tmp_path):
    from chad_tsdf_b200 import build
    exe = build.build_dag_reader()
    w = synth.Workload("persist", synth.BOX_ROOM, 32, 2, 0.0, 6.0, 0.05, 0.10)  # two scans 6 m apart: two submaps
    o = oracle_lib.OracleMap(w.sdf_res, w.sdf_trunc)
    per_submap = []                 # voxels of each submap as the oracle holds them right before the submap is closed
    pts, pos = w.scan(0)
    o.insert(pts, pos)
    per_submap.append(o.voxels())
    pts, pos = w.scan(1)
    o.insert(pts, pos)              # closes the first submap (tsdf.cpp:51-58) ...
    per_submap.append(o.voxels())
    o.finalize_active()             # ... and save() closes the second one (tsdf.cpp:78-81)
    roots = o.roots()
    assert len(roots) == 2
    chad_file = tmp_path / "map.chad"
    _write_chaddag1(chad_file, o, w.sdf_res, w.sdf_trunc)
    for submap, (keys, sd_bits, _) in enumerate(per_submap):
        head, got = _query(exe, chad_file, submap, keys, tmp_path)
        assert (np.float32(head[0]), np.float32(head[1]), int(head[2])) == (np.float32(w.sdf_res), np.float32(w.sdf_trunc), 2)
        assert (int(head[3]), int(head[4])) == tuple(roots[submap])
        assert np.array_equal(got, _quantise(sd_bits, w.sdf_trunc)), f"submap {submap}: voxel bytes differ"
        far = np.setdiff1d(keys + np.uint64(1 << 30), keys)[:5000]
        assert np.all(_query(exe, chad_file, submap, far, tmp_path)[1] == 0xFF)
        pair = np.setdiff1d(keys ^ np.uint64(1), keys)          # absent leaves of present clusters
        assert np.all(_query(exe, chad_file, submap, pair, tmp_path)[1] == 0xFF)
    # the other submap's voxels are a different tree: a voxel only one submap holds is absent from the other
    only0 = np.setdiff1d(per_submap[0][0], per_submap[1][0])[:5000]
    assert len(only0) and np.all(_query(exe, chad_file, 1, only0, tmp_path)[1] == 0xFF)
    o.close()


@pytest.mark.parametrize("damage", ["magic", "truncate", "trailing"])
def test_malformed_files_are_rejected(chad_lib, oracle_lib, tmp_path, damage):
    from chad_tsdf_b200 import build
    exe = build.build_dag_reader()
    o = oracle_lib.OracleMap(0.05, 0.1)
    pts = synth.sphere_demo_points(5000)
    o.insert(pts, np.zeros(3, np.float32))
    o.finalize_active()
    good = tmp_path / "good.chad"
    _write_chaddag1(good, o, 0.05, 0.1)
    o.close()
    blob = bytearray(open(good, "rb").read())
    if damage == "magic":
        blob[:8] = b"CHADDAG2"
    elif damage == "truncate":
        blob = blob[: len(blob) // 2]
    else:
        blob += b"\0\0\0\0"
    bad = tmp_path / "bad.chad"
    open(bad, "wb").write(bytes(blob))
    (tmp_path / "k.u64").write_bytes(b"")
    r = subprocess.run([exe, str(bad), "0", str(tmp_path / "k.u64"), str(tmp_path / "o.u8")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "chad::load_dag" in r.stderr


def test_grid_file_of_a_loaded_map_matches_the_oracle(chad_lib, oracle_lib, tmp_path):
    """chad::write_grid = the reference's ChadGrid constructor + saveGrid (lvr2.cpp:32-130,170-200) on a map read back with
    load_dag: one query point per voxel of the submap, at the voxel's lower corner (lvr2.cpp:81-85), with the decoded distance
    (cluster.hpp:46-50); a cell is kept only when all eight of its corner voxels exist (lvr2.cpp:115-129), its corners listed
    at the reference's offsets (lvr2.cpp:88-98)."""
    from chad_tsdf_b200 import build, capi
    exe = build.build_dag_reader()
    lib = capi.load()
    w = synth.Workload("grid", synth.BOX_ROOM, 16, 1, 0.0, 0.0, 0.05, 0.10)
    o = oracle_lib.OracleMap(w.sdf_res, w.sdf_trunc)
    # a densely sampled 1 m x 1 m wall patch 3 m in front of the sensor: a solid slab of band voxels, so complete cells exist
    rng = np.random.default_rng(5)
    pts = np.empty((20000, 3), np.float32)
    pts[:, 0] = 3.0 + rng.normal(0.0, 0.005, len(pts))
    pts[:, 1:] = rng.random((len(pts), 2)) - 0.5
    pos = np.zeros(3, np.float32)
    o.insert(pts, pos)
    keys, sd_bits, _ = o.voxels()
    o.finalize_active()
    chad_file, grid_file = tmp_path / "m.chad", tmp_path / "m.grid"
    _write_chaddag1(chad_file, o, w.sdf_res, w.sdf_trunc)
    o.close()
    r = subprocess.run([exe, "grid", str(chad_file), "0", str(grid_file)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    g = open(grid_file, "rb").read()
    hdr, nq, nc = struct.unpack_from("<fQQ", g, 0)
    assert np.float32(hdr) == np.float32(w.sdf_trunc)  # the reference stores the truncation distance here
    assert nq == len(keys) and len(g) == 20 + nq * 16 + nc * 32
    qp = np.frombuffer(g, np.float32, nq * 4, 20).reshape(nq, 4)
    cells = np.frombuffer(g, np.uint32, nc * 8, 20 + nq * 16).reshape(nc, 8)
    # query points come in the tree's traversal order = ascending Morton key = the oracle's voxel order
    xyz = np.empty((len(keys), 3), np.int32)
    for i, k in enumerate(keys):
        x, y, z = (C.c_int32(), C.c_int32(), C.c_int32())
        lib.chad_morton_decode(int(k), C.byref(x), C.byref(y), C.byref(z))
        xyz[i] = (x.value, y.value, z.value)
    assert np.array_equal(qp[:, :3], xyz.astype(np.float32) * np.float32(w.sdf_res))
    byte = _quantise(sd_bits, w.sdf_trunc)
    decoded = (byte.astype(np.float32) - np.float32(127.0)) * np.float32(1.0 / 127.0) * np.float32(w.sdf_trunc)
    assert np.array_equal(qp[:, 3], decoded)
    # complete cells, recomputed here from the voxel set: cell c has corner i at voxel c - offset[i]
    off = np.array([[0, 0, 0], [-1, 0, 0], [-1, -1, 0], [0, -1, 0], [0, 0, -1], [-1, 0, -1], [-1, -1, -1], [0, -1, -1]], np.int32)
    index = {tuple(v): i for i, v in enumerate(xyz.tolist())}
    expect = {}
    for v in index:
        for o_ in off:  # voxel v is corner i of cell v + offset[i]
            c = (v[0] + o_[0], v[1] + o_[1], v[2] + o_[2])
            if c in expect:
                continue
            corners = [index.get((c[0] - q[0], c[1] - q[1], c[2] - q[2])) for q in off]
            expect[c] = corners if None not in corners else None
    complete = sorted((lib.chad_morton_encode(*c), corners) for c, corners in expect.items() if corners is not None)
    assert nc == len(complete) and nc > 0
    assert np.array_equal(cells, np.array([c for _, c in complete], np.uint32))  # ascending Morton order of the cell


def _write_chaddag2(path, m, res, trunc, positions):
    """Independent writer of the CHADDAG2 format: roots + poses per submap, counters + words per level."""
    with open(path, "wb") as f:
        f.write(b"CHADDAG2")
        roots = m.roots()
        f.write(struct.pack("<ffI", res, trunc, len(roots)))
        for r, p in zip(roots, positions):
            f.write(struct.pack("<III", r[0], r[1], len(p)))
            f.write(np.ascontiguousarray(p, dtype="<f4").tobytes())
        for lv in range(21):
            arr, u, d = m.level(lv)
            f.write(struct.pack("<IIQ", u, d, len(arr)))
            f.write(np.ascontiguousarray(arr, dtype="<u4" if lv < 20 else "<u8").tobytes())


def _two_submaps(oracle_lib):
    w = synth.Workload("persist", synth.BOX_ROOM, 32, 3, 0.0, 3.0, 0.05, 0.10)  # scans 3 m apart: the third one opens a second submap
    o = oracle_lib.OracleMap(w.sdf_res, w.sdf_trunc)
    per_submap, poses = [], [[], []]
    for s in range(3):
        pts, pos = w.scan(s)
        before = o.voxels()
        if o.insert(pts, pos) == 1 and not per_submap:
            per_submap.append(before)
        poses[len(per_submap)].append(pos)
    per_submap.append(o.voxels())
    o.finalize_active()
    assert len(o.roots()) == 2 and [len(p) for p in poses] == [2, 1]
    return w, o, per_submap, poses


def test_leaf_cursor_walks_every_voxel_in_morton_order(chad_lib, oracle_lib, tmp_path):
    """chad::LeafCursor -- the leaf iterator the reference sketches (tsdf.hpp:120-155, tsdf.cpp:88-159) -- over a saved map: exactly the
    submap's voxels, ascending, each with the byte cluster.hpp quantises its distance to."""
    from chad_tsdf_b200 import build
    exe = build.build_dag_reader()
    w, o, per_submap, poses = _two_submaps(oracle_lib)
    chad_file = tmp_path / "m.chad"
    _write_chaddag2(chad_file, o, w.sdf_res, w.sdf_trunc, poses)
    o.close()
    for submap, (keys, sd_bits, _) in enumerate(per_submap):
        kf, bf = tmp_path / "lk.u64", tmp_path / "lb.u8"
        r = subprocess.run([exe, "leaves", str(chad_file), str(submap), str(kf), str(bf)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        n, _, npos = r.stdout.split()
        assert int(n) == len(keys) and int(npos) == len(poses[submap])
        assert np.array_equal(np.fromfile(kf, np.uint64), keys)
        assert np.array_equal(np.fromfile(bf, np.uint8), _quantise(sd_bits, w.sdf_trunc))


def test_chaddag2_survives_a_load_and_save(chad_lib, oracle_lib, tmp_path):
    """save_dag(load_dag(file)) reproduces the file byte for byte: poses (submap.hpp:110), dedup counters (levels.hpp:90-91,141) and all."""
    from chad_tsdf_b200 import build
    exe = build.build_dag_reader()
    w, o, _, poses = _two_submaps(oracle_lib)
    a, b = tmp_path / "a.chad", tmp_path / "b.chad"
    _write_chaddag2(a, o, w.sdf_res, w.sdf_trunc, poses)
    o.close()
    r = subprocess.run([exe, "resave", str(a), str(b)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.split() == ["1", "2"], r.stdout + r.stderr
    assert open(a, "rb").read() == open(b, "rb").read()


def test_a_child_address_outside_its_level_is_rejected(chad_lib, oracle_lib, tmp_path):
    """A crafted / corrupt file must not become an out-of-bounds read in the readers: load_dag validates every address once."""
    from chad_tsdf_b200 import build
    exe = build.build_dag_reader()
    w, o, _, poses = _two_submaps(oracle_lib)
    good = tmp_path / "good.chad"
    _write_chaddag2(good, o, w.sdf_res, w.sdf_trunc, poses)
    blob = bytearray(open(good, "rb").read())
    # the first child address of the first record of level 19 (word 2 of the level) -> far beyond the cluster level
    at = 8 + 12 + sum(12 + 12 * len(p) for p in poses)
    for lv in range(19):
        (n,) = struct.unpack_from("<Q", blob, at + 8)
        at += 16 + 4 * n
    struct.pack_into("<I", blob, at + 16 + 4 * 2, 0x7FFFFFF0)
    o.close()
    bad = tmp_path / "bad.chad"
    open(bad, "wb").write(bytes(blob))
    (tmp_path / "k.u64").write_bytes(b"")
    r = subprocess.run([exe, str(bad), "0", str(tmp_path / "k.u64"), str(tmp_path / "o.u8")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "inconsistent DAG" in r.stderr, r.stderr


def _ray_walk_here(voxels, origin, direction, max_distance, res, trunc):
    """Independent restatement of chad::raycast: voxel-by-voxel walk with a dictionary instead of the tree (no octant skipping)."""
    import math
    o = [float(np.float32(c)) for c in origin]
    dr = [float(np.float32(c)) for c in direction]
    length = math.sqrt(dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2])
    d = [c / length for c in dr]
    res, limit = float(np.float32(res)), float(np.float32(max_distance))
    v = [math.floor(o[a] / res) for a in range(3)]
    step = [1 if d[a] > 0 else (-1 if d[a] < 0 else 0) for a in range(3)]
    t_step = [res / abs(d[a]) if step[a] else math.inf for a in range(3)]
    t_next = [((v[a] + (1 if step[a] > 0 else 0)) * res - o[a]) / d[a] if step[a] else math.inf for a in range(3)]
    walked, along, prev, t_enter = 0, [], None, 0.0
    while t_enter < limit:
        if any(c < -(1 << 20) or c >= (1 << 20) for c in v):
            break
        axis = (0 if t_next[0] < t_next[2] else 2) if t_next[0] < t_next[1] else (1 if t_next[1] < t_next[2] else 2)
        walked += 1
        byte = voxels.get(tuple(v))
        if byte is not None:
            along.append(tuple(v))
            sd = float((np.float32(byte) - np.float32(127.0)) * np.float32(1.0 / 127.0) * np.float32(trunc))  # cluster.hpp:46-50
            t = (v[0] * res - o[0]) * d[0] + (v[1] * res - o[1]) * d[1] + (v[2] * res - o[2]) * d[2]
            if prev is not None and prev[1] > 0.0 and sd <= 0.0:
                t_hit = max(0.0, prev[0] + (t - prev[0]) * (prev[1] / (prev[1] - sd)))
                return (t_hit <= limit), t_hit, walked, along
            prev = (t, sd)
        t_enter = t_next[axis]
        t_next[axis] += t_step[axis]
        v[axis] += step[axis]
    return False, 0.0, walked, along


def test_ray_cast_through_a_saved_map(chad_lib, oracle_lib, tmp_path):
    """chad::raycast -- "raycast to retrieve leaves along it + physics hit", the reader the reference lists after the leaf iterator
    (tsdf.hpp:157-160) -- on a map read back with load_dag: the voxels along each ray and the hit must equal an independent walk
    over the oracle's voxel set made here; rays from the sensor towards the measured points hit the surface where it was measured;
    empty octants are crossed without descending into them."""
    from chad_tsdf_b200 import build, capi
    exe = build.build_dag_reader()
    lib = capi.load()
    w = synth.Workload("ray", synth.BOX_ROOM, 32, 1, 0.0, 0.0, 0.05, 0.10)
    pts, pos = w.scan(0)
    o = oracle_lib.OracleMap(w.sdf_res, w.sdf_trunc)
    o.insert(pts, pos)
    keys, sd_bits, _ = o.voxels()
    o.finalize_active()
    chad_file = tmp_path / "m.chad"
    _write_chaddag1(chad_file, o, w.sdf_res, w.sdf_trunc)
    o.close()
    byte = _quantise(sd_bits, w.sdf_trunc)
    voxels = {}
    for k, b in zip(keys.tolist(), byte.tolist()):
        x, y, z = (C.c_int32(), C.c_int32(), C.c_int32())
        lib.chad_morton_decode(int(k), C.byref(x), C.byref(y), C.byref(z))
        voxels[(x.value, y.value, z.value)] = b
    rng = np.random.default_rng(11)
    pick = rng.choice(len(pts), 120, replace=False)
    rays = []
    for i in pick:  # towards measured points, from the sensor: must hit close to the measured range
        rays.append([*pos, *(pts[i] - pos), 150.0])
    for _ in range(40):  # arbitrary rays from arbitrary places, axis-parallel ones included
        d = rng.normal(size=3)
        if rng.random() < 0.3:
            d[rng.integers(3)] = 0.0
        if rng.random() < 0.2:
            d = np.eye(3)[rng.integers(3)] * rng.choice([-1.0, 1.0])
        rays.append([*rng.uniform(-15, 15, 3), *d, float(rng.uniform(5, 60))])
    rays.append([0.0, 0.0, 500.0, 0.0, 0.0, 1.0, 50.0])  # far from everything: a handful of descents for 1000 voxels
    rays = np.asarray(rays, np.float32)
    rf, of, kf = tmp_path / "rays.f32", tmp_path / "hits.f64", tmp_path / "along.u64"
    rays.tofile(rf)
    r = subprocess.run([exe, "ray", str(chad_file), "0", str(rf), str(of), str(kf)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(of, np.float64).reshape(len(rays), 9)
    along_keys = np.fromfile(kf, np.uint64)
    at = 0
    hits_towards_points = 0
    for i, ray in enumerate(rays):
        hit, t_hit, walked, along = _ray_walk_here(voxels, ray[:3], ray[3:6], ray[6], w.sdf_res, w.sdf_trunc)
        n_along = int(got[i, 8])
        assert (bool(got[i, 0]), int(got[i, 5]), int(got[i, 6]), n_along) == (hit, walked, len(along), len(along)), (i, got[i], hit, walked, len(along))
        expect_keys = [lib.chad_morton_encode(*v) for v in along]
        assert along_keys[at:at + n_along].tolist() == expect_keys
        at += n_along
        assert got[i, 7] <= got[i, 5]  # never more descents than voxels
        if hit:
            assert abs(got[i, 1] - t_hit) <= 1e-6 * max(1.0, t_hit)  # (float32 in the struct)
            d = ray[3:6].astype(np.float64) / np.linalg.norm(ray[3:6].astype(np.float64))
            assert np.allclose(got[i, 2:5], ray[:3] + t_hit * d, atol=1e-4)
        if i < len(pick):
            measured = float(np.linalg.norm((pts[pick[i]] - pos).astype(np.float64)))
            if hit and abs(t_hit - measured) < 2.0 * w.sdf_res:
                hits_towards_points += 1
    assert at == len(along_keys)
    assert hits_towards_points >= 0.9 * len(pick), hits_towards_points  # the surface is found where the scan measured it (to two voxels)
    far = got[-1]
    assert not far[0] and far[5] >= 999 and far[6] == 0 and far[7] <= 8, far  # 1000 voxels of empty space: a few descents
