"""Python mirror of the reference's public class chad::TSDFMap
(/root/reference/include/chad/tsdf.hpp:21-171) on top of the C ABI: same constructor arguments,
`insert(points, position)`, and the first half of `save()` (`finalize_active`); plus the state
exports the parity tests need. Everything computes on the GPU through libchad_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class TSDFMap:
    def __init__(self, sdf_res: float = 0.05, sdf_trunc: float = 0.1, device: int = 0, max_batch_scans: int = 0, pair_path: int | None = None,
                 shard: tuple[int, int, bytes] | None = None):
        """shard = (rank, world, id): this object is rank `rank` of ONE map cut into `world` Morton ranges, one GPU each (collective: every
        rank constructs it with the id rank 0 got from shard_unique_id(), and then makes the same calls with the same scans)."""
        self._lib = capi.load()
        self._h = C.c_void_p()
        self.shard_rank, self.shard_world = (shard[0], shard[1]) if shard else (0, 1)
        if shard and shard[1] > 1:
            ident = C.create_string_buffer(bytes(shard[2]), capi.SHARD_ID_BYTES)
            rc = self._lib.chad_create_sharded(sdf_res, sdf_trunc, device, max_batch_scans, shard[0], shard[1], ident, C.byref(self._h))
        else:
            rc = self._lib.chad_create(sdf_res, sdf_trunc, device, max_batch_scans, C.byref(self._h))
        if rc != capi.CHAD_OK:
            raise capi.ChadError(rc, self._lib.chad_last_error(None).decode())
        self._sdf_res, self._sdf_trunc = float(sdf_res), float(sdf_trunc)
        self.sdf_res, self.sdf_trunc = self._sdf_res, self._sdf_trunc
        if pair_path is not None:
            self.set_pair_path(pair_path)

    @staticmethod
    def shard_unique_id() -> bytes:
        """Rank 0: the id every rank of a sharded map passes to the constructor (hand it over with any transport)."""
        lib = capi.load()
        ident = C.create_string_buffer(capi.SHARD_ID_BYTES)
        rc = lib.chad_shard_unique_id(ident)
        if rc != capi.CHAD_OK:
            raise capi.ChadError(rc, lib.chad_last_error(None).decode())
        return ident.raw

    def shard_info(self) -> dict:
        r, w = C.c_int(), C.c_int()
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.chad_shard_info(self._h, C.byref(r), C.byref(w), C.byref(a), C.byref(b), C.byref(c)))
        return {"rank": r.value, "world": w.value, "sent_runs": a.value, "sent_records": b.value, "exchanges": c.value}

    # -- lifetime --
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.chad_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != capi.CHAD_OK:
            raise capi.ChadError(rc, self._lib.chad_last_error(self._h).decode())

    # -- the reference's API --
    def insert(self, points, position, wait_for_copy: bool = True) -> None:
        """TSDFMap::insert(points, position): points = (n, 3) float32 (any host memory; pinned memory is DMA'd directly).
        wait_for_copy=False (page-locked memory only, chad_insert_async): the transfer stays in flight and `points` must not change
        until flush()."""
        if hasattr(points, "data_ptr"):  # torch tensor (host, possibly pinned) -- no numpy round trip
            assert points.dtype.is_floating_point and points.element_size() == 4 and points.is_contiguous() and not points.is_cuda
            n, p = points.numel() // 3, C.c_void_p(points.data_ptr())
        else:
            points = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
            n, p = points.shape[0], capi.ptr(points)
        pos = np.ascontiguousarray(position, dtype=np.float32).reshape(3)
        f = self._lib.chad_insert if wait_for_copy else self._lib.chad_insert_async
        self._check(f(self._h, p, n, capi.ptr(pos)))

    def insert_many(self, scans, wait_for_copy: bool = True):
        """The caller's loop over TSDFMap::insert behind one C call (chad_insert_many). scans: [(points, pose)], points = host arrays /
        tensors as for insert(). Returns the prepared argument block; pass it back as `scans` to skip the marshalling next time."""
        if not isinstance(scans, tuple):
            k = len(scans)
            ptrs, counts, poses, keep = (C.c_void_p * k)(), (C.c_size_t * k)(), np.empty((k, 3), np.float32), []
            for i, (pts, pos) in enumerate(scans):
                if hasattr(pts, "data_ptr"):
                    ptrs[i], counts[i] = pts.data_ptr(), pts.numel() // 3
                else:
                    pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
                    ptrs[i], counts[i] = pts.ctypes.data, pts.shape[0]
                keep.append(pts)
                poses[i] = np.asarray(pos, np.float32).reshape(3)
            scans = (ptrs, counts, poses, k, keep)
        ptrs, counts, poses, k, _ = scans
        self._check(self._lib.chad_insert_many(self._h, ptrs, counts, capi.ptr(poses), k, 1 if wait_for_copy else 0))
        return scans

    def insert_device(self, device_ptr: int, n: int, position) -> None:
        pos = np.ascontiguousarray(position, dtype=np.float32).reshape(3)
        self._check(self._lib.chad_insert_device(self._h, C.c_void_p(device_ptr), n, capi.ptr(pos)))

    def flush(self) -> None:
        self._check(self._lib.chad_flush(self._h))

    def finalize_active(self) -> int:
        """What TSDFMap::save does before meshing (tsdf.cpp:78-81). Returns the number of finalised submaps."""
        self._check(self._lib.chad_finalize_active(self._h))
        return len(self.roots())

    # -- state export --
    def roots(self):
        n = C.c_uint32()
        self._check(self._lib.chad_submap_count(self._h, C.byref(n)))
        out = []
        for i in range(n.value):
            a, b = C.c_uint32(), C.c_uint32()
            self._check(self._lib.chad_submap_roots(self._h, i, C.byref(a), C.byref(b)))
            out.append((a.value, b.value))
        return out

    def voxels(self):
        """(keys u64 ascending, sd_bits u32, weights u32) of the active submap."""
        n = C.c_size_t()
        self._check(self._lib.chad_voxel_count(self._h, C.byref(n)))
        keys, sd, w = np.empty(n.value, np.uint64), np.empty(n.value, np.uint32), np.empty(n.value, np.uint32)
        got = C.c_size_t()
        self._check(self._lib.chad_export_voxels(self._h, capi.ptr(keys), capi.ptr(sd), capi.ptr(w), n.value, C.byref(got)))
        assert got.value == n.value
        return keys, sd, w

    def query_voxels(self, submap: int, keys) -> np.ndarray:
        """Quantised TSDF bytes (0xFF = absent) of the given Morton keys in finalised submap `submap`, read from the DAG on the device."""
        k = np.ascontiguousarray(keys, dtype=np.uint64)
        out = np.empty(k.shape[0], np.uint8)
        self._check(self._lib.chad_query_voxels(self._h, submap, capi.ptr(k), k.shape[0], capi.ptr(out)))
        return out

    def iterate_leaves(self, submap: int):
        """Every voxel of finalised submap `submap` as (Morton keys ascending, quantised bytes): the device-side leaf iterator."""
        n = C.c_size_t()
        self._check(self._lib.chad_iterate_leaves(self._h, submap, None, None, 0, C.byref(n)))
        keys, b = np.empty(n.value, np.uint64), np.empty(n.value, np.uint8)
        if n.value:
            self._check(self._lib.chad_iterate_leaves(self._h, submap, capi.ptr(keys), capi.ptr(b), n.value, C.byref(n)))
        return keys, b

    def submap_positions(self, submap: int) -> np.ndarray:
        """Submap::positions (submap.hpp:110): poses of the scans of finalised submap `submap` (== number of submaps: the active one)."""
        n = C.c_size_t()
        self._check(self._lib.chad_submap_positions(self._h, submap, None, 0, C.byref(n)))
        out = np.empty((n.value, 3), np.float32)
        if n.value:
            self._check(self._lib.chad_submap_positions(self._h, submap, capi.ptr(out), n.value, C.byref(n)))
        return out

    def export_image(self) -> dict:
        """Everything chad_import_dag needs to continue this map elsewhere: level arrays, counters, roots, poses (after finalize_active)."""
        levels = [self.level(lv) for lv in range(capi.NUM_LEVELS)]
        roots = self.roots()
        return {"levels": [a for a, _, _ in levels], "uniques": [u for _, u, _ in levels], "dupes": [d for _, _, d in levels], "roots": roots,
                "positions": [self.submap_positions(i) for i in range(len(roots))]}

    def import_image(self, image: dict) -> None:
        """chad_import_dag: restore a saved map into this (empty) map; inserts continue from there."""
        img = capi.DagImage()
        keep = []
        for lv in range(20):
            a = np.ascontiguousarray(image["levels"][lv], np.uint32)
            keep.append(a)
            img.node_words[lv] = a.ctypes.data_as(C.c_void_p)
            img.node_word_count[lv] = len(a)
        cl = np.ascontiguousarray(image["levels"][20], np.uint64)
        keep.append(cl)
        img.cluster_words, img.cluster_word_count = cl.ctypes.data_as(C.c_void_p), len(cl)
        for lv in range(capi.NUM_LEVELS):
            img.uniques[lv], img.dupes[lv] = image["uniques"][lv], image["dupes"][lv]
        roots = np.ascontiguousarray(np.array(image["roots"], np.uint32).reshape(-1, 2))
        counts = np.array([len(p) for p in image["positions"]], np.uint32)
        poses = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float32).reshape(-1, 3) for p in image["positions"]]) if len(counts) else np.zeros((0, 3), np.float32))
        keep += [roots, counts, poses]
        img.roots, img.n_submaps = roots.ctypes.data_as(C.c_void_p), len(roots)
        img.positions, img.position_counts = poses.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p)
        self._check(self._lib.chad_import_dag(self._h, C.byref(img)))

    def level(self, level: int):
        n = C.c_size_t()
        self._check(self._lib.chad_level_words(self._h, level, C.byref(n)))
        arr = np.empty(n.value, np.uint32 if level < capi.LEVEL_CLUSTERS else np.uint64)
        self._check(self._lib.chad_export_level(self._h, level, capi.ptr(arr), n.value))
        u, d = C.c_uint32(), C.c_uint32()
        self._check(self._lib.chad_level_counters(self._h, level, C.byref(u), C.byref(d)))
        return arr, u.value, d.value

    def stats(self) -> dict:
        s = capi.Stats()
        self._check(self._lib.chad_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def memory(self) -> dict:
        """Device memory behind the map and its growth so far (chad_memory_info)."""
        m = capi.Memory()
        self._check(self._lib.chad_memory_info(self._h, C.byref(m)))
        return m.as_dict()

    def set_pair_path(self, mode: int) -> None:
        """2 = tile runs + streaming fold (default), 0 = block-binned grouping of the voxel updates, 1 = global radix sort. Identical results."""
        self._check(self._lib.chad_set_pair_path(self._h, mode))

    def pipeline_info(self) -> dict:
        """Plan slots (batches in flight) and whether a batch's ray walk runs beside the next batch's point stage."""
        slots, overlapped = C.c_int(0), C.c_int(0)
        self._check(self._lib.chad_pipeline_info(self._h, C.byref(slots), C.byref(overlapped)))
        return {"plan_slots": slots.value, "walk_overlapped": bool(overlapped.value)}

    def reset(self) -> None:
        """Forget the whole map but keep the device buffers (== a freshly constructed map)."""
        self._check(self._lib.chad_reset(self._h))

    def profile_enable(self, on: bool) -> None:
        self._check(self._lib.chad_profile_enable(self._h, 1 if on else 0))

    def profile(self) -> dict:
        """{kernel class name: (milliseconds, launches)} accumulated since profile_enable(True)."""
        out = {}
        for c in range(self._lib.chad_profile_classes()):
            name, ms, n = C.c_char_p(), C.c_double(), C.c_uint64()
            self._check(self._lib.chad_profile_get(self._h, c, C.byref(name), C.byref(ms), C.byref(n)))
            if n.value:
                out[name.value.decode()] = (ms.value, int(n.value))
        return out

    def profile_timeline(self):
        """[(kernel class name, begin ms, end ms)] of the last profiled flush, device time since its first launch."""
        n = C.c_size_t()
        self._check(self._lib.chad_profile_timeline(self._h, None, None, None, 0, C.byref(n)))
        cls, t0, t1 = np.zeros(n.value, np.int32), np.zeros(n.value, np.float32), np.zeros(n.value, np.float32)
        self._check(self._lib.chad_profile_timeline(self._h, capi.ptr(cls), capi.ptr(t0), capi.ptr(t1), n.value, C.byref(n)))
        out = []
        for c, a, b in zip(cls, t0, t1):
            name = C.c_char_p()
            self._check(self._lib.chad_profile_get(self._h, int(c), C.byref(name), None, None))
            out.append((name.value.decode(), float(a), float(b)))
        return out

    def reset_stats(self) -> None:
        self._check(self._lib.chad_reset_stats(self._h))

    # -- stage entry points (kernel-by-kernel parity) --
    def stage_points(self, points, position):
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        pos = np.ascontiguousarray(position, dtype=np.float32).reshape(3)
        n = pts.shape[0]
        xyz, keys = np.empty((n, 3), np.float32), np.empty(n, np.uint64)
        order, nrm = np.empty(n, np.uint32), np.empty((n, 3), np.float32)
        self._check(self._lib.chad_stage_points(self._h, capi.ptr(pts), n, capi.ptr(pos), capi.ptr(xyz), capi.ptr(keys), capi.ptr(order), capi.ptr(nrm)))
        return xyz, keys, order, nrm

    def stage_pairs(self, xyz_sorted, normals, position):
        pts = np.ascontiguousarray(xyz_sorted, dtype=np.float32).reshape(-1, 3)
        nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
        pos = np.ascontiguousarray(position, dtype=np.float32).reshape(3)
        n = pts.shape[0]
        total = C.c_size_t()
        counts = np.empty(n, np.uint32)
        self._check(self._lib.chad_stage_pairs(self._h, capi.ptr(pts), capi.ptr(nrm), n, capi.ptr(pos), capi.ptr(counts), None, None, 0, C.byref(total)))
        keys, sd = np.empty(total.value, np.uint64), np.empty(total.value, np.float32)
        self._check(self._lib.chad_stage_pairs(self._h, capi.ptr(pts), capi.ptr(nrm), n, capi.ptr(pos), capi.ptr(counts), capi.ptr(keys), capi.ptr(sd),
                                               total.value, C.byref(total)))
        return keys, sd, counts

    def stage_sort(self, keys, values, nbits: int):
        k = np.ascontiguousarray(keys, dtype=np.uint64).copy()
        v = np.ascontiguousarray(values, dtype=np.uint32).copy()
        self._check(self._lib.chad_stage_sort(self._h, capi.ptr(k), capi.ptr(v), k.shape[0], nbits))
        return k, v

    def stage_morton(self, voxels):
        vx = np.ascontiguousarray(voxels, dtype=np.int32).reshape(-1, 3)
        keys = np.empty(vx.shape[0], np.uint64)
        self._check(self._lib.chad_stage_morton(self._h, capi.ptr(vx), vx.shape[0], capi.ptr(keys)))
        return keys

    # -- device helpers for benchmarks --
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.chad_device_alloc(self._h, nbytes, C.byref(p)))
        return p.value

    def device_free(self, p: int) -> None:
        self._check(self._lib.chad_device_free(self._h, C.c_void_p(p)))

    def upload(self, device_ptr: int, host: np.ndarray) -> None:
        host = np.ascontiguousarray(host)
        self._check(self._lib.chad_upload(self._h, C.c_void_p(device_ptr), capi.ptr(host), host.nbytes))

    def timer_begin(self) -> None:
        self._check(self._lib.chad_timer_begin(self._h))

    def timer_end(self) -> float:
        ms = C.c_float()
        self._check(self._lib.chad_timer_end(self._h, C.byref(ms)))
        return ms.value
