"""TEST INFRASTRUCTURE ONLY. ctypes bindings for

* oracle/_ref/libchad_ref_{verbatim,stable}.so -- the reference's own sources (built by
  oracle/Makefile from /root/reference, see oracle/ref_capi.cpp), class `RefMap`;
* oracle/_build/liboracle.so -- the C restatement (oracle/chad_oracle.c), class `OracleMap`.

Both expose the same Python surface so parity tests can swap them.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(target: str = "all") -> None:
    """Run oracle/Makefile (the ref target is a no-op when /root/reference is absent)."""
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def ref_available(variant: str = "stable") -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", f"libchad_ref_{variant}.so"))


_REF_LIBS: dict[str, C.CDLL] = {}


def _ref_lib(variant: str) -> C.CDLL:
    if variant not in _REF_LIBS:
        lib = C.CDLL(os.path.join(HERE, "_ref", f"libchad_ref_{variant}.so"))
        lib.chadref_variant.restype = C.c_char_p
        lib.chadref_create.restype = C.c_void_p
        lib.chadref_create.argtypes = [C.c_float, C.c_float]
        lib.chadref_destroy.argtypes = [C.c_void_p]
        lib.chadref_insert.restype = C.c_uint32
        lib.chadref_insert.argtypes = [C.c_void_p, _f32p, C.c_size_t, _f32p]
        lib.chadref_finalize_active.restype = C.c_uint32
        lib.chadref_finalize_active.argtypes = [C.c_void_p]
        lib.chadref_voxel_count.restype = C.c_size_t
        lib.chadref_voxel_count.argtypes = [C.c_void_p]
        lib.chadref_export_voxels.restype = C.c_size_t
        lib.chadref_export_voxels.argtypes = [C.c_void_p, _u64p, _u32p, _u32p]
        lib.chadref_submap_count.restype = C.c_uint32
        lib.chadref_submap_count.argtypes = [C.c_void_p]
        lib.chadref_submap_roots.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.chadref_level_words.restype = C.c_size_t
        lib.chadref_level_words.argtypes = [C.c_void_p, C.c_int]
        lib.chadref_level_counters.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.chadref_export_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.chadref_stage_points.argtypes = [_f32p, C.c_size_t, _f32p, C.c_float, _f32p, _u64p, _f32p]
        lib.chadref_phase_count.restype = C.c_int
        lib.chadref_phase_tag.restype = C.c_char_p
        lib.chadref_phase_tag.argtypes = [C.c_int]
        lib.chadref_phase_sum_ms.restype = C.c_double
        lib.chadref_phase_sum_ms.argtypes = [C.c_int]
        assert lib.chadref_variant().decode() == variant
        _REF_LIBS[variant] = lib
    return _REF_LIBS[variant]


class _MapBase:
    """Shared Python surface of RefMap / OracleMap (prefix = C symbol prefix)."""
    _prefix = ""

    def __init__(self, lib, sdf_res: float, sdf_trunc: float):
        self._lib = lib
        self._f = lambda name: getattr(lib, self._prefix + name)
        self._h = C.c_void_p(self._f("create")(sdf_res, sdf_trunc))
        self.sdf_res, self.sdf_trunc = sdf_res, sdf_trunc

    def close(self):
        if self._h:
            self._f("destroy")(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def insert(self, points: np.ndarray, pos) -> int:
        pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
        p = np.ascontiguousarray(pos, dtype=np.float32).reshape(3)
        return int(self._f("insert")(self._h, pts, pts.shape[0], p))

    def finalize_active(self) -> int:
        return int(self._f("finalize_active")(self._h))

    def voxels(self):
        """(keys u64 ascending, sd_bits u32, weights u32) of the active submap's working set."""
        n = int(self._f("voxel_count")(self._h))
        keys, sd, w = np.empty(n, np.uint64), np.empty(n, np.uint32), np.empty(n, np.uint32)
        got = int(self._f("export_voxels")(self._h, keys, sd, w))
        assert got == n
        return keys, sd, w

    def roots(self):
        out = []
        for i in range(int(self._f("submap_count")(self._h))):
            a, b = C.c_uint32(), C.c_uint32()
            self._f("submap_roots")(self._h, i, C.byref(a), C.byref(b))
            out.append((a.value, b.value))
        return out

    def level(self, level: int):
        """(words, uniques, dupes): u32 words for node levels 0..19, u64 for level 20 (leaf clusters)."""
        n = int(self._f("level_words")(self._h, level))
        arr = np.empty(n, np.uint32 if level < 20 else np.uint64)
        self._f("export_level")(self._h, level, arr.ctypes.data_as(C.c_void_p))
        u, d = C.c_uint32(), C.c_uint32()
        self._f("level_counters")(self._h, level, C.byref(u), C.byref(d))
        return arr, u.value, d.value


class RefMap(_MapBase):
    """The reference itself (chad::TSDFMap compiled from /root/reference)."""
    _prefix = "chadref_"

    def __init__(self, sdf_res: float = 0.05, sdf_trunc: float = 0.1, variant: str = "stable"):
        super().__init__(_ref_lib(variant), sdf_res, sdf_trunc)
        self.variant = variant

    def phases_ms(self) -> dict[str, float]:
        lib = self._lib
        return {lib.chadref_phase_tag(i).decode(): lib.chadref_phase_sum_ms(i) for i in range(lib.chadref_phase_count())}

    def phases_reset(self) -> None:
        self._lib.chadref_phase_reset()


def ref_stage_points(points: np.ndarray, pos, sdf_res: float, variant: str = "stable"):
    """Reference point stage: (sorted xyz f32 (n,3), keys u64 (n,), normals f32 (n,3))."""
    lib = _ref_lib(variant)
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    n = pts.shape[0]
    xyz, keys, nrm = np.empty((n, 3), np.float32), np.empty(n, np.uint64), np.empty((n, 3), np.float32)
    lib.chadref_stage_points(pts, n, np.ascontiguousarray(pos, dtype=np.float32), sdf_res, xyz, keys, nrm)
    return xyz, keys, nrm


_ORACLE_LIB: C.CDLL | None = None


def _oracle_lib() -> C.CDLL:
    global _ORACLE_LIB
    if _ORACLE_LIB is None:
        path = os.path.join(HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build("oracle")
        lib = C.CDLL(path)
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_create.argtypes = [C.c_float, C.c_float]
        lib.oracle_destroy.argtypes = [C.c_void_p]
        lib.oracle_insert.restype = C.c_uint32
        lib.oracle_insert.argtypes = [C.c_void_p, _f32p, C.c_size_t, _f32p]
        lib.oracle_finalize_active.restype = C.c_uint32
        lib.oracle_finalize_active.argtypes = [C.c_void_p]
        lib.oracle_voxel_count.restype = C.c_size_t
        lib.oracle_voxel_count.argtypes = [C.c_void_p]
        lib.oracle_export_voxels.restype = C.c_size_t
        lib.oracle_export_voxels.argtypes = [C.c_void_p, _u64p, _u32p, _u32p]
        lib.oracle_submap_count.restype = C.c_uint32
        lib.oracle_submap_count.argtypes = [C.c_void_p]
        lib.oracle_submap_roots.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.oracle_level_words.restype = C.c_size_t
        lib.oracle_level_words.argtypes = [C.c_void_p, C.c_int]
        lib.oracle_level_counters.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.oracle_export_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_last_scan_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.oracle_stage_points.argtypes = [_f32p, C.c_size_t, _f32p, C.c_float, _f32p, _u64p, _u32p, C.c_void_p]
        lib.oracle_stage_pairs.restype = C.c_size_t
        lib.oracle_stage_pairs.argtypes = [_f32p, _f32p, C.c_size_t, _f32p, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_morton_encode.restype = C.c_uint64
        lib.oracle_morton_encode.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        lib.oracle_morton_decode.argtypes = [C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        lib.oracle_quantise_cluster.restype = C.c_uint64
        lib.oracle_quantise_cluster.argtypes = [_f32p, C.c_uint32, C.c_float]
        _ORACLE_LIB = lib
    return _ORACLE_LIB


class OracleMap(_MapBase):
    """The C restatement (oracle/chad_oracle.c)."""
    _prefix = "oracle_"

    def __init__(self, sdf_res: float = 0.05, sdf_trunc: float = 0.1):
        super().__init__(_oracle_lib(), sdf_res, sdf_trunc)

    def last_scan_stats(self) -> tuple[int, int]:
        u, v = C.c_uint64(), C.c_uint64()
        self._lib.oracle_last_scan_stats(self._h, C.byref(u), C.byref(v))
        return u.value, v.value


def oracle_stage_points(points: np.ndarray, pos, sdf_res: float):
    """Restatement point stage: (sorted xyz (n,3), keys u64, order u32 (input index of sorted i), normals (n,3))."""
    lib = _oracle_lib()
    pts = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, 3)
    n = pts.shape[0]
    xyz, keys = np.empty((n, 3), np.float32), np.empty(n, np.uint64)
    order, nrm = np.empty(n, np.uint32), np.empty((n, 3), np.float32)
    lib.oracle_stage_points(pts, n, np.ascontiguousarray(pos, dtype=np.float32), sdf_res, xyz, keys, order, nrm.ctypes.data_as(C.c_void_p))
    return xyz, keys, order, nrm


def oracle_stage_pairs(xyz_sorted: np.ndarray, normals: np.ndarray, pos, sdf_res: float, sdf_trunc: float):
    """Restatement band enumeration: (keys u64 (U,), sd f32 (U,), counts u32 (n,)) in (point, ray step) order."""
    lib = _oracle_lib()
    pts = np.ascontiguousarray(xyz_sorted, dtype=np.float32).reshape(-1, 3)
    nrm = np.ascontiguousarray(normals, dtype=np.float32).reshape(-1, 3)
    p = np.ascontiguousarray(pos, dtype=np.float32)
    n = pts.shape[0]
    total = int(lib.oracle_stage_pairs(pts, nrm, n, p, sdf_res, sdf_trunc, None, None, None))
    keys, sd, counts = np.empty(total, np.uint64), np.empty(total, np.float32), np.empty(n, np.uint32)
    lib.oracle_stage_pairs(pts, nrm, n, p, sdf_res, sdf_trunc, keys.ctypes.data_as(C.c_void_p), sd.ctypes.data_as(C.c_void_p),
                           counts.ctypes.data_as(C.c_void_p))
    return keys, sd, counts


def morton_encode(x: int, y: int, z: int) -> int:
    return int(_oracle_lib().oracle_morton_encode(x, y, z))


def morton_decode(key: int) -> tuple[int, int, int]:
    x, y, z = C.c_int32(), C.c_int32(), C.c_int32()
    _oracle_lib().oracle_morton_decode(key, C.byref(x), C.byref(y), C.byref(z))
    return x.value, y.value, z.value


def quantise_cluster(sd8: np.ndarray, present_mask: int, sdf_trunc: float) -> int:
    return int(_oracle_lib().oracle_quantise_cluster(np.ascontiguousarray(sd8, dtype=np.float32), present_mask, sdf_trunc))


def map_digest(m: _MapBase) -> dict:
    """sha256 digests of everything parity compares: active voxels, all 21 levels, counters, roots."""
    import hashlib

    def h(*arrays):
        d = hashlib.sha256()
        for a in arrays:
            d.update(np.ascontiguousarray(a).tobytes())
        return d.hexdigest()

    k, sd, w = m.voxels()
    out = {"voxels_n": int(len(k)), "voxels_keys": h(k), "voxels_sd_bits": h(sd), "voxels_weights": h(w), "weight_sum": int(w.astype(np.uint64).sum()),
           "roots": [list(r) for r in m.roots()], "levels": []}
    for lv in range(21):
        arr, u, d = m.level(lv)
        out["levels"].append({"words": int(len(arr)), "uniques": int(u), "dupes": int(d), "sha256": h(arr)})
    return out
