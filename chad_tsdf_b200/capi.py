"""ctypes binding of the C ABI declared in include/chad_b200.h (the drop-in boundary).

There is no CPU fallback: if the shared library has not been built, or no CUDA device exists,
loading / creating a map raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB

CHAD_OK = 0
NUM_LEVELS = 21
LEVEL_CLUSTERS = 20
SHARD_ID_BYTES = 512


class ChadError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"chad_b200 error {code}: {message}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("scans", "points", "updates", "scan_voxels", "batches", "submaps", "kernel_launches",
                                           "h2d_bytes", "d2h_bytes", "key_bits_points", "key_bits_pairs", "resident_clusters")]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Memory(C.Structure):
    """chad_memory of include/chad_b200.h."""
    _fields_ = [(n, C.c_uint64) for n in ("dag_words_bytes", "dag_arena_bytes", "dedup_bytes", "dedup_records", "dedup_max_load_permille", "chunk_table_bytes",
                                           "batch_buffer_bytes", "grow_events", "grow_bytes", "grow_host_us", "device_used_bytes", "device_total_bytes")]

    def as_dict(self) -> dict:
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class DagImage(C.Structure):
    """chad_dag_image of include/chad_b200.h."""
    _fields_ = [("node_words", C.c_void_p * 20), ("node_word_count", C.c_size_t * 20), ("cluster_words", C.c_void_p), ("cluster_word_count", C.c_size_t),
                ("uniques", C.c_uint32 * NUM_LEVELS), ("dupes", C.c_uint32 * NUM_LEVELS), ("roots", C.c_void_p), ("n_submaps", C.c_uint32),
                ("positions", C.c_void_p), ("position_counts", C.c_void_p)]


# every symbol include/chad_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "chad_create": (C.c_int, [C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(_P)]),
    "chad_destroy": (None, [_P]),
    "chad_last_error": (C.c_char_p, [_P]),
    "chad_insert": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "chad_insert_async": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "chad_insert_many": (C.c_int, [_P, _P, _P, _P, C.c_size_t, C.c_int]),
    "chad_insert_device": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "chad_flush": (C.c_int, [_P]),
    "chad_finalize_active": (C.c_int, [_P]),
    "chad_submap_count": (C.c_int, [_P, C.POINTER(C.c_uint32)]),
    "chad_submap_roots": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "chad_voxel_count": (C.c_int, [_P, C.POINTER(C.c_size_t)]),
    "chad_export_voxels": (C.c_int, [_P, _P, _P, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "chad_level_words": (C.c_int, [_P, C.c_int, C.POINTER(C.c_size_t)]),
    "chad_level_counters": (C.c_int, [_P, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "chad_export_level": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "chad_set_pair_path": (C.c_int, [_P, C.c_int]),
    "chad_pipeline_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "chad_reset": (C.c_int, [_P]),
    "chad_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "chad_memory_info": (C.c_int, [_P, C.POINTER(Memory)]),
    "chad_reset_stats": (C.c_int, [_P]),
    "chad_stage_points": (C.c_int, [_P, _P, C.c_size_t, _P, _P, _P, _P, _P]),
    "chad_stage_pairs": (C.c_int, [_P, _P, _P, C.c_size_t, _P, _P, _P, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "chad_stage_sort": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_int]),
    "chad_stage_morton": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "chad_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "chad_host_free": (C.c_int, [_P]),
    "chad_host_register": (C.c_int, [_P, C.c_size_t]),
    "chad_host_unregister": (C.c_int, [_P]),
    "chad_device_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "chad_device_free": (C.c_int, [_P, _P]),
    "chad_upload": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "chad_timer_begin": (C.c_int, [_P]),
    "chad_timer_end": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "chad_shard_unique_id": (C.c_int, [_P]),
    "chad_create_sharded": (C.c_int, [C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, _P, C.POINTER(_P)]),
    "chad_shard_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "chad_shard_export_chunks": (C.c_int, [_P, C.POINTER(C.c_size_t), C.POINTER(_P), C.POINTER(_P)]),
    "chad_shard_finalize_from": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_int]),
    "chad_shard_clear": (C.c_int, [_P]),
    "chad_query_voxels": (C.c_int, [_P, C.c_uint32, _P, C.c_size_t, _P]),
    "chad_iterate_leaves": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "chad_submap_positions": (C.c_int, [_P, C.c_uint32, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "chad_import_dag": (C.c_int, [_P, _P]),
    "chad_profile_timeline": (C.c_int, [_P, _P, _P, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "chad_morton_encode": (C.c_uint64, [C.c_int32, C.c_int32, C.c_int32]),
    "chad_morton_decode": (None, [C.c_uint64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "chad_key_compact": (C.c_uint64, [C.c_uint64, C.c_uint]),
    "chad_key_expand": (C.c_uint64, [C.c_uint64, C.c_uint]),
    "chad_profile_enable": (C.c_int, [_P, C.c_int]),
    "chad_profile_classes": (C.c_int, []),
    "chad_profile_get": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
}

_LIB: C.CDLL | None = None


def load() -> C.CDLL:
    """Load chad_tsdf_b200/libchad_b200.so and bind every declared symbol; raises if it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB):
            raise ChadError(-2, f"{LIB} has not been built (run `python -m chad_tsdf_b200.build`); there is no CPU fallback")
        lib = C.CDLL(LIB)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _LIB = lib
    return _LIB


def ptr(a: np.ndarray | None):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
