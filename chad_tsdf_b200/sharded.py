"""ONE TSDF map on several GPUs, one process per GPU (torch.distributed is plumbing only).

* Morton-range sharding (north_star): the map is cut into `world` contiguous Morton ranges and the C++ library
  does everything, including the NCCL exchanges (chad_create_sharded, csrc/shard.cu, csrc/runs.cu, csrc/context.cu). This file only
  hands the communicator id around (`create_sharded_map`) and gathers the shards' state for parity checks (`sharded_digest`).
* Submap-parallel mode (`SubmapParallelTSDFMap`): the map's submaps -- independent octrees -- are integrated on different ranks and
  only their sorted leaf chunks travel; kept as the second decomposition.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch
import torch.distributed as dist

CELL_WORDS = 8   # a leaf chunk's cells: 8 x (sd bits, weight) = 8 x int64


class _DevView:
    """Zero-copy view of raw device memory for torch (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def device_tensor(ptr: int, n_words: int, device) -> torch.Tensor:
    if n_words == 0:
        return torch.empty((0,), dtype=torch.int64, device=device)
    return torch.as_tensor(_DevView(ptr, n_words * 8), device=device).view(torch.int64)


class CudaShardEngine:
    """The per-rank numerical engine on a GPU: chad_shard_* of include/chad_b200.h."""

    def __init__(self, sdf_res: float, sdf_trunc: float, device: int, max_batch_scans: int = 1):
        from .tsdf_map import TSDFMap
        self.map = TSDFMap(sdf_res, sdf_trunc, device=device, max_batch_scans=max_batch_scans)
        self.device = torch.device("cuda", device)
        self._lib, self._h = self.map._lib, self.map._h

    def concat(self, scans: list) -> torch.Tensor:
        """The batch's points as one device tensor. A scan may be a numpy array (staged synchronously), a page-locked CPU
        tensor (asynchronous DMA) or a tensor that already lives on this device."""
        parts = []
        for p in scans:
            t = p if isinstance(p, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(p, np.float32))
            parts.append(t.reshape(-1, 3).to(self.device, dtype=torch.float32, non_blocking=True))
        xyz = parts[0] if len(parts) == 1 else torch.cat(parts)
        torch.cuda.current_stream(self.device).synchronize()  # the engine works on its own stream
        return xyz.contiguous()

    # the ordinary single-GPU insert path (SubmapParallelTSDFMap): host memory, or a tensor on this device
    def insert(self, points, position) -> None:
        if isinstance(points, torch.Tensor) and points.is_cuda:
            self.map.insert_device(points.data_ptr(), points.numel() // 3, position)
        else:
            self.map.insert(points, position)

    def flush(self) -> None:
        self.map.flush()

    def export_chunks(self):
        n, k, c = C.c_size_t(), C.c_void_p(), C.c_void_p()
        self.map._check(self._lib.chad_shard_export_chunks(self._h, C.byref(n), C.byref(k), C.byref(c)))
        keys = device_tensor(k.value or 0, n.value, self.device)
        cells = device_tensor(c.value or 0, n.value * CELL_WORDS, self.device).view(-1, CELL_WORDS)
        return keys, cells

    def finalize_from(self, keys: torch.Tensor, cells: torch.Tensor, clear_local: bool = True) -> None:
        keys, cells = keys.contiguous(), cells.contiguous()
        self._keep_fin = (keys, cells)  # the engine copies out of them asynchronously
        self.map._check(self._lib.chad_shard_finalize_from(self._h, C.c_void_p(keys.data_ptr()), C.c_void_p(cells.data_ptr()), keys.shape[0],
                                                           1 if clear_local else 0))

    def clear(self) -> None:
        self.map._check(self._lib.chad_shard_clear(self._h))

    def empty(self, shape, dtype=torch.int64) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype, device=self.device)

    def sync(self) -> None:
        """The collectives run on torch's stream, the engine on its own: order them."""
        torch.cuda.current_stream(self.device).synchronize()

    # state export: this rank's shard / the (replicated) DAG
    def voxels(self):
        return self.map.voxels()

    def roots(self):
        return self.map.roots()

    def level(self, lv):
        return self.map.level(lv)

    def stats(self):
        return self.map.stats()

    def close(self):
        self.map.close()


def create_sharded_map(sdf_res: float, sdf_trunc: float, device: int, group=None, max_batch_scans: int = 0):
    """Rank `dist.get_rank(group)` of ONE chad::TSDFMap cut into `world` Morton ranges (chad_create_sharded, include/chad_b200.h).
    torch.distributed only carries the 512-byte communicator id from rank 0 to the others; every exchange of the map itself is issued
    by the C++ library (NCCL send / recv on its own streams)."""
    from .tsdf_map import TSDFMap
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [TSDFMap.shard_unique_id() if rank == 0 and world > 1 else None]
    if world > 1:
        src = 0 if group is None else dist.get_global_rank(group, 0)
        dist.broadcast_object_list(box, src=src, group=group)
    return TSDFMap(sdf_res, sdf_trunc, device=device, max_batch_scans=max_batch_scans, shard=(rank, world, box[0]) if world > 1 else None)


def gather_voxels(local, group=None):
    """(keys, sd_bits, weights) of every rank, concatenated in rank order on rank 0 (None elsewhere). The ranges ascend with the rank,
    so the result is the whole active submap in ascending Morton order -- asserted here, because it is what makes the shards ONE map."""
    if not dist.is_initialized():  # a single process holds the whole map
        parts, rank = [tuple(np.ascontiguousarray(a) for a in local)], 0
    else:
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        parts = [None] * world if rank == 0 else None
        dst = 0 if group is None else dist.get_global_rank(group, 0)
        dist.gather_object(tuple(np.ascontiguousarray(a) for a in local), parts, dst=dst, group=group)
    if rank != 0:
        return None
    keys = np.concatenate([p[0] for p in parts])
    if len(keys) > 1 and not np.all(keys[1:] > keys[:-1]):
        raise AssertionError("the ranks' voxel ranges overlap or are out of order: the shards do not form one map")
    return keys, np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts])


def sharded_digest(m, group=None, with_dag: bool = True):
    """The digest of oracle.bindings.map_digest (active voxels, the 21 DAG levels, counters, roots) of a sharded map, on rank 0 (None
    elsewhere): voxels gathered from all ranks, the DAG read from rank 0, which holds it. Collective."""
    import hashlib

    def h(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()

    vox = gather_voxels(m.voxels(), group)
    roots = m.roots()
    if vox is None:
        return None
    k, sd, w = vox
    out = {"voxels_n": int(len(k)), "voxels_keys": h(k), "voxels_sd_bits": h(sd), "voxels_weights": h(w), "weight_sum": int(w.astype(np.uint64).sum()),
           "roots": [list(r) for r in roots], "levels": []}
    if with_dag:
        for lv in range(21):
            arr, u, d = m.level(lv)
            out["levels"].append({"words": int(len(arr)), "uniques": int(u), "dupes": int(d), "sha256": h(arr)})
    return out


class SubmapParallelTSDFMap:
    """chad::TSDFMap semantics with the map's SUBMAPS integrated on different GPUs.

    A submap is an independent octree: TSDFMap::insert clears it at every switch (tsdf.cpp:51-58) and only
    Submap::finalize couples it to the rest of the map, through the global DAG, in submap order (submap.hpp:10-106,
    levels.hpp). So submap s is integrated by rank s mod world alone, through the ordinary single-GPU insert path and
    without any exchange; when it closes, its owner broadcasts the submap's sorted leaf chunks (NCCL over NVLink) and
    every rank folds them into its replica of the DAG, in submap order. The result -- every submap's voxels and the
    whole DAG -- is bit-identical to the single-GPU map. The submap rule only needs the poses, so every rank evaluates
    it for every scan and simply skips the scans of the submaps it does not own (their points are never copied).
    """

    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self._first_pose: np.ndarray | None = None
        self._submap = 0          # index of the active submap
        self._closed = 0          # submaps already folded into the DAG (closes are collective and lag world - 1 submaps behind)
        self.owned_scans = 0
        self.broadcast_chunks = 0

    def owner(self, submap: int) -> int:
        return submap % self.world

    def insert(self, points, position) -> None:
        pos = np.ascontiguousarray(position, np.float32).reshape(3)
        if self._first_pose is None:
            self._first_pose = pos.copy()
        else:  # tsdf.cpp:46-61: strictly more than 5 m from the submap's first pose, fp32 ((dx*dx + dy*dy) + dz*dz, then sqrt)
            d = self._first_pose - pos
            t = d * d
            if np.sqrt(np.float32(np.float32(t[0] + t[1]) + t[2])) > np.float32(5.0):
                self._submap += 1
                self._first_pose = pos.copy()
                # Closing a submap is collective (its owner broadcasts): do it `world - 1` switches late, so that every rank has
                # queued the inserts of its own submap of the round first and the ranks integrate concurrently. The owner of the
                # submap entered now closes its previous one right here, before it reuses its table.
                self._close_until(self._submap - (self.world - 1))
        if self.owner(self._submap) == self.rank:
            self.engine.insert(points, pos)
            self.owned_scans += 1

    def flush(self) -> None:
        """Everything inserted so far is applied and every submap before the active one is in the DAG (collective)."""
        self._close_until(self._submap)
        self.engine.flush()

    def finalize_active(self) -> None:
        """The part of TSDFMap::save before meshing (tsdf.cpp:78-81)."""
        if self._first_pose is not None:
            self._submap += 1
            self._close_until(self._submap)
            self._first_pose = None

    def _close_until(self, end: int) -> None:
        while self._closed < end:
            self._close_submap(self._closed)
            self._closed += 1

    def _close_submap(self, submap: int) -> None:
        src = self.owner(submap)
        n = self.engine.empty((1,))
        if src == self.rank:
            keys, cells = self.engine.export_chunks()
            n.fill_(keys.shape[0])
        if self.world > 1:
            src_global = src if self.group is None else dist.get_global_rank(self.group, src)
            dist.broadcast(n, src_global, group=self.group)
            c = int(n.item())
            if src != self.rank:
                keys, cells = self.engine.empty((c,)), self.engine.empty((c, CELL_WORDS))
            if c:
                dist.broadcast(keys, src_global, group=self.group)
                dist.broadcast(cells, src_global, group=self.group)
            self.engine.sync()
        self.broadcast_chunks += int(keys.shape[0])
        self.engine.finalize_from(keys, cells, clear_local=False)
        if src == self.rank:
            self.engine.clear()  # octree.clear(), tsdf.cpp:57: the owner's table is free for its next submap
