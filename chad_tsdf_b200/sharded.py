"""One TSDF map sharded across GPUs by contiguous Morton ranges.

One process per GPU (torch.distributed, NCCL over NVLink). Every rank receives every scan (3 MB), runs the small point
stage on the whole batch, enumerates the band voxels of its 1/world slice of the sorted rays and sends each update to
the rank that owns the voxel (`all_to_all_single` of 16-byte tuples {Morton key, sorted-point rank, sd}). The receiver
bins / sorts / folds the tuples into its shard; the rank carried by every tuple restores the reference's fold order, so
the union of the shards is bit-identical to the single-GPU map. At a submap switch the shards' leaf chunks are
all-gathered (rank order == Morton order) and every rank builds the identical global DAG.

The numerical work is behind a small engine interface: `CudaShardEngine` drives the C ABI (chad_shard_*); the CPU tests
plug in a numpy/oracle engine to exercise this file's host logic (submap rule, batching, splits, exchange) with gloo.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch
import torch.distributed as dist

TUPLE_WORDS = 2  # a tuple is 2 x int64 on the wire: (key, rank | sd << 32)
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

    def front(self, xyz: torch.Tensor, offsets: np.ndarray, poses: np.ndarray, rank: int, world: int, new_submap: bool):
        counts = np.zeros(8, np.uint64)
        offsets = np.ascontiguousarray(offsets, np.uint32)
        poses = np.ascontiguousarray(poses, np.float32)
        self._keep = xyz  # until the next batch: the engine copies out of it asynchronously
        self.map._check(self._lib.chad_shard_front(self._h, C.c_void_p(xyz.data_ptr()), offsets.ctypes.data_as(C.c_void_p),
                                                   poses.ctypes.data_as(C.c_void_p), len(offsets) - 1, rank, world, int(new_submap),
                                                   counts.ctypes.data_as(C.c_void_p)))
        p = C.c_void_p()
        self.map._check(self._lib.chad_shard_send_buffer(self._h, C.byref(p)))
        counts = [int(c) for c in counts[:world]]
        send = device_tensor(p.value or 0, sum(counts) * TUPLE_WORDS, self.device).view(-1, TUPLE_WORDS)
        return counts, send

    # the ordinary single-GPU insert path (SubmapParallelTSDFMap): host memory, or a tensor on this device
    def insert(self, points, position) -> None:
        if isinstance(points, torch.Tensor) and points.is_cuda:
            self.map.insert_device(points.data_ptr(), points.numel() // 3, position)
        else:
            self.map.insert(points, position)

    def flush(self) -> None:
        self.map.flush()

    def ingest(self, tuples: torch.Tensor) -> None:
        tuples = tuples.contiguous()
        self.map._check(self._lib.chad_shard_ingest(self._h, C.c_void_p(tuples.data_ptr()), tuples.shape[0]))

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


class ShardedTSDFMap:
    """chad::TSDFMap semantics (insert / submap rule / finalize) over `world` Morton-range shards."""

    def __init__(self, engine, group=None, max_batch_scans: int = 16):
        self.engine = engine
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.max_batch = max(1, min(int(max_batch_scans), 64))
        self._scans: list = []   # (points, pose) of the batch being assembled
        self._first_pose: np.ndarray | None = None
        self._new_submap = True
        self.exchanged_tuples = 0
        self.phase_s = {"stage": 0.0, "front": 0.0, "exchange": 0.0, "ingest": 0.0, "close": 0.0}  # host wall clock per phase

    # -- the reference's API --
    def insert(self, points, position) -> None:
        """points: (n, 3) float32 as a numpy array, a page-locked CPU tensor or a tensor on the engine's device."""
        pts = points if isinstance(points, torch.Tensor) else np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        pos = np.ascontiguousarray(position, np.float32).reshape(3)
        # tsdf.cpp:46-61: strictly more than 5 m from the submap's first pose, fp32 ((dx*dx + dy*dy) + dz*dz, then sqrt)
        if self._first_pose is None:
            self._first_pose = pos.copy()
        else:
            d = self._first_pose - pos
            t = d * d
            if np.sqrt(np.float32(np.float32(t[0] + t[1]) + t[2])) > np.float32(5.0):
                self._close_submap()
                self._first_pose = pos.copy()
        if len(pts):
            self._scans.append((pts, pos))
        if len(self._scans) >= self.max_batch:
            self._process_batch()

    def flush(self) -> None:
        self._process_batch()

    def finalize_active(self) -> None:
        """The part of TSDFMap::save before meshing (tsdf.cpp:78-81)."""
        if self._first_pose is not None:
            self._close_submap()
            self._first_pose = None

    # -- internals --
    def _process_batch(self) -> None:
        if not self._scans:
            return
        t0 = time.perf_counter()
        xyz = self.engine.concat([p for p, _ in self._scans])
        offsets = np.concatenate([[0], np.cumsum([len(p) for p, _ in self._scans])]).astype(np.uint32)
        poses = np.stack([q for _, q in self._scans]).astype(np.float32)
        self._scans = []
        self.phase_s["stage"] += time.perf_counter() - t0
        t0 = time.perf_counter()
        counts, send = self.engine.front(xyz, offsets, poses, self.rank, self.world, self._new_submap)
        t1 = time.perf_counter()
        self._new_submap = False
        send_counts = self.engine.empty((self.world,))
        send_counts.copy_(torch.tensor(counts, dtype=torch.int64))
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        recv_list = [int(c) for c in recv_counts.tolist()]
        recv = self.engine.empty((sum(recv_list), TUPLE_WORDS))
        dist.all_to_all_single(recv, send, output_split_sizes=recv_list, input_split_sizes=counts, group=self.group)
        self.exchanged_tuples += sum(c for r, c in enumerate(counts) if r != self.rank)
        self.engine.sync()
        t2 = time.perf_counter()
        self.engine.ingest(recv)
        t3 = time.perf_counter()
        self.phase_s["front"] += t1 - t0
        self.phase_s["exchange"] += t2 - t1
        self.phase_s["ingest"] += t3 - t2

    def _close_submap(self) -> None:
        self._process_batch()
        t0 = time.perf_counter()
        keys, cells = self.engine.export_chunks()
        n = self.engine.empty((1,))
        n.fill_(keys.shape[0])
        all_n = self.engine.empty((self.world,))
        dist.all_gather_into_tensor(all_n, n, group=self.group)
        counts = [int(c) for c in all_n.tolist()]
        pad = max(max(counts), 1)
        kbuf = self.engine.empty((pad,))
        cbuf = self.engine.empty((pad, CELL_WORDS))
        kbuf[: keys.shape[0]] = keys
        cbuf[: keys.shape[0]] = cells
        all_k = self.engine.empty((self.world * pad,))
        all_c = self.engine.empty((self.world * pad, CELL_WORDS))
        dist.all_gather_into_tensor(all_k, kbuf, group=self.group)
        dist.all_gather_into_tensor(all_c, cbuf, group=self.group)
        # rank order == ascending Morton order (contiguous ranges)
        gk = torch.cat([all_k[r * pad: r * pad + c] for r, c in enumerate(counts)])
        gc = torch.cat([all_c[r * pad: r * pad + c] for r, c in enumerate(counts)])
        self.engine.sync()
        self.engine.finalize_from(gk, gc)
        self._new_submap = True
        self.phase_s["close"] += time.perf_counter() - t0


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
