"""TEST INFRASTRUCTURE: a CPU stand-in for CudaShardEngine built on the oracle's stage functions, so that the host logic
of chad_tsdf_b200/sharded.py (submap rule, batching, splitters, routing, order restoration, chunk gathering) can run
with world_size > 1 on gloo without a GPU."""
import numpy as np
import torch

from oracle import bindings as ob


class NumpyShardEngine:
    def __init__(self, sdf_res, sdf_trunc):
        self.res, self.trunc = float(sdf_res), float(sdf_trunc)
        self.vox = {}  # key -> (np.float32 sd, int weight)
        self.splitters = None
        self.finalized = []  # (keys, cells) streams handed to finalize_from

    def empty(self, shape, dtype=torch.int64):
        return torch.empty(shape, dtype=dtype)

    def sync(self):
        pass

    def concat(self, scans):
        return np.concatenate([np.ascontiguousarray(p, np.float32).reshape(-1, 3) for p in scans])

    def front(self, xyz, offsets, poses, rank, world, new_submap):
        pts_sorted, keys_sorted, normals, scan_of = [], [], [], []
        for s in range(len(offsets) - 1):
            pts = xyz[offsets[s]:offsets[s + 1]]
            x, k, order, nrm = ob.oracle_stage_points(pts, poses[s], self.res)
            pts_sorted.append(x); keys_sorted.append(k); normals.append(nrm); scan_of.append(np.full(len(x), s))
        P, K, Nn, S = map(np.concatenate, (pts_sorted, keys_sorted, normals, scan_of))
        n = len(P)
        if new_submap or self.splitters is None:
            k0 = keys_sorted[0]  # descending Morton
            n0 = len(k0)
            self.splitters = np.array([k0[n0 - 1 - ((g + 1) * n0 // world)] >> np.uint64(9) for g in range(world - 1)], dtype=np.uint64)
        i0, i1 = rank * n // world, (rank + 1) * n // world
        tk, tr, tsd = [], [], []
        for s in np.unique(S[i0:i1]):
            idx = np.nonzero(S[i0:i1] == s)[0] + i0
            pk, sd, counts = ob.oracle_stage_pairs(P[idx], Nn[idx], poses[s], self.res, self.trunc)
            tk.append(pk); tsd.append(sd.view(np.uint32)); tr.append(np.repeat(idx, counts).astype(np.uint64))
        tk, tr, tsd = (np.concatenate(a) if a else np.zeros(0, np.uint64) for a in (tk, tr, tsd))
        owner = np.searchsorted(self.splitters, tk >> np.uint64(9), side="right")
        order = np.argsort(owner, kind="stable")
        counts = [int((owner == d).sum()) for d in range(world)]
        send = np.stack([tk[order].astype(np.int64), (tr[order] | (tsd[order].astype(np.uint64) << np.uint64(32))).astype(np.int64)], axis=1) if len(tk) else np.zeros((0, 2), np.int64)
        return counts, torch.from_numpy(np.ascontiguousarray(send))

    # the whole insert of one scan on this engine (SubmapParallelTSDFMap)
    def insert(self, points, position):
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        if len(pts) == 0:
            return
        _, send = self.front(pts, np.array([0, len(pts)], np.uint32), np.asarray(position, np.float32).reshape(1, 3), 0, 1, True)
        self.ingest(send)

    def flush(self):
        pass

    def ingest(self, tuples):
        t = tuples.numpy().view(np.uint64)
        keys, rank, sd = t[:, 0], t[:, 1] & np.uint64(0xFFFFFFFF), (t[:, 1] >> np.uint64(32)).astype(np.uint32).view(np.float32)
        order = np.lexsort((rank, keys))
        for i in order:
            k = int(keys[i])
            acc, w = self.vox.get(k, (np.float32(0), 0))
            acc = np.float32(np.float32(acc * np.float32(w)) + sd[i])  # octree.hpp:161-163
            w += 1
            acc = np.float32(acc / np.float32(w))
            self.vox[k] = (acc, w)

    def voxels(self):
        keys = np.array(sorted(self.vox), dtype=np.uint64)
        sd = np.array([self.vox[int(k)][0] for k in keys], np.float32).view(np.uint32)
        w = np.array([self.vox[int(k)][1] for k in keys], np.uint32)
        return keys, sd, w

    def export_chunks(self):
        keys, sd, w = self.voxels()
        ck = np.unique(keys >> np.uint64(3))
        cells = np.zeros((len(ck), 8), np.uint64)
        pos = np.searchsorted(ck, keys >> np.uint64(3))
        cells[pos, (keys & np.uint64(7)).astype(np.int64)] = sd.astype(np.uint64) | (w.astype(np.uint64) << np.uint64(32))
        return torch.from_numpy(ck.astype(np.int64)), torch.from_numpy(cells.astype(np.int64))

    def finalize_from(self, keys, cells, clear_local=True):
        self.finalized.append((keys.numpy().view(np.uint64).copy(), cells.numpy().view(np.uint64).copy()))
        if clear_local:
            self.vox = {}

    def clear(self):
        self.vox = {}
