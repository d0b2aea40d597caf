"""CPU, world_size > 1 on gloo: the host logic of the multi-GPU modes (chad_tsdf_b200/sharded.py) without a GPU. The Morton-range
sharded map itself lives in the C++ library (its kernels and NCCL calls cannot run here); what is Python is the gather + digest that
proves the shards form ONE map, tested here on range shards cut from the oracle's map. The submap-parallel driver runs on a
numpy/oracle engine: every rank must be handed, in order, exactly the chunk stream of every submap of the single map."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _RangeShard:
    """A rank of a Morton-range sharded map, faked from the oracle's single map: the voxels of this rank's key range, the roots on
    every rank, the DAG levels on rank 0 only (what chad_create_sharded contexts expose)."""

    def __init__(self, oracle_map, rank, lo, hi, overlap=0):
        k, sd, w = oracle_map.voxels()
        sel = (k >= np.uint64(lo)) & (k < np.uint64(hi + overlap))
        self._vox = (k[sel], sd[sel], w[sel])
        self._o, self._rank = oracle_map, rank

    def voxels(self):
        return self._vox

    def roots(self):
        return self._o.roots()

    def level(self, lv):
        assert self._rank == 0, "the DAG levels live on rank 0"
        return self._o.level(lv)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import json
    from chad_tsdf_b200 import synth
    from chad_tsdf_b200.sharded import sharded_digest
    from oracle import bindings as ob
    w = synth.Workload("t", synth.BOX_ROOM, 16, 5, -2.0, 1.8, 0.05, 0.10, seed=5)  # 5 scans, 1.8 m apart: one switch at scan 3
    o = ob.OracleMap(w.sdf_res, w.sdf_trunc)
    for s in range(w.scans):
        o.insert(*w.scan(s))
    keys = o.voxels()[0]
    cuts = [0] + [int(keys[len(keys) * g // world]) for g in range(1, world)] + [2**63]
    before = sharded_digest(_RangeShard(o, rank, cuts[rank], cuts[rank + 1]), with_dag=False)
    try:  # ranges that overlap do not form one map: the gather must say so
        sharded_digest(_RangeShard(o, rank, cuts[rank], cuts[rank + 1], overlap=4096 if rank == 0 else 0), with_dag=False)
        overlap_caught = False
    except AssertionError:
        overlap_caught = True
    want_before = ob.map_digest(o)
    o.finalize_active()
    after = sharded_digest(_RangeShard(o, rank, cuts[rank], cuts[rank + 1]))
    if rank == 0:
        want = ob.map_digest(o)
        json.dump({"before": all(before[k] == want_before[k] for k in ("voxels_n", "voxels_keys", "voxels_sd_bits", "voxels_weights", "weight_sum")),
                   "after": after == want, "overlap_caught": overlap_caught, "submaps": len(after["roots"])}, open(os.path.join(out_dir, "digest.json"), "w"))
    else:
        assert before is None and after is None
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_digest_of_range_shards_is_the_single_maps_digest(tmp_path, oracle_lib, world):
    """The parity check bench.py --gpus N runs on every rank count: voxels gathered in rank order + the DAG of rank 0 hash to exactly what
    the single map hashes to (so the golden pins of the reference apply to the sharded map unchanged)."""
    import json
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = json.load(open(tmp_path / "digest.json"))
    assert r == {"before": True, "after": True, "overlap_caught": True, "submaps": 2}


def _worker_submaps(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from chad_tsdf_b200 import synth
    from chad_tsdf_b200.sharded import SubmapParallelTSDFMap
    from oracle import bindings as ob
    from tests.shard_cpu_engine import NumpyShardEngine
    w = synth.Workload("t", synth.BOX_ROOM, 16, 7, -3.0, 1.8, 0.05, 0.10, seed=6)  # 7 scans, 1.8 m apart: switches at scans 3 and 6
    eng = NumpyShardEngine(w.sdf_res, w.sdf_trunc)
    m = SubmapParallelTSDFMap(eng)
    o = ob.OracleMap(w.sdf_res, w.sdf_trunc)
    closed = []  # the oracle's voxels of every closed submap
    nsub = 0
    for s in range(w.scans):
        pts, pos = w.scan(s)
        before = o.voxels()
        got = o.insert(pts, pos)  # = submaps finalised so far
        if got != nsub:
            closed.append(before)
            nsub = got
        m.insert(pts, pos)
    m.flush()
    closed.append(o.voxels())
    m.finalize_active()
    out = {"n": len(eng.finalized), "owned": m.owned_scans}
    for i, (fk, fc) in enumerate(eng.finalized):
        out[f"fk{i}"], out[f"fc{i}"] = fk, fc
    np.savez(os.path.join(out_dir, f"sp{rank}.npz"), **out)
    if rank == 0:
        oo = {"n": len(closed)}
        for i, (k, sd, wt) in enumerate(closed):
            oo[f"k{i}"], oo[f"sd{i}"], oo[f"w{i}"] = k, sd, wt
        np.savez(os.path.join(out_dir, "sp_oracle.npz"), **oo)
    dist.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 3])
def test_submap_parallel_reproduces_every_submap(tmp_path, oracle_lib, world):
    """Submaps integrated round-robin on the ranks (closes lag world - 1 switches): every rank must be handed, in order, exactly
    the chunk stream of every submap of the single map (that stream is all Submap::finalize consumes)."""
    port = 29300 + os.getpid() % 300 + world
    mp.spawn(_worker_submaps, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"sp{i}.npz") for i in range(world)]
    o = np.load(tmp_path / "sp_oracle.npz")
    n = int(o["n"])
    assert n == 3 and all(int(x["n"]) == n for x in r)
    assert sum(int(x["owned"]) for x in r) == 7 and all(int(x["owned"]) > 0 for x in r)
    for i in range(n):
        fk, fc = r[0][f"fk{i}"], r[0][f"fc{i}"]
        for other in r[1:]:
            assert np.array_equal(fk, other[f"fk{i}"]) and np.array_equal(fc, other[f"fc{i}"])
        assert np.all(np.diff(fk.astype(np.int64)) > 0)
        present = (fc >> np.uint64(32)) != 0
        vk = ((fk[:, None] << np.uint64(3)) | np.arange(8, dtype=np.uint64)[None, :])[present]
        assert np.array_equal(vk, o[f"k{i}"])
        assert np.array_equal((fc[present] & np.uint64(0xFFFFFFFF)).astype(np.uint32), o[f"sd{i}"])
        assert np.array_equal((fc[present] >> np.uint64(32)).astype(np.uint32), o[f"w{i}"])
