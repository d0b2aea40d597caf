"""CPU, world_size 2 on gloo: the host logic of the Morton-range sharded map (chad_tsdf_b200/sharded.py) with a
numpy/oracle engine in place of the GPU: the union of the shards must equal the single map bit for bit, every shard
must hold only keys of its range, and the chunk streams gathered at a submap switch must be the whole submap, sorted."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from chad_tsdf_b200 import synth
    from chad_tsdf_b200.sharded import ShardedTSDFMap
    from oracle import bindings as ob
    from tests.shard_cpu_engine import NumpyShardEngine
    w = synth.Workload("t", synth.BOX_ROOM, 16, 5, -2.0, 1.8, 0.05, 0.10, seed=5)  # 5 scans, 1.8 m apart: one switch at scan 3
    eng = NumpyShardEngine(w.sdf_res, w.sdf_trunc)
    m = ShardedTSDFMap(eng, max_batch_scans=2)
    o = ob.OracleMap(w.sdf_res, w.sdf_trunc)
    oracle_before_switch = None
    for s in range(w.scans):
        pts, pos = w.scan(s)
        before = o.voxels()
        if o.insert(pts, pos) == 1 and oracle_before_switch is None:
            oracle_before_switch = before
        m.insert(pts, pos)
    m.flush()
    keys, sd, wt = eng.voxels()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), keys=keys, sd=sd, w=wt, splitters=eng.splitters, fin_keys=eng.finalized[0][0],
             fin_cells=eng.finalized[0][1], exchanged=m.exchanged_tuples)
    if rank == 0:
        ok, osd, ow = o.voxels()
        np.savez(os.path.join(out_dir, "oracle.npz"), keys=ok, sd=osd, w=ow, bk=oracle_before_switch[0], bsd=oracle_before_switch[1], bw=oracle_before_switch[2])
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_sharding_reproduces_the_single_map(tmp_path, oracle_lib):
    world = 2
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{i}.npz") for i in range(world)]
    o = np.load(tmp_path / "oracle.npz")
    # every rank derived the same splitters; shards are disjoint key ranges in rank order
    assert np.array_equal(r[0]["splitters"], r[1]["splitters"])
    spl = r[0]["splitters"][0]
    assert np.all((r[0]["keys"] >> np.uint64(9)) < spl) and np.all((r[1]["keys"] >> np.uint64(9)) >= spl)
    assert len(r[0]["keys"]) > 0 and len(r[1]["keys"]) > 0 and r[0]["exchanged"] > 0
    # union of the shards == the single map, bit for bit (keys, fp32 sd bits, weights)
    keys = np.concatenate([r[0]["keys"], r[1]["keys"]])
    assert np.array_equal(keys, o["keys"])
    assert np.array_equal(np.concatenate([r[0]["sd"], r[1]["sd"]]), o["sd"])
    assert np.array_equal(np.concatenate([r[0]["w"], r[1]["w"]]), o["w"])
    # the chunk stream gathered at the submap switch is identical on both ranks and equals the closed submap
    assert np.array_equal(r[0]["fin_keys"], r[1]["fin_keys"]) and np.array_equal(r[0]["fin_cells"], r[1]["fin_cells"])
    fk, fc = r[0]["fin_keys"], r[0]["fin_cells"]
    assert np.all(np.diff(fk.astype(np.int64)) > 0)
    present = (fc >> np.uint64(32)) != 0
    vk = ((fk[:, None] << np.uint64(3)) | np.arange(8, dtype=np.uint64)[None, :])[present]
    assert np.array_equal(vk, o["bk"])
    assert np.array_equal((fc[present] & np.uint64(0xFFFFFFFF)).astype(np.uint32), o["bsd"])
    assert np.array_equal((fc[present] >> np.uint64(32)).astype(np.uint32), o["bw"])


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
