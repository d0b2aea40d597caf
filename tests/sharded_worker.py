"""Multi-GPU worker (one rank per GPU, NCCL): ONE map on all ranks (Morton-range shards, or submaps round-robin) vs the CPU oracle.
Launch: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/sharded_worker.py <out_dir>"""
import json
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before CUDA starts: see bench.py

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from chad_tsdf_b200 import synth  # noqa: E402
from chad_tsdf_b200.sharded import CudaShardEngine, SubmapParallelTSDFMap, create_sharded_map, sharded_digest  # noqa: E402
from oracle import bindings as ob  # noqa: E402


def morton(out_dir, rank, world, local, case):
    """ONE map cut into `world` Morton ranges (chad_create_sharded): voxels of the active submap gathered from the ranks, the DAG read
    from rank 0, both against the CPU oracle's digests."""
    if case == "fine":   # truncation / voxel size = 3: longer bands, more runs cross the range boundaries
        w = synth.Workload("t", synth.INDOOR, 32, 5, -1.0, 1.4, 0.04, 0.12, seed=11)
    else:                # 9 scans, 1.3 m apart: submap switches at scans 4 and 8
        w = synth.Workload("t", synth.BOX_ROOM, 32, 9, -3.0, 1.3, 0.05, 0.10, seed=9)
    m = create_sharded_map(w.sdf_res, w.sdf_trunc, local, max_batch_scans=3)
    o = ob.OracleMap(w.sdf_res, w.sdf_trunc)
    for s in range(w.scans):
        pts, pos = w.scan(s)
        m.insert(pts, pos)
        if rank == 0:
            o.insert(pts, pos)
    m.flush()
    before = sharded_digest(m, with_dag=False)
    local_voxels = int(len(m.voxels()[0]))
    m.finalize_active()
    after = sharded_digest(m)
    result = {"rank": rank, "world": world, "info": m.shard_info(), "local_voxels": local_voxels, "roots": m.roots()}
    if rank == 0:
        ob_before = ob.map_digest(o)
        o.finalize_active()
        ob_after = ob.map_digest(o)
        result["voxels_match"] = all(before[k] == ob_before[k] for k in ("voxels_n", "voxels_keys", "voxels_sd_bits", "voxels_weights", "weight_sum"))
        result["roots_match"] = after["roots"] == ob_after["roots"]
        result["dag_matches_oracle"] = after["levels"] == ob_after["levels"]
        result["voxels_n"] = before["voxels_n"]
        result["submaps"] = len(after["roots"])
    with open(os.path.join(out_dir, f"result{rank}.json"), "w") as f:
        json.dump(result, f)
    m.close()


def submaps(out_dir, rank, world, local):
    w = synth.Workload("t", synth.BOX_ROOM, 32, 7, -3.0, 1.3, 0.05, 0.10, seed=9)  # 7 scans, 1.3 m apart: a switch at scan 4
    eng = CudaShardEngine(w.sdf_res, w.sdf_trunc, local, max_batch_scans=3)
    m = SubmapParallelTSDFMap(eng)
    o = ob.OracleMap(w.sdf_res, w.sdf_trunc)
    for s in range(w.scans):
        pts, pos = w.scan(s)
        m.insert(pts, pos)
        o.insert(pts, pos)
    m.flush()
    result = {"rank": rank, "world": world, "exchanged": m.broadcast_chunks}
    keys, sd, wt = eng.voxels()
    ok, osd, ow = o.voxels()
    np.savez(os.path.join(out_dir, f"shard{rank}.npz"), keys=keys, sd=sd, w=wt)
    m.finalize_active()
    o.finalize_active()
    result["roots_match"] = eng.roots() == o.roots()
    lv_ok = True
    for lv in range(21):
        ga, gu, gd = eng.level(lv)
        oa, ou, od = o.level(lv)
        lv_ok &= (gu, gd) == (ou, od) and np.array_equal(ga, oa)
    result["dag_matches_oracle"] = bool(lv_ok)
    result["roots"] = eng.roots()
    if rank == 0:
        np.savez(os.path.join(out_dir, "oracle.npz"), keys=ok, sd=osd, w=ow)
    with open(os.path.join(out_dir, f"result{rank}.json"), "w") as f:
        json.dump(result, f)
    eng.close()


def main():
    out_dir = sys.argv[1]
    mode = sys.argv[2] if len(sys.argv) > 2 else "morton"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if mode == "submaps":
        submaps(out_dir, rank, world, local)
    else:
        morton(out_dir, rank, world, local, mode.partition(":")[2])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
