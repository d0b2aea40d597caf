"""Multi-GPU worker (one rank per GPU, NCCL): Morton-range sharded map vs the CPU oracle.
Launch: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/sharded_worker.py <out_dir>"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from chad_tsdf_b200 import synth  # noqa: E402
from chad_tsdf_b200.sharded import CudaShardEngine, ShardedTSDFMap, SubmapParallelTSDFMap  # noqa: E402
from oracle import bindings as ob  # noqa: E402


def main():
    out_dir = sys.argv[1]
    mode = sys.argv[2] if len(sys.argv) > 2 else "morton"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = synth.Workload("t", synth.BOX_ROOM, 32, 7, -3.0, 1.3, 0.05, 0.10, seed=9)  # 7 scans, 1.3 m apart: a switch at scan 4
    eng = CudaShardEngine(w.sdf_res, w.sdf_trunc, local, max_batch_scans=3 if mode == "submaps" else 1)
    m = SubmapParallelTSDFMap(eng) if mode == "submaps" else ShardedTSDFMap(eng, max_batch_scans=3)
    o = ob.OracleMap(w.sdf_res, w.sdf_trunc)
    for s in range(w.scans):
        pts, pos = w.scan(s)
        m.insert(pts, pos)
        o.insert(pts, pos)
    m.flush()
    result = {"rank": rank, "world": world, "exchanged": m.broadcast_chunks if mode == "submaps" else m.exchanged_tuples}
    # this rank's shard must be exactly the oracle's voxels of its key range
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
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
