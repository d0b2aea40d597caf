import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, "/root/repo")
os.environ.setdefault("MASTER_ADDR","127.0.0.1"); os.environ.setdefault("MASTER_PORT","29533")
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda",0))
from chad_tsdf_b200 import synth
from chad_tsdf_b200.sharded import CudaShardEngine, ShardedTSDFMap
w = synth.WORKLOADS["cfg1_traj100_128beam"]
eng = CudaShardEngine(w.sdf_res, w.sdf_trunc, 0)
for nb in (1, 4, 16):
    eng.map.reset()
    m = ShardedTSDFMap(eng, max_batch_scans=nb)
    try:
        for s in range(nb):
            m.insert(*w.scan(s))
        m.flush()
        print(nb, "ok", eng.stats()["updates"], eng.stats()["resident_clusters"])
    except Exception as e:
        print(nb, "FAIL", e)
dist.destroy_process_group()
