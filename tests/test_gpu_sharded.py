"""GPU: the Morton-range sharded map (chad_shard_* + chad_tsdf_b200/sharded.py over NCCL) against the CPU oracle.
world = 1 runs everywhere; world = 2 needs two GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_sharded.py -m gpu`)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, tmp_path, mode="morton"):
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "sharded_worker.py"), str(tmp_path), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    res = [json.load(open(tmp_path / f"result{i}.json")) for i in range(world)]
    shards = [np.load(tmp_path / f"shard{i}.npz") for i in range(world)]
    o = np.load(tmp_path / "oracle.npz")
    for r_ in res:
        assert r_["roots_match"] and r_["dag_matches_oracle"], r_  # every rank holds the identical, exact DAG
    keys = np.concatenate([s["keys"] for s in shards])
    assert np.array_equal(keys, o["keys"])  # shards are disjoint ascending ranges whose union is the single map
    assert np.array_equal(np.concatenate([s["sd"] for s in shards]), o["sd"])
    assert np.array_equal(np.concatenate([s["w"] for s in shards]), o["w"])
    return res


def test_sharded_world1(chad_lib, oracle_lib, tmp_path):
    _run(1, tmp_path)


def test_sharded_world2(chad_lib, oracle_lib, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = _run(2, tmp_path)
    assert all(r["exchanged"] > 0 for r in res)  # updates really crossed NVLink


def test_submap_parallel_world1(chad_lib, oracle_lib, tmp_path):
    _run(1, tmp_path, "submaps")


def test_submap_parallel_world2(chad_lib, oracle_lib, tmp_path):
    """The map's submaps integrated on alternating GPUs, closed submaps broadcast over NCCL: same voxels, same DAG."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = _run(2, tmp_path, "submaps")
    assert all(r["exchanged"] > 0 for r in res)
