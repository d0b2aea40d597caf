"""Host-to-device bandwidth of page-locked 3 MB transfers (one LiDAR scan) on the box: alone / with a synchronisation after every
transfer / while the GPU is busy with a random-gather kernel (the insert path's access pattern). Run: python profiles/h2d_probe.py"""
import time

import torch

torch.cuda.init()
big = torch.empty(300 * 1024 * 1024 // 4, dtype=torch.float32, pin_memory=True)
dev = torch.empty_like(big, device="cuda")
s = torch.cuda.Stream()
busy = torch.cuda.Stream()
table = torch.randn(1 << 27, device="cuda")            # 512 MB
idx = torch.randint(0, 1 << 27, (1 << 24,), device="cuda")


def run(n_mb, mode, load):
    n = n_mb * 1024 * 1024 // 4
    reps = max(1, 300 // n_mb)
    torch.cuda.synchronize()
    if load:
        with torch.cuda.stream(busy):
            for _ in range(60):
                table.index_add_(0, idx, table[idx])   # random read-modify-write traffic for ~tens of ms
    with torch.cuda.stream(s):
        t0 = time.perf_counter()
        for r in range(reps):
            dev[r * n:(r + 1) * n].copy_(big[r * n:(r + 1) * n], non_blocking=True)
            if mode == "sync_each":
                s.synchronize()
        s.synchronize()
        dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{n_mb:4d} MB x {reps:3d} {mode:10s} load={int(load)} {reps * n * 4 / dt / 1e9:6.1f} GB/s")


for load in (False, True):
    for n_mb in (3, 48):
        for mode in ("sync_each", "async_all"):
            run(n_mb, mode, load)
            run(n_mb, mode, load)
