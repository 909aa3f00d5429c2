#!/usr/bin/env python
"""Pinned device-to-host copy rate of this box — the ceiling of bench.py's e2e number (5.9 GB of Surface Temperature per
rank and step).  Alone:  python tools/d2h_probe.py.  All ranks at once (the rate the N-GPU e2e line competes for):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_probe.py
Prints one JSON line on rank 0: per-rank GB/s (slowest, fastest) and the aggregate."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
numa = None
if world > 1:
    import torch.distributed as dist
    import bench
    numa = bench.bind_to_gpu_numa(local)          # the same first-touch placement bench.py uses
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 5888802816 // 8
d = torch.empty(n, dtype=torch.float64, device="cuda").normal_()
h = torch.empty(n, dtype=torch.float64).pin_memory()
for _ in range(2):
    h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
rate = n * 8 / dt / 1e9
if world > 1:
    t = torch.tensor([rate], dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    rates = [float(x.item()) for x in allr]
    if rank == 0:
        print(json.dumps({"n_gpus": world, "bytes_per_rank": n * 8, "per_rank_GBps_min": min(rates), "per_rank_GBps_max": max(rates),
                          "aggregate_GBps": sum(rates), "numa": numa, "host_cpus": os.cpu_count()}))
    dist.destroy_process_group()
else:
    print(json.dumps({"n_gpus": 1, "bytes_per_rank": n * 8, "per_rank_GBps_min": rate, "per_rank_GBps_max": rate, "aggregate_GBps": rate,
                      "ms": dt * 1e3}))
