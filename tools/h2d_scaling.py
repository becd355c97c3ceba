#!/usr/bin/env python
"""Host -> device copy bandwidth with N concurrent ranks (one per GPU), pinned host memory: is the end-to-end step at
N = 8 bound by the platform (PCIe root complexes / host DRAM) or by us?  Every rank copies its own pinned buffer to its
GPU back to back for a fixed time; prints per-GPU and aggregate GB/s.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_scaling.py
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {}
    for mb in (64, 256, 1024):
        nbytes = mb << 20
        host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        host.fill_(1)
        d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        for _ in range(3):
            d.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = max(4, int(8e9 // nbytes))
        t0 = time.perf_counter()
        for _ in range(reps):
            d.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gbs = torch.tensor([reps * nbytes / dt / 1e9], device=dev, dtype=torch.float64)
        lo, hi, tot = gbs.clone(), gbs.clone(), gbs.clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        out[f"{mb}MB"] = {"per_gpu_min_gbs": float(lo), "per_gpu_max_gbs": float(hi), "aggregate_gbs": float(tot)}
        del host, d
    if rank == 0:
        print(json.dumps({"n_gpus": world, "h2d_pinned": out, "cpus": os.cpu_count()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
