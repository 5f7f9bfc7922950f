#!/usr/bin/env python
"""Box probe (torchrun, one rank per GPU): bidirectional pinned memcpy bandwidth per GPU when all
ranks copy at the same time — the platform ceiling for the end-to-end figure at N > 1."""
import os, time, json
import torch, torch.distributed as dist
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
n = 1 << 31
h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device=dev); d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def both():
    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
    with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
best = 1e9
for _ in range(4):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter(); both(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
t = torch.tensor([best], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0: print(json.dumps(dict(n_gpus=world, bidir_each_GBs_per_gpu=n / t.item() / 1e9, slowest_rank_s=t.item())))
if world > 1: dist.destroy_process_group()
