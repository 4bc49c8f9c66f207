#!/usr/bin/env python
"""Developer tool (torchrun): what does the per-step gradient exchange cost on this box?
Times NCCL all-reduce of the render path's gradient set (P x 61 floats) as separate tensors and
as one flat buffer, reduce-scatter + all-gather, and checks symmetric-memory availability."""
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
sizes = {"pos": 3, "op": 1, "sc": 3, "rot": 4, "shs": 48, "ndc": 2}
parts = {k: torch.randn(P * v, device=dev) for k, v in sizes.items()}
flat = torch.randn(P * 61, device=dev)
radii = torch.randint(0, 50, (P,), device=dev, dtype=torch.int32)


def timeit(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def separate():
    ws = [dist.all_reduce(t, async_op=True) for t in parts.values()]
    ws.append(dist.all_reduce(radii, op=dist.ReduceOp.MAX, async_op=True))
    for w in ws:
        w.wait()


def one_flat():
    dist.all_reduce(flat)
    dist.all_reduce(radii, op=dist.ReduceOp.MAX)


shard = torch.empty(flat.numel() // world, device=dev)


def rs_ag():
    dist.reduce_scatter_tensor(shard, flat[: shard.numel() * world])
    dist.all_gather_into_tensor(flat[: shard.numel() * world], shard)


res = {"separate_ms": timeit(separate), "flat_ms": timeit(one_flat), "rs_ag_ms": timeit(rs_ag)}
for mb in (1, 16, 64, 244):
    t = torch.randn(mb * 250_000, device=dev)
    res[f"allreduce_{mb}MB_ms"] = timeit(lambda: dist.all_reduce(t))
try:
    import torch.distributed._symmetric_memory as symm

    t = symm.empty(1024, dtype=torch.float32, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD.group_name)
    res["symm_mem"] = {"ok": True, "multicast_ptr": bool(getattr(h, "multicast_ptr", 0)), "world": h.world_size}
except Exception as e:  # noqa: BLE001
    res["symm_mem"] = {"ok": False, "err": repr(e)[:300]}
if rank == 0:
    print(res, file=sys.stderr)
dist.destroy_process_group()
