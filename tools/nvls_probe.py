#!/usr/bin/env python
"""Developer tool (torchrun, >= 2 GPUs): correctness and cost of pxb_nvls_allreduce against NCCL."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from pointrix_b200 import parallel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
mode = sys.argv[2] if len(sys.argv) > 2 else 'auto'
ex = parallel.NvlsGradExchange(P, dev, mode=mode)
g = torch.Generator(device="cpu").manual_seed(7 + rank)
src = torch.randn(61 * P, generator=g).to(dev)
radii = torch.randint(0, 60, (P,), generator=g, dtype=torch.int32).to(dev)
ref = src.clone(); dist.all_reduce(ref)
rref = radii.clone(); dist.all_reduce(rref, op=dist.ReduceOp.MAX)
flat = ex.next_buffer(61 * P); flat.copy_(src)
r2 = radii.clone()
ex.exchange(r2)
torch.cuda.synchronize()
err = (flat - ref).abs().max().item()
ok_r = bool(torch.equal(r2, rref))
# timing
def run():
    ex.next_buffer(61 * P)
    ex.exchange(r2)
for _ in range(5):
    run()
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
t2 = []
for _ in range(5):
    dist.all_reduce(ref)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0.record()
for _ in range(20):
    dist.all_reduce(ref); dist.all_reduce(rref, op=dist.ReduceOp.MAX)
e1.record(); torch.cuda.synchronize()
tn = torch.tensor([e0.elapsed_time(e1) / 20], device=dev); dist.all_reduce(tn, op=dist.ReduceOp.MAX)
if rank == 0:
    print({"world": world, "mode": ex.mode, "max_abs_err_vs_nccl": err, "radii_equal": ok_r, "nvls_ms": t.item(), "nccl_ms": tn.item(),
           "nvls_algbw_GBs": (61 * P * 4 + 4 * P) / t.item() / 1e6}, file=sys.stderr)
dist.destroy_process_group()
