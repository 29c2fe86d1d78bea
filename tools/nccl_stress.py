"""NCCL-only stress at the data-parallel step's collective pattern (no aldi_b200 kernels): 5 async sum-all-reduces of
the bucket sizes of the 165 MB flat gradient per iteration, overlapped with torch matmuls on the compute stream.
Used to tell a box / NCCL problem from an interaction with our kernels.   torchrun ... tools/nccl_stress.py [iters]"""
import os
import sys

import torch
import torch.distributed as dist

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
sizes = [14_300_000, 3_300_000, 15_000_000, 7_100_000, 1_200_000]
grad = torch.randn(sum(sizes), device=dev)
a = torch.randn(4096, 4096, device=dev, dtype=torch.bfloat16)
for it in range(iters):
    off, works = 0, []
    for s in sizes:
        for _ in range(4):
            a = (a @ a).clamp_(-1, 1)
        works.append(dist.all_reduce(grad[off:off + s], async_op=True))
        off += s
    for w in works:
        w.wait()
    grad.mul_(1.0 / dist.get_world_size())
torch.cuda.synchronize()
if rank == 0:
    print("nccl_stress ok: %d iterations, world %d, |grad| %.4g" % (iters, dist.get_world_size(), float(grad.norm())))
dist.destroy_process_group()
