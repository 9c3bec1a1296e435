#!/usr/bin/env python
"""All-reduce of the 264 MB flat gradient buffer (66 floats x 1 M Gaussians) under torchrun: time per call and bus bandwidth.
NCCL_* env variables select the algorithm/protocol being compared."""
import os
import torch
import torch.distributed as dist


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 66 * 1000000
    buf = torch.ones(n, dtype=torch.float32, device=dev)
    for _ in range(5):
        dist.all_reduce(buf)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 30
    for _ in range(iters):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        t = float(ms.item())
        print("world %d  %s  all_reduce 264 MB: %.3f ms  busbw %.0f GB/s" % (
            world, {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}, t, 4.0 * n * 2 * (world - 1) / world / t / 1e6))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
