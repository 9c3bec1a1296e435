#!/usr/bin/env python
"""Caching-allocator diagnostics for the view-batched training step: cudaMalloc/cudaFree counts and wall time per step."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))
import bench  # noqa: E402


class A:
    pass


def main():
    args = A()
    args.P, args.W, args.H, args.views, args.streams, args.impl = 1000000, 1920, 1080, 8, int(os.environ.get("STREAMS", "4")), "ours"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    host, cams, bg, grad = bench.build_workload(args, 0, 1, dev)
    leaves, opt, targets = bench.make_train_state(host, dev, "ours", len(cams), args.W, args.H)
    for step in range(int(os.environ.get("STEPS", "14"))):
        s0 = torch.cuda.memory_stats()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bench.train_step(leaves, opt, targets, cams, bg, args, 1)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        s1 = torch.cuda.memory_stats()
        print("step %2d  %.2f ms  cudaMalloc %d  cudaFree %d  reserved %.2f GB  allocated peak %.2f GB  retries %d" % (
            step, dt, s1["num_device_alloc"] - s0["num_device_alloc"], s1["num_device_free"] - s0["num_device_free"],
            s1["reserved_bytes.all.current"] / 2 ** 30, s1["allocated_bytes.all.peak"] / 2 ** 30,
            s1["num_alloc_retries"] - s0["num_alloc_retries"]))


if __name__ == "__main__":
    main()
