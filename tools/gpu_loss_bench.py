#!/usr/bin/env python
"""Times the fused training loss + fused Adam against the torch implementations the reference uses
(utils/loss_utils.py ops via oracle/loss_port.py; torch.optim.Adam) on the same GPU."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))
import fused_adam  # noqa: E402
import loss_utils  # noqa: E402
from oracle import loss_port  # noqa: E402


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    H, W, P = 1080, 1920, 1000000
    g = torch.Generator().manual_seed(0)
    img = torch.rand((3, H, W), generator=g).cuda()
    gt = torch.rand((3, H, W), generator=g).cuda()

    def ours():
        x = img.clone().requires_grad_(True)
        loss_utils.l1_ssim_loss(x, gt, 0.2).backward()

    def ref():
        x = img.clone().requires_grad_(True)
        loss_port.training_loss(x, gt, 0.2).backward()

    out = {"loss_fwd_bwd_ms": {"ours": timeit(ours), "torch_reference_ops": timeit(ref)}}
    shapes = [(P, 3), (P, 1, 3), (P, 15, 3), (P, 1), (P, 3), (P, 4), (P, 6), (P,)]
    lrs = [1.6e-4, 2.5e-3, 1.25e-4, 0.05, 5e-3, 1e-3, 1e-3, 1e-3]
    base = [torch.randn(s, generator=g).cuda() for s in shapes]
    pa = [p.clone().requires_grad_(True) for p in base]
    pb = [p.clone().requires_grad_(True) for p in base]
    oa = fused_adam.FusedAdam([{"params": [p], "lr": lr} for p, lr in zip(pa, lrs)], lr=0.0, eps=1e-15)
    ob = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(pb, lrs)], lr=0.0, eps=1e-15)
    for p, q in zip(pa, pb):
        gr = torch.randn(p.shape, generator=g).cuda() * 0.01
        p.grad.copy_(gr)
        q.grad = gr.clone()
    n_par = sum(p.numel() for p in pa)
    ta, tb = timeit(oa.step), timeit(ob.step)
    out["adam_step_ms"] = {"ours": ta, "torch_optim_Adam": tb, "params": n_par,
                           "ours_GBps": 28.0 * n_par / ta / 1e6}
    import gsr_runtime as rt
    rt.profile_enable(True)
    for _ in range(5):
        ours()
        oa.step()
    torch.cuda.synchronize()
    prof = rt.profile_dump()
    rt.profile_enable(False)
    out["kernels_ms"] = {k: round(t / max(c, 1), 5) for k, (c, t) in prof.items()}
    px = 3 * H * W
    if "ssim_l1_fwd" in out["kernels_ms"]:      # algorithmic bytes: fwd reads 2 images, writes 3 maps; bwd reads 3 maps + 2 images, writes 1
        out["loss_alg_GBps"] = {"fwd": 4.0 * px * 5 / out["kernels_ms"]["ssim_l1_fwd"] / 1e6,
                                "bwd": 4.0 * px * 6 / out["kernels_ms"]["ssim_l1_bwd"] / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
