#!/usr/bin/env python
"""Deformation network at P points: tensor-core path vs the reference's torch classes on the same GPU.
Prints one JSON line: ms forward, ms forward+backward, per-kernel times, tensor-pipe roofline of the hidden-layer GEMM."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))
import deform_mlp, gsr_runtime as rt

ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=1000000)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--no-reference", action="store_true")
ap.add_argument("--only", default="", help="fwd: forward only (for ncu)")
a = ap.parse_args()
P = a.P
torch.manual_seed(0)
ours = deform_mlp.DirectTemporalNeRF().cuda()
g = torch.Generator().manual_seed(1)
x0 = ((torch.rand((P, 3), generator=g) * 2 - 1) * 1.3).cuda()
proj = [torch.randn((P, c), generator=g).cuda() for c in (3, 3, 4, 48)]
ts = torch.full((P, 1), 0.4, device="cuda")


def timed(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fwd(net):
    with torch.no_grad():
        return net(x0, 0.4 if net is ours else ts, 5000)


def fwd_bwd(net):
    x = x0.clone().requires_grad_(True)
    for p in net.parameters():
        p.grad = None
    outs = net(x, 0.4 if net is ours else ts, 5000)
    torch.autograd.backward(outs, proj)


out = {"P": P}
if a.only == "fwd":
    fwd(ours); torch.cuda.synchronize(); print(json.dumps({"ok": 1})); sys.exit(0)
out["ours_fwd_ms"] = timed(lambda: fwd(ours), a.steps)
out["ours_fwd_bwd_ms"] = timed(lambda: fwd_bwd(ours), a.steps)
rt.profile_enable(True)
fwd_bwd(ours); torch.cuda.synchronize()
out["kernels_fwd_bwd"] = {k: {"launches": c, "total_ms": round(t, 4)} for k, (c, t) in rt.profile_dump().items()}
rt.profile_enable(False)
flops_fwd = 2.0 * P * (64 * 256 + 6 * 256 * 256 + 320 * 256 + 256 * 58)
out["fwd_useful_tflops"] = flops_fwd / out["ours_fwd_ms"] / 1e9
out["fwd_tensor_tflops_3x"] = 3 * out["fwd_useful_tflops"]
if not a.no_reference:
    from oracle import ref_py
    gm = ref_py.gaussian_model()
    ref = gm.DirectTemporalNeRF().cuda()
    ref.load_state_dict(ours.state_dict())
    out["reference_fwd_ms"] = timed(lambda: fwd(ref), max(2, a.steps // 2))
    out["reference_fwd_bwd_ms"] = timed(lambda: fwd_bwd(ref), max(2, a.steps // 2))
    out["speedup_fwd"] = out["reference_fwd_ms"] / out["ours_fwd_ms"]
    out["speedup_fwd_bwd"] = out["reference_fwd_bwd_ms"] / out["ours_fwd_bwd_ms"]
out["peak_mem_GB"] = torch.cuda.max_memory_allocated() / 2 ** 30
print(json.dumps(out))
