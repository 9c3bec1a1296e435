#!/usr/bin/env python
"""BASELINE configs[3] (C4): 6 M Gaussians at 3840x2160, fwd+bwd, ours vs reference, with parity."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synthetic
from _gpu_util import make_view_settings, run_ours, rel_to_max
from oracle import ref_driver
P, W, H = int(sys.argv[1]) if len(sys.argv) > 1 else 6000000, 3840, 2160
sc, cam, rs = make_view_settings(P, W, H, bg=(0, 0, 0))
grad = synthetic.make_image_grad(W, H, device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2], out
ms_o, o = t(lambda: run_ours(rs, sc, grad))
def ref():
    kw = dict(shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    f = ref_driver.forward(rs, sc["means3D"], sc["opacities"], **kw)
    return f, ref_driver.backward(rs, f, grad, sc["means3D"], **kw)
ms_r, (f, b) = t(ref)
res = dict(P=P, W=W, H=H, R=f["num_rendered"], ours_ms=ms_o, ref_ms=ms_r, speedup=ms_r / ms_o,
           radii_equal=bool(torch.equal(o["radii"], f["radii"])), img_max_abs=float((o["color"] - f["color"]).abs().max()),
           grads_rel={k: rel_to_max(o["grads"][k], b[k]) for k in ("means3D", "opacities", "shs", "scales", "rotations")},
           peak_mem_GB=torch.cuda.max_memory_allocated() / 2**30)
print(json.dumps(res))
