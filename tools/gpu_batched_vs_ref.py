#!/usr/bin/env python
"""Full-size check of the view-batched step against the REFERENCE rasterizer (oracle/_ref), un-deformed scene:
gradients summed over V views by (a) the reference per view, (b) this library per view, (c) this library batched."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import synthetic, view_parallel as vp
from oracle import ref_driver
from diff_gaussian_rasterization import GaussianRasterizer, GaussianBackwardBatch, GaussianForwardBatch

P, W, H, V = int(os.environ.get("P", 1000000)), 1920, 1080, int(os.environ.get("V", 4))
sc = synthetic.make_scene(P, device="cuda")
grad = synthetic.make_image_grad(W, H, device="cuda")
bg = torch.zeros(3, device="cuda")
settings = [synthetic.raster_settings(synthetic.make_camera(k, 64, W, H, device="cuda"), bg, sh_degree=3) for k in range(V)]
names = ("means3D", "opacities", "shs", "scales", "rotations")

def ours(batched):
    lv = {k: sc[k].clone().requires_grad_(True) for k in names}
    vp.FlatGradBuffer(list(lv.values()))
    sinks = {k: v.grad for k, v in lv.items()}
    fwd = GaussianForwardBatch(settings, **lv) if batched else None
    batch = GaussianBackwardBatch(sinks) if batched else None
    for k in range(V):
        m2 = torch.zeros(P, 3, device="cuda", requires_grad=True)
        c, _ = GaussianRasterizer(settings[k])(means3D=lv["means3D"], means2D=m2, opacities=lv["opacities"], shs=lv["shs"], scales=lv["scales"],
                                               rotations=lv["rotations"], accumulate_grads=batch if batched else sinks,
                                               prepared=fwd.prepared(k) if batched else None)
        c.backward(grad)
    if batched:
        batch.flush()
    torch.cuda.synchronize()
    return {k: v.grad.clone() for k, v in lv.items()}

def ref():
    tot = None
    for k in range(V):
        f = ref_driver.forward(settings[k], sc["means3D"], sc["opacities"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
        g = ref_driver.backward(settings[k], f, grad, sc["means3D"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
        g = {n: g[n].double() for n in names}
        tot = g if tot is None else {n: tot[n] + g[n] for n in names}
    return tot

def stats(a, b):
    a, b = a.double().reshape(b.shape), b.double()
    d = (a - b).abs()
    mx = float(b.abs().max())
    rows = b.reshape(b.shape[0], -1)
    rn = rows.norm(dim=1)
    rel = (a.reshape(rows.shape) - rows).norm(dim=1) / (rn + 1e-30)
    big = rn > 1e-3 * rn.max()
    return {"max_abs_over_max": float(d.max()) / mx, "viol_1e-4rel+5e-5max": int((d > 1e-4 * b.abs() + 5e-5 * mx).sum()),
            "row_rel_p999": float(torch.quantile(rel[big][:4000000].float(), 0.999)) if int(big.sum()) else 0.0}

r = ref(); r2 = ref()
a = ours(False); b = ours(True)
for n in names:
    print(n, json.dumps({"ref_vs_ref": stats(r2[n], r[n]), "per_view_vs_ref": stats(a[n], r[n]), "batched_vs_ref": stats(b[n], r[n]),
                         "batched_vs_per_view": stats(b[n], a[n])}))
