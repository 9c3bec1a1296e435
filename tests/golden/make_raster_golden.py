#!/usr/bin/env python
"""Generate tests/golden/raster_golden.npz with the REFERENCE's own CUDA code.

Runs on a GPU box (gpurun): drives oracle/_ref/ref_dgr_C.so + ref_knn_C.so - the
reference rasterizer and simple-knn compiled unmodified from /root/reference by
oracle/build_ref.py - on small seeded scenes and stores inputs, every intermediate
(sliced out of the reference's geometry / binning / image buffers) and all
gradients.  These vectors pin oracle/gsr_oracle.c (tests/test_oracle_cpu.py) and
are compared directly with the sm_100a kernels (tests/test_gpu_parity.py).

    gpurun -- python tests/golden/make_raster_golden.py gpurun_out/raster_golden.npz
    cp gpurun_out/raster_golden.npz tests/golden/
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))
import synthetic  # noqa: E402
from oracle import ref_driver  # noqa: E402

CASES = {
    # name: (P, W, H, sh_degree, scale_mult, bg, scale_modifier, mode)
    "sh3": (900, 120, 72, 3, 3.0, (0.1, 0.2, 0.3), 1.0, "sh"),
    "sh1": (500, 64, 64, 1, 4.0, (0.0, 0.0, 0.0), 0.8, "sh"),
    "precomp": (600, 100, 50, 0, 3.0, (1.0, 1.0, 1.0), 1.0, "precomp"),
}


def run_case(name, out):
    P, W, H, deg, smult, bg, smod, mode = CASES[name]
    dev = "cuda"
    sc = synthetic.make_scene(P, seed=11 + len(name), device=dev, scale_mult=smult)
    cam = synthetic.make_camera(1, 7, W, H, device=dev)
    rs = synthetic.raster_settings(cam, torch.tensor(bg, device=dev), sh_degree=deg, scale_modifier=smod)
    grad = synthetic.make_image_grad(W, H, seed=5, device=dev)
    kw = {}
    if mode == "sh":
        kw = dict(shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    else:
        g = torch.Generator().manual_seed(3)
        colors = torch.rand((P, 3), generator=g).to(dev)
        # a valid precomputed covariance: take the reference's own cov3D from a scale/rot run
        f0 = ref_driver.forward(rs, sc["means3D"], sc["opacities"], shs=sc["shs"], scales=sc["scales"],
                                rotations=sc["rotations"])
        cov = ref_driver.slice_geom(f0["geom"], P)["cov3D"].clone()
        # Gaussians culled by the near plane never got a covariance written: give them one
        cov[f0["radii"] == 0] = torch.tensor([1e-3, 0, 0, 1e-3, 0, 1e-3], device=dev)
        kw = dict(colors_precomp=colors, cov3D_precomp=cov)
    f = ref_driver.forward(rs, sc["means3D"], sc["opacities"], **kw)
    b = ref_driver.backward(rs, f, grad, sc["means3D"], **kw)
    torch.cuda.synchronize()
    R = f["num_rendered"]
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    gs = ref_driver.slice_geom(f["geom"], P)
    bs = ref_driver.slice_binning(f["binning"], R)
    im = ref_driver.slice_img(f["img"], W, H)
    vis = (f["radii"] > 0)

    def put(k, t):
        out["%s/%s" % (name, k)] = t.detach().cpu().numpy()

    put("cfg", torch.tensor([P, W, H, deg, R], dtype=torch.int64))
    put("scale_modifier", torch.tensor([smod]))
    put("tanfov", torch.tensor([rs.tanfovx, rs.tanfovy], dtype=torch.float64))
    for k in ("bg", "viewmatrix", "projmatrix", "campos"):
        put(k, getattr(rs, k))
    put("means3D", sc["means3D"]); put("opacities", sc["opacities"]); put("grad_image", grad)
    for k, v in kw.items():
        put(k, v)
    put("radii", f["radii"]); put("color", f["color"])
    put("depths", torch.where(vis, gs["depths"], torch.zeros_like(gs["depths"])))
    put("means2D", gs["means2D"] * vis[:, None]); put("conic_opacity", gs["conic_opacity"] * vis[:, None])
    put("rgb", gs["rgb"] * vis[:, None]); put("clamped", gs["clamped"] & vis[:, None])
    put("tiles_touched", gs["tiles_touched"])
    if mode == "sh":
        put("cov3D", gs["cov3D"] * vis[:, None])
    put("keys_sorted", bs["point_list_keys"]); put("point_list", bs["point_list"])
    put("ranges", im["ranges"][:tiles]); put("n_contrib", im["n_contrib"]); put("final_T", im["accum_alpha"])
    for k, v in b.items():
        put("grad_" + k, v)
    put("mark_visible", ref_driver.mark_visible(sc["means3D"], rs.viewmatrix, rs.projmatrix))
    print(name, "P", P, "R", R, "visible", int(vis.sum()))


def main():
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "raster_golden.npz")
    out = {}
    for name in CASES:
        run_case(name, out)
    pts = synthetic.make_scene(700, seed=21, device="cuda")["means3D"]
    pts[10] = pts[11]                       # a duplicate point (distance 0 counts)
    out["knn/points"] = pts.cpu().numpy()
    out["knn/dist2"] = ref_driver.dist_cuda2(pts).cpu().numpy()
    os.makedirs(os.path.dirname(os.path.abspath(dst)), exist_ok=True)
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
