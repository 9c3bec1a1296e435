#!/usr/bin/env python
"""Stage-by-stage comparison of libgsr_b200 against the rebuilt reference CUDA
rasterizer (oracle/_ref) on a B200, with mismatch statistics for every
intermediate.  Diagnostic companion of tests/test_gpu_parity.py.

usage: python tools/gpu_diag.py [--P 100000] [--W 640] [--H 360] [--time] [--out gpurun_out/diag.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))

import gsr_runtime as rt  # noqa: E402
import synthetic  # noqa: E402
from diff_gaussian_rasterization import GaussianRasterizer, _RasterizeGaussians  # noqa: E402
from oracle import ref_driver, rigid_body_port  # noqa: E402


def bits_equal(a, b):
    return (a.contiguous().view(torch.int32) == b.contiguous().view(torch.int32))


def stat(name, a, b, exact=False, log=None):
    a = a.detach()
    b = b.detach()
    if a.dtype.is_floating_point:
        eq = bits_equal(a.float(), b.float())
        nm = int((~eq).sum())
        diff = (a.double() - b.double()).abs()
        mx = float(diff.max()) if diff.numel() else 0.0
        ref = float(b.double().abs().max()) if b.numel() else 0.0
        line = dict(name=name, n=a.numel(), bit_mismatch=nm, max_abs=mx, ref_max=ref)
    else:
        nm = int((a != b).sum())
        line = dict(name=name, n=a.numel(), mismatch=nm)
    print(json.dumps(line))
    if log is not None:
        log.append(line)
    return line


def run_mine(rs, sc, grad, twists=None, keep=False):
    ras = GaussianRasterizer(rs)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sc.items()}
    means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
    kw = {}
    if twists is not None:
        S = twists[0].clone().requires_grad_(True)
        th = twists[1].clone().requires_grad_(True)
        kw = dict(se3_S=S, se3_theta=th)
    color, radii = ras(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"],
                       shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"], **kw)
    (color * grad).sum().backward()
    out = dict(color=color.detach(), radii=radii, means2D_grad=means2D.grad,
               grads={k: v.grad for k, v in leaves.items()}, deformed=ras.deformed_means)
    if twists is not None:
        out["dS"], out["dtheta"] = S.grad, th.grad
    return out


def mine_intermediates(rs, sc):
    """Forward only, returning the workspaces for slicing."""
    lib = rt.load()
    P = sc["means3D"].shape[0]
    W, H = rs.image_width, rs.image_height
    dev = sc["means3D"].device
    view = rt.make_view(rs)
    geom = torch.zeros(lib.gsr_geom_bytes(P), dtype=torch.uint8, device=dev)
    img = torch.zeros(lib.gsr_image_bytes(W, H), dtype=torch.uint8, device=dev)
    radii = torch.zeros(P, dtype=torch.int32, device=dev)
    color = torch.zeros((3, H, W), device=dev)
    mb = rt.pinned_u32(dev)
    d = rt.gsr_deform()
    st = rt.stream_ptr(dev)
    rt.check(lib.gsr_forward_preprocess(view, P, 16, rt.ptr(sc["means3D"]), rt.ptr(sc["scales"]), rt.ptr(sc["rotations"]),
                                        rt.ptr(sc["opacities"]), rt.ptr(sc["shs"]), None, None, d, None, rt.ptr(radii),
                                        rt.ptr(geom), geom.numel(), mb.data_ptr(), 1, st))
    R = int(mb.item())
    binning = torch.zeros(lib.gsr_binning_bytes(R, W, H), dtype=torch.uint8, device=dev)
    rt.check(lib.gsr_forward_render(view, P, R, rt.ptr(radii), rt.ptr(geom), rt.ptr(binning), binning.numel(),
                                    rt.ptr(img), rt.ptr(color), 1, st))
    torch.cuda.synchronize()
    gl, il, bl = rt.geom_layout(P), rt.image_layout(W, H), rt.binning_layout(R, W, H)
    tiles = ((W + 15) // 16) * ((H + 15) // 16)

    def sl(buf, off, n, dt):
        es = torch.empty((), dtype=dt).element_size()
        return buf[off:off + n * es].view(dt)
    recs = sl(geom, gl["recs"], 12 * P, torch.float32).view(P, 12)
    return dict(R=R, radii=radii, color=color,
                depths=sl(geom, gl["depths"], P, torch.float32),
                tiles=sl(geom, gl["tiles_touched"], P, torch.int32),
                
                cov3D=sl(geom, gl["cov3D"], 6 * P, torch.float32).view(P, 6),
                clamped=sl(geom, gl["clamped"], P, torch.uint8),
                means2D=recs[:, 0:2], conic_opacity=torch.cat([recs[:, 2:5], recs[:, 5:6]], 1),
                rgb=recs[:, 6:9],
                keys=sl(binning, bl["keys_sorted"], R, torch.int64),
                point_list=sl(binning, bl["point_list"], R, torch.int32),
                final_T=sl(img, il["final_T"], W * H, torch.float32),
                n_contrib=sl(img, il["n_contrib"], W * H, torch.int32),
                ranges=sl(img, il["ranges"], 2 * tiles, torch.int32).view(tiles, 2))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=100000)
    ap.add_argument("--W", type=int, default=640)
    ap.add_argument("--H", type=int, default=360)
    ap.add_argument("--scale_mult", type=float, default=1.0)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = "cuda"
    log = []
    print("device:", torch.cuda.get_device_name(0), "lib:", rt.lib_path())
    sc = synthetic.make_scene(args.P, seed=0, device=dev, scale_mult=args.scale_mult)
    cam = synthetic.make_camera(0, 1, args.W, args.H, device=dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    rs = synthetic.raster_settings(cam, bg)
    grad = synthetic.make_image_grad(args.W, args.H, device=dev)
    P, W, H = args.P, args.W, args.H

    # ---- reference ----
    ref_f = ref_driver.forward(rs, sc["means3D"], sc["opacities"], shs=sc["shs"], scales=sc["scales"],
                               rotations=sc["rotations"])
    torch.cuda.synchronize()
    R = ref_f["num_rendered"]
    rg, rb, ri = ref_driver.slice_geom(ref_f["geom"], P), ref_driver.slice_binning(ref_f["binning"], R), \
        ref_driver.slice_img(ref_f["img"], W, H)
    ref_b = ref_driver.backward(rs, ref_f, grad, sc["means3D"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    torch.cuda.synchronize()
    vis = ref_f["radii"] > 0
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    print(json.dumps(dict(P=P, W=W, H=H, R_ref=R, visible=int(vis.sum()), n_contrib_sum=int(ri["n_contrib"].sum()))))

    # ---- mine: intermediates ----
    mi = mine_intermediates(rs, sc)
    print(json.dumps(dict(R_mine=mi["R"])))
    stat("radii", mi["radii"], ref_f["radii"], log=log)
    stat("tiles_touched", mi["tiles"], rg["tiles_touched"], log=log)
    stat("depths[vis]", mi["depths"][vis], rg["depths"][vis], log=log)
    stat("means2D[vis]", mi["means2D"][vis], rg["means2D"][vis], log=log)
    stat("cov3D[z>0.2]", mi["cov3D"][rg["depths"] != 0] if False else mi["cov3D"][vis], rg["cov3D"][vis], log=log)
    stat("conic_opacity[vis]", mi["conic_opacity"][vis], rg["conic_opacity"][vis], log=log)
    stat("rgb[vis]", mi["rgb"][vis], rg["rgb"][vis], log=log)
    cl_ref = (rg["clamped"][:, 0].int() | (rg["clamped"][:, 1].int() << 1) | (rg["clamped"][:, 2].int() << 2))
    stat("clamped[vis]", mi["clamped"][vis].int(), cl_ref[vis], log=log)
    if mi["R"] == R:
        stat("keys_sorted", mi["keys"], rb["point_list_keys"], log=log)
        stat("point_list", mi["point_list"], rb["point_list"], log=log)
    stat("ranges", mi["ranges"], ri["ranges"][:tiles], log=log)
    stat("n_contrib", mi["n_contrib"], ri["n_contrib"], log=log)
    stat("final_T", mi["final_T"], ri["accum_alpha"], log=log)
    stat("color", mi["color"], ref_f["color"], log=log)

    # ---- mine: autograd path ----
    m = run_mine(rs, sc, grad)
    torch.cuda.synchronize()
    stat("color(autograd)", m["color"], ref_f["color"], log=log)
    for k_m, k_r in (("means3D", "means3D"), ("opacities", "opacities"), ("shs", "shs"), ("scales", "scales"),
                     ("rotations", "rotations")):
        a, b = m["grads"][k_m], ref_b[k_r]
        line = stat("grad_" + k_m, a, b.view_as(a), log=log)
        rel = float((a.double() - b.view_as(a).double()).abs().max() / (b.double().abs().max() + 1e-30))
        print(json.dumps(dict(name="grad_" + k_m, rel_to_max=rel)))
        log.append(dict(name="grad_" + k_m + "_rel", rel_to_max=rel))
    a, b = m["means2D_grad"], ref_b["means2D"]
    stat("grad_means2D", a, b, log=log)
    print(json.dumps(dict(name="grad_means2D", rel_to_max=float((a - b).abs().max() / (b.abs().max() + 1e-30)))))

    # reference run-to-run nondeterminism (atomics) for scale
    ref_b2 = ref_driver.backward(rs, ref_f, grad, sc["means3D"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    print(json.dumps(dict(name="ref_vs_ref grad_means3D", rel_to_max=float(
        (ref_b2["means3D"] - ref_b["means3D"]).abs().max() / ref_b["means3D"].abs().max()))))

    # ---- SE3 fused vs port ----
    S, th = synthetic.make_twists(P, device=dev)
    mt = run_mine(rs, sc, grad, twists=(S, th))
    x = sc["means3D"].clone().requires_grad_(True)
    S_ = S.clone().requires_grad_(True)
    th_ = th.clone().requires_grad_(True)
    y = rigid_body_port.deform_points(x, S_, th_)
    stat("se3 deformed means", mt["deformed"], y, log=log)
    # feed OUR deformed means to the reference for bit-exact downstream comparison
    yd = mt["deformed"].detach()
    rf2 = ref_driver.forward(rs, yd, sc["opacities"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    stat("se3 radii", mt["radii"], rf2["radii"], log=log)
    stat("se3 color", mt["color"], rf2["color"], log=log)
    rb2 = ref_driver.backward(rs, rf2, grad, yd, shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    (y * rb2["means3D"]).sum().backward()
    for nm, a, b in (("se3 dx", mt["grads"]["means3D"], x.grad), ("se3 dS", mt["dS"], S_.grad), ("se3 dtheta", mt["dtheta"], th_.grad)):
        stat(nm, a, b, log=log)
        print(json.dumps(dict(name=nm, rel_to_max=float((a - b).abs().max() / (b.abs().max() + 1e-30)))))

    # ---- kNN ----
    from simple_knn._C import distCUDA2
    pts = sc["means3D"]
    d_m = distCUDA2(pts)
    d_r = ref_driver.dist_cuda2(pts)
    stat("distCUDA2", d_m, d_r, log=log)
    print(json.dumps(dict(name="distCUDA2", rel=float(((d_m - d_r).abs() / d_r.clamp_min(1e-30)).max()))))

    # ---- timing ----
    if args.time:
        def timeit(fn, n=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(n):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            return ts[len(ts) // 2]

        def ref_fb():
            f = ref_driver.forward(rs, sc["means3D"], sc["opacities"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
            ref_driver.backward(rs, f, grad, sc["means3D"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])

        def ref_f_only():
            ref_driver.forward(rs, sc["means3D"], sc["opacities"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])

        def mine_fb():
            run_mine(rs, sc, grad)

        def mine_f_only():
            with torch.no_grad():
                GaussianRasterizer(rs)(means3D=sc["means3D"], means2D=None, opacities=sc["opacities"], shs=sc["shs"],
                                       scales=sc["scales"], rotations=sc["rotations"])
        tm = dict(ref_fwd_ms=timeit(ref_f_only), ref_fwd_bwd_ms=timeit(ref_fb), mine_fwd_ms=timeit(mine_f_only),
                  mine_fwd_bwd_ms=timeit(mine_fb))
        print(json.dumps(tm))
        log.append(tm)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(log, f, indent=1)


if __name__ == "__main__":
    main()
