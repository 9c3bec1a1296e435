#!/usr/bin/env python
"""Workload counters of the blend stage (gsr_debug_blend_stats) for a synthetic view."""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))
import gsr_runtime as rt, synthetic
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gpu_diag import mine_intermediates

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=1000000); ap.add_argument("--W", type=int, default=1920); ap.add_argument("--H", type=int, default=1080)
    a = ap.parse_args()
    lib = rt.load(); dev = "cuda"
    sc = synthetic.make_scene(a.P, device=dev); cam = synthetic.make_camera(0, 1, a.W, a.H, device=dev)
    rs = synthetic.raster_settings(cam, torch.zeros(3, device=dev))
    # rerun forward keeping the workspaces
    P, W, H = a.P, a.W, a.H
    view = rt.make_view(rs)
    geom = torch.zeros(lib.gsr_geom_bytes(P), dtype=torch.uint8, device=dev); img = torch.zeros(lib.gsr_image_bytes(W, H), dtype=torch.uint8, device=dev)
    radii = torch.zeros(P, dtype=torch.int32, device=dev); color = torch.zeros((3, H, W), device=dev); mb = rt.pinned_u32(dev); st = rt.stream_ptr(dev)
    rt.check(lib.gsr_forward_preprocess(view, P, 16, rt.ptr(sc["means3D"]), rt.ptr(sc["scales"]), rt.ptr(sc["rotations"]), rt.ptr(sc["opacities"]), rt.ptr(sc["shs"]), None, None, rt.gsr_deform(), None, rt.ptr(radii), rt.ptr(geom), geom.numel(), mb.data_ptr(), 0, st))
    R = int(mb.item())
    binning = torch.zeros(lib.gsr_binning_bytes(R, W, H), dtype=torch.uint8, device=dev)
    rt.check(lib.gsr_forward_render(view, P, R, rt.ptr(radii), rt.ptr(geom), rt.ptr(binning), binning.numel(), rt.ptr(img), rt.ptr(color), 1, st))
    out = torch.zeros(8, dtype=torch.int64, device=dev)
    rt.check(lib.gsr_debug_blend_stats(view, P, R, rt.ptr(geom), rt.ptr(binning), rt.ptr(img), rt.ptr(out), st))
    torch.cuda.synchronize()
    names = ["staged_entries", "cull_survivors", "warp_iterations", "warp_iter_any_power", "warp_iter_any_blend", "pixel_pairs_evaluated", "pixel_pairs_blended", "list_entries_total"]
    d = dict(zip(names, out.tolist())); d.update(P=P, W=W, H=H, R=R, visible=int((radii > 0).sum()))
    print(json.dumps(d))
    out = torch.zeros(32, dtype=torch.int64, device=dev)
    rt.check(lib.gsr_debug_blend_group_stats(view, P, R, rt.ptr(geom), rt.ptr(binning), rt.ptr(img), rt.ptr(out), st))
    torch.cuda.synchronize()
    o = out.tolist()
    cfg = ["1x(8x8)", "2x(8x4)", "4x(4x4)", "4x(8x2)", "2x(4x8)", "8x(4x2)"]
    base = o[0]
    for c, nm in enumerate(cfg):
        print("%-8s sum_of_groups %11d  lockstep/batch %11d (%.3f of now)  ideal queues %11d (%.3f)  batches %d" % (
            nm, o[4 * c], o[4 * c + 1], o[4 * c + 1] / base, o[4 * c + 2], o[4 * c + 2] / base, o[4 * c + 3]))
    print("lane-iterations with a blending pixel %d (%.3f of 32 x iterations), pixels blended %d, evaluated %d" % (
        o[24], o[24] / (32.0 * base), o[25], o[26]))

if __name__ == "__main__":
    main()
