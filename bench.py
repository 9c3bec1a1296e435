#!/usr/bin/env python
"""Benchmark of the deformable Gaussian-splatting hot path (BASELINE.json metric:
fwd+bwd ms/view and Gaussian-views/s, 1 M Gaussians @ 1080p, 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one training-style pass over a batch of V views (default 8 per GPU, so
8 GPUs render the 64 cameras of config C5): for every view, SE3 exp-map deformation
of all Gaussians + rasterize forward + dL/dimage injection + backward down to the
leaf parameters and twists; per-Gaussian gradients are summed over the views of the
GPU and (N > 1) all-reduced over NCCL.  Synthetic scene/cameras.

Arms
  ours       GaussianRasterizer (drop-in API) -> libgsr_b200.so, SE3 fused in-kernel.
  reference  the reference's own CUDA rasterizer rebuilt unmodified (oracle/_ref) +
             its torch rigid_body op graph (oracle/rigid_body_port.py) on the same
             GPU: the comparator BASELINE.json's north_star names.  The reference has
             no CPU rasterizer; its torch-CPU deformation stage is reported as
             `cpu_baseline`.

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))

PARAM_KEYS = ("means3D", "scales", "rotations", "opacities", "shs")
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` at the default workload (profiles/)
NCU_TRAFFIC = {"preprocess_bwd": 581.1e6, "preprocess_fwd": 233.9e6, "blend_bwd": 110.1e6, "blend_fwd": 54.0e6,   # profiles/r02_ncu_full_metrics.csv
               # per launch = per 8-view step (profiles/r03_ncu_batched_kernels_metrics.csv)
               "preprocess_bwd_batched": 1484.9e6, "preprocess_fwd_batched": 753.9e6}


# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
CONFIGS = {
    # BASELINE.json configs[1..4]; C2 is the configuration the metric is quoted on (the default)
    "C2": dict(P=1000000, W=1920, H=1080, views=8, bodies=False,
               what="C2: 1 M Gaussians, SH degree 3, 1920x1080, per-Gaussian SE3 exp-map deform + rasterize fwd+bwd"),
    "C3": dict(P=500000, W=800, H=800, views=4, bodies=True,
               what="C3: 500 k Gaussians in 64 rigid bodies (body twist table of frame f of 300, a new frame every step), "
                    "800x800, fused SE3 + rasterize fwd+bwd"),
    "C4": dict(P=6000000, W=3840, H=2160, views=1, bodies=False,
               what="C4: 6 M Gaussians at 3840x2160 (tile/sort-heavy stress), per-Gaussian SE3 + rasterize fwd+bwd"),
}


def build_workload(args, rank, world, dev):
    import synthetic
    sc = synthetic.make_scene(args.P, seed=0, device="cpu")
    host = {k: sc[k].pin_memory() for k in PARAM_KEYS}
    if args.bodies:
        body_id, S, theta = synthetic.make_bodies(sc["means3D"], frame=150, frames=300, device="cpu")
        args.body_id = body_id.to(dev)
    else:
        S, theta = synthetic.make_twists(args.P, seed=2, device="cpu")
        args.body_id = None
    host["S"], host["theta"] = S.pin_memory(), theta.pin_memory()
    total_views = args.views * world
    K = max(total_views, 64)
    cams = [synthetic.make_camera(rank * args.views + v, K, args.W, args.H, device=dev) for v in range(args.views)]
    bg = torch.zeros(3, device=dev)
    grad = synthetic.make_image_grad(args.W, args.H, seed=1, device=dev)
    return host, cams, bg, grad


def to_device(host, dev):
    return {k: v.to(dev, non_blocking=True).requires_grad_(True) for k, v in host.items()}


def zero_grads(leaves):
    for t in leaves.values():
        t.grad = None


_FLAT = {}


def flat_buffer(leaves):
    """One flat fp32 gradient buffer for all leaves (view_parallel.FlatGradBuffer); `.grad`s are
    views into it, the backward kernel accumulates into it and the all-reduce runs on it."""
    import view_parallel
    key = tuple(id(t) for t in leaves.values())
    if _FLAT.get("key") != key:
        _FLAT["key"] = key
        shapes = [tuple(t.shape) for t in leaves.values()]
        if _FLAT.get("shapes") == shapes and _FLAT["buf"].flat.device == next(iter(leaves.values())).device:
            _FLAT["buf"].rebind(list(leaves.values()))          # a step's freshly staged parameters: reuse the buffer
        else:
            _FLAT["buf"] = view_parallel.FlatGradBuffer(list(leaves.values()))
            _FLAT["shapes"] = shapes
    return _FLAT["buf"]


_STREAMS = {}
_SETTINGS = {}


def view_settings(cam, bg):
    """GaussianRasterizationSettings of a camera, built once per (camera, background) - both arms: the cameras of a run are
    fixed, and a training loop would not rebuild them every step either (8 x 19 us of host time per step otherwise)."""
    import synthetic
    key = (id(cam), id(bg))
    hit = _SETTINGS.get(key)
    if hit is None or hit[0] is not cam or hit[1] is not bg:
        hit = (cam, bg, synthetic.raster_settings(cam, bg, sh_degree=3))
        _SETTINGS[key] = hit
    return hit[2]


def step_ours(leaves, cams, bg, grad, args):
    """One view-batched step: forward+backward of every local view, gradients summed in the flat buffer.
    The views of a step are independent, so they are issued round-robin on `--streams` CUDA streams:
    one view's latency-bound binning kernels, launch gaps and kernel tails overlap another view's
    issue-bound blend kernels (view_parallel.render_views)."""
    import synthetic
    import view_parallel
    from diff_gaussian_rasterization import GaussianRasterizer, GaussianBackwardBatch
    buf = flat_buffer(leaves)
    buf.zero_()
    sinks = {"means3D": leaves["means3D"].grad, "opacities": leaves["opacities"].grad, "shs": leaves["shs"].grad,
             "scales": leaves["scales"].grad, "rotations": leaves["rotations"].grad,
             "se3_S": leaves["S"].grad, "se3_theta": leaves["theta"].grad}
    no_deform = getattr(args, "no_deform", False)
    if no_deform:
        sinks = {k: v for k, v in sinks.items() if not k.startswith("se3_")}
    # the per-Gaussian half of the backward runs once for all views of the step (--batched-backward 0: once per view)
    batch = None
    overlap = False
    many = len(cams) > 1                  # one view per step: nothing to share, the per-view kernels are the faster ones
    if getattr(args, "batched_backward", 1) and many:
        # N > 1: the batched kernel runs over 4 ranges of Gaussians and each range's all-reduce starts as soon as the range
        # is enqueued (NCCL's stream), so only the last range's collective is exposed
        world = int(os.environ.get("WORLD_SIZE", "1"))
        overlap = world > 1 and bool(getattr(args, "overlap_allreduce", 0))
        P = leaves["means3D"].shape[0]
        batch = GaussianBackwardBatch(sinks, chunks=4 if overlap else 1,
                                      after_chunk=(lambda first, count: buf.all_reduce_rows(first, count, P)) if overlap else None)
        sinks = batch
    args._reduced_in_step = overlap

    settings = [view_settings(cam, bg) for cam in cams]
    # the per-Gaussian half of the forward once for all views of the step (--batched-forward 0: once per view)
    fwd = None
    if getattr(args, "batched_forward", 1) and many:
        from diff_gaussian_rasterization import GaussianForwardBatch
        extra = {} if no_deform else dict(se3_S=leaves["S"], se3_theta=leaves["theta"], body_id=args.body_id)
        fwd = GaussianForwardBatch(settings, means3D=leaves["means3D"], opacities=leaves["opacities"], shs=leaves["shs"],
                                   scales=leaves["scales"], rotations=leaves["rotations"], **extra)

    def render_view(i):
        rs = settings[i]
        ras = GaussianRasterizer(rs)
        means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
        prepared = fwd.prepared(i) if fwd is not None else None
        if no_deform:
            snk = sinks
            color, radii = ras(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"],
                               shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"], accumulate_grads=snk,
                               prepared=prepared)
        else:
            color, radii = ras(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"],
                               shs=leaves["shs"], scales=leaves["scales"], rotations=leaves["rotations"],
                               se3_S=leaves["S"], se3_theta=leaves["theta"], body_id=args.body_id, accumulate_grads=sinks,
                               prepared=prepared)
        # dL/dimage = grad is injected directly (as the reference arm does below): one reduction for the loss value, no
        # autograd graph through a multiply
        loss = torch.dot(color.detach().reshape(-1), grad.reshape(-1))
        color.backward(grad)
        return loss

    total = view_parallel.render_views(render_view, range(len(cams)), num_streams=args.streams, batch=batch)
    if overlap:
        buf.wait()
    return total


def _ref_deform(leaves, args):
    """The reference's torch op graph for the deformation: exp_se3 + the apply recipe (rigid bodies: the body's twist
    gathered per Gaussian first, which is what the reference API shape S[N,6], theta[N] asks for)."""
    from oracle import rigid_body_port
    if args.body_id is not None:
        idx = args.body_id.long()
        return rigid_body_port.deform_points(leaves["means3D"], leaves["S"][idx], leaves["theta"][idx])
    return rigid_body_port.deform_points(leaves["means3D"], leaves["S"], leaves["theta"])


def step_reference(leaves, cams, bg, grad, args):
    """Reference arm: torch rigid_body op graph (autograd) + the reference's CUDA
    rasterizer driven the way its own autograd.Function drives `_C`."""
    import synthetic
    from oracle import ref_driver, rigid_body_port
    loss_total = None
    for cam in cams:
        rs = view_settings(cam, bg)
        no_deform = getattr(args, "no_deform", False)
        y = leaves["means3D"] if no_deform else _ref_deform(leaves, args)
        yd = y.detach()
        f = ref_driver.forward(rs, yd, leaves["opacities"].detach(), shs=leaves["shs"].detach(),
                               scales=leaves["scales"].detach(), rotations=leaves["rotations"].detach())
        loss = torch.dot(f["color"].reshape(-1), grad.reshape(-1))
        g = ref_driver.backward(rs, f, grad, yd, shs=leaves["shs"].detach(), scales=leaves["scales"].detach(),
                                rotations=leaves["rotations"].detach())
        if no_deform:
            y.grad = g["means3D"] if y.grad is None else y.grad + g["means3D"]
        else:
            y.backward(g["means3D"])
        for k, gk in (("opacities", "opacities"), ("shs", "shs"), ("scales", "scales"), ("rotations", "rotations")):
            t = leaves[k]
            t.grad = g[gk].view_as(t) if t.grad is None else t.grad + g[gk].view_as(t)
        loss_total = loss.detach() if loss_total is None else loss_total + loss.detach()
    return loss_total


def counters_ours(leaves, cam, bg, args):
    """Data-dependent counters of one view, read from the library's own workspaces after a
    forward through the raw C ABI: visible Gaussians, num_rendered R, the sum of n_contrib (list entries every pixel
    walks before it terminates; bit-equal to the reference's buffer, tests/test_gpu_parity.py), and the blend
    stage's pair counts from gsr_debug_blend_stats (a replay of the forward loop)."""
    import gsr_runtime as rt
    import synthetic
    lib = rt.load()
    dev = leaves["means3D"].device
    P, W, H = args.P, args.W, args.H
    rs = view_settings(cam, bg)
    view = rt.make_view(rs)
    d = {k: v.detach() for k, v in leaves.items()}
    geom = torch.empty(lib.gsr_geom_bytes(P), dtype=torch.uint8, device=dev)
    img = torch.empty(lib.gsr_image_bytes(W, H), dtype=torch.uint8, device=dev)
    radii = torch.empty(P, dtype=torch.int32, device=dev)
    color = torch.empty((3, H, W), device=dev)
    means_def = torch.empty((P, 3), device=dev)
    mb = rt.pinned_u32(dev)
    st = rt.stream_ptr(dev)
    df = rt.gsr_deform()
    df.mode = rt.DEFORM_RIGID_BODIES if args.body_id is not None else rt.DEFORM_PER_GAUSSIAN
    df.num_bodies = int(d["S"].shape[0]) if args.body_id is not None else 0
    df.S, df.theta = d["S"].data_ptr(), d["theta"].data_ptr()
    df.body_id = args.body_id.data_ptr() if args.body_id is not None else None
    rt.check(lib.gsr_forward_preprocess(view, P, 16, rt.ptr(d["means3D"]), rt.ptr(d["scales"]), rt.ptr(d["rotations"]),
                                        rt.ptr(d["opacities"]), rt.ptr(d["shs"]), None, None, df, rt.ptr(means_def),
                                        rt.ptr(radii), rt.ptr(geom), geom.numel(), mb.data_ptr(), 0, st))
    R = int(mb.item()) & 0xFFFFFFFF
    binning = torch.empty(lib.gsr_binning_bytes(R, W, H), dtype=torch.uint8, device=dev)
    rt.check(lib.gsr_forward_render(view, P, R, rt.ptr(radii), rt.ptr(geom), rt.ptr(binning), binning.numel(),
                                    rt.ptr(img), rt.ptr(color), 0, st))
    out = torch.zeros(8, dtype=torch.int64, device=dev)
    rt.check(lib.gsr_debug_blend_stats(view, P, R, rt.ptr(geom), rt.ptr(binning), rt.ptr(img), rt.ptr(out), st))
    torch.cuda.synchronize()
    il = rt.image_layout(W, H)
    n_contrib = img[il["n_contrib"]:il["n_contrib"] + 4 * W * H].view(torch.int32)
    o = out.tolist()
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    return {"P": P, "P_visible": int((radii > 0).sum()), "R": R, "sum_n_contrib": int(n_contrib.long().sum()),
            "pairs_evaluated": o[5], "pairs_blended": o[6], "list_entries_after_cull": o[1], "tiles": tiles,
            "sort_passes_reference_scheme": (32 + max(tiles, 1).bit_length() + 7) // 8,
            "source": "libgsr_b200 workspaces of the step's last view (n_contrib / R / radii are bit-equal to the reference's, "
                      "see tests); pairs_* from gsr_debug_blend_stats"}


def counters_reference(leaves, cam, bg, args):
    """The same counters from the reference rasterizer's own buffers (oracle/_ref)."""
    import synthetic
    from oracle import ref_driver
    rs = view_settings(cam, bg)
    with torch.no_grad():
        y = _ref_deform(leaves, args)
        f = ref_driver.forward(rs, y, leaves["opacities"].detach(), shs=leaves["shs"].detach(),
                               scales=leaves["scales"].detach(), rotations=leaves["rotations"].detach())
    im = ref_driver.slice_img(f["img"], args.W, args.H)
    return {"P": args.P, "P_visible": int((f["radii"] > 0).sum()), "R": f["num_rendered"],
            "sum_n_contrib": int(im["n_contrib"].long().sum()),
            "source": "reference rasterizer buffers (radii, num_rendered, ImageState::n_contrib) of the step's last view"}


def allreduce_grads(leaves, world):
    if world == 1:
        return
    import torch.distributed as dist
    works = [dist.all_reduce(t.grad, async_op=True) for t in leaves.values() if t.grad is not None]
    for w in works:
        w.wait()


# learning rates of arguments/__init__.py:73-80 in the group order of scene/gaussian_model.py:840-860 (+ the twists),
# scaled by 1e-3: the reference applies them to PRE-activation tensors (log-scales, opacity logits); this step
# optimises the activated tensors the rasterizer consumes, where the unscaled rates would double every Gaussian's
# extent per step and the workload (num_rendered) would drift from step to step instead of staying C5's.
TRAIN_LR_SCALE = 1e-3
TRAIN_LRS = {"means3D": 0.00016, "shs": 0.0025 / 20.0, "opacities": 0.05, "scales": 0.005, "rotations": 0.001,
             "S": 0.00016, "theta": 0.00016}


def make_train_state(host, dev, impl, n_views, W, H):
    leaves = {k: v.to(dev).requires_grad_(True) for k, v in host.items()}
    groups = [{"params": [leaves[k]], "lr": TRAIN_LRS[k] * TRAIN_LR_SCALE, "name": k} for k in leaves]
    if impl == "ours":
        import fused_adam
        opt = fused_adam.FusedAdam(groups, lr=0.0, eps=1e-15)        # re-homes params + grads into flat buffers
    else:
        opt = torch.optim.Adam(groups, lr=0.0, eps=1e-15)            # scene/gaussian_model.py:846
    g = torch.Generator().manual_seed(7)
    targets = [torch.rand((3, H, W), generator=g).to(dev) for _ in range(n_views)]
    return leaves, opt, targets


def train_step(leaves, opt, targets, cams, bg, args, world):
    """One whole optimizer step over the rank's views (C5): render, loss, backward, all-reduce, Adam."""
    import synthetic
    if args.impl == "ours":
        import loss_utils
        import view_parallel
        from diff_gaussian_rasterization import GaussianRasterizer
        opt.zero_grad()
        sinks = {"means3D": leaves["means3D"].grad, "opacities": leaves["opacities"].grad, "shs": leaves["shs"].grad,
                 "scales": leaves["scales"].grad, "rotations": leaves["rotations"].grad,
                 "se3_S": leaves["S"].grad, "se3_theta": leaves["theta"].grad}
        batch = None
        if getattr(args, "batched_backward", 1):
            from diff_gaussian_rasterization import GaussianBackwardBatch
            batch = sinks = GaussianBackwardBatch(sinks)

        settings = [view_settings(cam, bg) for cam in cams]
        fwd = None
        if getattr(args, "batched_forward", 1):
            from diff_gaussian_rasterization import GaussianForwardBatch
            fwd = GaussianForwardBatch(settings, means3D=leaves["means3D"], opacities=leaves["opacities"], shs=leaves["shs"],
                                       scales=leaves["scales"], rotations=leaves["rotations"], se3_S=leaves["S"],
                                       se3_theta=leaves["theta"], body_id=args.body_id)

        def render_view(i):
            ras = GaussianRasterizer(settings[i])
            means2D = torch.zeros_like(leaves["means3D"], requires_grad=True)
            color, _ = ras(means3D=leaves["means3D"], means2D=means2D, opacities=leaves["opacities"], shs=leaves["shs"],
                           scales=leaves["scales"], rotations=leaves["rotations"], se3_S=leaves["S"],
                           se3_theta=leaves["theta"], body_id=args.body_id, accumulate_grads=sinks,
                           prepared=fwd.prepared(i) if fwd is not None else None)
            loss = loss_utils.l1_ssim_loss(color, targets[i], 0.2)
            loss.backward()
            return loss.detach()

        total = view_parallel.render_views(render_view, range(len(cams)), num_streams=args.streams, batch=batch)
        opt.all_reduce_and_step(chunks=8)          # N > 1: all-reduce in 8 pieces, Adam on each piece as it arrives
        return total
    from oracle import loss_port, ref_driver, rigid_body_port
    opt.zero_grad(set_to_none=True)
    total = None
    for i, cam in enumerate(cams):
        rs = view_settings(cam, bg)
        y = _ref_deform(leaves, args)
        yd = y.detach()
        f = ref_driver.forward(rs, yd, leaves["opacities"].detach(), shs=leaves["shs"].detach(),
                               scales=leaves["scales"].detach(), rotations=leaves["rotations"].detach())
        img = f["color"].detach().requires_grad_(True)
        loss = loss_port.training_loss(img, targets[i], 0.2)           # utils/loss_utils.py ops, train.py:529
        loss.backward()
        g = ref_driver.backward(rs, f, img.grad, yd, shs=leaves["shs"].detach(), scales=leaves["scales"].detach(),
                                rotations=leaves["rotations"].detach())
        y.backward(g["means3D"])
        for k in ("opacities", "shs", "scales", "rotations"):
            t = leaves[k]
            t.grad = g[k].view_as(t) if t.grad is None else t.grad + g[k].view_as(t)
        total = loss.detach() if total is None else total + loss.detach()
    allreduce_grads(leaves, world)
    opt.step()
    return total


def deform_network_timing(impl, P, dev, steps=3):
    """SURVEY 8f row f1 beside the headline: the deformation network DirectTemporalNeRF (scene/gaussian_model.py:242-316)
    that the reference's render() queries for every view, forward and forward+backward over all P Gaussians.
    ours: deform_mlp.DirectTemporalNeRF (tcgen05 hi/lo-split TF32 GEMMs); reference: the same network as torch fp32
    linears (oracle/deform_mlp_port.DeformMLP, bit-identical to the reference class - tests/test_oracle_cpu.py)."""
    torch.manual_seed(0)
    if impl == "ours":
        import deform_mlp
        net = deform_mlp.DirectTemporalNeRF().to(dev)
    else:
        from oracle import deform_mlp_port
        net = deform_mlp_port.DeformMLP().to(dev)
    g = torch.Generator().manual_seed(1)
    x0 = ((torch.rand((P, 3), generator=g) * 2 - 1) * 1.3).to(dev)
    ts = torch.full((P, 1), 0.4, device=dev)
    proj = [torch.randn((P, c), generator=g).to(dev) for c in (3, 3, 4, 48)]

    def fwd():
        with torch.no_grad():
            net(x0, ts, 5000)

    def fwd_bwd():
        x = x0.clone().requires_grad_(True)
        for p in net.parameters():
            p.grad = None
        torch.autograd.backward(net(x, ts, 5000), proj)

    def timed(fn):
        fn(); fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / steps
    t_f, t_fb = timed(fwd), timed(fwd_bwd)
    # ... and with the activation glue of render() (gaussian_renderer/__init__.py:79,116,122,140) between the network and the
    # rasterizer: ours = deform_mlp.deform_glue (one kernel each way), reference = the torch ops render() uses
    mk = lambda *sh: torch.randn(*sh, generator=g).to(dev).requires_grad_(True)
    scaling, rotation, f_dc, f_rest = mk(P, 3), mk(P, 4), mk(P, 1, 3), mk(P, 15, 3)
    gout = [torch.randn(sh, generator=g).to(dev) for sh in ((P, 3), (P, 3), (P, 4), (P, 16, 3))]

    def fwd_bwd_glue():
        x = x0.clone().requires_grad_(True)
        for p in list(net.parameters()) + [scaling, rotation, f_dc, f_rest]:
            p.grad = None
        if impl == "ours":
            outs = deform_mlp.deform_glue(net.heads(x, ts, 5000), x, scaling, rotation, f_dc, f_rest)
        else:
            dx, ds, dr, dsh = net(x, ts, 5000)
            outs = (x + dx, torch.exp(scaling + ds), torch.nn.functional.normalize(rotation + dr),
                    torch.cat((f_dc, f_rest), dim=1) + dsh.reshape(-1, 16, 3))
        torch.autograd.backward(outs, gout)
    t_fbg = timed(fwd_bwd_glue)
    flops = 2.0 * P * (64 * 256 + 6 * 256 * 256 + 320 * 256 + 256 * 58)
    out = {"P": P, "fwd_ms": round(t_f, 3), "fwd_bwd_ms": round(t_fb, 3), "fwd_bwd_with_activation_glue_ms": round(t_fbg, 3),
           "fwd_fp32_equivalent_TFLOPs": round(flops / t_f / 1e9, 1),
           "what": "DirectTemporalNeRF (84 -> 8 x 256 ReLU, skip at 4 -> heads 3/3/4/48) over all P Gaussians, one time value"}
    if impl == "ours":
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        tf32_peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0
        out["roofline"] = {"bound": "tensor", "achieved": round(3 * flops / t_f / 1e9, 1), "peak": tf32_peak, "unit": "TFLOP/s",
                           "frac": round(3 * flops / t_f / 1e9 / tf32_peak, 4),
                           "traffic": 2.0006e9 if P == 1000000 else None,      # per hidden-layer GEMM launch (ncu, profiles/r02_ncu_gemm_metrics.csv) = its algorithmic 2 x P x 256 x 4 B
                           "note": "forward; achieved counts the 3 TF32 tensor-core products issued per fp32-grade product; peak = half of "
                                   "the measured sustained bf16 cuBLAS rate (TF32 runs at half the bf16 rate on tcgen05)"}
    return out


def cpu_baseline(n=100000, iters=30):
    """Reference torch-CPU deformation stage (config C1): exp_se3 + apply, fwd+bwd."""
    import synthetic
    from oracle import rigid_body_port
    torch.set_num_threads(os.cpu_count() or 1)
    x0 = synthetic.make_scene(n, seed=0)["means3D"]
    S0, th0 = synthetic.make_twists(n, seed=2)
    gy = torch.randn_like(x0)
    ts = []
    for it in range(5 + iters):
        x, S, th = x0.clone().requires_grad_(True), S0.clone().requires_grad_(True), th0.clone().requires_grad_(True)
        t0 = time.perf_counter()
        y = rigid_body_port.deform_points(x, S, th)
        (y * gy).sum().backward()
        dt = time.perf_counter() - t0
        if it >= 5:
            ts.append(dt)
    ts.sort()
    med = ts[len(ts) // 2]
    return {"value": n / med, "unit": "Gaussian-views/s (SE3 deformation stage only; the reference has no CPU rasterizer)",
            "cores": torch.get_num_threads(), "kind": "port",
            "sample": "C1: rigid_body exp_se3+apply fwd+bwd, N=%d, fp32 torch-CPU, median of %d (%.1f ms/iter)" % (n, iters, med * 1e3)}


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS),
                    help="BASELINE.json config: C2 (default, the one the metric is quoted on), C3 (500 k Gaussians in 64 rigid "
                         "bodies, 800x800; add --train for the whole training step), C4 (6 M Gaussians at 3840x2160)")
    ap.add_argument("--views", type=int, default=None, help="views per GPU per step (default: the config's)")
    ap.add_argument("--P", type=int, default=None)
    ap.add_argument("--W", type=int, default=None)
    ap.add_argument("--H", type=int, default=None)
    ap.add_argument("--streams", type=int, default=4, help="CUDA streams the views of a step are spread over (ours)")
    ap.add_argument("--batched-forward", type=int, default=1, dest="batched_forward",
                    help="ours: 1 = the per-Gaussian half of the forward (preprocess) runs once per step for all views "
                         "(GaussianForwardBatch), 0 = once per view")
    ap.add_argument("--overlap-allreduce", type=int, default=0, dest="overlap_allreduce",
                    help="ours, N > 1, with --batched-backward 1: all-reduce the gradients by ranges of Gaussians while the "
                         "batched kernel computes the next range (0: one all-reduce after the step).  Measured at N = 2: "
                         "9.59 vs 9.65 ms/step - the step's exposed cost is rank skew, not the collective - so off by default")
    ap.add_argument("--batched-backward", type=int, default=1, dest="batched_backward",
                    help="ours: 1 = the per-Gaussian half of the backward runs once per step for all views "
                         "(GaussianBackwardBatch), 0 = once per view")
    ap.add_argument("--train", action="store_true",
                    help="also time the whole C5-style training step: per-view loss 0.8 L1 + 0.2 (1 - SSIM) against a fixed "
                         "target, gradient all-reduce, Adam step (ours: fused kernels; reference: its torch ops)")
    ap.add_argument("--sync-free", type=int, default=1,
                    help="ours: 1 = diff_gaussian_rasterization.set_sync_free(True): the duplicate list is sized from the previous "
                         "steps' num_rendered x 1.25 and the forward has no host round trip; 0 = read num_rendered back in every forward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mlp", action="store_true", help="skip the deformation-network timing (SURVEY 8f f1) reported beside the headline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = CONFIGS[args.config]
    for k in ("P", "W", "H", "views"):
        if getattr(args, k) is None:
            setattr(args, k, cfg[k])
    args.bodies = cfg["bodies"]
    args.no_deform = False

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    if args.impl == "reference":
        from oracle import ref_driver
        if not ref_driver.available():
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_dgr_C.so not built (needs /root/reference)"}))
            return
        step_fn = step_reference
    else:
        import gsr_runtime as rt
        rt.load()
        step_fn = step_ours
        import diff_gaussian_rasterization as dgr
        dgr.set_sync_free(bool(args.sync_free))

    host, cams, bg, grad = build_workload(args, rank, world, dev)
    leaves = to_device(host, dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def run_step():
        if args.impl == "ours":
            loss = step_fn(leaves, cams, bg, grad, args)     # zeroes + fills the flat gradient buffer
            if world > 1 and not getattr(args, "_reduced_in_step", False):
                flat_buffer(leaves).all_reduce()                # ONE collective over 66 floats/Gaussian
        else:
            zero_grads(leaves)
            loss = step_fn(leaves, cams, bg, grad, args)
            allreduce_grads(leaves, world)
        return loss

    # ---------------- device-resident timing ----------------
    for _ in range(args.warmup):
        run_step()
    launches = 0
    if args.impl == "ours":
        rt.launch_count(reset=True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    if args.impl == "ours":
        dgr.check_sync_free()          # raises if any timed view overflowed its capacity-sized workspace (its work would be missing)
        launches = rt.launch_count(reset=True)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    views_total = args.views * world
    value = args.P * views_total / (ms_per_step * 1e-3)

    # N > 1: the step's one collective, timed on its own (ranks aligned by a barrier first): it is not overlapped with the
    # views, so this is its exposed time per step; what the step loses beyond it is rank skew (different cameras per rank)
    collective = None
    if world > 1 and args.impl == "ours":
        fb = flat_buffer(leaves)
        for _ in range(2):
            fb.all_reduce()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            fb.all_reduce()
        c1.record()
        torch.cuda.synchronize()
        cms = torch.tensor([c0.elapsed_time(c1) / 5.0], device=dev)
        torch.distributed.all_reduce(cms, op=torch.distributed.ReduceOp.MAX)
        nbytes = fb.flat.numel() * 4
        collective = {"op": "all-reduce (NCCL) of the flat fp32 gradient buffer, once per step, after the views",
                      "bytes": nbytes, "ms": round(float(cms.item()), 4), "exposed_ms_per_step": round(float(cms.item()), 4),
                      "bus_GBps": round(nbytes * 2.0 * (world - 1) / world / (float(cms.item()) * 1e-3) / 1e9, 1)}

    # ---------------- breakdown: the rasterizer alone (no deformation), and for the reference arm its torch SE3 graph ----
    # north_star's target is ">= 2x the reference CUDA rasterizer's fwd+bwd throughput": both arms print
    # `rasterizer_only_ms_per_view` for the same un-deformed scene, so the ratio can be read off the two driver records.
    breakdown = None
    if rank == 0 or world > 1:
        def timed(fn, n):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n
        nb = max(2, min(args.steps, 5))
        args.no_deform = True

        def rast_only():
            if args.impl != "ours":
                zero_grads(leaves)
            step_fn(leaves, cams, bg, grad, args)
        t_rast = timed(rast_only, nb) / args.views
        args.no_deform = False
        breakdown = {"rasterizer_only_ms_per_view": round(t_rast, 4),
                     "what": "fwd+bwd of the rasterizer on the un-deformed scene, same views, no SE3 stage"}
        if args.impl == "reference":
            def se3_only():
                zero_grads(leaves)
                y = _ref_deform(leaves, args)
                y.backward(grad_y)
            grad_y = torch.randn_like(leaves["means3D"])
            t_se3 = timed(se3_only, nb)
            breakdown["se3_graph_ms_per_view"] = round(t_se3, 4)
            breakdown["what"] += "; se3_graph = torch exp_se3 + apply recipe fwd+bwd (scene/rigid_body.py op graph)"
        if world > 1:
            torch.distributed.barrier()

    # ---------------- end-to-end: host buffers in, loss out ----------------
    # Every step copies its own inputs from pinned host memory (264 MB of parameters + twists)
    # and reads its loss back.  As any input pipeline would, the copy for step i+1 is issued on
    # a side stream while step i computes (both arms use this same loop).
    h2d = sum(t.numel() * t.element_size() for t in host.values())
    copy_stream = torch.cuda.Stream(device=dev)

    # N > 1 (ours): every replica needs the same parameters, so each byte should cross PCIe ONCE per node, not once per
    # GPU: a rank copies its 1/N slice of one flat pinned buffer and the slices are all-gathered over NVLink on the copy
    # stream (own communicator).  Eight ranks each pulling the full 264 MB per step contend on the host (round 1: e2e
    # scaling 0.85 at N = 8 against 0.91 device-resident).  The reference arm keeps its stock per-process copy.
    sharded = world > 1 and args.impl == "ours"
    if sharded:
        import view_parallel
        offs, total = view_parallel.flat_layout(list(host.values()))
        total = (total + world * 8 - 1) // (world * 8) * (world * 8)
        host_flat = torch.zeros(total, dtype=torch.float32).pin_memory()
        for (k, v), o in zip(host.items(), offs):
            host_flat[o:o + v.numel()].copy_(v.reshape(-1))
        stage_pg = torch.distributed.new_group(backend="nccl")
        shard_n = total // world
        h2d = shard_n * 4

    def stage_inputs():
        with torch.cuda.stream(copy_stream):
            if sharded:
                flat = torch.empty(total, dtype=torch.float32, device=dev)
                shard = host_flat[rank * shard_n:(rank + 1) * shard_n].to(dev, non_blocking=True)
                torch.distributed.all_gather_into_tensor(flat, shard, group=stage_pg)
                staged = {k: flat[o:o + v.numel()].view(v.shape) for (k, v), o in zip(host.items(), offs)}
            else:
                staged = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return staged, ev

    def take(staged_ev):
        staged, ev = staged_ev
        torch.cuda.current_stream(dev).wait_event(ev)
        for t in staged.values():
            t.record_stream(torch.cuda.current_stream(dev))
        return {k: v.requires_grad_(True) for k, v in staged.items()}

    nxt = stage_inputs()
    for _ in range(2):
        leaves = take(nxt)
        nxt = stage_inputs()
        float(run_step().item())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        leaves = take(nxt)                            # this step's inputs: pinned host -> device
        nxt = stage_inputs()                          # next step's copy overlaps this step's compute
        loss = run_step()
        _ = float(loss.item())                        # device -> host read of the step's result
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if args.impl == "ours":
        dgr.check_sync_free()
    if world > 1:
        torch.distributed.all_reduce(e2e_s, op=torch.distributed.ReduceOp.MAX)
    e2e_val = args.P * views_total / (float(e2e_s.item()) / args.steps)

    # ---------------- optional: the whole training step (C5) ----------------
    train = None
    if args.train:
        del leaves
        t_leaves, t_opt, t_targets = make_train_state(host, dev, args.impl, len(cams), args.W, args.H)
        for _ in range(args.warmup):
            train_step(t_leaves, t_opt, t_targets, cams, bg, args, world)
        barrier()
        te0, te1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        te0.record()
        for _ in range(args.steps):
            t_loss = train_step(t_leaves, t_opt, t_targets, cams, bg, args, world)
        te1.record()
        barrier()
        tms = torch.tensor([te0.elapsed_time(te1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX)
        t_ms = float(tms.item()) / args.steps
        train = {"ms_per_step": t_ms, "ms_per_view": t_ms / args.views,
                 "gaussian_views_per_s": args.P * views_total / (t_ms * 1e-3), "loss_last": float(t_loss),
                 "what": "C5 step: %d views/GPU x (SE3 + rasterize fwd, 0.8 L1 + 0.2 (1-SSIM) vs a fixed random target, bwd) + "
                         "gradient all-reduce + Adam (eps 1e-15, reference lrs x 1e-3 so the scene stays stationary) on all %d M parameters" % (args.views, 66 * args.P // 1000000)}
        leaves = t_leaves

    # ---------------- per-kernel profile (ours) -> roofline ----------------
    roofline, roofline_hbm, kernels, counters = None, None, None, None
    if args.impl == "reference" and rank == 0:
        counters = counters_reference(leaves, cams[-1], bg, args)
    if args.impl == "ours" and rank == 0:
        rt.profile_enable(True)
        zero_grads(leaves)
        n_streams, args.streams = args.streams, 1       # kernels timed one at a time, not overlapping another view's
        ov, args.overlap_allreduce = getattr(args, "overlap_allreduce", 0), 0
        step_fn(leaves, cams, bg, grad, args)          # rank-local: NO collective here (other ranks are done)
        args.streams, args.overlap_allreduce = n_streams, ov
        torch.cuda.synchronize()
        prof = rt.profile_dump()
        rt.profile_enable(False)
        counters = counters_ours(leaves, cams[-1], bg, args)
        R = counters["R"]
        tiles = counters["tiles"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        twist_bytes = 28.0 if args.body_id is None else 4.0
        alg = {  # algorithmic bytes per launch (SURVEY.md 8d per-unit figures x units per launch; R = last view's)
            "preprocess_fwd": (288.0 + twist_bytes) * args.P,
            "preprocess_bwd": (516.0 + 2 * twist_bytes) * args.P,
            # view-batched forward (GaussianForwardBatch): parameters read once per step; per view the 20 bytes every
            # Gaussian gets (radius, tile count, rect, depth key) and the 53 bytes of a visible one (record, depth, clamp bits)
            "preprocess_fwd_batched": (236.0 + twist_bytes + 12.0) * args.P + len(cams) * 73.0 * args.P,
            # view-batched form (GaussianBackwardBatch): parameters read and gradients read-modify-written ONCE per step,
            # per view only the radius, the 48-byte moment record, conic + opacity, clamp bits and the view-space gradient
            "preprocess_bwd_batched": (252.0 + twist_bytes + 2 * (236.0 + twist_bytes)) * args.P + len(cams) * (4.0 + 48.0 + 24.0 + 1.0 + 12.0) * args.P,
            # radix fallback path (GSR_BINNING_RADIX=1)
            "duplicate_with_keys": 8.0 * R + 28.0 * args.P,
            "tile_sort_onesweep_pass": 16.0 * R,
            "tile_ranges": 4.0 * R + 8.0 * tiles,
        }
        # ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch at the default workload
        # (profiles/): only valid for P = 1 M @ 1920x1080, SH 3, per-Gaussian twists
        ncu_traffic = NCU_TRAFFIC
        default_workload = (args.P, args.W, args.H, args.bodies) == (1000000, 1920, 1080, False)
        kernels = {}
        for name, (cnt, tot) in prof.items():
            avg = tot / max(cnt, 1)
            k = {"launches": cnt, "avg_ms": round(avg, 5), "total_ms": round(tot, 4)}
            if name in alg and avg > 0:
                k["alg_GBps"] = round(alg[name] / avg / 1e6, 1)
                k["frac_of_hbm_peak"] = round(alg[name] / avg / 1e6 / peak, 4)
            kernels[name] = k
        # the whole binning stage against the reference scheme's traffic model (12 B/dup
        # duplicate + 152 B/dup 6-pass pair sort + 8 B/dup ranges); an implementation that moves fewer bytes
        # reads above what HBM could deliver to that scheme
        bin_names = ("scan_block_sums", "depth_sort_hist", "depth_sort_scan", "depth_sort_scatter", "depth_sort_local", "tile_count",
                     "tile_sweep_count", "tile_column_scan", "tile_base_scan", "tile_scatter", "tile_sweep_scatter",
                     "sorted_block_sums", "duplicate_with_keys", "tile_sort_onesweep_pass", "tile_ranges")
        bin_ms = sum(kernels[n]["total_ms"] for n in bin_names if n in kernels) / max(len(cams), 1)
        if bin_ms > 0 and R > 0:
            kernels["binning_stage"] = {"launches": 1, "avg_ms": round(bin_ms, 5), "total_ms": round(bin_ms * len(cams), 4),
                                        "alg_GBps": round(172.0 * R / bin_ms / 1e6, 1),
                                        "frac_of_hbm_peak": round(172.0 * R / bin_ms / 1e6 / peak, 4),
                                        "note": "sum of the binning kernels of one view vs the reference scheme's 172 B/duplicate"}
        hbm_kernels = [n for n in kernels if n in alg]
        dom = max(hbm_kernels, key=lambda n: kernels[n]["total_ms"]) if hbm_kernels else None
        if dom:
            a = alg[dom] / kernels[dom]["avg_ms"] / 1e6
            roofline_hbm = {"kernel": dom, "bound": "hbm", "achieved": round(a, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(a / peak, 4),
                            "traffic": ncu_traffic.get(dom) if default_workload else None, "peak_source": peak_src,
                            "note": "largest HBM-bound kernel; achieved = SURVEY 8d bytes x P / CUDA-event time"}
        # The time-dominant kernel is the blend backward: FP32-issue bound, not HBM or tensor bound (north_star: "FP32-pipe
        # utilisation for blend").  Roofline per SURVEY 8d: the reference's SASS spends 72 FP32-pipe instructions per BLENDED
        # (pixel, Gaussian) pair in the backward (28 in the forward); peak = 148 SMs x 128 FP32 lanes x SM clock thread-
        # instructions/s at the clock sampled during the timed region.
        sm_mhz = (clocks or {}).get("sm_mhz") or float(peaks.get("sm_max_mhz", 1965.0))
        fp32_peak = 148 * 128 * sm_mhz * 1e6
        for kname, per_pair in (("blend_bwd", 72.0), ("blend_fwd", 28.0)):
            if kname in kernels and kernels[kname]["avg_ms"] > 0:
                ach = per_pair * counters["pairs_blended"] / (kernels[kname]["avg_ms"] * 1e-3)
                kernels[kname]["alg_thread_instr_per_s"] = ach
                kernels[kname]["frac_of_fp32_issue_peak"] = round(ach / fp32_peak, 4)
        if "blend_bwd" in kernels and "frac_of_fp32_issue_peak" in kernels["blend_bwd"]:
            ach = kernels["blend_bwd"]["alg_thread_instr_per_s"]
            roofline = {"kernel": "blend_bwd", "bound": "fp32-issue", "achieved": round(ach / 1e12, 3), "peak": round(fp32_peak / 1e12, 3),
                        "unit": "T thread-instr/s", "frac": round(ach / fp32_peak, 4),
                        "traffic": ncu_traffic.get("blend_bwd") if default_workload else None,
                        "peak_source": "148 SM x 128 lanes x %.0f MHz (clock sampled during the timed region)" % sm_mhz,
                        "note": "time-dominant kernel (share = %.0f %% of the kernel sum); algorithmic work = 72 FP32 instr (SURVEY 8d, the "
                                "reference's own SASS count) x %d blended pairs of this view; DRAM traffic of the kernel is < 2 %% of HBM peak. "
                                "`roofline_hbm` is the largest HBM-bound kernel."
                                % (100.0 * kernels["blend_bwd"]["total_ms"] / max(1e-9, sum(v["total_ms"] for n, v in kernels.items() if n != "binning_stage")),
                                   counters["pairs_blended"])}
        else:
            roofline = roofline_hbm

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    out = {
        "metric": "gaussian_views_per_s_fwd_bwd",
        "value": value,
        "unit": "Gaussian-views/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "ms_per_view": ms_per_step / args.views,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "impl": args.impl,
        "config": {"workload": "%s; %d views/GPU/step on a camera circle (C5's 64 cameras at 8 GPUs)" % (cfg["what"], args.views),
                   "name": args.config,
                   "P": args.P, "width": args.W, "height": args.H, "views_per_gpu_per_step": args.views, "streams": args.streams,
                   "sync_free_forward": bool(args.sync_free) if args.impl == "ours" else False,
                   # the views of a step share the parameters: their per-Gaussian work (preprocess; the chain rule behind the
                   # blend backward) runs once per step for all views.  --batched-forward 0 --batched-backward 0: once per view
                   "view_batched": ({"forward_preprocess": bool(args.batched_forward) and args.views > 1,
                                     "gaussian_backward": bool(args.batched_backward) and args.views > 1}
                                    if args.impl == "ours" else None),
                   "parallelism": "view-parallel x%d, per-Gaussian grad all-reduce (NCCL)" % world if world > 1 else "single GPU",
                   "l2": "inputs (264 MB params+twists, 85 MB geometry state, 190 MB keys) exceed the 126 MB L2 every view"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "Gaussian-views/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "staging": ("each rank copies 1/%d of the parameters from pinned host memory and the slices are all-gathered over "
                            "NVLink (every byte crosses PCIe once per node)" % world) if sharded else "every rank copies all parameters from its pinned host buffer"},
        "gpu_launches": launches if args.impl == "ours" else 0,
    }
    if collective is not None:
        out["collective"] = collective
    if train is not None:
        out["train_step"] = train
    out["breakdown"] = breakdown
    out["counters"] = counters
    if args.impl == "ours":
        out["roofline"] = roofline
        out["roofline_hbm"] = roofline_hbm
        out["kernels"] = kernels
    if not args.no_mlp and world == 1:
        try:
            del leaves
            torch.cuda.empty_cache()
            out["deform_network"] = deform_network_timing(args.impl, min(args.P, 1000000), dev)
        except Exception as ex:  # pragma: no cover
            out["deform_network"] = {"error": repr(ex)}
    if not args.no_cpu_baseline and world == 1:
        try:
            out["cpu_baseline"] = cpu_baseline()
        except Exception as ex:  # pragma: no cover
            out["cpu_baseline"] = {"error": repr(ex)}
    if args.impl == "reference":
        out["reference_breakdown"] = breakdown
        out["reference_arm"] = ("reference CUDA rasterizer (oracle/_ref, sm_100, unmodified sources) + torch-CUDA "
                                "rigid_body op graph; the reference's torch-CPU deformation is `cpu_baseline`")
    print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
