"""Densification of the Gaussian set on the flat optimizer buffers.

The reference grows and prunes its point set every `densification_interval` iterations (train.py:610-648):
`add_densification_stats` accumulates the view-space gradient norms, `densify_and_prune` clones small Gaussians with
large gradients, splits large ones into N samples, prunes transparent / oversized ones, and `reset_opacity` clamps the
opacities - each followed by surgery on torch.optim.Adam's state (scene/gaussian_model.py:960-963, 1027-1257).

Here the statistics and the decisions are single kernels over the per-point arrays (csrc/densify.cu) and the surgery is
`fused_adam.FusedAdam.prune / append / replace` on the flat parameter, gradient and moment buffers.  Same order of
operations, same row order of the resulting tensors, same random numbers (the split draws its samples with the caller's
torch generator exactly as `torch.normal` does), so a replica that densifies here and one that runs the reference's
methods hold the same state - and, with `view_parallel.reduce_densification_stats` and a common seed, so do all ranks.

The optimizer must have the reference's group names: "xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"
(scene/gaussian_model.py:840-860); other groups (e.g. "offset_model") are left alone.
"""
import torch

import gsr_runtime as _rt

POINT_GROUPS = ("xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation")


class DensificationStats:
    """xyz_gradient_accum [P,1], xyz_gradient_accum_3vec [P,3], denom [P,1], max_radii2D [P]
    (scene/gaussian_model.py:836-838, 831)."""

    def __init__(self, num_points, device):
        self.device = torch.device(device)
        self.reset(num_points)

    def reset(self, num_points):
        # densification_postfix (gaussian_model.py:1123-1127) zeroes all four for the enlarged set
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=self.device)
        self.xyz_gradient_accum, self.xyz_gradient_accum_3vec = z(num_points, 1), z(num_points, 3)
        self.denom, self.max_radii2D = z(num_points, 1), z(num_points)

    def keep(self, valid_mask):
        # prune_points (gaussian_model.py:1076-1080)
        self.xyz_gradient_accum = self.xyz_gradient_accum[valid_mask]
        self.xyz_gradient_accum_3vec = self.xyz_gradient_accum_3vec[valid_mask]
        self.denom = self.denom[valid_mask]
        self.max_radii2D = self.max_radii2D[valid_mask]

    def add(self, viewspace_point_grad, radii):
        """train.py:613,618: `max_radii2D[vis] = max(max_radii2D[vis], radii[vis])` and `add_densification_stats(viewspace, vis)`
        with vis = radii > 0, in one pass.  `viewspace_point_grad` is `render()["viewspace_points"].grad` [P,3]."""
        lib = _rt.load()
        g = viewspace_point_grad.detach().float().contiguous()
        r = radii.to(torch.int32).contiguous()
        P = int(g.shape[0])
        if not g.is_cuda:
            raise _rt.GsrError("densify: CUDA tensors only (no CPU fallback)")
        with torch.cuda.device(g.device):
            _rt.check(lib.gsr_densify_stats(P, g.data_ptr(), r.data_ptr(), self.xyz_gradient_accum.data_ptr(),
                                            self.xyz_gradient_accum_3vec.data_ptr(), self.denom.data_ptr(),
                                            self.max_radii2D.data_ptr(), _rt.stream_ptr(g.device)))

    def all_reduce(self, group=None):
        import view_parallel
        view_parallel.reduce_densification_stats(self.xyz_gradient_accum, self.denom, self.max_radii2D, group=group)


def _named(opt):
    out = {}
    for g in opt.param_groups:
        if g.get("name") in POINT_GROUPS and len(g["params"]) == 1:
            out[g["name"]] = g["params"][0]
    missing = [n for n in POINT_GROUPS if n not in out]
    if missing:
        raise ValueError("densify: optimizer has no parameter group named %s" % missing)
    return out


def decide(stats, scaling_raw, grad_threshold, percent_dense, extent):
    """(clone mask, split mask) over the current points (gaussian_model.py:1186-1190, 1129-1137)."""
    lib = _rt.load()
    P = int(scaling_raw.shape[0])
    dev = scaling_raw.device
    flags = torch.empty(P, dtype=torch.uint8, device=dev)
    sc = scaling_raw.detach().contiguous()
    with torch.cuda.device(dev):
        _rt.check(lib.gsr_densify_decide(P, stats.xyz_gradient_accum.data_ptr(), stats.denom.data_ptr(), sc.data_ptr(),
                                         float(grad_threshold), float(percent_dense) * float(extent), flags.data_ptr(),
                                         _rt.stream_ptr(dev)))
    return (flags & 1).bool(), (flags & 2).bool()


def densify_and_prune(opt, stats, max_grad, min_opacity, extent, max_screen_size, percent_dense=0.01, N=2, generator=None):
    """GaussianModel.densify_and_prune (gaussian_model.py:1219-1233) on a FusedAdam + DensificationStats pair.
    Returns {group name: parameter} (the same Python objects, re-homed with their new row counts)."""
    lib = _rt.load()
    p = _named(opt)
    dev = p["xyz"].device
    P = int(p["xyz"].shape[0])
    clone, split = decide(stats, p["scaling"], max_grad, percent_dense, extent)
    # ---- densify_and_clone (:1186-1200): copies of the selected rows, appended ----
    idx_c = clone.nonzero(as_tuple=True)[0]
    opt.append({k: p[k].detach()[idx_c] for k in POINT_GROUPS})
    n_c = int(idx_c.numel())
    stats.reset(P + n_c)
    # ---- densify_and_split (:1129-1152): N samples of each selected ORIGINAL row (the padded gradient of a clone is 0) ----
    p = _named(opt)
    idx_s = split.nonzero(as_tuple=True)[0]
    n_s = int(idx_s.numel())
    # torch.normal(mean=0, std=stds) is randn * std: drawing randn with the same generator consumes the same random numbers
    z = torch.randn((N * n_s, 3), generator=generator, device=dev, dtype=torch.float32)
    xyz_s, sc_s, rot_s = p["xyz"].detach()[idx_s].contiguous(), p["scaling"].detach()[idx_s].contiguous(), p["rotation"].detach()[idx_s].contiguous()
    new_xyz = torch.empty((N * n_s, 3), dtype=torch.float32, device=dev)
    new_scaling = torch.empty((N * n_s, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _rt.check(lib.gsr_densify_split(n_s, N, xyz_s.data_ptr(), sc_s.data_ptr(), rot_s.data_ptr(), z.data_ptr(),
                                        new_xyz.data_ptr(), new_scaling.data_ptr(), _rt.stream_ptr(dev)))
    d = {"xyz": new_xyz, "scaling": new_scaling, "rotation": rot_s.repeat(N, 1),
         "f_dc": p["f_dc"].detach()[idx_s].repeat(N, 1, 1), "f_rest": p["f_rest"].detach()[idx_s].repeat(N, 1, 1),
         "opacity": p["opacity"].detach()[idx_s].repeat(N, 1)}
    opt.append(d)
    stats.reset(P + n_c + N * n_s)
    prune_filter = torch.cat((split, torch.zeros(n_c + N * n_s, dtype=torch.bool, device=dev)))
    opt.prune(~prune_filter)
    stats.keep(~prune_filter)
    # ---- prune (:1226-1231) on the resulting set ----
    p = _named(opt)
    P2 = int(p["xyz"].shape[0])
    mask = torch.empty(P2, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _rt.check(lib.gsr_densify_prune(P2, p["opacity"].detach().contiguous().data_ptr(), p["scaling"].detach().contiguous().data_ptr(),
                                        stats.max_radii2D.data_ptr(), float(min_opacity),
                                        float(max_screen_size) if max_screen_size else 0.0, 0.1 * float(extent),
                                        1 if max_screen_size else 0, mask.data_ptr(), _rt.stream_ptr(dev)))
    keep = ~mask.bool()
    opt.prune(keep)
    stats.keep(keep)
    return _named(opt)


def reset_opacity(opt):
    """GaussianModel.reset_opacity (gaussian_model.py:960-963): opacity <- inverse_sigmoid(min(sigmoid(opacity), 0.01)), moments reset."""
    o = _named(opt)["opacity"].detach()
    new = torch.min(torch.sigmoid(o), torch.ones_like(o) * 0.01)
    new = torch.log(new / (1 - new))
    opt.replace("opacity", new)
    return _named(opt)
