"""On-disk formats of the reference next to the point cloud.

* `save_point_cloud` / `load_point_cloud` - what `GaussianModel.save_ply` / `load_ply` write and read
  (scene/gaussian_model.py:905-929, 965-1025): `point_cloud.ply` (ply_io) plus the five network state-dicts
  `offset_model.pth`, `offset_model_rot.pth`, `offset_model_scaling.pth`, `opacity_mask.pth`, `shs_model.pth` beside it
  (the reference's `load_ply` torch.load()s all five unconditionally).
* `capture` / `restore` - the 12-tuple of `GaussianModel.capture()` / `.restore()` (:686-728), whose optimizer entry is
  `torch.optim.Adam.state_dict()` (FusedAdam emits and accepts that layout).
* `save_checkpoint` / `load_checkpoint` - train.py:685-697 / :99-113: `ckpt_save/chkpnt_<it>.pth` = `(capture(), iteration)`
  and the five networks as `<name>_<it>.pth`.

Only `offset_model` (DirectTemporalNeRF) is evaluated by the reference's render(); the other four networks are
constructed, saved and loaded but never called (their call sites are commented out, gaussian_renderer/__init__.py:97-140),
so here they are plain containers with the reference's parameter names and shapes.
"""
import os

import torch
import torch.nn as nn

import ply_io

NETWORKS = ("offset_model", "offset_model_rot", "offset_model_scaling", "opacity_mask", "shs_model")


class _Container(nn.Module):
    """`_time` (ModuleList of Linear) + `_time_out` (Linear): parameter names / shapes of the reference's dormant
    DirectTemporalNeRF_{rot,scaling,opacitymask,shs} (scene/gaussian_model.py:386-630)."""

    def __init__(self, D, W, in_ch, in_time, skips, out_ch):
        super().__init__()
        layers = [nn.Linear(in_ch + in_time, W)]
        for i in range(D - 1):
            layers.append(nn.Linear(W + (in_ch if i in skips else 0), W))
        self._time = nn.ModuleList(layers)
        self._time_out = nn.Linear(W, out_ch)


def make_networks(device="cpu"):
    """The five networks GaussianModel.__init__ builds (scene/gaussian_model.py:680-684), in its construction order."""
    import deform_mlp
    nets = {
        "offset_model": deform_mlp.DirectTemporalNeRF(),
        "offset_model_rot": _Container(3, 256, 7, 21, (4,), 4),        # :441-474 (time embedded: 21 channels; D = 3: no skip reached)
        "offset_model_scaling": _Container(8, 256, 6, 1, (4,), 3),     # :386-415
        "opacity_mask": _Container(8, 256, 3, 1, (4,), 1),             # :505-534
        "shs_model": _Container(8, 256, 3, 1, (4,), 48),               # :561-594
    }
    return {k: v.to(device) for k, v in nets.items()}


def save_point_cloud(path, xyz, features_dc, features_rest, opacity, scaling, rotation, networks):
    """GaussianModel.save_ply: `path` = .../point_cloud.ply; the networks go beside it."""
    ply_io.save_ply(path, xyz, features_dc, features_rest, opacity, scaling, rotation)
    d = os.path.dirname(path)
    for name in NETWORKS:
        torch.save(networks[name].state_dict(), os.path.join(d, name + ".pth"))


def load_point_cloud(path, networks=None, device="cuda", max_sh_degree=3):
    """GaussianModel.load_ply: returns (tensors dict, networks) with the networks' weights loaded from beside the PLY."""
    t = ply_io.load_ply(path, max_sh_degree=max_sh_degree, device=device)
    networks = networks if networks is not None else make_networks(device)
    d = os.path.dirname(path)
    for name in NETWORKS:
        networks[name].load_state_dict(torch.load(os.path.join(d, name + ".pth"), map_location=device))
    return t, networks


def capture(active_sh_degree, params, stats, optimizer, spatial_lr_scale):
    """GaussianModel.capture() (:686-700).  `params`: {"xyz","f_dc","f_rest","scaling","rotation","opacity"}; `stats`: densify.DensificationStats."""
    return (active_sh_degree, params["xyz"], params["f_dc"], params["f_rest"], params["scaling"], params["rotation"], params["opacity"],
            stats.max_radii2D, stats.xyz_gradient_accum, stats.denom, optimizer.state_dict(), spatial_lr_scale)


def restore(model_args, make_optimizer, device="cuda"):
    """GaussianModel.restore() (:702-728): `make_optimizer(params_dict)` plays training_setup (builds the FusedAdam over the
    restored tensors, reference group order); the optimizer state is then loaded from the tuple.
    Returns (active_sh_degree, params, stats, optimizer, spatial_lr_scale)."""
    import densify
    (active_sh_degree, xyz, f_dc, f_rest, scaling, rotation, opacity, max_radii2D, accum, denom, opt_dict, spatial_lr_scale) = model_args
    leaf = lambda t: t.detach().to(device).float().clone().requires_grad_(True)
    params = {"xyz": leaf(xyz), "f_dc": leaf(f_dc), "f_rest": leaf(f_rest), "scaling": leaf(scaling), "rotation": leaf(rotation),
              "opacity": leaf(opacity)}
    opt = make_optimizer(params)
    stats = densify.DensificationStats(params["xyz"].shape[0], device)
    stats.max_radii2D = max_radii2D.detach().to(device).float().clone()
    stats.xyz_gradient_accum = accum.detach().to(device).float().clone()
    stats.denom = denom.detach().to(device).float().clone()
    opt.load_state_dict(opt_dict)
    return active_sh_degree, params, stats, opt, spatial_lr_scale


def save_checkpoint(model_path, iteration, captured, networks):
    """train.py:685-697."""
    d = os.path.join(model_path, "ckpt_save")
    os.makedirs(d, exist_ok=True)
    torch.save((captured, iteration), os.path.join(d, "chkpnt_%d.pth" % iteration))
    for name in NETWORKS:
        torch.save(networks[name].state_dict(), os.path.join(d, "%s_%d.pth" % (name, iteration)))
    return os.path.join(d, "chkpnt_%d.pth" % iteration)


def load_checkpoint(checkpoint, networks=None, device="cuda"):
    """train.py:99-113: returns (model_args, first_iter, networks)."""
    d = os.path.dirname(checkpoint)
    ckpt_id = os.path.basename(checkpoint).split(".")[0].split("_")[-1]
    networks = networks if networks is not None else make_networks(device)
    for name in NETWORKS:
        networks[name].load_state_dict(torch.load(os.path.join(d, "%s_%s.pth" % (name, ckpt_id)), map_location=device))
    model_args, first_iter = torch.load(checkpoint, map_location=device, weights_only=False)
    return model_args, first_iter, networks
