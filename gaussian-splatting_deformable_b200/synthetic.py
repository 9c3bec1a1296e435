"""Synthetic scenes, cameras, twists and upstream gradients.

Datasets are unavailable offline, so every test and benchmark uses these
generators.  Everything is drawn on a CPU torch.Generator (device independent)
and then moved to the requested device.

Shapes and conventions follow the reference:
  * cloud: uniform in the cube U(-1.3,1.3)^3 the reference initialises from
    (scene/dataset_readers.py:577-585);
  * parameters and activations as scene/gaussian_model.py:825-831 and
    gaussian_renderer/__init__.py:116,122,140 (exp / normalize / sigmoid);
  * cameras as scene/cameras.py:49-58 with utils/graphics_utils.py:38-77
    (world_view_transform and full_proj_transform are TRANSPOSED matrices).
"""
import math
from types import SimpleNamespace

import numpy as np
import torch

FOVX = 0.6911112   # Blender camera_angle_x of the D-NeRF scenes
ZNEAR, ZFAR = 0.01, 100.0
CAM_RADIUS = 4.0


def make_scene(P, seed=0, device="cpu", sh_degree=3, scale_mult=1.0):
    """Activated Gaussian parameters: dict of means3D[P,3], scales[P,3], rotations[P,4],
    opacities[P,1], shs[P,16,3]."""
    g = torch.Generator().manual_seed(seed)
    xyz = (torch.rand((P, 3), generator=g) * 2.0 - 1.0) * 1.3
    s0 = 0.25 * (2.6 ** 3 / max(P, 1)) ** (1.0 / 3.0) * scale_mult
    log_scale = math.log(s0) + 0.5 * torch.randn((P, 3), generator=g)
    rot = torch.randn((P, 4), generator=g)
    opacity_logit = 2.0 * torch.randn((P, 1), generator=g)
    f_dc = torch.randn((P, 1, 3), generator=g)
    f_rest = 0.2 * torch.randn((P, 15, 3), generator=g)
    out = dict(
        means3D=xyz,
        scales=torch.exp(log_scale),
        rotations=torch.nn.functional.normalize(rot),
        opacities=torch.sigmoid(opacity_logit),
        shs=torch.cat([f_dc, f_rest], dim=1).contiguous(),
    )
    return {k: v.to(device=device, dtype=torch.float32).contiguous() for k, v in out.items()}


def _projection(znear, zfar, fovx, fovy):
    # utils/graphics_utils.py:51-71
    t = math.tan(fovy / 2) * znear
    r = math.tan(fovx / 2) * znear
    Pm = torch.zeros(4, 4)
    Pm[0, 0] = 2.0 * znear / (2 * r)
    Pm[1, 1] = 2.0 * znear / (2 * t)
    Pm[0, 2] = 0.0
    Pm[1, 2] = 0.0
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


def make_camera(k=0, K=1, W=1920, H=1080, device="cpu", radius=CAM_RADIUS, fovx=FOVX):
    """Camera k of K on a circle in the xz-plane looking at the origin (COLMAP
    convention: z forward, y down).  K=1 -> position (0,0,-radius), R=I, T=(0,0,radius)."""
    phi = 2.0 * math.pi * k / K
    C = np.array([radius * math.sin(phi), 0.0, -radius * math.cos(phi)])
    right = np.array([math.cos(phi), 0.0, math.sin(phi)])
    down = np.array([0.0, 1.0, 0.0])
    fwd = np.array([-math.sin(phi), 0.0, math.cos(phi)])
    Rw2c = np.stack([right, down, fwd])          # rows
    t = -Rw2c @ C
    Rt = np.eye(4)
    Rt[:3, :3] = Rw2c
    Rt[:3, 3] = t
    focal = W / (2 * math.tan(fovx / 2))
    fovy = 2 * math.atan(H / (2 * focal))
    wvt = torch.tensor(np.float32(Rt)).transpose(0, 1).contiguous()
    proj = _projection(ZNEAR, ZFAR, fovx, fovy).transpose(0, 1)
    full = wvt.unsqueeze(0).bmm(proj.unsqueeze(0)).squeeze(0).contiguous()
    center = wvt.inverse()[3, :3].contiguous()
    return SimpleNamespace(
        image_width=W, image_height=H, FoVx=fovx, FoVy=fovy, znear=ZNEAR, zfar=ZFAR,
        world_view_transform=wvt.to(device), full_proj_transform=full.to(device),
        camera_center=center.to(device), time=(k / (K - 1) if K > 1 else 0.0))


def make_twists(N, seed=2, device="cpu"):
    """Per-Gaussian screw axes S[N,6]=(w,v) and magnitudes theta[N]
    (mirrors scene/gaussian_model.py:161-164: theta=|w_raw|, S=(w_raw,v_raw)/theta)."""
    g = torch.Generator().manual_seed(seed)
    w_raw = 0.2 * torch.randn((N, 3), generator=g)
    v_raw = 0.05 * torch.randn((N, 3), generator=g)
    theta = w_raw.norm(dim=1)
    bad = theta < 1e-3
    while bad.any():
        w_raw[bad] = 0.2 * torch.randn((int(bad.sum()), 3), generator=g)
        theta = w_raw.norm(dim=1)
        bad = theta < 1e-3
    S = torch.cat([w_raw / theta[:, None], v_raw / theta[:, None]], dim=1)
    return S.to(device).contiguous(), theta.to(device).contiguous()


def make_bodies(means, frame=0, frames=300, seed=2, device="cpu"):
    """64 rigid bodies on a 4x4x4 grid over the cube; body twist xi[b] scaled by frame."""
    q = ((means.detach().cpu() + 1.3) / 2.6 * 4).floor().clamp(0, 3).long()
    body_id = (16 * q[:, 0] + 4 * q[:, 1] + q[:, 2]).to(torch.int32)
    g = torch.Generator().manual_seed(seed)
    xi_w = 0.5 * torch.randn((64, 3), generator=g)
    xi_v = 0.2 * torch.randn((64, 3), generator=g)
    s = max(frame, 1) / max(frames - 1, 1)
    w_raw, v_raw = s * xi_w, s * xi_v
    theta = w_raw.norm(dim=1).clamp_min(1e-6)
    S = torch.cat([w_raw / theta[:, None], v_raw / theta[:, None]], dim=1)
    return body_id.to(device), S.to(device).contiguous(), theta.to(device).contiguous()


def make_image_grad(W, H, seed=1, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    return torch.randn((3, H, W), generator=g).to(device)


def raster_settings(cam, bg, sh_degree=3, scale_modifier=1.0, debug=False):
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    return GaussianRasterizationSettings(
        image_height=int(cam.image_height), image_width=int(cam.image_width),
        tanfovx=math.tan(cam.FoVx * 0.5), tanfovy=math.tan(cam.FoVy * 0.5), bg=bg,
        scale_modifier=scale_modifier, viewmatrix=cam.world_view_transform,
        projmatrix=cam.full_proj_transform, sh_degree=sh_degree, campos=cam.camera_center,
        prefiltered=False, debug=debug)
