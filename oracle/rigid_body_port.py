"""CPU/torch restatement of the reference's scene/rigid_body.py (TEST INFRASTRUCTURE).

Used as (a) the checker for the fused SE3 kernels and the rigid_body drop-in and
(b) the torch-CPU / torch-CUDA deformation baseline in bench.py (the reference file
itself cannot travel to the GPU box).  It keeps the reference's OP GRAPH - a skew
matrix built with stack, an identity built with eye().repeat, three bmm's and two
cat's per exp_se3 (scene/rigid_body.py:16-24,41-45,61-65,86-93) - because the
baseline is meant to cost what the reference costs; the text is a restatement, not
a copy.  Pinned against the real file by tests/golden/make_se3_golden.py
(bit-exact on CPU) -> tests/golden/se3_golden.pt.
"""
import torch


def _b(x):
    """[N] -> [N,1,1] for broadcasting against [N,3,3]."""
    return x.unsqueeze(-1).unsqueeze(-1)


def skew(w):
    # rigid_body.py:16-24
    w = w.reshape(w.shape[0], 3)
    z = torch.zeros(w.shape[0], device=w.device)
    rows = [torch.stack([z, -w[:, 2], w[:, 1]], dim=1),
            torch.stack([w[:, 2], z, -w[:, 0]], dim=1),
            torch.stack([-w[:, 1], w[:, 0], z], dim=1)]
    return torch.stack(rows, dim=1)


def _eye(n, device):
    return torch.eye(3, device=device).unsqueeze(0).repeat(n, 1, 1)


def exp_so3(w, theta):
    # rigid_body.py:61-65
    W = skew(w)
    return _eye(w.shape[0], w.device) + _b(torch.sin(theta)) * W + (1.0 - _b(torch.cos(theta))) * torch.bmm(W, W)


def rp_to_se3(R, p):
    # rigid_body.py:41-45
    p = p.view(p.shape[0], 3, 1)
    bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]], device=p.device).repeat(p.shape[0], 1).unsqueeze(1)
    return torch.cat([torch.cat([R, p], dim=2), bottom], dim=1)


def exp_se3(S, theta):
    # rigid_body.py:86-93
    w, v = torch.split(S, 3, dim=1)
    W = skew(w)
    R = exp_so3(w, theta)
    A = _b(theta) * _eye(S.shape[0], S.device) + (1.0 - _b(torch.cos(theta))) * W + \
        (_b(theta) - _b(torch.sin(theta))) * torch.bmm(W, W)
    p = torch.bmm(A, v.unsqueeze(-1)).squeeze(-1)
    return rp_to_se3(R, p)


def to_homogenous(v):
    return torch.cat([v, torch.ones_like(v[..., :1])], dim=-1)


def from_homogenous(v):
    return v[..., :3] / v[..., -1:]


def deform_points(x, S, theta):
    """The apply recipe commented at gaussian_renderer/__init__.py:92-95:
    y = from_homogenous(T @ to_homogenous(x))."""
    T = exp_se3(S, theta)
    return from_homogenous(torch.bmm(T, to_homogenous(x).unsqueeze(-1)).squeeze(-1))
