"""Drop-in for the reference's scene/rigid_body.py (same six functions, same
signatures and return shapes: skew :16, rp_to_se3 :41, exp_so3 :61, exp_se3 :86,
to_homogenous :96, from_homogenous :99).

`exp_se3` runs one hand-written sm_100a kernel forward and one backward
(gsr_exp_se3 / gsr_exp_se3_backward in include/gsr_b200.h) instead of the
reference's ~25 small torch kernels and three bmm's.  For the render hot path,
prefer the fully fused form: `GaussianRasterizer(...)(..., se3_S=S, se3_theta=theta)`
never materialises the [N,4,4] transforms at all.

The helpers that are pure data movement (skew, rp_to_se3, homogeneous
conversions) and exp_so3 stay thin torch code - they are not on the hot path.
"""
import torch

import gsr_runtime as _rt


class _ExpSE3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S, theta):
        if not S.is_cuda:
            raise _rt.GsrError("exp_se3 runs on CUDA tensors only (no CPU fallback)")
        lib = _rt.load()
        N = int(S.shape[0])
        Sc = S.detach().float().contiguous()
        th = theta.detach().float().reshape(-1).contiguous()
        if th.numel() != N or Sc.shape[1] != 6:
            raise RuntimeError("exp_se3 expects S[N,6] and theta[N]")
        T = torch.empty((N, 4, 4), dtype=torch.float32, device=S.device)
        with torch.cuda.device(S.device):
            _rt.check(lib.gsr_exp_se3(N, _rt.ptr(Sc), _rt.ptr(th), _rt.ptr(T), _rt.stream_ptr(S.device)))
        ctx.save_for_backward(Sc, th)
        ctx.theta_shape = theta.shape
        return T

    @staticmethod
    def backward(ctx, dT):
        lib = _rt.load()
        Sc, th = ctx.saved_tensors
        N = int(Sc.shape[0])
        dS = torch.empty_like(Sc)
        dth = torch.empty_like(th)
        g = dT.float().contiguous()
        with torch.cuda.device(Sc.device):
            _rt.check(lib.gsr_exp_se3_backward(N, _rt.ptr(Sc), _rt.ptr(th), _rt.ptr(g), _rt.ptr(dS), _rt.ptr(dth),
                                               _rt.stream_ptr(Sc.device)))
        return dS, dth.reshape(ctx.theta_shape)


def skew(w):
    """Batch of cross-product matrices: skew(w) @ u == w x u.  [N,3] -> [N,3,3]."""
    w = w.reshape(w.shape[0], 3)
    W = w.new_zeros((w.shape[0], 3, 3))
    W[:, 0, 1], W[:, 0, 2] = -w[:, 2], w[:, 1]
    W[:, 1, 0], W[:, 1, 2] = w[:, 2], -w[:, 0]
    W[:, 2, 0], W[:, 2, 1] = -w[:, 1], w[:, 0]
    return W


def rp_to_se3(R, p):
    """Batch of rotations [N,3,3] and translations [N,3] -> homogeneous [N,4,4]."""
    N = p.shape[0]
    X = R.new_zeros((N, 4, 4))
    X[:, :3, :3] = R
    X[:, :3, 3] = p.reshape(N, 3)
    X[:, 3, 3] = 1.0
    return X


def exp_so3(w, theta):
    """Rodrigues: I + sin(theta) W + (1 - cos(theta)) W^2.  [N,3], [N] -> [N,3,3]."""
    W = skew(w)
    s = torch.sin(theta)[:, None, None]
    c = torch.cos(theta)[:, None, None]
    eye = torch.eye(3, device=w.device, dtype=W.dtype)[None]
    return eye + s * W + (1.0 - c) * torch.bmm(W, W)


def exp_se3(S, theta):
    """Exponential map from a batch of screw axes S[N,6] = (w, v) and magnitudes
    theta[N] to homogeneous transforms [N,4,4]."""
    return _ExpSE3.apply(S, theta)


def to_homogenous(v):
    return torch.cat([v, torch.ones_like(v[..., :1])], dim=-1)


def from_homogenous(v):
    return v[..., :3] / v[..., -1:]
