"""Drop-in for the reference's `utils/loss_utils.py` (same names and signatures), backed by the fused
sm_100a loss kernels of libgsr_b200 (csrc/loss.cu) - plus `l1_ssim_loss`, the whole training loss of
train.py:323,529 as ONE forward and ONE backward kernel.

    loss = (1.0 - opt.lambda_dssim) * l1_loss(image, gt) + opt.lambda_dssim * (1.0 - ssim(image, gt))   # reference
    loss = l1_ssim_loss(image, gt, opt.lambda_dssim)                                                     # fused

There is no CPU path: `ssim` / `l1_ssim_loss` on CPU tensors raise (`l1_loss` / `l2_loss` are the reference's
own one-line torch expressions and work anywhere).
"""
from math import exp

import torch

import gsr_runtime as _rt


def l1_loss(network_output, gt):
    """Mean absolute error (utils/loss_utils.py:17-18)."""
    return (network_output - gt).abs().mean()


def l2_loss(network_output, gt):
    """Mean squared error (utils/loss_utils.py:20-21)."""
    diff = network_output - gt
    return (diff * diff).mean()


def gaussian(window_size, sigma):
    """Normalised 1-D Gaussian taps as a float32 tensor - the values utils/loss_utils.py:23-25 produces
    (double-precision exp per tap, rounded to float32, then normalised in float32)."""
    centre = window_size // 2
    denom = float(2 * sigma ** 2)
    taps = torch.tensor([exp(-((x - centre) ** 2) / denom) for x in range(window_size)], dtype=torch.float32)
    return taps / taps.sum()


def create_window(window_size, channel):
    """[channel, 1, k, k] window = outer product of the 1-D taps (sigma 1.5), utils/loss_utils.py:27-31."""
    w1 = gaussian(window_size, 1.5)
    w2 = torch.outer(w1, w1).float()
    return w2.expand(channel, 1, window_size, window_size).contiguous()


_WINDOW11 = None


def _window11():
    global _WINDOW11
    if _WINDOW11 is None:
        _WINDOW11 = gaussian(11, 1.5).contiguous()          # host tensor, passed by pointer
    return _WINDOW11


class _SsimL1(torch.autograd.Function):
    """out = (1 - lam) * mean|x - gt| + lam * (1 - ssim(x, gt)); also leaves the L1 and SSIM means in ctx."""

    @staticmethod
    def forward(ctx, image, gt, lam):
        if not image.is_cuda or not gt.is_cuda:
            raise _rt.GsrError("libgsr_b200 loss kernels run on CUDA tensors only (no CPU fallback)")
        if image.dim() != 3 or image.shape[0] != 3 or image.shape != gt.shape:
            raise RuntimeError("ssim/l1_ssim_loss expect image and gt of the same shape [3,H,W]")
        lib = _rt.load()
        x = image.detach().float().contiguous()
        y = gt.detach().float().contiguous()
        C, H, W = (int(s) for s in x.shape)
        dev = x.device
        with torch.cuda.device(dev):
            dmaps = torch.empty((3, C, H, W), dtype=torch.float32, device=dev)
            out = torch.empty(8, dtype=torch.float32, device=dev)
            win = _window11()
            _rt.check(lib.gsr_ssim_l1_loss_forward(_rt.ptr(x), _rt.ptr(y), C, H, W, win.data_ptr(), float(lam),
                                                   _rt.ptr(dmaps), _rt.ptr(out), _rt.stream_ptr(dev)))
        if gt.requires_grad:
            raise _rt.GsrError("l1_ssim_loss / ssim: the target image must not require grad (only `image` is differentiated)")
        ctx.save_for_backward(x, y, dmaps)
        ctx.lam = float(lam)
        ctx.in_dtype = image.dtype
        loss, l1, ss = out[3], out[4], out[5]
        ctx.mark_non_differentiable(l1, ss)          # logging values: differentiating them raises instead of returning zeros
        return loss, l1, ss

    @staticmethod
    def backward(ctx, g_loss, g_l1, g_ssim):
        x, y, dmaps = ctx.saved_tensors
        lib = _rt.load()
        C, H, W = (int(s) for s in x.shape)
        dev = x.device
        if g_loss is None:
            return torch.zeros_like(x).to(ctx.in_dtype), None, None
        up = g_loss.detach().float().contiguous().reshape(1)
        with torch.cuda.device(dev):
            grad = torch.empty_like(x)
            _rt.check(lib.gsr_ssim_l1_loss_backward(_rt.ptr(x), _rt.ptr(y), C, H, W, _window11().data_ptr(), ctx.lam,
                                                    _rt.ptr(dmaps), _rt.ptr(up), _rt.ptr(grad), _rt.stream_ptr(dev)))
        return grad.to(ctx.in_dtype), None, None


def l1_ssim_loss(image, gt, lambda_dssim=0.2, return_parts=False):
    """(1 - lambda) * L1 + lambda * (1 - SSIM) of train.py:529 as one fused op.  Differentiable w.r.t. `image`.
    With return_parts=True also returns the (detached) L1 and SSIM means for logging (train.py:548)."""
    loss, l1, ss = _SsimL1.apply(image, gt, lambda_dssim)
    if return_parts:
        return loss, l1.detach(), ss.detach()
    return loss


def ssim(img1, img2, window_size=11, size_average=True):
    """utils/loss_utils.py:33-42 (mean SSIM of two [3,H,W] images).  Only the reference's own call shape
    (window 11, size_average=True) is implemented natively."""
    if window_size != 11 or not size_average:
        raise NotImplementedError("libgsr_b200 ssim: window_size=11, size_average=True (the reference's only use, train.py:529)")
    loss, _, _ = _SsimL1.apply(img1, img2, 1.0)        # lam = 1: loss = 1 - ssim
    return 1.0 - loss
