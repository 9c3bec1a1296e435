"""TEST INFRASTRUCTURE ONLY (oracle): torch restatement of the reference's training loss and optimizer step.

* `l1_loss`, `gaussian`, `create_window`, `ssim`, `training_loss` follow utils/loss_utils.py:17-64 and
  train.py:323,529 op for op (grouped `F.conv2d` with the 11x11 window, zero padding 5).
* `adam_reference` drives `torch.optim.Adam(lr=0.0, eps=1e-15)` with per-tensor groups the way
  scene/gaussian_model.py:834-846 builds it.  torch.optim is a third-party dependency that is not vendored in
  /root/reference: torch 2.11.0 (this image), algorithm = Kingma & Ba with torch's `_single_tensor_adam` op order.
Pinned: tests/golden/loss_golden.pt holds outputs of the REAL utils/loss_utils.py imported from /root/reference
(tests/golden/make_loss_golden.py); tests/test_oracle_cpu.py holds this port to them.
Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
from math import exp

import torch
import torch.nn.functional as F


def l1_loss(network_output, gt):                      # loss_utils.py:17-18
    return torch.abs((network_output - gt)).mean()


def gaussian(window_size, sigma):                     # loss_utils.py:23-25
    gauss = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return gauss / gauss.sum()


def create_window(window_size, channel):              # loss_utils.py:27-31
    w1 = gaussian(window_size, 1.5).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1, img2, window_size=11):                 # loss_utils.py:33-64, size_average=True
    channel = img1.size(-3)
    window = create_window(window_size, channel).to(img1.device).type_as(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def training_loss(image, gt, lambda_dssim=0.2):       # train.py:323,529
    return (1.0 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1.0 - ssim(image, gt))


def adam_reference(params, grads_per_step, lrs, eps=1e-15):
    """Run torch.optim.Adam the way gaussian_model.training_setup builds it: one group per tensor, lr=0.0 default,
    eps=1e-15.  `grads_per_step` is a list (steps) of lists (per tensor) of gradients.  Returns the final params."""
    ps = [p.clone().requires_grad_(True) for p in params]
    groups = [{"params": [p], "lr": lr, "name": "g%d" % i} for i, (p, lr) in enumerate(zip(ps, lrs))]
    opt = torch.optim.Adam(groups, lr=0.0, eps=eps)
    for grads in grads_per_step:
        for p, g in zip(ps, grads):
            p.grad = g.clone()
        opt.step()
    return [p.detach() for p in ps]
