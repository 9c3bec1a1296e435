"""Generates tests/golden/loss_golden.pt by importing the REAL reference utils/loss_utils.py (pure torch) from
/root/reference and running it on CPU, and torch.optim.Adam configured as scene/gaussian_model.py:834-846 does.
Run in the build container (the reference does not exist on the GPU box):  python tests/golden/make_loss_golden.py"""
import importlib.util
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_loss_utils", "/root/reference/utils/loss_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {"cases": []}
g = torch.Generator().manual_seed(0)
for (H, W, kind) in ((37, 53, "rand"), (16, 16, "rand"), (64, 96, "smooth"), (11, 5, "rand"), (40, 40, "equal")):
    if kind == "smooth":
        yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
        img = torch.stack([xx, yy, (xx * yy)]) + 0.05 * torch.rand((3, H, W), generator=g)
        gt = torch.stack([xx * 0.9, yy * 1.1, (xx + yy) * 0.5]).clamp(0, 1)
    else:
        img = torch.rand((3, H, W), generator=g)
        gt = torch.rand((3, H, W), generator=g)
    if kind == "equal":
        gt = img.clone()
        gt[:, ::3, ::2] += 0.25            # some exact zeros of (img - gt): torch.abs backward gives 0 there
    x = img.clone().requires_grad_(True)
    lam = 0.2                               # arguments/__init__.py:83
    Ll1 = ref.l1_loss(x, gt)
    s = ref.ssim(x, gt)
    loss = (1.0 - lam) * Ll1 + lam * (1.0 - s)          # train.py:529
    loss.backward()
    x2 = img.clone().requires_grad_(True)
    ref.ssim(x2, gt).backward()
    out["cases"].append({"image": img, "gt": gt, "lambda_dssim": lam, "l1": Ll1.detach(), "ssim": s.detach(),
                         "loss": loss.detach(), "dloss_dimage": x.grad.clone(), "dssim_dimage": x2.grad.clone()})

# Adam: three tensors with the reference's group structure (one group per tensor, eps=1e-15, lr per group), 5 steps
ps = [torch.randn((101, 3), generator=g), torch.randn((101, 16, 3), generator=g) * 0.1, torch.randn((101, 1), generator=g)]
lrs = [0.00016, 0.0025 / 20.0, 0.05]                                    # arguments/__init__.py position/feature/opacity lrs
grads = [[torch.randn(p.shape, generator=g) * (10.0 ** (-(s % 3))) for p in ps] for s in range(5)]
params = [p.clone().requires_grad_(True) for p in ps]
opt = torch.optim.Adam([{"params": [p], "lr": lr, "name": "g%d" % i} for i, (p, lr) in enumerate(zip(params, lrs))],
                       lr=0.0, eps=1e-15)
for s in range(5):
    for p, gr in zip(params, grads[s]):
        p.grad = gr.clone()
    if s == 3:
        opt.param_groups[0]["lr"] = 0.00008                               # update_learning_rate writes the group's lr
    opt.step()
out["adam"] = {"params": ps, "lrs": lrs, "grads": grads, "lr0_from_step3": 0.00008, "final": [p.detach().clone() for p in params],
               "torch": torch.__version__}
# learning-rate schedule (utils/general_utils.py:29-62) with the two configurations scene/gaussian_model.py:857-864 uses
spec2 = importlib.util.spec_from_file_location("ref_general_utils", "/root/reference/utils/general_utils.py")
gu = importlib.util.module_from_spec(spec2)
spec2.loader.exec_module(gu)
steps = [-1, 0, 1, 10, 100, 999, 1000, 5000, 20000, 39999, 40000, 50000]
cfgs = {"xyz": dict(lr_init=0.00016 * 5.0, lr_final=0.0000016 * 5.0, lr_delay_mult=0.01, max_steps=40000),
        "offset": dict(lr_init=8e-4, lr_final=1.6e-6, max_steps=40000),
        "delayed": dict(lr_init=1e-2, lr_final=1e-4, lr_delay_steps=500, lr_delay_mult=0.1, max_steps=3000),
        "disabled": dict(lr_init=0.0, lr_final=0.0)}
out["lr_schedule"] = {"steps": steps, "cfgs": cfgs,
                      "values": {k: [float(gu.get_expon_lr_func(**c)(s)) for s in steps] for k, c in cfgs.items()}}
torch.save(out, os.path.join(HERE, "loss_golden.pt"))
print("wrote", os.path.join(HERE, "loss_golden.pt"), [float(c["loss"]) for c in out["cases"]])
