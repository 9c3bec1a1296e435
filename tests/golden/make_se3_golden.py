#!/usr/bin/env python
"""Generate tests/golden/se3_golden.pt by running the REAL reference
scene/rigid_body.py (loaded by file path from /root/reference; build container only).

Also asserts that oracle/rigid_body_port.py reproduces it bit for bit on CPU -
this is what pins the SE3 oracle.  Run: python tests/golden/make_se3_golden.py
"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))
from oracle import rigid_body_port as port  # noqa: E402
import synthetic  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_rigid_body", "/root/reference/scene/rigid_body.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

torch.manual_seed(0)
N = 512
S, theta = synthetic.make_twists(N, seed=2)
x = synthetic.make_scene(N, seed=0)["means3D"]
g = torch.Generator().manual_seed(7)
gy = torch.randn((N, 3), generator=g)
gT = torch.randn((N, 4, 4), generator=g)

out = {}
for name, mod in (("ref", ref), ("port", port)):
    S_ = S.clone().requires_grad_(True)
    th_ = theta.clone().requires_grad_(True)
    x_ = x.clone().requires_grad_(True)
    T = mod.exp_se3(S_, th_)
    y = mod.from_homogenous(torch.bmm(T, mod.to_homogenous(x_).unsqueeze(-1)).squeeze(-1))
    (y * gy).sum().backward()
    res = dict(T=T.detach(), y=y.detach(), dS=S_.grad.clone(), dtheta=th_.grad.clone(), dx=x_.grad.clone())
    # gradient of exp_se3 alone for an arbitrary upstream dT
    S2 = S.clone().requires_grad_(True)
    th2 = theta.clone().requires_grad_(True)
    (mod.exp_se3(S2, th2) * gT).sum().backward()
    res["dS_T"], res["dtheta_T"] = S2.grad.clone(), th2.grad.clone()
    res["skew"] = mod.skew(S[:, :3])
    res["exp_so3"] = mod.exp_so3(S[:, :3], theta)
    out[name] = res

for k in out["ref"]:
    assert torch.equal(out["ref"][k], out["port"][k]), "port differs from reference in %s" % k
print("oracle/rigid_body_port.py == /root/reference/scene/rigid_body.py bit-exact on", sorted(out["ref"]))
torch.save(dict(S=S, theta=theta, x=x, gy=gy, gT=gT, **out["ref"]), os.path.join(ROOT, "tests", "golden", "se3_golden.pt"))
print("wrote tests/golden/se3_golden.pt")
