"""Generates tests/golden/mlp_golden.pt from the REAL deformation-network classes of the reference.
scene/gaussian_model.py cannot be imported here (plyfile / FrEIA / simple_knn are absent), so the source of exactly
the classes and functions needed (Embedder, get_embedder, DirectTemporalNeRF, DirectTemporalNeRF_se3) is cut out of
the file with `ast` and executed unmodified; `rigid` is the real scene/rigid_body.py loaded by path.
Run in the build container:  python tests/golden/make_mlp_golden.py"""
import ast
import importlib.util
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/scene/gaussian_model.py"
spec = importlib.util.spec_from_file_location("ref_rigid_body", "/root/reference/scene/rigid_body.py")
rigid = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rigid)

text = open(SRC).read()
tree = ast.parse(text)
want = {"Embedder", "get_embedder", "DirectTemporalNeRF", "DirectTemporalNeRF_se3"}
ns = {"torch": torch, "nn": nn, "F": F, "rigid": rigid}
for node in tree.body:
    if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in want:
        exec(compile(ast.get_source_segment(text, node), SRC, "exec"), ns)

torch.manual_seed(0)
net = ns["DirectTemporalNeRF"]()                     # D=8, W=256, embedders 10/10 (gaussian_model.py:680)
g = torch.Generator().manual_seed(1)
N = 257
x = (torch.rand((N, 3), generator=g) * 2.6 - 1.3).requires_grad_(True)
ts = torch.full((N, 1), 0.37)
outs = net(x, ts, 5000)
proj = [torch.randn(o.shape, generator=g) for o in outs]
sum((o * p).sum() for o, p in zip(outs, proj)).backward()
def digest(t):           # enough to pin a 65k-element gradient without storing it
    f = t.detach().double().reshape(-1)
    return {"shape": tuple(t.shape), "sum": float(f.sum()), "abs_sum": float(f.abs().sum()), "head": t.detach().reshape(-1)[:16].clone(),
            "tail": t.detach().reshape(-1)[-16:].clone()}


# weights are NOT stored: nn.Linear initialises from the global RNG in construction order, which the port reproduces
# (torch.manual_seed(0), same layer order), so the same seed gives the same 513 338 parameters; their digests pin that
gold = {"seed": 0, "x": x.detach().clone(), "ts": ts, "iteration": 5000,
        "outs": [o.detach().clone() for o in outs], "proj": proj, "dx": x.grad.clone(),
        "params": {k: digest(p) for k, p in net.named_parameters()},
        "dparams": {k: digest(p.grad) for k, p in net.named_parameters()},
        "outs_early": [o.detach().clone() for o in net(x.detach(), ts, 100)]}
# the embedding alone
emb, dim = ns["get_embedder"](10, 3, 0)
gold["embed_x"] = emb(x.detach())
gold["embed_dim"] = dim
# the se3 variant: raw heads -> (S, theta) -> transform
torch.manual_seed(2)
se3 = ns["DirectTemporalNeRF_se3"](input_ch=63, input_ch_time=21)
xe = gold["embed_x"]
te = ns["get_embedder"](10, 1, 0)[0](ts)
w, v = se3.query_time(xe, te, se3._time, se3._w, se3._v)
theta = torch.norm(w, dim=-1)
S = torch.cat([w / theta[..., None], v / theta[..., None]], dim=-1)
gold["se3"] = {"w_raw": w.detach().clone(), "v_raw": v.detach().clone(), "S": S.detach().clone(), "theta": theta.detach().clone(),
               "transform": se3(xe, te, 5000).detach().clone()}
torch.save(gold, os.path.join(HERE, "mlp_golden.pt"))
print("wrote mlp_golden.pt", [tuple(o.shape) for o in outs], float(outs[0].abs().mean()))
