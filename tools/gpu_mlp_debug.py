import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))
import deform_mlp
from oracle import ref_py
gm = ref_py.gaussian_model()
P = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
torch.manual_seed(5)
ref = gm.DirectTemporalNeRF().cuda()
ours = deform_mlp.DirectTemporalNeRF().cuda()
ours.load_state_dict(ref.state_dict())
ref64 = gm.DirectTemporalNeRF().cuda().double()
ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
g = torch.Generator().manual_seed(P)
x0 = ((torch.rand((P, 3), generator=g) * 2 - 1) * 1.3).cuda()
ts = torch.full((P, 1), 0.61, device="cuda")
proj = [torch.randn((P, c), generator=g).cuda() for c in (3, 3, 4, 48)]
res = {}
for name, net, dt in (("ref", ref, torch.float32), ("ours", ours, torch.float32), ("f64", ref64, torch.float64)):
    x = x0.to(dt).clone().requires_grad_(True)
    outs = net(x, ts.to(dt), 5000)
    sum((o * p.to(dt)).sum() for o, p in zip(outs, proj)).backward()
    torch.cuda.synchronize()
    res[name] = {k: p.grad.double().clone() for k, p in net.named_parameters()}
def r(a, b): return float((a - b).abs().max() / b.abs().max())
for k in res["ref"]:
    t = res["f64"][k]
    line = "%-28s ours-f64 %.2e  ref-f64 %.2e  ours-ref %.2e" % (k, r(res["ours"][k], t), r(res["ref"][k], t), r(res["ours"][k], res["ref"][k]))
    if k == "_time.0.weight":
        line += "  | pos part ours-f64 %.2e time part %.2e" % (r(res["ours"][k][:, :63], t[:, :63]), r(res["ours"][k][:, 63:], t[:, 63:]))
    if k == "_time.5.weight":
        line += "  | embed part ours-f64 %.2e hidden part %.2e" % (r(res["ours"][k][:, :63], t[:, :63]), r(res["ours"][k][:, 63:], t[:, 63:]))
    print(line)
