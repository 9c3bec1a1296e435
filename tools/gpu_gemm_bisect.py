import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))
import gsr_runtime as rt, deform_mlp
from deform_mlp import _Gemm, _planes
dev = "cuda"; P = 1000000
torch.manual_seed(0)
net = deform_mlp.DirectTemporalNeRF().cuda()
W = net._time[1].weight.detach().contiguous(); b = net._time[1].bias.detach().contiguous()
Wp = _planes(W)
g0 = torch.Generator().manual_seed(0)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 4)
out = torch.empty(P, 256, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
res = {}
for name, A in (("randn", torch.randn(P, 256, generator=g0).to(dev)), ("relu_randn", torch.relu(torch.randn(P, 256, generator=g0)).to(dev)),
                ("small_pos", (torch.rand(P, 256, generator=g0) * 0.3).to(dev)), ("relu_small", torch.relu(torch.randn(P, 256, generator=g0) * 0.2).to(dev))):
    res["A=" + name] = t(lambda: _Gemm.run(dev, P, 256, A, 256, Wp, rt.GEMM_RELU_SPLIT, bias=b, out=out, err=err))
A = torch.relu(torch.randn(P, 256, generator=g0) * 0.2).to(dev)
res["no_err"] = t(lambda: _Gemm.run(dev, P, 256, A, 256, Wp, rt.GEMM_RELU_SPLIT, bias=b, out=out))
res["no_bias"] = t(lambda: _Gemm.run(dev, P, 256, A, 256, Wp, rt.GEMM_RELU_SPLIT, out=out, err=err))
Wr = _planes((torch.randn(256, 256, generator=g0) / 16).to(dev))
res["randnW"] = t(lambda: _Gemm.run(dev, P, 256, A, 256, Wr, rt.GEMM_RELU_SPLIT, bias=b, out=out, err=err))
res["bias_randn"] = t(lambda: _Gemm.run(dev, P, 256, A, 256, Wp, rt.GEMM_RELU_SPLIT, bias=torch.randn(256, generator=g0).to(dev), out=out, err=err))
print(json.dumps(res))
# the real forward, per layer timings with events
x0 = ((torch.rand((P, 3), generator=g0) * 2 - 1) * 1.3).cuda()
rt.profile_enable(True)
with torch.no_grad():
    net(x0, 0.4, 5000)
torch.cuda.synchronize()
print(json.dumps({k: (c, round(tt, 3)) for k, (c, tt) in rt.profile_dump().items()}))
