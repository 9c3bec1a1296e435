import ctypes, json, os, sys, subprocess, threading
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))
import gsr_runtime as rt
lib = rt.load(); dev = "cuda"
st = lambda: rt.stream_ptr(dev)
M, N, K = 1000000, 256, 256
g0 = torch.Generator().manual_seed(0)
A = torch.randn(M, K, generator=g0).to(dev); B = (torch.randn(N, K, generator=g0) / 16).to(dev); bias = torch.randn(N, generator=g0).to(dev)
b_hi, b_lo = torch.empty_like(B), torch.empty_like(B)
rt.check(lib.gsr_mlp_split(B.data_ptr(), B.numel(), b_hi.data_ptr(), b_lo.data_ptr(), st()))
outs = [torch.empty(M, N, device=dev) for _ in range(2)]
def make(out, a):
    g = rt.gsr_gemm(); g.M, g.N = M, N
    g.A0_hi, g.A0_lo, g.K0, g.ldA0 = a.data_ptr(), None, K, K
    g.B_hi, g.B_lo, g.ldB = b_hi.data_ptr(), b_lo.data_ptr(), K
    g.mode, g.k_splits, g.bias = rt.GEMM_RELU_SPLIT, 1, bias.data_ptr()
    g.out_hi, g.out_lo, g.ld_out = out.data_ptr(), None, N
    return g
gs = [make(outs[0], A), make(outs[1], outs[0]), make(outs[0], outs[1])]
def run(n, chain):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        rt.check(lib.gsr_mlp_gemm(ctypes.byref(gs[(1 + i % 2) if (chain and i) else 0]), st()))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
q = "clocks.sm,power.draw"
for n, chain in ((10, False), (100, False), (400, False), (10, True), (200, True)):
    ms = run(n, chain)
    smi = subprocess.run(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    print(json.dumps(dict(n=n, chain=chain, ms=ms, smi_after=smi)), flush=True)
