#!/usr/bin/env python
"""Stand-alone check of gsr_mlp_gemm (tcgen05 3xTF32) against fp64, next to torch's own fp32 matmul."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_deformable_b200"))
import gsr_runtime as rt

lib = rt.load()
dev = "cuda"
st = lambda: rt.stream_ptr(dev)


def split(x):
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    rt.check(lib.gsr_mlp_split(x.data_ptr(), x.numel(), hi.data_ptr(), lo.data_ptr(), st()))
    return hi, lo


def gemm(A, B, mode, bias=None, A1=None, mask=None, k_splits=1, want_T=False, colsum=False, out=None, singleA=False,
         singleB=False, plain_out=False):
    M, K0 = A.shape
    N = B.shape[0]
    a_hi, a_lo = (A, None) if singleA else split(A)
    b_hi, b_lo = (B, None) if singleB else split(B)
    g = rt.gsr_gemm()
    g.M, g.N = M, N
    g.A0_hi, g.A0_lo, g.K0, g.ldA0 = a_hi.data_ptr(), (a_lo.data_ptr() if a_lo is not None else None), K0, A.stride(0)
    keep = [a_hi, a_lo, b_hi, b_lo]
    if A1 is not None:
        h, l = (A1, None) if singleA else split(A1); keep += [h, l]
        g.A1_hi, g.A1_lo, g.K1, g.ldA1 = h.data_ptr(), (l.data_ptr() if l is not None else None), A1.shape[1], A1.stride(0)
    g.B_hi, g.B_lo, g.ldB = b_hi.data_ptr(), (b_lo.data_ptr() if b_lo is not None else None), B.stride(0)
    g.mode, g.k_splits = mode, k_splits
    if bias is not None:
        g.bias = bias.data_ptr()
    if mask is not None:
        g.mask_src, g.ld_mask = mask.data_ptr(), mask.stride(0)
    o_hi = torch.zeros(M, N, device=dev) if out is None else out
    o_lo = torch.zeros(M, N, device=dev)
    g.out_hi, g.out_lo, g.ld_out = o_hi.data_ptr(), (None if (plain_out or mode in (rt.GEMM_PLAIN, rt.GEMM_ATOMIC)) else o_lo.data_ptr()), N
    Mp = (M + 3) // 4 * 4
    t_hi = t_lo = None
    if want_T:
        t_hi, t_lo = torch.zeros(N, Mp, device=dev), torch.zeros(N, Mp, device=dev)
        g.outT_hi, g.outT_lo, g.ld_outT = t_hi.data_ptr(), (None if plain_out else t_lo.data_ptr()), Mp
    cs = None
    if colsum:
        cs = torch.zeros(N, device=dev)
        g.colsum = cs.data_ptr()
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    g.error_flag = err.data_ptr()
    rt.check(lib.gsr_mlp_gemm(ctypes.byref(g), st()))
    torch.cuda.synchronize()
    return dict(hi=o_hi, lo=o_lo, t_hi=t_hi, t_lo=t_lo, colsum=cs, err=int(err.item()), Mp=Mp)


def rel(a, b):
    return float((a.double() - b).abs().max() / (b.abs().max() + 1e-300))


res = []
gen = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=gen).to(dev)
for (M, N, K) in ((128, 256, 32), (128, 256, 64), (300, 256, 64), (4096, 256, 256), (1000, 64, 256), (777, 58, 256), (5000, 128, 96)):
    A, B, bias = rn(M, K), rn(N, K) / K ** 0.5, rn(N)
    ref = A.double() @ B.double().t() + bias.double()
    o = gemm(A, B, rt.GEMM_PLAIN, bias=bias)
    t32 = A @ B.t() + bias
    res.append(dict(case="plain M%d N%d K%d" % (M, N, K), err_flag=o["err"], ours=rel(o["hi"], ref), torch_fp32=rel(t32, ref)))
    print(res[-1], flush=True)
# single-plane operands (split in shared memory by the converter warps)
for (M, N, K, sa, sb) in ((4096, 256, 256, True, False), (1000, 64, 256, True, False), (256, 256, 50000, True, True), (3000, 256, 64, True, True)):
    A, B, bias = rn(M, K), rn(N, K) / K ** 0.5, rn(N)
    ref = torch.relu(A.double() @ B.double().t() + bias.double())
    o = gemm(A, B, rt.GEMM_RELU_SPLIT, bias=bias, singleA=sa, singleB=sb, plain_out=True, want_T=True, k_splits=1)
    res.append(dict(case="single-plane relu M%d N%d K%d A%d B%d" % (M, N, K, sa, sb), err_flag=o["err"], ours=rel(o["hi"], ref),
                    transposed=rel(o["t_hi"][:, :M].t(), ref)))
    print(res[-1], flush=True)
# relu + planes + transposed + colsum
M, N, K = 1000, 256, 256
A, B, bias = rn(M, K), rn(N, K) / K ** 0.5, rn(N)
ref = torch.relu(A.double() @ B.double().t() + bias.double())
o = gemm(A, B, rt.GEMM_RELU_SPLIT, bias=bias, want_T=True, colsum=True)
v = o["hi"].double() + o["lo"].double()
vt = (o["t_hi"].double() + o["t_lo"].double())[:, :M].t()
res.append(dict(case="relu_split", err_flag=o["err"], ours=rel(v, ref), transposed=rel(vt, ref), colsum=rel(o["colsum"], ref.sum(0)),
                hi_is_tf32=bool(((o["hi"].view(torch.int32) & 0x1fff) == 0).all())))
print(res[-1], flush=True)
# two K segments + mask
M, N, K0, K1 = 900, 256, 64, 256
A0, A1 = rn(M, K0), rn(M, K1)
B = rn(N, K0 + K1) / (K0 + K1) ** 0.5
mask = rn(M, N)
ref = (torch.cat([A0, A1], 1).double() @ B.double().t()) * (mask.double() > 0)
o = gemm(A0, B, rt.GEMM_SPLIT, A1=A1, mask=mask, colsum=True)
res.append(dict(case="two_segments_mask", err_flag=o["err"], ours=rel(o["hi"].double() + o["lo"].double(), ref), colsum=rel(o["colsum"], ref.sum(0))))
print(res[-1], flush=True)
# split-K atomic (weight gradient shape: M = 256 outputs, K = many points)
for (M, N, K, S) in ((256, 256, 100000, 74), (256, 64, 33333 // 4 * 4, 50), (64, 256, 20000, 16)):
    A, B = rn(M, K), rn(N, K)
    ref = A.double() @ B.double().t()
    o = gemm(A, B, rt.GEMM_ATOMIC, k_splits=S)
    res.append(dict(case="atomic M%d N%d K%d S%d" % (M, N, K, S), err_flag=o["err"], ours=rel(o["hi"], ref), torch_fp32=rel(A @ B.t(), ref)))
    print(res[-1], flush=True)
# MN-major operands: C += A^T B with A [K x M], B [K x N] row-major (weight gradients from row-major activations)
for (K, M, N, S) in ((5000, 256, 256, 8), (100000, 256, 256, 100), (33333, 256, 64, 40), (20000, 58, 256, 20), (777, 128, 64, 3)):
    A, B = rn(K, (M + 3) // 4 * 4)[:, :M], rn(K, N)
    ref = A.double().t() @ B.double()
    out = torch.zeros(M, N, device=dev)
    g = rt.gsr_gemm(); g.M, g.N = M, N
    g.A0_hi, g.K0, g.ldA0 = A.data_ptr(), K, A.stride(0)
    g.B_hi, g.ldB = B.data_ptr(), B.stride(0)
    g.mode, g.k_splits, g.mn_major = rt.GEMM_ATOMIC, S, 1
    g.out_hi, g.ld_out = out.data_ptr(), N
    err = torch.zeros(1, dtype=torch.int32, device=dev); g.error_flag = err.data_ptr()
    rt.check(lib.gsr_mlp_gemm(ctypes.byref(g), st())); torch.cuda.synchronize()
    res.append(dict(case="mn-major atomic K%d M%d N%d S%d" % (K, M, N, S), err_flag=int(err.item()), ours=rel(out, ref), torch_fp32=rel(A.t() @ B, ref)))
    print(res[-1], flush=True)
# timing: one hidden layer at 1 M points
M, N, K = 1000000, 256, 256
A, B, bias = rn(M, K), rn(N, K) / 16, rn(N)
b_hi, b_lo = split(B)
o_hi = torch.empty(M, N, device=dev)
g = rt.gsr_gemm(); g.M, g.N = M, N
g.A0_hi, g.A0_lo, g.K0, g.ldA0 = A.data_ptr(), None, K, K
g.B_hi, g.B_lo, g.ldB = b_hi.data_ptr(), b_lo.data_ptr(), K
g.mode, g.k_splits, g.bias = rt.GEMM_RELU_SPLIT, 1, bias.data_ptr()
g.out_hi, g.out_lo, g.ld_out = o_hi.data_ptr(), None, N
for _ in range(3):
    rt.check(lib.gsr_mlp_gemm(ctypes.byref(g), st()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    rt.check(lib.gsr_mlp_gemm(ctypes.byref(g), st()))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
e0.record()
for _ in range(10):
    y = torch.relu(torch.addmm(bias, A, B.t()))
e1.record(); torch.cuda.synchronize()
ms_t = e0.elapsed_time(e1) / 10
res.append(dict(case="timing 1M x 256 x 256 relu layer", ours_ms=ms, torch_fp32_ms=ms_t, ours_tflops_3x=3 * 2 * M * N * K / ms / 1e9,
                useful_tflops=2 * M * N * K / ms / 1e9))
print(res[-1], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gemm_check.json"), "w"), indent=1)
