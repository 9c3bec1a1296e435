"""the deformation network on the tensor cores (csrc/mlp_gemm.cu, deform_mlp.py) against
  (1) golden vectors produced by the REAL reference classes (tests/golden/mlp_golden.pt),
  (2) the reference's live torch classes on the same GPU (oracle/_ref/py/scene/gaussian_model.py), at up to 1 M points,
  (3) a float64 evaluation of the same network, next to the reference's own fp32 error.
Tolerance: the reference computes in fp32 (cuBLAS SGEMM).  The tensor-core path must reproduce that fp32 result to
|a - b| <= 1e-4 |b| + 2e-5 max|b| per element, outputs and every gradient - and be no further from the float64 truth than
4x the reference's own fp32 error + 1e-6 of max (hi/lo-split TF32 products: ~2^-22 per product; a single TF32 or BF16
product would be ~1e-3 and fails both)."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL, ATOL = 1e-4, 2e-5
OUT_L2 = 1e-5        # relative L2 distance of the outputs from the float64 evaluation (measured 3.1e-6; the reference's fp32: 2.4e-7)
KINK_L2 = 1e-2       # ... of the gradients at the reference's initialisation, where ReLU flips dominate (measured <= 3e-3; reference <= 2e-3)
_REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "mlp_report.jsonl")


def _close(a, b, what, kink_frac=0.0):
    """`kink_frac`: fraction of elements allowed outside the bar.  Only the gradient w.r.t. the POSITIONS uses it: that
    gradient is discontinuous where a hidden unit's pre-activation crosses zero (ReLU), so for the ~0.1 % of points that
    have one of their 2048 pre-activations within rounding distance of 0, two correct fp32 evaluations pick different
    sides (the reference's own fp32 result shows the same against float64 - asserted below)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu().reshape(a.shape)
    d = (a - b).abs()
    bmax = float(b.abs().max()) + 1e-300
    st = dict(what=what, rel_to_max=float(d.max()) / bmax, violations=int((d > RTOL * b.abs() + ATOL * bmax).sum()), elements=d.numel(),
              median_rel_to_max=float(d.median()) / bmax)
    if os.path.isdir(os.path.dirname(_REPORT)):
        with open(_REPORT, "a") as f:
            f.write(json.dumps(st) + "\n")
    assert st["violations"] <= kink_frac * d.numel(), st
    return st


def _digest_close(t, d, what):
    f = t.detach().double().reshape(-1).cpu()
    assert tuple(t.shape) == d["shape"], what
    assert abs(float(f.sum()) - d["sum"]) <= 2e-5 * max(1e-30, d["abs_sum"]), (what, float(f.sum()), d["sum"], d["abs_sum"])
    assert abs(float(f.abs().sum()) - d["abs_sum"]) <= 1e-4 * d["abs_sum"], what
    scale = d["abs_sum"] / max(1, f.numel())
    for got, want in ((f[:16], d["head"].double()), (f[-16:], d["tail"].double())):
        assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-3 * scale, what


def test_embedding_is_the_reference_embedding():
    import ctypes
    import gsr_runtime as rt
    g = torch.load(os.path.join(GOLD, "mlp_golden.pt"), weights_only=False)
    x = g["x"].cuda()
    P = x.shape[0]
    e_hi, e_lo = torch.empty(P, 64, device="cuda"), torch.empty(P, 64, device="cuda")
    rt.check(rt.load().gsr_mlp_embed(x.data_ptr(), P, e_hi.data_ptr(), e_lo.data_ptr(), None, None, 0, rt.stream_ptr("cuda")))
    e = (e_hi + e_lo)[:, :63].cpu()
    assert float((e - g["embed_x"]).abs().max()) <= 2.4e-7          # sinf / cosf: the same libdevice routines torch uses, <= 2 ulp of 1
    assert bool(((e_hi.view(torch.int32) & 0x1fff) == 0).all()) and float(e_hi[:, 63].abs().max() + e_lo[:, 63].abs().max()) == 0.0


def test_deform_mlp_vs_reference_golden():
    import deform_mlp
    g = torch.load(os.path.join(GOLD, "mlp_golden.pt"), weights_only=False)
    torch.manual_seed(g["seed"])
    net = deform_mlp.DirectTemporalNeRF()
    assert sum(p.numel() for p in net.parameters()) == 513338
    sd = net.state_dict()
    assert list(sd.keys()) == list(g["params"].keys())             # the reference's state_dict names, in its order
    for k, p in net.named_parameters():                              # and the same initial values for the same seed
        assert torch.equal(p.detach().reshape(-1)[:16], g["params"][k]["head"]), k
    net = net.cuda()
    x = g["x"].cuda().requires_grad_(True)
    outs = net(x, g["ts"].cuda(), g["iteration"])
    assert [tuple(o.shape) for o in outs] == [tuple(o.shape) for o in g["outs"]]
    for i, (o, want) in enumerate(zip(outs, g["outs"])):
        _close(o, want, "golden out%d" % i)
    sum((o * p.cuda()).sum() for o, p in zip(outs, g["proj"])).backward()
    _close(x.grad, g["dx"], "golden dx")
    for k, p in net.named_parameters():
        _digest_close(p.grad, g["dparams"][k], k)
    early = net(x.detach(), g["ts"].cuda(), 100)                     # iteration < 3000: zeros of the reference's shapes
    for o, want in zip(early, g["outs_early"]):
        assert o.shape == want.shape and float(o.abs().sum()) == 0.0
    with pytest.raises(AssertionError, match="same time"):
        ts = g["ts"].clone().cuda(); ts[3] = 0.9
        net(x.detach(), ts, 5000)
    # the _se3 producer (row a3): raw heads -> (S, theta) -> transform
    torch.manual_seed(2)
    se3 = deform_mlp.DirectTemporalNeRF_se3(input_ch=63, input_ch_time=21).cuda()
    xe = g["embed_x"].cuda()
    w, v = se3.raw_heads(xe, g["ts"].cuda())
    _close(w, g["se3"]["w_raw"], "golden se3 w_raw")
    _close(v, g["se3"]["v_raw"], "golden se3 v_raw")
    S, th = se3.screw(xe, g["ts"].cuda())
    _close(S, g["se3"]["S"], "golden se3 S")
    _close(th, g["se3"]["theta"], "golden se3 theta")
    T = se3(xe, g["ts"].cuda(), 5000)
    assert T.shape == g["se3"]["transform"].shape
    assert float((T.cpu() - g["se3"]["transform"]).abs().max()) <= 2e-6
    # theta -> 0 policy: the reference divides by |w| unguarded (NaN); with theta_eps the point gets the identity
    S0, t0 = deform_mlp.screw_from_raw(torch.zeros(2, 3, device="cuda"), torch.ones(2, 3, device="cuda"), eps=1e-12)
    assert float(S0.abs().sum()) == 0.0 and float(t0.abs().sum()) == 0.0
    Sn, _ = deform_mlp.screw_from_raw(torch.zeros(2, 3, device="cuda"), torch.ones(2, 3, device="cuda"))
    assert bool(torch.isnan(Sn).any() | torch.isinf(Sn).any())


def _run_net(net, x0, ts, proj, dtype=torch.float32):
    x = x0.to(dtype).clone().requires_grad_(True)
    for p in net.parameters():
        p.grad = None
    outs = net(x, ts.to(dtype), 5000)
    sum((o * p.to(dtype)).sum() for o, p in zip(outs, proj)).backward()
    torch.cuda.synchronize()
    return [o.detach() for o in outs], x.grad, {k: p.grad.clone() for k, p in net.named_parameters()}


def _hidden_activations(net, x0, ts):
    hs = []
    hooks = [l.register_forward_hook(lambda m, i, o: hs.append(o.detach())) for l in net._time]
    with torch.no_grad():
        net(x0[:20000], ts[:20000], 5000)
    for h in hooks:
        h.remove()
    return hs


def _l2(a, b):
    a, b = a.double(), b.double().reshape(a.shape)
    return float((a - b).norm() / (b.norm() + 1e-300))


@pytest.mark.parametrize("P", [1000, 100003, 1000000])
def test_deform_mlp_vs_live_reference_classes(P):
    """Forward: the bar above on every output.  Backward, twice:
    (a) with every hidden unit active (biases shifted so that no pre-activation is near zero) the gradients are smooth
        functions of the inputs and must meet the bar element by element - this is the test of the backward kernels;
    (b) with the reference's own initialisation the gradients are NOT smooth: a unit whose pre-activation is within
        rounding distance of zero is switched on in one fp32 evaluation and off in another, and one such flip moves a
        bias gradient by a whole term of its sum (~1 / sqrt(P) of its value).  The reference's own fp32 gradients sit
        4e-4 .. 2e-3 (relative L2) from the float64 gradients for that reason (measured, profiles/r02_mlp_parity.md).  The
        number of flips grows with the rounding error of the pre-activations (tensor-core path: 1.4e-6 of max per layer,
        cuBLAS SGEMM: 4.5e-7 - the tensor core accumulates with truncation) and the L2 error with its square root: measured
        1.5e-3 .. 3e-3 where the reference has 1e-4 .. 2e-3.  Required: relative L2 distance from float64 <= 1e-2 (both
        numbers are recorded); the heads (no ReLU between them and the loss) still element by element."""
    import deform_mlp
    from oracle import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/py not staged")
    gm = ref_py.gaussian_model()
    torch.manual_seed(5)
    ref = gm.DirectTemporalNeRF().cuda()
    ours = deform_mlp.DirectTemporalNeRF().cuda()
    ours.load_state_dict(ref.state_dict())                           # the reference's offset_model.pth layout
    g = torch.Generator().manual_seed(P)
    x0 = ((torch.rand((P, 3), generator=g) * 2 - 1) * 1.3).cuda()
    ts = torch.full((P, 1), 0.61, device="cuda")
    proj = [torch.randn((P, c), generator=g).cuda() for c in (3, 3, 4, 48)]
    r, o = _run_net(ref, x0, ts, proj), _run_net(ours, x0, ts, proj)
    for i in range(4):
        _close(o[0][i], r[0][i], "P=%d out%d" % (P, i))
    for k in r[2]:
        if k.startswith("_time_out"):
            _close(o[2][k], r[2][k], "P=%d d%s" % (P, k))
    # (b) against float64
    ref64 = gm.DirectTemporalNeRF().cuda().double()
    ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    t = _run_net(ref64, x0, ts, proj, torch.float64)
    rows = []
    for k in r[2]:
        e_ref, e_our = _l2(r[2][k], t[2][k]), _l2(o[2][k], t[2][k])
        rows.append(dict(what="P=%d kinked d%s vs float64 (rel L2)" % (P, k), ours=e_our, reference_fp32=e_ref))
        assert e_our <= KINK_L2, rows[-1]
    e_ref, e_our = _l2(r[1], t[1]), _l2(o[1], t[1])
    rows.append(dict(what="P=%d kinked dx vs float64 (rel L2)" % P, ours=e_our, reference_fp32=e_ref))
    assert e_our <= KINK_L2, rows[-1]
    for i in range(4):
        e_ref, e_our = _l2(r[0][i], t[0][i]), _l2(o[0][i], t[0][i])
        rows.append(dict(what="P=%d out%d vs float64 (rel L2)" % (P, i), ours=e_our, reference_fp32=e_ref))
        assert e_our <= OUT_L2, rows[-1]
    del t, ref64
    # (a) every unit active (weights / 4, biases + 4: pre-activations stay 7 sigma above zero): smooth gradients,
    # element-by-element bar
    with torch.no_grad():
        for net in (ref, ours):
            for layer in net._time:
                layer.weight.mul_(0.25)
                layer.bias.add_(4.0)
    r, o = _run_net(ref, x0, ts, proj), _run_net(ours, x0, ts, proj)
    assert all(float(h.min()) > 0 for h in _hidden_activations(ref, x0, ts))
    for i in range(4):
        _close(o[0][i], r[0][i], "P=%d all-active out%d" % (P, i))
    _close(o[1], r[1], "P=%d all-active dx" % P)
    for k in r[2]:
        _close(o[2][k], r[2][k], "P=%d all-active d%s" % (P, k))
    if os.path.isdir(os.path.dirname(_REPORT)):
        with open(_REPORT, "a") as f:
            for row in rows:
                f.write(json.dumps(row) + "\n")


def test_deform_mlp_inside_the_reference_render():
    """The network as GaussianModel.offset_model under the reference's real render() (gaussian_renderer/__init__.py:79)."""
    import deform_mlp
    import diff_gaussian_rasterization as ours_ras
    import synthetic
    from types import SimpleNamespace
    from oracle import ref_py
    from test_gpu_contract import _leaves, _real_gaussian_model
    if not ref_py.available():
        pytest.skip("oracle/_ref/py not staged")
    gm = ref_py.gaussian_model()
    render = ref_py.render_fn(ours_ras)
    P, W, H = 40000, 480, 320
    pc = _real_gaussian_model(gm, P, seed=9)
    native = deform_mlp.DirectTemporalNeRF().cuda()
    native.load_state_dict(pc.offset_model.state_dict())
    cam = synthetic.make_camera(2, 8, W, H, device="cuda")
    cam.time = 0.25
    pipe = SimpleNamespace(debug=False, convert_SHs_python=False, compute_cov3D_python=False)
    bg = torch.zeros(3, device="cuda")
    gimg = synthetic.make_image_grad(W, H, device="cuda")
    res = []
    torch_net = pc.offset_model
    for net in (torch_net, native):
        pc.offset_model = net
        for _, p in _leaves(pc):
            p.grad = None
        out = render(cam, pc, pipe, bg, iteration=5000)
        (out["render"] * gimg).sum().backward()
        res.append((out["render"].detach().clone(), out["radii"].clone(), out["means3D"].detach().clone(),
                    {n: p.grad.clone() for n, p in _leaves(pc)}))
    # the rasterizer's tile decisions are discontinuous in the means: compare what is continuous
    _close(res[1][2], res[0][2], "render() means3D with the native network")
    assert float((res[1][1] != res[0][1]).float().mean()) <= 1e-3              # radii: a handful of ceil() flips at most
    assert float((res[1][0] - res[0][0]).abs().mean()) <= 1e-5
    for n in ("xyz", "scaling", "opacity"):
        a, b = res[1][3][n].double(), res[0][3][n].double()
        assert float((a - b).abs().mean() / b.abs().mean()) <= 1e-3, n


def test_fused_activation_glue_matches_the_reference_render_ops():
    """gaussian_renderer/__init__.py:79,116,122,140 as one kernel each way, against the same torch ops."""
    import deform_mlp
    P = 50001
    g = torch.Generator().manual_seed(4)
    mk = lambda *s: torch.randn(*s, generator=g).cuda()
    xyz, scaling, rotation, f_dc, f_rest = mk(P, 3), mk(P, 3) * 0.5 - 4, mk(P, 4), mk(P, 1, 3), mk(P, 15, 3) * 0.2
    heads = torch.zeros(P, 64, device="cuda")
    heads[:, :58] = mk(P, 58) * 0.1
    gout = [mk(P, 3), mk(P, 3), mk(P, 4), mk(P, 16, 3)]
    res = []
    for fused in (False, True):
        leaves = [t.clone().requires_grad_(True) for t in (heads, xyz, scaling, rotation, f_dc, f_rest)]
        h, x, sc, ro, dc, fr = leaves
        if fused:
            outs = deform_mlp.deform_glue(h, x, sc, ro, dc, fr)
        else:
            outs = (x + h[:, 0:3], torch.exp(sc + h[:, 3:6]), torch.nn.functional.normalize(ro + h[:, 6:10]),
                    torch.cat((dc, fr), dim=1) + h[:, 10:58].reshape(-1, 16, 3))
        torch.autograd.backward(outs, gout)
        res.append(([o.detach() for o in outs], [l.grad for l in leaves]))
    for a, b in zip(res[1][0], res[0][0]):
        assert a.shape == b.shape and float((a - b).abs().max()) <= 2e-6 * float(b.abs().max())
    for a, b in zip(res[1][1], res[0][1]):
        assert a.shape == b.shape and float((a - b).abs().max()) <= 2e-6 * float(b.abs().max()) + 1e-9
    # the network's raw output feeds it directly
    torch.manual_seed(1)
    net = deform_mlp.DirectTemporalNeRF().cuda()
    hh = net.heads(xyz, 0.5)
    dx = net(xyz, 0.5, 5000)[0]
    assert hh.shape == (P, 64) and torch.equal(hh[:, 0:3], dx) and float(hh[:, 58:].abs().max()) == 0.0
    assert float(net.heads(xyz, 0.5, 100).abs().max()) == 0.0
