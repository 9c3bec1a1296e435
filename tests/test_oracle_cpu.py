"""Pin the CPU oracle (oracle/gsr_oracle.c, oracle/rigid_body_port.py) against golden
vectors produced by the REFERENCE itself:
  tests/golden/raster_golden.npz  <- reference CUDA rasterizer + simple-knn run on a B200
                                     (tests/golden/make_raster_golden.py)
  tests/golden/se3_golden.pt      <- reference scene/rigid_body.py (tests/golden/make_se3_golden.py)
No GPU needed."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle_c, rigid_body_port

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "raster_golden.npz"))


class _RS:
    pass


def _settings(g, case):
    P, W, H, deg, R = [int(x) for x in g[case + "/cfg"]]
    rs = _RS()
    rs.image_width, rs.image_height, rs.sh_degree = W, H, deg
    rs.tanfovx, rs.tanfovy = [float(x) for x in g[case + "/tanfov"]]
    rs.scale_modifier = float(g[case + "/scale_modifier"][0])
    rs.bg, rs.viewmatrix, rs.projmatrix, rs.campos = (g[case + "/" + k] for k in ("bg", "viewmatrix", "projmatrix", "campos"))
    return rs, P, W, H, R


def _inputs(g, case):
    kw = {}
    for k in ("shs", "scales", "rotations", "colors_precomp", "cov3D_precomp"):
        if case + "/" + k in g.files:
            kw[k] = g[case + "/" + k]
    return kw


def _run(g, case):
    rs, P, W, H, R = _settings(g, case)
    return rs, oracle_c.forward(rs, g[case + "/means3D"], g[case + "/opacities"], **_inputs(g, case)), R


@pytest.mark.parametrize("case", ["sh3", "sh1", "precomp"])
def test_oracle_forward_matches_reference(gold, case):
    g = gold
    rs, f, R = _run(g, case)
    vis = g[case + "/radii"] > 0
    # integer / bit work: exact
    assert f["num_rendered"] == R
    np.testing.assert_array_equal(f["radii"], g[case + "/radii"])
    np.testing.assert_array_equal(f["tiles_touched"].astype(np.int64), g[case + "/tiles_touched"].astype(np.int64))
    np.testing.assert_array_equal(f["keys_sorted"].view(np.int64), g[case + "/keys_sorted"])
    np.testing.assert_array_equal(f["point_list"].astype(np.int64), g[case + "/point_list"].astype(np.int64))
    np.testing.assert_array_equal(f["ranges"].astype(np.int64), g[case + "/ranges"].astype(np.int64))
    # the geometry chain reproduces the reference's roundings: bit-exact floats
    for k in ("depths", "means2D", "conic_opacity"):
        a = f[k][vis].view(np.int32)
        b = g[case + "/" + k][vis].view(np.int32)
        assert (a == b).all(), k
    if case != "precomp":
        assert (f["cov3D"][vis].view(np.int32) == g[case + "/cov3D"][vis].view(np.int32)).all()
        np.testing.assert_allclose(f["rgb"][vis], g[case + "/rgb"][vis], rtol=0, atol=2e-6)
        np.testing.assert_array_equal(f["clamped"][vis].astype(bool), g[case + "/clamped"][vis].astype(bool))
    # blending uses the host libm's expf: tolerance, not bit-exact
    np.testing.assert_allclose(f["color"], g[case + "/color"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(f["final_T"], g[case + "/final_T"], rtol=0, atol=1e-6)
    mism = (f["n_contrib"].astype(np.int64) != g[case + "/n_contrib"].astype(np.int64)).mean()
    assert mism < 1e-3, "n_contrib mismatch fraction %g" % mism


@pytest.mark.parametrize("case", ["sh3", "sh1", "precomp"])
def test_oracle_backward_matches_reference(gold, case):
    g = gold
    rs, f, R = _run(g, case)
    b = oracle_c.backward(f, g[case + "/grad_image"])
    names = ["means2D", "opacities", "means3D"]
    if case == "precomp":
        names += ["colors", "cov3D"]
    else:
        names += ["shs", "scales", "rotations"]
    for k in names:
        ref = g[case + "/grad_" + k].reshape(b[k].shape)
        scale = max(float(np.abs(ref).max()), 1e-12)
        err = float(np.abs(b[k] - ref).max()) / scale
        assert err < 1e-4, "%s: relative-to-max error %g" % (k, err)


def test_oracle_mark_visible_and_msb(gold):
    g = gold
    got = oracle_c.mark_visible(g["sh3/means3D"], g["sh3/viewmatrix"], g["sh3/projmatrix"])
    np.testing.assert_array_equal(got, g["sh3/mark_visible"].astype(bool))
    # rasterizer_impl.cu:35-50 at the tile counts of the benchmark configs
    assert [oracle_c.higher_msb(n) for n in (8160, 2500, 32400)] == [13, 12, 15]


def test_oracle_knn_matches_reference(gold):
    got = oracle_c.knn_dist2(gold["knn/points"])
    np.testing.assert_allclose(got, gold["knn/dist2"], rtol=2e-6, atol=0)
    assert got[10] < got.mean()           # the duplicated point has a zero nearest distance


def test_rigid_body_port_matches_reference_golden():
    g = torch.load(os.path.join(GOLD, "se3_golden.pt"))
    S = g["S"].clone().requires_grad_(True)
    th = g["theta"].clone().requires_grad_(True)
    x = g["x"].clone().requires_grad_(True)
    T = rigid_body_port.exp_se3(S, th)
    y = rigid_body_port.from_homogenous(torch.bmm(T, rigid_body_port.to_homogenous(x).unsqueeze(-1)).squeeze(-1))
    (y * g["gy"]).sum().backward()
    # same torch build as the generator -> bit-exact; allow an ulp across builds
    for name, got in (("T", T), ("y", y), ("dS", S.grad), ("dtheta", th.grad), ("dx", x.grad)):
        torch.testing.assert_close(got.detach(), g[name], rtol=1e-6, atol=1e-6, msg=name)
    torch.testing.assert_close(rigid_body_port.skew(g["S"][:, :3]), g["skew"], rtol=0, atol=0)
    torch.testing.assert_close(rigid_body_port.exp_so3(g["S"][:, :3], g["theta"]), g["exp_so3"], rtol=1e-6, atol=1e-6)
    # bottom row / rotation sanity of the SE3 matrices
    assert torch.equal(T[:, 3].detach(), torch.tensor([0.0, 0.0, 0.0, 1.0]).expand(T.shape[0], 4))
    R = T[:, :3, :3].detach()
    torch.testing.assert_close(torch.bmm(R, R.transpose(1, 2)), torch.eye(3).expand_as(R), rtol=0, atol=1e-5)


# ---------------------------------------------------------------------------
# Training loss + optimizer step: the torch port against outputs of the REAL
# utils/loss_utils.py and torch.optim.Adam (tests/golden/make_loss_golden.py)
# ---------------------------------------------------------------------------
@pytest.fixture(scope="module")
def loss_gold():
    return torch.load(os.path.join(GOLD, "loss_golden.pt"), weights_only=False)


def test_loss_port_matches_the_reference_loss_utils(loss_gold):
    from oracle import loss_port
    for c in loss_gold["cases"]:
        x = c["image"].clone().requires_grad_(True)
        loss = loss_port.training_loss(x, c["gt"], c["lambda_dssim"])
        loss.backward()
        assert torch.equal(loss.detach(), c["loss"])                     # same ops, same machine class: bit-exact
        assert torch.equal(loss_port.l1_loss(c["image"], c["gt"]), c["l1"])
        assert torch.equal(loss_port.ssim(c["image"], c["gt"]), c["ssim"])
        assert torch.allclose(x.grad, c["dloss_dimage"], rtol=0, atol=1e-9)


def test_adam_port_matches_torch_adam_as_the_reference_configures_it(loss_gold):
    from oracle import loss_port
    a = loss_gold["adam"]
    # the port takes fixed lrs; replay the golden's lr change at step 3 by splitting the run
    first = loss_port.adam_reference(a["params"], a["grads"][:3], a["lrs"])
    assert all(torch.isfinite(p).all() for p in first)
    ps = [p.clone().requires_grad_(True) for p in a["params"]]
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(ps, a["lrs"])], lr=0.0, eps=1e-15)
    for s, grads in enumerate(a["grads"]):
        for p, g in zip(ps, grads):
            p.grad = g.clone()
        if s == 3:
            opt.param_groups[0]["lr"] = a["lr0_from_step3"]
        opt.step()
    for p, f in zip(ps, a["final"]):
        assert torch.allclose(p.detach(), f, rtol=0, atol=1e-7)


# ---------------------------------------------------------------------------
# Deformation network (oracle only so far): the port against outputs of the REAL classes
# ---------------------------------------------------------------------------
def test_deformation_mlp_port_matches_the_reference_classes():
    from oracle import deform_mlp_port as port
    g = torch.load(os.path.join(GOLD, "mlp_golden.pt"), weights_only=False)

    def same_digest(t, d):
        f = t.detach().double().reshape(-1)
        return (tuple(t.shape) == d["shape"] and abs(float(f.sum()) - d["sum"]) <= 1e-9 * max(1.0, d["abs_sum"]) and
                torch.equal(t.detach().reshape(-1)[:16], d["head"]) and torch.equal(t.detach().reshape(-1)[-16:], d["tail"]))

    assert torch.equal(port.embed(g["x"]), g["embed_x"]) and g["embed_dim"] == 63
    torch.manual_seed(g["seed"])
    net = port.DeformMLP()
    assert sum(p.numel() for p in net.parameters()) == 513338
    for k, p in net.named_parameters():                      # same names, same initial values as the reference module
        assert same_digest(p, g["params"][k]), k
    x = g["x"].clone().requires_grad_(True)
    outs = net(x, g["ts"], g["iteration"])
    for o, want in zip(outs, g["outs"]):
        assert torch.equal(o.detach(), want)
    sum((o * p).sum() for o, p in zip(outs, g["proj"])).backward()
    assert torch.equal(x.grad, g["dx"])
    for k, p in net.named_parameters():
        assert same_digest(p.grad, g["dparams"][k]), k
    for o, want in zip(net(g["x"], g["ts"], 100), g["outs_early"]):     # iteration < 3000: zeros of the right shapes
        assert torch.equal(o, want) and float(o.abs().sum()) == 0.0
    S, th = port.screw_from_raw(g["se3"]["w_raw"], g["se3"]["v_raw"])
    assert torch.equal(S, g["se3"]["S"]) and torch.equal(th, g["se3"]["theta"])
    from oracle import rigid_body_port
    assert torch.allclose(rigid_body_port.exp_se3(S, th), g["se3"]["transform"], rtol=0, atol=1e-6)
