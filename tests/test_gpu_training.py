"""densification decisions and checkpoint formats against the reference's REAL GaussianModel
(oracle/_ref/py/scene/gaussian_model.py, staged unmodified) on the GPU."""
import os
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu

OPT = SimpleNamespace(percent_dense=0.01, position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01,
                      position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001)


def _gm():
    from oracle import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/py not staged (needs /root/reference at build time)")
    return ref_py.gaussian_model()


def _reference_model(gm, P, seed):
    from test_gpu_contract import _real_gaussian_model
    pc = _real_gaussian_model(gm, P, seed)
    pc.spatial_lr_scale = 1.0
    pc.training_setup(OPT)
    return pc


def _ours_from(pc):
    """FusedAdam + DensificationStats mirroring a reference GaussianModel's state (same group order and names)."""
    import copy
    import densify
    import fused_adam
    t = {"xyz": pc._xyz, "f_dc": pc._features_dc, "f_rest": pc._features_rest, "opacity": pc._opacity, "scaling": pc._scaling,
         "rotation": pc._rotation}
    mine = {k: v.detach().clone().requires_grad_(True) for k, v in t.items()}
    net = copy.deepcopy(pc.offset_model)
    groups = []
    for g in pc.optimizer.param_groups:
        if g["name"] == "offset_model":
            groups.append({"params": list(net.parameters()), "lr": g["lr"], "name": "offset_model"})
        else:
            groups.append({"params": [mine[g["name"]]], "lr": g["lr"], "name": g["name"]})
    opt = fused_adam.FusedAdam(groups, lr=0.0, eps=1e-15)
    stats = densify.DensificationStats(pc._xyz.shape[0], "cuda")
    return mine, net, opt, stats


def _step_both(pc, mine, opt, g, scale=0.01):
    opt.zero_grad()
    ref_t = {"xyz": pc._xyz, "f_dc": pc._features_dc, "f_rest": pc._features_rest, "opacity": pc._opacity, "scaling": pc._scaling,
             "rotation": pc._rotation}
    for k, v in ref_t.items():
        gr = torch.randn(v.shape, generator=g).cuda() * scale
        v.grad = gr.clone()
        mine[k].grad.copy_(gr)
    opt.step(skip=("offset_model",))                       # the network gets no gradient in this test: torch skips it too
    pc.optimizer.step()
    pc.optimizer.zero_grad(set_to_none=True)


def _same_state(pc, mine, opt, stats, tag):
    ref_t = {"xyz": pc._xyz, "f_dc": pc._features_dc, "f_rest": pc._features_rest, "opacity": pc._opacity, "scaling": pc._scaling,
             "rotation": pc._rotation}
    for k, v in ref_t.items():
        assert mine[k].shape == v.shape, (tag, k, mine[k].shape, v.shape)
        assert torch.allclose(mine[k].detach(), v.detach(), rtol=2e-6, atol=1e-7), (tag, k, float((mine[k] - v).abs().max()))
    sd = opt.state_dict()
    for gi, g in enumerate(pc.optimizer.param_groups):
        if g["name"] == "offset_model":
            continue
        st = pc.optimizer.state.get(g["params"][0])
        idx = sd["param_groups"][gi]["params"][0]
        assert torch.allclose(sd["state"][idx]["exp_avg"], st["exp_avg"], rtol=1e-5, atol=1e-9), (tag, g["name"])
        assert torch.allclose(sd["state"][idx]["exp_avg_sq"], st["exp_avg_sq"], rtol=1e-5, atol=1e-12), (tag, g["name"])
    assert torch.equal(stats.max_radii2D, pc.max_radii2D) and torch.equal(stats.denom, pc.denom)
    assert torch.allclose(stats.xyz_gradient_accum, pc.xyz_gradient_accum, rtol=1e-6, atol=0)
    assert torch.allclose(stats.xyz_gradient_accum_3vec, pc.xyz_gradient_accum_3vec, rtol=1e-6, atol=0)


@pytest.mark.parametrize("P,size_threshold", [(100000, None), (100000, 20), (3001, 20)])
def test_densify_and_prune_vs_reference_gaussian_model(P, size_threshold):
    import densify
    gm = _gm()
    pc = _reference_model(gm, P, seed=21)
    mine, net, opt, stats = _ours_from(pc)
    g = torch.Generator().manual_seed(P)
    extent = 2.6
    for it in range(3):
        _step_both(pc, mine, opt, g)
    # statistics of a few "views": random view-space gradients, ~60 % of the points visible per view
    for view in range(4):
        vs = torch.randn((P, 3), generator=g).cuda() * 3e-4
        radii = (torch.rand(P, generator=g) * 40 - 15).clamp_min(0).to(torch.int32).cuda()
        vis = radii > 0
        pc.max_radii2D[vis] = torch.max(pc.max_radii2D[vis], radii[vis])                    # train.py:613
        pc.add_densification_stats(SimpleNamespace(grad=vs), vis)                         # train.py:618
        stats.add(vs, radii)
    _same_state(pc, mine, opt, stats, "stats")
    thr = 0.0002
    clone, split = densify.decide(stats, mine["scaling"], thr, OPT.percent_dense, extent)
    grads = pc.xyz_gradient_accum / pc.denom
    grads[grads.isnan()] = 0.0
    big = torch.max(pc.get_scaling, dim=1).values > OPT.percent_dense * extent
    assert torch.equal(clone, (torch.norm(grads, dim=-1) >= thr) & ~big) and torch.equal(split, (grads.squeeze() >= thr) & big)
    assert int(clone.sum()) + int(split.sum()) > 100 and (P < 50000 or (int(clone.sum()) > 100 and int(split.sum()) > 100))
    # the same random numbers for the split samples
    torch.manual_seed(77)
    pc.densify_and_prune(thr, 0.005, extent, size_threshold)
    gen = torch.Generator(device="cuda").manual_seed(77)
    named = densify.densify_and_prune(opt, stats, thr, 0.005, extent, size_threshold, percent_dense=OPT.percent_dense, generator=gen)
    assert named["xyz"] is mine["xyz"] and mine["xyz"].shape[0] != P
    _same_state(pc, mine, opt, stats, "densified")
    for it in range(2):                                                                    # and the optimizers keep stepping in lockstep
        _step_both(pc, mine, opt, g)
    _same_state(pc, mine, opt, stats, "stepped")
    pc.reset_opacity()
    densify.reset_opacity(opt)
    _same_state(pc, mine, opt, stats, "opacity reset")
    _step_both(pc, mine, opt, g)
    _same_state(pc, mine, opt, stats, "after reset")
    # the densified set renders through flat buffers of an arbitrary row count (32-byte aligned tensor starts)
    assert all(mine[k].data_ptr() % 32 == 0 for k in mine)


def test_checkpoint_tuple_and_sidecars_round_trip_with_the_reference(tmp_path):
    """f4: chkpnt_<it>.pth = (GaussianModel.capture(), iteration) + the five network state-dicts (train.py:685-697),
    both directions: the reference's capture restores into FusedAdam, ours restores into the reference's GaussianModel."""
    import checkpoint_io
    import densify
    import fused_adam
    gm = _gm()
    P = 5003
    pc = _reference_model(gm, P, seed=31)
    mine, net, opt, stats = _ours_from(pc)
    g = torch.Generator().manual_seed(1)
    for it in range(3):
        _step_both(pc, mine, opt, g)
    # ---- reference -> ours ----
    ref_nets = {"offset_model": pc.offset_model, "offset_model_rot": pc.offset_model_rot, "offset_model_scaling": pc.offset_model_scaling,
                "opacity_mask": pc.opacity_mask, "shs_model": pc.shs_model}
    d = tmp_path / "ckpt_save"
    d.mkdir()
    torch.save((pc.capture(), 3), str(d / "chkpnt_3.pth"))
    for k, m in ref_nets.items():
        torch.save(m.state_dict(), str(d / ("%s_3.pth" % k)))
    model_args, first_iter, nets = checkpoint_io.load_checkpoint(str(d / "chkpnt_3.pth"))
    assert first_iter == 3
    for k in checkpoint_io.NETWORKS:                          # every parameter name and shape of the reference's networks
        a, b = nets[k].state_dict(), ref_nets[k].state_dict()
        assert list(a.keys()) == list(b.keys()) and all(torch.equal(a[n].cpu(), b[n].cpu()) for n in a), k

    def make_opt(params):
        groups = []
        for grp in pc.optimizer.param_groups:
            ps = list(nets["offset_model"].parameters()) if grp["name"] == "offset_model" else [params[grp["name"]]]
            groups.append({"params": ps, "lr": grp["lr"], "name": grp["name"]})
        return fused_adam.FusedAdam(groups, lr=0.0, eps=1e-15)
    sh, params, st2, opt2, lr_scale = checkpoint_io.restore(model_args, make_opt)
    assert sh == pc.active_sh_degree and lr_scale == pc.spatial_lr_scale
    _same_state(pc, params, opt2, st2, "restored from the reference's checkpoint")
    _step_both(pc, params, opt2, g)
    _same_state(pc, params, opt2, st2, "stepped after restore")
    # ---- ours -> reference ----
    path = checkpoint_io.save_checkpoint(str(tmp_path / "out"), 4, checkpoint_io.capture(sh, params, st2, opt2, lr_scale), nets)
    (model_params, it4) = torch.load(path, map_location="cuda:0", weights_only=False)
    pc2 = gm.GaussianModel(3)
    pc2.restore(model_params, OPT)                              # the reference's own restore (training_setup + load_state_dict)
    for k in checkpoint_io.NETWORKS:
        getattr(pc2, k).load_state_dict(torch.load(os.path.join(os.path.dirname(path), "%s_4.pth" % k)))
    assert it4 == 4
    _same_state(pc2, params, opt2, st2, "the reference restored our checkpoint")
    _step_both(pc2, params, opt2, g)
    _same_state(pc2, params, opt2, st2, "stepped after the reference's restore")
    # ---- save_ply's sidecars ----
    checkpoint_io.save_point_cloud(str(tmp_path / "pc" / "point_cloud.ply"), params["xyz"], params["f_dc"], params["f_rest"],
                                   params["opacity"], params["scaling"], params["rotation"], nets)
    t, nets2 = checkpoint_io.load_point_cloud(str(tmp_path / "pc" / "point_cloud.ply"))
    assert torch.equal(t["xyz"], params["xyz"].detach()) and torch.equal(t["features_rest"], params["f_rest"].detach())
    for k in checkpoint_io.NETWORKS:                          # and the reference's own classes accept every file strictly
        getattr(pc2, k).load_state_dict(torch.load(str(tmp_path / "pc" / (k + ".pth"))), strict=True)
