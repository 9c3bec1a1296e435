"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every
symbol include/gsr_b200.h declares, workspace layouts are sane, and the Python
operator mirrors the reference's interface (names, field order, error behaviour).
No compute call is made (there is no GPU here and no CPU fallback)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

import gsr_runtime as rt
import diff_gaussian_rasterization as dgr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "gsr_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(gsr_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(rt.lib_path())
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libgsr_b200.so does not export %s" % n
    # and the Python binding knows every one of them
    assert set(names) == set(rt.EXPORTED_SYMBOLS)
    assert rt.load().gsr_version() >= 100


def test_struct_mirrors_match_header():
    txt = open(os.path.join(ROOT, "include", "gsr_b200.h")).read()
    body = re.search(r"typedef struct gsr_view \{(.*?)\} gsr_view;", txt, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [re.sub(r"\[.*", "", f.split()[-1]) for f in body.split(";") if f.strip()]
    assert fields == [f[0] for f in rt.gsr_view._fields_]
    assert ctypes.sizeof(rt.gsr_view) == 4 * (2 + 2 + 3 + 1 + 16 + 16 + 1 + 3 + 2)
    assert [f[0] for f in rt.gsr_deform._fields_] == ["mode", "num_bodies", "S", "theta", "body_id"]


def test_workspace_sizes_and_layouts():
    lib = rt.load()
    P, W, H, R = 1000000, 1920, 1080, 7600000
    gl = rt.geom_layout(P)
    offs = sorted(gl.values())
    assert offs[0] == 0 and all(o % 256 == 0 for o in offs) and len(set(offs)) == len(offs)
    assert gl["recs"] - gl["tiles_touched"] >= 4 * P and gl["clamped"] - gl["recs"] >= 48 * P
    assert lib.gsr_geom_bytes(P) > max(offs)
    assert lib.gsr_geom_bytes(2 * P) > lib.gsr_geom_bytes(P)
    il = rt.image_layout(W, H)
    assert il["n_contrib"] - il["final_T"] >= 4 * W * H and lib.gsr_image_bytes(W, H) >= il["ranges"] + 8 * 8160
    bl = rt.binning_layout(R, W, H)
    assert len(set(bl.values())) == 4
    # 2 tile-id arrays + 2 value arrays + onesweep state + reference-format keys: 24..32 B per duplicate
    # + the counting sort's chunk x tile count matrix (512 x 8160 x 4 B at 1080p), independent of R
    matrix = 512 * 8160 * 4
    assert 24 * R <= lib.gsr_binning_bytes(R, W, H) - matrix <= 32 * R
    assert lib.gsr_binning_bytes(0, W, H) < matrix + (1 << 20)
    assert lib.gsr_grad_bytes(P) >= 48 * P
    # 45 key bits at 1080p -> 6 digit passes (even): the sorted data ends in the first buffer pair
    assert lib.gsr_sort_bytes(R, 0, 45) > 6 * ((R + 4095) // 4096) * 256 * 4
    assert lib.gsr_knn_bytes(100000) > 100000 * (2 * 8 + 2 * 4 + 16)


def test_settings_tuple_is_the_reference_one():
    # diff_gaussian_rasterization/__init__.py:157-169
    assert dgr.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")
    sig = inspect.signature(dgr.GaussianRasterizer.forward)
    assert list(sig.parameters)[:9] == ["self", "means3D", "means2D", "opacities", "shs", "colors_precomp", "scales",
                                        "rotations", "cov3D_precomp"]
    for extra in ("se3_S", "se3_theta", "body_id"):         # additions are keyword-only, default None
        p = sig.parameters[extra]
        assert p.kind is inspect.Parameter.KEYWORD_ONLY and p.default is None
    assert list(inspect.signature(dgr.rasterize_gaussians).parameters)[:9] == [
        "means3D", "means2D", "sh", "colors_precomp", "opacities", "scales", "rotations", "cov3Ds_precomp",
        "raster_settings"]
    assert hasattr(dgr.GaussianRasterizer, "markVisible") and hasattr(dgr, "_RasterizeGaussians")


def _settings():
    z = torch.zeros(3)
    return dgr.GaussianRasterizationSettings(32, 32, 0.5, 0.5, z, 1.0, torch.eye(4), torch.eye(4), 3, z, False, False)


def test_argument_validation_matches_reference_messages():
    ras = dgr.GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    o = torch.zeros(4, 1)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        ras(m, m, o, scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        ras(m, m, o, shs=torch.zeros(4, 16, 3), colors_precomp=torch.zeros(4, 3), scales=torch.ones(4, 3),
            rotations=torch.ones(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        ras(m, m, o, shs=torch.zeros(4, 16, 3))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        ras(m, m, o, shs=torch.zeros(4, 16, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4),
            cov3D_precomp=torch.zeros(4, 6))
    with pytest.raises(Exception, match="both se3_S and se3_theta"):
        ras(m, m, o, shs=torch.zeros(4, 16, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4),
            se3_S=torch.zeros(4, 6))
    with pytest.raises(RuntimeError, match=r"means3D must have dimensions \(num_points, 3\)"):
        ras(torch.zeros(4, 4), m, o, shs=torch.zeros(4, 16, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))


def test_no_cpu_fallback():
    """CPU tensors must fail loudly: the product has no eager/CPU path."""
    import rigid_body
    from simple_knn._C import distCUDA2
    ras = dgr.GaussianRasterizer(_settings())
    m = torch.zeros(4, 3)
    with pytest.raises(rt.GsrError, match="no CPU fallback"):
        ras(m, m, torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), scales=torch.ones(4, 3), rotations=torch.ones(4, 4))
    with pytest.raises(rt.GsrError, match="no CPU fallback"):
        distCUDA2(torch.zeros(8, 3))
    with pytest.raises(rt.GsrError, match="no CPU fallback"):
        rigid_body.exp_se3(torch.zeros(2, 6), torch.ones(2))
    with pytest.raises(rt.GsrError, match="no CPU fallback"):       # the view-batched preprocess has no CPU path either
        dgr.GaussianForwardBatch([_settings()], means3D=m, opacities=torch.zeros(4, 1), shs=torch.zeros(4, 16, 3),
                                 scales=torch.ones(4, 3), rotations=torch.ones(4, 4))


def test_view_batch_objects_host_logic():
    """GaussianBackwardBatch: dict-like target set, flush of an empty batch is a no-op, chunk count sanitised."""
    t = {"means3D": torch.zeros(4, 3)}
    b = dgr.GaussianBackwardBatch(t, chunks=0)
    assert len(b) == 0 and b.chunks == 1 and dict(b.items()) == t and b.viewspace_grads == []
    b.flush()                                                        # nothing pending: must not touch the library
    assert len(b) == 0
    # the batched entry points are part of the C ABI the header declares
    for name in ("gsr_backward_blend", "gsr_backward_batched_fill_slots", "gsr_backward_gaussians_batched",
                 "gsr_forward_batched_fill_slots", "gsr_forward_preprocess_batched", "gsr_read_num_rendered", "gsr_depth_order"):
        assert name in rt.EXPORTED_SYMBOLS


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "gaussian-splatting_deformable_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), "%s mentions the oracle" % f


def test_host_view_cache_tracks_identity_and_version():
    t = torch.arange(3, dtype=torch.float32)
    assert rt.host_values(t, 3) == [0.0, 1.0, 2.0]
    t.add_(1)                                   # in-place edit bumps _version -> must be re-read
    assert rt.host_values(t, 3) == [1.0, 2.0, 3.0]
    u = torch.tensor([9.0, 9.0, 9.0])
    assert rt.host_values(u, 3) == [9.0, 9.0, 9.0]
    v = rt.make_view(_settings())
    assert v.image_width == 32 and list(v.viewmatrix)[:5] == [1.0, 0.0, 0.0, 0.0, 0.0]


def test_rigid_body_helpers_match_port():
    import rigid_body
    from oracle import rigid_body_port as port
    g = torch.Generator().manual_seed(0)
    w = torch.randn(5, 3, generator=g)
    th = torch.rand(5, generator=g) + 0.1
    p = torch.randn(5, 3, generator=g)
    assert torch.equal(rigid_body.skew(w), port.skew(w))
    torch.testing.assert_close(rigid_body.exp_so3(w, th), port.exp_so3(w, th))
    R = port.exp_so3(w, th)
    assert torch.equal(rigid_body.rp_to_se3(R, p), port.rp_to_se3(R, p))
    assert torch.equal(rigid_body.to_homogenous(p), port.to_homogenous(p))
    h = port.to_homogenous(p) * 2.0
    assert torch.equal(rigid_body.from_homogenous(h), port.from_homogenous(h))


def test_loss_and_optimizer_modules_mirror_the_reference_surface_and_refuse_cpu():
    """same names/signatures as utils/loss_utils.py; no CPU fallback behind them."""
    import fused_adam
    import loss_utils
    for name in ("l1_loss", "l2_loss", "gaussian", "create_window", "ssim"):          # utils/loss_utils.py:17-42
        assert callable(getattr(loss_utils, name))
    assert list(inspect.signature(loss_utils.ssim).parameters) == ["img1", "img2", "window_size", "size_average"]
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "loss_golden.pt"), weights_only=False)
    c = gold["cases"][0]
    # the window the kernels receive is the reference's own (loss_utils.py:23-25), value for value
    w = loss_utils.gaussian(11, 1.5)
    assert w.shape == (11,) and abs(float(w.sum()) - 1.0) < 1e-6 and torch.equal(w, w.flip(0))
    assert torch.equal(loss_utils.create_window(11, 3)[1, 0], torch.outer(w, w))
    # l1/l2 are the reference's torch one-liners and work anywhere; ssim / the fused loss / Adam need the GPU library
    assert torch.equal(loss_utils.l1_loss(c["image"], c["gt"]), c["l1"])
    with pytest.raises(rt.GsrError):
        loss_utils.ssim(c["image"], c["gt"])
    with pytest.raises(rt.GsrError):
        loss_utils.l1_ssim_loss(c["image"], c["gt"], 0.2)
    with pytest.raises(NotImplementedError):
        loss_utils.ssim(c["image"], c["gt"], window_size=7)
    with pytest.raises(rt.GsrError):
        fused_adam.FusedAdam([{"params": [torch.zeros(4, requires_grad=True)], "lr": 0.1}], lr=0.0, eps=1e-15)
    with pytest.raises(ValueError):
        fused_adam.FusedAdam([])


def test_flat_grad_buffer_rebind_and_sequential_render_views():
    import view_parallel
    a = [torch.zeros(5, 3, requires_grad=True), torch.zeros(7, requires_grad=True)]
    buf = view_parallel.FlatGradBuffer(a)
    # every tensor starts on a 32-byte boundary (8 floats): 15 -> 16, 7 -> 8
    assert buf.offsets == [0, 16] and buf.flat.numel() == 24 and a[1].grad.data_ptr() == buf.flat[16:].data_ptr()
    a[0].grad += 2.0
    b = [torch.zeros(5, 3, requires_grad=True), torch.zeros(7, requires_grad=True)]
    buf.rebind(b)                                               # same memory, new owners, nothing cleared
    assert float(b[0].grad.sum()) == 30.0 and b[0].grad.data_ptr() == buf.flat.data_ptr()
    with pytest.raises(ValueError):
        buf.rebind([torch.zeros(4, 3), torch.zeros(7)])
    # without CUDA render_views degrades to the plain sequential loop (host logic only; no rasterizer involved)
    seen = []
    total = view_parallel.render_views(lambda i: (seen.append(i), torch.tensor(float(i)))[1], [3, 1, 2], num_streams=4)
    assert seen == [3, 1, 2] and float(total) == 6.0


def test_lr_schedule_matches_the_reference_function():
    import fused_adam
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "loss_golden.pt"), weights_only=False)["lr_schedule"]
    for name, cfg in gold["cfgs"].items():
        f = fused_adam.get_expon_lr_func(**cfg)
        for s, want in zip(gold["steps"], gold["values"][name]):
            assert abs(f(s) - want) <= 1e-12 * max(1.0, abs(want)) + 1e-18, (name, s)
    assert list(inspect.signature(fused_adam.get_expon_lr_func).parameters) == [
        "lr_init", "lr_final", "lr_delay_steps", "lr_delay_mult", "max_steps"]


def test_batched_slot_packing_is_host_only_and_validates():
    """gsr_forward/backward_batched_fill_slots are pure host functions (no GPU needed): sizes, argument checks and the
    frozen conventions they enforce (one scale_modifier / sh_degree per batch, no prefiltered / debug views)."""
    import ctypes
    lib = rt.load()
    P, n = 1000, 3
    views = [rt.make_view(_settings()) for _ in range(n)]
    nb_f, nb_b = lib.gsr_forward_batched_slots_bytes(n), lib.gsr_backward_batched_slots_bytes(n)
    assert nb_f > 0 and nb_b > 0 and nb_f % n == 0 and nb_b % n == 0
    assert lib.gsr_forward_batched_slots_bytes(0) == 0 and lib.gsr_forward_batched_slots_bytes(2 * n) == 2 * nb_f
    host = (ctypes.c_uint8 * max(nb_f, nb_b))()
    fake = 0x10000                                       # never dereferenced: the functions only do pointer arithmetic
    fa = (rt.gsr_view_fwd * n)()
    ba = (rt.gsr_view_grads * n)()
    for j in range(n):
        fa[j].view = ctypes.pointer(views[j]); fa[j].radii = fake; fa[j].geom_ws = fake
        ba[j].view = ctypes.pointer(views[j]); ba[j].radii = fake; ba[j].geom_ws = fake; ba[j].grad_ws = fake; ba[j].dL_dmeans2D = fake
    assert lib.gsr_forward_batched_fill_slots(n, fa, P, 16, host, nb_f) == 0
    assert lib.gsr_backward_batched_fill_slots(n, ba, P, 16, host, nb_b) == 0
    assert any(host[i] for i in range(nb_b))             # something was written
    assert lib.gsr_forward_batched_fill_slots(n, fa, P, 16, host, nb_f - 1) != 0        # buffer too small
    assert b"too small" in lib.gsr_last_error_string()
    assert lib.gsr_backward_batched_fill_slots(n, ba, P, 16, host, nb_b - 1) != 0
    fa[1].geom_ws = None
    assert lib.gsr_forward_batched_fill_slots(n, fa, P, 16, host, nb_f) != 0            # NULL workspace
    fa[1].geom_ws = fake
    views[2].scale_modifier = 2.0
    assert lib.gsr_forward_batched_fill_slots(n, fa, P, 16, host, nb_f) != 0            # mixed scale_modifier
    assert b"scale_modifier" in lib.gsr_last_error_string()
    assert lib.gsr_backward_batched_fill_slots(n, ba, P, 16, host, nb_b) != 0
    views[2].scale_modifier = views[0].scale_modifier
    views[0].prefiltered = 1
    assert lib.gsr_forward_batched_fill_slots(n, fa, P, 16, host, nb_f) != 0            # prefiltered views take the per-view path
    views[0].prefiltered = 0
    assert lib.gsr_forward_batched_fill_slots(n, fa, P, 16, host, nb_f) == 0
    # depth-sort workspace query grows with P and is part of the geometry workspace
    assert lib.gsr_depth_order_ws_bytes(1000) < lib.gsr_depth_order_ws_bytes(1000000) < lib.gsr_geom_bytes(1000000)
