"""Run by tests/test_gpu_contract.py::test_kernel_variant_in_subprocess with GSR_* variables set: the kernel
variants are chosen once per process, so each one is checked in a process of its own - forward stages bit-exact
and gradients within the bar against the reference's rasterizer (oracle/_ref)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gaussian-splatting_deformable_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import synthetic  # noqa: E402
from _gpu_util import assert_grad_close, intermediates, make_view_settings, run_ours  # noqa: E402
from oracle import ref_driver as rd  # noqa: E402

for (P, W, H, smult) in ((200000, 800, 600, 1.0), (40000, 333, 177, 4.0)):
    sc, cam, rs = make_view_settings(P, W, H, scale_mult=smult)
    grad = synthetic.make_image_grad(W, H, device="cuda")
    kw = dict(shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    f = rd.forward(rs, sc["means3D"], sc["opacities"], **kw)
    b = rd.backward(rs, f, grad, sc["means3D"], **kw)
    R = f["num_rendered"]
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    m = intermediates(rs, sc)
    bn = rd.slice_binning(f["binning"], R)
    im = rd.slice_img(f["img"], W, H)
    assert m["R"] == R
    assert torch.equal(m["keys_sorted"], bn["point_list_keys"])
    assert torch.equal(m["point_list"], bn["point_list"])
    assert torch.equal(m["ranges"], im["ranges"][:tiles])
    assert torch.equal(m["n_contrib"], im["n_contrib"])
    assert float((m["color"] - f["color"]).abs().max()) <= 1e-5
    o = run_ours(rs, sc, grad)
    assert torch.equal(o["radii"], f["radii"])
    assert float((o["color"] - f["color"]).abs().max()) <= 1e-5
    for k in ("means3D", "opacities", "shs", "scales", "rotations"):
        assert_grad_close(o["grads"][k], b[k], "variant %s %s" % (sorted((k_, v) for k_, v in os.environ.items() if k_.startswith("GSR_")), k))
print("VARIANT_OK")
