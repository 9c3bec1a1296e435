"""the reference's PLY layout, pinned by the attribute list the REAL
GaussianModel.construct_list_of_attributes produces (tests/golden/make_ply_golden.py executes its source) and by an
independently packed byte image of a small cloud.  Pure host code: CPU only."""
import os
import struct

import numpy as np
import pytest
import torch

import ply_io

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cloud(P=7, K=15, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(s, generator=g)
    return dict(xyz=r(P, 3), features_dc=r(P, 1, 3), features_rest=r(P, K, 3), opacity=r(P, 1), scaling=r(P, 3), rotation=r(P, 4))


def test_attribute_order_is_the_reference_one():
    gold = torch.load(os.path.join(GOLD, "ply_golden.pt"), weights_only=False)
    assert ply_io.attribute_names() == gold["attributes_sh3"] and len(gold["attributes_sh3"]) == 62
    assert ply_io.attribute_names(num_rest=0) == gold["attributes_sh0"]


def test_file_bytes_and_round_trip(tmp_path):
    c = _cloud()
    path = str(tmp_path / "point_cloud" / "iteration_7" / "point_cloud.ply")       # directories are created (mkdir_p)
    ply_io.save_ply(path, **c)
    raw = open(path, "rb").read()
    names = ply_io.attribute_names()
    header = "ply\nformat binary_little_endian 1.0\nelement vertex 7\n" + "".join("property float %s\n" % n for n in names) + "end_header\n"
    assert raw.startswith(header.encode())
    body = raw[len(header):]
    assert len(body) == 7 * 62 * 4
    # row 3, packed by hand: xyz, zero normals, f_dc channel-major, f_rest channel-major, opacity, scales, quaternion
    i = 3
    row = list(c["xyz"][i]) + [0.0] * 3 + list(c["features_dc"][i].t().reshape(-1)) + list(c["features_rest"][i].t().reshape(-1)) + \
        list(c["opacity"][i]) + list(c["scaling"][i]) + list(c["rotation"][i])
    assert body[i * 62 * 4:(i + 1) * 62 * 4] == struct.pack("<62f", *[float(v) for v in row])
    back = ply_io.load_ply(path, max_sh_degree=3)
    for k, v in c.items():
        assert back[k].shape == v.shape and torch.equal(back[k], v), k


def test_reader_accepts_what_plyfile_variants_write_and_rejects_the_rest(tmp_path):
    c = _cloud(P=3, K=0)
    p = str(tmp_path / "dc_only.ply")
    ply_io.save_ply(p, **c)
    back = ply_io.load_ply(p, max_sh_degree=0)
    assert back["features_rest"].shape == (3, 0, 3) and torch.equal(back["features_dc"], c["features_dc"])
    # comments and big-endian bodies (plyfile can produce both) are read; a wrong SH degree is refused like the reference's assert
    raw = open(p, "rb").read()
    head, body = raw.split(b"end_header\n")
    be = head.replace(b"binary_little_endian", b"binary_big_endian").replace(b"ply\n", b"ply\ncomment made elsewhere\n") + b"end_header\n" + \
        np.frombuffer(body, "<f4").astype(">f4").tobytes()
    q = str(tmp_path / "be.ply")
    open(q, "wb").write(be)
    assert torch.equal(ply_io.load_ply(q, max_sh_degree=0)["xyz"], c["xyz"])
    try:
        ply_io.load_ply(p, max_sh_degree=3)
        assert False
    except ValueError:
        pass


def test_load_accepts_double_properties_and_reports_truncation(tmp_path):
    import numpy as np
    import ply_io
    P = 7
    g = torch.Generator().manual_seed(3)
    t = dict(xyz=torch.randn(P, 3, generator=g), features_dc=torch.randn(P, 1, 3, generator=g), features_rest=torch.randn(P, 15, 3, generator=g),
             opacity=torch.randn(P, 1, generator=g), scaling=torch.randn(P, 3, generator=g), rotation=torch.randn(P, 4, generator=g))
    path = str(tmp_path / "a" / "point_cloud.ply")
    ply_io.save_ply(path, t["xyz"], t["features_dc"], t["features_rest"], t["opacity"], t["scaling"], t["rotation"])
    raw = open(path, "rb").read()
    head, body = raw[:raw.index(b"end_header\n") + 11], raw[raw.index(b"end_header\n") + 11:]
    # the same table written by a tool that uses doubles (plyfile reads those; so does load_ply)
    dbl = str(tmp_path / "double.ply")
    with open(dbl, "wb") as f:
        f.write(head.replace(b"property float ", b"property double "))
        f.write(np.frombuffer(body, dtype="<f4").astype("<f8").tobytes())
    back = ply_io.load_ply(dbl)
    for k in t:
        assert torch.equal(back[k], t[k]), k
    cut = str(tmp_path / "cut.ply")
    open(cut, "wb").write(raw[:-10])
    with pytest.raises(ValueError, match="truncated"):
        ply_io.load_ply(cut)


def test_sidecar_networks_have_the_reference_parameter_names():
    """checkpoint_io.make_networks against the REAL classes of scene/gaussian_model.py (staged by oracle/build_ref.py)."""
    from oracle import ref_py
    if not os.path.exists(os.path.join(ref_py.PY, "scene", "gaussian_model.py")):
        pytest.skip("oracle/_ref/py not staged")
    import checkpoint_io
    gm = ref_py.gaussian_model()
    ref = gm.GaussianModel(3)
    mine = checkpoint_io.make_networks()
    for name in checkpoint_io.NETWORKS:
        a = {k: tuple(v.shape) for k, v in mine[name].state_dict().items()}
        b = {k: tuple(v.shape) for k, v in getattr(ref, name).state_dict().items()}
        assert a == b and list(a) == list(b), name
