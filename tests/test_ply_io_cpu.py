"""the reference's PLY layout, pinned by the attribute list the REAL
GaussianModel.construct_list_of_attributes produces (tests/golden/make_ply_golden.py executes its source) and by an
independently packed byte image of a small cloud.  Pure host code: CPU only."""
import os
import struct

import numpy as np
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
