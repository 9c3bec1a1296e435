"""Driver for oracle/_ref/ref_dgr_C.so and ref_knn_C.so: the reference's own CUDA
rasterizer and kNN, compiled unmodified by oracle/build_ref.py.

TEST INFRASTRUCTURE ONLY (checker / reference arm of the benchmark).

It restates, in a few lines, what the reference's Python wrapper does around `_C`
(diff_gaussian_rasterization/__init__.py:46-155: argument order of
rasterize_gaussians / rasterize_gaussians_backward and the order of the returned
gradients) and slices the three opaque byte buffers with the `obtain()` layout of
rasterizer_impl.h:21-27 / rasterizer_impl.cu:155-194 (every sub-array start rounded
up to 128 B) so intermediate stages can be compared bit for bit.
"""
import importlib.util
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_mods = {}


def available():
    return os.path.exists(os.path.join(_REF, "ref_dgr_C.so"))


def _load(name):
    if name not in _mods:
        path = os.path.join(_REF, name + ".so")
        if not os.path.exists(path):
            raise FileNotFoundError("%s not built (run python oracle/build_ref.py where /root/reference exists)" % path)
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mods[name] = m
    return _mods[name]


def _e():
    return torch.Tensor([])


def forward(rs, means3D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
    """_C.rasterize_gaussians with the reference's 19-argument order.  Returns a dict."""
    C = _load("ref_dgr_C")
    args = (rs.bg, means3D, _e() if colors_precomp is None else colors_precomp, opacities,
            _e() if scales is None else scales, _e() if rotations is None else rotations,
            float(rs.scale_modifier), _e() if cov3D_precomp is None else cov3D_precomp,
            rs.viewmatrix, rs.projmatrix, float(rs.tanfovx), float(rs.tanfovy),
            int(rs.image_height), int(rs.image_width), _e() if shs is None else shs, int(rs.sh_degree),
            rs.campos, bool(rs.prefiltered), bool(rs.debug))
    R, color, radii, geom, binning, img = C.rasterize_gaussians(*args)
    return dict(num_rendered=int(R), color=color, radii=radii, geom=geom, binning=binning, img=img)


def backward(rs, fwd, grad_color, means3D, shs=None, colors_precomp=None, scales=None, rotations=None,
             cov3D_precomp=None):
    """_C.rasterize_gaussians_backward; returns grads keyed by name."""
    C = _load("ref_dgr_C")
    args = (rs.bg, means3D, fwd["radii"], _e() if colors_precomp is None else colors_precomp,
            _e() if scales is None else scales, _e() if rotations is None else rotations,
            float(rs.scale_modifier), _e() if cov3D_precomp is None else cov3D_precomp,
            rs.viewmatrix, rs.projmatrix, float(rs.tanfovx), float(rs.tanfovy), grad_color,
            _e() if shs is None else shs, int(rs.sh_degree), rs.campos, fwd["geom"], fwd["num_rendered"],
            fwd["binning"], fwd["img"], bool(rs.debug))
    g2d, gcol, gop, g3d, gcov, gsh, gsc, grot = C.rasterize_gaussians_backward(*args)
    return dict(means2D=g2d, colors=gcol, opacities=gop, means3D=g3d, cov3D=gcov, shs=gsh, scales=gsc,
                rotations=grot)


def mark_visible(means3D, viewmatrix, projmatrix):
    return _load("ref_dgr_C").mark_visible(means3D, viewmatrix, projmatrix)


def dist_cuda2(points):
    return _load("ref_knn_C").distCUDA2(points)


def _carve(buf, spec):
    """spec: list of (name, count, torch dtype, elems_per_item). 128-byte aligned starts."""
    out, off = {}, 0
    for name, count, dtype, width in spec:
        off = (off + 127) & ~127
        nbytes = count * width * torch.empty((), dtype=dtype).element_size()
        t = buf[off:off + nbytes].view(dtype)
        out[name] = t.view(count, width) if width > 1 else t
        off += nbytes
    return out


def slice_geom(geom, P):
    return _carve(geom, [("depths", P, torch.float32, 1), ("clamped", P, torch.bool, 3),
                         ("internal_radii", P, torch.int32, 1), ("means2D", P, torch.float32, 2),
                         ("cov3D", P, torch.float32, 6), ("conic_opacity", P, torch.float32, 4),
                         ("rgb", P, torch.float32, 3), ("tiles_touched", P, torch.int32, 1)])


def slice_binning(binning, R):
    return _carve(binning, [("point_list", R, torch.int32, 1), ("point_list_unsorted", R, torch.int32, 1),
                            ("point_list_keys", R, torch.int64, 1), ("point_list_keys_unsorted", R, torch.int64, 1)])


def slice_img(img, W, H):
    N = W * H
    return _carve(img, [("accum_alpha", N, torch.float32, 1), ("n_contrib", N, torch.int32, 1),
                        ("ranges", N, torch.int32, 2)])
