"""ctypes/numpy wrapper of oracle/liboracle.so (gsr_oracle.c, the CPU restatement of the
reference rasterizer).  TEST INFRASTRUCTURE ONLY - never imported by the product."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


class orc_view(ctypes.Structure):
    _fields_ = [("P", ctypes.c_int), ("D", ctypes.c_int), ("M", ctypes.c_int), ("W", ctypes.c_int),
                ("H", ctypes.c_int), ("tan_fovx", ctypes.c_float), ("tan_fovy", ctypes.c_float),
                ("scale_modifier", ctypes.c_float), ("bg", ctypes.c_float * 3), ("view", ctypes.c_float * 16),
                ("proj", ctypes.c_float * 16), ("campos", ctypes.c_float * 3)]


def available():
    if os.path.exists(_LIB):
        return True
    try:
        from . import build_oracle
        build_oracle.build()
        return True
    except Exception:
        return False


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            from . import build_oracle
            build_oracle.build()
        _lib = ctypes.CDLL(_LIB)
        _lib.orc_count.restype = ctypes.c_uint32
        _lib.orc_bin.restype = ctypes.c_uint32
        _lib.orc_higher_msb.restype = ctypes.c_uint32
    return _lib


def _np(t, dtype=np.float32):
    if t is None:
        return None
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def make_view(P, D, M, W, H, tanfovx, tanfovy, bg, viewmatrix, projmatrix, campos, scale_modifier=1.0):
    v = orc_view()
    v.P, v.D, v.M, v.W, v.H = int(P), int(D), int(M), int(W), int(H)
    v.tan_fovx, v.tan_fovy, v.scale_modifier = float(tanfovx), float(tanfovy), float(scale_modifier)
    v.bg[:] = [float(x) for x in _np(bg).reshape(-1)]
    v.view[:] = [float(x) for x in _np(viewmatrix).reshape(-1)]
    v.proj[:] = [float(x) for x in _np(projmatrix).reshape(-1)]
    v.campos[:] = [float(x) for x in _np(campos).reshape(-1)]
    return v


def view_from_settings(rs, P, M):
    return make_view(P, rs.sh_degree, M, rs.image_width, rs.image_height, rs.tanfovx, rs.tanfovy, rs.bg,
                     rs.viewmatrix, rs.projmatrix, rs.campos, rs.scale_modifier)


def forward(rs, means3D, opacities, shs=None, scales=None, rotations=None, colors_precomp=None, cov3D_precomp=None,
            **view_override):
    """Full forward on the CPU.  Returns a dict of numpy arrays with every intermediate."""
    L = lib()
    means = _np(means3D)
    P = means.shape[0]
    shs_n, colors_n = _np(shs), _np(colors_precomp)
    M = shs_n.shape[1] if shs_n is not None else 0
    if view_override:
        v = make_view(P, rs.sh_degree, M, rs.image_width, rs.image_height, rs.tanfovx, rs.tanfovy,
                      view_override.get("bg", rs.bg), view_override.get("viewmatrix", rs.viewmatrix),
                      view_override.get("projmatrix", rs.projmatrix), view_override.get("campos", rs.campos),
                      rs.scale_modifier)
    else:
        v = view_from_settings(rs, P, M)
    W, H = v.W, v.H
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    scales_n, rots_n, cov_pre = _np(scales), _np(rotations), _np(cov3D_precomp)
    opac = _np(opacities).reshape(-1)
    o = dict(radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
             cov3D=np.zeros((P, 6), np.float32), rgb=np.zeros((P, 3), np.float32),
             conic_opacity=np.zeros((P, 4), np.float32), tiles_touched=np.zeros(P, np.uint32),
             clamped=np.zeros((P, 3), np.uint8))
    L.orc_preprocess(ctypes.byref(v), _p(means), _p(scales_n), _p(rots_n), _p(opac), _p(shs_n), _p(cov_pre),
                     _p(colors_n), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["rgb"]),
                     _p(o["conic_opacity"]), _p(o["tiles_touched"]), _p(o["clamped"]))
    o["point_offsets"] = np.zeros(P, np.uint32)
    R = int(L.orc_count(ctypes.byref(v), _p(o["tiles_touched"]), _p(o["point_offsets"])))
    o["num_rendered"] = R
    o["keys_unsorted"] = np.zeros(max(R, 1), np.uint64)
    o["keys_sorted"] = np.zeros(max(R, 1), np.uint64)
    o["point_list"] = np.zeros(max(R, 1), np.uint32)
    o["ranges"] = np.zeros((tiles, 2), np.uint32)
    L.orc_bin(ctypes.byref(v), _p(o["radii"]), _p(o["means2D"]), _p(o["depths"]), _p(o["tiles_touched"]),
              _p(o["keys_unsorted"]), _p(o["keys_sorted"]), _p(o["point_list"]), _p(o["ranges"]))
    for k in ("keys_unsorted", "keys_sorted", "point_list"):
        o[k] = o[k][:R]
    feats = colors_n if colors_n is not None else o["rgb"]
    o["final_T"] = np.zeros(W * H, np.float32)
    o["n_contrib"] = np.zeros(W * H, np.uint32)
    o["color"] = np.zeros((3, H, W), np.float32)
    L.orc_render(ctypes.byref(v), _p(o["ranges"]), _p(o["point_list"]), _p(o["means2D"]), _p(feats),
                 _p(o["conic_opacity"]), _p(o["final_T"]), _p(o["n_contrib"]), _p(o["color"]))
    o["_view"] = v
    o["_inputs"] = dict(means=means, scales=scales_n, rots=rots_n, shs=shs_n, colors=colors_n, cov_pre=cov_pre,
                        feats=feats)
    return o


def backward(fwd, dL_dpix):
    """Backward for a forward() result.  Returns grads keyed like the reference's outputs."""
    L = lib()
    v = fwd["_view"]
    inp = fwd["_inputs"]
    P, M = v.P, v.M
    g = _np(dL_dpix)
    d2d = np.zeros((P, 3), np.float32)
    dcon = np.zeros((P, 4), np.float32)
    dop = np.zeros((P, 1), np.float32)
    dcol = np.zeros((P, 3), np.float32)
    L.orc_render_backward(ctypes.byref(v), _p(fwd["ranges"]), _p(fwd["point_list"]), _p(fwd["means2D"]),
                          _p(fwd["conic_opacity"]), _p(inp["feats"]), _p(fwd["final_T"]), _p(fwd["n_contrib"]),
                          _p(g), _p(d2d), _p(dcon), _p(dop), _p(dcol))
    d3d = np.zeros((P, 3), np.float32)
    dcov = np.zeros((P, 6), np.float32)
    dsh = np.zeros((P, max(M, 1), 3), np.float32)
    dsc = np.zeros((P, 3), np.float32)
    drot = np.zeros((P, 4), np.float32)
    cov3D = inp["cov_pre"] if inp["cov_pre"] is not None else fwd["cov3D"]
    L.orc_preprocess_backward(ctypes.byref(v), _p(inp["means"]), _p(fwd["radii"]), _p(inp["shs"]),
                              _p(fwd["clamped"]), _p(inp["scales"]), _p(inp["rots"]), _p(cov3D), _p(d2d), _p(dcon),
                              _p(dcol), _p(d3d), _p(dcov), _p(dsh), _p(dsc), _p(drot))
    return dict(means2D=d2d, conic=dcon, opacities=dop, colors=dcol, means3D=d3d, cov3D=dcov,
                shs=dsh[:, :M] if M else dsh[:, :0], scales=dsc, rotations=drot)


def mark_visible(means3D, viewmatrix, projmatrix):
    means = _np(means3D)
    v = make_view(means.shape[0], 0, 0, 16, 16, 1.0, 1.0, [0, 0, 0], viewmatrix, projmatrix, [0, 0, 0])
    out = np.zeros(means.shape[0], np.uint8)
    lib().orc_mark_visible(ctypes.byref(v), _p(means), _p(out))
    return out.astype(bool)


def knn_dist2(points):
    pts = _np(points)
    out = np.zeros(pts.shape[0], np.float32)
    lib().orc_knn_dist2(int(pts.shape[0]), _p(pts), _p(out))
    return out


def higher_msb(n):
    return int(lib().orc_higher_msb(ctypes.c_uint32(int(n))))
