"""ctypes binding of libgsr_b200.so (C ABI in include/gsr_b200.h) + small host helpers.

PyTorch is used for device memory, streams and autograd plumbing only; all compute
goes through the hand-written sm_100a kernels in csrc/.  There is NO CPU or eager
fallback: if the library is missing or a call fails this module raises.
"""
import ctypes
import os
import weakref

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgsr_b200.so")
_lib = None


class GsrError(RuntimeError):
    pass


class gsr_view(ctypes.Structure):
    """Mirror of `struct gsr_view` (include/gsr_b200.h)."""
    _fields_ = [
        ("image_height", ctypes.c_int32),
        ("image_width", ctypes.c_int32),
        ("tanfovx", ctypes.c_float),
        ("tanfovy", ctypes.c_float),
        ("bg", ctypes.c_float * 3),
        ("scale_modifier", ctypes.c_float),
        ("viewmatrix", ctypes.c_float * 16),
        ("projmatrix", ctypes.c_float * 16),
        ("sh_degree", ctypes.c_int32),
        ("campos", ctypes.c_float * 3),
        ("prefiltered", ctypes.c_int32),
        ("debug", ctypes.c_int32),
    ]


class gsr_deform(ctypes.Structure):
    """Mirror of `struct gsr_deform` (include/gsr_b200.h)."""
    _fields_ = [
        ("mode", ctypes.c_int32),
        ("num_bodies", ctypes.c_int32),
        ("S", ctypes.c_void_p),
        ("theta", ctypes.c_void_p),
        ("body_id", ctypes.c_void_p),
    ]


class gsr_gemm(ctypes.Structure):
    """Mirror of `struct gsr_gemm` (include/gsr_b200.h)."""
    _fields_ = [
        ("M", ctypes.c_int32), ("N", ctypes.c_int32),
        ("A0_hi", ctypes.c_void_p), ("A0_lo", ctypes.c_void_p), ("K0", ctypes.c_int32), ("ldA0", ctypes.c_int64),
        ("A1_hi", ctypes.c_void_p), ("A1_lo", ctypes.c_void_p), ("K1", ctypes.c_int32), ("ldA1", ctypes.c_int64),
        ("B_hi", ctypes.c_void_p), ("B_lo", ctypes.c_void_p), ("ldB", ctypes.c_int64),
        ("mode", ctypes.c_int32), ("k_splits", ctypes.c_int32),
        ("bias", ctypes.c_void_p),
        ("mask_src", ctypes.c_void_p), ("ld_mask", ctypes.c_int32),
        ("out_hi", ctypes.c_void_p), ("out_lo", ctypes.c_void_p), ("ld_out", ctypes.c_int32),
        ("outT_hi", ctypes.c_void_p), ("outT_lo", ctypes.c_void_p), ("ld_outT", ctypes.c_int64),
        ("colsum", ctypes.c_void_p),
        ("error_flag", ctypes.c_void_p),
        ("mn_major", ctypes.c_int32),
    ]


class gsr_view_grads(ctypes.Structure):
    """include/gsr_b200.h: one view of a view-batched backward."""
    _fields_ = [("view", ctypes.POINTER(gsr_view)), ("radii", ctypes.c_void_p), ("geom_ws", ctypes.c_void_p),
                ("grad_ws", ctypes.c_void_p), ("dL_dmeans2D", ctypes.c_void_p)]


class gsr_view_fwd(ctypes.Structure):
    """include/gsr_b200.h: one view of a view-batched forward preprocess."""
    _fields_ = [("view", ctypes.POINTER(gsr_view)), ("radii", ctypes.c_void_p), ("geom_ws", ctypes.c_void_p)]


GEMM_RELU_SPLIT, GEMM_SPLIT, GEMM_PLAIN, GEMM_ATOMIC = 0, 1, 2, 3
DEFORM_NONE, DEFORM_PER_GAUSSIAN, DEFORM_RIGID_BODIES = 0, 1, 2

_P = ctypes.c_void_p
_SIGNATURES = {
    "gsr_last_error_string": (ctypes.c_char_p, []),
    "gsr_version": (ctypes.c_int, []),
    "gsr_profile_enable": (None, [ctypes.c_int]),
    "gsr_launch_count": (ctypes.c_ulonglong, [ctypes.c_int]),
    "gsr_profile_dump": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t]),
    "gsr_geom_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "gsr_image_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "gsr_binning_bytes": (ctypes.c_size_t, [ctypes.c_uint32, ctypes.c_int, ctypes.c_int]),
    "gsr_grad_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "gsr_geom_layout": (None, [ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]),
    "gsr_image_layout": (None, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]),
    "gsr_binning_layout": (None, [ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_size_t)]),
    "gsr_forward_preprocess": (ctypes.c_int, [ctypes.POINTER(gsr_view), ctypes.c_int, ctypes.c_int,
                                              _P, _P, _P, _P, _P, _P, _P, ctypes.POINTER(gsr_deform), _P,
                                              _P, _P, ctypes.c_size_t, _P, ctypes.c_int, _P]),
    "gsr_forward_render": (ctypes.c_int, [ctypes.POINTER(gsr_view), ctypes.c_int, ctypes.c_uint32, _P, _P, _P,
                                          ctypes.c_size_t, _P, _P, ctypes.c_int, _P]),
    "gsr_forward_preprocess_async": (ctypes.c_int, [ctypes.POINTER(gsr_view), ctypes.c_int, ctypes.c_int,
                                                    _P, _P, _P, _P, _P, _P, _P, ctypes.POINTER(gsr_deform), _P,
                                                    _P, _P, ctypes.c_size_t, _P]),
    "gsr_forward_render_capacity": (ctypes.c_int, [ctypes.POINTER(gsr_view), ctypes.c_int, ctypes.c_uint32, _P, _P, _P,
                                                   ctypes.c_size_t, _P, _P, _P, _P]),
    "gsr_backward": (ctypes.c_int, [ctypes.POINTER(gsr_view), ctypes.c_int, ctypes.c_int, ctypes.c_uint32,
                                    _P, _P, _P, _P, _P, _P, _P, ctypes.POINTER(gsr_deform), _P,
                                    _P, _P, _P, _P, _P,
                                    _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, ctypes.c_int, _P]),
    "gsr_forward_batched_slots_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "gsr_forward_batched_fill_slots": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(gsr_view_fwd), ctypes.c_int, ctypes.c_int, _P, ctypes.c_size_t]),
    "gsr_forward_preprocess_batched": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(gsr_view_fwd), _P, ctypes.c_int, ctypes.c_int, _P, _P, _P, _P, _P,
                                                       ctypes.POINTER(gsr_deform), _P, ctypes.c_size_t, _P]),
    "gsr_read_num_rendered": (ctypes.c_int, [_P, ctypes.c_int, _P, _P]),
    "gsr_backward_blend": (ctypes.c_int, [ctypes.POINTER(gsr_view), ctypes.c_int, ctypes.c_uint32, _P, _P, _P, _P, _P, _P]),
    "gsr_backward_batched_slots_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "gsr_backward_batched_fill_slots": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(gsr_view_grads), ctypes.c_int, ctypes.c_int, _P, ctypes.c_size_t]),
    "gsr_backward_gaussians_batched": (ctypes.c_int, [ctypes.c_int, _P, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, _P, _P, _P, _P,
                                                       ctypes.POINTER(gsr_deform), _P, _P, _P, _P, _P, _P, _P, ctypes.c_int, _P]),
    "gsr_debug_blend_stats": (ctypes.c_int, [ctypes.POINTER(gsr_view), ctypes.c_int, ctypes.c_uint32, _P, _P, _P, _P, _P]),
    "gsr_depth_order_ws_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "gsr_depth_order": (ctypes.c_int, [_P, ctypes.c_int, _P, ctypes.c_size_t, _P, _P, _P]),
    "gsr_debug_blend_group_stats": (ctypes.c_int, [ctypes.POINTER(gsr_view), ctypes.c_int, ctypes.c_uint32, _P, _P, _P, _P, _P]),
    "gsr_debug_exp_check": (ctypes.c_int, [ctypes.c_float, _P, _P]),
    "gsr_ssim_l1_loss_forward": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, ctypes.c_float, _P, _P, _P]),
    "gsr_ssim_l1_loss_backward": (ctypes.c_int, [_P, _P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, ctypes.c_float, _P, _P,
                                                 _P, _P]),
    "gsr_adam_step": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "gsr_mlp_gemm": (ctypes.c_int, [ctypes.POINTER(gsr_gemm), _P]),
    "gsr_mlp_split": (ctypes.c_int, [_P, ctypes.c_int64, _P, _P, _P]),
    "gsr_mlp_split_transpose": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.c_int, _P, _P, ctypes.c_int, _P]),
    "gsr_mlp_prepare": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_int, ctypes.c_int64, _P, _P, ctypes.c_int64, _P, _P, ctypes.c_int64, _P, _P]),
    "gsr_mlp_embed": (ctypes.c_int, [_P, ctypes.c_int, _P, _P, _P, _P, ctypes.c_int64, _P]),
    "gsr_mlp_embed_backward": (ctypes.c_int, [_P, ctypes.c_int, _P, _P, ctypes.c_int, _P]),
    "gsr_densify_stats": (ctypes.c_int, [ctypes.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "gsr_densify_decide": (ctypes.c_int, [ctypes.c_int, _P, _P, _P, ctypes.c_float, ctypes.c_float, _P, _P]),
    "gsr_densify_split": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "gsr_densify_prune": (ctypes.c_int, [ctypes.c_int, _P, _P, _P, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, _P, _P]),
    "gsr_deform_glue_forward": (ctypes.c_int, [ctypes.c_int] + [_P] * 11),
    "gsr_deform_glue_backward": (ctypes.c_int, [ctypes.c_int] + [_P] * 14),
    "gsr_mark_visible": (ctypes.c_int, [ctypes.POINTER(gsr_view), ctypes.c_int, _P, _P, _P]),
    "gsr_knn_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "gsr_knn_dist2": (ctypes.c_int, [ctypes.c_int, _P, _P, _P, ctypes.c_size_t, _P]),
    "gsr_exp_se3": (ctypes.c_int, [ctypes.c_int, _P, _P, _P, _P]),
    "gsr_exp_se3_backward": (ctypes.c_int, [ctypes.c_int, _P, _P, _P, _P, _P, _P]),
    "gsr_sort_bytes": (ctypes.c_size_t, [ctypes.c_uint32, ctypes.c_int, ctypes.c_int]),
    "gsr_sort_pairs": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, _P,
                                      ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), _P]),
    "gsr_sort_pairs32": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, _P,
                                        ctypes.c_size_t, ctypes.POINTER(ctypes.c_int), _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib_path():
    return _LIB_PATH


def load():
    """Load libgsr_b200.so (raises GsrError if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise GsrError("libgsr_b200.so is not built: run `python gaussian-splatting_deformable_b200/build.py` "
                       "(or __graft_entry__.build()).  There is no fallback path.")
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise GsrError("libgsr_b200: %s" % load().gsr_last_error_string().decode())


def ptr(t):
    """Device pointer of a tensor, or NULL for None / empty tensors (the reference passes
    empty CPU tensors for absent optionals, diff_gaussian_rasterization/__init__.py:197-207)."""
    if t is None or t.numel() == 0:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


# ---------------------------------------------------------------------------
# Host copies of the (tiny) camera tensors.  The kernels take the matrices in
# the constant bank, so the host needs their values; Camera objects keep the same
# tensors for the whole run, so this is a one-time D2H per camera, not per step.
# Keyed on object identity (validated through a weakref) and the in-place
# version counter, never on data_ptr (addresses are recycled by the allocator).
# ---------------------------------------------------------------------------
_host_cache = {}


def host_values(t, n):
    key = id(t)
    hit = _host_cache.get(key)
    if hit is not None and hit[0]() is t and hit[1] == t._version:
        return hit[2]
    vals = t.detach().to(dtype=torch.float32, device="cpu").contiguous().reshape(-1).tolist()
    if len(vals) != n:
        raise GsrError("expected a tensor with %d elements, got %d" % (n, len(vals)))
    if len(_host_cache) > 8192:
        _host_cache.clear()
    try:
        _host_cache[key] = (weakref.ref(t), t._version, vals)
    except TypeError:
        pass
    return vals


def make_view(rs, sh_degree=None):
    """Build the host-side gsr_view from a GaussianRasterizationSettings tuple."""
    v = gsr_view()
    v.image_height = int(rs.image_height)
    v.image_width = int(rs.image_width)
    v.tanfovx = float(rs.tanfovx)
    v.tanfovy = float(rs.tanfovy)
    v.bg[:] = host_values(rs.bg, 3)
    v.scale_modifier = float(rs.scale_modifier)
    v.viewmatrix[:] = host_values(rs.viewmatrix, 16)
    v.projmatrix[:] = host_values(rs.projmatrix, 16)
    v.sh_degree = int(rs.sh_degree if sh_degree is None else sh_degree)
    v.campos[:] = host_values(rs.campos, 3)
    v.prefiltered = int(bool(rs.prefiltered))
    v.debug = int(bool(rs.debug))
    return v


_pinned = {}


def pinned_u32(device):
    """One pinned 4-byte mailbox per device for the num_rendered read-back."""
    key = str(device)
    buf = _pinned.get(key)
    if buf is None:
        buf = torch.zeros(1, dtype=torch.int32).pin_memory()
        _pinned[key] = buf
    return buf


_status_ring = {}


def pinned_status(device, slots=512):
    """A 4-word pinned slot for the asynchronous status read-back of a sync-free forward, from a per-device ring (a slot is
    reused `slots` calls later, long after its copy has landed)."""
    key = str(device)
    ring = _status_ring.get(key)
    if ring is None:
        ring = [torch.zeros((slots, 4), dtype=torch.int32).pin_memory(), 0]
        _status_ring[key] = ring
    i = ring[1]
    ring[1] = (i + 1) % slots
    return ring[0][i]


def geom_layout(P):
    out = (ctypes.c_size_t * 6)()
    load().gsr_geom_layout(int(P), out)
    return dict(zip(("depths", "tiles_touched", "recs", "clamped", "depth_order", "cov3D"), list(out)))


def image_layout(W, H):
    out = (ctypes.c_size_t * 3)()
    load().gsr_image_layout(int(W), int(H), out)
    return dict(zip(("final_T", "n_contrib", "ranges"), list(out)))


def binning_layout(R, W, H):
    out = (ctypes.c_size_t * 4)()
    load().gsr_binning_layout(int(R), int(W), int(H), out)
    return dict(zip(("keys_sorted", "point_list", "tile_ids_sorted", "scratch"), list(out)))


def profile_enable(on=True):
    load().gsr_profile_enable(1 if on else 0)


def launch_count(reset=False):
    return int(load().gsr_launch_count(1 if reset else 0))


def profile_dump():
    """{kernel name: (launches, total_ms)} measured with CUDA events on the launching stream."""
    buf = ctypes.create_string_buffer(8192)
    load().gsr_profile_dump(buf, 8192)
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split()
        out[name] = (int(cnt), float(ms))
    return out
