"""Drop-in replacement for the reference's `diff_gaussian_rasterization` package.

Same names, argument meaning, return values and error behaviour as
submodules/diff-gaussian-rasterization/diff_gaussian_rasterization/__init__.py
(GaussianRasterizationSettings :157-169, GaussianRasterizer :171-220,
rasterize_gaussians :21-42, _RasterizeGaussians :44-155), backed by
libgsr_b200.so (hand-written sm_100a kernels) through ctypes instead of the
reference's pybind `_C` module.

Additions that do not disturb reference call sites (keyword-only, default None):
`GaussianRasterizer.forward(..., se3_S=, se3_theta=, body_id=)` fuses the SE3
exponential-map deformation of scene/rigid_body.py into the preprocess kernel,
forward and backward.  The deformed means of the last call are available as
`rasterizer.deformed_means`.
`accumulate_grads={"means3D": buf, "opacities": buf, "shs": buf, "scales": buf,
"rotations": buf, "se3_S": buf, "se3_theta": buf}` (any subset): the backward kernel
adds this view's gradient straight into the given fp32 buffers (e.g. views into one
flat all-reduce buffer, see view_parallel.FlatGradBuffer) and autograd receives no
gradient for those inputs - view-batched training then needs no per-view
accumulation pass.
"""
import ctypes
import os
import warnings
from typing import NamedTuple

import torch
import torch.nn as nn

import gsr_runtime as _rt


def _c(t):
    """Contiguous fp32 view of an optional tensor (None for absent/empty)."""
    if t is None or t.numel() == 0:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _empty_like_arg(t):
    return torch.Tensor([]) if t is None else t


class _Deform:
    """Optional fused SE3 deformation (see gsr_deform in include/gsr_b200.h)."""
    __slots__ = ("mode", "S", "theta", "body_id", "num_bodies")

    def __init__(self, S, theta, body_id):
        self.S, self.theta, self.body_id = S, theta, body_id
        if S is None:
            self.mode, self.num_bodies = _rt.DEFORM_NONE, 0
        elif body_id is None:
            self.mode, self.num_bodies = _rt.DEFORM_PER_GAUSSIAN, 0
        else:
            self.mode, self.num_bodies = _rt.DEFORM_RIGID_BODIES, int(S.shape[0])

    def c_struct(self):
        d = _rt.gsr_deform()
        d.mode = self.mode
        d.num_bodies = self.num_bodies
        d.S = self.S.data_ptr() if self.S is not None else None
        d.theta = self.theta.data_ptr() if self.theta is not None else None
        d.body_id = self.body_id.data_ptr() if self.body_id is not None else None
        return d


_ACC_BITS = {"means3D": 1, "opacities": 2, "shs": 4, "scales": 8, "rotations": 16, "se3_S": 32, "se3_theta": 32}


def rasterize_gaussians(
    means3D,
    means2D,
    sh,
    colors_precomp,
    opacities,
    scales,
    rotations,
    cov3Ds_precomp,
    raster_settings,
    se3_S=None,
    se3_theta=None,
    body_id=None,
    accumulate_grads=None,
    aux=None,
):
    return _RasterizeGaussians.apply(
        means3D,
        means2D,
        sh,
        colors_precomp,
        opacities,
        scales,
        rotations,
        cov3Ds_precomp,
        raster_settings,
        se3_S,
        se3_theta,
        body_id,
        accumulate_grads,
        aux,
    )


# ---- sync-free forward (opt-in): size the duplicate list from a capacity instead of reading num_rendered back ----------
# The reference reads num_rendered back in the middle of every forward (rasterizer_impl.cu:281) and so does the default
# path here: the list workspace is sized from it.  With set_sync_free(True) (or GSR_SYNC_FREE=1) the first forward of a
# (device, P, W, H) configuration measures num_rendered the usual way; later ones size the workspace from the largest
# num_rendered seen so far times `margin` and enqueue the whole forward - and its backward - without any host round
# trip.  A view's real counters arrive asynchronously in pinned memory and are looked at, without waiting, at the next
# forward of the same configuration (or, waiting, by check_sync_free()).  If a view overflowed its workspace the device
# emptied every tile - background image, zero gradients - the capacity is raised and that later call raises GsrError.
_SYNC_FREE = {"on": os.environ.get("GSR_SYNC_FREE", "0") == "1", "margin": float(os.environ.get("GSR_SYNC_FREE_MARGIN", "1.25"))}
_capacity = {}
_pending = {}          # key -> list of (event, pinned status words, capacity) of forwards not looked at yet


def set_sync_free(on=True, margin=None):
    _SYNC_FREE["on"] = bool(on)
    if margin is not None:
        _SYNC_FREE["margin"] = float(margin)
    if not on:
        _capacity.clear()
        _pending.clear()


def _note_num_rendered(key, R):
    cap = int(R * _SYNC_FREE["margin"]) + (1 << 16)
    if cap > _capacity.get(key, 0):
        _capacity[key] = cap


def _check_pending(key, blocking):
    """Look at the status words of earlier sync-free forwards of configuration `key` (all of them if blocking, else those
    whose copy has landed); raises GsrError if one of them overflowed its workspace."""
    todo = _pending.get(key)
    if not todo:
        return
    rest, overflowed = [], None
    for ev, status, cap in todo:
        if not blocking and not ev.query():
            rest.append((ev, status, cap))
            continue
        ev.synchronize()
        ov, _, _, R = (int(x) & 0xFFFFFFFF for x in status.tolist())
        _note_num_rendered(key, R)
        if ov:
            overflowed = (R, cap)
    _pending[key] = rest
    if overflowed is not None:
        raise _rt.GsrError("sync-free forward: an earlier view emitted %d (Gaussian, tile) pairs but its workspace was sized for %d, "
                           "so its image was the background only and its gradients zero.  The capacity has been raised - repeat "
                           "that step (or call diff_gaussian_rasterization.set_sync_free(False))." % overflowed)


def check_sync_free():
    """Wait for every sync-free forward issued so far and raise if one overflowed (call it wherever a step's results are
    consumed on the host anyway, e.g. where the loss is logged)."""
    for key in list(_pending):
        _check_pending(key, blocking=True)


def cpu_deep_copy_tuple(input_tuple):
    # diff_gaussian_rasterization/__init__.py:17-19
    copied_tensors = [item.cpu().clone() if isinstance(item, torch.Tensor) else item for item in input_tuple]
    return tuple(copied_tensors)


def _bucket(nbytes):
    """Binning workspace size for `nbytes` of need: num_rendered differs from view to view, and exact sizes make the
    caching allocator split and re-cudaMalloc blocks (milliseconds of stall once several streams' pools fragment).
    Round up to a power of two below 64 MB (small scenes stay small) and to a multiple of 64 MB above."""
    if nbytes >= (64 << 20):
        return ((nbytes + (64 << 20) - 1) // (64 << 20)) * (64 << 20)
    b = 1 << 20
    while b < nbytes:
        b <<= 1
    return b


class GaussianForwardBatch:
    """The forward's per-Gaussian half (SE3 deform, projection, covariance, SH colour: preprocessCUDA) for SEVERAL views of
    the same inputs in one pass (libgsr_b200: gsr_forward_preprocess_batched): every parameter record is read once per step
    instead of once per view, SE3 and cov3D are evaluated once.  Each view's workspace receives byte for byte what the
    per-view call writes, so everything downstream - and every result - is unchanged.

        fwd = GaussianForwardBatch(settings_list, means3D=xyz, opacities=op, shs=shs, scales=s, rotations=r, se3_S=S, se3_theta=th)
        color, radii = GaussianRasterizer(settings_list[i])(means3D=xyz, ..., prepared=fwd.prepared(i))

    The rasterizer call must be given the same tensors.  Requirements (checked): SH colours with 16 coefficients (32-byte
    aligned), scales + rotations, one scale_modifier / sh_degree for the batch, no prefiltered / debug settings.  The
    workspaces are allocated on the current stream: join the streams the views ran on into it before the batch is dropped
    (view_parallel.render_views does)."""

    def __init__(self, settings_list, means3D, opacities, shs, scales, rotations, se3_S=None, se3_theta=None, body_id=None):
        lib = _rt.load()
        if not means3D.is_cuda:
            raise _rt.GsrError("libgsr_b200 runs on CUDA tensors only (no CPU fallback)")
        dev = means3D.device
        P = int(means3D.shape[0])
        n = len(settings_list)
        if n == 0 or P == 0:
            raise _rt.GsrError("GaussianForwardBatch: needs at least one view and one Gaussian")
        if (se3_S is None) != (se3_theta is None):
            raise Exception('Please provide both se3_S and se3_theta, or neither!')
        for name, t in (("means3D", means3D), ("opacities", opacities), ("shs", shs), ("scales", scales), ("rotations", rotations),
                        ("se3_S", se3_S), ("se3_theta", se3_theta)):
            if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
                # a converted copy would not be the tensor the rasterizer call is given later
                raise _rt.GsrError("GaussianForwardBatch: %s must be a contiguous float32 tensor" % name)
        self.tensors = tuple(_c(t) for t in (means3D, opacities, shs, scales, rotations))
        means_c, opac_c, sh_c, scales_c, rots_c = self.tensors
        if sh_c is None or sh_c.dim() != 3 or int(sh_c.shape[1]) != 16 or sh_c.data_ptr() & 31:
            raise _rt.GsrError("GaussianForwardBatch needs SH colours [P, 16, 3] in a 32-byte aligned tensor")
        self.deform = _Deform(_c(se3_S), _c(se3_theta), body_id.to(torch.int32).contiguous() if body_id is not None else None)
        self.P, self.M, self.device = P, 16, dev
        self.settings = list(settings_list)
        self.views = [_rt.make_view(rs) for rs in self.settings]
        with torch.cuda.device(dev):
            stream = _rt.stream_ptr(dev)
            gbytes = lib.gsr_geom_bytes(P)
            gstride = (gbytes + 255) // 256 * 256                      # one allocation each, sliced per view (host time)
            geoms = torch.empty(n * gstride, dtype=torch.uint8, device=dev)
            self.geoms = [geoms[j * gstride: j * gstride + gbytes] for j in range(n)]
            radii = torch.empty((n, P), dtype=torch.int32, device=dev)
            self.radii = [radii[j] for j in range(n)]
            self.means_def = torch.empty((P, 3), dtype=torch.float32, device=dev) if self.deform.mode else None
            arr = (_rt.gsr_view_fwd * n)()
            for j in range(n):
                arr[j].view = ctypes.pointer(self.views[j])
                arr[j].radii = self.radii[j].data_ptr()
                arr[j].geom_ws = self.geoms[j].data_ptr()
            nbytes = lib.gsr_forward_batched_slots_bytes(n)
            host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            _rt.check(lib.gsr_forward_batched_fill_slots(n, arr, P, 16, host.data_ptr(), nbytes))
            slots = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            slots.copy_(host, non_blocking=True)
            _rt.check(lib.gsr_forward_preprocess_batched(
                n, arr, _rt.ptr(slots), P, 16, _rt.ptr(means_c), _rt.ptr(scales_c), _rt.ptr(rots_c), _rt.ptr(opac_c), _rt.ptr(sh_c),
                self.deform.c_struct(), _rt.ptr(self.means_def), gbytes, stream))
            self.event = torch.cuda.Event()
            self.event.record(torch.cuda.current_stream(dev))

    def __len__(self):
        return len(self.views)

    def prepared(self, i):
        """The handle to pass as `prepared=` to the rasterizer call of view i."""
        return (self, int(i))


class GaussianBackwardBatch:
    """`accumulate_grads=` target that also BATCHES the per-Gaussian half of the backward over the views of a step.

    A training step renders several views of the same parameters and sums their gradients.  Passed as `accumulate_grads`
    (in place of the plain {input name: gradient buffer} dict it wraps), this object makes each view's backward run only
    the blend backward (per-pixel work -> the view's 48-byte per-Gaussian records); `flush()` then runs ONE kernel over
    the Gaussians for all pending views (libgsr_b200: gsr_backward_gaussians_batched): every parameter record is read
    once and the running gradients are updated once per step instead of once per view.  The result equals the per-view
    path up to fp32 summation order.

    Requirements (checked): SH colours with 16 coefficients, scales + rotations (no precomputed colours / covariances),
    every differentiable Gaussian input present in `targets`, 32-byte aligned SH tensors, and all views of a batch given
    the SAME input tensors (a view with other inputs flushes the pending ones first).  The view-space gradients
    (`means2D`'s, which the reference reads for its densification statistics) are not returned through autograd in this
    mode: after `flush()` they are in `viewspace_grads`, one [P, 3] tensor per view in the order the backwards ran."""

    def __init__(self, targets, chunks=1, after_chunk=None):
        """`chunks` > 1: flush() runs the kernel over that many consecutive ranges of Gaussians and calls
        `after_chunk(first, count)` after enqueueing each - e.g. to start the all-reduce of that range of the gradient
        buffers (view_parallel.FlatGradBuffer.all_reduce_rows) while the next range is computed."""
        self.targets = dict(targets)
        self.chunks = max(1, int(chunks))
        self.after_chunk = after_chunk
        self.viewspace_grads = []
        self._items = []
        self._shared = None
        self._key = None

    def __len__(self):
        return len(self._items)

    def __del__(self):
        if getattr(self, "_items", None):
            warnings.warn("GaussianBackwardBatch dropped with %d view(s) pending: flush() was never called, their per-Gaussian "
                          "gradients were not computed" % len(self._items))

    # dict protocol used by the forward's argument checks
    def items(self):
        return self.targets.items()

    def _add(self, key, shared, item):
        if self._items and key != self._key:
            self.flush()
        if not self._items:
            self.viewspace_grads = []
        self._key, self._shared = key, shared
        self._items.append(item)

    def flush(self):
        """Run the batched per-Gaussian backward for the pending views on the current stream (which is made to wait for
        the streams the views ran on).  No-op when nothing is pending."""
        items, sh = self._items, self._shared
        self._items, self._shared, self._key = [], None, None
        if not items:
            return
        lib = _rt.load()
        dev = sh["means3D"].device
        P, M = sh["P"], sh["M"]
        n = len(items)
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream(dev)
            arr = (_rt.gsr_view_grads * n)()
            for j, it in enumerate(items):
                cur.wait_event(it["event"])
                arr[j].view = ctypes.pointer(it["view"])
                arr[j].radii = it["radii"].data_ptr()
                arr[j].geom_ws = it["geom"].data_ptr()
                arr[j].grad_ws = it["grad_ws"].data_ptr()
                arr[j].dL_dmeans2D = it["grad_means2D"].data_ptr()
            nbytes = lib.gsr_backward_batched_slots_bytes(n)
            host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            _rt.check(lib.gsr_backward_batched_fill_slots(n, arr, P, M, host.data_ptr(), nbytes))
            slots = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            slots.copy_(host, non_blocking=True)
            t = self.targets
            mask = 0
            for k in t:
                mask |= _ACC_BITS[k]
            deform = _Deform(sh["tw_S"], sh["tw_theta"], sh["body_id"])
            means_def = items[0]["means_def"]
            chunks = min(self.chunks, max(1, P // 4096))
            step = ((P + chunks - 1) // chunks + 127) // 128 * 128
            for first in range(0, P, step):
                count = min(step, P - first)
                _rt.check(lib.gsr_backward_gaussians_batched(
                    n, _rt.ptr(slots), float(items[0]["view"].scale_modifier), P, first, count, M, _rt.ptr(sh["means3D"]),
                    _rt.ptr(means_def), _rt.ptr(sh["scales"]), _rt.ptr(sh["rotations"]), _rt.ptr(sh["sh"]), deform.c_struct(),
                    _rt.ptr(t["means3D"]), _rt.ptr(t["opacities"]), _rt.ptr(t["shs"]), _rt.ptr(t["scales"]), _rt.ptr(t["rotations"]),
                    _rt.ptr(t.get("se3_S")), _rt.ptr(t.get("se3_theta")), mask, _rt.stream_ptr(dev)))
                if self.after_chunk is not None:
                    self.after_chunk(first, count)
            # The views' workspaces were allocated on the views' streams and are read here on another one.  Instead of
            # record_stream (which parks the blocks until an event is polled and sends the allocator to cudaMalloc when
            # the host runs ahead), every view stream is ordered after this kernel: whatever reuses the blocks there
            # comes after their last use.
            done = torch.cuda.Event()
            done.record(cur)
            for st in {it["stream"] for it in items}:
                if st != cur:
                    st.wait_event(done)
        self.viewspace_grads = [it["grad_means2D"] for it in items]


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(
        ctx,
        means3D,
        means2D,
        sh,
        colors_precomp,
        opacities,
        scales,
        rotations,
        cov3Ds_precomp,
        raster_settings,
        se3_S=None,
        se3_theta=None,
        body_id=None,
        accumulate_grads=None,
        aux=None,
    ):
        lib = _rt.load()
        debug = bool(raster_settings.debug)
        if debug:
            # reference :83-90: copy the arguments before they can be corrupted; dumped if the forward fails
            cpu_args = cpu_deep_copy_tuple((raster_settings.bg, means3D, colors_precomp, opacities, scales, rotations,
                                            raster_settings.scale_modifier, cov3Ds_precomp, raster_settings.viewmatrix,
                                            raster_settings.projmatrix, raster_settings.tanfovx, raster_settings.tanfovy,
                                            raster_settings.image_height, raster_settings.image_width, sh,
                                            raster_settings.sh_degree, raster_settings.campos, raster_settings.prefiltered,
                                            raster_settings.debug, se3_S, se3_theta, body_id))
            try:
                out = _RasterizeGaussians._forward(ctx, lib, means3D, means2D, sh, colors_precomp, opacities, scales,
                                                   rotations, cov3Ds_precomp, raster_settings, se3_S, se3_theta, body_id,
                                                   accumulate_grads, aux)
                torch.cuda.synchronize(means3D.device)    # CHECK_CUDA(.., debug) syncs after every launch (auxiliary.h:166-172)
                return out
            except Exception as ex:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise ex
        return _RasterizeGaussians._forward(ctx, lib, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                            cov3Ds_precomp, raster_settings, se3_S, se3_theta, body_id, accumulate_grads, aux)

    @staticmethod
    def _forward(ctx, lib, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                 raster_settings, se3_S, se3_theta, body_id, accumulate_grads, aux):
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            # rasterize_points.cu:57-59
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        if not means3D.is_cuda:
            raise _rt.GsrError("libgsr_b200 runs on CUDA tensors only (no CPU fallback)")
        dev = means3D.device
        P = int(means3D.shape[0])
        H, W = int(raster_settings.image_height), int(raster_settings.image_width)

        means3D_c = _c(means3D) if P else means3D
        sh_c, colors_c, opac_c = _c(sh), _c(colors_precomp), _c(opacities)
        scales_c, rots_c, cov_c = _c(scales), _c(rotations), _c(cov3Ds_precomp)
        M = int(sh_c.shape[1]) if sh_c is not None else 0
        deform = _Deform(_c(se3_S), _c(se3_theta),
                         body_id.to(torch.int32).contiguous() if body_id is not None else None)

        color = torch.empty((3, H, W), dtype=torch.float32, device=dev)
        prep = aux.get("prepared") if aux else None
        if prep is not None:
            # the per-Gaussian half was done for the whole batch of views (GaussianForwardBatch): take its results
            fb, vi = prep
            same = (means3D_c, opac_c, sh_c, scales_c, rots_c)
            if (fb.P != P or fb.device != dev or any(a is None or a.data_ptr() != b.data_ptr() for a, b in zip(same, fb.tensors)) or
                    fb.deform.mode != deform.mode or (deform.mode and (fb.deform.S.data_ptr() != deform.S.data_ptr() or
                                                                       fb.deform.theta.data_ptr() != deform.theta.data_ptr()))):
                raise _rt.GsrError("prepared=: the rasterizer call must be given the tensors the GaussianForwardBatch was built from")
            if (colors_c is not None and colors_c.numel()) or (cov_c is not None and cov_c.numel()):
                raise _rt.GsrError("prepared=: precomputed colours / covariances take the per-view path")
            if fb.settings[vi] is not raster_settings and tuple(fb.settings[vi][:4]) != tuple(raster_settings[:4]):
                raise _rt.GsrError("prepared=: view %d of the batch was prepared for other raster settings" % vi)
            radii, means_def, view = fb.radii[vi], fb.means_def, fb.views[vi]
            torch.cuda.current_stream(dev).wait_event(fb.event)
        else:
            radii = torch.empty((P,), dtype=torch.int32, device=dev)
            means_def = torch.empty((P, 3), dtype=torch.float32, device=dev) if deform.mode else None
            view = _rt.make_view(raster_settings)
        with torch.cuda.device(dev):
            stream = _rt.stream_ptr(dev)
            geom = fb.geoms[vi] if prep is not None else torch.empty(lib.gsr_geom_bytes(P), dtype=torch.uint8, device=dev)
            img = torch.empty(lib.gsr_image_bytes(W, H), dtype=torch.uint8, device=dev)
            mailbox = _rt.pinned_u32(dev)
            num_rendered = 0
            dstruct = deform.c_struct()
            key = (dev.index, P, W, H)
            cap = None
            if _SYNC_FREE["on"] and P != 0 and not raster_settings.debug and not raster_settings.prefiltered:
                _check_pending(key, blocking=False)
                cap = _capacity.get(key)
            if cap:
                if prep is None:
                    _rt.check(lib.gsr_forward_preprocess_async(
                        view, P, M, _rt.ptr(means3D_c), _rt.ptr(scales_c), _rt.ptr(rots_c), _rt.ptr(opac_c),
                        _rt.ptr(sh_c), _rt.ptr(cov_c), _rt.ptr(colors_c), dstruct, _rt.ptr(means_def),
                        _rt.ptr(radii), _rt.ptr(geom), geom.numel(), stream))
                nbytes = lib.gsr_binning_bytes(cap, W, H)
                binning = torch.empty(_bucket(nbytes), dtype=torch.uint8, device=dev)
                status = _rt.pinned_status(dev)
                _rt.check(lib.gsr_forward_render_capacity(view, P, cap, _rt.ptr(radii), _rt.ptr(geom), _rt.ptr(binning),
                                                          binning.numel(), _rt.ptr(img), _rt.ptr(color), status.data_ptr(), stream))
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                _pending.setdefault(key, []).append((ev, status, cap))
                num_rendered = cap
            elif P != 0 and prep is not None:
                _rt.check(lib.gsr_read_num_rendered(_rt.ptr(geom), P, mailbox.data_ptr(), stream))
                torch.cuda.current_stream(dev).synchronize()
                num_rendered = int(mailbox.item()) & 0xFFFFFFFF
            elif P != 0:
                _rt.check(lib.gsr_forward_preprocess(
                    view, P, M, _rt.ptr(means3D_c), _rt.ptr(scales_c), _rt.ptr(rots_c), _rt.ptr(opac_c),
                    _rt.ptr(sh_c), _rt.ptr(cov_c), _rt.ptr(colors_c), dstruct, _rt.ptr(means_def),
                    _rt.ptr(radii), _rt.ptr(geom), geom.numel(), mailbox.data_ptr(), 0, stream))
                num_rendered = int(mailbox.item()) & 0xFFFFFFFF
            if not cap:
                if _SYNC_FREE["on"]:
                    _note_num_rendered(key, num_rendered)
                nbytes = lib.gsr_binning_bytes(num_rendered, W, H)
                binning = torch.empty(_bucket(nbytes), dtype=torch.uint8, device=dev)
                _rt.check(lib.gsr_forward_render(view, P, num_rendered, _rt.ptr(radii), _rt.ptr(geom), _rt.ptr(binning),
                                                 binning.numel(), _rt.ptr(img), _rt.ptr(color), 0, stream))

        batch = accumulate_grads if isinstance(accumulate_grads, GaussianBackwardBatch) else None
        acc = dict(batch.targets) if batch is not None else (dict(accumulate_grads) if accumulate_grads else {})
        shapes = {"means3D": means3D, "opacities": opacities, "shs": sh, "scales": scales, "rotations": rotations,
                  "se3_S": se3_S, "se3_theta": se3_theta}
        for k, buf in acc.items():
            src = shapes.get(k)
            if k not in _ACC_BITS or src is None or src.numel() == 0:
                raise _rt.GsrError("accumulate_grads: no such differentiable input: %r" % k)
            if (not buf.is_cuda) or buf.dtype != torch.float32 or not buf.is_contiguous() or buf.numel() != src.numel():
                raise _rt.GsrError("accumulate_grads[%r] must be a contiguous CUDA float32 tensor of %d elements"
                                   % (k, src.numel()))
            if src.requires_grad and not src.is_leaf:
                # autograd receives None for an accumulated input: a non-leaf (e.g. scales = exp(_scaling)) would
                # silently lose the gradient of whatever produced it
                raise _rt.GsrError("accumulate_grads[%r]: the input is not a leaf tensor; accumulate only into the "
                                   "gradients of leaves (pass the activated tensor without accumulate_grads instead)" % k)
        if ("se3_S" in acc) != ("se3_theta" in acc):
            raise _rt.GsrError("accumulate_grads: give both se3_S and se3_theta or neither")
        if batch is not None:
            need = {"means3D", "opacities", "shs", "scales", "rotations"} | ({"se3_S", "se3_theta"} if deform.mode else set())
            if (sh_c is None or sh_c.numel() == 0 or M != 16 or scales_c is None or scales_c.numel() == 0 or
                    (colors_c is not None and colors_c.numel()) or (cov_c is not None and cov_c.numel())):
                raise _rt.GsrError("GaussianBackwardBatch needs SH colours with 16 coefficients and scales + rotations")
            if not need <= set(acc):
                raise _rt.GsrError("GaussianBackwardBatch: targets must hold a gradient buffer for each of %s" % sorted(need))
            if (sh_c.data_ptr() | acc["shs"].data_ptr()) & 31:
                raise _rt.GsrError("GaussianBackwardBatch: shs and its gradient buffer must be 32-byte aligned")
            if raster_settings.debug:
                raise _rt.GsrError("GaussianBackwardBatch: not available with debug=True")
        ctx.batch = batch
        ctx.acc = acc
        ctx.raster_settings = raster_settings
        ctx.num_rendered = num_rendered
        ctx.M = M
        ctx.deform_mode = deform.mode
        ctx.num_bodies = deform.num_bodies
        ctx.view = view
        ctx.flags = (sh is not None and sh.numel() > 0, colors_precomp is not None and colors_precomp.numel() > 0,
                     scales is not None and scales.numel() > 0, cov3Ds_precomp is not None and cov3Ds_precomp.numel() > 0)
        none = torch.empty(0, device=dev)
        ctx.save_for_backward(
            colors_c if colors_c is not None else none, means3D_c, scales_c if scales_c is not None else none,
            rots_c if rots_c is not None else none, cov_c if cov_c is not None else none, radii,
            sh_c if sh_c is not None else none, geom, binning, img,
            means_def if means_def is not None else none,
            deform.S if deform.S is not None else none, deform.theta if deform.theta is not None else none,
            deform.body_id if deform.body_id is not None else none)
        ctx.mark_non_differentiable(radii)
        if aux is not None:          # per-call results for the caller (no class-level state: rasterizers may interleave)
            aux["deformed_means"] = means_def
            aux["num_rendered"] = num_rendered if not cap else None        # sync-free: not known on the host yet
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _):
        if ctx.raster_settings.debug:
            # reference :132-139
            cpu_args = cpu_deep_copy_tuple((grad_out_color,) + tuple(ctx.saved_tensors))
            try:
                out = _RasterizeGaussians._backward(ctx, grad_out_color)
                torch.cuda.synchronize(grad_out_color.device)
                return out
            except Exception as ex:
                torch.save(cpu_args, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise ex
        return _RasterizeGaussians._backward(ctx, grad_out_color)

    @staticmethod
    def _backward(ctx, grad_out_color):
        lib = _rt.load()
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geom, binning, img,
         means_def, tw_S, tw_theta, body_id) = ctx.saved_tensors
        dev = means3D.device
        P = int(means3D.shape[0])
        M = ctx.M
        has_sh, has_colors, has_scales, has_cov = ctx.flags
        f32 = dict(dtype=torch.float32, device=dev)
        acc = ctx.acc
        if ctx.batch is not None and P != 0:
            # view-batched backward: the blend half now, the per-Gaussian half at GaussianBackwardBatch.flush()
            g = grad_out_color
            if g.dtype != torch.float32:
                g = g.float()
            g = g.contiguous()
            with torch.cuda.device(dev):
                grad_ws = torch.empty(lib.gsr_grad_bytes(P), dtype=torch.uint8, device=dev)
                _rt.check(lib.gsr_backward_blend(ctx.view, P, ctx.num_rendered, _rt.ptr(geom), _rt.ptr(binning), _rt.ptr(img),
                                                 _rt.ptr(grad_ws), _rt.ptr(g), _rt.stream_ptr(dev)))
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
            dm = ctx.deform_mode
            shared = dict(P=P, M=M, means3D=means3D, scales=scales, rotations=rotations, sh=sh,
                          tw_S=tw_S if dm else None, tw_theta=tw_theta if dm else None,
                          body_id=body_id if dm == _rt.DEFORM_RIGID_BODIES else None)
            key = (P, M, dm, means3D.data_ptr(), scales.data_ptr(), rotations.data_ptr(), sh.data_ptr(),
                   tw_S.data_ptr() if dm else 0, tw_theta.data_ptr() if dm else 0,
                   body_id.data_ptr() if dm == _rt.DEFORM_RIGID_BODIES else 0, float(ctx.view.scale_modifier))
            ctx.batch._add(key, shared, dict(view=ctx.view, radii=radii, geom=geom, grad_ws=grad_ws, event=ev,
                                             stream=torch.cuda.current_stream(dev),
                                             grad_means2D=torch.empty((P, 3), **f32), means_def=means_def if dm else None))
            return (None,) * 14
        mask = 0
        for k in acc:
            mask |= _ACC_BITS[k]

        def out(key, shape, needed=True):
            if not needed:
                return None
            return acc[key] if key in acc else torch.empty(shape, **f32)
        grad_means3D = out("means3D", (P, 3))
        grad_means2D = torch.empty((P, 3), **f32)
        grad_opacities = out("opacities", (P, 1))
        grad_colors = torch.empty((P, 3), **f32) if has_colors else None
        grad_cov3D = torch.empty((P, 6), **f32) if has_cov else None
        grad_sh = out("shs", (P, M, 3), has_sh)
        grad_scales = out("scales", (P, 3), has_scales)
        grad_rots = out("rotations", (P, 4), has_scales)
        grad_S = grad_theta = None
        deform = _Deform(tw_S if ctx.deform_mode else None, tw_theta if ctx.deform_mode else None,
                         body_id if ctx.deform_mode == _rt.DEFORM_RIGID_BODIES else None)
        if ctx.deform_mode == _rt.DEFORM_PER_GAUSSIAN:
            grad_S = out("se3_S", (P, 6))
            grad_theta = out("se3_theta", (P,))
        elif ctx.deform_mode == _rt.DEFORM_RIGID_BODIES:     # the kernel always accumulates body twists
            grad_S = acc["se3_S"] if "se3_S" in acc else torch.zeros((ctx.num_bodies, 6), **f32)
            grad_theta = acc["se3_theta"] if "se3_theta" in acc else torch.zeros((ctx.num_bodies,), **f32)
        if P != 0:
            g = grad_out_color
            if g.dtype != torch.float32:
                g = g.float()
            g = g.contiguous()
            with torch.cuda.device(dev):
                grad_ws = torch.empty(lib.gsr_grad_bytes(P), dtype=torch.uint8, device=dev)
                _rt.check(lib.gsr_backward(
                    ctx.view, P, M, ctx.num_rendered, _rt.ptr(means3D), _rt.ptr(means_def),
                    _rt.ptr(scales), _rt.ptr(rotations), _rt.ptr(sh), _rt.ptr(cov3Ds_precomp), _rt.ptr(colors_precomp),
                    deform.c_struct(), _rt.ptr(radii), _rt.ptr(geom), _rt.ptr(binning), _rt.ptr(img), _rt.ptr(grad_ws),
                    _rt.ptr(g), _rt.ptr(grad_means3D), _rt.ptr(grad_means2D), _rt.ptr(grad_opacities),
                    _rt.ptr(grad_colors), _rt.ptr(grad_cov3D), _rt.ptr(grad_sh), _rt.ptr(grad_scales),
                    _rt.ptr(grad_rots), _rt.ptr(grad_S), _rt.ptr(grad_theta), mask, _rt.stream_ptr(dev)))
        # Same order as the reference (__init__.py:143-153), then the SE3 extras.
        def ret(key, g):          # inputs whose gradient was accumulated in place get None from autograd
            return None if key in acc else g
        grads = (
            ret("means3D", grad_means3D),
            grad_means2D,
            ret("shs", grad_sh),
            grad_colors if has_colors else None,
            ret("opacities", grad_opacities),
            ret("scales", grad_scales),
            ret("rotations", grad_rots),
            grad_cov3D if has_cov else None,
            None,
            ret("se3_S", grad_S),
            ret("se3_theta", grad_theta),
            None,
            None,
            None,
        )
        return grads


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings
        self.deformed_means = None       # of this rasterizer's last forward with se3_S/se3_theta
        self.num_rendered = 0            # of this rasterizer's last forward

    def markVisible(self, positions):
        # Mark visible points (based on frustum culling for camera) with a boolean
        with torch.no_grad():
            lib = _rt.load()
            raster_settings = self.raster_settings
            P = int(positions.shape[0])
            pos = positions.float().contiguous()
            visible = torch.zeros((P,), dtype=torch.bool, device=positions.device)
            if P != 0:
                view = _rt.make_view(raster_settings._replace(
                    bg=raster_settings.bg if raster_settings.bg is not None else torch.zeros(3),
                    campos=raster_settings.campos if raster_settings.campos is not None else torch.zeros(3)))
                with torch.cuda.device(positions.device):
                    _rt.check(lib.gsr_mark_visible(view, P, _rt.ptr(pos), _rt.ptr(visible),
                                                   _rt.stream_ptr(positions.device)))
        return visible

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, *, se3_S=None, se3_theta=None, body_id=None, accumulate_grads=None, prepared=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        if (se3_S is None) != (se3_theta is None):
            raise Exception('Please provide both se3_S and se3_theta, or neither!')
        if body_id is not None and se3_S is None:
            raise Exception('body_id needs se3_S/se3_theta (one twist per rigid body)!')

        aux = {"prepared": prepared} if prepared is not None else {}
        out = rasterize_gaussians(
            means3D,
            means2D,
            _empty_like_arg(shs),
            _empty_like_arg(colors_precomp),
            opacities,
            _empty_like_arg(scales),
            _empty_like_arg(rotations),
            _empty_like_arg(cov3D_precomp),
            raster_settings,
            se3_S,
            se3_theta,
            body_id,
            accumulate_grads,
            aux,
        )
        self.deformed_means = aux.get("deformed_means")
        self.num_rendered = aux.get("num_rendered", 0)
        return out
