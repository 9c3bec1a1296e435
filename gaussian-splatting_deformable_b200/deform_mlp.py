"""Drop-in for the reference's deformation networks, evaluated on the tensor cores.

`DirectTemporalNeRF` (scene/gaussian_model.py:242-316) is the network `GaussianModel.get_xyz_all` queries every iteration
(:761-769): positional embedding of the positions (63) and of the view's time (21) -> 8 x 256 ReLU layers, the embedded
position re-concatenated in front of the activations after layer 4 -> four linear heads (dx 3, dx_scale 3, dx_rot 4,
mlp_shs 48).  `DirectTemporalNeRF_se3` (:99-173) has the same trunk with heads (w 3, v 3) that become the screw axis and
angle of the SE3 exponential map.

Same class names, constructor defaults, parameter names (`_time.N.weight`, `_time_out.weight`, ...: the reference's
`offset_model.pth` state-dicts load unchanged), call signature and return values.  The linears run as tcgen05
(`kind::tf32`) GEMMs with hi/lo operand splitting (csrc/mlp_gemm.cu): three tensor-core products per term give the
reference's fp32 result to ~1e-6 (plain TF32 or BF16 would miss it by ~1e-3), forward and backward.

Data flow (P points, everything stays on the device, no torch compute on P-sized tensors):
  gsr_mlp_embed       xyz -> embedding [P x 64]
  gsr_mlp_gemm x 8    hidden layers: relu(A W^T + b) -> fp32 [P x 256] = next layer's A operand
                      layer 0 folds the time embedding (one value per view) into its bias; layer 5 reads [embedding | hidden]
  gsr_mlp_gemm        the four heads as one [58 x 256] GEMM -> fp32 [P x 64]
backward, per layer:  dA = dZ W masked by the ReLU of the layer below (+ column sums = bias gradient), and
                      dW += dZ^T A as a split-K GEMM over the points (both operands read row-major as MN-major tensor-core
                      operands: no transposed copies) with an atomic epilogue;
  gsr_mlp_embed_backward  d embedding -> d xyz.
"""
import ctypes

import torch
import torch.nn as nn

import gsr_runtime as _rt

EMBED = 64          # 63 embedded position channels + 1 zero column (TMA rows are 16-byte multiples)
HEADS = 58          # 3 + 3 + 4 + 48
_SPLIT_K = 1024         # points per tile of a weight-gradient GEMM: the tensor core accumulates with truncation, so a long
                        # accumulation drifts (measured 1e-5 of max at K = 1350, 6e-5 at 6800); 1024 keeps it at fp32 level


def _planes(x):
    """(hi, lo) planes of a small fp32 tensor (weights)."""
    lib = _rt.load()
    x = x.detach().contiguous()
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    _rt.check(lib.gsr_mlp_split(x.data_ptr(), x.numel(), hi.data_ptr(), lo.data_ptr(), _rt.stream_ptr(x.device)))
    return hi, lo


def time_embedding(t, num_freqs=10):
    """Embedder.embed (gaussian_model.py:62-63) of one time value: [t, sin(t 2^0), cos(t 2^0), ...] (21 values)."""
    t = torch.as_tensor(t, dtype=torch.float32).reshape(1)
    out = [t]
    for k in range(num_freqs):
        f = torch.tensor(2.0 ** k, dtype=torch.float32, device=t.device)
        out += [torch.sin(t * f), torch.cos(t * f)]
    return torch.cat(out)


class _Gemm:
    """One gsr_mlp_gemm call; keeps the argument struct readable."""

    @staticmethod
    def run(dev, M, N, A0, K0, B, mode, *, A1=None, K1=0, bias=None, mask=None, out=None, outT=None, ld_out=None,
            ld_outT=0, colsum=None, k_splits=1, err=None, ldA0=None, ldA1=None, ldB=None, out_ptr_offset=0, mn_major=False):
        lib = _rt.load()
        g = _rt.gsr_gemm()
        g.M, g.N = int(M), int(N)
        pl = lambda t: (t[0], t[1]) if isinstance(t, tuple) else (t, None)       # (hi, lo) planes, or ONE fp32 plane
        a0, a0l = pl(A0)
        g.A0_hi, g.A0_lo, g.K0 = a0.data_ptr(), (a0l.data_ptr() if a0l is not None else None), int(K0)
        g.ldA0 = int(ldA0 if ldA0 is not None else a0.stride(0))
        if A1 is not None:
            a1, a1l = pl(A1)
            g.A1_hi, g.A1_lo, g.K1 = a1.data_ptr(), (a1l.data_ptr() if a1l is not None else None), int(K1)
            g.ldA1 = int(ldA1 if ldA1 is not None else a1.stride(0))
        b, bl = pl(B)
        g.B_hi, g.B_lo = b.data_ptr(), (bl.data_ptr() if bl is not None else None)
        g.ldB = int(ldB if ldB is not None else b.stride(0))
        g.mode, g.k_splits = int(mode), int(k_splits)
        if bias is not None:
            g.bias = bias.data_ptr()
        if mask is not None:
            g.mask_src, g.ld_mask = mask.data_ptr(), int(mask.stride(0))
        if isinstance(out, tuple):
            g.out_hi, g.out_lo = out[0].data_ptr(), out[1].data_ptr()
            g.ld_out = int(ld_out if ld_out is not None else out[0].stride(0))
        else:
            g.out_hi = out.data_ptr() + 4 * int(out_ptr_offset)
            g.ld_out = int(ld_out if ld_out is not None else out.stride(0))
        if outT is not None:
            t, tl = pl(outT)
            g.outT_hi, g.outT_lo, g.ld_outT = t.data_ptr(), (tl.data_ptr() if tl is not None else None), int(ld_outT)
        if colsum is not None:
            g.colsum = colsum.data_ptr()
        if err is not None:
            g.error_flag = err.data_ptr()
        g.mn_major = 1 if mn_major else 0
        _rt.check(lib.gsr_mlp_gemm(ctypes.byref(g), _rt.stream_ptr(dev)))


def _pack_weights(weights, biases, te, in_pts, n_layers):
    """The network's parameters in the layout the GEMMs read.  Returns (per-layer dicts, packed head weight, head bias).
    Layer 0: [256 x 84] -> position part [256 x 64] (column 63 zero) with the time part folded into the bias;
    the skip layer [256 x (63 + 256)] -> [256 x (64 + 256)] with a zero column at 63."""
    packed = []
    for i in range(n_layers):
        W, b = weights[i].detach(), biases[i].detach()
        if i == 0:
            Wp = torch.zeros((W.shape[0], EMBED), dtype=torch.float32, device=W.device)
            Wp[:, :in_pts] = W[:, :in_pts]
            beff = b + W[:, in_pts:] @ te
            packed.append(dict(W=Wp, b=beff.contiguous(), skip=False))
        elif W.shape[1] > W.shape[0]:
            Wp = torch.zeros((W.shape[0], EMBED + W.shape[0]), dtype=torch.float32, device=W.device)
            Wp[:, :in_pts] = W[:, :in_pts]
            Wp[:, EMBED:] = W[:, in_pts:]
            packed.append(dict(W=Wp, b=b.contiguous(), skip=True))
        else:
            packed.append(dict(W=W.contiguous(), b=b.contiguous(), skip=False))
    Wh = torch.cat([w.detach() for w in weights[n_layers:]], 0).contiguous()      # [58 x 256]
    bh = torch.cat([b.detach() for b in biases[n_layers:]], 0).contiguous()
    return packed, Wh, bh


class _DeformMLPFn(torch.autograd.Function):
    """forward(x [P,3], te [21], head_sizes, n_layers, grad_mode, *weights (L layers + heads), *biases) -> fp32 [P x 64] (heads in the first
    sum(head_sizes) columns)."""

    @staticmethod
    def forward(ctx, x, te, head_sizes, n_layers, grad_mode, *wb):
        lib = _rt.load()
        if not x.is_cuda:
            raise _rt.GsrError("deform_mlp runs on CUDA tensors only (no CPU fallback)")
        n_w = n_layers + len(head_sizes)
        weights, biases = wb[:n_w], wb[n_w:]
        dev = x.device
        P = int(x.shape[0])
        Wd = int(weights[0].shape[0])                       # 256
        in_pts = int(weights[0].shape[1]) - int(te.numel())
        need_grad = bool(grad_mode) and any(ctx.needs_input_grad)      # (grad mode is always off INSIDE forward)
        f32 = dict(dtype=torch.float32, device=dev)
        Pp = (P + 3) // 4 * 4
        x_c = x.detach().float().contiguous()
        packed, Wh, bh = _pack_weights(weights, biases, te.detach().to(dev), in_pts, n_layers)
        n_heads = int(Wh.shape[0])
        st = _rt.stream_ptr(dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            # activations travel as ONE fp32 plane each (the consumer GEMM splits its tiles in shared memory)
            E = torch.empty((P, EMBED), **f32)
            _rt.check(lib.gsr_mlp_embed(x_c.data_ptr(), P, E.data_ptr(), None, None, None, 0, st))
            acts = []
            prev = None
            pool = [] if need_grad else [torch.empty((P, Wd), **f32) for _ in range(2)]
            for i, L in enumerate(packed):
                Wp = _planes(L["W"])
                H = torch.empty((P, Wd), **f32) if need_grad else pool[i & 1]
                if i == 0:
                    _Gemm.run(dev, P, Wd, E, EMBED, Wp, _rt.GEMM_RELU_SPLIT, bias=L["b"], out=H, err=err)
                elif L["skip"]:
                    _Gemm.run(dev, P, Wd, E, EMBED, Wp, _rt.GEMM_RELU_SPLIT, A1=prev, K1=Wd, bias=L["b"], out=H, err=err)
                else:
                    _Gemm.run(dev, P, Wd, prev, Wd, Wp, _rt.GEMM_RELU_SPLIT, bias=L["b"], out=H, err=err)
                prev = H
                if need_grad:
                    acts.append(H)
            out = torch.empty((P, EMBED), **f32)
            if n_heads < EMBED:
                out[:, n_heads:].zero_()
            _Gemm.run(dev, P, n_heads, prev, Wd, _planes(Wh), _rt.GEMM_PLAIN, bias=bh, out=out, err=err)
        ctx.err = err
        if need_grad:
            ctx.state = dict(x=x_c, te=te.detach().to(dev), E=E, acts=acts, packed=packed, Wh=Wh, P=P, Pp=Pp,
                             Wd=Wd, in_pts=in_pts, n_layers=n_layers, n_heads=n_heads, head_sizes=tuple(head_sizes),
                             shapes=[tuple(w.shape) for w in weights])
        return out

    @staticmethod
    def backward(ctx, g_out):
        lib = _rt.load()
        s = ctx.state
        dev = g_out.device
        P, Pp, Wd, n_heads, n_layers = s["P"], s["Pp"], s["Wd"], s["n_heads"], s["n_layers"]
        f32 = dict(dtype=torch.float32, device=dev)
        st = _rt.stream_ptr(dev)
        err = ctx.err
        g = g_out.float()
        if g.stride(1) != 1:
            g = g.contiguous()
        splits = max(74, (P + _SPLIT_K - 1) // _SPLIT_K)
        with torch.cuda.device(dev):
            # dL/d heads is already a K-major fp32 plane [P x 64]; its column sums are the head biases' gradient
            if g.stride(0) % 4 or g.data_ptr() % 16:
                g = g.contiguous()
            dO = g
            db_heads = torch.zeros(EMBED, **f32)
            _rt.check(lib.gsr_mlp_prepare(g.data_ptr(), P, EMBED, g.stride(0), None, None, 0, None, None, 0, db_heads.data_ptr(), st))
            acts, E = s["acts"], s["E"]
            # Weight gradients dW = dZ^T X read dZ [P x 256] and X [P x K_in] as they lie (row-major, the reduction runs over
            # rows: MN-major tensor-core operands), split over the points with an atomic epilogue.
            def weight_grad(dZ_, M_, X, N_, out, **kw):
                _Gemm.run(dev, M_, N_, dZ_, P, X, _rt.GEMM_ATOMIC, out=out, k_splits=splits, mn_major=True, err=err, **kw)
            # heads: dWh = dO^T H_last ; dZ_last = (dO Wh) masked by relu(H_last)
            dWh = torch.zeros((n_heads, Wd), **f32)
            weight_grad(dO, n_heads, acts[-1], Wd, dWh)
            WhT = torch.zeros((Wd, EMBED), **f32)
            WhT[:, :n_heads] = s["Wh"].t()
            dZ, dZ2 = torch.empty((P, Wd), **f32), torch.empty((P, Wd), **f32)
            db = [torch.zeros(Wd, **f32) for _ in range(n_layers)]
            dWp = [torch.zeros_like(L["W"]) for L in s["packed"]]
            dE = torch.zeros((P, EMBED), **f32)
            _Gemm.run(dev, P, Wd, dO, EMBED, _planes(WhT), _rt.GEMM_SPLIT, mask=acts[-1], out=dZ, colsum=db[n_layers - 1], err=err)
            for i in range(n_layers - 1, -1, -1):
                L = s["packed"][i]
                if i == 0:
                    weight_grad(dZ, Wd, E, EMBED, dWp[i])
                elif L["skip"]:
                    weight_grad(dZ, Wd, E, EMBED, dWp[i], ld_out=EMBED + Wd)
                    weight_grad(dZ, Wd, acts[i - 1], Wd, dWp[i], ld_out=EMBED + Wd, out_ptr_offset=EMBED)
                else:
                    weight_grad(dZ, Wd, acts[i - 1], Wd, dWp[i])
                # input gradient of layer i
                WT = L["W"].t().contiguous()                      # [K_in x 256]
                if i == 0:
                    _Gemm.run(dev, P, EMBED, dZ, Wd, _planes(WT), _rt.GEMM_ATOMIC, out=dE, err=err)
                else:
                    if L["skip"]:
                        _Gemm.run(dev, P, EMBED, dZ, Wd, _planes(WT[:EMBED].contiguous()), _rt.GEMM_ATOMIC, out=dE, err=err)
                        WT = WT[EMBED:].contiguous()
                    _Gemm.run(dev, P, Wd, dZ, Wd, _planes(WT), _rt.GEMM_SPLIT, mask=acts[i - 1], out=dZ2, colsum=db[i - 1], err=err)
                    dZ, dZ2 = dZ2, dZ
            dx = torch.empty((P, 3), **f32)
            _rt.check(lib.gsr_mlp_embed_backward(s["x"].data_ptr(), P, dE.data_ptr(), dx.data_ptr(), 0, st))
        # unpack the padded layouts into the parameters' own shapes (0.5 M values in total)
        in_pts, te = s["in_pts"], s["te"]
        gw, gb = [], []
        for i, L in enumerate(s["packed"]):
            if i == 0:
                gw.append(torch.cat([dWp[i][:, :in_pts], torch.outer(db[i], te)], 1))
            elif L["skip"]:
                gw.append(torch.cat([dWp[i][:, :in_pts], dWp[i][:, EMBED:]], 1))
            else:
                gw.append(dWp[i])
            gb.append(db[i])
        o = 0
        for hs in s["head_sizes"]:
            gw.append(dWh[o:o + hs]); gb.append(db_heads[o:o + hs]); o += hs
        ctx.state = None
        return (dx if ctx.needs_input_grad[0] else None, None, None, None, None, *gw, *gb)


class _GlueFn(torch.autograd.Function):
    """(heads [P x 64], _xyz, _scaling, _rotation, _features_dc, _features_rest) -> (means3D, scales, rotations, shs): the
    reference's activation glue (gaussian_renderer/__init__.py:79,116,122,140) as one kernel each way."""

    @staticmethod
    def forward(ctx, heads, xyz, scaling, rotation, f_dc, f_rest):
        lib = _rt.load()
        if not heads.is_cuda:
            raise _rt.GsrError("deform_glue runs on CUDA tensors only (no CPU fallback)")
        dev = heads.device
        P = int(xyz.shape[0])
        c = lambda t: t.detach().float().contiguous()
        h, x, sc, ro, dc, fr = c(heads), c(xyz), c(scaling), c(rotation), c(f_dc), c(f_rest)
        if h.shape != (P, EMBED) or fr.numel() != P * 45 or dc.numel() != P * 3:
            raise _rt.GsrError("deform_glue: heads must be [P, 64] and the features [P,1,3] / [P,15,3] (SH degree 3)")
        f32 = dict(dtype=torch.float32, device=dev)
        means, scales, rots, shs = torch.empty((P, 3), **f32), torch.empty((P, 3), **f32), torch.empty((P, 4), **f32), torch.empty((P, 16, 3), **f32)
        with torch.cuda.device(dev):
            _rt.check(lib.gsr_deform_glue_forward(P, h.data_ptr(), x.data_ptr(), sc.data_ptr(), ro.data_ptr(), dc.data_ptr(), fr.data_ptr(),
                                                  means.data_ptr(), scales.data_ptr(), rots.data_ptr(), shs.data_ptr(), _rt.stream_ptr(dev)))
        ctx.save_for_backward(h, ro, scales)
        ctx.shapes = (tuple(f_dc.shape), tuple(f_rest.shape))
        return means, scales, rots, shs

    @staticmethod
    def backward(ctx, g_means, g_scales, g_rots, g_shs):
        lib = _rt.load()
        h, ro, scales = ctx.saved_tensors
        dev = h.device
        P = int(h.shape[0])
        f32 = dict(dtype=torch.float32, device=dev)
        c = lambda t: None if t is None else t.detach().float().contiguous()
        gm, gs, gr, gh = c(g_means), c(g_scales), c(g_rots), c(g_shs)
        need = ctx.needs_input_grad
        d_heads = torch.empty((P, EMBED), **f32)
        d_xyz = torch.empty((P, 3), **f32) if need[1] else None
        d_sc = torch.empty((P, 3), **f32) if need[2] else None
        d_ro = torch.empty((P, 4), **f32) if need[3] else None
        d_dc = torch.empty(ctx.shapes[0], **f32) if need[4] else None
        d_fr = torch.empty(ctx.shapes[1], **f32) if need[5] else None
        with torch.cuda.device(dev):
            _rt.check(lib.gsr_deform_glue_backward(P, h.data_ptr(), ro.data_ptr(), scales.data_ptr(), _rt.ptr(gm), _rt.ptr(gs), _rt.ptr(gr),
                                                   _rt.ptr(gh), d_heads.data_ptr(), _rt.ptr(d_xyz), _rt.ptr(d_sc), _rt.ptr(d_ro),
                                                   _rt.ptr(d_dc), _rt.ptr(d_fr), _rt.stream_ptr(dev)))
        return (d_heads if need[0] else None, d_xyz, d_sc, d_ro, d_dc, d_fr)


def deform_glue(heads, xyz, scaling, rotation, features_dc, features_rest):
    """What render() does between `pc.get_xyz_all` and the rasterizer call (gaussian_renderer/__init__.py:79,116,122,140),
    fused: returns (means3D, scales, rotations, shs) from the network's raw output `heads` [P x 64] (see
    DirectTemporalNeRF.heads) and the model's pre-activation tensors."""
    return _GlueFn.apply(heads, xyz, scaling, rotation, features_dc, features_rest)


def _check_pipeline(err):
    if int(err.item()) != 0:
        raise _rt.GsrError("deform_mlp: a tensor-core pipeline wait timed out (workspace corrupted?)")


class _Trunk(nn.Module):
    """Shared construction: `_time` = 8 linears (the skip layer takes the embedded position in front), same order of
    nn.Linear construction as the reference so that the same seed gives the same initial weights."""

    def _build(self, D, W, in_pts, in_time, skips):
        layers = [nn.Linear(in_pts + in_time, W)]
        for i in range(D - 1):
            layers += [nn.Linear(W + (in_pts if i in skips else 0), W)]
        return nn.ModuleList(layers)

    def _run(self, x, te, heads, check=False):
        weights = [l.weight for l in self._time] + [h.weight for h in heads]
        biases = [l.bias for l in self._time] + [h.bias for h in heads]
        sizes = tuple(int(h.weight.shape[0]) for h in heads)
        out = _DeformMLPFn.apply(x, te, sizes, len(self._time), torch.is_grad_enabled(), *weights, *biases)
        return out


class DirectTemporalNeRF(_Trunk):
    """scene/gaussian_model.py:242-316 (constructor arguments kept; only D = 8, W = 256, skips = [4] are built natively)."""

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, input_ch_time=1, output_ch=4, skips=[4],
                 use_viewdirs=False, memory=[], embed_fn=None, zero_canonical=True):
        super().__init__()
        if memory:
            raise NotImplementedError
        if W % 32 or (len(skips) > 1):
            raise _rt.GsrError("deform_mlp: width must be a multiple of 32 with at most one skip layer")
        self.D, self.W, self.skips = D, W, list(skips)
        self.input_ch = input_ch * 21                  # get_embedder(10, input_ch, 0)
        self.input_ch_time = 21                        # get_embedder(10, 1, 0)
        if self.input_ch != 63:
            raise _rt.GsrError("deform_mlp: the native embedding handles 3-D positions")
        self.input_ch_views, self.use_viewdirs, self.memory, self.zero_canonical = input_ch_views, use_viewdirs, memory, zero_canonical
        self._time = self._build(D, W, self.input_ch, self.input_ch_time, self.skips)
        self._time_out = nn.Linear(W, 3)
        self._time_out_scale = nn.Linear(W, 3)
        self._time_out_rot = nn.Linear(W, 4)
        self._time_out_shs = nn.Linear(W, 48)

    def forward(self, x, ts, iteration):
        # gaussian_model.py:303: "Only accepts all points from same time"
        if torch.is_tensor(ts):
            t0 = ts.reshape(-1)[:1]
            assert bool((ts[:, :1] == t0).all()), "Only accepts all points from same time"
            te = time_embedding(t0.detach().float().to(x.device))         # same torch sin / cos kernels as the reference's embedder
        else:
            te = time_embedding(float(ts)).to(x.device)
        n = x.shape[0]
        if iteration < 3000:                            # gaussian_model.py:305-310: the network's output is discarded
            z = lambda c: torch.zeros((n, c), dtype=torch.float32, device=x.device)
            return z(3), z(3), z(4), z(48)
        out = self._run(x, te, [self._time_out, self._time_out_scale, self._time_out_rot, self._time_out_shs])
        return out[:, 0:3], out[:, 3:6], out[:, 6:10], out[:, 10:58]

    def heads(self, x, ts, iteration=1 << 30):
        """The raw [P x 64] output (dx 0-2, dscale 3-5, drot 6-9, dshs 10-57, 6 zero columns) for `deform_glue`; zeros while
        iteration < 3000 like forward()."""
        if iteration < 3000:
            return torch.zeros((x.shape[0], EMBED), dtype=torch.float32, device=x.device)
        t0 = ts.reshape(-1)[:1].detach().float().to(x.device) if torch.is_tensor(ts) else float(ts)
        return self._run(x, time_embedding(t0).to(x.device), [self._time_out, self._time_out_scale, self._time_out_rot, self._time_out_shs])


def screw_from_raw(w_raw, v_raw, eps=0.0):
    """gaussian_model.py:161-164: theta = |w|, S = (w, v) / theta.  The reference has no guard (theta = 0 gives NaN);
    with eps > 0 points whose |w| <= eps get the identity (S = 0, theta = 0) instead."""
    theta = torch.norm(w_raw, dim=-1)
    if eps > 0.0:
        ok = theta > eps
        safe = torch.where(ok, theta, torch.ones_like(theta))
        S = torch.cat([w_raw, v_raw], -1) / safe[..., None] * ok[..., None]
        return S, theta * ok
    return torch.cat([w_raw / theta[..., None], v_raw / theta[..., None]], dim=-1), theta


class DirectTemporalNeRF_se3(_Trunk):
    """scene/gaussian_model.py:99-173.  `forward` returns the [N,4,4] transforms like the reference (through the drop-in
    rigid_body.exp_se3); `screw(x, ts)` returns (S [N,6], theta [N]) for the rasterizer's fused SE3 inputs
    (GaussianRasterizer(..., se3_S=S, se3_theta=theta)), which never materialises the transforms.
    The native trunk takes 3-D positions and one time per call and embeds them itself (input_ch 63 / input_ch_time 21 is
    what the reference's own embedders produce for it)."""

    def __init__(self, D=8, W=256, input_ch=63, input_ch_views=3, input_ch_time=21, output_ch=4, skips=[4],
                 use_viewdirs=False, memory=[], embed_fn=None, zero_canonical=True, theta_eps=0.0):
        super().__init__()
        if memory:
            raise NotImplementedError
        if input_ch != 63 or input_ch_time != 21:
            raise _rt.GsrError("deform_mlp: the native se3 trunk is built for the embedded inputs (63 + 21 channels)")
        self.D, self.W, self.skips, self.input_ch, self.input_ch_time = D, W, list(skips), input_ch, input_ch_time
        self.theta_eps = theta_eps
        self._time = self._build(D, W, input_ch, input_ch_time, self.skips)
        self._w = nn.Linear(W, 3)
        self._v = nn.Linear(W, 3)

    def raw_heads(self, x, ts):
        """(w_raw [N,3], v_raw [N,3]) = query_time (gaussian_model.py:141-150).  `x` is either the positions [N,3] or, as
        the reference's forward takes it, their embedding [N,63] (whose first three channels ARE the positions:
        include_input=True, gaussian_model.py:43-45); likewise `ts` is a time, [N,1] times or their [N,21] embedding."""
        if x.shape[1] == self.input_ch and self.input_ch != 3:
            x = x[:, :3]
        t0 = ts.reshape(-1)[:1].detach().float().to(x.device) if torch.is_tensor(ts) else float(ts)
        te = time_embedding(t0).to(x.device)
        out = self._run(x, te, [self._w, self._v])
        return out[:, 0:3], out[:, 3:6]

    def screw(self, x, ts):
        w, v = self.raw_heads(x, ts)
        return screw_from_raw(w, v, self.theta_eps)

    def forward(self, x, ts, iteration):
        import rigid_body
        S, theta = self.screw(x, ts)
        if iteration < 3000:                             # gaussian_model.py:169-171 (returns [N,3] zeros: the reference's shape bug, kept)
            return torch.zeros((x.shape[0], 3), dtype=torch.float32, device=x.device)
        return rigid_body.exp_se3(S, theta)
