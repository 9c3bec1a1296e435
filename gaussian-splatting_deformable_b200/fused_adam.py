"""Fused Adam over one flat parameter buffer (csrc/adam.cu) with the reference's optimizer surface.

The reference builds `torch.optim.Adam(l, lr=0.0, eps=1e-15)` from a list of per-tensor groups
`{'params': [tensor], 'lr': ..., 'name': ...}` (scene/gaussian_model.py:834-846) and drives the learning rates by
writing `param_group['lr']` (gaussian_model.py:869-886).  `FusedAdam` takes the same list, re-homes every
parameter as a view of ONE flat fp32 buffer (gradients: `view_parallel.FlatGradBuffer`, which is also the
all-reduce buffer; moments: two more flat buffers) and runs the whole step as one kernel.
Arithmetic is torch.optim.Adam's (torch 2.11; no amsgrad, weight decay or maximize).  `state_dict()` /
`load_state_dict()` use torch.optim.Adam's own layout, so the reference's checkpoints (scene/gaussian_model.py:698,725)
go both ways.  Differences: the step count is kept per GROUP (torch: per parameter; the reference always steps the
parameters of a group together), and a group whose gradients were not produced must be named in `step(skip=...)`
(torch skips parameters whose `.grad` is None; here gradients always exist as views of the flat buffer).
"""
import ctypes
import math

import torch

import gsr_runtime as _rt
import view_parallel


def get_expon_lr_func(lr_init, lr_final, lr_delay_steps=0, lr_delay_mult=1.0, max_steps=1000000):
    """Learning-rate schedule of utils/general_utils.py:29-62 (same name, signature and values): log-linear
    interpolation from lr_init (step 0) to lr_final (step max_steps), optionally eased in by a quarter sine that
    starts at lr_delay_mult; 0 for negative steps or when both rates are 0.  The reference writes the result into
    `param_group['lr']` every iteration (scene/gaussian_model.py:875-886); FusedAdam reads it from there."""
    log_init = math.log(lr_init) if lr_init > 0 else None
    log_final = math.log(lr_final) if lr_final > 0 else None

    def schedule(step):
        if step < 0 or (lr_init == 0.0 and lr_final == 0.0):
            return 0.0
        ease = 1.0
        if lr_delay_steps > 0:
            frac = min(max(step / lr_delay_steps, 0.0), 1.0)
            ease = lr_delay_mult + (1 - lr_delay_mult) * math.sin(0.5 * math.pi * frac)
        t = min(max(step / max_steps, 0.0), 1.0)
        return ease * math.exp(log_init * (1 - t) + log_final * t)

    return schedule


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.param_groups = []
        flat_params = []
        for g in groups:
            ps = [g["params"]] if isinstance(g["params"], torch.Tensor) else list(g["params"])
            d = dict(g)
            d["params"] = ps
            d.setdefault("lr", lr)
            d.setdefault("betas", betas)
            d.setdefault("eps", eps)
            self.param_groups.append(d)
            flat_params.extend(ps)
        if not flat_params:
            raise ValueError("optimizer got an empty parameter list")
        for p in flat_params:
            if not p.is_cuda or p.dtype != torch.float32:
                raise _rt.GsrError("FusedAdam: parameters must be CUDA float32 tensors (no CPU fallback)")
        self._params = flat_params
        dev = flat_params[0].device
        # same 32-byte-aligned layout as the gradient buffer (view_parallel.flat_layout): the Adam kernel walks all four
        # flat buffers with one index, and the rasterizer's vector accesses need aligned tensor starts whatever P is
        offs, total = view_parallel.flat_layout(flat_params)
        self.flat_params = torch.zeros(total, dtype=torch.float32, device=dev)
        self._begin = []
        it = iter(offs)
        for g in self.param_groups:
            first = True
            for p in g["params"]:
                off = next(it)
                if first:
                    self._begin.append(off)
                    first = False
                n = p.numel()
                self.flat_params[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_params[off:off + n].view_as(p)
        self._begin.append(total)
        self.grads = view_parallel.FlatGradBuffer(flat_params)       # p.grad = views of grads.flat
        self.exp_avg = torch.zeros_like(self.grads.flat)
        self.exp_avg_sq = torch.zeros_like(self.grads.flat)
        # torch.optim.Adam keeps one step count per parameter and skips parameters whose .grad is None (a dormant group,
        # e.g. the deformation network before its warm-up ends, keeps its moments and bias correction).  Gradients here
        # always exist (views of the flat buffer), so dormancy is stated explicitly: step(skip=[group names]).
        self._steps = [0] * len(self.param_groups)

    @property
    def state_step(self):
        return max(self._steps) if self._steps else 0

    # -- torch.optim surface the reference uses --
    def zero_grad(self, set_to_none=True):
        # the reference calls zero_grad(set_to_none=True) (train.py:683); gradients live in the flat buffer, so
        # they are zeroed in place and stay attached (accumulate_grads / the all-reduce use the same memory)
        self.grads.zero_()

    def _hyper(self, skip):
        """Per-group step sizes / bias corrections of THIS step (advances the step counts of the groups that step)."""
        ss, bc, b1s, b2s, es = [], [], [], [], []
        for i, g in enumerate(self.param_groups):
            if i in skip or g.get("name") in skip:
                # lr 0 and betas 1: m = lerp(m, g, 0), v = 1 v + 0 g g, p -= 0 -> the kernel leaves the group as it is
                ss.append(0.0); bc.append(1.0); b1s.append(1.0); b2s.append(1.0); es.append(g["eps"])
                continue
            self._steps[i] += 1
            t = self._steps[i]
            b1, b2 = g["betas"]
            bias_correction1 = 1 - b1 ** t
            bias_correction2 = 1 - b2 ** t
            ss.append(g["lr"] / bias_correction1)
            bc.append(bias_correction2 ** 0.5)
            b1s.append(b1); b2s.append(b2); es.append(g["eps"])
        return ss, bc, b1s, b2s, es

    def _launch(self, hyper, lo=0, hi=None):
        """The Adam kernel on elements [lo, hi) of the flat buffers (lo, hi multiples of 4)."""
        lib = _rt.load()
        n = len(self.param_groups)
        total = self._begin[-1]
        hi = total if hi is None else hi
        ss, bc, b1s, b2s, es = hyper
        clip = [min(max(b, lo), hi) - lo for b in self._begin]
        begin = (ctypes.c_ulonglong * (n + 1))(*clip)
        arr = lambda v: (ctypes.c_float * n)(*v)
        darr = lambda v: (ctypes.c_double * n)(*v)
        dev = self.flat_params.device
        off = 4 * lo
        with torch.cuda.device(dev):
            _rt.check(lib.gsr_adam_step(self.flat_params.data_ptr() + off, self.grads.flat.data_ptr() + off,
                                        self.exp_avg.data_ptr() + off, self.exp_avg_sq.data_ptr() + off, n, begin,
                                        arr(ss), arr(bc), darr(b1s), darr(b2s), arr(es), _rt.stream_ptr(dev)))

    def step(self, skip=()):
        """One Adam step on every group except those named (or indexed) in `skip`: a skipped group is what
        torch.optim.Adam does for parameters whose .grad is None - values, moments and step count untouched."""
        self._launch(self._hyper(skip))

    def all_reduce_and_step(self, chunks=8, group=None, skip=()):
        """View-parallel training: sum the flat gradient buffer over the ranks and step, PIPELINED - the buffer is
        all-reduced in `chunks` pieces issued back to back on NCCL's stream, and the Adam kernel runs on each piece as soon
        as it has arrived, so the optimizer pass hides under the collective instead of following it."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return self.step(skip)
        total = self._begin[-1]
        per = (total // max(1, chunks) + 7) // 8 * 8
        bounds = [(lo, min(total, lo + per)) for lo in range(0, total, per)] if per > 0 else [(0, total)]
        flat = self.grads.flat
        works = [dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True) for lo, hi in bounds]
        hyper = self._hyper(skip)
        for w, (lo, hi) in zip(works, bounds):
            w.wait()                                   # the current stream waits for this piece (no host block)
            self._launch(hyper, lo, hi)

    # -- optimizer-state surgery of densification (scene/gaussian_model.py:1027-1100) on the flat buffers --
    def _group_views(self, buf):
        """Per group: list of views of `buf` shaped like the group's parameters."""
        out = []
        it = iter(view_parallel.flat_layout(self._params)[0])
        for g in self.param_groups:
            vs = []
            for p in g["params"]:
                off = next(it)
                vs.append(buf[off:off + p.numel()].view_as(p))
            out.append(vs)
        return out

    def _rebuild(self, new_p, new_m, new_v):
        """Re-home every parameter (same Python objects, possibly new shapes) into fresh flat buffers."""
        dev = self.flat_params.device
        offs, total = view_parallel.flat_layout([t for grp in new_p for t in grp])
        flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        begin, params = [], []
        it = iter(offs)
        for g, ps, ms, vs in zip(self.param_groups, new_p, new_m, new_v):
            first = True
            for p, val, m, v in zip(g["params"], ps, ms, vs):
                off = next(it)
                if first:
                    begin.append(off)
                    first = False
                n = val.numel()
                flat_p[off:off + n].copy_(val.reshape(-1))
                flat_m[off:off + n].copy_(m.reshape(-1))
                flat_v[off:off + n].copy_(v.reshape(-1))
                p.grad = None
                p.data = flat_p[off:off + n].view(val.shape)
                params.append(p)
        begin.append(total)
        self.flat_params, self.exp_avg, self.exp_avg_sq, self._begin, self._params = flat_p, flat_m, flat_v, begin, params
        self.grads = view_parallel.FlatGradBuffer(params)
        return {g.get("name", str(i)): g["params"][0] for i, g in enumerate(self.param_groups) if len(g["params"]) == 1}

    def _per_point(self, g, n):
        return len(g["params"]) == 1 and g["params"][0].dim() >= 1 and g["params"][0].shape[0] == n

    def prune(self, valid_mask):
        """`_prune_optimizer` (gaussian_model.py:1042-1063): keep rows `valid_mask` of every single-tensor group whose
        first dimension is the point count, together with their moments.  Returns {group name: parameter}."""
        n = valid_mask.numel()
        P, M, V = self._group_views(self.flat_params), self._group_views(self.exp_avg), self._group_views(self.exp_avg_sq)
        if not any(self._per_point(g, n) for g in self.param_groups):
            raise ValueError("prune: no parameter group has %d rows" % n)
        keep = lambda grp, g: [t[valid_mask] if self._per_point(g, n) else t for t in grp]
        return self._rebuild([keep(a, g) for a, g in zip(P, self.param_groups)], [keep(a, g) for a, g in zip(M, self.param_groups)],
                             [keep(a, g) for a, g in zip(V, self.param_groups)])

    def append(self, tensors_dict):
        """`cat_tensors_to_optimizer` (gaussian_model.py:1084-1107): append rows to the groups named in `tensors_dict`;
        the new rows start with zero moments.  Returns {group name: parameter}."""
        P, M, V = self._group_views(self.flat_params), self._group_views(self.exp_avg), self._group_views(self.exp_avg_sq)
        np_, nm, nv = [], [], []
        for g, ps, ms, vs in zip(self.param_groups, P, M, V):
            ext = tensors_dict.get(g.get("name")) if len(g["params"]) == 1 else None
            if ext is None:
                np_.append(ps); nm.append(ms); nv.append(vs)
            else:
                ext = ext.detach().to(ps[0].device, torch.float32)
                np_.append([torch.cat((ps[0], ext), dim=0)])
                nm.append([torch.cat((ms[0], torch.zeros_like(ext)), dim=0)])
                nv.append([torch.cat((vs[0], torch.zeros_like(ext)), dim=0)])
        return self._rebuild(np_, nm, nv)

    def replace(self, name, tensor):
        """`replace_tensor_to_optimizer` (gaussian_model.py:1027-1040): new values for one group, moments reset to zero."""
        P, M, V = self._group_views(self.flat_params), self._group_views(self.exp_avg), self._group_views(self.exp_avg_sq)
        for i, g in enumerate(self.param_groups):
            if g.get("name") == name and len(g["params"]) == 1:
                t = tensor.detach().to(P[i][0].device, torch.float32)
                P[i], M[i], V[i] = [t], [torch.zeros_like(t)], [torch.zeros_like(t)]
        return self._rebuild(P, M, V)

    # -- checkpoint surface: torch.optim.Adam's own state_dict layout (the reference saves `optimizer.state_dict()` in
    # capture() and restores it in restore(), scene/gaussian_model.py:686-728), so chkpnt*.pth files go both ways --
    def state_dict(self):
        state, packed, idx = {}, [], 0
        M, V = self._group_views(self.exp_avg), self._group_views(self.exp_avg_sq)
        for gi, g in enumerate(self.param_groups):
            ids = []
            for m, v in zip(M[gi], V[gi]):
                if self._steps[gi] > 0:         # torch creates a parameter's state at its first step
                    state[idx] = {"step": torch.tensor(float(self._steps[gi])), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
                ids.append(idx)
                idx += 1
            d = {k: v for k, v in g.items() if k != "params"}
            for k, v in (("weight_decay", 0), ("amsgrad", False), ("maximize", False), ("foreach", None), ("capturable", False),
                         ("differentiable", False), ("fused", None), ("decoupled_weight_decay", False)):
                d.setdefault(k, v)
            d["params"] = ids
            packed.append(d)
        return {"state": state, "param_groups": packed}

    def load_state_dict(self, sd):
        if "state" not in sd:                   # round-1 private layout: one global step + the flat moment buffers
            if sd["exp_avg"].numel() != self.exp_avg.numel() or sd["exp_avg_sq"].numel() != self.exp_avg_sq.numel():
                raise ValueError("load_state_dict: flat moment buffers have %d elements, this optimizer %d"
                                 % (sd["exp_avg"].numel(), self.exp_avg.numel()))
            self._steps = [int(sd["step"])] * len(self.param_groups)
            self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
            for g, s_ in zip(self.param_groups, sd["param_groups"]):
                g.update({k: v for k, v in s_.items() if k != "params"})
            return
        groups = sd["param_groups"]
        if len(groups) != len(self.param_groups):
            raise ValueError("loaded state dict has a different number of parameter groups")
        M, V = self._group_views(self.exp_avg), self._group_views(self.exp_avg_sq)
        for gi, (g, s_) in enumerate(zip(self.param_groups, groups)):
            if len(s_["params"]) != len(g["params"]):
                raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
            steps = set()
            for pid, p, m, v in zip(s_["params"], g["params"], M[gi], V[gi]):
                st = sd["state"].get(pid)
                if st is None:
                    m.zero_(); v.zero_(); steps.add(0)
                    continue
                if tuple(st["exp_avg"].shape) != tuple(p.shape) or tuple(st["exp_avg_sq"].shape) != tuple(p.shape):
                    raise ValueError("load_state_dict: moments of parameter %d have shape %s, the parameter %s"
                                     % (pid, tuple(st["exp_avg"].shape), tuple(p.shape)))
                m.copy_(st["exp_avg"]); v.copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
            if len(steps) > 1:
                raise ValueError("load_state_dict: parameters of group %r carry different step counts %s; FusedAdam keeps "
                                 "one step count per group" % (g.get("name", gi), sorted(steps)))
            self._steps[gi] = steps.pop() if steps else 0
            g.update({k: v for k, v in s_.items() if k not in ("params", "foreach", "capturable", "differentiable", "fused",
                                                               "weight_decay", "amsgrad", "maximize", "decoupled_weight_decay")})
            if s_.get("weight_decay", 0) or s_.get("amsgrad", False) or s_.get("maximize", False):
                raise ValueError("FusedAdam implements plain Adam only (no weight decay / amsgrad / maximize)")
