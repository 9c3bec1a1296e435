"""Fused Adam over one flat parameter buffer (csrc/adam.cu) with the reference's optimizer surface.

The reference builds `torch.optim.Adam(l, lr=0.0, eps=1e-15)` from a list of per-tensor groups
`{'params': [tensor], 'lr': ..., 'name': ...}` (scene/gaussian_model.py:834-846) and drives the learning rates by
writing `param_group['lr']` (gaussian_model.py:869-886).  `FusedAdam` takes the same list, re-homes every
parameter as a view of ONE flat fp32 buffer (gradients: `view_parallel.FlatGradBuffer`, which is also the
all-reduce buffer; moments: two more flat buffers) and runs the whole step as one kernel.
Arithmetic is torch.optim.Adam's (torch 2.11; no amsgrad, weight decay or maximize).
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
        total = sum(p.numel() for p in flat_params)
        pad = (-total) % 4
        self.flat_params = torch.empty(total + pad, dtype=torch.float32, device=dev)
        off = 0
        self._begin = []
        for g in self.param_groups:
            self._begin.append(off)
            for p in g["params"]:
                n = p.numel()
                self.flat_params[off:off + n].copy_(p.detach().reshape(-1))
                p.data = self.flat_params[off:off + n].view_as(p)
                off += n
        self._begin.append(off)
        self.grads = view_parallel.FlatGradBuffer(flat_params)       # p.grad = views of grads.flat
        self.exp_avg = torch.zeros_like(self.grads.flat)
        self.exp_avg_sq = torch.zeros_like(self.grads.flat)
        self.state_step = 0

    # -- torch.optim surface the reference uses --
    def zero_grad(self, set_to_none=True):
        # the reference calls zero_grad(set_to_none=True) (train.py:683); gradients live in the flat buffer, so
        # they are zeroed in place and stay attached (accumulate_grads / the all-reduce use the same memory)
        self.grads.zero_()

    def step(self):
        lib = _rt.load()
        self.state_step += 1
        t = self.state_step
        n = len(self.param_groups)
        begin = (ctypes.c_ulonglong * (n + 1))(*self._begin)
        ss, bc, b1s, b2s, es = [], [], [], [], []
        for g in self.param_groups:
            b1, b2 = g["betas"]
            bias_correction1 = 1 - b1 ** t
            bias_correction2 = 1 - b2 ** t
            ss.append(g["lr"] / bias_correction1)
            bc.append(bias_correction2 ** 0.5)
            b1s.append(b1); b2s.append(b2); es.append(g["eps"])
        arr = lambda v: (ctypes.c_float * n)(*v)
        dev = self.flat_params.device
        with torch.cuda.device(dev):
            darr = lambda v: (ctypes.c_double * n)(*v)
            _rt.check(lib.gsr_adam_step(self.flat_params.data_ptr(), self.grads.flat.data_ptr(), self.exp_avg.data_ptr(),
                                        self.exp_avg_sq.data_ptr(), n, begin, arr(ss), arr(bc), darr(b1s), darr(b2s), arr(es),
                                        _rt.stream_ptr(dev)))

    # -- optimizer-state surgery of densification (scene/gaussian_model.py:1027-1100) on the flat buffers --
    def _group_views(self, buf):
        """Per group: list of views of `buf` shaped like the group's parameters."""
        out, off = [], 0
        for g in self.param_groups:
            vs = []
            for p in g["params"]:
                n = p.numel()
                vs.append(buf[off:off + n].view_as(p))
                off += n
            out.append(vs)
        return out

    def _rebuild(self, new_p, new_m, new_v):
        """Re-home every parameter (same Python objects, possibly new shapes) into fresh flat buffers."""
        dev = self.flat_params.device
        total = sum(t.numel() for grp in new_p for t in grp)
        flat_p = torch.empty(total + ((-total) % 4), dtype=torch.float32, device=dev)
        flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        off, begin, params = 0, [], []
        for g, ps, ms, vs in zip(self.param_groups, new_p, new_m, new_v):
            begin.append(off)
            for p, val, m, v in zip(g["params"], ps, ms, vs):
                n = val.numel()
                flat_p[off:off + n].copy_(val.reshape(-1))
                flat_m[off:off + n].copy_(m.reshape(-1))
                flat_v[off:off + n].copy_(v.reshape(-1))
                p.grad = None
                p.data = flat_p[off:off + n].view(val.shape)
                params.append(p)
                off += n
        begin.append(off)
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

    def state_dict(self):
        return {"step": self.state_step, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.state_step = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
