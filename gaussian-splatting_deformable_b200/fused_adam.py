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

    def state_dict(self):
        return {"step": self.state_step, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.state_step = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
