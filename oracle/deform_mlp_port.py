"""TEST INFRASTRUCTURE ONLY (oracle): torch restatement of the reference's deformation network, the caller immediately
upstream of the hot path (SURVEY.md 8f row f1 - NOT built natively yet; this oracle and its golden vectors are the
first step of that row).

* `embed`          - positional embedding of scene/gaussian_model.py:32-81 (`get_embedder(10, d, 0)`: the input, then
                     sin and cos of x * 2^k for k = 0..9, concatenated in that order).
* `DeformMLP`      - `DirectTemporalNeRF` (gaussian_model.py:242-316): 8 x 256 ReLU trunk on [embed(x) (63), embed(t) (21)],
                     the embedded point re-concatenated IN FRONT of the activations after layer 4, four linear heads
                     (dx 3, dx_scale 3, dx_rot 4, mlp_shs 48); zeros for iteration < 3000.
* `screw_from_raw` - how `DirectTemporalNeRF_se3.forward` (gaussian_model.py:153-173) turns raw (w, v) head outputs into
                     the (S, theta) the fused SE3 path consumes: theta = |w|, S = (w, v) / theta (no epsilon guard).
Pinned: tests/golden/mlp_golden.pt holds outputs of the REAL classes (their source executed from
/root/reference by tests/golden/make_mlp_golden.py); tests/test_oracle_cpu.py holds this port to them bit for bit.
Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def embed(x, num_freqs=10):
    """[N, d] -> [N, d * (1 + 2 * num_freqs)]: x, sin(x f0), cos(x f0), sin(x f1), ...; f_k = 2^k (log sampling)."""
    freqs = 2.0 ** torch.linspace(0.0, num_freqs - 1, steps=num_freqs)
    parts = [x]
    for f in freqs:
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, -1)


class DeformMLP(nn.Module):
    """Parameter names match the reference's state_dict (`_time.N.weight`, `_time_out.weight`, ...), so a checkpoint
    of `offset_model.pth` (gaussian_model.py:925) loads directly."""

    def __init__(self, D=8, W=256, input_ch=3, skips=(4,)):
        super().__init__()
        self.in_pts = input_ch * 21
        self.in_time = 21
        self.skips = tuple(skips)
        layers = [nn.Linear(self.in_pts + self.in_time, W)]
        for i in range(D - 1):
            layers.append(nn.Linear(W + (self.in_pts if i in self.skips else 0), W))
        self._time = nn.ModuleList(layers)
        self._time_out = nn.Linear(W, 3)
        self._time_out_scale = nn.Linear(W, 3)
        self._time_out_rot = nn.Linear(W, 4)
        self._time_out_shs = nn.Linear(W, 48)

    def forward(self, x, ts, iteration):
        pts = embed(x)
        t = embed(ts)
        h = torch.cat([pts, t], dim=-1)
        for i, layer in enumerate(self._time):
            h = F.relu(layer(h))
            if i in self.skips:
                h = torch.cat([pts, h], -1)
        dx, ds, dr, dsh = self._time_out(h), self._time_out_scale(h), self._time_out_rot(h), self._time_out_shs(h)
        if iteration < 3000:                                  # gaussian_model.py:305-310
            n = pts.shape[0]
            return (torch.zeros_like(pts[:, :3]), torch.zeros_like(pts[:, :3]), torch.zeros(n, 4).to(pts.device),
                    torch.zeros(n, 48).to(pts.device))
        return dx, ds, dr, dsh


def screw_from_raw(w_raw, v_raw):
    """gaussian_model.py:161-164."""
    theta = torch.norm(w_raw, dim=-1)
    return torch.cat([w_raw / theta[..., None], v_raw / theta[..., None]], dim=-1), theta
