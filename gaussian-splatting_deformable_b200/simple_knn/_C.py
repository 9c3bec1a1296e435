"""Drop-in for `simple_knn._C` (submodules/simple-knn/ext.cpp:15-17): `distCUDA2`.

`from simple_knn._C import distCUDA2` keeps working (scene/gaussian_model.py:21,818).
"""
import torch

import gsr_runtime as _rt


def distCUDA2(points):
    """Mean squared distance to the 3 nearest other points (spatial.cu:15-26).

    points: float tensor [P,3] on CUDA -> float32 tensor [P]."""
    if not points.is_cuda:
        raise _rt.GsrError("distCUDA2 runs on CUDA tensors only (no CPU fallback)")
    lib = _rt.load()
    P = int(points.shape[0])
    pts = points.detach().float().contiguous()
    out = torch.zeros((P,), dtype=torch.float32, device=points.device)
    if P == 0:
        return out
    with torch.cuda.device(points.device):
        nbytes = lib.gsr_knn_bytes(P)
        temp = torch.empty(nbytes, dtype=torch.uint8, device=points.device)
        _rt.check(lib.gsr_knn_dist2(P, _rt.ptr(pts), _rt.ptr(out), _rt.ptr(temp), nbytes,
                                    _rt.stream_ptr(points.device)))
    return out
