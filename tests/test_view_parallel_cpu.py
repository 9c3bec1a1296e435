"""World-size-2 (gloo, CPU) test of the view-parallel path: each rank renders its share
of the cameras (with the CPU oracle standing in for the GPU kernels), gradients are
summed through ONE flat all-reduce, and the result must equal the single-process sum
over all cameras."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synthetic
import view_parallel as vp
from oracle import oracle_c

P, W, H, NVIEWS = 300, 48, 32, 4
KEYS = ("means3D", "scales", "rotations", "opacities", "shs")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _view_grads(sc, k):
    cam = synthetic.make_camera(k, NVIEWS, W, H)
    rs = synthetic.raster_settings(cam, torch.tensor([0.1, 0.0, 0.2]))
    f = oracle_c.forward(rs, sc["means3D"], sc["opacities"], shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"])
    b = oracle_c.backward(f, synthetic.make_image_grad(W, H, seed=1))
    loss = float((f["color"] * synthetic.make_image_grad(W, H, seed=1).numpy()).sum())
    return {k2: torch.from_numpy(np.ascontiguousarray(b[k2])) for k2 in KEYS}, loss, f["radii"]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = synthetic.make_scene(P, seed=0, scale_mult=3.0)
    params = [sc[k].clone().requires_grad_(True) for k in KEYS]
    buf = vp.FlatGradBuffer(params)
    views = vp.partition_views(NVIEWS, world, rank)
    stats = dict(accum=torch.zeros(P, 1), denom=torch.zeros(P, 1), radii=torch.zeros(P))

    def render_view(i):
        grads, loss, radii = _view_grads(sc, i)
        for p, k in zip(params, KEYS):
            p.grad += grads[k].view_as(p)                 # what autograd's AccumulateGrad does
        vis = torch.from_numpy(radii > 0)
        stats["accum"][vis] += 1.0
        stats["denom"][vis] += 1.0
        stats["radii"] = torch.maximum(stats["radii"], torch.from_numpy(radii).float())
        return torch.tensor(loss)

    total = vp.render_step(render_view, views, buf)
    dist.all_reduce(total)
    vp.reduce_densification_stats(stats["accum"], stats["denom"], stats["radii"])
    # every .grad is still a view into the flat buffer (no pack/unpack)
    assert all(p.grad.data_ptr() >= buf.flat.data_ptr() for p in params)
    torch.save(dict(flat=buf.flat.clone(), loss=total, views=views, stats=stats), os.path.join(out_dir, "r%d.pt" % rank))
    dist.destroy_process_group()


def test_partition_covers_all_views_once():
    for n, w in ((64, 8), (7, 2), (3, 4), (0, 2)):
        parts = [vp.partition_views(n, w, r) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_rank_allreduce_equals_single_process_sum(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(tmp_path, "r0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "r1.pt"))
    assert r0["views"] == [0, 2] and r1["views"] == [1, 3]
    # replicas agree bit for bit after the collective
    assert torch.equal(r0["flat"], r1["flat"]) and torch.equal(r0["loss"], r1["loss"])
    assert all(torch.equal(r0["stats"][k], r1["stats"][k]) for k in r0["stats"])
    # and equal the single-process accumulation over all four views
    sc = synthetic.make_scene(P, seed=0, scale_mult=3.0)
    ref = {k: torch.zeros_like(sc[k]) for k in KEYS}
    loss = 0.0
    seen = torch.zeros(P)
    for k in range(NVIEWS):
        g, l, radii = _view_grads(sc, k)
        for kk in KEYS:
            ref[kk] += g[kk].view_as(ref[kk])
        loss += l
        seen += torch.from_numpy(radii > 0).float()
    flat_ref = torch.cat([ref[k].reshape(-1) for k in KEYS])
    assert flat_ref.numel() == P * 59                    # 59 floats (236 B) per Gaussian
    # the flat buffer starts every tensor on a 32-byte boundary (view_parallel.flat_layout); padding stays zero
    import view_parallel as vp
    offs, total = vp.flat_layout([ref[k] for k in KEYS])
    assert r0["flat"].numel() == total
    got = torch.cat([r0["flat"][o:o + ref[k].numel()] for o, k in zip(offs, KEYS)])
    pad = r0["flat"].clone()
    for o, k in zip(offs, KEYS):
        pad[o:o + ref[k].numel()] = 0
    assert float(pad.abs().max()) == 0.0
    torch.testing.assert_close(got, flat_ref, rtol=1e-5, atol=1e-5 * float(flat_ref.abs().max()))
    assert abs(float(r0["loss"]) - loss) <= 1e-4 * abs(loss)
    assert torch.equal(r0["stats"]["denom"].reshape(-1), seen)


def _rows_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    rows = 1003                                                     # not a multiple of anything
    shapes = [(rows, 3), (rows, 1), (rows, 16, 3), (rows, 4), (5, 6), (5,)]     # per-Gaussian tensors + a few body twists
    params = [torch.zeros(s, requires_grad=True) for s in shapes]
    a, b = vp.FlatGradBuffer(params), None
    a.flat.copy_(torch.randn(a.flat.numel(), generator=g))
    whole = a.flat.clone()
    # by ranges of rows, asynchronously (what GaussianBackwardBatch's after_chunk does), against one all-reduce of everything
    n = 0
    for first in range(0, rows, 256):
        n += a.all_reduce_rows(first, min(256, rows - first), rows)
    a.wait()
    dist.all_reduce(whole)
    offs, _ = vp.flat_layout(params)
    used = torch.zeros_like(whole, dtype=torch.bool)
    for p, o in zip(params, offs):
        used[o:o + p.numel()] = True
    assert n == int(used.sum())                                     # every element exactly once, padding never sent
    torch.save(dict(rows=a.flat.clone(), whole=whole, used=used), os.path.join(out_dir, "rows%d.pt" % rank))
    dist.destroy_process_group()


def test_all_reduce_by_row_ranges_equals_one_all_reduce(tmp_path):
    world = 2
    mp.spawn(_rows_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(tmp_path, "rows0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rows1.pt"))
    u = r0["used"]
    assert torch.equal(r0["rows"][u], r0["whole"][u]) and torch.equal(r1["rows"][u], r1["whole"][u])
    assert torch.equal(r0["rows"][u], r1["rows"][u])
