"""View-parallel (data-parallel over cameras) training support.

The reference is single-process / single-GPU (utils/general_utils.py:133 pins cuda:0)
and has no distributed code at all.  The render of each (camera, time) is independent
given the full parameter set and gradients are additive, so the hot path shards by
VIEW: every rank keeps a full replica of the Gaussian parameters (and twists), renders
cameras {rank, rank+G, ...} of the step's batch, sums their per-Gaussian gradients
locally, and ONE all-reduce of a single flat fp32 buffer makes the replicas agree
(59 floats = 236 B per Gaussian for xyz/f_dc/f_rest/opacity/scaling/rotation,
scene/gaussian_model.py:825-831, + 7 per twist).  There is no other exchange step -
no collective on the data path - so nothing else is communicated.

One process per GPU, `torch.distributed` (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def partition_views(num_views, world_size, rank):
    """Indices of the views rank `rank` renders: a round-robin split, so that
    neighbouring (similar-cost) cameras land on different ranks."""
    return list(range(rank, num_views, world_size))


ALIGN = 8        # every tensor of a flat buffer starts on a 32-byte boundary (8 floats): the kernels move quaternion
                 # records as 128-bit and SH records as 256-bit accesses, and a packed layout would only be aligned
                 # when the point count happens to be a multiple of 4 (it is not after a prune / densify)


def flat_layout(params, align=ALIGN):
    """Start offsets (in floats) of `params` laid out back to back with `align`-float alignment, and the total
    length (itself a multiple of `align`).  Padding elements are never written: their gradients stay zero."""
    offs, off = [], 0
    for p in params:
        offs.append(off)
        off += (p.numel() + align - 1) // align * align
    return offs, off


class FlatGradBuffer:
    """Gradients of a set of leaf tensors laid out in ONE contiguous fp32 buffer.

    Each parameter's `.grad` is a view into the buffer, so autograd accumulates the
    views' gradients in place across the views a rank renders and the all-reduce runs
    on the buffer itself - no pack/unpack kernels around the collective."""

    def __init__(self, params):
        self.params = list(params)
        self.offsets, total = flat_layout(self.params)
        first = self.params[0]
        self.flat = torch.zeros(total, dtype=torch.float32, device=first.device)
        self._works = []
        self._attach()

    def _attach(self, only_if_replaced=False):
        for p, off in zip(self.params, self.offsets):
            g = self.flat[off:off + p.numel()].view_as(p)
            if only_if_replaced and p.grad is not None and p.grad.data_ptr() == g.data_ptr():
                continue
            p.grad = g

    def rebind(self, params):
        """Attach the same flat buffer to a new set of parameter tensors of identical shapes (e.g. a step's freshly
        staged copies): no allocation, no fill."""
        params = list(params)
        if [tuple(p.shape) for p in params] != [tuple(p.shape) for p in self.params]:
            raise ValueError("rebind: parameter shapes differ")
        self.params = params
        self._attach()
        return self

    def zero_(self):
        self.flat.zero_()
        self._attach(only_if_replaced=True)      # re-attach: callers may have replaced .grad

    def all_reduce(self, group=None, average=False):
        """Sum (or average) the flat buffer over all ranks.  Returns bytes reduced."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat.div_(dist.get_world_size(group))
        return self.flat.numel() * 4


    # ---- all-reduce by ranges of Gaussians, overlapped with the kernel that produces the gradients -------------------
    def all_reduce_rows(self, first, count, rows, group=None):
        """Start (asynchronously) the all-reduce of rows [first, first + count) of every per-Gaussian gradient (the tensors
        whose leading dimension is `rows`); the other tensors (e.g. the twists of a few rigid bodies) go with the range
        that ends at `rows`.  Meant as GaussianBackwardBatch's `after_chunk`: the range just enqueued on the current
        stream is reduced on the collective's stream while the next range is computed.  Call wait() before the gradients
        are used.  Returns the number of elements reduced."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return 0
        pieces = []
        for p, off in zip(self.params, self.offsets):
            if p.dim() >= 1 and p.shape[0] == rows:
                per = p.numel() // rows
                piece = self.flat[off + first * per: off + (first + count) * per]
            elif first + count == rows:
                piece = self.flat[off: off + p.numel()]
            else:
                continue
            if piece.numel():
                pieces.append(piece)
        if not pieces:
            return 0
        cm = getattr(dist, "_coalescing_manager", None) if self.flat.is_cuda else None
        if cm is not None:
            # one NCCL group (one kernel) for the range's pieces instead of one collective per tensor
            with cm(group=group, device=self.flat.device, async_ops=True) as work:
                for piece in pieces:
                    dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=group)
            self._works.append(work)
        else:
            for piece in pieces:
                self._works.append(dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=group, async_op=True))
        return sum(piece.numel() for piece in pieces)

    def wait(self):
        """Order the current stream after every all-reduce started by all_reduce_rows."""
        for w in self._works:
            w.wait()
        self._works = []


def reduce_densification_stats(xyz_gradient_accum, denom, max_radii2D, group=None):
    """Make the densification statistics identical on all ranks (SUM for the
    accumulated view-space gradient norms and their counts, MAX for the radii;
    scene/gaussian_model.py:1252-1257, train.py:613), so every replica takes the
    same clone/split/prune decisions."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    dist.all_reduce(xyz_gradient_accum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(denom, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(max_radii2D, op=dist.ReduceOp.MAX, group=group)


_streams = {}


def render_views(render_view, views, num_streams=2, batch=None):
    """Run `render_view(i)` (forward + backward of local view i, accumulating into shared gradient
    buffers) for every i in `views`, spreading the views round-robin over `num_streams` CUDA streams.
    The views of one step only meet in the gradient sums (atomic adds), so they may overlap: the
    latency-bound part of one view (depth sort passes, scans, launch gaps, kernel tails) runs under
    the issue-bound blend kernels of another.  Returns the summed loss (tensor on the current stream).
    `render_view` must allocate per-view state itself (the rasterizer does) and return a detached loss.
    `batch`: the diff_gaussian_rasterization.GaussianBackwardBatch the views were given as `accumulate_grads`, if any - its
    pending per-Gaussian backward runs here, once for all views, after the views' streams have been joined."""
    views = list(views)
    if num_streams <= 1 or len(views) <= 1 or not torch.cuda.is_available():
        total = None
        for i in views:
            loss = render_view(i)
            total = loss if total is None else total + loss
        if batch is not None:
            batch.flush()
        return total
    dev = torch.cuda.current_device()
    key = (dev, num_streams)
    if key not in _streams:
        _streams[key] = [torch.cuda.Stream(device=dev) for _ in range(num_streams)]
    side = _streams[key]
    main = torch.cuda.current_stream(dev)
    for s in side:
        s.wait_stream(main)                 # parameters / zeroed gradient buffer are ready
    losses = []
    for n, i in enumerate(views):
        with torch.cuda.stream(side[n % num_streams]):
            losses.append(render_view(i))
    for s in side:
        main.wait_stream(s)
    if batch is not None:
        batch.flush()
    total = losses[0]
    for l in losses[1:]:
        total = total + l
    for l in losses:
        l.record_stream(main)
    return total


def render_step(render_view, views, buffer, group=None):
    """One view-parallel step: `render_view(i)` must run forward+backward for local view
    i (accumulating into the parameters' .grad, i.e. into `buffer`) and return the loss.
    Returns the summed loss of the local views (a tensor)."""
    buffer.zero_()
    total = None
    for i in views:
        loss = render_view(i)
        total = loss.detach() if total is None else total + loss.detach()
    buffer.all_reduce(group=group)
    return total
