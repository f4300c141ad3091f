"""Data parallelism for the training step: one process per GPU, replicated weights, batch sharded by
rank, and ONE exchange per step - the all-reduce of the trainable-gradient arena (Depth_f + Task,
44.25 M fp32 = 177 MB) - replacing the reference's single-process ``torch.nn.DataParallel``
(models/networks.py:113-116: per-forward re-broadcast of every parameter, scatter/gather, losses on
GPU 0).  Every per-sample computation of the step is independent (InstanceNorm / GroupNorm, per-sample
K / crop / rectangles).  Each *mean* loss term is a mean over the local shard, so averaging the gradients
over ranks reproduces the full-batch gradient when shards have equal size; the
four total-variation terms are *sums* over the batch (main_model.py:15-19), so a rank weighs them by the
world size (``model.tv_scale``) before the average.

The arena is laid out in gradient-ready order (``main_model.ParamArena``), so bucket i is complete
while the backward pass is still producing buckets i+1..; ``GradBuckets`` issues each bucket's
all-reduce as soon as its last gradient kernel has been enqueued (``ops.GRAD_READY`` hook) on a
dedicated communication stream behind an event of the gradient-writing streams, so it overlaps the
rest of the backward pass; ``finish()`` joins before Adam.
The same sequence is legal under CUDA-graph capture (an event fork / join inside the captured step),
so the graph-replayed step - the one bench.py times - overlaps exactly like the eager one.
"""
import torch
import torch.distributed as dist

from . import ops


class GradBuckets:
    def __init__(self, model, bucket_mb=32, group=None, overlap=True):
        self.arena = model.arena
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.overlap = overlap
        # contiguous buckets over the arena, cut at parameter boundaries
        limit = int(bucket_mb * (1 << 20) // 4)
        self.buckets = []          # (start, end, [param data_ptrs])
        start, ptrs = 0, []
        for p, o in zip(self.arena.params, self.arena.offsets):
            end = o + (p.numel() + 3) // 4 * 4
            ptrs.append(p.data_ptr())
            if end - start >= limit:
                self.buckets.append((start, end, ptrs))
                start, ptrs = end, []
        if ptrs:
            self.buckets.append((start, self.arena.total, ptrs))
        self.ptr_to_bucket = {q: i for i, (_, _, ps) in enumerate(self.buckets) for q in ps}
        self.comm, self.comm_used = None, False
        self._rearm()
        model.grad_sync = self
        model.optimizer_G.grad_scale = 1.0 / self.world      # all-reduce(sum), averaged inside Adam
        model.tv_scale = float(self.world)                   # batch-SUM loss terms (see the module docstring)
        self.overlap_hooks = self.world > 1 and self.overlap
        if self.overlap_hooks:
            ops.GRAD_READY = self._ready

    def _rearm(self):
        self.pending = [set(ps) for _, _, ps in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.next = 0

    def _launch(self, i):
        s, e, _ = self.buckets[i]
        self.launched[i] = True
        if self.world == 1:
            return
        view = self.arena.grad[s:e]
        if not view.is_cuda:                       # gloo / CPU arena (tests): nothing to overlap with
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
            return
        # The collective is enqueued on a dedicated communication stream behind everything that has written gradients so
        # far (the weight-gradient stream of ops._on_side and the compute stream); the compute stream does not wait for it
        # until finish().  An event fork / join, so the same sequence is legal inside a CUDA-graph capture.
        if self.comm is None:
            self.comm = torch.cuda.Stream()
        ops.order_after_gradient_writers()
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        self.comm_used = True

    def _ready(self, ptr):
        i = self.ptr_to_bucket.get(ptr)
        if i is None or self.launched[i]:
            return
        self.pending[i].discard(ptr)
        # buckets are launched strictly in order so every rank issues the same collective sequence
        while self.next < len(self.buckets) and not self.pending[self.next]:
            self._launch(self.next)
            self.next += 1

    def all_at_once(self):
        """ONE in-order all-reduce of the whole gradient arena on the current stream (overlap=False)"""
        if self.world > 1:
            dist.all_reduce(self.arena.grad, op=dist.ReduceOp.SUM, group=self.group)

    def finish(self):
        """Launch whatever is left (parameters that got no gradient this step), join, re-arm."""
        ops.join_side()
        if not self.overlap_hooks:
            return self.all_at_once()
        for i in range(self.next, len(self.buckets)):
            if not self.launched[i]:
                self._launch(i)
        if self.comm_used:
            torch.cuda.current_stream().wait_stream(self.comm)      # an event join under capture
            self.comm_used = False
        self._rearm()

    def detach(self, model):
        """back to single-process semantics (used by the data-parallel self-check)"""
        model.grad_sync = None
        model.optimizer_G.grad_scale = 1.0
        model.tv_scale = 1.0
        if ops.GRAD_READY == self._ready:
            ops.GRAD_READY = None

    def attach(self, model):
        model.grad_sync = self
        model.optimizer_G.grad_scale = 1.0 / self.world
        model.tv_scale = float(self.world)
        if self.overlap_hooks:
            ops.GRAD_READY = self._ready
        self._rearm()


def broadcast_weights(model, src=0, group=None):
    """Make every rank start from rank `src`'s weights (all five networks)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for name in model.model_names:
        for t in model._unwrap(getattr(model, "net" + name)).state_dict().values():
            dist.broadcast(t, src=src, group=group)
    ops.WEIGHT_EPOCH += 1


def weights_in_sync(model, group=None):
    """max |w_rank - w_rank0| over the trainable arena, maximised over ranks (0.0 = bit-identical replicas)"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0.0
    ref = model.arena.flat.detach().clone()
    dist.broadcast(ref, src=0, group=group)
    d = (ref - model.arena.flat.detach()).abs().max().reshape(1)
    dist.all_reduce(d, op=dist.ReduceOp.MAX, group=group)
    return float(d)


def shutdown(models=()):
    """Orderly exit of a data-parallel run: drop the captured graphs (they hold the communicator's collectives), drain
    the device, then destroy the process group."""
    for m in models:
        if hasattr(m, "reset_graph"):
            m.reset_graph()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dist.destroy_process_group()


def _cat_batches(batches):
    out = {}
    for k in batches[0]:
        vs = [b[k] for b in batches]
        out[k] = torch.cat(vs, 0) if torch.is_tensor(vs[0]) else sum((list(v) for v in vs), [])
    return out


def dp_self_check(model, sync, batches, rects, group=None):
    """Numerical check of the data-parallel step: the rank-averaged gradient arena of one
    sharded step against the gradient of ONE process on the concatenated batch, plus replica consistency of the weights.

    batches[r] / rects[r] = rank r's host batch (set_input dict) and rectangle tables (rects_real, counts_real, rects_syn,
    counts_syn) - every rank holds all of them (they are generated from per-rank seeds).  Call before the step is captured
    into a CUDA graph (the check changes the batch shape).  No optimizer update is made.
    -> dict(cos, rel_l2, weights_max_diff), identical on every rank."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = len(batches)

    def grads(batch, rect):
        model.rect_override = rect
        model.set_input(batch)
        ops.zero_pool_reset(model.device)
        model.forward()
        model.optimizer_G.zero_grad()
        model.backward_G()
        if model.grad_sync is not None:
            model.grad_sync.finish()
        torch.cuda.synchronize()
        return model.arena.grad.detach().clone() * model.optimizer_G.grad_scale, float(model.loss_G)

    # Everything runs on the stream the training step will later be captured on: the autograd engine remembers the stream each
    # node (incl. the parameters' gradient accumulators) was first built on and orders later backward passes against it -
    # a first pass on another stream would plant a dependency on un-captured work into the capture.
    if getattr(model, "_gstream", None) is None:
        model._gstream = torch.cuda.Stream()
    outer = torch.cuda.current_stream()
    model._gstream.wait_stream(outer)
    try:
        with torch.cuda.stream(model._gstream):
            g_dp, loss_dp = grads(batches[rank], rects[rank])
            sync.detach(model)
            import numpy as np
            full_rect = tuple(np.concatenate([r[i] for r in rects], 0) for i in range(4))
            g_full, loss_full = grads(_cat_batches(batches), full_rect)
        outer.wait_stream(model._gstream)
    finally:
        sync.attach(model)
        model.rect_override = None
        model._in = None
        model._rect = None
    a, b = g_dp.double(), g_full.double()
    cos = float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))
    rel = float((a - b).norm() / b.norm().clamp_min(1e-300))
    stats = torch.tensor([cos, -cos, rel, loss_dp], device=model.device, dtype=torch.float64)
    if dist.is_initialized() and world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM, group=group)
        cos_min, rel_max, loss_mean = -float(mx[1]), float(mx[2]), float(sm[3]) / world
    else:
        cos_min, rel_max, loss_mean = cos, rel, loss_dp
    return dict(cos=cos_min, rel_l2=rel_max, loss_rank_mean=loss_mean, loss_full_batch=loss_full,
                weights_max_diff=weights_in_sync(model, group), world=world)
