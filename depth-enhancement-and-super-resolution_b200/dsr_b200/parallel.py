"""Data parallelism for the training step: one process per GPU, replicated weights, batch sharded by
rank, and ONE collective per step - the all-reduce of the trainable-gradient arena (Depth_f + Task,
44.25 M fp32 = 177 MB) - replacing the reference's single-process ``torch.nn.DataParallel``
(models/networks.py:113-116: per-forward re-broadcast of every parameter, scatter/gather, losses on
GPU 0).  Every per-sample computation of the step is independent (InstanceNorm / GroupNorm, per-sample
K / crop / rectangles); each loss is a mean over the local shard, so averaging the gradients over
ranks reproduces the full-batch gradient when shards have equal size.

The arena is laid out in gradient-ready order (``main_model.ParamArena``), so bucket i is complete
while the backward pass is still producing buckets i+1..; ``GradBuckets`` launches each bucket's
all-reduce on a side stream as soon as its last gradient kernel has been enqueued
(``ops.GRAD_READY`` hook) and ``finish()`` makes the compute stream wait before Adam.
"""
import torch
import torch.distributed as dist

from . import ops


class GradBuckets:
    def __init__(self, model, bucket_mb=32, group=None, overlap=True):
        self.arena = model.arena
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.group = group
        self.overlap = overlap and self.arena.grad.is_cuda
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
        self.pending = [set(ps) for _, _, ps in self.buckets]
        self.launched = [False] * len(self.buckets)
        self.stream = torch.cuda.Stream() if self.overlap else None
        self.handles = []
        model.grad_sync = self
        model.optimizer_G.grad_scale = 1.0 / self.world      # all-reduce(sum), averaged inside Adam
        # bucketed overlap through the ops.GRAD_READY hook for eager steps; the graph-replayed step uses one collective
        self.overlap_hooks = self.world > 1 and not getattr(model, "use_graph", False)
        if self.overlap_hooks:
            ops.GRAD_READY = self._ready

    def _launch(self, i):
        s, e, _ = self.buckets[i]
        view = self.arena.grad[s:e]
        self.launched[i] = True
        if self.world == 1:
            return
        if self.overlap:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                self.handles.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.handles.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _ready(self, ptr):
        i = self.ptr_to_bucket.get(ptr)
        if i is None or self.launched[i]:
            return
        self.pending[i].discard(ptr)
        # buckets are launched strictly in order so every rank issues the same collective sequence
        while True:
            j = self.launched.index(False) if False in self.launched else -1
            if j < 0 or self.pending[j]:
                break
            self._launch(j)

    def all_at_once(self):
        """ONE all-reduce of the whole gradient arena on the current stream (no side stream, no hooks): what the
        CUDA-graph step records - 177 MB over NVLink 5 is ~0.5 ms, under 2 % of the step, so overlap buys little there
        and a single in-order collective is trivially capturable."""
        if self.world > 1:
            dist.all_reduce(self.arena.grad, op=dist.ReduceOp.SUM, group=self.group)

    def finish(self):
        """Launch whatever is left (parameters that got no gradient this step), wait, re-arm."""
        if not self.overlap_hooks:
            return self.all_at_once()
        for i in range(len(self.buckets)):
            if not self.launched[i]:
                self._launch(i)
        for h in self.handles:
            h.wait()
        if self.overlap and self.world > 1:
            torch.cuda.current_stream().wait_stream(self.stream)
        self.handles = []
        self.pending = [set(ps) for _, _, ps in self.buckets]
        self.launched = [False] * len(self.buckets)


def broadcast_weights(model, src=0, group=None):
    """Make every rank start from rank `src`'s weights (all five networks)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for name in model.model_names:
        for t in model._unwrap(getattr(model, "net" + name)).state_dict().values():
            dist.broadcast(t, src=src, group=group)
    ops.WEIGHT_EPOCH += 1
