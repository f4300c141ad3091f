"""Data parallelism, checked numerically on the CPU with a world_size-2 gloo group (SURVEY.md section 4(iii):
"N-GPU result vs 1-GPU result on the concatenated batch"; replaces models/networks.py:113-116 DataParallel).

Each rank computes the oracle's gradients on ITS shard of a batch (one sample per rank, its own rows of the rectangle
tables, the batch-sum TV terms weighted by the world size as ``MainModel.tv_scale`` does), writes them into the product's
``ParamArena`` in gradient-ready order through the ``ops.GRAD_READY`` hook, and ``parallel.GradBuckets`` all-reduces the
buckets; the averaged arena must equal the oracle's gradient on the concatenated batch, and one Adam update from it
must leave bit-identical weights on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, bucket_mb):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    root = os.path.dirname(HERE)
    for p in (root, os.path.join(root, "depth-enhancement-and-super-resolution_b200"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from types import SimpleNamespace
    from dsr_b200 import main_model, ops, parallel
    from oracle import ref_ops, ref_step
    from util import build_host_model, cosine, state_dicts

    H = W = 128
    host = build_host_model(world, H, W)                       # same seed on every rank: replicated weights
    sds = state_dicts(host)
    full = ref_step.synthetic_batch(world, H, W, seed=3, depth_kind="smooth")
    np.random.seed(5)
    state = np.random.get_state()
    rects_real = ref_ops.draw_rects(world, H, W, "train")     # the single-process draw order: real loop, then syn loop
    rects_syn = ref_ops.draw_rects(world, H, W, "train")

    # --- single process on the concatenated batch (the reference's DataParallel semantics)
    np.random.set_state(state)
    ref_full = ref_step.OracleStep(sds, lr=1e-4).step(full, update=False)

    # --- this rank's shard: its sample, its rows of the rectangle tables, TV (a batch SUM) weighted by the world size
    shard = {k: (v[rank:rank + 1] if torch.is_tensor(v) else v[rank:rank + 1]) for k, v in full.items()}
    queue = [rects_real[rank:rank + 1], rects_syn[rank:rank + 1]]
    orig_draw, orig_tv = ref_ops.draw_rects, ref_ops.tv_loss
    ref_ops.draw_rects = lambda *a, **k: queue.pop(0)
    ref_ops.tv_loss = lambda x: orig_tv(x) * float(world)
    try:
        orc = ref_step.OracleStep(sds, lr=1e-4)
        ref_shard = orc.step(shard, update=False)
    finally:
        ref_ops.draw_rects, ref_ops.tv_loss = orig_draw, orig_tv

    # --- the product's arena + bucketed all-reduce, driven in gradient-ready order through the hook
    arena = main_model.ParamArena([host.netDepth_f, host.netTask], torch.device("cpu"))
    shim = SimpleNamespace(arena=arena, optimizer_G=SimpleNamespace(grad_scale=1.0), grad_sync=None, tv_scale=1.0)
    sync = parallel.GradBuckets(shim, bucket_mb=bucket_mb)
    assert shim.grad_sync is sync and shim.optimizer_G.grad_scale == 1.0 / world and shim.tv_scale == float(world)
    assert len(sync.buckets) >= (2 if bucket_mb < 100 else 1)
    assert sync.buckets[0][0] == 0 and sync.buckets[-1][1] == arena.total
    assert all(a[1] == b[0] for a, b in zip(sync.buckets, sync.buckets[1:]))
    names = {}
    for net in ("Depth_f", "Task"):
        for n, p in getattr(host, "net" + net).named_parameters():
            names[p.data_ptr()] = (net, n)
    arena.zero_grad()
    for p in arena.params:                                     # = the order the backward pass produces them in
        ops.DIRECT_GRADS[p.data_ptr()].add_(ref_shard["grads"][names[p.data_ptr()]])
        ops._grad_ready(p)
    launched_early = sum(sync.launched)
    sync.finish()
    g_dp = arena.grad * shim.optimizer_G.grad_scale

    flat_full = torch.zeros_like(arena.grad)
    for p, o in zip(arena.params, arena.offsets):
        flat_full[o:o + p.numel()] = ref_full["grads"][names[p.data_ptr()]].flatten()
    cos = cosine(g_dp, flat_full)
    rel = float((g_dp - flat_full).norm() / flat_full.norm())

    # --- one Adam update from the averaged gradient: replicas stay bit-identical
    m, v = torch.zeros_like(arena.flat), torch.zeros_like(arena.flat)
    with torch.no_grad():
        ref_ops.adam_update(arena.flat, g_dp, m, v, 1, 1e-4)
    mine = arena.flat.detach().clone()
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    same = bool(torch.equal(ref0, mine))
    in_sync = parallel.weights_in_sync(shim)
    arena.release()
    assert not any(p.data_ptr() in ops.DIRECT_GRADS for p in arena.params)
    torch.save(dict(cos=cos, rel=rel, same=same, in_sync=in_sync, launched_early=launched_early, buckets=len(sync.buckets),
                    loss_full=ref_full["losses"]["G"], loss_shard=ref_shard["losses"]["G"]),
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    parallel.shutdown()


@pytest.mark.parametrize("bucket_mb", [32, 1000])
def test_dp_gradients_equal_concatenated_batch_gloo(tmp_path, bucket_mb):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), bucket_mb), nprocs=world, join=True)
    res = [torch.load(str(tmp_path / f"r{r}.pt")) for r in range(world)]
    for r in res:
        assert r["cos"] >= 0.99999 and r["rel"] <= 2e-3, r
        assert r["same"] and r["in_sync"] == 0.0, r
    if bucket_mb == 32:
        # buckets complete in gradient-ready order: all but the tail are on the wire before the backward pass has ended
        assert res[0]["buckets"] >= 5 and res[0]["launched_early"] == res[0]["buckets"], res[0]
    # the rank-mean of the shard losses is the full-batch loss (the TV terms, weighted by the world size, included)
    mean_shard = sum(r["loss_shard"] for r in res) / world
    assert abs(mean_shard - res[0]["loss_full"]) <= 1e-4 * abs(res[0]["loss_full"]), (mean_shard, res[0]["loss_full"])
