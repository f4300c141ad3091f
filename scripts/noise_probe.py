#!/usr/bin/env python
"""Run-to-run variability of the gradient arena: the same step (same weights, same batch, same rectangle draws) executed
several times eagerly; prints the flat cosine and the worst per-tensor cosine between runs.  Split-K / statistics atomics
sum in a different order every run; the nets amplify that ~1e4-fold.  GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "depth-enhancement-and-super-resolution_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from oracle import ref_step
from util import build_host_model, cosine, rehome


def main():
    from dsr_b200 import ops
    host = build_host_model(2, 128, 128)
    batch = ref_step.synthetic_batch(2, 128, 128, seed=1, depth_kind="smooth")
    m = rehome(host, host.opt, [0])
    m._train()
    np.random.seed(7)
    for it in range(3):
        m.set_input(batch); m.optimize_parameters(it, 1)
    state = (m.arena.flat, m.arena.exp_avg, m.arena.exp_avg_sq, m.optimizer_G.step_dev)
    snap = [t.clone() for t in state]
    grads = []
    for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
        for t, s0 in zip(state, snap):
            t.copy_(s0)
        ops.WEIGHT_EPOCH += 1
        np.random.seed(11)
        m.set_input(batch); m.optimize_parameters(9, 1)
        grads.append(m.arena.grad.clone())
    names = {}
    for net in ("Depth_f", "Task"):
        for n, p in m._unwrap(getattr(m, "net" + net)).named_parameters():
            names[p.data_ptr()] = f"{net}.{n}"
    for r in range(1, len(grads)):
        worst, wn = 1.0, ""
        for p, o in zip(m.arena.params, m.arena.offsets):
            if p.dim() == 4:
                c = cosine(grads[0][o:o + p.numel()].cpu(), grads[r][o:o + p.numel()].cpu())
                if c < worst:
                    worst, wn = c, names[p.data_ptr()]
        print(f"run 0 vs {r}: flat cos {cosine(grads[0].cpu(), grads[r].cpu()):.8f}  worst weight tensor {worst:.6f} ({wn})", flush=True)


if __name__ == "__main__":
    main()
