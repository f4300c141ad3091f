#!/usr/bin/env python
"""Soak test of the pipelined training loop: N steps over a rotating set of batches with the host running far ahead of the
device (no synchronisation inside the loop), pipelined vs single-graph replay from the same initial weights.  Prints the loss
of every 25th step for both, the device-side count of non-finite losses, and the step time."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "depth-enhancement-and-super-resolution_b200"))
import numpy as np
import torch

import bench


def main(steps=300):
    from dsr_b200 import ops
    wl = dict(bench.WORKLOADS["c2"])
    batches = [bench.make_batch(wl, s) for s in range(5)]
    for b in batches:
        for k in ("A_i", "B_i", "A_d", "B_d"):
            b[k] = b[k].pin_memory()
    out = {}
    for pipe in (True, False):
        ops.CONFIG["pipeline_frozen"] = pipe
        torch.manual_seed(0); np.random.seed(0)
        m = bench.make_model(wl, [0], True, name=f"soak{int(pipe)}")
        losses = []
        for i in range(8):
            m.set_input(batches[i % 5]); m.optimize_parameters(i, 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            m.set_input(batches[i % 5])
            m.optimize_parameters(i, 1)
            if i % 25 == 0:
                losses.append(m.loss_G.detach().clone())
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps * 1e3
        out[pipe] = ([round(float(x), 2) for x in losses], m.nonfinite_steps(), round(dt, 3))
        m.reset_graph()
        del m
        torch.cuda.empty_cache()
    for pipe, (ls, bad, dt) in out.items():
        print(f"pipelined={pipe}: {dt} ms/step (wall, host inputs), non-finite steps {bad}, loss every 25 steps: {ls}")
    a, b = out[True][0], out[False][0]
    assert out[True][1] == 0 and out[False][1] == 0
    assert abs(a[0] - b[0]) <= 1e-3 * abs(b[0]), "first compared step must agree"
    assert a[-1] < 0.8 * a[0] and b[-1] < 0.8 * b[0], "the loss must go down in both modes"
    print("soak OK")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300)
