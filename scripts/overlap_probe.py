#!/usr/bin/env python
"""Upper bound of what more concurrency can buy on one GPU: TWO independent MainModel instances replay their captured steps on
two streams at the same time; aggregate pair-samples/s against one instance alone.  (Decides whether cross-step pipelining of
the frozen networks is worth building.)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "depth-enhancement-and-super-resolution_b200"))
import torch

import bench


def main():
    import numpy as np
    wl = bench.WORKLOADS["c2"]
    torch.manual_seed(0); np.random.seed(0)
    models = [bench.make_model(wl, [0], True, name=f"probe{i}") for i in range(2)]
    batches = [{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in bench.make_batch(wl, s).items()} for s in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]

    def run(n_models, steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(steps):
            for m in range(n_models):
                with torch.cuda.stream(streams[m]):
                    models[m].set_input(batches[(i + m) % 2])
                    models[m].optimize_parameters(i, 1)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / steps * 1e3

    run(2, 6)                     # warm-up + graph capture of both
    one = run(1, 20)
    two = run(2, 20)
    print(f"one model: {one:.2f} ms/step; two models concurrently: {two:.2f} ms per pair of steps = {two / 2:.2f} ms/step "
          f"({100 * (one / (two / 2) - 1):.1f} % more throughput)")


if __name__ == "__main__":
    main()
