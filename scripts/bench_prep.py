#!/usr/bin/env python
"""Micro-benchmark of the operand preparation (dsr_tc_prep / dsr_tc_prep_fin) on the shapes the C2 step launches
(gpurun_out/layers_step_*.json): time per launch, bytes moved, fraction of the measured HBM copy peak.  Each shape runs on a
ring of input tensors larger than L2 (cold) and on ONE tensor (L2-resident, the situation inside the step where the producing
GEMM has just written it).

  python scripts/bench_prep.py [--iters 20] [--out gpurun_out/prep.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "depth-enhancement-and-super-resolution_b200"))

import torch

# N, H, W, C, Ca, layout, pad, prm, lo, bf16 copy, csum
SHAPES = [
    (12, 64, 64, 128, 128, 0, 2, 0, 0, 0, 1),      # dY of a ResNet-block conv (bias-gradient sums ride along)
    (12, 64, 64, 128, 128, 0, 1, 0, 1, 0, 0),      # block input, reflect pad
    (12, 64, 64, 128, 128, 0, 1, 1, 1, 1, 0),      # IN + ReLU prologue, bf16 copy for the weight gradient
    (6, 64, 64, 256, 256, 0, 1, 1, 1, 0, 0),       # G_A_d block (frozen)
    (12, 256, 256, 128, 128, 0, 6, 0, 0, 0, 1),    # dY of the 7x7 head
    (12, 256, 256, 32, 64, 0, 3, 1, 1, 1, 0),      # 32-channel head input
    (12, 128, 128, 64, 64, 0, 1, 0, 0, 0, 1),
    (12, 128, 128, 64, 256, 2, 1, 1, 1, 1, 0),     # space-to-depth operand of a stride-2 conv
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from dsr_b200 import ops
    peak = 6540.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    rows = []
    for (N, H, W, C, Ca, layout, pad, has_prm, lo, bf, cs) in SHAPES:
        plan = dict(layout=layout, Cp=(C + 7) // 8 * 8, Ca=Ca)
        nbuf = max(2, int(400e6 // (N * H * W * C * 4)) + 1)
        xs = [torch.randn(N, H, W, C, device="cuda") for _ in range(nbuf)]
        sums = torch.zeros(N * C * 2, dtype=torch.float64, device="cuda")
        ops._call("dsr_channel_sums", ops._p(xs[0]), N, H * W, C, ops._p(sums, torch.float64))
        csum_buf = torch.zeros(C * 8, dtype=torch.float64, device="cuda")
        for fold in ((False, True) if has_prm else ((1, 8) if cs else (False,))):
            for resident in (False, True):
                if cs:
                    ops.CONFIG['csum_reps'] = fold                      # A/B of the replicated bias-gradient accumulator
                else:
                    ops.CONFIG["fold_finalize"] = fold
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

                def one(i):
                    x = xs[0] if resident else xs[i % nbuf]
                    prm = ops._norm_params(x, 0, None, None, 1e-5, sums, lazy=True) if has_prm else None
                    csum = csum_buf if cs else None            # never re-zeroed: only the time matters here
                    return ops._tc_prep(x, plan, pad, ops.PAD_ZERO if cs else ops.PAD_REFLECT, prm, ops.ACT_RELU if has_prm else ops.ACT_NONE,
                                        need_lo=bool(lo), also_bf16=bool(bf), csum=csum, dtype="f16" if lo else "bf16")
                for i in range(3):
                    r = one(i)
                torch.cuda.synchronize()
                # the launches are replayed from a CUDA graph: 15 us kernels launched one by one from Python time the host
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(a.iters):
                        r = one(i)
                g.replay()
                torch.cuda.synchronize()
                ev[0].record()
                g.replay()
                ev[1].record()
                torch.cuda.synchronize()
                us = ev[0].elapsed_time(ev[1]) * 1e3 / a.iters
                out_b = r[0].numel() * 2 * (1 + bool(lo) + bool(bf))
                byts = N * H * W * C * 4 + out_b
                rows.append(dict(shape=[N, H, W, C, Ca, layout, pad, has_prm, lo, bf, cs], fold_or_reps=fold, resident=resident, us=round(us, 2),
                                 mb=round(byts / 1e6, 1), gbps=round(byts / us / 1e3, 0), frac=round(byts / us / 1e3 / peak, 3)))
                print(rows[-1], flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
