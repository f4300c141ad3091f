#!/usr/bin/env python
"""The grouped data gradient of the 7x7 32 -> 128 head (ops._tc_dgrad_group) alone: GEMM time for g = 2 (pairs, conv_tc2)
and g = 4 (quads, conv_tc3<1>), with the kernels' wait counters (DSR_BENCH_WAITS=1).
  python scripts/bench_dgrad_quad.py [--iters 10] [--g 4]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "depth-enhancement-and-super-resolution_b200"))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--g", default="2,4")
    a = ap.parse_args()
    from dsr_b200 import _lib, ops
    N, H, W, Co, Ci = 12, 256, 256, 128, 32
    dy = torch.randn(N, H, W, Co, device="cuda")
    w = (torch.randn(Co, Ci, 7, 7, device="cuda") * 0.05)
    for g in [int(x) for x in a.g.split(",")]:
        gP = ops._Prepared(dy)
        calls, orig = [], _lib.call
        _lib.call = lambda nm, *aa: (calls.append((nm, aa)), orig(nm, *aa))[1]
        y, Wc = ops._tc_dgrad_group(gP, w, 6, H + 6, W + 6, "bf16", g)
        _lib.call = orig
        torch.cuda.synchronize()
        gemms = [(nm, aa) for nm, aa in calls if nm.startswith("dsr_tc_gemm")]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for nm, aa in gemms:
            orig(nm, *aa)
        e0.record()
        for _ in range(a.iters):
            for nm, aa in gemms:
                orig(nm, *aa)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.iters
        macs = N * (H + 6) * (W + 6) * Co * Ci * 49
        out = dict(g=g, kernel=[nm for nm, _ in gemms], ms=round(ms, 4), alg_tflops=round(2 * macs / ms / 1e9, 1))
        if os.environ.get("DSR_BENCH_WAITS") and gemms[0][0] in ("dsr_tc_gemm2", "dsr_tc_gemm3"):
            lib = _lib.load()
            buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
            setdbg = lib.dsr_tc2_set_debug if gemms[0][0] == "dsr_tc_gemm2" else lib.dsr_tc3_set_debug
            setdbg(buf.data_ptr())
            orig(gemms[0][0], *gemms[0][1])
            torch.cuda.synchronize()
            setdbg(None)
            b = buf.view(148, 16).double()
            names = ["mma_wait_patch", "mma_wait_w", "mma_wait_acc", "mma_total", "epi_wait_acc", "epi_total", "pprod_wait", "wprod_wait", "epi_ld", "setup"]
            out["waits_mean_cycles"] = {n: round(float(b[:, i].mean())) for i, n in enumerate(names)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
