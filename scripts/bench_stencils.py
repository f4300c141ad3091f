#!/usr/bin/env python
"""HBM roofline of the bandwidth-bound kernels of the path (loss / normal stencils, masks, resize, norm layers, Adam)
at a size that cannot live in the 126 MB L2: B planes of 512x640 (the full-size 640x480 frame as the nets see it).

achieved GB/s = ALGORITHMIC bytes (fp32 I/O of the op, each tensor counted once) / mean launch
time (CUDA events on the launching stream, 3 warm-up + `--iters` launches); peak = MEASURED_PEAKS.json hbm_gbs.

  python scripts/bench_stencils.py [--batch 96] [--iters 20] [--out gpurun_out/stencils.json]
"""
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
    ap.add_argument("--batch", type=int, default=96)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="", help="comma-separated substrings of the case names to run")
    args = ap.parse_args()
    from dsr_b200 import stencil_bench
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    res = stencil_bench.run(args.batch, args.iters, peak, verbose=True, only=[o for o in args.only.split(',') if o] or None)
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
