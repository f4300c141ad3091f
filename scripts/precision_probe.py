#!/usr/bin/env python
"""Margins of the parity gates (pred rel-L2 <= 1e-2, losses 1e-3, gradient cosine >= 0.999) for the first golden step
(B = 2, 128x128) under different precision settings of ops.CONFIG.  GPU box only; test infrastructure.

  python scripts/precision_probe.py "wgrad_passes=2" "wgrad_passes=1" ...
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "depth-enhancement-and-super-resolution_b200")):
    sys.path.insert(0, p)

import numpy as np
import torch

from oracle import ref_step
from util import build_host_model, cosine, grad_is_informative, load_golden, rehome, rel_l2, state_dicts


def main():
    from dsr_b200 import ops
    import __graft_entry__ as ge
    ge.build()
    g = load_golden("step_b2_128.npz")
    host = build_host_model(2, 128, 128)
    sds = state_dicts(host)
    batch = ref_step.synthetic_batch(2, 128, 128, seed=1, depth_kind="smooth")
    np.random.seed(0)
    ref = ref_step.OracleStep(sds, lr=1e-4).step(batch)
    base = dict(ops.CONFIG)
    for spec in ["default"] + sys.argv[1:]:
        ops.CONFIG.clear(); ops.CONFIG.update(base)
        if spec != "default":
            for kv in spec.split(","):
                k, v = kv.split("=")
                ops.CONFIG[k] = type(base[k])(v) if not isinstance(base[k], bool) else v in ("1", "True")
        ops.WEIGHT_EPOCH += 1
        model = rehome(host, host.opt, [0])
        model._train()
        np.random.seed(0)
        model.set_input(batch)
        model.optimize_parameters(0, 1)
        pred = max(rel_l2(getattr(model, k).detach().cpu(), g["s0/" + k]) for k in ("pred_syn_depth", "pred_real_depth"))
        froz = {k: rel_l2(getattr(model, k).detach().cpu(), g["s0/" + k]) for k in ("syn2real_depth", "real_depth_by_image")}
        losses = model.get_current_losses()
        lerr = max(abs(losses[k] - float(g["s0/loss/" + k])) / max(abs(float(g["s0/loss/" + k])), 1e-3) for k in losses)
        worst, wname, fa, fb = 1.0, "", [], []
        for net in ("Depth_f", "Task"):
            params = dict(model._unwrap(getattr(model, "net" + net)).named_parameters())
            for n, gr in ((n, ref["grads"][(net, n)]) for n in sds[net]):
                if grad_is_informative(net, n):
                    mine = params[n].grad.detach().cpu()
                    c = cosine(mine, gr)
                    if c < worst:
                        worst, wname = c, f"{net}.{n}"
                    fa.append(mine.flatten()); fb.append(gr.flatten())
        print(f"{spec:40s} pred {pred:.2e}  loss {lerr:.2e}  worst cos {worst:.6f} ({wname})  flat cos {cosine(torch.cat(fa), torch.cat(fb)):.7f}  frozen outs {froz['syn2real_depth']:.2e} {froz['real_depth_by_image']:.2e}", flush=True)
        model.arena.release()


if __name__ == "__main__":
    main()
