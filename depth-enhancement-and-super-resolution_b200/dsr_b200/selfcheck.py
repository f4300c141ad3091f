"""smoke(): one small optimize_parameters step on cuda:0 checked against the CPU oracle.
(The oracle is imported here as the CHECKER only - see oracle/__init__.py.)"""
import os
import sys

import numpy as np
import torch


def smoke_step(B=1, H=128, W=128, verbose=True):
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import ref_step
    from . import _lib, main_model, options

    if not torch.cuda.is_available():
        raise RuntimeError("dsr_b200.smoke needs a CUDA device (there is no CPU path)")
    _lib.load()
    torch.cuda.set_device(0)
    opt = options.main_flags(gpu_ids=[], batch_size=B, crop_size_h=H, crop_size_w=W, name="smoke",
                             checkpoints_dir="/tmp/dsr_smoke")
    torch.manual_seed(0)
    host = main_model.MainModel(opt)                       # CPU-seeded reference initialisation
    sds = {n: {k: v.detach().clone() for k, v in getattr(host, "net" + n).state_dict().items()} for n in host.model_names}
    opt.gpu_ids = [0]
    model = main_model.MainModel(opt)
    for n, sd in sds.items():
        model._unwrap(getattr(model, "net" + n)).load_state_dict(sd)
    model._train()
    batch = ref_step.synthetic_batch(B, H, W, seed=1, depth_kind="smooth")
    launches0 = _lib.LAUNCHES
    np.random.seed(0)
    model.set_input(batch)
    model.optimize_parameters(0, 1)
    torch.cuda.synchronize()
    launches = _lib.LAUNCHES - launches0
    np.random.seed(0)
    ref = ref_step.OracleStep(sds, lr=opt.lr).step(batch)
    worst = 0.0
    for k in ("pred_syn_depth", "pred_real_depth"):
        a, b = getattr(model, k).detach().cpu().double(), ref["tensors"][k].detach().double()
        worst = max(worst, float((a - b).norm() / b.norm()))
    lg, lr_ = float(model.loss_G), ref["losses"]["G"]
    ok = worst <= 1e-2 and abs(lg - lr_) <= 1e-3 * abs(lr_)
    if verbose:
        print(f"[dsr_b200 smoke] kernels launched={launches} pred rel-L2={worst:.3e} loss_G={lg:.6f} oracle={lr_:.6f} "
              f"-> {'OK' if ok else 'MISMATCH'}")
    if not ok:
        raise AssertionError(f"smoke step does not match the oracle: rel-L2 {worst:.3e}, loss {lg} vs {lr_}")
    return dict(launches=launches, rel_l2=worst, loss_G=lg)
