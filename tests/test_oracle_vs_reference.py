"""The oracle against the LIVE reference (not stored vectors): the unmodified reference, byte-compiled from /root/reference into
``oracle/_ref`` by oracle/build_ref.py, runs one ``set_input`` + ``optimize_parameters`` on the host through its own
``TrainOptions`` / ``MainModel`` (oracle/ref_live.py); the oracle restatement takes the reference's OWN initial weights and
the same batch and must reproduce masks bit for bit, predictions, every loss term and the trainable gradients.
CPU only.  Skipped where neither /root/reference nor a staged ``oracle/_ref`` exists."""
import sys

import numpy as np
import pytest
import torch

from oracle import build_ref, ref_live, ref_step


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


class _reference_modules:
    """the reference has a top-level package called ``util`` - so has tests/ (util.py): swap them for the duration"""

    def __enter__(self):
        self.saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "util" or k.startswith("util.")}

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k == "util" or k.startswith("util.")]:
            del sys.modules[k]
        sys.modules.update(self.saved)
        return False


@pytest.mark.skipif(not (build_ref.build_ref() or build_ref.available()), reason="no reference tree and no staged oracle/_ref")
def test_oracle_step_matches_live_reference():
    B, H, W = 1, 128, 128
    batch = ref_step.synthetic_batch(B, H, W, seed=5, depth_kind="smooth")
    with _reference_modules():
        ref = ref_live.make_model(B, H, W, seed=3)
        sds = {n: {k: v.detach().clone() for k, v in getattr(ref, "net" + n).state_dict().items()} for n in ref.model_names}
        np.random.seed(11)
        ref.set_input(batch)
        ref.optimize_parameters(0, 1)
    orc = ref_step.OracleStep(sds, lr=1e-4)
    np.random.seed(11)
    out = orc.step(batch)
    t = out["tensors"]
    for k in ("syn_mask", "real_mask", "real_hole_mask", "gt_mask_syn", "gt_mask_real"):
        assert np.array_equal(t[k].numpy().astype(np.uint8), getattr(ref, k).numpy().astype(np.uint8)), k      # bit-exact
    for k in ("syn2real_depth", "syn_depth_by_image", "real_depth_by_image", "pred_syn_depth", "pred_real_depth"):
        assert rel_l2(t[k].detach(), getattr(ref, k).detach()) <= 2e-5, k
    live = ref.get_current_losses()
    for k, v in out["losses"].items():
        if k in live:
            assert abs(v - live[k]) <= 2e-5 * max(abs(live[k]), 1e-12), (k, v, live[k])
    assert abs(out["losses"]["G"] - float(ref.loss_G)) <= 2e-5 * abs(float(ref.loss_G))
    n_checked = 0
    for net in ("Depth_f", "Task"):
        for n, prm in getattr(ref, "net" + net).named_parameters():
            g_ref, g_orc = prm.grad.detach().double().flatten(), out["grads"][(net, n)].detach().double().flatten()
            # (a conv bias in front of an affine-less InstanceNorm has an exactly-zero gradient in exact arithmetic: both sides
            # hold rounding noise there - tests/util.grad_is_informative)
            if n.endswith("weight") and float(g_ref.norm()) > 1e-12:
                cos = float(g_ref @ g_orc / (g_ref.norm() * g_orc.norm()))
                assert cos >= 0.99999, (net, n, cos)
                n_checked += 1
    assert n_checked >= 30
