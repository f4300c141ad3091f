"""The super-resolution part of the oracle (oracle/ref_ops.bicubic / nearest / loss_stack_sr, ref_step.OracleSRStep)
against golden vectors recorded from the LIVE reference (tests/golden/make_golden.py resize sr).  CPU only."""
import numpy as np
import torch

from oracle import ref_ops, ref_step
from tests_proj import proj_vec
from util import build_host_model, grad_is_informative, load_golden, rel_l2, state_dicts


def test_resize_against_reference_vectors():
    g = load_golden("resize.npz")
    x = torch.from_numpy(g["x"])
    for name, size in (("down", (12, 18)), ("up", (48, 72)), ("odd", (10, 50))):
        assert np.abs(ref_ops.bicubic_restated(x, size).numpy() - g["bicubic_" + name]).max() <= 5e-6, name
        assert np.array_equal(ref_ops.bicubic(x, size).numpy(), g["bicubic_" + name]), name
        assert np.array_equal(ref_ops.nearest(x, size).numpy(), g["nearest_" + name]), name       # index work: bit-exact
    xg = x.clone().requires_grad_(True)
    (ref_ops.bicubic_restated(xg, (12, 18)) * torch.from_numpy(g["gy_down"])).sum().backward()
    assert np.abs(xg.grad.numpy() - g["gx_down"]).max() <= 5e-6


def test_oracle_sr_step_matches_reference_two_steps():
    g = load_golden("sr_step_b1_128.npz")
    m = build_host_model(1, 128, 128, sr=True)
    sds = state_dicts(m)
    for name, sd in sds.items():                       # same seed + same constructor order => the golden run's weights
        a = float(sum(v.double().abs().sum() for v in sd.values()))
        assert abs(a - g["wsum/" + name][1]) <= 1e-9 * a, name
    orc = ref_step.OracleSRStep(sds, (128, 128), lr=2e-5)
    batch = ref_step.synthetic_sr_batch(1, 128, 128, seed=1, depth_kind="smooth")
    np.random.seed(0)
    for it in range(2):
        out = orc.step(batch)
        t, p = out["tensors"], f"s{it}/"
        if it == 0:
            for k in ("syn_mask", "gt_mask_syn", "gt_mask_real"):
                assert np.array_equal(t[k].numpy().astype(np.uint8), g[p + k]), k                 # bit-exact
            for k in ("real_mask", "real_hole_mask"):                                             # LR (nearest) after backward_G
                assert np.array_equal(out["visuals"][k].numpy().astype(np.uint8), g[p + k]), k
            assert rel_l2(out["visuals"]["real_depth"], g[p + "real_depth"].astype(np.float32)) <= 1e-3   # fp16-stored
        for k in ("pred_syn_depth", "pred_real_depth", "pred_real_depth_hr"):
            assert rel_l2(t[k].detach(), g[p + k]) <= (2e-5 if it == 0 else 2e-3), (k, it)
        for k, v in out["losses"].items():
            ref = float(g[p + "loss/" + k])
            tol = 2e-5 if it == 0 else 2e-3
            assert abs(v - ref) <= tol * max(abs(ref), 1e-3), (k, it, v, ref)
        if it == 0:
            gi = 0
            for net in ("Depth_f", "Task"):
                for n in orc.sd[net]:
                    gr = out["grads"][(net, n)].double().flatten()
                    ref_norm, ref_proj = g[p + f"gstat/{net}/{n}"]
                    proj = float(gr @ proj_vec(gr.numel(), 1000 + gi))
                    gi += 1
                    if grad_is_informative(net, n):
                        assert abs(float(gr.norm()) - ref_norm) <= 2e-3 * ref_norm, (net, n)
                        assert abs(proj - ref_proj) <= 2e-3 * ref_norm, (net, n)
