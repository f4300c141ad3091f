"""The oracle (oracle/) against the golden vectors generated from the LIVE reference
(tests/golden/make_golden.py).  CPU only.  This is the pin that lets the GPU parity tests trust it."""
import numpy as np
import torch

from oracle import ref_ops, ref_step
from util import build_host_model, grad_is_informative, load_golden, rel_l2, state_dicts


def test_ops_against_reference_vectors():
    g = load_golden("ops.npz")
    d, img = torch.from_numpy(g["d"]), torch.from_numpy(g["img"])
    K, crop = torch.from_numpy(g["K"]), torch.from_numpy(g["crop"])
    hole, valid = ref_ops.hole_valid_masks(d)
    assert np.array_equal(hole.numpy().astype(np.uint8), g["hole"])          # bit-exact
    assert np.array_equal(valid.numpy().astype(np.uint8), g["valid"])        # bit-exact
    assert np.abs(ref_ops.surface_normals_old(d).numpy() - g["normals_old"]).max() <= 1e-6   # fp32: ulp-level
    n_new = ref_ops.surface_normals_new(d, K, crop).numpy()
    assert np.abs(n_new - g["normals_new"]).max() <= 1e-7
    tv = float(ref_ops.tv_loss(ref_ops.surface_normals_old(d) * 100))
    assert abs(tv - float(g["tv"])) <= 1e-6 * abs(float(g["tv"]))
    sm = float(ref_ops.smooth_loss(torch.from_numpy(g["smooth_d"]), torch.from_numpy(g["smooth_img"]), 3))
    assert abs(sm - float(g["smooth"])) <= 1e-6 * abs(float(g["smooth"]))
    ss = float(ref_ops.ssim(torch.from_numpy(g["ssim_a"]), torch.from_numpy(g["ssim_b"])))
    assert abs(ss - float(g["ssim"])) <= 1e-6


def test_constructors_reproduce_reference_weights_and_keys():
    g = load_golden("step_b2_128.npz")
    m = build_host_model(2, 128, 128)
    for name, sd in state_dicts(m).items():
        assert list(sd.keys()) == list(g["wkeys/" + name])
        assert [str(tuple(v.shape)) for v in sd.values()] == list(g["wshapes/" + name])
        s = float(sum(v.double().sum() for v in sd.values()))
        a = float(sum(v.double().abs().sum() for v in sd.values()))
        assert abs(s - g["wsum/" + name][0]) <= 1e-9 * max(1.0, abs(a))
        assert abs(a - g["wsum/" + name][1]) <= 1e-9 * a
        assert sum(v.numel() for v in sd.values()) == int(g["wsum/" + name][2])


def test_oracle_step_matches_reference_two_steps():
    g = load_golden("step_b2_128.npz")
    torch.set_num_threads(max(1, torch.get_num_threads()))
    m = build_host_model(2, 128, 128)
    orc = ref_step.OracleStep(state_dicts(m), lr=1e-4)
    batch = ref_step.synthetic_batch(2, 128, 128, seed=1, depth_kind="smooth")
    for k in ("A_i", "B_i", "A_d", "B_d"):
        assert np.array_equal(batch[k].numpy(), g["in/" + k])
    np.random.seed(0)
    for it in range(2):
        out = orc.step(batch)
        t, p = out["tensors"], f"s{it}/"
        if it == 0:
            for k in ("syn_mask", "real_mask", "real_hole_mask", "gt_mask_syn", "gt_mask_real"):
                assert np.array_equal(t[k].numpy().astype(np.uint8), g[p + k]), k      # integer work: bit-exact
        for k in ("syn2real_depth", "syn_depth_by_image", "real_depth_by_image", "pred_syn_depth", "pred_real_depth"):
            assert rel_l2(t[k].detach(), g[p + k]) <= (2e-5 if it == 0 else 2e-3), (k, it)
        for k, v in out["losses"].items():
            ref = float(g[p + "loss/" + k])
            tol = 2e-5 if it == 0 else 2e-3
            assert abs(v - ref) <= tol * max(abs(ref), 1e-3), (k, it, v, ref)
        if it == 0:
            gi, wnorm = 0, 1.0
            for net in ("Depth_f", "Task"):
                for n in orc.sd[net]:
                    gr = out["grads"][(net, n)].double().flatten()
                    ref_norm, ref_proj = g[p + f"gstat/{net}/{n}"]
                    from tests_proj import proj_vec
                    proj = float(gr @ proj_vec(gr.numel(), 1000 + gi))
                    gi += 1
                    if grad_is_informative(net, n):
                        assert abs(float(gr.norm()) - ref_norm) <= 2e-3 * ref_norm, (net, n)
                        assert abs(proj - ref_proj) <= 2e-3 * ref_norm, (net, n)
                        wnorm = ref_norm
                    else:   # pure rounding noise in both implementations: only check it is negligible
                        assert float(gr.norm()) <= 1e-2 * max(wnorm, 1e-3) and ref_norm <= 1e-2 * max(wnorm, 1e-3), (net, n)
