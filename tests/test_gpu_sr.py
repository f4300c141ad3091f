"""GPU parity of the super-resolution path (csrc/resize.cu, dsr_b200.main_sr_model.MainSRModel) against golden vectors
of the live reference (tests/golden/resize.npz, sr_step_b1_128.npz) and the oracle on the same seeded inputs.
Gates as for the main step: integer / index work bit-exact, pred rel-L2 <= 1e-2, losses within 1e-3 relative,
parameter-gradient cosine >= 0.999."""
import numpy as np
import pytest
import torch

from oracle import ref_ops, ref_step
from util import build_host_model, cosine, grad_is_informative, load_golden, rehome, rel_l2, state_dicts

pytestmark = pytest.mark.gpu


def test_bicubic_nearest_match_reference_vectors(built_lib):
    from dsr_b200 import ops
    g = load_golden("resize.npz")
    x = torch.from_numpy(g["x"]).cuda()
    for name, size in (("down", (12, 18)), ("up", (48, 72)), ("odd", (10, 50))):
        assert np.abs(ops.bicubic(x, size).cpu().numpy() - g["bicubic_" + name]).max() <= 5e-6, name     # fp32 tolerance
        assert np.array_equal(ops.nearest(x, size).cpu().numpy(), g["nearest_" + name]), name            # bit-exact
    xg = x.clone().requires_grad_(True)
    (ops.bicubic(xg, (12, 18)) * torch.from_numpy(g["gy_down"]).cuda()).sum().backward()
    assert np.abs(xg.grad.cpu().numpy() - g["gx_down"]).max() <= 5e-6


@pytest.mark.parametrize("shape,size", [((2, 128, 20, 28), (40, 56)), ((1, 128, 33, 17), (16, 8)), ((3, 1, 64, 80), (32, 40)),
                                        ((2, 3, 1, 5), (2, 10)), ((1, 6, 40, 40), (40, 40)),
                                        ((1, 128, 9, 13), (18, 26)), ((2, 32, 16, 12), (32, 24)), ((2, 4, 3, 300), (6, 600)),
                                        ((2, 1, 30, 36), (15, 18)), ((1, 2, 70, 264), (35, 132)), ((1, 1, 2, 4), (1, 2))])
def test_bicubic_layouts_and_adjoint(built_lib, shape, size):
    """channels-last activations (C = 128 feature maps, float4 path), planes, degenerate sizes; backward = exact adjoint"""
    from dsr_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.rand(shape, generator=g) * 2 - 1
    ref = ref_ops.bicubic_restated(x, size)
    for cl in (False, True):
        xd = x.cuda()
        if cl:
            xd = xd.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)          # NHWC-backed like the network outputs
        xd.requires_grad_(True)
        y = ops.bicubic(xd, size)
        assert float((y.detach().cpu() - ref).abs().max()) <= 1e-5
        gy = torch.rand(y.shape, generator=g)
        (y * gy.cuda()).sum().backward()
        xr = x.clone().requires_grad_(True)
        (ref_ops.bicubic_restated(xr, size) * gy).sum().backward()
        assert float((xd.grad.cpu() - xr.grad).abs().max()) <= 1e-5
    assert torch.equal(ops.nearest(x.cuda(), size).cpu(), ref_ops.nearest(x, size))


@pytest.fixture(scope="module")
def sr_model(built_lib):
    host = build_host_model(1, 128, 128, sr=True)
    return rehome(host, host.opt, [0]), state_dicts(host)


def test_sr_step_matches_reference_golden_and_oracle(sr_model):
    model, sds = sr_model
    g = load_golden("sr_step_b1_128.npz")
    batch = ref_step.synthetic_sr_batch(1, 128, 128, seed=1, depth_kind="smooth")
    orc = ref_step.OracleSRStep(sds, (128, 128), lr=2e-5)
    model._train()
    np.random.seed(0)
    rng_state = np.random.get_state()
    for it in range(2):
        np.random.set_state(rng_state)
        ref = orc.step(batch)
        np.random.set_state(rng_state)
        model.set_input(batch)
        model.optimize_parameters(it, 1)
        rng_state = np.random.get_state()
        p = f"s{it}/"
        if it == 0:                                                   # integer / index work: bit-exact
            for k in ("syn_mask", "real_mask", "real_hole_mask"):     # real_* are the nearest-resized LR masks (:394-395)
                assert np.array_equal(getattr(model, k).cpu().numpy().astype(np.uint8), g[p + k]), k
            assert np.array_equal(model.gt_mask_real.cpu().numpy(), g[p + "gt_mask_real"])
            assert np.array_equal(model.gt_mask_syn.cpu().numpy(), g[p + "gt_mask_syn"])
            assert tuple(model.pred_real_depth_hr.shape[2:]) == (256, 256) and tuple(model.pred_real_depth.shape[2:]) == (128, 128)
            for k in ("syn2real_depth", "syn_depth_by_image", "real_depth_by_image", "real_depth"):
                assert rel_l2(getattr(model, k).detach().cpu(), g[p + k].astype(np.float32)) <= 1e-2, k
        for k in ("pred_syn_depth", "pred_real_depth", "pred_real_depth_hr"):
            assert rel_l2(getattr(model, k).detach().cpu(), g[p + k]) <= 1e-2, (k, it)                    # the gate
            assert rel_l2(getattr(model, k).detach().cpu(), ref["tensors"][k].detach()) <= (2e-3 if it == 0 else 1e-2), (k, it)
        losses = model.get_current_losses()
        tol = 1e-3 if it == 0 else 1e-2
        for k in g.files:
            if not k.startswith(p + "loss/"):
                continue
            name, want = k[len(p) + 5:], float(g[k])
            if name == "G":
                v = float(model.loss_G)
            elif name.startswith("mean_of_abs"):
                v = float(getattr(model, "loss_" + name))
            else:
                v = losses[name]
            assert abs(v - want) <= tol * max(abs(want), 1e-3), (name, it, v, want)
        if it == 0:
            flat_a, flat_b = [], []
            for net in ("Depth_f", "Task"):
                params = dict(model._unwrap(getattr(model, "net" + net)).named_parameters())
                for n in orc.sd[net]:
                    if grad_is_informative(net, n):
                        mine, gr = params[n].grad.detach().cpu(), ref["grads"][(net, n)]
                        assert cosine(mine, gr) >= 0.999, (net, n, cosine(mine, gr))
                        flat_a.append(mine.flatten()); flat_b.append(gr.flatten())
            assert cosine(torch.cat(flat_a), torch.cat(flat_b)) >= 0.999


def test_sr_graph_replay_and_test_stage(built_lib):
    """the SR step under CUDA-graph replay (same machinery as MainModel) and the forward-only test stage (:289-293)"""
    host = build_host_model(1, 128, 128, sr=True)
    model = rehome(host, host.opt, [0])             # graph mode from the first step (the capture stream owns the autograd nodes)
    model._train()
    batch = ref_step.synthetic_sr_batch(1, 128, 128, seed=2, depth_kind="smooth")
    model.use_graph = True
    np.random.seed(3)
    vals = []
    for it in range(4):
        model.set_input(batch)
        model.optimize_parameters(it, 1)
        vals.append(float(model.loss_G))
    assert model._graph is not None and all(np.isfinite(v) for v in vals)
    model.use_graph = False
    model.reset_graph()
    model.eval()
    with torch.no_grad():
        model.set_input(batch)
        model.calculate("test")
    assert tuple(model.pred_real_depth_hr.shape) == (1, 1, 256, 256)
    assert torch.equal(model.depth_masked.cpu(), batch["B_d"])           # p = 0 in the test stage: no artificial holes
    model._train()
