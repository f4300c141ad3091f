"""Image Guidance Network step (dsr_b200.I2D_model.I2DModel, SURVEY.md section 8f rank 2 / BASELINE configs[0]):
oracle vs golden vectors of the live reference (CPU), and the CUDA step vs golden + oracle (GPU) at the main-step gates."""
import numpy as np
import pytest
import torch

from oracle import ref_step
from tests_proj import proj_vec
from util import cosine, grad_is_informative, load_golden, rel_l2

# conv biases of the Task U-Net that feed an affine-less InstanceNorm carry only rounding noise (see util.grad_is_informative);
# its outermost up-convolution (model.model.3) and the first down-convolution without a norm (model.model.0) are informative
I2D_INFORMATIVE_BIASES = {"model.model.0.bias", "model.model.3.bias",
                          "model.model.1.model.3.model.3.model.3.model.3.model.3.model.1.bias"}


def _informative(n):
    return n.endswith("weight") or n in I2D_INFORMATIVE_BIASES


def _host_model(B=2, H=128, W=128, gpu_ids=()):
    from dsr_b200 import I2D_model, options
    opt = options.i2d_flags(gpu_ids=[], batch_size=B, crop_size_h=H, crop_size_w=W, name="t", checkpoints_dir="/tmp/dsr_ck")
    torch.manual_seed(0)
    np.random.seed(0)
    host = I2D_model.I2DModel(opt)
    if not gpu_ids:
        return host, None
    sds = {n: {k: v.detach().clone() for k, v in getattr(host, "net" + n).state_dict().items()} for n in host.model_names}
    opt.gpu_ids = list(gpu_ids)
    dev = I2D_model.I2DModel(opt)
    for n, sd in sds.items():
        dev._unwrap(getattr(dev, "net" + n)).load_state_dict(sd)
    return dev, sds


def test_i2d_oracle_matches_reference_two_steps():
    g = load_golden("i2d_step_b2_128.npz")
    host, _ = _host_model()
    sds = {n: getattr(host, "net" + n).state_dict() for n in host.model_names}
    for name, sd in sds.items():
        assert list(sd.keys()) == list(g["wkeys/" + name])
        a = float(sum(v.double().abs().sum() for v in sd.values()))
        assert abs(a - g["wsum/" + name][1]) <= 1e-9 * a, name
    orc = ref_step.OracleI2DStep(sds, lr=2e-4)
    batch = ref_step.synthetic_batch(2, 128, 128, seed=1, depth_kind="smooth")
    for it in range(2):
        out = orc.step(batch)
        p = f"s{it}/"
        for k in ("pred_syn_depth", "pred_real_depth"):
            assert rel_l2(out["tensors"][k].detach(), g[p + k]) <= (2e-5 if it == 0 else 2e-3), (k, it)
        for k, v in out["losses"].items():
            ref = float(g[p + "loss/" + k])
            assert abs(v - ref) <= (2e-5 if it == 0 else 2e-3) * max(abs(ref), 1e-3), (k, it, v, ref)
        if it == 0:
            for gi, n in enumerate(orc.sd["Task"]):
                gr = out["grads"][n].double().flatten()
                ref_norm, ref_proj = g[p + f"gstat/Task/{n}"]
                if _informative(n):
                    assert abs(float(gr.norm()) - ref_norm) <= 2e-3 * ref_norm, n
                    assert abs(float(gr @ proj_vec(gr.numel(), 1000 + gi)) - ref_proj) <= 2e-3 * ref_norm, n


@pytest.mark.gpu
def test_i2d_step_matches_reference_golden_and_oracle(built_lib):
    g = load_golden("i2d_step_b2_128.npz")
    model, sds = _host_model(gpu_ids=[0])
    model._train()
    orc = ref_step.OracleI2DStep(sds, lr=2e-4)
    batch = ref_step.synthetic_batch(2, 128, 128, seed=1, depth_kind="smooth")
    for it in range(2):
        ref = orc.step(batch)
        model.set_input(batch)
        model.optimize_parameters(it)
        p = f"s{it}/"
        for k in ("pred_syn_depth", "pred_real_depth"):
            assert rel_l2(getattr(model, k).detach().cpu(), g[p + k]) <= 1e-2, (k, it)                 # the gate
            assert rel_l2(getattr(model, k).detach().cpu(), ref["tensors"][k].detach()) <= (2e-3 if it == 0 else 1e-2), (k, it)
        losses = dict(model.get_current_losses(), G=float(model.loss_G))
        for k, v in losses.items():
            want = float(g[p + "loss/" + k])
            assert abs(v - want) <= (1e-3 if it == 0 else 1e-2) * max(abs(want), 1e-3), (k, it, v, want)
        if it == 0:
            params = dict(model._unwrap(model.netTask).named_parameters())
            fa, fb = [], []
            for n in orc.sd["Task"]:
                if _informative(n):
                    mine, gr = params[n].grad.detach().cpu(), ref["grads"][n]
                    assert cosine(mine, gr) >= 0.999, (n, cosine(mine, gr))
                    fa.append(mine.flatten()); fb.append(gr.flatten())
            assert cosine(torch.cat(fa), torch.cat(fb)) >= 0.999
    vis = model.get_current_visuals()
    assert all(k in vis for k in model.visual_names)


@pytest.mark.gpu
def test_i2d_graph_replay_matches_eager(built_lib):
    """same weights, same batch: the replayed step's loss equals the eager step's"""
    from dsr_b200 import ops
    batch = ref_step.synthetic_batch(2, 128, 128, seed=3, depth_kind="smooth")
    model, _ = _host_model(gpu_ids=[0])
    model._train()
    model.use_graph = True
    for it in range(4):                                  # 2 eager warm-up steps, capture + replay, one more replay
        model.set_input(batch)
        model.optimize_parameters(it)
    assert model._graph is not None and model.optimizer_G.n_steps == 4
    state = (model.arena.flat, model.arena.exp_avg, model.arena.exp_avg_sq, model.optimizer_G.step_dev)
    snap = [t.clone() for t in state]
    out = []
    for use_graph in (True, False):
        for t, s0 in zip(state, snap):
            t.copy_(s0)
        ops.WEIGHT_EPOCH += 1
        model.use_graph = use_graph
        model.set_input(batch)
        model.optimize_parameters(9)
        out.append((float(model.loss_G), model.pred_real_depth.detach().clone()))
    assert abs(out[0][0] - out[1][0]) <= 1e-5 * abs(out[1][0])
    assert rel_l2(out[0][1].cpu(), out[1][1].cpu()) <= 1e-5


@pytest.mark.gpu
def test_i2d_save_all_writes_pngs(built_lib, tmp_path):
    """--save_all in the test stage (I2D_model.py:171-182): uint16 PNGs of the real prediction, rows [16, H-16), bit-exact
    against the numpy restatement of the export"""
    from dsr_b200 import I2D_model, io, options
    from oracle import ref_io
    opt = options.i2d_flags(gpu_ids=[0], batch_size=2, crop_size_h=128, crop_size_w=128, name="t", checkpoints_dir="/tmp/dsr_ck",
                            save_all=True, save_image_folder=str(tmp_path) + "/")
    torch.manual_seed(0)
    model = I2D_model.I2DModel(opt)
    model.eval()
    batch = ref_step.synthetic_batch(2, 128, 128, seed=3, depth_kind="smooth")
    batch["B_paths"] = ["/somewhere/real_0042.png", "x/real_0043.jpg"]
    with torch.no_grad():
        model.set_input(batch)
        model.forward()                      # train stage: nothing is written
        assert not list(tmp_path.iterdir())
        model.forward("test")
    want = ref_io.depth_to_u16(model.pred_real_depth.detach().cpu().numpy(), 16)
    for i, name in enumerate(("real_0042.png", "real_0043.png")):
        out = io.read_png_u16(str(tmp_path / name))
        assert out.shape == (96, 128) and out.dtype == np.uint16 and np.array_equal(out, want[i])
