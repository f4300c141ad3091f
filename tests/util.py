"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def _model_cls(sr):
    if sr:
        from dsr_b200 import main_sr_model
        return main_sr_model.MainSRModel
    from dsr_b200 import main_model
    return main_model.MainModel


SR_FLAGS = dict(w_real_l1_d=90.0, w_syn_norm=3.0, w_syn_holes=1600.0, w_real_holes=1600.0, lr=0.00002, SR=True)   # README.md:86


def build_host_model(B, H, W, gpu_ids=(), seed=0, sr=False, **kw):
    """MainModel with the seeded reference initialisation (torch.manual_seed(seed) before the
    constructors == the weights the golden run had; checked by test_host_logic)."""
    from dsr_b200 import options
    if sr:
        kw = dict(SR_FLAGS, **kw)
    opt = options.main_flags(gpu_ids=list(gpu_ids), batch_size=B, crop_size_h=H, crop_size_w=W, name="t",
                             checkpoints_dir="/tmp/dsr_ck", **kw)
    torch.manual_seed(seed)
    np.random.seed(seed)
    dev_state = None
    if gpu_ids:
        # the reference initialises on the device when gpu_ids is set (CUDA RNG); for parity we want
        # the CPU-seeded weights, so build on the host and move afterwards
        opt.gpu_ids = []
    m = _model_cls(sr)(opt)
    if gpu_ids:
        m = rehome(m, opt, list(gpu_ids))
    return m


def rehome(host_model, opt, gpu_ids):
    """Build a device model and load the host model's (CPU-seeded) weights into it."""
    from dsr_b200 import main_model
    sds = state_dicts(host_model)
    opt.gpu_ids = gpu_ids
    m = type(host_model)(opt)
    for name, sd in sds.items():
        main_model.MainModel._unwrap(getattr(m, "net" + name)).load_state_dict(sd)
    return m


def state_dicts(model):
    return {name: {k: v.detach().cpu().clone() for k, v in model._unwrap(getattr(model, "net" + name)).state_dict().items()}
            for name in model.model_names}


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cosine(a, b):
    a, b = torch.as_tensor(a).double().flatten(), torch.as_tensor(b).double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


INFORMATIVE_BIASES = {
    ("Depth_f", "model.23.bias"),
    ("Task", "model.model.0.bias"),
    ("Task", "model.model.3.bias"),
    ("Task", "model.model.1.model.3.model.3.model.3.model.3.model.3.model.1.bias"),
}


def grad_is_informative(net, name):
    """A conv bias that feeds an affine-less InstanceNorm has an exactly-zero gradient in exact
    arithmetic (the reference produces ~1e-9..1e-4 rounding noise there): only the
    weights and the four biases with no norm behind them carry signal."""
    return name.endswith("weight") or (net, name) in INFORMATIVE_BIASES


def _loss_value(model, losses, name):
    if name == "G":
        return float(model.loss_G)
    if name.startswith("mean_of_abs"):
        return float(getattr(model, "loss_" + name))
    return losses[name]


def compare_step_with_oracle(model, ref, orc, pred_keys=("pred_syn_depth", "pred_real_depth"), tag=""):
    """The BASELINE.json gates of ONE step of `model` (already run) against the oracle's result `ref` on the same inputs:
    input-derived masks bit-exact, pred rel-L2 <= 1e-2, every loss term within 1e-3 relative, gradient cosine >= 0.999 per
    informative tensor and flattened.  Returns the measured margins (worst values) so the caller can log them."""
    rep = dict(tag=tag)

    def want(k):       # the SR step overwrites some real_* tensors with their LR versions (main_sr_model.py:394-398)
        v = ref.get("visuals", {}).get(k)
        return (v if v is not None else ref["tensors"][k]).detach()

    for k in ("syn_mask", "real_mask", "real_hole_mask", "gt_mask_real", "gt_mask_syn"):
        a = getattr(model, k).detach().cpu().numpy()
        b = want(k).cpu().numpy()
        assert a.shape == b.shape and np.array_equal(a.astype(np.int64), b.astype(np.int64)), (tag, k)
    rep["pred_rel_l2"] = {}
    for k in ("syn2real_depth", "syn_depth_by_image", "real_depth_by_image") + tuple(pred_keys):
        r = rel_l2(getattr(model, k).detach().cpu(), want(k))
        rep["pred_rel_l2"][k] = r
        assert r <= 1e-2, (tag, k, r)
    losses = model.get_current_losses()
    worst_loss = 0.0
    for k, want in ref["losses"].items():
        v = _loss_value(model, losses, k)
        err = abs(v - want) / max(abs(want), 1e-3)
        worst_loss = max(worst_loss, err)
        assert err <= 1e-3, (tag, k, v, want)
    rep["worst_loss_rel"] = worst_loss
    flat_a, flat_b, worst, worst_name = [], [], 1.0, None
    for net in ("Depth_f", "Task"):
        params = dict(model._unwrap(getattr(model, "net" + net)).named_parameters())
        for n in orc.sd[net]:
            if not grad_is_informative(net, n):
                continue
            mine, gr = params[n].grad.detach().cpu(), ref["grads"][(net, n)]
            c = cosine(mine, gr)
            if c < worst:
                worst, worst_name = c, f"{net}.{n}"
            flat_a.append(mine.flatten()); flat_b.append(gr.flatten())
    rep["worst_tensor_cosine"], rep["worst_tensor"] = worst, worst_name
    rep["flat_cosine"] = cosine(torch.cat(flat_a), torch.cat(flat_b))
    return rep


def log_margins(rep, name="parity_margins.jsonl"):
    """append the measured gate margins to gpurun_out/ (scratch; merged back from the GPU box) - best effort"""
    import json
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
        with open(os.path.join(root, "gpurun_out", name), "a") as f:
            f.write(json.dumps(rep) + "\n")
    except OSError:
        pass
