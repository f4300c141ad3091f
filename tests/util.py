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
