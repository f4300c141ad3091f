"""GPU batch evaluator of the reference's depth metrics (``new_metrics.py``).

``calc_metrics`` mirrors ``new_metrics.calc_metrics`` (:193-201) for a whole batch at once: inputs are depth maps in
millimetres (what the uint16 PNGs hold), ``hole_map = input < 50``, ``target_hole_map = target < 50`` (:14, :224-225),
values clipped to ``[0, max_depth]`` (:207-208).  Returns per-image metrics; ``mean_over_images`` applies the
reference's nan-aware average (:246-248).  All sums are taken by ``csrc/metrics.cu`` in fp64.
"""
import math

import numpy as np
import torch

from . import ops

HOLES_THRESHOLD = 50          # new_metrics.py:14
ALL_METRICS = ["rmse", "mae", "rmse_h", "rmse_d", "psnr", "ssim", "mae_h", "mae_d", "mse_v"]      # :269


def calc_metrics(pred, target, input_orig, K=None, max_depth=5100, metric_names=ALL_METRICS, device="cuda"):
    """pred / target / input_orig: (B, H, W) arrays or tensors in millimetres; K: (B, 3, 3) or (3, 3) intrinsics (needed for
    mse_v).  A target twice the size of the prediction is sub-sampled ``[0::2, 0::2]`` as in :217-218; the input must
    already have the target's size (the reference resizes it with skimage, :222).  -> {name: float64 array (B,)}."""
    def dev(a):
        t = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a)
        return t.to(device=device, dtype=torch.float32).contiguous()
    p, t = dev(pred), dev(target)
    if t.shape[-2] == 2 * p.shape[-2]:
        t = t[..., 0::2, 0::2].contiguous()
    x = dev(input_orig)
    if p.dim() == 2:
        p, t, x = p[None], t[None], x[None]
    if not (p.shape == t.shape == x.shape):
        raise ValueError(f"prediction {tuple(p.shape)}, target {tuple(t.shape)} and input {tuple(x.shape)} must have one size")
    B, H, W = p.shape
    kinv = None
    if "mse_v" in metric_names:
        if K is None:
            raise ValueError("mse_v needs the camera intrinsics K")
        Kt = torch.as_tensor(np.asarray(K), dtype=torch.float64)
        if Kt.dim() == 2:
            Kt = Kt[None].expand(B, 3, 3)
        kinv = torch.linalg.inv(Kt).reshape(B, 9).contiguous().to(device)
    out = torch.empty((B, 16), device=device, dtype=torch.float64)
    ops._call("dsr_eval_metric_sums", ops._p(p), ops._p(t), ops._p(x), ops._p(kinv, torch.float64) if kinv is not None else None,
              B, H, W, float(HOLES_THRESHOLD), float(max_depth), int("ssim" in metric_names), ops._p(out, torch.float64))
    s = out.cpu().numpy()
    nan = np.full(B, np.nan)

    def ratio(num, den):
        return np.where(den > 0, num / np.where(den > 0, den, 1), nan)
    res = {}
    for name in metric_names:
        if name == "mae":
            res[name] = ratio(s[:, 1], s[:, 0])
        elif name == "rmse":
            res[name] = np.sqrt(ratio(s[:, 2], s[:, 0]))
        elif name == "psnr":
            mse = ratio(s[:, 2], s[:, 0]) / float(max_depth) ** 2
            if np.any(mse == 0):
                raise NotImplementedError("Same img")                  # new_metrics.py:124-125
            res[name] = 20.0 * math.log10(1.0) - 10.0 * np.log10(mse)
        elif name == "mae_h":
            res[name] = ratio(s[:, 4], s[:, 3])
        elif name == "rmse_h":
            res[name] = np.sqrt(ratio(s[:, 5], s[:, 3]))
        elif name == "mae_d":
            res[name] = ratio(s[:, 7], s[:, 6])
        elif name == "rmse_d":
            res[name] = np.sqrt(ratio(s[:, 8], s[:, 6]))
        elif name == "mse_v":
            res[name] = ratio(s[:, 10], s[:, 9])
        elif name == "ssim":
            res[name] = ratio(s[:, 12], s[:, 11])
        else:
            raise KeyError(name)
    return res


def mean_over_images(per_image):
    """new_metrics.py:246-248: the mean of every metric over the images where it is defined."""
    return {k: float(np.mean(v[~np.isnan(v)])) for k, v in per_image.items()}
