"""Depth -> surface normals, drop-in for the reference's ``models/norms.py``.

``SurfaceNormals`` (image-space, fp32; norms.py:185-235) and ``SurfaceNormals_new`` (camera-space,
fp64 inside, per-sample intrinsics + crop; norms.py:6-158) keep their call signatures; the math
runs in one fused stencil kernel each (forward and analytic backward), instead of ~20 / ~60 ATen
launches plus a host-built fp64 ray grid copied H2D per call.
"""
import torch
import torch.nn as nn

from . import ops


def camera_table(K, crop, shift=0.5, device=None):
    """(B, 11) float64 table the kernels read: K^-1 row-major, w0 + shift, h0 + shift.

    Mirrors norms.py:75-89,103-107 (K.inverse() in fp64, u = w0 + j + shift, v = h0 + i + shift) and
    the asserts of batch_arange (norms.py:14,16)."""
    K = torch.as_tensor(K, dtype=torch.float64).detach().cpu()
    crop = torch.as_tensor(crop).detach().cpu()
    h, h_, w, w_ = crop[:, 0], crop[:, 1], crop[:, 2], crop[:, 3]
    assert bool((h_ >= h).all()) and bool((w_ >= w).all()), "stop value should be greater or equal to start value"
    assert bool(((h_ - h) == (h_ - h)[0]).all()) and bool(((w_ - w) == (w_ - w)[0]).all()), \
        "all ranges have to be same length"
    kinv = torch.linalg.inv(K).reshape(-1, 9)
    tab = torch.cat([kinv, (w.to(torch.float64) + shift)[:, None], (h.to(torch.float64) + shift)[:, None]], dim=1)
    tab = tab.contiguous()
    if device is not None:
        tab = tab.pin_memory().to(device, non_blocking=True) if torch.cuda.is_available() else tab.to(device)
    return tab


class SurfaceNormals_new(nn.Module):
    def forward(self, depth, K, crop, depth_type="orthogonal", shift=.5):
        if depth_type != "orthogonal":
            raise ValueError(f"Unknown type {depth_type}")              # norms.py:100-101
        B, _, H, W = depth.shape
        crop_t = torch.as_tensor(crop)
        assert int(crop_t[0, 1] - crop_t[0, 0]) == H and int(crop_t[0, 3] - crop_t[0, 2]) == W, \
            "crop extent must equal the depth size"
        cams = camera_table(K, crop, shift, depth.device)
        return ops.normals_new(depth.float(), cams)


class SurfaceNormals(nn.Module):
    def forward(self, depth):
        return ops.normals_old(depth, 1.0)
