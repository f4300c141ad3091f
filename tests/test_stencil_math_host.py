"""Host build of csrc/stencil_math.cuh (the header the CUDA kernels include) against the oracle:
forward values and the hand-derived analytic backward of both normal estimators.  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import ref_ops
from util import load_golden

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostcheck") / "libstencil_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", os.path.join(HERE, "hostcheck", "stencil_host.cpp"),
                    "-o", out], check=True)
    return ctypes.CDLL(out)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _cams(K, crop):
    from dsr_b200.norms import camera_table
    return camera_table(K, crop).numpy()


def test_old_normals_fwd_bwd(hostlib):
    g = load_golden("ops.npz")
    d = np.ascontiguousarray(g["d"])
    B, _, H, W = d.shape
    out = np.empty((B, 3, H, W), np.float32)
    hostlib.host_normals_old_fwd(_ptr(d), B, H, W, ctypes.c_float(1.0), _ptr(out))
    assert np.abs(out - g["normals_old"]).max() <= 1e-6
    dt = torch.from_numpy(d).requires_grad_(True)
    go = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(3))
    (ref_ops.surface_normals_old(dt) * 100 * go).sum().backward()
    gd = np.empty_like(d)
    hostlib.host_normals_old_bwd(_ptr(d), _ptr(go.numpy()), B, H, W, ctypes.c_float(100.0), _ptr(gd))
    ref = dt.grad.numpy()
    assert np.abs(gd - ref).max() <= 2e-4 * np.abs(ref).max()


def test_new_normals_fwd_bwd(hostlib):
    g = load_golden("ops.npz")
    d = np.ascontiguousarray(g["d"])
    B, _, H, W = d.shape
    cams = np.ascontiguousarray(_cams(g["K"], g["crop"]))
    out = np.empty((B, 3, H, W), np.float32)
    hostlib.host_normals_new_fwd(_ptr(d), _ptr(cams), B, H, W, _ptr(out))
    assert np.abs(out - g["normals_new"]).max() <= 1e-6
    dt = torch.from_numpy(d).requires_grad_(True)
    go = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(4))
    (ref_ops.surface_normals_new(dt, torch.from_numpy(g["K"]), torch.from_numpy(g["crop"])) * go).sum().backward()
    gd = np.empty_like(d)
    hostlib.host_normals_new_bwd(_ptr(d), _ptr(go.numpy()), _ptr(cams), B, H, W, _ptr(gd))
    ref = dt.grad.numpy()
    assert np.isfinite(ref).all() and np.isfinite(gd).all()
    # degenerate pixels (zero-length normal, clamped by 1e-12) carry 1e12-scale gradients in the
    # reference as well: compare relatively, element by element
    assert np.allclose(gd, ref, rtol=2e-3, atol=1e-3 * np.median(np.abs(ref)))


@pytest.mark.parametrize("kind", ["golden", "smooth", "steep", "skewed", "holes", "tiny"])
def test_new_normals_affine_closed_form(hostlib, kind):
    """fp32 closed form for pin-hole cameras (csrc/stencil_math.cuh aff_*) against the fp64 oracle: forward and backward."""
    g = load_golden("ops.npz")
    if kind == "golden":
        d, K, crop = np.ascontiguousarray(g["d"]), g["K"], g["crop"]
    else:
        B, H, W = 2, 37, 52
        yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
        amp = 0.002 if kind == "smooth" else 0.05
        d = np.stack([np.clip(0.1 * b + amp * (xx * (1 + b) + 0.5 * yy) + 0.05 * np.sin(xx / 5.0 + b) * np.cos(yy / 7.0), -0.95, 0.95)
                      for b in range(B)])[:, None].astype(np.float32)
        K = np.repeat(np.array([[[577.87, 0, 319.5], [0, 577.87, 239.5], [0, 0, 1]]], np.float64), B, 0)
        crop = np.array([[100, 100 + H, 200, 200 + W], [0, H, 0, W]], np.int64)
        if kind == "skewed":          # both off-diagonal terms of K^-1 non-zero (k1, k3): still affine, every coefficient in play
            K[:, 0, 1], K[:, 1, 0] = 23.0, -17.0
        if kind == "holes":           # d = -1 blocks: z = 0, zero-length normals (clamped denominator) and their neighbours
            d[:, :, 5:12, 7:20] = -1.0
            d[1, :, :, :3] = -1.0
        if kind == "tiny":            # every pixel on a border: one-sided differences only
            d, crop = np.ascontiguousarray(d[:, :, :2, :3]), np.array([[100, 102, 200, 203], [0, 2, 0, 3]], np.int64)
    B, _, H, W = d.shape
    cams = np.ascontiguousarray(_cams(K, crop))
    out = np.empty((B, 3, H, W), np.float32)
    assert hostlib.host_normals_aff_fwd(_ptr(d), _ptr(cams), B, H, W, _ptr(out)) == 1
    dt = torch.from_numpy(d).requires_grad_(True)
    ref = ref_ops.surface_normals_new(dt, torch.from_numpy(np.asarray(K)), torch.from_numpy(np.asarray(crop)))
    assert np.abs(out - ref.detach().numpy()).max() <= 2e-6
    go = torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(4))
    (ref * go).sum().backward()
    gd = np.empty_like(d)
    assert hostlib.host_normals_aff_bwd(_ptr(d), _ptr(go.numpy()), _ptr(cams), B, H, W, _ptr(gd)) == 1
    refg = dt.grad.numpy()
    assert np.isfinite(gd).all()
    assert np.allclose(gd, refg, rtol=2e-3, atol=1e-3 * np.median(np.abs(refg)))


def test_bilinear_align_corners_indices(hostlib):
    for n_in, n_out in ((256, 64), (256, 128), (640, 160), (96, 24), (7, 3)):
        x = torch.arange(n_in, dtype=torch.float32)[None, None, None, :]
        ref = torch.nn.functional.interpolate(x, size=(1, n_out), mode="bilinear", align_corners=True)[0, 0, 0]
        for o in range(n_out):
            i0, i1 = ctypes.c_int(), ctypes.c_int()
            l0, l1 = ctypes.c_float(), ctypes.c_float()
            hostlib.host_bilin_ac(o, n_out, n_in, ctypes.byref(i0), ctypes.byref(i1), ctypes.byref(l0), ctypes.byref(l1))
            v = l0.value * i0.value + l1.value * i1.value
            assert abs(v - float(ref[o])) <= 1e-3, (n_in, n_out, o)
