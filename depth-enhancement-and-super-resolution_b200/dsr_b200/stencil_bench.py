"""HBM roofline of the bandwidth-bound kernels of the path (loss / normal stencils, masks, resize, norm layers, Adam)
at a size that cannot live in the 126 MB L2: `batch` planes of 512x640 (the full-size 640x480 frame as the nets see it).

achieved GB/s = ALGORITHMIC bytes (fp32 I/O of the op, each tensor counted once) / mean launch
time (CUDA events on the launching stream, 3 warm-up + `iters` launches back to back).  Used by bench.py
(`roofline_hbm`) and scripts/bench_stencils.py."""
import torch


def run(batch=96, iters=20, peak=6540.8, device=0, verbose=False, only=None):
    from . import _lib, ops
    from .norms import camera_table
    _lib.load()
    B, H, W = batch, 512, 640
    P = B * H * W
    dev = torch.device("cuda", device)
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: torch.rand(*s, device=dev, generator=g) * 1.8 - 0.9
    d, d2 = rnd(B, 1, H, W), rnd(B, 1, H, W)
    d[:, :, 100:140, 200:260] = -1.0
    img = rnd(B, 3, H, W)
    n3, g3 = rnd(B, 3, H, W), rnd(B, 3, H, W)
    m1 = (rnd(B, 1, H, W) > -0.5).float()
    m2 = (rnd(B, 1, H, W) > -0.5).float()
    out1, out1b, out3 = torch.empty_like(d), torch.empty_like(d), torch.empty_like(n3)
    K = torch.tensor([[577.87, 0, 319.5], [0, 577.87, 239.5], [0, 0, 1]], dtype=torch.float64).repeat(B, 1, 1)
    cams = camera_table(K, torch.tensor([[0, H, 0, W]] * B), 0.5).to(dev)
    acc = torch.zeros(8, device=dev, dtype=torch.float64)
    one = torch.ones(2, device=dev)
    rects = torch.randint(0, 400, (B, 64, 4), device=dev, dtype=torch.int32)
    rects[:, :, 2:] = rects[:, :, 2:] % 60
    counts = torch.full((B,), 40, device=dev, dtype=torch.int32)
    gt = torch.empty((B, 1, H, W), device=dev, dtype=torch.uint8)
    C = 128
    Bn = max(1, B // 16)
    act = rnd(Bn, H, W, C)                       # an NHWC activation of the 128-channel layers
    act2, actg = torch.empty_like(act), rnd(Bn, H, W, C)
    Pn = Bn * H * W
    prm = torch.empty(3 * Bn * C, device=dev)
    sums = torch.zeros(Bn * C * 2, device=dev, dtype=torch.float64)
    npar = 44_250_000
    par, grd, ma, va = (torch.zeros(npar, device=dev) for _ in range(4))
    grd.normal_(generator=g)
    hyper = torch.tensor([1e-4, 0.9, 0.999, 1e-8], device=dev, dtype=torch.float64)
    step = torch.zeros(1, device=dev, dtype=torch.int32)
    lr_h, lr_w = H // 2, W // 2
    dlr = torch.empty((B, 1, lr_h, lr_w), device=dev)
    feat_lr = rnd(Bn, lr_h, lr_w, C)
    p, p64, pu8, pi32 = ops._p, (lambda t: ops._p(t, torch.float64)), (lambda t: ops._p(t, torch.uint8)), (lambda t: ops._p(t, torch.int32))
    call = ops._call
    f4 = 4.0
    cases = [  # name, algorithmic bytes, thunk
        ("hole_valid_masks (4P -> 8P)", 12 * P, lambda: call("dsr_hole_valid_masks", p(d), B, H, W, -0.97, p(out1), p(out1b))),
        ("rect_holes (8P -> 9P)", 17 * P, lambda: call("dsr_rect_holes", p(m1), p(d), pi32(rects), pi32(counts), 64, B, H, W,
                                                       -0.97, pu8(gt), p(out1), p(out1b))),
        ("normals_old fwd (4P -> 12P)", 16 * P, lambda: call("dsr_normals_old_fwd", p(d), B, H, W, 100.0, p(out3))),
        ("normals_old bwd (16P -> 4P)", 20 * P, lambda: call("dsr_normals_old_bwd", p(d), p(g3), B, H, W, 100.0, p(out1))),
        ("normals_new fwd (4P -> 12P, closed fp32 form)", 16 * P, lambda: call("dsr_normals_new_fwd", p(d), p64(cams), B, H, W, p(out3))),
        ("normals_new bwd (16P -> 4P)", 20 * P, lambda: call("dsr_normals_new_bwd", p(d), p(g3), p64(cams), B, H, W, p(out1))),
        ("tv fwd (12P -> scalar)", 12 * P, lambda: call("dsr_tv_fwd", p(n3), B * 3, H, W, p64(acc))),
        ("tv bwd (12P -> 12P)", 24 * P, lambda: call("dsr_tv_bwd", p(n3), B * 3, H, W, p(one), 1.0, p(out3))),
        ("masked L1/L2 fwd, 3 ch, 2 masks (32P -> 2 scalars)", 32 * P,
         lambda: call("dsr_masked_diff_fwd", p(n3), p(g3), p(m1), p(m2), B, 3, H * W, p64(acc))),
        ("masked L1/L2 bwd, 3 ch, 2 masks (32P -> 12P)", 44 * P,
         lambda: call("dsr_masked_diff_bwd", p(n3), p(g3), p(m1), p(m2), B, 3, H * W, p(one), p(one) + 4, 1e-6, 1e-6, p(out3))),
        ("masked_sums (12P -> 3 scalars)", 12 * P, lambda: call("dsr_masked_sums", p(d), p(d2), p(m1), P, p64(acc))),
        ("smooth level fwd (16P -> 2 scalars)", 16 * P, lambda: call("dsr_smooth_level_fwd", p(d), p(img), B, 3, H, W, p64(acc))),
        ("smooth level bwd (16P -> 4P)", 20 * P, lambda: call("dsr_smooth_level_bwd", p(d), p(img), B, 3, H, W, p(one), 1e-6, 1e-6, p(out1), 0)),
        ("ssim fwd (8P -> scalar)", 8 * P, lambda: call("dsr_ssim_fwd", p(d), p(d2), B, H, W, p64(acc), None)),
        ("bicubic x0.5 planes (4P -> P)", 5 * P, lambda: call("dsr_bicubic_fwd", p(d), B, H, W, 1, lr_h, lr_w, p(dlr))),
        ("bicubic x2 NHWC C=128 (Pn*C -> 4 Pn*C)", (Pn // 4 + Pn) * C * f4,
         lambda: call("dsr_bicubic_fwd", p(feat_lr), Bn, lr_h, lr_w, C, H, W, p(act2))),
        ("channel_sums NHWC C=128 (4 B/elt)", Pn * C * f4, lambda: call("dsr_channel_sums", p(act), Bn, H * W, C, p64(sums))),
        ("norm_apply fwd NHWC C=128 (8 B/elt)", 2 * Pn * C * f4,
         lambda: call("dsr_norm_apply_fwd", p(act), p(prm), None, p(act2), Bn, H * W, C, 1)),
        ("in_bwd_sums NHWC C=128 (8 B/elt)", 2 * Pn * C * f4,
         lambda: call("dsr_in_bwd_sums", p(act), p(actg), p(prm), Bn, H * W, C, 1, p64(sums))),
        ("in_bwd_apply NHWC C=128 (12 B/elt)", 3 * Pn * C * f4,
         lambda: call("dsr_in_bwd_apply", p(act), p(actg), p(prm), p64(sums), p(act2), Bn, H * W, C, 1)),
        ("adam over the 44.25 M-parameter arena (28 B/param)", 28.0 * npar,
         lambda: call("dsr_adam_step_dev", p(par), p(grd), p(ma), p(va), npar, p64(hyper), pi32(step), 1.0)),
    ]
    call("dsr_channel_sums", p(act), Bn, H * W, C, p64(sums))
    call("dsr_norm_finalize", p64(sums), Bn, C, H * W, 0, None, None, 1e-5, p(prm))
    rows = []
    for name, nbytes, fn in cases:
        if only and not any(o in name for o in only):
            continue
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / iters
        gbs = nbytes / (us * 1e-6) / 1e9
        rows.append(dict(kernel=name, algorithmic_mbytes=round(nbytes / 1e6, 1), us=round(us, 1), gbs=round(gbs, 1), frac=round(gbs / peak, 3)))
        if verbose:
            print(f"{name:58s} {nbytes / 1e6:9.1f} MB {us:9.1f} us {gbs:8.1f} GB/s {100 * gbs / peak:5.1f}% of {peak:.0f}")
    res = dict(shape=dict(planes=B, H=H, W=W, nhwc_images=Bn, C=C), peak_gbs=peak, iters=iters, rows=rows,
               note="inputs far larger than the 126 MB L2 (one fp32 plane set = %.0f MB); back-to-back launches of the same "
                    "kernel, CUDA events on the launching stream" % (4 * P / 1e6))
    return res
