"""Oracle (CPU) restatement of the depth-derived ops of the hot path.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  Plain CPU torch / numpy, written from the
behaviour of the reference (citations are into ``/root/reference``); pinned against the live
reference through ``tests/golden`` (see ``tests/golden/make_golden.py``).

All tensors are NCHW, float32 unless stated, exactly as in the reference.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

BORDER = -0.97  # models/main_model.py:133  (self.border)


# --------------------------------------------------------------------------------------------
# masks
# --------------------------------------------------------------------------------------------
def hole_valid_masks(depth, border=BORDER):
    """hole = 1[d <= border]; valid = 1[3x3 box-dilation(hole) < 1].

    models/main_model.py:208-230: four sequential in-place shifted adds on a clone of the hole
    mask.  Row pass: r1[i] = h[i] + h[i+1]; r2[i] = r1[i] + r1[i-1] - i.e. non-zero iff one of
    h[i-1], h[i], h[i+1] is set; the column pass does the same on r2, so the support of the
    result is the 3x3 dilation of ``hole`` (clipped at the image border).  ``valid`` thresholds
    it with ``< 1``.  Returns (hole, valid) as float32 {0,1}.
    """
    d = depth.detach()
    hole = (d <= border)
    h = hole.numpy()
    v = h.copy()
    v[:, :, :-1, :] |= h[:, :, 1:, :]
    v[:, :, 1:, :] |= h[:, :, :-1, :]
    w = v.copy()
    w[:, :, :, :-1] |= v[:, :, :, 1:]
    w[:, :, :, 1:] |= v[:, :, :, :-1]
    valid = torch.from_numpy(~w).to(torch.float32)
    return hole.to(torch.float32), valid


def draw_rects(batch, H, W, stage="train", rng=np.random, p_train=0.90, div=8):
    """Host RNG stream of the random rectangle holes for ONE of the two loops.

    models/main_model.py:257-273 (real) and :278-294 (syn): per sample, in this order:
    randint(10, n) -> choice(W, number, replace=False) -> choice(H, number, replace=False) ->
    randint(W//150, W//8, number) * binomial(1, p) -> randint(H//150, H//8, number) *
    binomial(1, p); n, p = (60, 0.9) for 'train', (11, 0) otherwise.
    models/main_sr_model.py:298-305, :319-326: same stream with ``// 10`` instead of ``// 8`` and
    p = 0.95 for the real loop (``p_train`` / ``div``).
    Returns a list (len batch) of int arrays (number, 4) = [x, y, size_x, size_y].
    """
    n = 60 if stage == "train" else 11
    p = p_train if stage == "train" else 0
    out = []
    for _ in range(batch):
        number = rng.randint(10, n)
        xs = rng.choice(W, number, replace=False)
        ys = rng.choice(H, number, replace=False)
        sx = rng.randint(W // 150, W // div, number) * rng.binomial(1, p)
        sy = rng.randint(H // 150, H // div, number) * rng.binomial(1, p)
        out.append(np.stack([xs, ys, sx, sy], axis=1).astype(np.int64))
    return out


def rect_gt_mask(valid, rects):
    """gt_mask (int64 {0,1}, (B,1,H,W)): 0 where the pixel is valid AND covered by a rectangle.

    models/main_model.py:268-276: ``ones[y:y+s_y, x:x+s_x] = 0`` for each rectangle (numpy slicing
    clips at the border; an empty slice when a size is 0), then
    ``np.where((valid > 0.05) & (ones < 0.05), 0, 1)``.
    """
    B, _, H, W = valid.shape
    v = valid.detach().numpy()
    out = np.ones((B, 1, H, W), dtype=np.int64)
    for i in range(B):
        cover = np.zeros((H, W), dtype=bool)
        for x, y, sx, sy in rects[i]:
            cover[y:y + sy, x:x + sx] = True
        out[i, 0] = np.where((v[i, 0] > 0.05) & cover, 0, 1)
    return torch.from_numpy(out)


def apply_gt_mask(depth, gt_mask):
    """``where(gt_mask < 0.05, -1, depth)``  (models/main_model.py:276, :298)."""
    return torch.where(gt_mask < 0.05, torch.tensor(-1.0), depth)


# --------------------------------------------------------------------------------------------
# normals
# --------------------------------------------------------------------------------------------
def _grad_axis(f, axis):
    """np.gradient-style derivative with unit spacing: central /2 inside, one-sided at the ends.
    models/norms.py:115-158 (and :192-235 for the fp32 copy)."""
    n = f.shape[axis]
    idx = lambda s: tuple([slice(None)] * axis + [s] + [slice(None)] * (f.dim() - axis - 1))
    inner = (f[idx(slice(2, None))] - f[idx(slice(None, -2))]) / 2.0
    first = f[idx(slice(1, 2))] - f[idx(slice(0, 1))]
    last = f[idx(slice(n - 1, n))] - f[idx(slice(n - 2, n - 1))]
    return torch.cat([first, inner, last], dim=axis)


def surface_normals_old(depth):
    """Image-space normals: n = (-dd/dH, -dd/dW, 1) / (||.||_2 + 1e-6).  models/norms.py:185-190.
    The caller multiplies by 100 (models/main_model.py:345-349)."""
    if depth.dtype != torch.float32:
        raise TypeError("Input shold be torch.float32")  # norms.py:205-208
    gh = -_grad_axis(depth, 2)
    gw = -_grad_axis(depth, 3)
    v = torch.cat([gh, gw, torch.ones_like(depth)], dim=1)
    nrm = torch.sqrt((v * v).sum(dim=1, keepdim=True))
    return v / (nrm + 1e-6)


def surface_normals_new(depth, K, crop, shift=0.5):
    """Camera-space normals from a pin-hole back projection, float64 inside, float32 out.

    models/norms.py:103-108 (forward), :75-101 (batch_pc), :29-73 (pc_to_normals):
    z = (d+1)/2; ray = K^-1 [u, v, 1]^T with u = w0 + j + shift, v = h0 + i + shift; the ray is
    divided by its z component and multiplied by z; derivatives along u (W) and v (H) as in
    ``_grad_axis``; n = dP/dv x dP/du; F.normalize (eps 1e-12).
    """
    B, _, H, W = depth.shape
    z = (depth.to(torch.float64) + 1.0) / 2.0
    K = torch.as_tensor(K, dtype=torch.float64)
    crop = torch.as_tensor(crop)
    h0 = crop[:, 0].to(torch.float64)
    h1 = crop[:, 1].to(torch.float64)
    w0 = crop[:, 2].to(torch.float64)
    w1 = crop[:, 3].to(torch.float64)
    assert bool((h1 >= h0).all()) and bool((w1 >= w0).all())        # norms.py:14
    nh = torch.div(h1 - h0, 1.0, rounding_mode="floor")
    nw = torch.div(w1 - w0, 1.0, rounding_mode="floor")
    assert bool((nh == nh[0]).all()) and bool((nw == nw[0]).all())  # norms.py:16
    vv = h0[:, None] + torch.arange(int(nh[0]), dtype=torch.float64)[None] + shift   # (B, H)
    uu = w0[:, None] + torch.arange(int(nw[0]), dtype=torch.float64)[None] + shift   # (B, W)
    v = vv[:, :, None].expand(-1, -1, uu.shape[1])
    u = uu[:, None, :].expand(-1, vv.shape[1], -1)
    pix = torch.stack([u, v, torch.ones_like(u)], dim=1)                             # (B,3,H,W)
    rays = torch.einsum("blk,bkij->blij", torch.linalg.inv(K), pix)
    rays = rays / rays[:, 2:3]
    P = rays * z
    Xu, Yu, Zu = (_grad_axis(P[:, c], 2) for c in range(3))     # axis 2 of (B,H,W) = W = u
    Xv, Yv, Zv = (_grad_axis(P[:, c], 1) for c in range(3))     # axis 1 = H = v
    nx = Yv * Zu - Yu * Zv
    ny = Zv * Xu - Zu * Xv
    nz = Xv * Yu - Xu * Yv
    n = torch.stack([nx, ny, nz], dim=1)
    # F.normalize (norms.py:72): v / max(||v||_2, 1e-12); vector_norm has a zero sub-gradient at 0
    n = n / torch.linalg.vector_norm(n, dim=1, keepdim=True).clamp_min(1e-12)
    return n.to(torch.float32)


# --------------------------------------------------------------------------------------------
# scalar losses
# --------------------------------------------------------------------------------------------
def tv_loss(x):
    """Sum (not mean) of squared forward differences along W and H.  models/main_model.py:15-19."""
    dw = x[:, :, :, :-1] - x[:, :, :, 1:]
    dh = x[:, :, :-1, :] - x[:, :, 1:, :]
    return (dh * dh).sum() + (dw * dw).sum()


def l1_mean(a, b):
    """torch.nn.L1Loss (mean over all elements).  models/main_model.py:171."""
    return (a - b).abs().mean()


def mse_mean(a, b):
    """torch.nn.MSELoss.  models/main_model.py:172."""
    d = a - b
    return (d * d).mean()


def _bilinear_ac(img, nh, nw):
    """Bilinear resize with align_corners=True (F.upsample in models/main_model.py:34)."""
    return F.interpolate(img, size=(nh, nw), mode="bilinear", align_corners=True)


def smooth_loss(depth, image, num_scales=3):
    """Edge-aware smoothness.  models/main_model.py:22-73.

    Pyramid = [H/4, H/2, H] (index 0 is the COARSEST level); 'x' differences run along H and 'y'
    along W (main_model.py:41-48); weight = exp(-mean_c |grad I|); term_i = mean|grad d * w| / 2^i.
    """
    H, W = depth.shape[2], depth.shape[3]
    total = 0.0
    levels = list(range(num_scales - 1, -1, -1))     # ratio exponents, coarsest first
    for i, e in enumerate(levels):
        if e == 0:
            d, im = depth, image
        else:
            d = _bilinear_ac(depth, H // 2 ** e, W // 2 ** e)
            im = _bilinear_ac(image, H // 2 ** e, W // 2 ** e)
        dgx = d[:, :, :-1, :] - d[:, :, 1:, :]
        dgy = d[:, :, :, :-1] - d[:, :, :, 1:]
        igx = im[:, :, :-1, :] - im[:, :, 1:, :]
        igy = im[:, :, :, :-1] - im[:, :, :, 1:]
        wx = torch.exp(-igx.abs().mean(dim=1, keepdim=True))
        wy = torch.exp(-igy.abs().mean(dim=1, keepdim=True))
        total = total + (dgx * wx).abs().mean() / 2 ** i + (dgy * wy).abs().mean() / 2 ** i
    return total


def ssim(img1, img2, window_size=11, sigma=1.5):
    """Gaussian-window SSIM, zero 'same' padding, mean over everything.
    models/pytorch_ssim/__init__.py:7-37 (gaussian :7-9, window :11-15, map :17-37)."""
    C = img1.shape[1]
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2))
                      for x in range(window_size)])
    g = g / g.sum()
    w2 = (g[:, None] @ g[None, :]).float()
    win = w2.expand(C, 1, window_size, window_size).contiguous()
    pad = window_size // 2
    blur = lambda t: F.conv2d(t, win, padding=pad, groups=C)
    mu1, mu2 = blur(img1), blur(img2)
    s1 = blur(img1 * img1) - mu1 * mu1
    s2 = blur(img2 * img2) - mu2 * mu2
    s12 = blur(img1 * img2) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return m.mean()


# --------------------------------------------------------------------------------------------
# loss assembly  (models/main_model.py:340-417)
# --------------------------------------------------------------------------------------------
DEFAULT_WEIGHTS = dict(w_syn_l1=15.0, w_real_l1_d=40.0, w_real_l1_i=0.1, w_syn_norm=2.0,
                       w_smooth=1.0, w_syn_holes=800.0, w_real_holes=1600.0, scale_G=1.0)


def loss_stack(t, w=None):
    """All named loss terms and loss_G from the tensors of one forward pass.

    ``t`` needs: syn_depth, real_depth, syn_mask, real_mask, real_hole_mask, gt_mask_syn,
    gt_mask_real, syn2real_depth_masked, real_depth_by_image, pred_syn_depth, pred_real_depth,
    real_image, K_A, K_B, crop_A, crop_B.   Returns (loss_G, dict of terms, dict of visuals).
    """
    w = dict(DEFAULT_WEIGHTS, **(w or {}))
    L, V = {}, {}
    ms, mr, hr = t["syn_mask"], t["real_mask"], t["real_hole_mask"]
    ps, pr = t["pred_syn_depth"], t["pred_real_depth"]
    sd, rd = t["syn_depth"], t["real_depth"]
    s2r = t["syn2real_depth_masked"]

    # old (image-space) normals x100   main_model.py:343-352
    n_s = surface_normals_old(sd) * 100
    n_sp = surface_normals_old(ps) * 100
    n_rp = surface_normals_old(pr) * 100
    L["tv_syn_norm_old"] = tv_loss(n_sp) * (10 ** -7)
    L["tv_real_norm_old"] = tv_loss(n_rp) * (10 ** -7)
    L["syn_norms_old"] = mse_mean(n_s * ms, n_sp * ms)

    # extra-hole masks   main_model.py:354-357, :396
    a_s = ((s2r < BORDER) | (t["gt_mask_syn"] < 0.1)).to(torch.float32)
    a_r = torch.where(t["gt_mask_real"] > 0.1, torch.tensor(0.0), torch.tensor(1.0))
    V["a_s"], V["a_r"] = a_s, a_r

    # camera-space normals   main_model.py:360-372
    N_s = surface_normals_new(sd, t["K_A"], t["crop_A"])
    N_s2r = surface_normals_new(s2r, t["K_A"], t["crop_A"])
    N_sp = surface_normals_new(ps, t["K_A"], t["crop_A"])
    N_r = surface_normals_new(rd, t["K_B"], t["crop_B"])
    N_rp = surface_normals_new(pr, t["K_B"], t["crop_B"])
    V.update(norm_syn=N_s, norm_syn2real=N_s2r, norm_syn_pred=N_sp, norm_real=N_r, norm_real_pred=N_rp)
    L["tv_syn_norm"] = tv_loss(N_sp) * (10 ** -7)
    L["tv_real_norm"] = tv_loss(N_rp) * (10 ** -7)
    L["syn_norms"] = l1_mean(N_s * ms, N_sp * ms)
    L["syn_norms_holes"] = l1_mean(N_s * ms * a_s, N_sp * ms * a_s)

    # depth terms   main_model.py:383-390
    L["holes_syn"] = l1_mean(sd * ms * a_s, ps * ms * a_s)
    L["holes_syn_l2"] = mse_mean(sd * ms * a_s, ps * ms * a_s) * 5
    L["task_syn"] = l1_mean(sd * ms, ps * ms)
    L["task_real_by_depth"] = l1_mean(rd * mr, pr * mr)
    L["task_real_by_image"] = l1_mean(t["real_depth_by_image"] * hr, pr * hr)

    G = (L["task_syn"] * w["w_syn_l1"] + L["holes_syn"] * w["w_syn_holes"]
         + w["w_syn_holes"] * L["holes_syn_l2"] + L["task_real_by_depth"] * w["w_real_l1_d"]
         + L["task_real_by_image"] * w["w_real_l1_i"] + L["tv_syn_norm"] * 1
         + L["syn_norms_holes"] * w["w_syn_norm"] * 5 + L["tv_real_norm"] * 1
         + L["syn_norms_old"] * w["w_syn_norm"] + L["tv_real_norm_old"] * 1
         + L["tv_syn_norm_old"] * 1)                                     # main_model.py:393
    L["holes_real"] = l1_mean(rd * a_r, pr * a_r)                       # :397
    L["holes_real_l2"] = mse_mean(rd * a_r, pr * a_r) * 5               # :398
    G = G + L["holes_real"] * w["w_real_holes"] + L["holes_real_l2"] * w["w_real_holes"]   # :399
    G = G + L["syn_norms"] * w["w_syn_norm"]                            # :404
    L["smooth"] = smooth_loss(pr, t["real_image"], 3)                   # :407
    G = G + L["smooth"] * w["w_smooth"]                                 # :408
    G = G * w["scale_G"]                                                # :417
    return G, L, V


# --------------------------------------------------------------------------------------------
# super-resolution step  (models/main_sr_model.py)
# --------------------------------------------------------------------------------------------
def bicubic(x, size):
    """The reference's own call: F.interpolate(x, size, mode='bicubic').  models/main_sr_model.py:279-293, :361,
    :368-372, :396-398.  Used by the SR step oracle so that it follows the reference bit for bit: the HR nets amplify
    a 1e-6 resampling difference to 6e-3 on the bottleneck weight gradients (measured with bicubic_restated)."""
    return F.interpolate(x, size=tuple(size), mode="bicubic")


def bicubic_restated(x, size):
    """What that call computes (align_corners=False; cubic convolution A = -0.75 over border-clamped taps, source
    coordinate (o + 0.5) * in/out - 0.5), restated tap by tap with no call into torch's interpolate; pinned against
    tests/golden/resize.npz and used to check the CUDA kernel."""
    A = -0.75

    def taps(n_in, n_out):
        o = torch.arange(n_out, dtype=torch.float32)
        r = (float(n_in) / float(n_out)) * (o + 0.5) - 0.5
        f = torch.floor(r)
        t = r - f
        c1 = lambda v: ((A + 2) * v - (A + 3)) * v * v + 1
        c2 = lambda v: ((A * v - 5 * A) * v + 8 * A) * v - 4 * A
        wts = torch.stack([c2(t + 1), c1(t), c1(1 - t), c2(2 - t)], 1)                 # (n_out, 4)
        idx = (f.long()[:, None] + torch.arange(-1, 3)[None, :]).clamp(0, n_in - 1)      # (n_out, 4)
        return idx, wts

    B, C, H, W = x.shape
    iy, wy = taps(H, size[0])
    ix, wx = taps(W, size[1])
    rows = (x[:, :, :, ix] * wx.to(x.dtype)).sum(-1)                     # (B, C, H, Wo)
    return (rows[:, :, iy, :] * wy.to(x.dtype)[None, None, :, :, None]).sum(3)        # (B, C, Ho, Wo)


def nearest(x, size):
    """F.interpolate(x, size, mode='nearest'): src = min(floor(dst * in/out), in - 1).
    models/main_sr_model.py:394-395, :452, :459."""
    H, W = x.shape[2], x.shape[3]
    iy = torch.floor(torch.arange(size[0], dtype=torch.float32) * (float(H) / size[0])).long().clamp(max=H - 1)
    ix = torch.floor(torch.arange(size[1], dtype=torch.float32) * (float(W) / size[1])).long().clamp(max=W - 1)
    return x[:, :, iy][:, :, :, ix]


SR_DEFAULT_WEIGHTS = dict(w_syn_l1=15.0, w_real_l1_d=90.0, w_real_l1_i=0.1, w_syn_norm=3.0,
                          w_smooth=1.0, w_syn_holes=1600.0, w_real_holes=1600.0, scale_G=1.0)   # README.md:86


def loss_stack_sr(t, lr_size, w=None):
    """Loss stack of the SR step, models/main_sr_model.py:391-482.  ``t`` as for loss_stack plus
    pred_real_depth_hr (HR) and pred_real_depth (= bicubic x0.5 of it); real_* tensors at HR on entry.
    Returns (loss_G, terms, visuals)."""
    w = dict(SR_DEFAULT_WEIGHTS, **(w or {}))
    L, V = {}, {}
    h, wd = lr_size
    mr = nearest(t["real_mask"], (h, wd))                                # :394
    hr = nearest(t["real_hole_mask"], (h, wd))                           # :395
    rd = bicubic(t["real_depth"], (h, wd))                               # :396
    ri = bicubic(t["real_image"], (h, wd))                               # :397
    V["real_depth_by_image"] = bicubic(t["real_depth_by_image"], (h, wd))   # :398
    V.update(real_mask=mr, real_hole_mask=hr, real_depth=rd, real_image=ri)
    ms = t["syn_mask"]
    ps, pr, pr_hr = t["pred_syn_depth"], t["pred_real_depth"], t["pred_real_depth_hr"]
    sd, s2r = t["syn_depth"], t["syn2real_depth_masked"]
    n_s = surface_normals_old(sd) * 100                                  # :402-410
    n_sp = surface_normals_old(ps) * 100
    n_rp_hr = surface_normals_old(pr_hr) * 100
    L["tv_syn_norm_old"] = tv_loss(n_sp) * (10 ** -7)
    L["tv_real_norm_old"] = tv_loss(n_rp_hr) * (10 ** -7)
    L["syn_norms_old"] = l1_mean(n_s, n_sp)
    a_s = ((s2r < BORDER) | (t["gt_mask_syn"] < 0.1)).to(torch.float32)  # :412-415
    a_r = nearest(torch.where(t["gt_mask_real"] > 0.1, torch.tensor(0.0), torch.tensor(1.0)), (h, wd))   # :458-459
    V["a_s"], V["a_r"] = a_s, a_r
    N_s = surface_normals_new(sd, t["K_A"], t["crop_A"])                 # :425-431
    N_s2r = surface_normals_new(s2r, t["K_A"], t["crop_A"])
    N_sp = surface_normals_new(ps, t["K_A"], t["crop_A"])
    N_r = surface_normals_new(rd, t["K_B"], t["crop_B"])
    N_rp = surface_normals_new(pr, t["K_B"], t["crop_B"])
    N_rp_hr = surface_normals_new(pr_hr, t["K_A"], t["crop_A"])
    V.update(norm_syn=N_s, norm_syn2real=N_s2r, norm_syn_pred=N_sp, norm_real=N_r, norm_real_pred=N_rp,
             norm_real_pred_hr=N_rp_hr)
    L["tv_syn_norm"] = tv_loss(N_sp) * (10 ** -7)
    L["tv_real_norm"] = tv_loss(N_rp) * (10 ** -7)
    L["syn_norms"] = mse_mean(N_s * ms, N_rp_hr * ms)                    # :434
    L["syn_norms_holes"] = l1_mean(N_s * ms * a_s, N_sp * ms * a_s)      # :435
    L["holes_syn"] = l1_mean(sd * ms * a_s, ps * ms * a_s)               # :445-452
    L["holes_syn_l2"] = mse_mean(sd * ms * a_s, ps * ms * a_s) * 5
    L["task_syn"] = l1_mean(sd * ms, ps * ms)
    L["task_real_by_depth"] = l1_mean(rd * mr, pr * mr)
    L["task_real_by_image"] = l1_mean(nearest(sd, (h, wd)) * hr, pr * hr)
    G = (L["task_syn"] * w["w_syn_l1"] + L["holes_syn"] * w["w_syn_holes"]
         + w["w_syn_holes"] * L["holes_syn_l2"] + L["task_real_by_depth"] * w["w_real_l1_d"]
         + L["task_real_by_image"] * w["w_real_l1_i"] + L["tv_syn_norm"] * 1
         + L["syn_norms_holes"] * w["w_syn_norm"] * 5 + L["tv_real_norm"] * 2
         + L["syn_norms_old"] * w["w_syn_norm"] * 5 + L["tv_real_norm_old"] * 2
         + L["tv_syn_norm_old"] * 1)                                     # :455
    L["holes_real"] = l1_mean(rd * a_r, pr * a_r)                       # :460
    L["holes_real_l2"] = mse_mean(rd * a_r, pr * a_r) * 5               # :461
    G = G + L["holes_real"] * w["w_real_holes"] + L["holes_real_l2"] * w["w_real_holes"]   # :462
    G = G + L["syn_norms"] * w["w_syn_norm"]                            # :469
    L["smooth"] = smooth_loss(pr, ri, 3)                                # :472
    G = G + L["smooth"] * w["w_smooth"]
    G = G * w["scale_G"]                                                # :482
    return G, L, V


def monitor_scalars_sr(t, lr_size):
    """models/main_sr_model.py:362-372 (real side: bicubic-resized depth AND mask)."""
    out = {}
    d, m, p = t["syn_depth"], t["syn_mask"], t["pred_syn_depth"].detach()
    out["syn_mean_diff"] = float((d * m).mean() - (p * m).mean())
    out["mean_of_abs_diff_syn"] = float((d * m - p * m).abs().mean())
    d, m, p = bicubic(t["real_depth"], lr_size), bicubic(t["real_mask"], lr_size), t["pred_real_depth"].detach()
    out["real_mean_diff"] = float((d * m).mean() - (p * m).mean())
    out["mean_of_abs_diff_real"] = float((d * m - p * m).abs().mean())
    return out


def monitor_scalars(t):
    """syn/real mean differences (models/main_model.py:308-318)."""
    out = {}
    for dom in ("syn", "real"):
        d, m, p = t[f"{dom}_depth"], t[f"{dom}_mask"], t[f"pred_{dom}_depth"].detach()
        out[f"{dom}_mean_diff"] = float((d * m).mean() - (p * m).mean())
        out[f"mean_of_abs_diff_{dom}"] = float((d * m - p * m).abs().mean())
    return out


# --------------------------------------------------------------------------------------------
# Adam  (torch.optim.Adam defaults, models/main_model.py:176: betas (0.9, 0.999), eps 1e-8, wd 0)
# --------------------------------------------------------------------------------------------
def adam_update(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """One Adam update in place on (p, m, v); ``step`` is the 1-based step count."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)
