"""Oracle (numpy) restatement of the training-time augmentation of ``MyUnalignedDataset.trasform``
(data/my_main_dataset.py:56-90): Resize(INTER_AREA) -> Rotate(+-30 deg, p = 0.9) -> RandomCrop -> HorizontalFlip(p = 0.5)
(or Resize -> PadIfNeeded(512, 640) -> HorizontalFlip with --no_aug / in the test stage), then clip to [-1, 1].

TEST INFRASTRUCTURE ONLY.  The transforms live in two third-party dependencies that /root/reference does not vendor:
albumentations 0.4.6 (requirements.txt:5; NOT in this image) and OpenCV (requirements.txt:6; cv2 IS in this image).

* The pixel work is OpenCV's: ``cv2.warpAffine(img, getRotationMatrix2D((w/2, h/2), angle, 1), (w, h), INTER_LINEAR,
  BORDER_REFLECT_101)`` (albumentations/augmentations/functional.py ``rotate``), ``cv2.resize(INTER_AREA)``,
  ``cv2.copyMakeBorder(BORDER_REFLECT_101)``.  ``warp_affine`` below restates warpAffine's published fixed-point scheme
  (coordinates in 1/1024 px, rounded to 1/32 px, a 32 x 32 table of float bilinear weights) and is PINNED against
  cv2.warpAffine itself by tests/test_augment.py (index work bit-exact: identical source taps and table entries).
* The random draws (``draw_params``) restate albumentations 0.4.6 ``Compose`` / ``BasicTransform.__call__`` /
  ``Rotate.get_params`` / ``RandomCrop.get_params`` over Python's ``random`` module, call by call.  That library is absent,
  so this order is "parity unpinned" (restated from its published source, not executed).
"""
import math
import random as _random

import numpy as np

AB_BITS, INTER_BITS = 10, 5
AB_SCALE, INTER_TAB = 1 << AB_BITS, 1 << INTER_BITS


def rotation_matrix(angle_deg, w, h):
    """cv2.getRotationMatrix2D((w / 2, h / 2), angle, 1.0) (2 x 3, float64) - forward map src -> dst"""
    a = math.radians(angle_deg)
    alpha, beta = math.cos(a), math.sin(a)
    cx, cy = w / 2, h / 2
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy], [-beta, alpha, beta * cx + (1 - alpha) * cy]], dtype=np.float64)


def invert_affine(M):
    """cv2.invertAffineTransform in double, as warpAffine does when WARP_INVERSE_MAP is not set"""
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22, A12, A21 = M[1, 1] * D, M[0, 0] * D, -M[0, 1] * D, -M[1, 0] * D
    b1 = -A11 * M[0, 2] - A12 * M[1, 2]
    b2 = -A21 * M[0, 2] - A22 * M[1, 2]
    return np.array([[A11, A12, b1], [A21, A22, b2]], dtype=np.float64)


def reflect101(p, n):
    """cv2.borderInterpolate(p, n, BORDER_REFLECT_101)"""
    p = np.asarray(p).copy()
    if n == 1:
        return np.zeros_like(p)
    for _ in range(64):
        lo, hi = p < 0, p >= n
        if not (lo.any() or hi.any()):
            break
        p = np.where(lo, -p, p)
        p = np.where(p >= n, 2 * (n - 1) - p, p)
    return p


def warp_coords(Minv, w, h):
    """warpAffine's fixed-point source coordinates of every destination pixel -> (sx, sy, ax, ay): integer tap origin and the
    1/32-pixel fractions (imgwarp.cpp WarpAffineInvoker: adelta / bdelta per column, X0 / Y0 per row, round_delta = 16)."""
    x = np.arange(w, dtype=np.float64)
    y = np.arange(h, dtype=np.float64)
    sat = lambda v: np.rint(v).astype(np.int64)                      # saturate_cast<int>(double) = cvRound (half to even)
    adelta, bdelta = sat(Minv[0, 0] * x * AB_SCALE), sat(Minv[1, 0] * x * AB_SCALE)
    rd = AB_SCALE // INTER_TAB // 2
    X0 = sat((Minv[0, 1] * y + Minv[0, 2]) * AB_SCALE) + rd
    Y0 = sat((Minv[1, 1] * y + Minv[1, 2]) * AB_SCALE) + rd
    X = (X0[:, None] + adelta[None, :]) >> (AB_BITS - INTER_BITS)
    Y = (Y0[:, None] + bdelta[None, :]) >> (AB_BITS - INTER_BITS)
    return X >> INTER_BITS, Y >> INTER_BITS, X & (INTER_TAB - 1), Y & (INTER_TAB - 1)


def bilinear_tab():
    """initInterTab2D(INTER_LINEAR): float weights w[ay][ax] = (vy0 vx0, vy0 vx1, vy1 vx0, vy1 vx1), v(f) = (1 - f, f), f = a / 32"""
    f = (np.arange(INTER_TAB, dtype=np.float32) * np.float32(1.0 / INTER_TAB)).astype(np.float32)
    v = np.stack([np.float32(1.0) - f, f], 1)                        # [32][2]
    return (v[:, None, :, None] * v[None, :, None, :]).astype(np.float32).reshape(INTER_TAB, INTER_TAB, 4)


def warp_affine(img, M, border_reflect101=True):
    """cv2.warpAffine(img, M, (w, h), flags=INTER_LINEAR, borderMode=BORDER_REFLECT_101) for float32 (H, W) or (H, W, C)"""
    img = np.asarray(img, dtype=np.float32)
    h, w = img.shape[:2]
    sx, sy, ax, ay = warp_coords(invert_affine(np.asarray(M, dtype=np.float64)), w, h)
    tab = bilinear_tab()[ay, ax]                                     # (h, w, 4)
    x0, x1, y0, y1 = reflect101(sx, w), reflect101(sx + 1, w), reflect101(sy, h), reflect101(sy + 1, h)
    im = img if img.ndim == 3 else img[:, :, None]
    out = (im[y0, x0] * tab[..., 0:1] + im[y0, x1] * tab[..., 1:2] + im[y1, x0] * tab[..., 2:3] + im[y1, x1] * tab[..., 3:4])
    out = out.astype(np.float32)
    return out if img.ndim == 3 else out[:, :, 0]


def resize_area(img, height, width):
    """cv2.resize(img, (width, height), interpolation=INTER_AREA): identity for equal sizes, the box mean of resizeAreaFast
    for integer down-scale factors (float32 running sum over the box rows, then * 1 / area); other ratios are outside the
    path (ScanNet / InteriorNet frames are stored at load_size 640 x 480 already)."""
    img = np.asarray(img, dtype=np.float32)
    h, w = img.shape[:2]
    if (h, w) == (height, width):
        return img
    if h % height or w % width:
        raise NotImplementedError("resize_area: only identity and integer down-scale factors")
    fy, fx = h // height, w // width
    acc = np.zeros((height, width) + img.shape[2:], dtype=np.float32)
    for dy in range(fy):
        for dx in range(fx):
            acc = (acc + img[dy::fy, dx::fx]).astype(np.float32)
    return (acc * np.float32(1.0 / (fy * fx))).astype(np.float32)


def draw_params(load_h, load_w, crop_h, crop_w, train=True, no_aug=False, rng=_random):
    """The Python-``random`` draws of ONE ``trasform`` call in albumentations 0.4.6 order: Compose (p = 1) draws once; every
    transform draws once for its own ``p``; Rotate draws ``uniform(-30, 30)`` when it fires; RandomCrop draws h_start then
    w_start.  -> dict(angle or None, top, left, flip, pad) (crop offsets as RandomCrop computes them: int((H - h) * start))."""
    rng.random()                                                     # Compose.__call__: random.random() < self.p
    rng.random()                                                     # Resize (p = 1)
    if not train or no_aug:
        rng.random()                                                 # PadIfNeeded (p = 1)
        flip = (rng.random() < 0.5) if train else False              # test stage has no flip in the list
        ph, pw = max(512 - load_h, 0), max(640 - load_w, 0)
        return dict(angle=None, top=-(ph // 2), left=-(pw // 2), flip=flip, out_h=load_h + ph, out_w=load_w + pw)
    angle = rng.uniform(-30, 30) if rng.random() < 0.9 else None
    rng.random()                                                     # RandomCrop (p = 1)
    h_start, w_start = rng.random(), rng.random()
    top, left = int((load_h - crop_h) * h_start), int((load_w - crop_w) * w_start)
    flip = rng.random() < 0.5
    return dict(angle=angle, top=top, left=left, flip=flip, out_h=crop_h, out_w=crop_w)


def augment(depth, img, p, load_h, load_w):
    """One sample through the chain.  depth (H, W) float32 in [-1, 1], img (H, W, 3) float32 -> (1, h, w), (3, h, w)."""
    out = []
    for a in (depth, img):
        a = resize_area(a, load_h, load_w)
        if p["angle"] is not None:
            a = warp_affine(a, rotation_matrix(p["angle"], load_w, load_h))
        ys = reflect101(np.arange(p["out_h"]) + p["top"], load_h)    # crop (top, left >= 0) or reflect-101 padding (< 0)
        xs = reflect101(np.arange(p["out_w"]) + p["left"], load_w)
        a = a[ys][:, xs]
        if p["flip"]:
            a = a[:, ::-1]
        a = np.clip(a, -1, 1).astype(np.float32)
        out.append(a[None] if a.ndim == 2 else np.ascontiguousarray(np.moveaxis(a, -1, 0)))
    return out[0], out[1]
