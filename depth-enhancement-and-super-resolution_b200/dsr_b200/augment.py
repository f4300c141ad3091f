"""GPU augmentation stage: the training-time transform chain of the reference's dataset,
``MyUnalignedDataset.trasform`` (data/my_main_dataset.py:56-90) - Resize(INTER_AREA) -> Rotate(+-30 deg, p 0.9) -> RandomCrop ->
HorizontalFlip(p 0.5), or Resize -> PadIfNeeded(512, 640) -> HorizontalFlip with ``--no_aug`` / in the test stage, then
``np.clip(-1, 1)`` - applied to whole batches that are already on the device (``io.depth_from_u16`` / ``io.image_from_u8``),
so that eight B200s are not fed by per-sample OpenCV calls on the host.

The random parameters are drawn on the host from Python's ``random`` module in the call order of albumentations 0.4.6
(requirements.txt:5: ``Compose`` / ``BasicTransform.__call__`` draw once per transform for ``p``; ``Rotate.get_params``
draws ``uniform(-30, 30)``; ``RandomCrop.get_params`` draws ``h_start`` then ``w_start``) - same idea as the rectangle tables
of the training step; the pixel work (csrc/augment.cu) restates ``cv2.warpAffine``'s fixed-point bilinear scheme bit for
bit."""
import math
import random as _random

import numpy as np
import torch

from . import ops


def draw_params(load_h, load_w, crop_h, crop_w, train=True, no_aug=False, rng=_random):
    """-> dict(angle | None, top, left, flip, out_h, out_w) for ONE ``trasform`` call (image and depth share it)"""
    rng.random()                                       # Compose (p = 1)
    rng.random()                                       # Resize (p = 1)
    if not train or no_aug:                            # my_main_dataset.py:71-74, :82-84
        rng.random()                                   # PadIfNeeded (p = 1)
        flip = (rng.random() < 0.5) if train else False
        ph, pw = max(512 - load_h, 0), max(640 - load_w, 0)
        return dict(angle=None, top=-(ph // 2), left=-(pw // 2), flip=flip, out_h=load_h + ph, out_w=load_w + pw)
    angle = rng.uniform(-30, 30) if rng.random() < 0.9 else None           # :77
    rng.random()                                       # RandomCrop (p = 1)
    h_start, w_start = rng.random(), rng.random()
    top, left = int((load_h - crop_h) * h_start), int((load_w - crop_w) * w_start)
    flip = rng.random() < 0.5                          # :79
    return dict(angle=angle, top=top, left=left, flip=flip, out_h=crop_h, out_w=crop_w)


def inverse_rotation(angle_deg, w, h):
    """cv2.getRotationMatrix2D((w / 2, h / 2), angle, 1) followed by cv2.invertAffineTransform, in double -> 6 numbers"""
    a = math.radians(angle_deg)
    al, be = math.cos(a), math.sin(a)
    cx, cy = w / 2, h / 2
    m00, m01, m02, m10, m11, m12 = al, be, (1 - al) * cx - be * cy, -be, al, be * cx + (1 - al) * cy
    D = m00 * m11 - m01 * m10
    D = 1.0 / D if D != 0 else 0.0
    A11, A22, A12, A21 = m11 * D, m00 * D, -m01 * D, -m10 * D
    return [A11, A12, -A11 * m02 - A12 * m12, A21, A22, -A21 * m02 - A22 * m12]


def param_tables(params, load_h, load_w, device):
    """per-sample parameter dicts -> (minv float64 [N, 6], ipar int32 [N, 4]) on `device`"""
    minv = np.zeros((len(params), 6), dtype=np.float64)
    ipar = np.zeros((len(params), 4), dtype=np.int32)
    for n, p in enumerate(params):
        if p["angle"] is not None:
            minv[n] = inverse_rotation(p["angle"], load_w, load_h)
        ipar[n] = (int(p["angle"] is not None), p["top"], p["left"], int(p["flip"]))
    return torch.from_numpy(minv).to(device, non_blocking=True), torch.from_numpy(ipar).to(device, non_blocking=True)


def resize_area(x, height, width):
    """A.Resize(height, width, interpolation=cv2.INTER_AREA) of NCHW planes: identity or an integer down-scale factor"""
    x = ops.planes(x)
    B, C, H, W = x.shape
    if (H, W) == (height, width):
        return x
    if H % height or W % width:
        raise NotImplementedError("dsr_b200.augment: INTER_AREA resize only for integer down-scale factors (the datasets of the "
                                  "reference are stored at load_size 640 x 480 already)")
    y = torch.empty((B, C, height, width), device=x.device, dtype=torch.float32)
    ops._call("dsr_resize_area_int", ops._p(x), B * C, H, W, H // height, W // width, ops._p(y))
    return y


def augment_batch(depth, img, params, load_h, load_w):
    """depth (B, 1, H, W), img (B, 3, H, W) float32 in [-1, 1] on the device, params = [draw_params(...)] per sample ->
    (depth (B, 1, h, w), img (B, 3, h, w)) after the whole chain."""
    out_h, out_w = params[0]["out_h"], params[0]["out_w"]
    if any((p["out_h"], p["out_w"]) != (out_h, out_w) for p in params):
        raise ValueError("all samples of a batch share one output size")
    outs = []
    minv = ipar = None
    for x in (depth, img):
        x = resize_area(x, load_h, load_w)
        B, C, H, W = x.shape
        if minv is None:
            minv, ipar = param_tables(params, load_h, load_w, x.device)
        y = torch.empty((B, C, out_h, out_w), device=x.device, dtype=torch.float32)
        ops._call("dsr_augment_gather", ops._p(x), B, C, H, W, ops._p(minv, torch.float64), ops._p(ipar, torch.int32), ops._p(y),
                  out_h, out_w)
        outs.append(y)
    return outs[0], outs[1]


def crop_tables(batch, train, no_aug, crop_h, crop_w):
    """crop_A / crop_B of the batch dict (my_main_dataset.py:183-190): [0, crop_h, 0, crop_w] when training with
    augmentation, [0, 512, 0, 640] otherwise - the camera-space normals read them"""
    row = [0, crop_h, 0, crop_w] if (train and not no_aug) else [0, 512, 0, 640]
    return torch.tensor([row] * batch, dtype=torch.int64)
