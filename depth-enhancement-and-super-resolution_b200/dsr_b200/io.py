"""On-disk formats either side of the hot path (citations into /root/reference).

Input  (data/my_main_dataset.py:35-52): depth = uint16 PNG in millimetres -> ``min(d, 5100) / 5100 * 2 - 1`` (float64 in
numpy, stored as float32), image = uint8 RGB -> ``(x - 127.5) / 127.5``.
Output (models/main_model.py:321-333, ``--save_all`` in the test stage): ``clip((pred + 1) / 2, 0, 1) * 5100`` truncated
to uint16, rows ``[16, H - 16)`` (the 480 rows of a 640x480 frame inside the 512-row network input), one PNG per sample
named after the real-domain file.

The arithmetic runs on the device (csrc/resize.cu); only the PNG container is coded on the host (cv2, or PIL when cv2 is
absent - the reference uses imageio, which this image does not ship).
"""
import os

import numpy as np
import torch

from . import ops

MAX_MM = 5100        # `meters = 5100` (sic) in my_main_dataset.py:38


def depth_from_u16(d_u16, device, max_mm=MAX_MM):
    """(B, H, W) or (B, 1, H, W) uint16 millimetres (numpy / CPU tensor) -> (B, 1, H, W) float32 in [-1, 1] on `device`."""
    a = np.ascontiguousarray(np.asarray(d_u16)).astype(np.uint16, copy=False)
    if a.ndim == 3:
        a = a[:, None]
    t = torch.from_numpy(a.view(np.int16)).to(device, non_blocking=True)          # torch has no uint16 arithmetic: raw bits
    out = torch.empty(a.shape, device=device, dtype=torch.float32)
    ops._call("dsr_u16_to_depth", ops._p(t, torch.int16), t.numel(), int(max_mm), ops._p(out))
    return out


def image_from_u8(img_u8, device):
    """(B, H, W, 3) uint8 RGB as decoded from the file -> (B, 3, H, W) float32 in [-1, 1] on `device`."""
    a = np.ascontiguousarray(np.asarray(img_u8)).astype(np.uint8, copy=False)
    B, H, W, C = a.shape
    t = torch.from_numpy(a).to(device, non_blocking=True)
    out = torch.empty((B, C, H, W), device=device, dtype=torch.float32)
    ops._call("dsr_u8_to_image", ops._p(t, torch.uint8), B, H, W, C, ops._p(out))
    return out


def depth_to_u16(pred, crop_rows=16, max_mm=MAX_MM):
    """(B, 1, H, W) float32 prediction in [-1, 1] -> (B, H - 2*crop_rows, W) uint16 millimetres (numpy, host)."""
    p = ops.planes(pred.detach())
    B, _, H, W = p.shape
    out = torch.empty((B, H - 2 * crop_rows, W), device=p.device, dtype=torch.int16)
    ops._call("dsr_depth_to_u16", ops._p(p), B, H, W, int(crop_rows), float(max_mm), ops._p(out, torch.int16))
    return out.cpu().numpy().view(np.uint16)


def write_png_u16(path, arr):
    arr = np.ascontiguousarray(arr, dtype=np.uint16)
    try:
        import cv2
        if not cv2.imwrite(path, arr):
            raise IOError(f"cv2.imwrite failed for {path}")
    except ImportError:
        from PIL import Image
        Image.fromarray(arr).save(path)


def read_png_u16(path):
    try:
        import cv2
        a = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if a is None:
            raise IOError(f"cannot read {path}")
        return a
    except ImportError:
        from PIL import Image
        return np.array(Image.open(path))


def save_predictions(pred, paths, folder, crop_rows=16, max_mm=MAX_MM):
    """main_model.py:321-333: ``file = f'{save_image_folder}{basename}.png'`` for every sample of the batch."""
    arrs = depth_to_u16(pred, crop_rows, max_mm)
    files = []
    for a, path in zip(arrs, paths):
        name = str(path).split("/")[-1].split(".")[0]
        file = f"{folder}{name}.png"
        d = os.path.dirname(file)
        if d:
            os.makedirs(d, exist_ok=True)
        write_png_u16(file, a)
        files.append(file)
    return files
