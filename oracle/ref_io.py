"""Oracle (numpy) restatement of the on-disk format conversions either side of the hot path.  TEST INFRASTRUCTURE ONLY.

Parity pin: these are three lines of numpy arithmetic each; the reference functions that hold them
(``MyUnalignedDataset.trasform`` data/my_main_dataset.py:32-52, the ``--save_all`` branch models/main_model.py:321-333) sit
inside classes that import albumentations / imageio, which this image does not have, so they are restated line by line
here and NOT executed ("parity unpinned by execution" for this file; the arithmetic is numpy's own).
"""
import numpy as np


def depth_from_u16(depth_u16, meters=5100):
    """data/my_main_dataset.py:38-42 (integer branch): ``np.where(depth > m, m, depth) / m``, ``* 2 - 1``, astype float32."""
    depth = np.asarray(depth_u16).astype(np.int32)
    depth = np.where(depth > meters, meters, depth) / meters
    depth = depth * 2 - 1
    return depth.astype(np.float32)


def image_from_u8(img_u8):
    """data/my_main_dataset.py:35-36: ``img.astype(np.float32)``, ``(img - 127.5) / 127.5``; HWC -> CHW as ToTensor does."""
    img = np.asarray(img_u8).astype(np.float32)
    img = (img - 127.5) / 127.5
    return np.ascontiguousarray(np.moveaxis(img, -1, -3))


def depth_to_u16(pred, crop_rows=16):
    """models/main_model.py:321,331-333: ``np.clip((img.permute(1,2,0).numpy() + 1) / 2, 0, 1)[:, :, 0] * 5100`` of
    ``pred[i][:, 16:-16, :]``, ``astype(np.uint16)``.  pred: (B, 1, H, W) float32."""
    pred = np.asarray(pred, dtype=np.float32)
    out = []
    for i in range(pred.shape[0]):
        img = np.transpose(pred[i][:, crop_rows:pred.shape[2] - crop_rows, :], (1, 2, 0))
        out_np = np.clip((img + 1) / 2, 0, 1)[:, :, 0] * 5100
        out.append(out_np.astype(np.uint16))
    return np.stack(out)
