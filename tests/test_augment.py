"""GPU augmentation stage (data/my_main_dataset.py:56-90): the numpy oracle against OpenCV itself
(CPU), the CUDA stage against the oracle (GPU, bit-exact), and the host parameter stream of the product against the oracle's."""
import random

import numpy as np
import pytest
import torch

from oracle import ref_augment as ra


def _frames(rng, h, w):
    depth = (rng.rand(h, w).astype(np.float32) * 2 - 1)
    depth[rng.rand(h, w) < 0.05] = -1.0
    return depth, (rng.rand(h, w, 3).astype(np.float32) * 2 - 1)


def test_oracle_warp_is_bit_exact_against_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    for (h, w) in ((480, 640), (37, 53), (64, 64)):
        depth, img = _frames(rng, h, w)
        for ang in (-30.0, -11.3, 0.0, 7.77, 29.999):
            M = cv2.getRotationMatrix2D((w / 2, h / 2), ang, 1.0)
            assert np.array_equal(M, ra.rotation_matrix(ang, w, h))
            for a in (depth, img):
                ref = cv2.warpAffine(a, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
                assert np.array_equal(ref, ra.warp_affine(a, M)), (h, w, ang)
    big = rng.rand(960, 1280, 3).astype(np.float32)
    assert np.array_equal(cv2.resize(big, (640, 480), interpolation=cv2.INTER_AREA), ra.resize_area(big, 480, 640))
    d = rng.rand(480, 640).astype(np.float32)
    assert np.array_equal(cv2.copyMakeBorder(d, 16, 16, 0, 0, cv2.BORDER_REFLECT_101), d[ra.reflect101(np.arange(512) - 16, 480)])


def test_host_parameter_stream_equals_the_oracle():
    from dsr_b200 import augment
    for kw in (dict(train=True, no_aug=False), dict(train=True, no_aug=True), dict(train=False, no_aug=False)):
        random.seed(11)
        a = [ra.draw_params(480, 640, 256, 256, **kw) for _ in range(20)]
        sa = random.random()
        random.seed(11)
        b = [augment.draw_params(480, 640, 256, 256, **kw) for _ in range(20)]
        assert a == b and sa == random.random()
    p = a[0]
    assert p["angle"] is None and (p["top"], p["left"], p["out_h"], p["out_w"]) == (-16, 0, 512, 640) and p["flip"] is False
    random.seed(3)
    ps = [augment.draw_params(480, 640, 256, 256) for _ in range(400)]
    rot = sum(p["angle"] is not None for p in ps) / 400.0
    assert 0.84 < rot < 0.96 and 0.4 < sum(p["flip"] for p in ps) / 400.0 < 0.6
    assert all(0 <= p["top"] <= 224 and 0 <= p["left"] <= 384 and (p["angle"] is None or -30 <= p["angle"] <= 30) for p in ps)
    for ang in (-30.0, 12.5):
        assert np.array_equal(np.array(augment.inverse_rotation(ang, 640, 480)).reshape(2, 3), ra.invert_affine(ra.rotation_matrix(ang, 640, 480)))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["train", "no_aug", "test", "downscale"])
def test_gpu_stage_bit_exact_against_oracle(built_lib, mode):
    from dsr_b200 import augment
    rng = np.random.RandomState(5)
    B = 5
    h, w = (960, 1280) if mode == "downscale" else (480, 640)
    frames = [_frames(rng, h, w) for _ in range(B)]
    kw = dict(train=mode != "test", no_aug=mode == "no_aug")
    random.seed(7)
    params = [augment.draw_params(480, 640, 256, 320, **kw) for _ in range(B)]
    if mode in ("train", "downscale"):
        params[1] = dict(params[1], angle=None)                      # the p = 0.1 branch: no rotation
        params[2] = dict(params[2], angle=-30.0, flip=True, top=0, left=0)
        params[3] = dict(params[3], angle=29.5, flip=False, top=224, left=320)
    depth = torch.from_numpy(np.stack([f[0] for f in frames])[:, None]).cuda()
    img = torch.from_numpy(np.stack([np.moveaxis(f[1], -1, 0) for f in frames])).cuda()
    d_out, i_out = augment.augment_batch(depth, img, params, 480, 640)
    for n in range(B):
        rd, ri = ra.augment(frames[n][0], frames[n][1], params[n], 480, 640)
        assert np.array_equal(d_out[n].cpu().numpy(), rd), (mode, n)
        assert np.array_equal(i_out[n].cpu().numpy(), ri), (mode, n)
    assert tuple(d_out.shape[2:]) == ((256, 320) if mode in ("train", "downscale") else (512, 640))
    assert float(d_out.abs().max()) <= 1.0 and float(i_out.abs().max()) <= 1.0
    assert augment.crop_tables(B, kw["train"], kw["no_aug"], 256, 320).tolist()[0] == ([0, 256, 0, 320] if mode in ("train", "downscale") else [0, 512, 0, 640])
