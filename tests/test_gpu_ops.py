"""GPU parity of every op of the C-ABI library against the oracle / a CPU fp32 torch reference,
on seeded inputs.  Integer-valued work (masks) must be BIT-EXACT; floating point within the
tolerance written at each assert.  Run on the B200 box: python -m pytest tests -m gpu."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_ops
from util import load_golden, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(built_lib):
    from dsr_b200 import ops as o
    assert torch.cuda.is_available()
    return o


def G(seed):
    return torch.Generator().manual_seed(seed)


def cl(x):
    """CPU NCHW tensor -> CUDA tensor with NHWC memory (what the nets pass around)."""
    return x.cuda().contiguous(memory_format=torch.channels_last)


# ---------------------------------------------------------------------------------------------
# masks (bit-exact)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 1, 40, 56), (1, 1, 2, 2), (3, 1, 128, 128), (6, 1, 256, 256), (1, 1, 7, 301)])
def test_hole_valid_masks_bit_exact(ops, shape):
    d = torch.rand(shape, generator=G(1)) * 2 - 1
    d[torch.rand(shape, generator=G(2)) < 0.07] = -1.0
    d[0, 0, 0, 0] = -0.97
    d[0, 0, -1, -1] = -0.9700001
    hole_r, valid_r = ref_ops.hole_valid_masks(d)
    hole, valid = ops.hole_valid_masks(d.cuda())
    assert torch.equal(hole.cpu(), hole_r) and torch.equal(valid.cpu(), valid_r)


def test_hole_valid_masks_golden_and_degenerate(ops):
    g = load_golden("ops.npz")
    hole, valid = ops.hole_valid_masks(torch.from_numpy(g["d"]).cuda())
    assert np.array_equal(hole.cpu().numpy().astype(np.uint8), g["hole"])
    assert np.array_equal(valid.cpu().numpy().astype(np.uint8), g["valid"])
    for fill, v in ((-1.0, 0.0), (0.3, 1.0)):                  # all-hole / no-hole
        d = torch.full((2, 1, 16, 24), fill)
        hole, valid = ops.hole_valid_masks(d.cuda())
        assert float(valid.min()) == v and float(valid.max()) == v and float(hole.max()) == 1.0 - v


@pytest.mark.parametrize("stage,HW", [("train", (128, 128)), ("train", (256, 256)), ("test", (128, 160)), ("train", (512, 640))])
def test_rect_holes_bit_exact(ops, stage, HW):
    from dsr_b200 import main_model
    H, W = HW
    B = 3
    d = torch.rand(B, 1, H, W, generator=G(3)) * 1.6 - 0.8
    d[torch.rand(B, 1, H, W, generator=G(4)) < 0.05] = -1.0
    _, valid = ref_ops.hole_valid_masks(d)
    np.random.seed(11)
    rects_o = ref_ops.draw_rects(B, H, W, stage)
    np.random.seed(11)
    rects, counts = main_model.draw_rects(B, H, W, stage)       # product host RNG == oracle RNG stream
    for i in range(B):
        assert counts[i] == len(rects_o[i]) and np.array_equal(rects[i, :counts[i]], rects_o[i])
    gt_r = ref_ops.rect_gt_mask(valid, rects_o)
    masked_r = ref_ops.apply_gt_mask(d, gt_r)
    extra_r = ((masked_r < -0.97) | (gt_r < 0.1)).float()
    gt, masked, extra = ops.rect_holes(valid.cuda(), d.cuda(), torch.from_numpy(rects).cuda(),
                                       torch.from_numpy(counts).cuda(), main_model.MAX_RECTS, extra_border=-0.97)
    assert torch.equal(gt.cpu().to(torch.int64), gt_r)
    assert torch.equal(masked.cpu(), masked_r) and torch.equal(extra.cpu(), extra_r)
    _, _, extra2 = ops.rect_holes(valid.cuda(), d.cuda(), torch.from_numpy(rects).cuda(),
                                  torch.from_numpy(counts).cuda(), main_model.MAX_RECTS)
    assert torch.equal(extra2.cpu(), torch.where(gt_r > 0.1, torch.tensor(0.0), torch.tensor(1.0)))


# ---------------------------------------------------------------------------------------------
# normals / tv / masked losses / smoothness / ssim
# ---------------------------------------------------------------------------------------------
def _smooth_depth(B, H, W, seed):
    from oracle.ref_step import synthetic_batch
    return synthetic_batch(B, H, W, seed=seed, depth_kind="smooth")["A_d"]


@pytest.mark.parametrize("HW", [(40, 56), (128, 128), (2, 2)])
def test_normals_old_fwd_bwd(ops, HW):
    H, W = HW
    d = (torch.rand(2, 1, H, W, generator=G(5)) * 1.8 - 0.9).requires_grad_(True)
    go = torch.randn(2, 3, H, W, generator=G(6))
    ref = ref_ops.surface_normals_old(d) * 100
    (ref * go).sum().backward()
    dc = d.detach().cuda().requires_grad_(True)
    out = ops.normals_old(dc, 100.0)
    (out * go.cuda()).sum().backward()
    assert (out.cpu() - ref.detach()).abs().max() <= 1e-4            # values are O(100)
    assert rel_l2(dc.grad.cpu(), d.grad) <= 1e-4


def test_normals_new_fwd_bwd_golden(ops):
    from dsr_b200.norms import SurfaceNormals_new, camera_table
    g = load_golden("ops.npz")
    d = torch.from_numpy(g["d"])
    K, crop = torch.from_numpy(g["K"]), torch.from_numpy(g["crop"])
    out = SurfaceNormals_new()(d.cuda(), K, crop)
    assert (out.cpu().numpy() - g["normals_new"]).__abs__().max() <= 1e-6
    # backward on a non-degenerate depth map
    d2 = _smooth_depth(2, 48, 64, 9).clamp_min(-0.9).requires_grad_(True)
    go = torch.randn(2, 3, 48, 64, generator=G(7))
    crop2 = torch.tensor([[10, 58, 20, 84], [0, 48, 0, 64]])
    (ref_ops.surface_normals_new(d2, K, crop2) * go).sum().backward()
    dc = d2.detach().cuda().requires_grad_(True)
    (ops.normals_new(dc, camera_table(K, crop2, 0.5, "cuda")) * go.cuda()).sum().backward()
    assert np.allclose(dc.grad.cpu().numpy(), d2.grad.numpy(), rtol=2e-3, atol=1e-3 * float(d2.grad.abs().median()))


def test_normals_new_mixed_cameras(ops):
    """one pin-hole camera (closed fp32 form, normals_aff_*) and one with a perspective row in K (generic fp64 kernels) in the
    SAME batch: every plane picks its kernel on the device; both against the fp64 oracle, forward and backward."""
    from dsr_b200.norms import camera_table
    H, W = 40, 72
    K = torch.tensor([[[577.87, 0, 319.5], [0, 577.87, 239.5], [0, 0, 1]], [[577.87, 0.3, 319.5], [0.1, 570.0, 239.5], [1e-5, 2e-5, 1]],
                      [[600.0, 0, 320], [0, 600.0, 240], [0, 0, 1]]], dtype=torch.float64)
    crop = torch.tensor([[10, 10 + H, 20, 20 + W], [0, H, 0, W], [100, 100 + H, 300, 300 + W]])
    assert float(torch.linalg.inv(K)[1, 2, :2].abs().max()) > 0          # sample 1 really is non-affine
    d = _smooth_depth(3, H, W, 21).clamp_min(-0.9).requires_grad_(True)
    go = torch.randn(3, 3, H, W, generator=G(22))
    ref = ref_ops.surface_normals_new(d, K, crop)
    (ref * go).sum().backward()
    dc = d.detach().cuda().requires_grad_(True)
    out = ops.normals_new(dc, camera_table(K, crop, 0.5, "cuda"))
    (out * go.cuda()).sum().backward()
    assert float((out.detach().cpu() - ref.detach()).abs().max()) <= 2e-6
    assert np.allclose(dc.grad.cpu().numpy(), d.grad.numpy(), rtol=2e-3, atol=1e-3 * float(d.grad.abs().median()))


# warp-strip / row-band rolling kernels (csrc/stencil_roll.cu): shapes with several 128-column strips per row, a last strip that
# is only partly filled, one-strip rows, and chunks of rows that do not divide H - all against the CPU oracle
ROLL_SHAPES = [(3, 24, 640), (2, 33, 260), (1, 70, 132), (2, 9, 1024), (2, 17, 4), (1, 300, 128)]


@pytest.mark.parametrize("BHW", ROLL_SHAPES)
def test_rolling_normals_and_tv_wide_rows(ops, BHW):
    from dsr_b200.norms import camera_table
    B, H, W = BHW
    d = _smooth_depth(B, max(H, 16), max(W, 16), 31)[:, :, :H, :W].contiguous().clamp_min(-0.9)
    d = (d + 0.02 * torch.randn(B, 1, H, W, generator=G(32))).requires_grad_(True)
    go = torch.randn(B, 3, H, W, generator=G(33))
    K = torch.tensor([[577.87, 0, 319.5], [0, 577.87, 239.5], [0, 0, 1]], dtype=torch.float64).repeat(B, 1, 1)
    crop = torch.tensor([[7, 7 + H, 11, 11 + W]] * B)
    # image-space normals
    ref = ref_ops.surface_normals_old(d) * 100
    (ref * go).sum().backward()
    g_ref, d.grad = d.grad.clone(), None
    dc = d.detach().cuda().requires_grad_(True)
    out = ops.normals_old(dc, 100.0)
    (out * go.cuda()).sum().backward()
    assert (out.detach().cpu() - ref.detach()).abs().max() <= 1e-4
    assert rel_l2(dc.grad.cpu(), g_ref) <= 1e-4
    # camera-space normals
    ref = ref_ops.surface_normals_new(d, K, crop)
    (ref * go).sum().backward()
    g_ref, d.grad = d.grad.clone(), None
    dc = d.detach().cuda().requires_grad_(True)
    out = ops.normals_new(dc, camera_table(K, crop, 0.5, "cuda"))
    (out * go.cuda()).sum().backward()
    assert float((out.detach().cpu() - ref.detach()).abs().max()) <= 2e-6
    assert np.allclose(dc.grad.cpu().numpy(), g_ref.numpy(), rtol=2e-3, atol=1e-3 * float(g_ref.abs().median()))
    # TV
    x = torch.randn(B, 3, H, W, generator=G(34))
    assert abs(float(ops.tv_loss(x.cuda())) - float(ref_ops.tv_loss(x))) <= 1e-5 * float(ref_ops.tv_loss(x))


def test_tv_fwd_bwd(ops):
    x = torch.randn(2, 3, 37, 53, generator=G(8)).requires_grad_(True)
    ref = ref_ops.tv_loss(x)
    ref.backward()
    xc = x.detach().cuda().requires_grad_(True)
    out = ops.tv_loss(xc)
    (out * 1.0).backward()
    assert abs(float(out) - float(ref)) <= 1e-5 * float(ref)
    assert rel_l2(xc.grad.cpu(), x.grad) <= 1e-6


@pytest.mark.parametrize("C,two", [(1, False), (1, True), (3, True)])
def test_masked_l1_l2_fwd_bwd(ops, C, two):
    a = torch.randn(2, C, 33, 47, generator=G(9))
    b = torch.randn(2, C, 33, 47, generator=G(10)).requires_grad_(True)
    m1 = (torch.rand(2, 1, 33, 47, generator=G(11)) < 0.7).float()
    m2 = (torch.rand(2, 1, 33, 47, generator=G(12)) < 0.5).float() if two else None
    am, bm = (a * m1, b * m1) if m2 is None else (a * m1 * m2, b * m1 * m2)
    l1, l2 = ref_ops.l1_mean(am, bm), ref_ops.mse_mean(am, bm)
    (3 * l1 + 7 * l2).backward()
    bc = b.detach().cuda().requires_grad_(True)
    out = ops.masked_l1_l2(a.cuda(), bc, m1.cuda(), m2.cuda() if two else None)
    (3 * out[0] + 7 * out[1]).backward()
    assert abs(float(out[0]) - float(l1)) <= 1e-5 * float(l1) and abs(float(out[1]) - float(l2)) <= 1e-5 * float(l2)
    assert rel_l2(bc.grad.cpu(), b.grad) <= 1e-5


@pytest.mark.parametrize("HW", [(64, 96), (128, 128), (256, 320)])
def test_smooth_fwd_bwd(ops, HW):
    H, W = HW
    d = (torch.rand(2, 1, H, W, generator=G(13)) * 2 - 1).requires_grad_(True)
    img = torch.rand(2, 3, H, W, generator=G(14)) * 2 - 1
    ref = ref_ops.smooth_loss(d, img, 3)
    ref.backward()
    dc = d.detach().cuda().requires_grad_(True)
    out = ops.smooth_loss(dc, img.cuda(), 3)
    out.backward()
    assert abs(float(out) - float(ref)) <= 1e-3 * float(ref)         # north_star: each loss within 1e-3 relative
    assert abs(float(out) - float(ref)) <= 2e-5 * float(ref)         # what fp32 actually achieves
    assert rel_l2(dc.grad.cpu(), d.grad) <= 1e-3


def test_smooth_golden(ops):
    g = load_golden("ops.npz")
    out = ops.smooth_loss(torch.from_numpy(g["smooth_d"]).cuda(), torch.from_numpy(g["smooth_img"]).cuda(), 3)
    assert abs(float(out) - float(g["smooth"])) <= 2e-5 * float(g["smooth"])


def test_ssim(ops):
    g = load_golden("ops.npz")
    out = ops.ssim(torch.from_numpy(g["ssim_a"]).cuda(), torch.from_numpy(g["ssim_b"]).cuda())
    assert abs(float(out) - float(g["ssim"])) <= 1e-5
    a = torch.rand(1, 1, 70, 100, generator=G(15))
    assert abs(float(ops.ssim(a.cuda(), a.cuda())) - 1.0) <= 1e-5     # identity property
    b = torch.rand(1, 1, 70, 100, generator=G(16))
    assert abs(float(ops.ssim(a.cuda(), b.cuda())) - float(ref_ops.ssim(a, b))) <= 1e-5


def test_masked_sums(ops):
    d, p = torch.randn(2, 1, 31, 45, generator=G(17)), torch.randn(2, 1, 31, 45, generator=G(18))
    m = (torch.rand(2, 1, 31, 45, generator=G(19)) < 0.6).float()
    s = ops.masked_sums(d.cuda(), p.cuda(), m.cuda()).cpu()
    ref = torch.stack([(d * m).double().sum(), (p * m).double().sum(), (d * m - p * m).abs().double().sum()])
    assert torch.allclose(s, ref, rtol=1e-6)


# ---------------------------------------------------------------------------------------------
# network plumbing
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C", [1, 3, 32, 261])
def test_layout_roundtrip(ops, C):
    x = torch.randn(2, C, 19, 23, generator=G(20))
    xh = ops.nhwc(x.cuda())                                           # NCHW-contiguous in -> transpose kernel
    assert torch.equal(xh.cpu(), x.permute(0, 2, 3, 1).contiguous())
    back = ops.planes(ops.nchw(xh))
    assert torch.equal(back.cpu(), x)


@pytest.mark.parametrize("mode,p", [("reflect", 3), ("reflect", 1), ("replicate", 3), ("replicate", 1), ("zeros", 2)])
def test_pad_fwd_bwd(ops, mode, p):
    x = torch.randn(2, 5, 9, 11, generator=G(21)).requires_grad_(True)
    ref = F.pad(x, (p, p, p, p), mode={"zeros": "constant"}.get(mode, mode))
    go = torch.randn(ref.shape, generator=G(22))
    (ref * go).sum().backward()
    xc = cl(x.detach()).requires_grad_(True)
    out = ops.pad2d(xc, p, mode)
    (out * go.cuda()).sum().backward()
    assert torch.equal(out.cpu(), ref.detach())
    assert torch.allclose(xc.grad.cpu(), x.grad, atol=1e-6)


def test_activations_and_cat(ops):
    x = torch.randn(2, 8, 7, 9, generator=G(23)).requires_grad_(True)
    go = torch.randn(2, 8, 7, 9, generator=G(24))
    for fn, ref_fn in ((ops.relu, F.relu), (lambda t: ops.leaky_relu(t, 0.2), lambda t: F.leaky_relu(t, 0.2)),
                       (ops.tanh, torch.tanh)):
        x.grad = None
        r = ref_fn(x)
        (r * go).sum().backward()
        xc = cl(x.detach()).requires_grad_(True)
        o = fn(xc)
        (o * go.cuda()).sum().backward()
        assert torch.allclose(o.cpu(), r.detach(), atol=1e-6) and torch.allclose(xc.grad.cpu(), x.grad, atol=1e-6)
    a, b, c = (torch.randn(2, n, 5, 6, generator=G(25 + n)) for n in (128, 1, 3))
    ac, bc, cc = cl(a).requires_grad_(True), cl(b), cl(c).requires_grad_(True)
    o = ops.cat([ac, bc, cc])
    assert torch.equal(o.cpu(), torch.cat([a, b, c], 1))
    go = torch.randn(o.shape, generator=G(30))
    (o * go.cuda()).sum().backward()
    assert torch.equal(ac.grad.cpu(), go[:, :128]) and torch.equal(cc.grad.cpu(), go[:, 129:])


@pytest.mark.parametrize("shape,act,res", [((2, 32, 16, 16), 1, False), ((3, 128, 8, 8), 0, True),
                                           ((2, 512, 2, 2), 0, False), ((1, 64, 33, 17), 1, False)])
def test_instance_norm_fwd_bwd(ops, shape, act, res):
    x = (torch.randn(shape, generator=G(31)) * 2 + 0.5).requires_grad_(True)
    r = torch.randn(shape, generator=G(32)).requires_grad_(True) if res else None
    ref = F.instance_norm(x, eps=1e-5)
    if act:
        ref = F.relu(ref)
    if res:
        ref = ref + r
    go = torch.randn(shape, generator=G(33))
    (ref * go).sum().backward()
    xc = cl(x.detach()).requires_grad_(True)
    rc = cl(r.detach()).requires_grad_(True) if res else None
    out = ops.instance_norm(xc, 1e-5, act, rc)
    (out * go.cuda()).sum().backward()
    assert torch.allclose(out.cpu(), ref.detach(), atol=2e-5)
    assert rel_l2(xc.grad.cpu(), x.grad) <= 1e-4
    if res:
        assert torch.allclose(rc.grad.cpu(), r.grad, atol=1e-6)


def test_group_norm_fwd(ops):
    x = torch.randn(2, 64, 12, 10, generator=G(34)) * 3 + 1
    w, b = torch.randn(64, generator=G(35)), torch.randn(64, generator=G(36))
    ref = F.relu(F.group_norm(x, 8, w, b, eps=1e-5))
    with torch.no_grad():
        out = ops.group_norm(cl(x), 8, w.cuda(), b.cuda(), 1e-5, 1)
    assert torch.allclose(out.cpu(), ref, atol=3e-5)          # (the backward: tests/test_translation_blocks.py)


# ---------------------------------------------------------------------------------------------
# convolutions: every layer-shape class of the five nets, small spatial sizes
# ---------------------------------------------------------------------------------------------
CONV_CASES = [  # Cin, Cout, k, stride, pad, H, W
    (3, 32, 7, 1, 0, 22, 22), (2, 32, 7, 1, 0, 22, 22), (32, 128, 7, 1, 0, 22, 22), (64, 1, 7, 1, 0, 22, 22),
    (32, 64, 3, 2, 1, 16, 16), (128, 128, 3, 1, 0, 18, 18), (256, 256, 3, 1, 0, 10, 10),
    (261, 64, 4, 2, 1, 16, 16), (128, 64, 4, 2, 1, 16, 16), (512, 512, 4, 2, 1, 4, 4), (512, 512, 4, 2, 1, 2, 2),
    (1, 32, 7, 1, 0, 22, 22), (32, 64, 4, 2, 1, 18, 18),
]


# (engine, passes, operand dtype, forward rel-L2 tolerance): what each precision mode is good for
ENGINES = [("simt", 3, "f16", 1e-5), ("tc", 3, "f16", 1e-5), ("tc", 1, "f16", 1e-3), ("tc", 3, "bf16", 5e-5),
           ("tc", 2, "bf16", 4e-3), ("tc", 1, "bf16", 8e-3)]


@pytest.fixture(params=ENGINES, ids=lambda e: f"{e[0]}{e[1]}{e[2]}")
def engine(request, ops):
    old = dict(ops.CONFIG)
    # wgrad_passes follows `passes` here: these op-level tolerances are those of the full-precision weight gradient;
    # the default single-pass weight gradient is covered by test_tc_wgrad_large[1-...] and by the step-level gates
    ops.CONFIG.update(engine=request.param[0], passes=request.param[1], dtype=request.param[2], wgrad_passes=request.param[1],
                      big_hw=0)
    # (engine, passes, forward tolerance, dgrad tolerance: the backward operands are bf16 hi/lo)
    xtol = 1e-5 if request.param[0] == "simt" else {3: 5e-5, 2: 4e-3, 1: 8e-3}[request.param[1]]
    yield (request.param[0], request.param[1], request.param[3], xtol)
    ops.CONFIG.update(old)


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_fwd_bwd(ops, case, engine):
    Ci, Co, k, s, p, H, W = case
    ftol = engine[2]
    x = torch.randn(2, Ci, H, W, generator=G(40)).requires_grad_(True)
    w = (torch.randn(Co, Ci, k, k, generator=G(41)) * 0.05).requires_grad_(True)
    b = torch.randn(Co, generator=G(42)).requires_grad_(True)
    ref = F.conv2d(x, w, b, stride=s, padding=p)
    go = torch.randn(ref.shape, generator=G(43))
    (ref * go).sum().backward()
    xc, wc, bc = cl(x.detach()).requires_grad_(True), w.detach().cuda().requires_grad_(True), b.detach().cuda().requires_grad_(True)
    out = ops.conv2d(xc, wc, bc, s, p)
    (out * go.cuda()).sum().backward()
    assert rel_l2(out.cpu(), ref.detach()) <= ftol
    assert rel_l2(xc.grad.cpu(), x.grad) <= engine[3] and rel_l2(wc.grad.cpu(), w.grad) <= max(1e-5, engine[3])
    assert rel_l2(bc.grad.cpu(), b.grad) <= 1e-5


@pytest.mark.parametrize("mode,k,Ci,Co,H,W", [("reflect", 7, 32, 128, 24, 40), ("reflect", 3, 128, 128, 16, 16),
                                              ("replicate", 3, 256, 256, 16, 16), ("replicate", 7, 64, 1, 32, 32),
                                              ("replicate", 4, 32, 64, 32, 32), ("reflect", 3, 128, 128, 20, 12),
                                              ("reflect", 4, 64, 128, 20, 12), ("replicate", 4, 3, 64, 16, 24),
                                              ("replicate", 4, 64, 128, 2, 2), ("reflect", 4, 32, 32, 2, 4)])
def test_conv2d_fused_padding_modes(ops, engine, mode, k, Ci, Co, H, W):
    """pad module folded into the conv (ReflectionPad2d + Conv2d, padding_mode='replicate')."""
    stride = 2 if k == 4 else 1
    p = 1 if k == 4 else k // 2
    x = torch.randn(3, Ci, H, W, generator=G(60)).requires_grad_(True)
    w = (torch.randn(Co, Ci, k, k, generator=G(61)) * 0.05).requires_grad_(True)
    b = torch.randn(Co, generator=G(62))
    ref = torch.tanh(F.conv2d(F.pad(x, (p, p, p, p), mode=mode), w, b, stride=stride))
    go = torch.randn(ref.shape, generator=G(63))
    (ref * go).sum().backward()
    xc, wc = cl(x.detach()).requires_grad_(True), w.detach().cuda().requires_grad_(True)
    out = ops.conv2d(xc, wc, b.cuda(), stride, p, act_out=ops.ACT_TANH, pad_mode=mode)
    (out * go.cuda()).sum().backward()
    assert rel_l2(out.cpu(), ref.detach()) <= engine[2]
    btol = max(2e-5, 2 * engine[2])        # the tanh backward reads the forward output
    assert rel_l2(xc.grad.cpu(), x.grad) <= max(btol, engine[3]) and rel_l2(wc.grad.cpu(), w.grad) <= max(btol, engine[3])


@pytest.mark.parametrize("mode,pad,H,W", [("zeros", 3, 32, 64), ("reflect", 3, 24, 40), ("reflect", 3, 40, 64), ("zeros", 0, 38, 70),
                                          ("replicate", 3, 17, 33)])
def test_dgrad_group_forms_agree(ops, mode, pad, H, W):
    """data gradient of the 7x7 32 -> 128 head (networks.py:378, :413) in its three GEMM forms - plain rows (N = 32), pixel
    pairs (N = 64, conv_tc2) and pixel quads (128 GEMM rows on the channel-major kernel, output rows rounded up to whole
    groups and folded by dsr_pad2d_bwd_pitch) - against the fp32 reference and against each other, weight gradient included
    (it reads the widened zero-padded dY operand the grouped forms leave behind)."""
    old = dict(ops.CONFIG)
    try:
        x = torch.randn(2, 32, H, W, generator=G(160)).requires_grad_(True)
        w = (torch.randn(128, 32, 7, 7, generator=G(161)) * 0.05).requires_grad_(True)
        b = torch.randn(128, generator=G(162)).requires_grad_(True)
        xp = F.pad(x, (pad,) * 4, mode={"zeros": "constant"}.get(mode, mode)) if pad else x
        ref = F.conv2d(xp, w, b)
        go = torch.randn(ref.shape, generator=G(163))
        (ref * go).sum().backward()
        got = {}
        for name, cfg in (("plain", dict(dgrad_pair=False)), ("pair", dict(dgrad_pair=True, dgrad_quad=False)),
                          ("quad", dict(dgrad_pair=True, dgrad_quad=True))):
            ops.CONFIG.update(old)
            ops.CONFIG.update(engine="tc", passes=3, dtype="f16", wgrad_passes=3, big_hw=0, **cfg)
            xc, wc, bc = cl(x.detach()).requires_grad_(True), w.detach().cuda().requires_grad_(True), b.detach().cuda().requires_grad_(True)
            out = ops.conv2d(xc, wc, bc, 1, pad, pad_mode=mode)
            (out * go.cuda()).sum().backward()
            got[name] = (xc.grad.cpu(), wc.grad.cpu(), bc.grad.cpu())
            assert rel_l2(got[name][0], x.grad) <= 5e-5, name
            assert rel_l2(got[name][1], w.grad) <= 5e-5 and rel_l2(got[name][2], b.grad) <= 1e-5, name
        assert rel_l2(got["quad"][0], got["plain"][0]) <= 2e-5 and rel_l2(got["pair"][0], got["plain"][0]) <= 2e-5
    finally:
        ops.CONFIG.update(old)


@pytest.mark.parametrize("kind,Ci,Co,k,stride,pad,mode,H,W", [
    ("conv", 128, 256, 3, 1, 1, "reflect", 32, 48), ("conv", 64, 128, 4, 2, 1, "zeros", 64, 64), ("conv", 32, 128, 7, 1, 3, "reflect", 32, 64),
    ("convT", 256, 128, 4, 2, 1, "zeros", 16, 24), ("conv", 192, 128, 3, 1, 1, "replicate", 33, 21)])
def test_tc3_single_pass_k64_stages(ops, kind, Ci, Co, k, stride, pad, mode, H, W):
    """single-pass channel-major GEMM with 64-channel pipeline stages (128-byte rows, SWIZZLE_128B: conv_tc3_kernel<1, 64>) against
    the 32-channel form (DSR_TC3_KB=32, same 16-bit operands, only the fp32 accumulation order differs) and the fp32 reference;
    forward and data gradient (the 32 -> 128 7x7 case runs the quad form, Ca = 512, 21 taps)."""
    import os
    old = dict(ops.CONFIG)
    try:
        ops.CONFIG.update(engine="tc", passes=1, dtype="f16", bwd_dtype="bf16", wgrad_passes=1, big_hw=0, halo_min_tiles=0)
        x = torch.randn(2, Ci, H, W, generator=G(170)).requires_grad_(True)
        if kind == "conv":
            w = (torch.randn(Co, Ci, k, k, generator=G(171)) * 0.05).requires_grad_(True)
            xp = F.pad(x, (pad,) * 4, mode={"zeros": "constant"}.get(mode, mode))
            ref = F.conv2d(xp, w, None, stride=stride)
        else:
            w = (torch.randn(Ci, Co, k, k, generator=G(171)) * 0.05).requires_grad_(True)
            ref = F.conv_transpose2d(x, w, None, stride=stride, padding=pad)
        go = torch.randn(ref.shape, generator=G(173))
        (ref * go).sum().backward()
        got = {}
        for kb in ("32", "64"):
            os.environ["DSR_TC3_KB"] = kb
            xc, wc = cl(x.detach()).requires_grad_(True), w.detach().cuda().requires_grad_(True)
            with _CallLog() as names:
                if kind == "conv":
                    out = ops.conv2d(xc, wc, None, stride, pad, pad_mode=mode)
                else:
                    out = ops.conv_transpose2d(xc, wc, None, stride, pad, 0)
                (out * go.cuda()).sum().backward()
            assert "dsr_tc_gemm3" in names, "the shape must reach the channel-major kernel"
            got[kb] = (out.detach().cpu(), xc.grad.cpu())
            assert rel_l2(got[kb][0], ref.detach()) <= 1e-3 and rel_l2(got[kb][1], x.grad) <= 8e-3
        assert rel_l2(got["64"][0], got["32"][0]) <= 1e-6 and rel_l2(got["64"][1], got["32"][1]) <= 1e-6
    finally:
        os.environ.pop("DSR_TC3_KB", None)
        ops.CONFIG.update(old)


class _CallLog:
    """records the names of the library calls made inside the `with` block"""

    def __enter__(self):
        from dsr_b200 import _lib
        self.lib, self.orig, self.names = _lib, _lib.call, []

        def logged(name, *a):
            self.names.append(name)
            return self.orig(name, *a)
        _lib.call = logged
        return self.names

    def __exit__(self, *exc):
        self.lib.call = self.orig
        return False


CONVT_CASES = [  # Cin, Cout, k, stride, pad, opad, H, W
    (128, 64, 3, 2, 1, 1, 8, 8), (64, 32, 3, 2, 1, 1, 9, 7), (512, 512, 4, 2, 1, 0, 2, 2), (1024, 256, 4, 2, 1, 0, 4, 4),
    (128, 1, 4, 2, 1, 0, 16, 16), (256, 128, 4, 2, 1, 0, 8, 8),
]


@pytest.mark.parametrize("case", CONVT_CASES)
def test_conv_transpose2d_fwd_bwd(ops, case, engine):
    Ci, Co, k, s, p, op, H, W = case
    x = torch.randn(2, Ci, H, W, generator=G(44)).requires_grad_(True)
    w = (torch.randn(Ci, Co, k, k, generator=G(45)) * 0.05).requires_grad_(True)
    b = torch.randn(Co, generator=G(46)).requires_grad_(True)
    ref = torch.tanh(F.conv_transpose2d(x, w, b, stride=s, padding=p, output_padding=op))
    go = torch.randn(ref.shape, generator=G(47))
    (ref * go).sum().backward()
    xc, wc, bc = cl(x.detach()).requires_grad_(True), w.detach().cuda().requires_grad_(True), b.detach().cuda().requires_grad_(True)
    out = ops.conv_transpose2d(xc, wc, bc, s, p, op, act_out=ops.ACT_TANH)
    (out * go.cuda()).sum().backward()
    assert rel_l2(out.cpu(), ref.detach()) <= max(engine[2], 3e-5 if Ci >= 1024 else 0.0)
    btol = max(3e-5, 2 * engine[2])        # the tanh backward reads the forward output
    assert rel_l2(xc.grad.cpu(), x.grad) <= max(btol, engine[3]) and rel_l2(wc.grad.cpu(), w.grad) <= max(btol, engine[3])
    assert rel_l2(bc.grad.cpu(), b.grad) <= btol


def test_tc_large_tiles_and_split_k(ops):
    """full 16x8 tiles over several images, ragged edges, and the split-K path of the tiny-M layers."""
    old = dict(ops.CONFIG)
    try:
        ops.CONFIG.update(engine="tc", passes=3, dtype="f16", wgrad_passes=3, big_hw=0)
        for (Ci, Co, k, s, p, N, H, W, split) in [(128, 128, 3, 1, 1, 5, 64, 64, 1), (64, 128, 3, 1, 1, 2, 40, 24, 1),
                                                  (512, 512, 4, 2, 1, 3, 8, 8, -1), (512, 512, 4, 2, 1, 12, 4, 4, 8),
                                                  (256, 512, 4, 2, 1, 2, 20, 12, -1)]:
            ops.CONFIG.update(split_k=split)
            x = torch.randn(N, Ci, H, W, generator=G(70))
            w = torch.randn(Co, Ci, k, k, generator=G(71)) * 0.03
            b = torch.randn(Co, generator=G(72))
            ref = F.conv2d(x, w, b, stride=s, padding=p)
            with torch.no_grad():
                out = ops.conv2d(cl(x), w.cuda(), b.cuda(), s, p)
            assert rel_l2(out.cpu(), ref) <= 1e-5, (Ci, Co, k, s, N, H, W, split)
    finally:
        ops.CONFIG.update(old)


@pytest.mark.parametrize("wpass,tol", [(3, 5e-5), (2, 4e-3), (1, 8e-3)])
def test_tc_wgrad_large(ops, wpass, tol):
    """weight gradient on the tcgen05 path: several 64-pixel K tiles, split-K, ragged edges, every layout
    (NORMAL / PAIR / S2D conv, S2D transposed conv incl. the 1-channel head)."""
    old = dict(ops.CONFIG)
    try:
        ops.CONFIG.update(engine="tc", passes=3, dtype="f16", wgrad_passes=wpass, big_hw=0)
        for (kind, Ci, Co, k, s, p, op, N, H, W) in [("conv", 128, 128, 3, 1, 1, 0, 3, 40, 24), ("conv", 32, 128, 7, 1, 3, 0, 2, 36, 20),
                                                     ("conv", 261, 64, 4, 2, 1, 0, 2, 32, 48), ("conv", 64, 128, 3, 2, 1, 0, 2, 24, 24),
                                                     ("conv", 512, 512, 4, 2, 1, 0, 3, 4, 4), ("convT", 1024, 256, 4, 2, 1, 0, 2, 8, 8),
                                                     ("convT", 128, 1, 4, 2, 1, 0, 2, 32, 32), ("convT", 128, 64, 3, 2, 1, 1, 2, 20, 12)]:
            x = torch.randn(N, Ci, H, W, generator=G(80)).requires_grad_(True)
            if kind == "conv":
                w = (torch.randn(Co, Ci, k, k, generator=G(81)) * 0.05).requires_grad_(True)
                ref = F.conv2d(x, w, None, stride=s, padding=p)
            else:
                w = (torch.randn(Ci, Co, k, k, generator=G(81)) * 0.05).requires_grad_(True)
                ref = F.conv_transpose2d(x, w, None, stride=s, padding=p, output_padding=op)
            go = torch.randn(ref.shape, generator=G(82))
            (ref * go).sum().backward()
            xc, wc = cl(x.detach()).requires_grad_(True), w.detach().cuda().requires_grad_(True)
            out = ops.conv2d(xc, wc, None, s, p) if kind == "conv" else ops.conv_transpose2d(xc, wc, None, s, p, op)
            (out * go.cuda()).sum().backward()
            assert rel_l2(wc.grad.cpu(), w.grad) <= tol, (kind, Ci, Co, k, s, N, H, W, rel_l2(wc.grad.cpu(), w.grad))
            assert rel_l2(xc.grad.cpu(), x.grad) <= 5e-5
    finally:
        ops.CONFIG.update(old)


def test_default_dgrad_policy_single_pass_on_big_layers(ops):
    """ops.CONFIG defaults: data-gradient GEMMs of layers with >= 32^2 output pixels run ONE bf16 pass (8e-3 class),
    smaller layers keep the three-pass 5e-5 class; forward results are unaffected."""
    assert ops.CONFIG["big_hw"] == 1024 and ops.CONFIG["big_bwd_passes"] == 1 and ops.CONFIG["big_fwd_passes"] == 0
    for (kind, Ci, Co, k, s, p, op, H, W, lo, hi) in [("conv", 128, 128, 3, 1, 1, 0, 40, 48, 1e-4, 8e-3), ("conv", 128, 128, 3, 1, 1, 0, 16, 16, 0, 5e-5),
                                                      ("conv", 32, 128, 7, 1, 3, 0, 64, 64, 1e-4, 8e-3), ("conv", 64, 128, 4, 2, 1, 0, 64, 64, 1e-4, 8e-3),
                                                      ("convT", 256, 128, 4, 2, 1, 0, 32, 32, 1e-4, 8e-3), ("convT", 128, 64, 3, 2, 1, 1, 8, 8, 0, 5e-5)]:
        x = torch.randn(2, Ci, H, W, generator=G(180)).requires_grad_(True)
        if kind == "conv":
            w = (torch.randn(Co, Ci, k, k, generator=G(181)) * 0.05).requires_grad_(True)
            ref = F.conv2d(x, w, None, stride=s, padding=p)
        else:
            w = (torch.randn(Ci, Co, k, k, generator=G(181)) * 0.05).requires_grad_(True)
            ref = F.conv_transpose2d(x, w, None, stride=s, padding=p, output_padding=op)
        go = torch.randn(ref.shape, generator=G(182))
        (ref * go).sum().backward()
        xc, wc = cl(x.detach()).requires_grad_(True), w.detach().cuda().requires_grad_(True)
        out = ops.conv2d(xc, wc, None, s, p) if kind == "conv" else ops.conv_transpose2d(xc, wc, None, s, p, op)
        (out * go.cuda()).sum().backward()
        e = rel_l2(xc.grad.cpu(), x.grad)
        assert rel_l2(out.detach().cpu(), ref.detach()) <= 1e-5 and lo <= e <= hi, (kind, Ci, Co, k, H, W, e)


def test_cat_conv2d_restricted_dgrad(ops):
    """conv over a lazily concatenated input: forward == conv(cat), gradients only for the parts that need one"""
    old = dict(ops.CONFIG)
    try:
        ops.CONFIG.update(engine="tc", passes=3, dtype="f16", wgrad_passes=3, big_hw=0)
        Cs, H, W = (128, 128, 2, 3), 32, 48
        xs = [torch.randn(2, c, H, W, generator=G(90 + i)) for i, c in enumerate(Cs)]
        xs[1].requires_grad_(True)
        w = (torch.randn(64, sum(Cs), 4, 4, generator=G(95)) * 0.05).requires_grad_(True)
        b = torch.randn(64, generator=G(96)).requires_grad_(True)
        ref = F.conv2d(torch.cat(xs, 1), w, b, stride=2, padding=1)
        go = torch.randn(ref.shape, generator=G(97))
        (ref * go).sum().backward()
        xc = [cl(x.detach()) for x in xs]
        xc[1].requires_grad_(True)
        wc, bc = w.detach().cuda().requires_grad_(True), b.detach().cuda().requires_grad_(True)
        out = ops.cat_conv2d(xc, wc, bc, 2, 1)
        (out * go.cuda()).sum().backward()
        assert rel_l2(out.cpu(), ref.detach()) <= 1e-5
        assert rel_l2(xc[1].grad.cpu(), xs[1].grad) <= 5e-5 and xc[0].grad is None
        assert rel_l2(wc.grad.cpu(), w.grad) <= 5e-5 and rel_l2(bc.grad.cpu(), b.grad) <= 1e-5
    finally:
        ops.CONFIG.update(old)


def test_fused_prologue_matches_unfused(ops):
    """[InstanceNorm, ReLU, ReflectionPad, Conv] as ONE fused unit == the same modules run one by one (fwd + bwd)"""
    from dsr_b200 import networks as nw
    old = dict(ops.CONFIG)
    try:
        ops.CONFIG.update(engine="tc", passes=3, dtype="f16", wgrad_passes=3, big_hw=0)
        torch.manual_seed(5)
        mods = [nw.Conv2d(64, 128, 3, stride=2, padding=1), nw.InstanceNorm2d(128), nw.ReLU(True), nw.ReflectionPad2d(1),
                nw.Conv2d(128, 128, 3, padding=0), nw.InstanceNorm2d(128), nw.ReLU(True),
                nw.ConvTranspose2d(128, 64, 3, stride=2, padding=1, output_padding=1), nw.Tanh()]
        for m in mods:
            m.cuda()
        x = torch.randn(2, 64, 32, 32, generator=G(98))
        go = torch.randn(2, 64, 32, 32, generator=G(99)).cuda()
        res = []
        for fused in (True, False):
            xc = cl(x).requires_grad_(True)
            for m in mods:
                for p_ in m.parameters():
                    p_.grad = None
            if fused:
                y = nw.run_fused(mods, xc)
            else:
                y = xc
                for m in mods:
                    y = m(y)
            (y * go).sum().backward()
            res.append((y.detach().cpu(), xc.grad.cpu(), [p_.grad.cpu().clone() for m in mods for p_ in m.parameters()]))
        assert rel_l2(res[0][0], res[1][0]) <= 2e-5 and rel_l2(res[0][1], res[1][1]) <= 2e-4
        for ga, gb in zip(res[0][2], res[1][2]):
            if gb.dim() == 4:                                   # weights (biases in front of an IN are ~0 noise)
                assert rel_l2(ga, gb) <= 2e-4
    finally:
        ops.CONFIG.update(old)


@pytest.mark.parametrize("groups,C,HW,layout_pad", [(0, 64, (24, 40), 1), (0, 128, (9, 7), 0), (8, 64, (12, 10), 1), (0, 20, (6, 5), 2)])
def test_folded_norm_finalize_kernels_bit_identical(ops, groups, C, HW, layout_pad):
    """dsr_tc_prep_fin / dsr_norm_apply_fwd_fin (statistics finalised inside the consumer's launch) against
    dsr_norm_finalize + dsr_tc_prep / dsr_norm_apply_fwd on the SAME sums: same arithmetic, so operands, outputs and the
    constants written for the backward pass must be EQUAL, not close.  InstanceNorm: networks.py:30, :380-381, :478-480;
    GroupNorm(8, C): translation_network.py:46."""
    H, W = HW
    N = 3
    xh = (torch.randn(N, H, W, C, generator=G(61)) * 2 + 0.7).cuda()
    gamma = torch.randn(C, generator=G(62)).cuda() if groups else None
    beta = torch.randn(C, generator=G(63)).cuda() if groups else None
    plan = dict(layout=0, Cp=(C + 7) // 8 * 8, Ca=(C + 63) // 64 * 64)
    sums = torch.zeros(N * C * 2, dtype=torch.float64, device="cuda")       # ONE set of statistics for both routes
    ops._call("dsr_channel_sums", ops._p(xh), N, H * W, C, ops._p(sums, torch.float64))
    out = []
    for fold in (True, False):
        old = dict(ops.CONFIG)
        try:
            ops.CONFIG.update(fold_finalize=fold, passes=3, dtype="f16")
            n0 = _launch_count()
            prm = ops._norm_params(xh, groups, gamma, beta, 1e-5, sums, lazy=True)
            assert (getattr(prm, "_fin", None) is not None) == fold
            ahi, alo, Ha, Wa = ops._tc_prep(xh, plan, layout_pad, ops.PAD_REFLECT, prm, ops.ACT_RELU)[:4]
            prm2 = ops._norm_params(xh, groups, gamma, beta, 1e-5, sums, lazy=True)
            y = torch.empty_like(xh)
            ops._norm_apply_fwd(xh, prm2, xh, y, ops.ACT_RELU)
            torch.cuda.synchronize()
            out.append((ahi.view(torch.int16).cpu(), alo.view(torch.int16).cpu(), prm.cpu(), prm2.cpu(), y.cpu(), _launch_count() - n0))
        finally:
            ops.CONFIG.update(old)
    for a, b in zip(out[0][:5], out[1][:5]):
        assert torch.equal(a, b)
    assert out[0][5] == out[1][5] - 2                               # two finalize launches fewer
    ref = F.relu(F.instance_norm(xh.permute(0, 3, 1, 2).cpu(), eps=1e-5) if groups == 0 else
                 F.group_norm(xh.permute(0, 3, 1, 2).cpu(), groups, gamma.cpu(), beta.cpu(), 1e-5)) + xh.permute(0, 3, 1, 2).cpu()
    assert torch.allclose(out[0][4].permute(0, 3, 1, 2), ref, atol=5e-5)


@pytest.mark.parametrize("layout,C,Ca,HW,pad,mode", [(0, 128, 128, (20, 33), 1, 1), (0, 32, 64, (17, 40), 3, 1), (0, 64, 64, (9, 7), 2, 0),
                                                     (2, 64, 256, (24, 18), 1, 0), (2, 32, 128, (13, 11), 1, 2), (1, 32, 64, (12, 30), 3, 1),
                                                     (0, 256, 256, (6, 5), 0, 0), (0, 512, 512, (4, 4), 1, 2), (0, 16, 64, (8, 300), 1, 0)])
@pytest.mark.parametrize("variant", ["plain", "prm_relu", "fin_lrelu_bf", "csum_nolo", "bf16"])
def test_prep_fast_kernel_is_bit_identical(ops, layout, C, Ca, HW, pad, mode, variant):
    """tc_prep_fast_kernel (one channel group per thread, constants in registers, two items in flight) against the general
    tc_prep_kernel (DSR_PREP_FAST=0) over the layouts / paddings / fused prologues the step uses: every output plane, the
    constants written for the backward pass and the bias-gradient sums must be EQUAL.  The operand of every nn.Conv2d /
    nn.ConvTranspose2d, networks.py:379-415, :544-616."""
    import os
    H, W = HW
    N = 3
    csum_wanted = variant == "csum_nolo"
    if csum_wanted and (layout == 1 or (pad and mode != 0)):
        pytest.skip("bias-gradient sums need every source element written exactly once")
    xh = (torch.randn(N, H, W, C, generator=G(71)) * 1.7 + 0.2).cuda()
    plan = dict(layout=layout, Cp=Ca // 4 if layout == 2 else (Ca // 2 if layout == 1 else C), Ca=Ca)
    sums = torch.zeros(N * C * 2, dtype=torch.float64, device="cuda")
    ops._call("dsr_channel_sums", ops._p(xh), N, H * W, C, ops._p(sums, torch.float64))
    out = []
    old = dict(ops.CONFIG)
    try:
        ops.CONFIG.update(passes=3, dtype="f16", fold_finalize=True, csum_reps=1)
        for fast in ("1", "0"):
            os.environ["DSR_PREP_FAST"] = fast
            prm, act, slope, kw = None, ops.ACT_NONE, 0.0, {}
            if variant == "prm_relu":
                prm, act = ops._norm_params(xh, 0, None, None, 1e-5, sums, lazy=False), ops.ACT_RELU
            elif variant == "fin_lrelu_bf":
                prm, act, slope, kw = ops._norm_params(xh, 0, None, None, 1e-5, sums, lazy=True), ops.ACT_LRELU, 0.2, dict(also_bf16=True)
            elif variant == "csum_nolo":
                kw = dict(csum=torch.zeros(C, dtype=torch.float64, device="cuda"), need_lo=False, dtype="bf16")
            elif variant == "bf16":
                kw = dict(dtype="bf16")
            r = ops._tc_prep(xh, plan, pad, mode, prm, act, slope, **kw)
            torch.cuda.synchronize()
            planes = [t.view(torch.int16).cpu() for t in r if torch.is_tensor(t)]
            extra = [prm.cpu()] if prm is not None else []
            out.append((planes, extra, kw.get("csum")))
        assert len(out[0][0]) == len(out[1][0]) >= 1
        for a, b in zip(out[0][0] + out[0][1], out[1][0] + out[1][1]):
            assert torch.equal(a, b)
        if csum_wanted:
            assert torch.allclose(out[0][2].cpu(), out[1][2].cpu(), rtol=1e-5, atol=1e-4)     # block partials: fp32 shared-memory atomics
            assert torch.allclose(out[0][2].cpu(), xh.double().sum(dim=(0, 1, 2)).cpu(), rtol=1e-5, atol=1e-3)
    finally:
        os.environ.pop("DSR_PREP_FAST", None)
        ops.CONFIG.update(old)


@pytest.mark.parametrize("train", [True, False])
def test_residual_block_tail_fused_with_next_operand(ops, train):
    """dsr_tc_prep_norm_res: the closing InstanceNorm + skip add of a residual block (networks.py:478-480) and the stand-alone
    norm + ReLU in front of the block stack also write the NEXT convolution's arranged operand.  Same arithmetic as the two
    separate passes: outputs and gradients must agree to rounding-order noise, with one launch less per block."""
    from dsr_b200 import networks as nw
    old = dict(ops.CONFIG)
    try:
        ops.CONFIG.update(engine="tc", passes=3, dtype="f16", wgrad_passes=3, big_hw=0, fuse_bwd_prep=False)   # (one fusion at a time)
        torch.manual_seed(21)
        norm = nw.get_norm_layer("instance")
        mods = [nw.Conv2d(16, 64, 3, stride=2, padding=1), norm(64), nw.ReLU(True)] + \
               [nw.ResnetBlock(64, "reflect", norm, False, True) for _ in range(3)] + \
               [nw.ConvTranspose2d(64, 32, 3, stride=2, padding=1, output_padding=1), norm(32), nw.ReLU(True),
                nw.ReflectionPad2d(3), nw.Conv2d(32, 8, 7, padding=0)]
        net = nw.FusedSequential(*mods).cuda()
        x = torch.randn(2, 16, 48, 40, generator=G(298))
        res = []
        for fuse in (True, False):
            ops.CONFIG.update(fuse_norm_prep=fuse)
            ops.zero_pool_reset("cuda")
            for p_ in net.parameters():
                p_.grad = None
            xc = cl(x).requires_grad_(train)
            with _CallLog() as names:
                if train:
                    y = net(xc)
                    (y * y).sum().backward()
                else:
                    with torch.no_grad():
                        y = net(xc)
            res.append((y.detach().cpu(), xc.grad.cpu() if train else None, [p_.grad.cpu().clone() for p_ in net.parameters()] if train else [],
                        names.count("dsr_tc_prep_norm_res"), sum(1 for nm in names if "pack" not in nm)))
        assert res[0][3] == 4 and res[1][3] == 0                    # stand-alone norm + three block tails
        assert res[0][4] == res[1][4] - 4
        assert rel_l2(res[0][0], res[1][0]) <= 1e-6
        if train:
            assert rel_l2(res[0][1], res[1][1]) <= 2e-4
            for ga, gb in zip(res[0][2], res[1][2]):
                if gb.dim() == 4:
                    assert rel_l2(ga, gb) <= 2e-4
    finally:
        ops.CONFIG.update(old)


@pytest.mark.parametrize("layout,C,Cp,Ca,HW,pad,wa_min", [(0, 128, 128, 128, (20, 24), 2, 0), (0, 64, 64, 64, (16, 16), 1, 24),
                                                          (2, 64, 64, 256, (18, 22), 1, 0), (0, 72, 72, 128, (9, 7), 0, 0),
                                                          (2, 32, 32, 128, (13, 11), 1, 0), (0, 512, 512, 512, (4, 4), 1, 0)])
@pytest.mark.parametrize("act,csum_wanted,need_lo", [(0, False, True), (1, True, False), (1, False, True), (0, True, False)])
def test_in_bwd_apply_fused_with_the_dy_operand_is_bit_identical(ops, layout, C, Cp, Ca, HW, pad, wa_min, act, csum_wanted, need_lo):
    """dsr_tc_prep_in_bwd: InstanceNorm2d(affine=False) [+ReLU] backward (networks.py:30, :380-381) that also writes the
    zero-framed 16-bit dY operand of the convolution in front of the norm layer.  Against the two passes it replaces
    (dsr_in_bwd_apply, then dsr_tc_prep of its output): dx and every operand plane must be EQUAL, the bias-gradient sums equal
    to summation-order noise."""
    H, W = HW
    N = 3
    ACT = ops.ACT_RELU if act else ops.ACT_NONE
    xh = (torch.randn(N, H, W, C, generator=G(171)) * 1.3 + 0.4).cuda()
    g = torch.randn(N, H, W, C, generator=G(172)).cuda()
    xn = xh.permute(0, 3, 1, 2)
    mean = xn.mean(dim=(2, 3))
    rstd = (xn.var(dim=(2, 3), unbiased=False) + 1e-5).rsqrt()
    prm = torch.stack([mean, rstd, torch.zeros_like(mean)]).contiguous().view(-1)            # [3][N][C]
    old = dict(ops.CONFIG)
    try:
        ops.CONFIG.update(passes=3, csum_reps=4)
        plan = dict(layout=layout, Cp=Cp, Ca=Ca)
        if wa_min:
            plan["Wa"] = wa_min
        sums2 = torch.zeros(N * C * 2, dtype=torch.float64, device="cuda")
        ops._call("dsr_in_bwd_sums", ops._p(xh), ops._p(g), ops._p(prm), N, H * W, C, ACT, ops._p(sums2, torch.float64))
        # the two passes
        gx_ref = torch.empty_like(xh)
        ops._call("dsr_in_bwd_apply", ops._p(xh), ops._p(g), ops._p(prm), ops._p(sums2, torch.float64), ops._p(gx_ref), N, H * W, C, ACT)
        cs_ref = torch.zeros(C * 4, dtype=torch.float64, device="cuda") if csum_wanted else None
        ahi_r, alo_r, Ha, Wa = ops._tc_prep(gx_ref, plan, pad, ops.PAD_ZERO, dtype="bf16", csum=cs_ref, need_lo=need_lo)
        # the fused pass
        gx = torch.full_like(xh, float("nan"))
        ahi = torch.full((N, Ha, Wa, Ca), -1.0, device="cuda", dtype=torch.bfloat16)
        alo = torch.full((N, Ha, Wa, Ca), -1.0, device="cuda", dtype=torch.bfloat16) if need_lo else None
        cs = torch.zeros(C * 4, dtype=torch.float64, device="cuda") if csum_wanted else None
        ops._call("dsr_tc_prep_in_bwd", ops._p(xh), ops._p(g), ops._p(prm), ops._p(sums2, torch.float64), ops._p(gx), N, H, W, C, ACT,
                  pad, layout, Cp, ops._p(ahi, torch.bfloat16), ops._p(alo, torch.bfloat16), Ha, Wa, Ca, 0,
                  ops._p(cs, torch.float64), 4 if csum_wanted else 1)
        torch.cuda.synchronize()
        assert torch.equal(gx, gx_ref)
        assert torch.equal(ahi.view(torch.int16), ahi_r.view(torch.int16))
        assert (alo_r is None) == (alo is None)
        if need_lo:
            assert torch.equal(alo.view(torch.int16), alo_r.view(torch.int16))
        if csum_wanted:
            tot, tot_r = cs.view(4, C).sum(0).cpu(), cs_ref.view(4, C).sum(0).cpu()
            assert torch.allclose(tot, tot_r, rtol=1e-5, atol=1e-4)
            assert torch.allclose(tot, gx_ref.double().sum(dim=(0, 1, 2)).cpu(), rtol=1e-5, atol=1e-3)
        # against the closed form (fp64 on the CPU): dx = rstd * (g' - mean(g') - xhat * mean(g' * xhat))
        xd, gd = xh.double().cpu().permute(0, 3, 1, 2), g.double().cpu().permute(0, 3, 1, 2)
        xhat = (xd - xd.mean(dim=(2, 3), keepdim=True)) * (xd.var(dim=(2, 3), unbiased=False, keepdim=True) + 1e-5).rsqrt()
        gp = gd * (xhat > 0) if act else gd
        ref = (xd.var(dim=(2, 3), unbiased=False, keepdim=True) + 1e-5).rsqrt() * (
            gp - gp.mean(dim=(2, 3), keepdim=True) - xhat * (gp * xhat).mean(dim=(2, 3), keepdim=True))
        assert rel_l2(gx.cpu().permute(0, 3, 1, 2).double(), ref) <= 1e-5
    finally:
        ops.CONFIG.update(old)


def test_norm_backward_writes_the_next_dy_operand_from_the_second_step(ops):
    """ops._in_bwd: from the second backward pass on (the first one records what each convolution asks of its dY) the
    InstanceNorm backward in front of a convolution writes that convolution's dY operand itself.  Same arithmetic as the
    separate passes: gradients agree to summation-order noise, with one launch less per normalised convolution."""
    from dsr_b200 import networks as nw
    old = dict(ops.CONFIG)
    try:
        ops.CONFIG.update(engine="tc", passes=3, dtype="f16", wgrad_passes=3, big_hw=0)
        torch.manual_seed(23)
        norm = nw.get_norm_layer("instance")
        mods = [nw.ReflectionPad2d(3), nw.Conv2d(8, 32, 7, padding=0), norm(32), nw.ReLU(True),
                nw.Conv2d(32, 64, 3, stride=2, padding=1), norm(64), nw.ReLU(True)] + \
               [nw.ResnetBlock(64, "reflect", norm, False, True) for _ in range(2)] + \
               [nw.ConvTranspose2d(64, 32, 3, stride=2, padding=1, output_padding=1), norm(32), nw.ReLU(True),
                nw.ReflectionPad2d(3), nw.Conv2d(32, 8, 7, padding=0)]
        net = nw.FusedSequential(*mods).cuda()
        x = torch.randn(2, 8, 48, 40, generator=G(299))
        res = []
        for fuse in (True, False):
            ops.CONFIG.update(fuse_bwd_prep=fuse)
            ops._DY_SPEC.clear()
            per_step = []
            for step in range(2):
                ops.zero_pool_reset("cuda")
                for p_ in net.parameters():
                    p_.grad = None
                xc = cl(x).requires_grad_(True)
                with _CallLog() as names:
                    y = net(xc)
                    (y * y).sum().backward()
                per_step.append((xc.grad.cpu(), [p_.grad.cpu().clone() for p_ in net.parameters()],
                                 names.count("dsr_tc_prep_in_bwd"), names.count("dsr_in_bwd_apply"),
                                 sum(1 for nm in names if "pack" not in nm)))
            res.append(per_step)
        (f1, f2), (u1, u2) = res
        assert f1[2] == 0 and u1[2] == 0 and u2[2] == 0            # first step: nothing recorded yet; switch off: never
        assert f2[2] >= 5 and f2[2] + f2[3] == u2[3]               # fused applies replace plain ones one for one ...
        assert f2[4] == u2[4] - f2[2]                              # ... and each saves the consumer's dsr_tc_prep launch
        for a, b in ((f2, u2), (f2, f1)):
            # (closeness, not equality: the norm statistics come from fp64 atomics in the GEMM epilogues, whose run-to-run
            # differences in the last bits the backward operands amplify - measured 2.0e-4 between two identical runs)
            assert rel_l2(a[0], b[0]) <= 1e-3
            for ga, gb in zip(a[1], b[1]):
                if gb.dim() == 4:          # (the bias gradients in front of a norm layer are sums that cancel: pure rounding noise)
                    assert rel_l2(ga, gb) <= 1e-3
        assert rel_l2(f2[1][-1], u2[1][-1]) <= 1e-3                # the last bias (no norm behind it) is a real gradient
    finally:
        ops.CONFIG.update(old)
        ops._DY_SPEC.clear()


@pytest.mark.parametrize("padding", ["reflect", "zero"])
def test_residual_skip_gradient_joins_the_first_conv_backward(ops, padding):
    """out = x + conv_block(x) (networks.py:478-480): with fuse_skip_grad the skip connection leaves through the block's first
    convolution node, so both gradients of x meet in its backward and are added by its padding adjoint
    (dsr_pad2d_bwd_pitch_add) - or by one add inside the node where that route does not apply - instead of autograd's own
    accumulation.  Same sums, same order of the two terms: gradients agree to the noise of the statistics' atomics."""
    from dsr_b200 import networks as nw
    old = dict(ops.CONFIG)
    try:
        ops.CONFIG.update(engine="tc", passes=3, dtype="f16", wgrad_passes=3, big_hw=0, fuse_bwd_prep=False)
        torch.manual_seed(29)
        norm = nw.get_norm_layer("instance")
        mods = [nw.Conv2d(16, 64, 3, stride=2, padding=1), norm(64), nw.ReLU(True)] + \
               [nw.ResnetBlock(64, padding, norm, False, True) for _ in range(3)] + [nw.Conv2d(64, 16, 3, padding=1)]
        net = nw.FusedSequential(*mods).cuda()
        x = torch.randn(2, 16, 40, 48, generator=G(301))
        res = []
        for fuse in (True, False):
            ops.CONFIG.update(fuse_skip_grad=fuse)
            ops.zero_pool_reset("cuda")
            for p_ in net.parameters():
                p_.grad = None
            xc = cl(x).requires_grad_(True)
            with _CallLog() as names:
                y = net(xc)
                (y * y).sum().backward()
            res.append((y.detach().cpu(), xc.grad.cpu(), [p_.grad.cpu().clone() for p_ in net.parameters()],
                        names.count("dsr_pad2d_bwd_pitch_add")))
        assert res[0][3] == (3 if padding == "reflect" else 0) and res[1][3] == 0
        assert rel_l2(res[0][0], res[1][0]) <= 1e-6
        assert rel_l2(res[0][1], res[1][1]) <= 1e-3
        for ga, gb in zip(res[0][2], res[1][2]):
            if gb.dim() == 4:
                assert rel_l2(ga, gb) <= 1e-3
        # forward under no_grad takes the plain route (no alias output)
        with torch.no_grad():
            assert rel_l2(net(cl(x)).cpu(), res[0][0]) <= 1e-6
    finally:
        ops.CONFIG.update(old)


@pytest.mark.parametrize("C,H,W,pad,mode,act_in,tanh", [(64, 70, 37, 3, "replicate", 0, True), (64, 128, 128, 3, "reflect", 1, False),
                                                        (20, 64, 16, 3, "zeros", 0, False), (6, 97, 50, 0, "zeros", 2, True),
                                                        (32, 256, 40, 3, "zeros", 0, False)])
def test_conv_out1_register_blocked_kernel(ops, C, H, W, pad, mode, act_in, tanh):
    """conv_out1_s1_rb_kernel (four output rows per thread, activations of a kernel column held in registers): the 7x7 -> 1
    channel heads (translation_network.py:495, G_A_d's Conv2d(64, 1, 7) + Tanh behind a replicate pad) at sizes that leave
    partial tiles in both directions, every padding mode, with the fused norm-apply / activation prologue, against a CPU fp64
    convolution and against the one-pixel-per-thread kernel (DSR_OUT1_RB=0)."""
    import os
    N = 2
    x = torch.randn(N, C, H, W, generator=G(311)) * 1.5 + 0.2
    w = torch.randn(1, C, 7, 7, generator=G(312)) * 0.05
    b = torch.randn(1, generator=G(313))
    old = dict(ops.CONFIG)
    try:
        ops.CONFIG.update(engine="tc", out1=True)
        xin = x.double()
        prm, act, slope = None, ops.ACT_NONE, 0.0
        xh = x.permute(0, 2, 3, 1).contiguous().cuda()
        if act_in:
            mean, var = xin.mean(dim=(2, 3), keepdim=True), xin.var(dim=(2, 3), unbiased=False, keepdim=True)
            rstd = (var + 1e-5).rsqrt()
            xin = (xin - mean) * rstd
            xin = F.relu(xin) if act_in == 1 else F.leaky_relu(xin, 0.2)
            act, slope = (ops.ACT_RELU, 0.0) if act_in == 1 else (ops.ACT_LRELU, 0.2)
            prm = torch.stack([mean.view(N, C), rstd.view(N, C), torch.zeros(N, C, dtype=torch.float64)]).float().contiguous().view(-1).cuda()
        pm = {"zeros": "constant"}.get(mode, mode)
        ref = F.conv2d(F.pad(xin, (pad,) * 4, mode=pm), w.double(), b.double())
        if tanh:
            ref = torch.tanh(ref)
        Ho, Wo = H + 2 * pad - 6, W + 2 * pad - 6
        outs = []
        wd, bd = w[0].contiguous().cuda(), b.cuda()                # (kept alive: _p hands out raw pointers)
        for rb in ("1", "0"):
            os.environ["DSR_OUT1_RB"] = rb
            y = torch.full((N, Ho, Wo, 1), float("nan"), device="cuda")
            ops._call("dsr_conv_out1", ops._p(xh), N, H, W, C, ops._p(prm), act, slope, ops._p(wd), ops._p(bd),
                      7, 7, pad, ops.PAD_MODES[mode], 0, ops.ACT_TANH if tanh else ops.ACT_NONE, ops._p(y))
            torch.cuda.synchronize()
            outs.append(y.cpu().view(N, 1, Ho, Wo))
        assert rel_l2(outs[0].double(), ref) <= 2e-6
        assert rel_l2(outs[0], outs[1]) <= 2e-6
    finally:
        os.environ.pop("DSR_OUT1_RB", None)
        ops.CONFIG.update(old)


@pytest.mark.parametrize("C,H,W,act_in,tanh", [(128, 40, 23, 1, True), (20, 33, 16, 0, False), (6, 64, 50, 2, True), (128, 128, 128, 1, True)])
def test_conv_transpose_out1_register_blocked_kernel(ops, C, H, W, act_in, tanh):
    """convT4_out1_rb_kernel (four input rows = 16 outputs per thread): the U-Net heads ConvTranspose2d(128, 1, 4, 2, 1) + Tanh
    behind ReLU (networks.py:553-555) at sizes with partial tiles, against a CPU fp64 transposed convolution and against the
    one-position-per-thread kernel (DSR_OUT1_RB=0)."""
    import os
    N = 2
    x = torch.randn(N, C, H, W, generator=G(321)) * 1.5 + 0.2
    w = torch.randn(C, 1, 4, 4, generator=G(322)) * 0.05
    b = torch.randn(1, generator=G(323))
    xin = x.double()
    prm, act, slope = None, ops.ACT_NONE, 0.0
    xh = x.permute(0, 2, 3, 1).contiguous().cuda()
    if act_in:
        mean, var = xin.mean(dim=(2, 3), keepdim=True), xin.var(dim=(2, 3), unbiased=False, keepdim=True)
        rstd = (var + 1e-5).rsqrt()
        xin = (xin - mean) * rstd
        xin = F.relu(xin) if act_in == 1 else F.leaky_relu(xin, 0.2)
        act, slope = (ops.ACT_RELU, 0.0) if act_in == 1 else (ops.ACT_LRELU, 0.2)
        prm = torch.stack([mean.view(N, C), rstd.view(N, C), torch.zeros(N, C, dtype=torch.float64)]).float().contiguous().view(-1).cuda()
    ref = F.conv_transpose2d(xin, w.double(), b.double(), stride=2, padding=1)
    if tanh:
        ref = torch.tanh(ref)
    outs = []
    wd, bd = w[:, 0].contiguous().cuda(), b.cuda()                 # (kept alive: _p hands out raw pointers)
    try:
        for rb in ("1", "0"):
            os.environ["DSR_OUT1_RB"] = rb
            y = torch.full((N, 2 * H, 2 * W, 1), float("nan"), device="cuda")
            ops._call("dsr_conv_out1", ops._p(xh), N, H, W, C, ops._p(prm), act, slope, ops._p(wd), ops._p(bd),
                      4, 4, 1, ops.PAD_ZERO, 1, ops.ACT_TANH if tanh else ops.ACT_NONE, ops._p(y))
            torch.cuda.synchronize()
            outs.append(y.cpu().view(N, 1, 2 * H, 2 * W))
        assert rel_l2(outs[0].double(), ref) <= 2e-6
        assert rel_l2(outs[0], outs[1]) <= 2e-6
    finally:
        os.environ.pop("DSR_OUT1_RB", None)


def test_folded_norm_finalize_through_the_layers(ops):
    """the same comparison through the layer stack (prologue route and stand-alone InstanceNorm), forward and backward; the
    statistics come from fp64 atomics in the GEMM epilogues here (run-to-run differences in the last bits, amplified by the
    single-pass bf16 backward operands), so this is a closeness check - the bitwise one is the kernel-level test above"""
    from dsr_b200 import networks as nw
    old = dict(ops.CONFIG)
    try:
        ops.CONFIG.update(engine="tc", passes=3, dtype="f16", wgrad_passes=3, big_hw=0)
        torch.manual_seed(11)
        mods = [nw.Conv2d(32, 64, 3, stride=2, padding=1), nw.InstanceNorm2d(64), nw.ReLU(True), nw.ReflectionPad2d(1),
                nw.Conv2d(64, 64, 3, padding=0), nw.InstanceNorm2d(64), nw.ReLU(True),
                nw.ConvTranspose2d(64, 32, 3, stride=2, padding=1, output_padding=1), nw.InstanceNorm2d(32), nw.ReLU(True),
                nw.ReflectionPad2d(3), nw.Conv2d(32, 128, 7, padding=0)]
        for m in mods:
            m.cuda()
        x = torch.randn(3, 32, 24, 40, generator=G(198)) * 2 + 0.3
        res = []
        for fold in (True, False):
            ops.CONFIG.update(fold_finalize=fold)
            n0 = _launch_count()
            xc = cl(x).requires_grad_(True)
            for m in mods:
                for p_ in m.parameters():
                    p_.grad = None
            y = nw.run_fused(mods, xc)                              # prologue route: dsr_tc_prep_fin
            z = ops.instance_norm(y, 1e-5, 1, None)                 # stand-alone route: dsr_norm_apply_fwd_fin
            (z * z).sum().backward()
            res.append((z.detach().cpu(), xc.grad.cpu(), [p_.grad.cpu().clone() for m in mods for p_ in m.parameters()],
                        _launch_count() - n0))
        assert rel_l2(res[0][0], res[1][0]) <= 1e-6 and rel_l2(res[0][1], res[1][1]) <= 2e-4
        for ga, gb in zip(res[0][2], res[1][2]):
            if gb.dim() == 4:
                assert rel_l2(ga, gb) <= 2e-4
    finally:
        ops.CONFIG.update(old)


def _launch_count():
    from dsr_b200 import _lib
    return _lib.LAUNCHES


def test_adam_matches_oracle(ops):
    n = 1003
    p, g = torch.randn(n, generator=G(50)), torch.randn(n, generator=G(51))
    m, v = torch.zeros(n), torch.zeros(n)
    pc, mc, vc = p.clone().cuda(), m.clone().cuda(), v.clone().cuda()
    # arena slices are 16-byte aligned: allocate padded
    for step in (1, 2, 3):
        gs = g * step
        ref_ops.adam_update(p, gs, m, v, step, 1e-4)
        ops.adam_step(pc, gs.cuda(), mc, vc, 1e-4, step)
    assert torch.allclose(pc.cpu(), p, atol=1e-7) and torch.allclose(vc.cpu(), v, rtol=1e-5, atol=1e-12)


def test_cpu_tensor_raises(ops):
    with pytest.raises(RuntimeError):
        ops.hole_valid_masks(torch.zeros(1, 1, 4, 4))
    with pytest.raises(RuntimeError):
        ops.conv2d(torch.zeros(1, 3, 8, 8), torch.zeros(4, 3, 3, 3))


# ---------------------------------------------------------------------------------------------
# ragged sizes for the tiled / register-quad stencil kernels (csrc/stencil_tiled.cu): planes smaller than one 16 x 64
# tile, widths that are not a multiple of 4 (scalar load / store paths), sizes that end one pixel past a tile border
# ---------------------------------------------------------------------------------------------
RAGGED = [(2, 2), (3, 5), (16, 64), (17, 65), (5, 130), (33, 66), (31, 127), (48, 192)]


@pytest.mark.parametrize("HW", RAGGED)
def test_stencils_ragged_sizes(ops, HW):
    from dsr_b200.norms import camera_table
    H, W = HW
    B = 3
    d = torch.rand(B, 1, H, W, generator=G(90)) * 1.8 - 0.9
    d[torch.rand(B, 1, H, W, generator=G(91)) < 0.1] = -1.0
    # masks: bit-exact
    hole_ref, valid_ref = ref_ops.hole_valid_masks(d)
    hole, valid = ops.hole_valid_masks(d.cuda())
    assert torch.equal(hole.cpu(), hole_ref) and torch.equal(valid.cpu(), valid_ref)
    # image-space normals, forward + backward
    x = d.clone().requires_grad_(True)
    go = torch.randn(B, 3, H, W, generator=G(92))
    ref = ref_ops.surface_normals_old(x) * 100
    (ref * go).sum().backward()
    xc = d.cuda().requires_grad_(True)
    out = ops.normals_old(xc, 100.0)
    (out * go.cuda()).sum().backward()
    assert (out.cpu() - ref.detach()).abs().max() <= 1e-4 and rel_l2(xc.grad.cpu(), x.grad) <= 1e-4
    # camera-space normals (smooth depth: the backward is ill-conditioned on noise), forward + backward
    ds = _smooth_depth(B, max(H, 10), max(W, 10), 93)[:, :, :H, :W].clamp_min(-0.9).contiguous().requires_grad_(True)
    K = torch.tensor([[577.87, 0.7, 319.5], [0, 571.3, 239.5], [0, 0, 1]], dtype=torch.float64).repeat(B, 1, 1)
    crop = torch.tensor([[7, 7 + H, 3, 3 + W]] * B)
    refn = ref_ops.surface_normals_new(ds, K, crop)
    (refn * go).sum().backward()
    dc = ds.detach().cuda().requires_grad_(True)
    outn = ops.normals_new(dc, camera_table(K, crop, 0.5, "cuda"))
    (outn * go.cuda()).sum().backward()
    assert (outn.cpu() - refn.detach()).abs().max() <= 2e-6
    assert np.allclose(dc.grad.cpu().numpy(), ds.grad.numpy(), rtol=2e-3, atol=1e-3 * float(ds.grad.abs().median()) + 1e-12)
    # TV, forward + backward
    n3 = torch.randn(B, 3, H, W, generator=G(94)).requires_grad_(True)
    reft = ref_ops.tv_loss(n3)
    reft.backward()
    nc = n3.detach().cuda().requires_grad_(True)
    outt = ops.tv_loss(nc)
    outt.backward()
    assert abs(float(outt) - float(reft)) <= 1e-5 * max(float(reft), 1e-6) and rel_l2(nc.grad.cpu(), n3.grad) <= 1e-6
    # masked L1 / MSE, forward + backward
    a = torch.randn(B, 3, H, W, generator=G(95))
    b = torch.randn(B, 3, H, W, generator=G(96)).requires_grad_(True)
    m1 = (torch.rand(B, 1, H, W, generator=G(97)) < 0.7).float()
    m2 = (torch.rand(B, 1, H, W, generator=G(98)) < 0.5).float()
    l1, l2 = ref_ops.l1_mean(a * m1 * m2, b * m1 * m2), ref_ops.mse_mean(a * m1 * m2, b * m1 * m2)
    (l1 + 3 * l2).backward()
    bc = b.detach().cuda().requires_grad_(True)
    o = ops.masked_l1_l2(a.cuda(), bc, m1.cuda(), m2.cuda())
    (o[0] + 3 * o[1]).backward()
    assert abs(float(o[0]) - float(l1)) <= 1e-5 * float(l1) and abs(float(o[1]) - float(l2)) <= 1e-5 * float(l2)
    assert rel_l2(bc.grad.cpu(), b.grad) <= 1e-5


@pytest.mark.parametrize("HW", [(8, 12), (17, 65), (36, 132)])
def test_smooth_ragged_sizes(ops, HW):
    H, W = HW
    d = (torch.rand(2, 1, H, W, generator=G(13)) * 2 - 1).requires_grad_(True)
    img = torch.rand(2, 3, H, W, generator=G(14)) * 2 - 1
    ref = ref_ops.smooth_loss(d, img, 3)
    ref.backward()
    dc = d.detach().cuda().requires_grad_(True)
    out = ops.smooth_loss(dc, img.cuda(), 3)
    out.backward()
    assert abs(float(out) - float(ref)) <= 2e-5 * float(ref) and rel_l2(dc.grad.cpu(), d.grad) <= 1e-3


@pytest.mark.parametrize("HW", [(5, 7), (17, 65), (40, 130)])
def test_rect_holes_ragged_sizes(ops, HW):
    """rectangle tables drawn by the host in the reference's RNG order, planes that straddle tile borders"""
    H, W = HW
    B = 2
    d = torch.rand(B, 1, H, W, generator=G(80)) * 1.8 - 0.9
    d[torch.rand(B, 1, H, W, generator=G(81)) < 0.1] = -1.0
    _, valid = ref_ops.hole_valid_masks(d)
    rng = np.random.RandomState(5)
    rects = []
    for _ in range(B):
        n = rng.randint(3, 12)
        rects.append(np.stack([rng.randint(0, W, n), rng.randint(0, H, n), rng.randint(0, max(2, W // 2), n),
                               rng.randint(0, max(2, H // 2), n)], 1).astype(np.int64))
    gt_ref = ref_ops.rect_gt_mask(valid, rects)
    masked_ref = ref_ops.apply_gt_mask(d, gt_ref)
    tab = np.zeros((B, 64, 4), dtype=np.int32)
    cnt = np.zeros((B,), dtype=np.int32)
    for i, r in enumerate(rects):
        tab[i, :len(r)] = r
        cnt[i] = len(r)
    gt, masked, extra = ops.rect_holes(valid.cuda(), d.cuda(), torch.from_numpy(tab).cuda(), torch.from_numpy(cnt).cuda(), 64)
    assert np.array_equal(gt.cpu().numpy(), gt_ref.numpy().astype(np.uint8))
    assert torch.equal(masked.cpu(), masked_ref)


@pytest.mark.parametrize("B,C,h,w", [(3, 3, 37, 64), (2, 3, 130, 128), (2, 1, 5, 256), (5, 3, 64, 640), (1, 3, 1, 64), (2, 4, 9, 68),
                                     (7, 3, 200, 320)])
def test_smooth_row_ring_matches_register_kernels(ops, B, C, h, w):
    """csrc/stencil_ring.cu (bulk-copy row ring, the path large plane sets take) against the register kernels that the
    oracle tests pin: forward sums within fp32 summation-order noise, backward bit-for-bit up to sign-of-zero."""
    from dsr_b200.ops import _call, _p
    g = torch.Generator().manual_seed(17)
    d = (torch.rand(B, 1, h, w, generator=g) * 1.8 - 0.9).cuda()
    img = (torch.rand(B, C, h, w, generator=g) * 2 - 1).cuda()
    gs = torch.tensor(0.7, device="cuda")
    ref_s, ring_s = torch.zeros(2, dtype=torch.float64, device="cuda"), torch.zeros(2, dtype=torch.float64, device="cuda")
    assert B * h * w < (2 << 20)                  # small enough that the default entry point takes the register kernels
    _call("dsr_smooth_level_fwd", _p(d), _p(img), B, C, h, w, _p(ref_s, torch.float64))
    _call("dsr_smooth_level_fwd_ring", _p(d), _p(img), B, C, h, w, _p(ring_s, torch.float64))
    assert torch.allclose(ref_s, ring_s, rtol=1e-5, atol=1e-6), (ref_s, ring_s)
    for acc in (0, 1):
        ref_g, ring_g = torch.full_like(d, 0.25), torch.full_like(d, 0.25)
        _call("dsr_smooth_level_bwd", _p(d), _p(img), B, C, h, w, _p(gs), 0.3, 0.6, _p(ref_g), acc)
        _call("dsr_smooth_level_bwd_ring", _p(d), _p(img), B, C, h, w, _p(gs), 0.3, 0.6, _p(ring_g), acc)
        assert float((ref_g - ring_g).abs().max()) <= 1e-6, acc


def test_smooth_row_ring_rejects_unsuitable_shapes(ops):
    from dsr_b200.ops import _call, _p
    d, img = torch.zeros(1, 1, 8, 30, device="cuda"), torch.zeros(1, 3, 8, 30, device="cuda")
    with pytest.raises(RuntimeError):
        _call("dsr_smooth_level_fwd_ring", _p(d), _p(img), 1, 3, 8, 30, _p(torch.zeros(2, dtype=torch.float64, device="cuda"), torch.float64))



def test_wgrad2_matches_first_generation_kernel(ops):
    """csrc/wgrad_tc2.cu (8 column blocks per CTA, N = 256 MMAs, 32-pixel K tiles, vector reductions) against the
    first-generation weight-gradient kernel on the SAME single-pass bf16 operands: identical products, different summation
    order.  Shapes: every layer class of the step (3x3 / 7x7 pair / first layer with 32 rows / s2d down convs / transposed
    convs / 2x2 and 4x4 bottleneck levels), ragged edges, short last column-block groups."""
    old = dict(ops.CONFIG)
    cases = [("conv", 128, 128, 3, 1, 1, 0, 3, 40, 24), ("conv", 32, 128, 7, 1, 3, 0, 2, 36, 20), ("conv", 2, 32, 7, 1, 3, 0, 2, 32, 48),
             ("conv", 261, 64, 4, 2, 1, 0, 2, 32, 48), ("conv", 64, 128, 3, 2, 1, 0, 2, 24, 24), ("conv", 512, 512, 4, 2, 1, 0, 3, 4, 4),
             ("conv", 512, 512, 4, 2, 1, 0, 5, 8, 8), ("convT", 1024, 256, 4, 2, 1, 0, 2, 8, 8), ("convT", 512, 512, 4, 2, 1, 0, 12, 2, 2),
             ("convT", 128, 64, 3, 2, 1, 1, 2, 20, 12), ("conv", 256, 256, 3, 1, 1, 0, 2, 16, 16)]
    try:
        for case in cases:
            kind, Ci, Co, k, s, p, op, N, H, W = case
            x = cl(torch.randn(N, Ci, H, W, generator=G(280)))
            wshape = (Co, Ci, k, k) if kind == "conv" else (Ci, Co, k, k)
            w0 = torch.randn(wshape, generator=G(281)) * 0.05
            grads = []
            for kernel in (1, 2):
                ops.CONFIG.update(engine="tc", passes=3, dtype="f16", wgrad_passes=1, wgrad_kernel=kernel)
                xc, wc = x.clone().requires_grad_(True), w0.cuda().requires_grad_(True)
                out = ops.conv2d(xc, wc, None, s, p) if kind == "conv" else ops.conv_transpose2d(xc, wc, None, s, p, op)
                go = torch.randn(out.shape, generator=G(282)).cuda()
                (out * go).sum().backward()
                grads.append(wc.grad.detach().cpu())
            assert rel_l2(grads[1], grads[0]) <= 2e-6, (case, rel_l2(grads[1], grads[0]))
    finally:
        ops.CONFIG.update(old)


# weight packing: the shared-memory tiled kernel against the gather kernel, byte for byte (csrc/conv_tc.cu)
PACK_CASES = [  # variant, (D0, D1, R, S), Cp, T, Ca, Cout, phase
    (0, (256, 256, 3, 3), 256, 9, 256, 256, (0, 0)),          # CONV 3x3
    (0, (130, 200, 3, 3), 200, 9, 256, 130, (0, 0)),          # CONV, ragged channels / rows, K padding
    (2, (512, 256, 4, 4), 256, 4, 1024, 512, (0, 0)),         # CONV_S2D 4x4 stride 2
    (2, (128, 72, 3, 3), 80, 4, 320, 128, (0, 0)),            # CONV_S2D 3x3 (zero taps), Cp > Cin
    (4, (256, 128, 3, 3), 256, 9, 256, 128, (0, 0)),          # CONV_DGRAD (K channel = dim 0)
    (3, (512, 256, 4, 4), 512, 4, 512, 256, (1, 0)),          # CONVT_PH, one phase
    (3, (300, 100, 4, 4), 304, 4, 320, 100, (-1, -1)),        # CONVT_PH, all four phases stacked, ragged
    (3, (256, 128, 3, 3), 256, 4, 256, 128, (-1, -1)),        # CONVT_PH 3x3 (output_padding 1 form)
]


@pytest.mark.parametrize("kind,Ci,Co,k,stride,pad,H,W", [("conv", 128, 128, 3, 1, 1, 64, 64), ("conv", 64, 128, 4, 2, 1, 64, 48),
                                                      ("convT", 128, 64, 4, 2, 1, 32, 32), ("conv", 32, 128, 7, 1, 3, 48, 64)])
def test_wgrad_slabs_match_atomics_and_are_reproducible(ops, kind, Ci, Co, k, stride, pad, H, W):
    """dsr_tc_wgrad2p with partial = 1 (every K split stores its own slab, dsr_tc_unpack_wgrad_splits sums them in slab order)
    against the red.add form of the same GEMM: same products, fp32 summation order differs; and two runs of the slab form
    must be bit-identical (the atomics form is not).  Weight gradients of networks.py:379-415, :544-616."""
    old = dict(ops.CONFIG)
    try:
        x = torch.randn(4, Ci, H, W, generator=G(180))
        w = (torch.randn(*((Co, Ci, k, k) if kind == "conv" else (Ci, Co, k, k)), generator=G(181)) * 0.05)
        res, go = {}, None
        for name, slabs in (("atomics", False), ("slabs", True), ("slabs2", True)):
            ops.CONFIG.update(old)
            ops.CONFIG.update(engine="tc", passes=3, dtype="f16", wgrad_passes=1, wgrad_slabs=slabs)
            xc, wc = cl(x).requires_grad_(True), w.cuda().requires_grad_(True)
            with _CallLog() as names:
                out = ops.conv2d(xc, wc, None, stride, pad) if kind == "conv" else ops.conv_transpose2d(xc, wc, None, stride, pad, 0)
                if go is None:
                    go = torch.randn(out.shape, generator=G(182)).cuda()     # a FIXED dY: small forward GEMMs use split-K atomics
                (out * go).sum().backward()
            assert "dsr_tc_wgrad2p" in names
            res[name] = wc.grad.cpu()
        assert rel_l2(res["slabs"], res["atomics"]) <= 1e-6
        assert torch.equal(res["slabs"], res["slabs2"])
    finally:
        ops.CONFIG.update(old)


@pytest.mark.parametrize("case", PACK_CASES)
@pytest.mark.parametrize("f16", [1, 0])
def test_tiled_weight_pack_matches_gather_kernel(ops, case, f16):
    import os
    variant, shape, Cp, T, Ca, Cout, (pa, pb) = case
    D0, D1, R, S = shape
    w = (torch.randn(shape, generator=G(41)) * 0.02).cuda()
    rows = Cout * (4 if pa < 0 else 1)
    outs = []
    for tiled in ("1", "0"):
        os.environ["DSR_PACK_TILED"] = tiled
        hi = torch.full((rows, T * Ca), -1, device="cuda", dtype=torch.int16)
        lo = torch.full((rows, T * Ca), -1, device="cuda", dtype=torch.int16)
        ops._call("dsr_tc_pack_weight", ops._p(w), D0, D1, R, S, variant, Cp, pa, pb, 1, Cout, T, Ca, ops._p(hi, torch.int16),
                  ops._p(lo, torch.int16), f16, 64.0 if f16 else 1.0)
        outs.append((hi.cpu(), lo.cpu()))
    os.environ.pop("DSR_PACK_TILED", None)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert int((outs[0][0] != 0).sum()) > 0


@pytest.mark.parametrize("variant,shape,Cp,T,Ca", [(0, (256, 256, 3, 3), 256, 9, 256), (0, (130, 200, 3, 3), 200, 9, 256),
                                                   (2, (512, 256, 4, 4), 256, 4, 1024), (2, (128, 72, 3, 3), 80, 4, 320),
                                                   (0, (96, 48, 4, 4), 48, 16, 64)])
@pytest.mark.parametrize("accumulate", [0, 1])
def test_tiled_wgrad_unpack_matches_gather_kernel(ops, variant, shape, Cp, T, Ca, accumulate):
    import os
    D0, D1, R, S = shape
    dwp = torch.randn((D0, T * Ca), generator=G(43)).cuda()
    base = torch.randn(shape, generator=G(44)).cuda()
    outs = []
    for tiled in ("1", "0"):
        os.environ["DSR_PACK_TILED"] = tiled
        g = base.clone()
        ops._call("dsr_tc_unpack_wgrad", ops._p(dwp), D0, D1, R, S, variant, Cp, T, Ca, ops._p(g), accumulate)
        outs.append(g.cpu())
    os.environ.pop("DSR_PACK_TILED", None)
    assert torch.equal(outs[0], outs[1])
    assert not torch.equal(outs[0], base.cpu())
