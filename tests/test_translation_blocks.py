"""Building blocks of the translation_block row (SURVEY.md section 8f rank 3 / BASELINE configs[4]): the GroupNorm generator
and the PatchGAN discriminator, forward + backward with the LSGAN terms of translation_model.py:199-214, against golden
vectors of the live reference (tests/golden/gan_blocks_b2_64.npz) and the oracle."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_nets
from tests_proj import proj_vec
from util import cosine, load_golden, rel_l2

G_OPT = dict(ngf_img=32, ngf_depth=32, ngf=64, norm="group", dropout=False, init_type="normal", gpu_ids=[], input_nc_img=3,
             n_downsampling=2, use_semantic=False, n_blocks=9, upsampling_type="transpose", output_nc_depth=1, input_nc_depth=1)
D_OPT = dict(ndf=64, n_layers_D=3, norm_d="none", netD="n_layers", init_type="normal", gpu_ids=[], use_spnorm=False)


def _nets():
    from dsr_b200 import translation_network as tn
    torch.manual_seed(0)
    G = tn.define_Gen(SimpleNamespace(**G_OPT), input_type="img_depth")
    D = tn.define_D(SimpleNamespace(**D_OPT), input_type="depth")
    return G, D


def test_constructors_and_oracle_match_reference():
    g = load_golden("gan_blocks_b2_64.npz")
    G, D = _nets()
    for name, net in (("G", G), ("D", D)):
        sd = net.state_dict()
        assert list(sd.keys()) == list(g["wkeys/" + name])
        a = float(sum(v.double().abs().sum() for v in sd.values()))
        assert abs(a - float(g["wsum/" + name][0])) <= 1e-9 * a
    sd_g = {k: v.detach().clone().requires_grad_(True) for k, v in G.state_dict().items()}
    sd_d = {k: v.detach().clone().requires_grad_(True) for k, v in D.state_dict().items()}
    out = ref_nets.gan_block_step(sd_g, sd_d, torch.from_numpy(g["in/depth"]), torch.from_numpy(g["in/img"]),
                                  torch.from_numpy(g["in/real"]))
    assert rel_l2(out["fake"], g["fake"]) <= 2e-5 and rel_l2(out["pred_fake"], g["pred_fake"]) <= 2e-5
    assert abs(out["loss_G"] - float(g["loss_G"])) <= 2e-5 * float(g["loss_G"])
    assert abs(out["loss_D"] - float(g["loss_D"])) <= 2e-5 * float(g["loss_D"])
    for gi, n in enumerate(sd_g):
        gr = out["grads_g"][n].double().flatten()
        ref_norm, ref_proj = g["gG/" + n]
        assert abs(float(gr.norm()) - ref_norm) <= 2e-3 * max(ref_norm, 1e-12), n
        assert abs(float(gr @ proj_vec(gr.numel(), 2000 + gi)) - ref_proj) <= 2e-3 * max(ref_norm, 1e-12), n
    for gi, n in enumerate(sd_d):
        gr = out["grads_d"][n].double().flatten()
        ref_norm, ref_proj = g["gD/" + n]
        assert abs(float(gr.norm()) - ref_norm) <= 2e-3 * ref_norm and abs(float(gr @ proj_vec(gr.numel(), 3000 + gi)) - ref_proj) <= 2e-3 * ref_norm, n


@pytest.mark.gpu
@pytest.mark.parametrize("C", [64, 24, 256])        # 24: the scalar kernels (C/4 does not divide the block), others float4
@pytest.mark.parametrize("act,res", [(0, False), (1, False), (0, True)])
def test_group_norm_backward(built_lib, act, res, C):
    from dsr_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(3, C, 12, 10, generator=g) * 2 + 0.5).requires_grad_(True)
    w = (torch.randn(C, generator=g) * 0.5 + 1).requires_grad_(True)
    b = torch.randn(C, generator=g).requires_grad_(True)
    r = torch.randn(3, C, 12, 10, generator=g).requires_grad_(True) if res else None
    go = torch.randn(3, C, 12, 10, generator=g)
    ref = F.group_norm(x, 8, w, b, eps=1e-5)
    ref = F.relu(ref) if act else ref
    ref = ref + r if res else ref
    (ref * go).sum().backward()
    xc = x.detach().cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    wc, bc = w.detach().cuda().requires_grad_(True), b.detach().cuda().requires_grad_(True)
    rc = r.detach().cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True) if res else None
    out = ops.group_norm(xc, 8, wc, bc, 1e-5, act, rc)
    (out * go.cuda()).sum().backward()
    assert rel_l2(out.detach().cpu(), ref.detach()) <= 1e-5
    assert rel_l2(xc.grad.cpu(), x.grad) <= 2e-5 and rel_l2(wc.grad.cpu(), w.grad) <= 2e-5 and rel_l2(bc.grad.cpu(), b.grad) <= 2e-5
    if res:
        assert rel_l2(rc.grad.cpu(), r.grad) <= 1e-6


@pytest.mark.gpu
def test_generator_and_discriminator_train_like_the_reference(built_lib):
    from dsr_b200 import ops
    g = load_golden("gan_blocks_b2_64.npz")
    G, D = _nets()
    sd_g = {k: v.detach().clone().requires_grad_(True) for k, v in G.state_dict().items()}
    sd_d = {k: v.detach().clone().requires_grad_(True) for k, v in D.state_dict().items()}
    depth, img, real = (torch.from_numpy(g[k]) for k in ("in/depth", "in/img", "in/real"))
    ref = ref_nets.gan_block_step(sd_g, sd_d, depth, img, real)
    G, D = G.cuda(), D.cuda()
    one = torch.ones(2, 1, 6, 6, device="cuda")

    def mse_to(pred, value):                 # GANLoss('lsgan') = MSELoss against a constant map (translation_network.py:161-188)
        tgt = torch.full_like(pred, value)
        return ops.masked_l1_l2(tgt, pred, torch.ones((pred.shape[0], 1) + tuple(pred.shape[2:]), device=pred.device))[1]

    fake = G(depth.cuda(), img.cuda())
    pred_fake = D(fake)
    assert tuple(pred_fake.shape) == tuple(ref["pred_fake"].shape) == tuple(one.shape)
    loss_G = 0.5 * mse_to(pred_fake, 1.0)
    loss_G.backward()
    assert rel_l2(fake.detach().cpu(), g["fake"]) <= 1e-3 and rel_l2(pred_fake.detach().cpu(), g["pred_fake"]) <= 1e-3
    assert abs(float(loss_G) - float(g["loss_G"])) <= 1e-3 * float(g["loss_G"])
    fa, fb = [], []
    for n, prm in G.named_parameters():
        c = cosine(prm.grad.detach().cpu(), ref["grads_g"][n])
        assert c >= 0.999, (n, c)
        fa.append(prm.grad.detach().cpu().flatten()); fb.append(ref["grads_g"][n].flatten())
    assert cosine(torch.cat(fa), torch.cat(fb)) >= 0.9999
    D.zero_grad()
    loss_D = 0.5 * (mse_to(D(real.cuda()), 1.0) + mse_to(D(fake.detach()), 0.0))
    loss_D.backward()
    assert abs(float(loss_D) - float(g["loss_D"])) <= 1e-3 * float(g["loss_D"])
    for n, prm in D.named_parameters():
        c = cosine(prm.grad.detach().cpu(), ref["grads_d"][n])
        assert c >= 0.999, (n, c)
