"""Generate the golden fixtures under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference, read-only):
    python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs are committed as small .npz files; this
script is the provenance of every number in them.  Recipe = SURVEY.md Appendix D (imageio stub,
TrainOptions with a synthetic argv, MainModel on --gpu_ids -1).
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True
sys.modules["imageio"] = types.ModuleType("imageio")

import numpy as np
import torch

from oracle.ref_step import synthetic_batch

torch.set_num_threads(8)


def ref_opt(B, H, W):
    sys.argv = ["main.py", "--gpu_ids", "-1", "--image_and_depth", "--custom_pathes", "--use_image_for_trans",
                "--w_syn_l1", "15", "--w_real_l1_d", "40", "--norm_loss", "--w_syn_norm", "2",
                "--use_smooth_loss", "--w_smooth", "1", "--w_syn_holes", "800", "--w_real_holes", "1600",
                "--use_masked", "--use_scannet", "--lr", "0.0001", "--model", "main_network_best",
                "--batch_size", str(B), "--name", "golden", "--do_train", "--model_type", "main",
                "--checkpoints_dir", "/tmp/golden/ckpt", "--crop_size_h", str(H), "--crop_size_w", str(W)]
    from options.train_options import TrainOptions
    return TrainOptions().parse()


def proj_vec(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, generator=g, dtype=torch.float64)


def golden_step(B=2, H=128, W=128, depth_kind="smooth", tag="step_b2_128"):
    opt = ref_opt(B, H, W)
    from models.main_model import MainModel
    torch.manual_seed(0)
    np.random.seed(0)
    model = MainModel(opt)
    model.setup(opt)
    model._train()
    out = {}
    # checksums of the initial weights: pins "same seed + same constructor order => same weights"
    for name in model.model_names:
        sd = getattr(model, "net" + name).state_dict()
        out[f"wsum/{name}"] = np.array([float(sum(v.double().sum() for v in sd.values())),
                                        float(sum(v.double().abs().sum() for v in sd.values())),
                                        float(sum(v.numel() for v in sd.values()))])
        out[f"wkeys/{name}"] = np.array(list(sd.keys()))
        out[f"wshapes/{name}"] = np.array([str(tuple(v.shape)) for v in sd.values()])
    batch = synthetic_batch(B, H, W, seed=1, depth_kind=depth_kind)
    for k in ("A_i", "B_i", "A_d", "B_d"):
        out[f"in/{k}"] = batch[k].numpy()
    out["in/K"] = batch["K_A"].numpy()
    out["in/crop"] = batch["crop_A"].numpy()
    np.random.seed(0)
    for it in range(2):
        model.set_input(batch)
        model.optimize_parameters(it, 1)
        p = f"s{it}/"
        for k, v in model.get_current_losses().items():
            out[p + "loss/" + k] = np.float64(v)
        out[p + "loss/G"] = np.float64(float(model.loss_G))
        out[p + "loss/mean_of_abs_diff_syn"] = np.float64(model.loss_mean_of_abs_diff_syn)
        out[p + "loss/mean_of_abs_diff_real"] = np.float64(model.loss_mean_of_abs_diff_real)
        for k in ("pred_syn_depth", "pred_real_depth", "syn2real_depth", "syn_depth_by_image",
                  "real_depth_by_image", "depth_masked", "syn2real_depth_masked"):
            out[p + k] = getattr(model, k).detach().numpy().astype(np.float32)
        if it == 0:
            for k in ("syn_mask", "real_mask", "real_hole_mask"):
                out[p + k] = getattr(model, k).numpy().astype(np.uint8)
            out[p + "gt_mask_syn"] = model.gt_mask_syn.numpy().astype(np.uint8)
            out[p + "gt_mask_real"] = model.gt_mask_real.numpy().astype(np.uint8)
            out[p + "norm_syn_pred"] = model.norm_syn_pred.detach().numpy().astype(np.float16)
            out[p + "norm_real"] = model.norm_real.detach().numpy().astype(np.float16)
            # gradients: per tensor (L2 norm, projection on a fixed random vector), a few in full
            gi = 0
            for net in ("Depth_f", "Task"):
                for n, prm in getattr(model, "net" + net).named_parameters():
                    g = prm.grad.detach().double().flatten()
                    out[p + f"gstat/{net}/{n}"] = np.array([float(g.norm()), float(g @ proj_vec(g.numel(), 1000 + gi))])
                    gi += 1
                    if g.numel() <= 8192:
                        out[p + f"gfull/{net}/{n}"] = prm.grad.detach().numpy()
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print("wrote", tag, {k: float(v) for k, v in out.items() if k.startswith("s0/loss/")})


def golden_sr_step(B=1, h=128, w=128, tag="sr_step_b1_128"):
    """One + one MainSRModel steps (models/main_sr_model.py) on a B=1, LR 128x128 -> HR 256x256 synthetic batch,
    flags of README.md:86.  The reference hard-codes gpu_ids=[0,1,2,3] for G_A_d (main_sr_model.py:166), so
    translation_network.init_net is wrapped to force gpu_ids=[]."""
    from oracle.ref_step import synthetic_sr_batch
    sys.argv = ["main.py", "--gpu_ids", "-1", "--image_and_depth", "--custom_pathes", "--use_image_for_trans",
                "--w_syn_l1", "15", "--w_real_l1_d", "90", "--norm_loss", "--w_syn_norm", "3",
                "--use_smooth_loss", "--w_smooth", "1", "--w_syn_holes", "1600", "--w_real_holes", "1600",
                "--use_masked", "--use_scannet", "--lr", "0.00002", "--model", "main_network_best",
                "--batch_size", str(B), "--name", "golden_sr", "--do_train", "--model_type", "main", "--SR",
                "--checkpoints_dir", "/tmp/golden/ckpt", "--crop_size_h", str(h), "--crop_size_w", str(w)]
    from options.train_options import TrainOptions
    opt = TrainOptions().parse()
    from models import translation_network as tn
    orig = tn.init_net
    tn.init_net = lambda net, init_type="normal", init_gain="relu", gpu_ids=[], param=None: orig(net, init_type, init_gain, [], param)
    from models.main_sr_model import MainSRModel
    torch.manual_seed(0)
    np.random.seed(0)
    model = MainSRModel(opt)
    model.setup(opt)
    model._train()
    out = {}
    for name in model.model_names:
        sd = getattr(model, "net" + name).state_dict()
        out[f"wsum/{name}"] = np.array([float(sum(v.double().sum() for v in sd.values())),
                                        float(sum(v.double().abs().sum() for v in sd.values())),
                                        float(sum(v.numel() for v in sd.values()))])
    batch = synthetic_sr_batch(B, h, w, seed=1, depth_kind="smooth")
    np.random.seed(0)
    for it in range(2):
        model.set_input(batch)
        model.optimize_parameters(it, 1)
        p = f"s{it}/"
        for k, v in model.get_current_losses().items():
            out[p + "loss/" + k] = np.float64(v)
        out[p + "loss/G"] = np.float64(float(model.loss_G))
        out[p + "loss/mean_of_abs_diff_syn"] = np.float64(model.loss_mean_of_abs_diff_syn)
        out[p + "loss/mean_of_abs_diff_real"] = np.float64(model.loss_mean_of_abs_diff_real)
        for k in ("pred_syn_depth", "pred_real_depth", "pred_real_depth_hr"):
            out[p + k] = getattr(model, k).detach().numpy().astype(np.float32)
        if it == 0:
            for k in ("syn2real_depth", "syn_depth_by_image", "real_depth_by_image", "real_depth"):
                out[p + k] = getattr(model, k).detach().numpy().astype(np.float16)
            for k in ("syn_mask", "real_mask", "real_hole_mask"):
                out[p + k] = getattr(model, k).numpy().astype(np.uint8)
            out[p + "gt_mask_syn"] = model.gt_mask_syn.numpy().astype(np.uint8)
            out[p + "gt_mask_real"] = model.gt_mask_real.numpy().astype(np.uint8)
            gi = 0
            for net in ("Depth_f", "Task"):
                for n, prm in getattr(model, "net" + net).named_parameters():
                    g = prm.grad.detach().double().flatten()
                    out[p + f"gstat/{net}/{n}"] = np.array([float(g.norm()), float(g @ proj_vec(g.numel(), 1000 + gi))])
                    gi += 1
    tn.init_net = orig
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print("wrote", tag, {k: float(v) for k, v in out.items() if k.startswith("s0/loss/")})


def golden_i2d_step(B=2, H=128, W=128, tag="i2d_step_b2_128"):
    """Two I2DModel steps (models/I2D_model.py), flags of README.md:28, B = 2 at 128x128."""
    sys.argv = ["main.py", "--gpu_ids", "-1", "--image_and_depth", "--custom_pathes", "--w_real_l1", "1", "--w_syn_l1", "1",
                "--lr", "0.0002", "--Imagef_outf", "128", "--Imagef_basef", "32", "--use_scannet", "--model", "I2D",
                "--norm_loss", "--do_train", "--batch_size", str(B), "--name", "golden_i2d",
                "--checkpoints_dir", "/tmp/golden/ckpt", "--crop_size_h", str(H), "--crop_size_w", str(W)]
    from options.train_options import TrainOptions
    opt = TrainOptions().parse()
    from models.I2D_model import I2DModel
    torch.manual_seed(0)
    np.random.seed(0)
    model = I2DModel(opt)
    model.setup(opt)
    model._train()
    out = {}
    for name in model.model_names:
        sd = getattr(model, "net" + name).state_dict()
        out[f"wsum/{name}"] = np.array([float(sum(v.double().sum() for v in sd.values())),
                                        float(sum(v.double().abs().sum() for v in sd.values())),
                                        float(sum(v.numel() for v in sd.values()))])
        out[f"wkeys/{name}"] = np.array(list(sd.keys()))
    batch = synthetic_batch(B, H, W, seed=1, depth_kind="smooth")
    for it in range(2):
        model.set_input(batch)
        model.optimize_parameters(it)
        p = f"s{it}/"
        for k, v in model.get_current_losses().items():
            out[p + "loss/" + k] = np.float64(v)
        out[p + "loss/G"] = np.float64(float(model.loss_G))
        for k in ("pred_syn_depth", "pred_real_depth"):
            out[p + k] = getattr(model, k).detach().numpy().astype(np.float32)
        if it == 0:
            gi = 0
            for n, prm in model.netTask.named_parameters():
                g = prm.grad.detach().double().flatten()
                out[p + f"gstat/Task/{n}"] = np.array([float(g.norm()), float(g @ proj_vec(g.numel(), 1000 + gi))])
                gi += 1
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print("wrote", tag, {k: float(v) for k, v in out.items() if k.startswith("s0/loss/")})


def metric_cases():
    """synthetic uint16-valued depth triples (input with holes, prediction, target with holes) + ScanNet intrinsics"""
    rs = np.random.RandomState(21)
    cases = []
    for (H, W) in ((48, 64), (33, 47)):
        yy, xx = np.mgrid[0:H, 0:W]
        target = np.round(1500 + 20 * xx + 12 * yy + 300 * np.sin(xx / 9.0) + rs.rand(H, W) * 15).astype(np.float64)
        target[5:9, 10:22] = 0
        target[rs.rand(H, W) < 0.02] = 0
        target[0, :3] = 6000                                  # clipped to max_depth
        pred = np.round(target + rs.randn(H, W) * 25 + 10).clip(0, 65535)
        pred[target == 0] = np.round(1400 + rs.rand(int((target == 0).sum())) * 100)
        inp = target.copy()
        inp[20:30, 30:44] = 0                                 # holes of the input that the target does not have
        inp[rs.rand(H, W) < 0.05] = 0
        cases.append((pred, target, inp))
    K = np.array([[577.87, 0, 319.5], [0, 577.87, 239.5], [0, 0, 1.0]])
    return cases, K


def golden_metrics():
    """new_metrics.calc_metrics of the live reference on synthetic triples (albumentations / skimage / imageio / tqdm are
    stubbed: calc_metrics itself only needs numpy, scipy and torch)."""
    for name in ("albumentations", "tqdm", "skimage", "skimage.transform"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["skimage.transform"].resize = lambda a, shape: a
    import new_metrics as nm
    cases, K = metric_cases()
    names = ["rmse", "mae", "rmse_h", "rmse_d", "psnr", "ssim", "mae_h", "mae_d", "mse_v"]
    out = {"names": np.array(names)}
    for i, (pred, target, inp) in enumerate(cases):
        p, t = pred.clip(0, 5100), target.clip(0, 5100)
        r = nm.calc_metrics(p, t, inp < nm.holes_threshold, t < nm.holes_threshold, K, 5100, names)
        out[f"c{i}/pred"], out[f"c{i}/target"], out[f"c{i}/input"] = pred, target, inp
        out[f"c{i}/values"] = np.array([float(r[n]) for n in names])
    out["K"] = K
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
    print("wrote metrics", {n: float(v) for n, v in zip(names, out["c0/values"])})


def golden_gan_blocks(B=2, H=64, W=64, tag="gan_blocks_b2_64"):
    """Generator (img_depth, GroupNorm) + NLayerDiscriminator (norm_d none) of the live reference, forward + backward with the
    LSGAN terms of translation_model.py:199-214 - the compute of BASELINE configs[4]."""
    from types import SimpleNamespace
    import torch.nn as nn
    from models import translation_network as tn
    torch.manual_seed(0)
    og = SimpleNamespace(ngf_img=32, ngf_depth=32, ngf=64, norm="group", dropout=False, init_type="normal", gpu_ids=[],
                         input_nc_img=3, n_downsampling=2, use_semantic=False, n_blocks=9, upsampling_type="transpose",
                         output_nc_depth=1, input_nc_depth=1)
    G = tn.define_Gen(og, input_type="img_depth")
    od = SimpleNamespace(ndf=64, n_layers_D=3, norm_d="none", netD="n_layers", init_type="normal", gpu_ids=[], use_spnorm=False)
    D = tn.define_D(od, input_type="depth")
    g = torch.Generator().manual_seed(3)
    depth = torch.rand(B, 1, H, W, generator=g) * 1.8 - 0.9
    img = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    real = torch.rand(B, 1, H, W, generator=g) * 1.8 - 0.9
    crit = tn.GANLoss("lsgan")
    fake = G(depth, img)
    pred_fake = D(fake)
    loss_G = 0.5 * crit(pred_fake, True)                                   # translation_model.py:214
    loss_G.backward()
    out = {"in/depth": depth.numpy(), "in/img": img.numpy(), "in/real": real.numpy(),
           "fake": fake.detach().numpy(), "pred_fake": pred_fake.detach().numpy(), "loss_G": np.float64(float(loss_G))}
    for name, net in (("G", G), ("D", D)):
        sd = net.state_dict()
        out[f"wkeys/{name}"] = np.array(list(sd.keys()))
        out[f"wsum/{name}"] = np.array([float(sum(v.double().abs().sum() for v in sd.values()))])
    for gi, (n, prm) in enumerate(G.named_parameters()):
        gr = prm.grad.detach().double().flatten()
        out[f"gG/{n}"] = np.array([float(gr.norm()), float(gr @ proj_vec(gr.numel(), 2000 + gi))])
    D.zero_grad()
    loss_D = 0.5 * (crit(D(real), True) + crit(D(fake.detach()), False))   # translation_model.py:199-205
    loss_D.backward()
    out["loss_D"] = np.float64(float(loss_D))
    for gi, (n, prm) in enumerate(D.named_parameters()):
        gr = prm.grad.detach().double().flatten()
        out[f"gD/{n}"] = np.array([float(gr.norm()), float(gr @ proj_vec(gr.numel(), 3000 + gi))])
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print("wrote", tag, float(loss_G), float(loss_D))


def translation_batch(B, H, W, seed=4):
    g = torch.Generator().manual_seed(seed)
    d = lambda: torch.rand(B, 1, H, W, generator=g) * 1.6 - 0.7
    A_d, B_d = d(), d()
    A_d[:, :, 5:9, 10:20] = -1.0                       # holes of the A domain (hole_mask_A = depth <= -0.98)
    return dict(A_name=["a"] * B, B_name=["b"] * B, A_img=torch.rand(B, 3, H, W, generator=g) * 2 - 1, A_depth=A_d,
                B_img=torch.rand(B, 3, H, W, generator=g) * 2 - 1, B_depth=B_d)


def golden_translation_step(B=1, H=64, W=64, n_gen=2, tag="translation_step_b1_64", extra_flags=(), extra_losses=()):
    """One TranslationModel.optimize_parameters (models/translation_model.py:274-291) of the live reference, default loss flags
    (plus `extra_flags`), --num_iter_gen 2, normal init; the loop is unrolled here to record the first generator iteration's
    gradients."""
    sys.argv = ["main.py", "--gpu_ids", "-1", "--custom_pathes", "--use_scannet", "--lr", "0.0002", "--model", "translation_block",
                "--batch_size", str(B), "--name", "golden_tr", "--netD", "n_layers", "--crop_size_h", str(H), "--crop_size_w", str(W),
                "--do_train", "--max_distance", "5100", "--init_type", "normal", "--model_type", "translation",
                "--num_iter_gen", str(n_gen), "--checkpoints_dir", "/tmp/golden/ckpt"] + list(extra_flags)
    from options.train_options import TrainOptions
    opt = TrainOptions().parse()
    from models.translation_model import TranslationModel
    torch.manual_seed(0)
    model = TranslationModel(opt)
    model.setup(opt)
    batch = translation_batch(B, H, W)
    out = {}
    nets = ["G_A", "G_B", "D_A_depth", "D_B_depth", "D_A_normal", "D_B_normal"]
    for name in nets:
        sd = getattr(model, "net" + name).state_dict()
        out[f"wkeys/{name}"] = np.array(list(sd.keys()))
        out[f"wsum/{name}"] = np.array([float(sum(v.double().abs().sum() for v in sd.values()))])
    model.set_input(batch)
    model.set_requires_grad(model.disc, False)
    for it in range(opt.num_iter_gen):
        model.forward()
        model.zero_grad([model.netG_A, model.netG_B])
        model.backward_G()
        if it == 0:
            for k in ("fake_depth_B", "fake_depth_A", "rec_depth_B", "idt_B", "fake_norm_B", "real_norm_A"):
                out["s0/" + k] = getattr(model, k).detach().numpy()
            for k in ("G_A", "G_B", "cycle_B", "cycle_n_B", "idt_B", "depth_range_A", "depth_range_B") + tuple(extra_losses):
                out["s0/loss/" + k] = np.float64(float(getattr(model, "loss_" + k)))
            out["s0/loss/G"] = np.float64(float(model.loss_G))
            if "cycle_A" in extra_losses:
                out["s0/rec_depth_A"] = model.rec_depth_A.detach().numpy()
            gi = 0
            for name in ("G_A", "G_B"):
                for n, prm in getattr(model, "net" + name).named_parameters():
                    gr = prm.grad.detach().double().flatten()
                    out[f"s0/g/{name}/{n}"] = np.array([float(gr.norm()), float(gr @ proj_vec(gr.numel(), 4000 + gi))])
                    gi += 1
        model.optimizer_G.step()
    model.set_requires_grad(model.disc, True)
    model.set_requires_grad([model.netG_A, model.netG_B], False)
    model.zero_grad(model.disc)
    model.backward_D_A()
    model.backward_D_B()
    gi = 0
    for name in ("D_A_depth", "D_A_normal", "D_B_depth", "D_B_normal"):
        out["end/loss/" + name] = np.float64(float(getattr(model, "loss_" + name)))
        for n, prm in getattr(model, "net" + name).named_parameters():
            gr = prm.grad.detach().double().flatten()
            out[f"end/g/{name}/{n}"] = np.array([float(gr.norm()), float(gr @ proj_vec(gr.numel(), 5000 + gi))])
            gi += 1
    model.optimizer_D.step()
    for k in ("G_A", "G_B", "cycle_B", "cycle_n_B", "idt_B", "depth_range_A", "depth_range_B") + tuple(extra_losses):
        out["end/loss/" + k] = np.float64(float(getattr(model, "loss_" + k)))
    out["end/fake_depth_B"] = model.fake_depth_B.detach().numpy()
    for name in nets:                                   # weights after the step: pins both Adam variants
        sd = getattr(model, "net" + name).state_dict()
        v = torch.cat([t.double().flatten() for t in sd.values()])
        out[f"end/w/{name}"] = np.array([float(v.norm()), float(v @ proj_vec(v.numel(), 6000))])
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    print("wrote", tag, {k: float(v) for k, v in out.items() if k.startswith("s0/loss/")})


def golden_resize():
    """F.interpolate bicubic / nearest vectors (the torch calls of main_sr_model.py:279-293, :361, :394-398)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2, 3, 24, 36, generator=g) * 2 - 1
    out = {"x": x.numpy()}
    for name, size in (("down", (12, 18)), ("up", (48, 72)), ("odd", (10, 50))):
        out["bicubic_" + name] = F.interpolate(x, size=size, mode="bicubic").numpy()
        out["nearest_" + name] = F.interpolate(x, size=size, mode="nearest").numpy()
    xg = x.clone().requires_grad_(True)
    gy = torch.rand(2, 3, 12, 18, generator=g)
    (F.interpolate(xg, size=(12, 18), mode="bicubic") * gy).sum().backward()
    out["gy_down"], out["gx_down"] = gy.numpy(), xg.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "resize.npz"), **out)
    print("wrote resize")


def golden_ops():
    """Per-op vectors from the reference's own functions (adversarial: skewed K, crop offset,
    non-square, all-hole rows, borders)."""
    from models import norms as rn
    from models import main_model as rm
    from models import pytorch_ssim as rs
    out = {}
    g = torch.Generator().manual_seed(7)
    B, H, W = 2, 40, 56
    d = torch.rand(B, 1, H, W, generator=g) * 1.8 - 0.9
    d[:, :, 5:9, 10:30] = -1.0
    d[0, 0, 0, :] = -1.0
    d[1, 0, :, W - 1] = -0.97
    d[1, 0, 20, 20] = -0.9700001
    img = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    K = torch.tensor([[[577.87, 0.7, 319.5], [0, 571.3, 239.5], [0, 0, 1]],
                      [[600.0, 0, 320.0], [0, 600.0, 240.0], [0, 0, 1]]], dtype=torch.float64)
    crop = torch.tensor([[64, 64 + H, 100, 100 + W], [0, H, 5, 5 + W]])
    out["d"], out["img"], out["K"], out["crop"] = d.numpy(), img.numpy(), K.numpy(), crop.numpy()
    out["normals_old"] = rn.SurfaceNormals()(d).numpy()
    out["normals_new"] = rn.SurfaceNormals_new()(d, K, crop).numpy()
    out["tv"] = np.float64(rm.tv_loss(out_t := rn.SurfaceNormals()(d) * 100))
    d2 = torch.rand(B, 1, 64, 96, generator=g) * 2 - 1
    im2 = torch.rand(B, 3, 64, 96, generator=g) * 2 - 1
    out["smooth_d"], out["smooth_img"] = d2.numpy(), im2.numpy()
    out["smooth"] = np.float64(rm.get_smooth_weight(d2, im2, 3))
    a = torch.rand(2, 3, 33, 47, generator=g)
    b = (a + 0.1 * torch.randn(2, 3, 33, 47, generator=g)).clamp(0, 1)
    out["ssim_a"], out["ssim_b"] = a.numpy(), b.numpy()
    out["ssim"] = np.float64(rs.ssim(a, b))
    # hole / valid masks exactly as main_model.py:208-230 computes them
    one, zero = torch.tensor(1).float(), torch.tensor(0).float()
    holl = torch.where(d <= -0.97, one, zero)
    r = holl.clone()
    r[:, :, :-1, :] += r[:, :, 1:, :].clone()
    r[:, :, 1:, :] += r[:, :, :-1, :].clone()
    r[:, :, :, :-1] += r[:, :, :, 1:].clone()
    r[:, :, :, 1:] += r[:, :, :, :-1].clone()
    out["hole"] = holl.numpy().astype(np.uint8)
    out["valid"] = torch.where(r < 1, one, zero).numpy().astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **out)
    print("wrote ops")


if __name__ == "__main__":
    which = sys.argv[1:] or ["ops", "step", "resize", "sr", "i2d", "metrics", "gan", "translation", "translation_tv", "translation_flags",
                             "translation_inpB"]
    sys.argv = sys.argv[:1]
    if "ops" in which:
        golden_ops()
    if "step" in which:
        golden_step()
    if "resize" in which:
        golden_resize()
    if "sr" in which:
        golden_sr_step()
    if "i2d" in which:
        golden_i2d_step()
    if "metrics" in which:
        golden_metrics()
    if "gan" in which:
        golden_gan_blocks()
    if "translation" in which:
        golden_translation_step()
    if "translation_tv" in which:
        # the one optional term the CUDA model wires so far (translation_model.py:247-249)
        golden_translation_step(tag="translation_tv_b1_64", extra_flags=["--l_tv_A", "2.0"], extra_losses=("tv_norm_A",))
    if "translation_flags" in which:
        # the optional loss terms (translation_model.py:222-249): cycle A (masked L1 + masked cosine), mean differences, TV of normals
        golden_translation_step(tag="translation_flags_b1_64",
                                extra_flags=["--use_cycle_A", "--l_mean_A", "0.5", "--l_mean_B", "0.7", "--l_tv_A", "2.0"],
                                extra_losses=("cycle_A", "cycle_n_A", "mean_dif_A", "mean_dif_B", "tv_norm_A"))
    if "translation_inpB" in which:
        # G_B as a depth-only generator (translation_model.py:146-147, :167-168, :185-186), together with cycle A through it
        golden_translation_step(tag="translation_inpB_b1_64", extra_flags=["--inp_B", "depth", "--use_cycle_A"],
                                extra_losses=("cycle_A", "cycle_n_A"))
