"""Oracle (CPU) restatement of one ``MainModel.optimize_parameters`` step.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  Follows models/main_model.py:204-429:
forward through the five nets, rectangle-hole masks drawn from ``np.random`` in the reference's
order (real loop first, then syn: main_model.py:257-298), the loss stack, autograd through Task and
Depth_f only, Adam.  Also the CPU baseline timed by ``bench.py`` (``cpu_baseline.kind == "port"``).
"""
import collections

import numpy as np
import torch

from . import ref_nets, ref_ops

NET_NAMES = ("G_A_d", "I2D_features", "Image2Depth", "Task", "Depth_f")   # main_model.py:127
TRAINABLE = ("Depth_f", "Task")                                           # main_model.py:176 (chain order)


class OracleStep:
    """Holds the five state_dicts (fp32 CPU) + Adam state; ``step(batch)`` = one training step."""

    def __init__(self, state_dicts, lr=1e-4, weights=None):
        self.sd = {k: collections.OrderedDict((n, t.detach().clone().float()) for n, t in v.items())
                   for k, v in state_dicts.items()}
        self.lr = lr
        self.weights = weights
        self.n_step = 0
        self.adam = {}
        for net in TRAINABLE:
            for n, p in self.sd[net].items():
                p.requires_grad_(True)
                self.adam[(net, n)] = (torch.zeros_like(p), torch.zeros_like(p))

    # main_model.py:204-318
    def forward(self, batch, stage="train"):
        t = {}
        t["syn_image"], t["real_image"] = batch["A_i"].float(), batch["B_i"].float()
        t["syn_depth"], t["real_depth"] = batch["A_d"].float(), batch["B_d"].float()
        t["K_A"], t["K_B"], t["crop_A"], t["crop_B"] = batch["K_A"], batch["K_B"], batch["crop_A"], batch["crop_B"]
        t["real_hole_mask"], t["real_mask"] = ref_ops.hole_valid_masks(t["real_depth"])
        _, t["syn_mask"] = ref_ops.hole_valid_masks(t["syn_depth"])
        with torch.no_grad():
            t["syn2real_depth"] = ref_nets.translation_generator(self.sd["G_A_d"], t["syn_depth"], t["syn_image"])
            f_syn = ref_nets.resnet_generator(self.sd["I2D_features"], t["syn_image"])
            f_real = ref_nets.resnet_generator(self.sd["I2D_features"], t["real_image"])
            t["syn_depth_by_image"] = ref_nets.unet_generator(self.sd["Image2Depth"], f_syn)
            t["real_depth_by_image"] = ref_nets.unet_generator(self.sd["Image2Depth"], f_real)
        B, _, H, W = t["real_depth"].shape
        t["rects_real"] = ref_ops.draw_rects(B, H, W, stage)
        t["gt_mask_real"] = ref_ops.rect_gt_mask(t["real_mask"], t["rects_real"])
        t["depth_masked"] = ref_ops.apply_gt_mask(t["real_depth"], t["gt_mask_real"])
        t["rects_syn"] = ref_ops.draw_rects(B, H, W, stage)
        t["gt_mask_syn"] = ref_ops.rect_gt_mask(t["syn_mask"], t["rects_syn"])
        t["syn2real_depth_masked"] = ref_ops.apply_gt_mask(t["syn2real_depth"], t["gt_mask_syn"])
        in_s = torch.cat([t["syn2real_depth_masked"], t["syn_depth_by_image"]], 1)
        in_r = torch.cat([t["depth_masked"], t["real_depth_by_image"]], 1)
        fd_s = ref_nets.resnet_generator(self.sd["Depth_f"], in_s)
        fd_r = ref_nets.resnet_generator(self.sd["Depth_f"], in_r)
        t["pred_syn_depth"] = ref_nets.unet_generator(self.sd["Task"], torch.cat([f_syn, fd_s, in_s, t["syn_image"]], 1))
        t["pred_real_depth"] = ref_nets.unet_generator(self.sd["Task"], torch.cat([f_real, fd_r, in_r, t["real_image"]], 1))
        t["monitor"] = ref_ops.monitor_scalars(t)
        return t

    # main_model.py:422-429
    def step(self, batch, stage="train", update=True):
        for net in TRAINABLE:
            for p in self.sd[net].values():
                p.grad = None
        t = self.forward(batch, stage)
        loss_G, terms, visuals = ref_ops.loss_stack(t, self.weights)
        loss_G.backward()
        grads = {(net, n): p.grad.detach().clone() for net in TRAINABLE for n, p in self.sd[net].items()}
        if update:
            self.n_step += 1
            with torch.no_grad():
                for net in TRAINABLE:
                    for n, p in self.sd[net].items():
                        m, v = self.adam[(net, n)]
                        ref_ops.adam_update(p, p.grad, m, v, self.n_step, self.lr)
        losses = {k: float(v) for k, v in terms.items()}
        losses.update(t["monitor"])
        losses["G"] = float(loss_G)
        return dict(tensors=t, losses=losses, visuals=visuals, grads=grads)


class OracleSRStep(OracleStep):
    """One ``MainSRModel.optimize_parameters`` step (models/main_sr_model.py:228-497).  ``lr_size`` =
    (opt.crop_size_h, opt.crop_size_w); the batch holds HR (2h x 2w) tensors."""

    def __init__(self, state_dicts, lr_size, lr=2e-5, weights=None):
        super().__init__(state_dicts, lr=lr, weights=weights)
        self.lr_size = tuple(lr_size)

    def forward(self, batch, stage="train"):
        h, w = self.lr_size
        t = {}
        t["syn_image"], t["real_image"] = batch["A_i"].float(), batch["B_i"].float()
        t["syn_depth"], t["real_depth"] = batch["A_d"].float(), batch["B_d"].float()
        t["K_A"], t["K_B"], t["crop_A"], t["crop_B"] = batch["K_A"], batch["K_B"], batch["crop_A"], batch["crop_B"]
        t["real_hole_mask"], t["real_mask"] = ref_ops.hole_valid_masks(t["real_depth"])
        _, t["syn_mask"] = ref_ops.hole_valid_masks(t["syn_depth"])
        B, _, H, W = t["real_depth"].shape
        with torch.no_grad():
            t["syn2real_depth"] = ref_nets.translation_generator(self.sd["G_A_d"], t["syn_depth"], t["syn_image"])
            f_real = ref_nets.resnet_generator(self.sd["I2D_features"], ref_ops.bicubic(t["real_image"], (h, w)))   # :279
            t["real_depth_by_image"] = ref_ops.bicubic(ref_nets.unet_generator(self.sd["Image2Depth"], f_real), (H, W))
            f_real = ref_ops.bicubic(f_real, (H, W))
            f_syn = ref_nets.resnet_generator(self.sd["I2D_features"], ref_ops.bicubic(t["syn_image"], (h, w)))     # :285
            t["syn_depth_by_image"] = ref_ops.bicubic(ref_nets.unet_generator(self.sd["Image2Depth"], f_syn), (H, W))
            f_syn = ref_ops.bicubic(f_syn, (H, W))
        t["rects_real"] = ref_ops.draw_rects(B, H, W, stage, p_train=0.95, div=10)           # :298-305
        t["gt_mask_real"] = ref_ops.rect_gt_mask(t["real_mask"], t["rects_real"])
        t["depth_masked"] = ref_ops.apply_gt_mask(t["real_depth"], t["gt_mask_real"])
        t["rects_syn"] = ref_ops.draw_rects(B, H, W, stage, p_train=0.90, div=10)            # :319-326
        t["gt_mask_syn"] = ref_ops.rect_gt_mask(t["syn_mask"], t["rects_syn"])
        t["syn2real_depth_masked"] = ref_ops.apply_gt_mask(t["syn2real_depth"], t["gt_mask_syn"])
        in_r = torch.cat([t["depth_masked"], t["real_depth_by_image"]], 1)
        fd_r = ref_nets.resnet_generator(self.sd["Depth_f"], in_r)
        t["pred_real_depth_hr"] = ref_nets.unet_generator(self.sd["Task"], torch.cat([f_real, fd_r, in_r, t["real_image"]], 1))
        in_s = torch.cat([t["syn2real_depth_masked"], t["syn_depth_by_image"]], 1)
        fd_s = ref_nets.resnet_generator(self.sd["Depth_f"], in_s)
        t["pred_syn_depth"] = ref_nets.unet_generator(self.sd["Task"], torch.cat([f_syn, fd_s, in_s, t["syn_image"]], 1))
        t["pred_real_depth"] = ref_ops.bicubic(t["pred_real_depth_hr"], (h, w))                # :361
        t["monitor"] = ref_ops.monitor_scalars_sr(t, (h, w))
        return t

    def step(self, batch, stage="train", update=True):
        for net in TRAINABLE:
            for p in self.sd[net].values():
                p.grad = None
        t = self.forward(batch, stage)
        loss_G, terms, visuals = ref_ops.loss_stack_sr(t, self.lr_size, self.weights)
        loss_G.backward()
        grads = {(net, n): p.grad.detach().clone() for net in TRAINABLE for n, p in self.sd[net].items()}
        if update:
            self.n_step += 1
            with torch.no_grad():
                for net in TRAINABLE:
                    for n, p in self.sd[net].items():
                        m, v = self.adam[(net, n)]
                        ref_ops.adam_update(p, p.grad, m, v, self.n_step, self.lr)
        losses = {k: float(v) for k, v in terms.items()}
        losses.update(t["monitor"])
        losses["G"] = float(loss_G)
        return dict(tensors=t, losses=losses, visuals=visuals, grads=grads)


class OracleI2DStep:
    """One ``I2DModel.optimize_parameters`` step (models/I2D_model.py:163-250): Image_f + Task on the syn and the real
    image, masked L1 (mask = depth >= -0.97, :223-227), Adam over netTask only (:143).  Image_f's own gradients are never
    used by the reference (no optimizer owns them), so they are not computed here."""

    def __init__(self, state_dicts, lr=2e-4, w_syn_l1=1.0, w_real_l1=1.0, scale_G=1.0):
        self.sd = {k: collections.OrderedDict((n, t.detach().clone().float()) for n, t in v.items())
                   for k, v in state_dicts.items()}
        self.lr, self.w = lr, (w_syn_l1, w_real_l1, scale_G)
        self.n_step = 0
        self.adam = {}
        for n, p in self.sd["Task"].items():
            p.requires_grad_(True)
            self.adam[n] = (torch.zeros_like(p), torch.zeros_like(p))

    def step(self, batch, update=True):
        for p in self.sd["Task"].values():
            p.grad = None
        t = {}
        t["syn_image"], t["real_image"] = batch["A_i"].float(), batch["B_i"].float()
        t["syn_depth"], t["real_depth"] = batch["A_d"].float(), batch["B_d"].float()
        with torch.no_grad():
            f_syn = ref_nets.resnet_generator(self.sd["Image_f"], t["syn_image"])
            f_real = ref_nets.resnet_generator(self.sd["Image_f"], t["real_image"])
        t["pred_syn_depth"] = ref_nets.unet_generator(self.sd["Task"], f_syn)
        t["pred_real_depth"] = ref_nets.unet_generator(self.sd["Task"], f_real)
        L = {}
        L["syn_norms"] = ref_ops.l1_mean(ref_ops.surface_normals_old(t["syn_depth"]),
                                         ref_ops.surface_normals_old(t["pred_syn_depth"].detach()))        # :217 (not in loss_G)
        m_s = torch.where(t["syn_depth"] < -0.97, torch.tensor(0.0), torch.tensor(1.0))                  # :223
        L["task_syn"] = ref_ops.l1_mean(t["syn_depth"] * m_s, t["pred_syn_depth"] * m_s)
        m_r = torch.where(t["real_depth"] < -0.97, torch.tensor(0.0), torch.tensor(1.0))                 # :226
        L["task_real"] = ref_ops.l1_mean(t["real_depth"] * m_r, t["pred_real_depth"] * m_r)
        G = (L["task_syn"] * self.w[0] + L["task_real"] * self.w[1]) * self.w[2]                         # :230-234
        G.backward()
        grads = {n: p.grad.detach().clone() for n, p in self.sd["Task"].items()}
        if update:
            self.n_step += 1
            with torch.no_grad():
                for n, p in self.sd["Task"].items():
                    m, v = self.adam[n]
                    ref_ops.adam_update(p, p.grad, m, v, self.n_step, self.lr)
        losses = {k: float(v) for k, v in L.items()}
        losses["G"] = float(G)
        return dict(tensors=t, losses=losses, grads=grads)


def synthetic_sr_batch(B, h, w, seed=1, depth_kind="smooth"):
    """HR (2h x 2w) synthetic batch with the K / crop conventions of data/my_naive_sr_dataset.py:190-207:
    K_A scaled by [[2,1,2],[1,2,2],[1,1,1]], crop_A = HR extent, crop_B = LR extent."""
    b = synthetic_batch(B, 2 * h, 2 * w, seed=seed, depth_kind=depth_kind)
    scale = torch.tensor([[2.0, 1, 2], [1, 2, 2], [1, 1, 1]], dtype=torch.float64)
    b["K_A"] = b["K_A"] * scale
    b["crop_B"] = torch.tensor([[0, h, 0, w]] * B)
    return b


def synthetic_batch(B, H, W, seed=1, depth_kind="noise"):
    """Synthetic RGB-D batch of SURVEY.md section 8(d) / Appendix D (CPU tensors, dataset dict keys of
    data/my_main_dataset.py:195)."""
    g = torch.Generator().manual_seed(seed)

    def depth():
        if depth_kind == "noise":
            d = torch.rand(B, 1, H, W, generator=g) * 1.6 - 0.8
            d[torch.rand(B, 1, H, W, generator=g) < 0.05] = -1.0
            return d
        yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
        c = torch.rand(B, 5, generator=g) - 0.5
        d = (c[:, 0, None, None] * xx + c[:, 1, None, None] * yy + 0.5 * c[:, 2, None, None]
             + 0.2 * torch.sin(3.0 * xx * (1 + c[:, 3, None, None]) + 2.0 * yy * (1 + c[:, 4, None, None])))
        d = d.clamp(-0.9, 0.9)[:, None].contiguous()
        for b in range(B):
            for _ in range(int(torch.randint(3, 9, (1,), generator=g))):
                y0 = int(torch.randint(0, H - 8, (1,), generator=g)); x0 = int(torch.randint(0, W - 8, (1,), generator=g))
                hh = int(torch.randint(2, max(3, H // 8), (1,), generator=g)); ww = int(torch.randint(2, max(3, W // 8), (1,), generator=g))
                d[b, 0, y0:y0 + hh, x0:x0 + ww] = -1.0
        return d

    A_i = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    B_i = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    A_d, B_d = depth(), depth()
    K = torch.tensor([[577.87, 0, 319.5], [0, 577.87, 239.5], [0, 0, 1]], dtype=torch.float64).repeat(B, 1, 1)
    crop = torch.tensor([[0, H, 0, W]] * B)
    return dict(A_i=A_i, B_i=B_i, A_d=A_d, B_d=B_d, A_paths=["a"] * B, B_paths=["b"] * B,
                K_A=K, K_B=K.clone(), crop_A=crop, crop_B=crop.clone())
