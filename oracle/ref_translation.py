"""Oracle (torch CPU fp32) restatement of the reference's translation block: models/translation_network.py (SurfaceNormals
:329-360, losses :281-327, GANLoss :139-205) and one ``TranslationModel.optimize_parameters`` call
(models/translation_model.py:140-291): the default flags (cycle B, depth + normal discriminators, identity B, depth range
losses) and, optionally, ``use_cycle_A`` / ``l_mean_A`` / ``l_mean_B`` / ``l_tv_A`` (``inp_B='depth'`` is not restated).
TEST INFRASTRUCTURE ONLY.  Pinned against the live reference by tests/golden/translation_step_*.npz (default flags) and
tests/golden/translation_flags_b1_64.npz (the optional terms)."""
import math

import torch
import torch.nn.functional as F

from . import ref_nets


def fov_normals(x):
    """SurfaceNormals.get_normal (:341-359)."""
    b, c, h, w = x.shape
    H0, W0, fov = 482, 642, 60
    gx = (torch.arange(1, W0 + 1) - (W0 + 1) / 2) / (W0 / 2) * math.tan(fov / 2 / 180 * math.pi)
    gy = -(torch.arange(1, H0 + 1) - (H0 + 1) / 2) / (H0 / 2) * math.tan(fov / 2 / 180 * math.pi) * (H0 / W0)
    grid = torch.stack([gx.repeat(H0, 1), gy.repeat(W0, 1).t(), torch.ones(H0, W0)], 0).float()
    ph, pw = (H0 - h) // 2, (W0 - w) // 2
    grid = grid[:, ph + 1:ph + 1 + h, pw + 1:pw + 1 + w]
    pv = F.pad(x.repeat(1, 3, 1, 1) * grid, (1, 1, 1, 1), mode="reflect")
    dgx = pv[:, :, 0:h, 0:w] / 2 - pv[:, :, 0:h, 2:w + 2] / 2
    dgy = pv[:, :, 2:h + 2, 0:w] / 2 - pv[:, :, 0:h, 0:w] / 2
    crs = torch.cross(dgx, dgy, dim=1)
    norm = crs.norm(2, 1, keepdim=True)
    return -crs / norm.clamp(min=1e-8)


def masked_l1(x, y, mask):              # MaskedL1Loss (:281-286)
    return (torch.abs(y - x) * mask).sum() / (mask.sum() + 1e-6)


def cos_sim_loss(x, y):                 # CosSimLoss (:310-316)
    return torch.mean(1 - F.cosine_similarity(x, y, dim=1))


def masked_mean_dif(x, y, mask):        # MaskedMeanDif (:288-293): per-sample masked mean difference, then mean of |.|
    return torch.mean(torch.abs(((y - x) * mask).sum(dim=(2, 3)) / (mask.sum(dim=(2, 3)) + 1e-6)))


def masked_cos_sim_loss(x, y, mask):    # MaskedCosSimLoss (:320-327) - the reference divides by sum(mask) + 1e+6 (sic)
    loss = 1 - F.cosine_similarity(x, y, dim=1)
    return (loss.unsqueeze(1) * mask).sum() / (mask.sum() + 1e+6)


def tv_norm(x):                         # TV_norm(surf_normal=True) (:302-311): first two components only
    x = x[:, :2]
    return ((x[:, :, 1:, :] - x[:, :, :-1, :]).pow(2).sum() + (x[:, :, :, 1:] - x[:, :, :, :-1]).pow(2).sum()) / x.numel()


def lsgan(pred, real):                  # GANLoss('lsgan') (:161-205)
    return F.mse_loss(pred, torch.ones_like(pred) if real else torch.zeros_like(pred))


class OracleTranslationStep:
    """state dicts of G_A, G_B (img_depth generators) and the four discriminators; Adam(lr, betas=(beta1, 0.999)) with
    weight_decay w_decay_G on the generators (translation_model.py:117-118)."""

    def __init__(self, sds, lr=2e-4, beta1=0.5, w_decay_G=1e-4, num_iter_gen=3, l_cycle_B=5.0, l_normal=1.0, l_identity=1.0,
                 l_depth_A=5.0, l_depth_B=5.0, use_cycle_A=False, l_cycle_A=10.0, l_mean_A=0.0, l_mean_B=0.0, l_tv_A=0.0,
                 inp_B="img_depth"):
        self.sd = {k: {n: t.detach().clone().float().requires_grad_(True) for n, t in v.items()} for k, v in sds.items()}
        self.cfg = dict(lr=lr, beta1=beta1, wd=w_decay_G, n_gen=num_iter_gen, l_cycle_B=l_cycle_B, l_normal=l_normal,
                        l_identity=l_identity, l_depth_A=l_depth_A, l_depth_B=l_depth_B, use_cycle_A=use_cycle_A,
                        l_cycle_A=l_cycle_A, l_mean_A=l_mean_A, l_mean_B=l_mean_B, l_tv_A=l_tv_A, inp_B=inp_B)
        self.adam = {k: {n: (torch.zeros_like(p), torch.zeros_like(p)) for n, p in v.items()} for k, v in self.sd.items()}
        self.steps = {"G": 0, "D": 0}

    def _G(self, name, depth, img):
        if name == "G_B" and self.cfg["inp_B"] == "depth":                  # (:146-147, :167-168, :185-186)
            img = None
        return ref_nets.translation_generator(self.sd[name], depth, img)

    def _D(self, name, x):
        return ref_nets.nlayer_discriminator(self.sd[name], x)

    def forward(self, b):               # translation_model.py:140-187
        t = {}
        A_d, A_i, B_d, B_i = b["A_depth"].float(), b["A_img"].float(), b["B_depth"].float(), b["B_img"].float()
        t["hole_mask_A"] = A_d <= -0.98
        t["fake_depth_B"] = self._G("G_A", A_d, A_i)
        t["fake_depth_A"] = self._G("G_B", B_d, B_i)
        for k in ("real_norm_A", "real_norm_B", "fake_norm_A", "fake_norm_B"):
            src = dict(real_norm_A=A_d, real_norm_B=B_d, fake_norm_A=t["fake_depth_A"], fake_norm_B=t["fake_depth_B"])[k]
            t[k] = fov_normals(src)
        t["hole_mask_B"] = t["fake_depth_A"] <= -0.98
        if self.cfg["use_cycle_A"]:                                          # (:165-172) A -> B -> A through G_B
            t["rec_depth_A"] = self._G("G_B", t["fake_depth_B"], A_i)
            t["rec_norm_A"] = fov_normals(t["rec_depth_A"])
        t["rec_depth_B"] = self._G("G_A", t["fake_depth_A"], B_i)          # (:176-178; the first, detached call is discarded)
        t["rec_norm_B"] = fov_normals(t["rec_depth_B"])
        t["idt_A"] = self._G("G_A", B_d, B_i)                               # (:181-187)
        t["idt_B"] = self._G("G_B", A_d, A_i)
        t.update(A_d=A_d, B_d=B_d)
        return t

    def _adam(self, nets, key, wd):
        c = self.cfg
        self.steps[key] += 1
        s = self.steps[key]
        with torch.no_grad():
            for net in nets:
                for n, p in self.sd[net].items():
                    g = p.grad if p.grad is not None else torch.zeros_like(p)
                    if wd:
                        g = g + wd * p
                    m, v = self.adam[net][n]
                    m.mul_(c["beta1"]).add_(g, alpha=1 - c["beta1"])
                    v.mul_(0.999).addcmul_(g, g, value=0.001)
                    denom = (v.sqrt() / math.sqrt(1 - 0.999 ** s)).add_(1e-8)
                    p.addcdiv_(m, denom, value=-c["lr"] / (1 - c["beta1"] ** s))

    def _zero(self, nets):
        for net in nets:
            for p in self.sd[net].values():
                p.grad = None

    def step(self, b):                  # translation_model.py:274-291
        c = self.cfg
        L, first = {}, None
        for it in range(c["n_gen"]):
            t = self.forward(b)
            self._zero(["G_A", "G_B"])
            # backward_G (:208-268): the discriminators are frozen (requires_grad False): only G gets gradients
            G_A = 0.5 * lsgan(self._D("D_A_depth", t["fake_depth_B"]), True) + 0.5 * lsgan(self._D("D_A_normal", t["fake_norm_B"]), True)
            G_B = 0.5 * lsgan(self._D("D_B_depth", t["fake_depth_A"]), True) + 0.5 * lsgan(self._D("D_B_normal", t["fake_norm_A"]), True)
            cyc_B = F.l1_loss(t["rec_depth_B"], t["B_d"]) * c["l_cycle_B"]
            cyc_n_B = cos_sim_loss(t["rec_norm_B"], t["real_norm_B"]) * c["l_normal"] * c["l_cycle_B"]
            idt_B = F.l1_loss(t["idt_B"], t["A_d"]) * c["l_identity"]
            rng_A = masked_l1(t["fake_depth_B"], t["A_d"], ~t["hole_mask_A"]) * c["l_depth_A"]
            rng_B = masked_l1(t["fake_depth_A"], t["B_d"], ~t["hole_mask_B"]) * c["l_depth_B"]
            loss_G = (G_A + rng_A) + (G_B + cyc_B + cyc_n_B + idt_B + rng_B)
            extra = {}
            if c["use_cycle_A"]:                                             # (:222-225)
                mA = ~t["hole_mask_A"]
                extra["cycle_A"] = masked_l1(t["rec_depth_A"], t["A_d"], mA) * c["l_cycle_A"]
                extra["cycle_n_A"] = masked_cos_sim_loss(t["rec_norm_A"], t["real_norm_A"], mA.repeat(1, 3, 1, 1)) * c["l_normal"] * c["l_cycle_A"]
            if c["l_mean_A"] > 0:                                            # (:240-245)
                extra["mean_dif_A"] = masked_mean_dif(t["fake_depth_B"], t["A_d"], ~t["hole_mask_A"]) * c["l_mean_A"]
            if c["l_mean_B"] > 0:
                extra["mean_dif_B"] = masked_mean_dif(t["fake_depth_A"], t["B_d"], ~t["hole_mask_B"]) * c["l_mean_B"]
            if c["l_tv_A"] > 0:                                              # (:247-249)
                extra["tv_norm_A"] = tv_norm(t["fake_norm_B"]) * c["l_tv_A"]
            for v in extra.values():
                loss_G = loss_G + v
            gs = torch.autograd.grad(loss_G, [p for n in ("G_A", "G_B") for p in self.sd[n].values()], allow_unused=True)
            k = 0
            for n in ("G_A", "G_B"):
                for p in self.sd[n].values():
                    p.grad = gs[k]; k += 1
            L = dict(G_A=float(G_A), G_B=float(G_B), cycle_B=float(cyc_B), cycle_n_B=float(cyc_n_B), idt_B=float(idt_B),
                     depth_range_A=float(rng_A), depth_range_B=float(rng_B), G=float(loss_G))
            L.update({k2: float(v) for k2, v in extra.items()})
            if first is None:
                first = dict(losses=dict(L), tensors={k2: v.detach().clone() for k2, v in t.items()},
                             grads={(n, pn): p.grad.detach().clone() for n in ("G_A", "G_B") for pn, p in self.sd[n].items() if p.grad is not None})
            self._adam(["G_A", "G_B"], "G", c["wd"])
        discs = ["D_A_depth", "D_A_normal", "D_B_depth", "D_B_normal"]
        self._zero(discs)
        pairs = dict(D_A_depth=(t["rec_depth_B"], t["fake_depth_B"]), D_A_normal=(t["rec_norm_B"], t["fake_norm_B"]),
                     D_B_depth=(t["A_d"], t["fake_depth_A"]), D_B_normal=(t["real_norm_A"], t["fake_norm_A"]))   # :189-206
        for n in discs:
            real, fake = pairs[n]
            loss_D = 0.5 * (lsgan(self._D(n, real.detach()), True) + lsgan(self._D(n, fake.detach()), False))
            gs = torch.autograd.grad(loss_D, list(self.sd[n].values()))
            for p, g in zip(self.sd[n].values(), gs):
                p.grad = g
            L[n] = float(loss_D)
        d_grads = {(n, pn): p.grad.detach().clone() for n in discs for pn, p in self.sd[n].items()}
        self._adam(discs, "D", 0.0)
        return dict(first=first, losses=L, tensors={k2: v.detach() for k2, v in t.items()}, d_grads=d_grads)
