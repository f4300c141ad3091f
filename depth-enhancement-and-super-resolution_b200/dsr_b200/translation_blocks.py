"""``GanBlockStep`` - the two networks of BASELINE configs[4] (``translation_block``: GroupNorm ResNet generator + n_layers
PatchGAN discriminator) trained for one step with the LSGAN terms of the reference's ``models/translation_model.py``:

  generator step      loss_G = 0.5 * MSE(D(G(depth, img)), 1)                       (translation_model.py:214)
  discriminator step  loss_D = 0.5 * (MSE(D(real), 1) + MSE(D(fake.detach()), 0))   (translation_model.py:199-205)
  Adam(lr, betas=(beta1, 0.999)) on each net                                         (translation_model.py:117-118)

This is NOT the full ``TranslationModel`` (two generators, four discriminators, cycle / identity / normal / range losses, the
3 : 1 schedule - SURVEY.md section 8f rank 3, not built yet); it is the measured, parity-tested fwd + bwd of the two network
families that model is made of, behind the same ``set_input`` / ``optimize_parameters`` surface.
"""
from types import SimpleNamespace

import torch

from . import ops, translation_network
from .base_model import BaseModel, GraphStepMixin
from .main_model import ArenaAdam, ParamArena


class GanBlockStep(GraphStepMixin, BaseModel):
    def __init__(self, opt):
        BaseModel.__init__(self, opt)
        self.loss_names = ["G_A", "D_A_depth"]
        self.model_names = ["G_A", "D_A_depth"]
        self.visual_names = ["real_img_A", "real_depth_A", "real_depth_B", "fake_depth_B"]
        g_opt = SimpleNamespace(ngf_img=32, ngf_depth=32, ngf=64, norm="group", dropout=False, init_type=opt.init_type,
                                gpu_ids=opt.gpu_ids, input_nc_img=3, n_downsampling=2, use_semantic=False, n_blocks=9,
                                upsampling_type="transpose", output_nc_depth=1, input_nc_depth=1)   # translation_model.py:84
        self.netG_A = translation_network.define_Gen(g_opt, input_type="img_depth")
        self.netD_A_depth = translation_network.define_D(opt, input_type="depth")
        self._graph_init(opt)
        self.grad_sync = None
        self._in = None
        if self.isTrain and self.gpu_ids:
            beta1 = getattr(opt, "beta1", 0.5)
            self.arena_G = ParamArena([self._unwrap(self.netG_A)], self.device)
            self.arena_D = ParamArena([self._unwrap(self.netD_A_depth)], self.device)
            self.optimizer_G, self.optimizer_D = ArenaAdam(self.arena_G, opt.lr), ArenaAdam(self.arena_D, opt.lr)
            for o in (self.optimizer_G, self.optimizer_D):
                o.param_groups[0]["betas"] = (beta1, 0.999)
            self.optimizers += [self.optimizer_G, self.optimizer_D]

    def set_input(self, input):                                     # translation_model.py:129-137 (A domain + a real depth)
        src = dict(real_img_A=input["A_i"], real_depth_A=input["A_d"], real_depth_B=input["B_d"])
        shapes = {k: tuple(v.shape) for k, v in src.items()}
        if self._in is None or self._in["shapes"] != shapes:
            self._in = dict(shapes=shapes)
            for k, v in src.items():
                self._in[k] = torch.empty(v.shape, device=self.device, dtype=torch.float32)
        for k, v in src.items():
            v = v if v.dtype == torch.float32 else v.float()
            if v.device.type == "cpu" and self.device.type == "cuda" and not v.is_pinned():
                v = v.pin_memory()
            self._in[k].copy_(v, non_blocking=True)
            setattr(self, k, self._in[k])

    @staticmethod
    def _lsgan(pred, value):              # GANLoss('lsgan'): MSELoss against a constant map (translation_network.py:161-188)
        ones = torch.ones((pred.shape[0], 1) + tuple(pred.shape[2:]), device=pred.device)
        return ops.masked_l1_l2(torch.full_like(pred, value), pred, ones)[1]

    def forward(self):
        self.fake_depth_B = self.netG_A(self.real_depth_A, self.real_img_A)

    def _step_body(self):
        if self.device.type == "cuda":
            ops.zero_pool_reset(self.device)
        self.set_requires_grad([self.netD_A_depth], False)
        self.forward()
        self.optimizer_G.zero_grad()
        self.loss_G_A = 0.5 * self._lsgan(self.netD_A_depth(self.fake_depth_B), 1.0)
        self.loss_G = self.loss_G_A
        self.loss_G.backward()
        ops.join_side()
        self.optimizer_G.step()
        self.set_requires_grad([self.netD_A_depth], True)
        self.optimizer_D.zero_grad()
        fake = self.fake_depth_B.detach()
        self.loss_D_A_depth = 0.5 * (self._lsgan(self.netD_A_depth(self.real_depth_B), 1.0) + self._lsgan(self.netD_A_depth(fake), 0.0))
        self.loss_D_A_depth.backward()
        ops.join_side()
        self.optimizer_D.step()

    def optimize_parameters(self, iters=0, fr=1):
        self._graph_optimize()
