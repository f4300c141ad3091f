"""``TranslationModel`` - the unpaired depth-translation step (``--model translation_block --model_type translation``), drop-in
for the reference's ``models/translation_model.py`` (citations into /root/reference).  SURVEY.md section 8f rank 3 /
BASELINE configs[4].

Two GroupNorm ResNet generators (G_A: A -> B, G_B: B -> A, ``input_type='img_depth'``), four PatchGAN discriminators (depth
and field-of-view normals for each domain), LSGAN, cycle B (L1 + cosine similarity of normals), identity B, masked depth-range
losses; ``optimize_parameters`` = ``num_iter_gen`` generator iterations with the discriminators frozen, then one discriminator
update (:274-291).  Same method surface, ``loss_*`` / visual names and checkpoint names as the reference.

Built around the default flag set (translation_model.py:14-43): ``use_cycle_B``, ``disc_for_depth``, ``disc_for_normals``,
``l_identity > 0`` with ``no_idt_A`` and the depth-range losses are always on; the optional terms ``use_cycle_A`` (masked L1 +
masked cosine similarity through G_B, :165-172, :222-225), ``l_mean_A`` / ``l_mean_B`` (:243-247), ``l_tv_A`` (:247-249) and the
depth-only ``inp_B='depth'`` generator (:146-147) are wired as well.  Deliberate difference: the reference runs ``netG_A`` on ``(fake_depth_A,
real_img_B)`` twice per forward and throws the first result away (:177-178); here it runs once.
"""
import torch

from . import ops, translation_network
from .base_model import BaseModel, GraphStepMixin
from .main_model import ArenaAdam, ParamArena
from .translation_blocks import GanBlockStep

from types import SimpleNamespace


def data_to_meters(x, max_distance):                                # util/util.py:8-12
    scale = max_distance / 2.0
    return (x * scale + scale) / 1000.0


class TranslationModel(GraphStepMixin, BaseModel):
    @staticmethod
    def modify_commandline_options(parser, is_train):               # translation_model.py:13-43
        for name, default in (("l_cycle_A_begin", 10.0), ("l_cycle_A_end", 10.0), ("l_cycle_B_begin", 5.0), ("l_cycle_B_end", 5.0),
                              ("l_identity", 1.0), ("l_normal", 1.0), ("l_reconstruction_semantic", 0.0), ("l_depth_A_begin", 5.0),
                              ("l_depth_A_end", 0.0), ("l_depth_B_begin", 5.0), ("l_depth_B_end", 0.0), ("l_mean_A", 0.0),
                              ("l_mean_B", 0.0), ("l_tv_A", 0.0), ("w_decay_G", 0.0001)):
            parser.add_argument("--" + name, type=float, default=default)
        for name, default in (("l_max_iter", 5000), ("l_num_iter", 5000), ("num_iter_gen", 3), ("num_iter_dis", 1)):
            parser.add_argument("--" + name, type=int, default=default)
        parser.add_argument("--no_idt_A", action="store_true", default=True)
        parser.add_argument("--use_cycle_A", action="store_true", default=False)
        parser.add_argument("--use_cycle_B", action="store_true", default=True)
        parser.add_argument("--disc_for_normals", action="store_true", default=True)
        parser.add_argument("--disc_for_depth", action="store_true", default=True)
        parser.add_argument("--inp_B", type=str, default="img_depth")
        parser.add_argument("--norm_d", type=str, default="none")
        return parser

    def __init__(self, opt):                                        # translation_model.py:45-127
        BaseModel.__init__(self, opt)
        if opt.inp_B not in ("img_depth", "depth") or \
                not (opt.use_cycle_B and opt.disc_for_depth and opt.disc_for_normals and opt.l_identity > 0 and opt.no_idt_A):
            raise NotImplementedError("dsr_b200.TranslationModel: cycle B, the depth + normal discriminators and identity B "
                                      "(the defaults of translation_model.py:14-43, README.md:51) cannot be switched off")
        if self.isTrain:                                             # translation_model.py:47-68, same order
            self.loss_names = ["G_A", "G_B", "depth_dif_A", "depth_dif_B"]
            if opt.l_mean_A > 0:
                self.loss_names.append("mean_dif_A")
            if opt.l_mean_B > 0:
                self.loss_names.append("mean_dif_B")
            if opt.use_cycle_A:
                self.loss_names += ["cycle_A", "cycle_n_A"]
            self.loss_names += ["cycle_B", "cycle_n_B", "D_A_depth", "D_B_depth", "D_A_normal", "D_B_normal", "idt_A", "idt_B"]
            if opt.l_depth_A_begin > 0:
                self.loss_names.append("depth_range_A")
            if opt.l_depth_B_begin > 0:
                self.loss_names.append("depth_range_B")
            if opt.l_tv_A > 0:
                self.loss_names.append("tv_norm_A")                  # translation_model.py:67-68
        self.loss_names_test = ["depth_dif_A", "depth_dif_B"]
        self.visual_names = ["real_img_A", "real_depth_A", "real_img_B", "real_depth_B", "fake_depth_B", "fake_depth_A", "name_A",
                             "name_B"] + (["rec_depth_A"] if opt.use_cycle_A else []) + ["rec_depth_B"]
        if self.isTrain:
            self.visual_names += ["idt_A", "idt_B"]
        self.model_names = ["G_A", "G_B"]
        g_opt = lambda: SimpleNamespace(ngf_img=32, ngf_depth=32, ngf=64, norm="group", dropout=False, init_type=opt.init_type,
                                        gpu_ids=opt.gpu_ids, input_nc_img=3, n_downsampling=2, use_semantic=False, n_blocks=9,
                                        upsampling_type="transpose", output_nc_depth=1, input_nc_depth=1)       # :84-88
        self.netG_A = translation_network.define_Gen(g_opt(), input_type="img_depth")
        self.netG_B = translation_network.define_Gen(g_opt(), input_type=opt.inp_B)
        self.disc = []
        if self.isTrain:
            self.model_names += ["D_A_depth", "D_B_depth", "D_A_normal", "D_B_normal"]
            self.netD_A_depth = translation_network.define_D(opt, input_type="depth")                         # :95-104
            self.netD_B_depth = translation_network.define_D(opt, input_type="depth")
            self.netD_A_normal = translation_network.define_D(opt, input_type="normal")
            self.netD_B_normal = translation_network.define_D(opt, input_type="normal")
            self.disc = [self.netD_A_depth, self.netD_B_depth, self.netD_A_normal, self.netD_B_normal]
            self.l_depth_A, self.l_depth_B = opt.l_depth_A_begin, opt.l_depth_B_begin
            self.l_cycle_A, self.l_cycle_B = opt.l_cycle_A_begin, opt.l_cycle_B_begin
            self.calc_l_step()
            if self.gpu_ids:
                self.arena_G = ParamArena([self._unwrap(self.netG_A), self._unwrap(self.netG_B)], self.device)
                self.arena_D = ParamArena([self._unwrap(m) for m in self.disc], self.device)
                self.optimizer_G, self.optimizer_D = ArenaAdam(self.arena_G, opt.lr), ArenaAdam(self.arena_D, opt.lr)
                for o in (self.optimizer_G, self.optimizer_D):
                    o.param_groups[0]["betas"] = (opt.beta1, 0.999)                                          # :117-118
            else:
                import itertools
                self.optimizer_G = torch.optim.Adam(itertools.chain(self.netG_A.parameters(), self.netG_B.parameters()), lr=opt.lr,
                                                    betas=(opt.beta1, 0.999), weight_decay=opt.w_decay_G)
                self.optimizer_D = torch.optim.Adam(itertools.chain(*[m.parameters() for m in self.disc]), lr=opt.lr,
                                                    betas=(opt.beta1, 0.999))
            self.optimizers += [self.optimizer_G, self.optimizer_D]
            self.opt_names = ["optimizer_G", "optimizer_D"]
        self._in = None
        self.loss_idt_A = 0
        self._graph_init(opt)        # optional CUDA-graph replay of the whole optimize_parameters call (opt.cuda_graph)
        # the scheduled loss weights (update_loss_weight, :300-305) live in a small device tensor: a captured step multiplies
        # by the tensor, so a replay sees the current schedule instead of the capture-time floats
        self._lw_dev = None
        if self.isTrain and self.gpu_ids:
            self._lw_dev = torch.zeros(4, device=self.device, dtype=torch.float32)
            self._lw_host = torch.zeros(4, dtype=torch.float32).pin_memory()
            self._sync_loss_weights()

    _LW = ("l_depth_A", "l_depth_B", "l_cycle_A", "l_cycle_B")

    def _sync_loss_weights(self):
        if self._lw_dev is not None:
            self._lw_host.copy_(torch.tensor([getattr(self, k) for k in self._LW], dtype=torch.float32))
            self._lw_dev.copy_(self._lw_host, non_blocking=True)

    def _w(self, name):
        """scheduled loss weight as a multiplier: a device scalar on the GPU path (graph-replay safe), the float otherwise"""
        return getattr(self, name) if self._lw_dev is None else self._lw_dev[self._LW.index(name)]

    def set_input(self, input):                                     # translation_model.py:129-137
        self.name_A, self.name_B = input["A_name"], input["B_name"]
        src = dict(real_img_A=input["A_img"], real_depth_A=input["A_depth"], real_img_B=input["B_img"], real_depth_B=input["B_depth"])
        shapes = {k: tuple(v.shape) for k, v in src.items()}
        if self._in is None or self._in["shapes"] != shapes:
            if getattr(self, "_graph", None) is not None:
                raise RuntimeError("dsr_b200: the captured CUDA graph is bound to the first batch shape; call reset_graph()")
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
    def _valid_mask(x):
        """~hole_mask of get_mask (:325-327, hole = depth <= -0.98) as a float {0,1} map.  ops.below_mask gives 0 where x < thr, so
        the threshold is the next float above -0.98 (x <= -0.98  <=>  x < nextafter(-0.98, +inf))."""
        import numpy as np
        thr = float(np.nextafter(np.float32(-0.98), np.float32(1.0)))
        return ops.below_mask(x, thr)

    _lsgan = staticmethod(GanBlockStep._lsgan)

    def forward(self):                                              # translation_model.py:140-187
        self.valid_A = self._valid_mask(self.real_depth_A)          # ~hole_mask_A
        self.fake_depth_B = self.netG_A(self.real_depth_A, self.real_img_A)
        self.fake_depth_A = self._G_B(self.real_depth_B, self.real_img_B)
        if self.isTrain:
            self.real_norm_A = ops.fov_normals(self.real_depth_A)
            self.real_norm_B = ops.fov_normals(self.real_depth_B)
            self.fake_norm_A = ops.fov_normals(self.fake_depth_A)
            self.fake_norm_B = ops.fov_normals(self.fake_depth_B)
        self.valid_B = self._valid_mask(self.fake_depth_A)          # ~hole_mask_B (no gradient through the mask)
        if self.opt.use_cycle_A:                                    # :165-172: A -> B -> A through G_B
            self.rec_depth_A = self._G_B(self.fake_depth_B, self.real_img_A)
            if self.isTrain:
                self.rec_norm_A = ops.fov_normals(self.rec_depth_A)
        self.rec_depth_B = self.netG_A(self.fake_depth_A, self.real_img_B)          # :176-178 (once, see the module docstring)
        if self.isTrain:
            self.rec_norm_B = ops.fov_normals(self.rec_depth_B)
            self.idt_A = self.netG_A(self.real_depth_B, self.real_img_B)            # :181-187
            self.idt_B = self._G_B(self.real_depth_A, self.real_img_A)

    def _G_B(self, depth, img):
        """netG_B is a depth-only generator with --inp_B depth (translation_model.py:146-147, :167-168, :185-186)"""
        return self.netG_B(depth) if self.opt.inp_B == "depth" else self.netG_B(depth, img)

    def _masked_l1(self, x, y, valid):
        """MaskedL1Loss (translation_network.py:281-286): sum(|y - x| * mask) / (sum(mask) + 1e-6); gradient to x."""
        n = float(valid.numel())
        s_mask = ops.masked_sums(valid, valid, valid)[0]            # sum(mask) (mask in {0,1})
        scale = (n / (s_mask + 1e-6)).to(torch.float32)
        return ops.masked_l1_l2(y, x, valid)[0] * scale

    def _l1(self, x, y):
        return ops.masked_l1_l2(y, x, torch.ones((x.shape[0], 1) + tuple(x.shape[2:]), device=x.device))[0]

    def _d_loss(self, netD, real, fake):                            # backward_D_base :189-194
        loss = 0.5 * (self._lsgan(netD(real.detach()), 1.0) + self._lsgan(netD(fake.detach()), 0.0))
        loss.backward()
        ops.join_side()
        return loss

    def backward_D_A(self):                                         # :196-200
        self.loss_D_A_depth = self._d_loss(self.netD_A_depth, self.rec_depth_B, self.fake_depth_B)
        self.loss_D_A_normal = self._d_loss(self.netD_A_normal, self.rec_norm_B, self.fake_norm_B)

    def backward_D_B(self):                                         # :202-206
        self.loss_D_B_depth = self._d_loss(self.netD_B_depth, self.real_depth_A, self.fake_depth_A)
        self.loss_D_B_normal = self._d_loss(self.netD_B_normal, self.real_norm_A, self.fake_norm_A)

    def backward_G(self):                                           # translation_model.py:208-272
        opt = self.opt
        self.loss_G_A = 0.5 * self._lsgan(self.netD_A_depth(self.fake_depth_B), 1.0) + \
            0.5 * self._lsgan(self.netD_A_normal(self.fake_norm_B), 1.0)
        self.loss_G_B = 0.5 * self._lsgan(self.netD_B_depth(self.fake_depth_A), 1.0) + \
            0.5 * self._lsgan(self.netD_B_normal(self.fake_norm_A), 1.0)
        loss_A, loss_B = self.loss_G_A, self.loss_G_B
        if opt.use_cycle_A:                                          # :222-225
            self.loss_cycle_A = self._masked_l1(self.rec_depth_A, self.real_depth_A, self.valid_A) * self._w("l_cycle_A")
            # MaskedCosSimLoss with the 3-channel mask: 3 * sum(mask * (1 - cos)) / (3 * sum(mask) + 1e+6)  (the 1e+6 is the
            # reference's, translation_network.py:327)
            s_mask = ops.masked_sums(self.valid_A, self.valid_A, self.valid_A)[0]
            scale = (3.0 / (3.0 * s_mask + 1e+6)).to(torch.float32)
            self.loss_cycle_n_A = ops.cos_sim_masked_sum(self.rec_norm_A, self.real_norm_A, self.valid_A) * scale * \
                opt.l_normal * self._w("l_cycle_A")
            loss_A = loss_A + self.loss_cycle_A + self.loss_cycle_n_A
        self.loss_cycle_B = self._l1(self.rec_depth_B, self.real_depth_B) * self._w("l_cycle_B")
        self.loss_cycle_n_B = ops.cos_sim_loss(self.rec_norm_B, self.real_norm_B) * opt.l_normal * self._w("l_cycle_B")
        loss_B = loss_B + self.loss_cycle_B + self.loss_cycle_n_B
        self.loss_idt_A = 0
        self.loss_idt_B = self._l1(self.idt_B, self.real_depth_A) * opt.l_identity
        loss_B = loss_B + self.loss_idt_B
        if opt.l_mean_A > 0:                                         # :240-245
            self.loss_mean_dif_A = ops.masked_mean_dif(self.fake_depth_B, self.real_depth_A, self.valid_A) * opt.l_mean_A
            loss_A = loss_A + self.loss_mean_dif_A
        if opt.l_mean_B > 0:
            self.loss_mean_dif_B = ops.masked_mean_dif(self.fake_depth_A, self.real_depth_B, self.valid_B) * opt.l_mean_B
            loss_B = loss_B + self.loss_mean_dif_B
        if self.l_depth_A > 0:
            self.loss_depth_range_A = self._masked_l1(self.fake_depth_B, self.real_depth_A, self.valid_A) * self._w("l_depth_A")
            loss_A = loss_A + self.loss_depth_range_A
        if self.l_depth_B > 0:
            self.loss_depth_range_B = self._masked_l1(self.fake_depth_A, self.real_depth_B, self.valid_B) * self._w("l_depth_B")
            loss_B = loss_B + self.loss_depth_range_B
        if opt.l_tv_A > 0:                                           # :247-249, TV_norm(surf_normal=True) translation_network.py:302-311:
            n2 = self.fake_norm_B[:, :2]                             # squared forward differences of the first two components / numel
            self.loss_tv_norm_A = ops.tv_loss(n2) * (opt.l_tv_A / n2.numel())
            loss_A = loss_A + self.loss_tv_norm_A
        self.loss_G = loss_A + loss_B
        self.loss_G.backward()
        ops.join_side()
        with torch.no_grad():                                       # :266-270 (metres)
            md = opt.max_distance
            self.loss_depth_dif_A = self._masked_l1(data_to_meters(self.fake_depth_B.detach(), md),
                                                    data_to_meters(self.real_depth_A, md), self.valid_A)
            self.loss_depth_dif_B = self._masked_l1(data_to_meters(self.fake_depth_A.detach(), md),
                                                    data_to_meters(self.real_depth_B, md), self.valid_B)

    def zero_grad(self, nets=None):                                 # nn.Module.zero_grad of the reference's double base class
        pass

    def _adam_G(self):
        if self.opt.w_decay_G and hasattr(self, "arena_G"):         # torch.optim.Adam(weight_decay): g += wd * p before the moments
            self.arena_G.grad.add_(self.arena_G.flat, alpha=self.opt.w_decay_G)
        self.optimizer_G.step()

    def optimize_parameters(self, iters=0, fr=1):                   # translation_model.py:274-291
        self._graph_optimize()

    def _step_body(self):
        if self.device.type == "cuda":
            ops.zero_pool_reset(self.device)
        self.set_requires_grad(self.disc, False)
        for _ in range(self.opt.num_iter_gen):
            self.forward()
            self.optimizer_G.zero_grad()
            self.backward_G()
            self._adam_G()
        self.set_requires_grad(self.disc, True)
        self.set_requires_grad([self.netG_A, self.netG_B], False)
        for j in range(self.opt.num_iter_dis):
            if j > 0:
                with torch.no_grad():
                    self.forward()
            self.optimizer_D.zero_grad()
            self.backward_D_A()
            self.backward_D_B()
            self.optimizer_D.step()
        self.set_requires_grad([self.netG_A, self.netG_B], True)

    def calc_l_step(self):                                          # :293-298
        o = self.opt
        self.l_depth_A_step = abs(o.l_depth_A_begin - o.l_depth_A_end) / o.l_num_iter
        self.l_depth_B_step = abs(o.l_depth_B_begin - o.l_depth_B_end) / o.l_num_iter
        self.l_cycle_A_step = abs(o.l_cycle_A_begin - o.l_cycle_A_end) / o.l_num_iter
        self.l_cycle_B_step = abs(o.l_cycle_B_begin - o.l_cycle_B_end) / o.l_num_iter

    def update_loss_weight(self, global_iter):                      # :300-305
        if global_iter > self.opt.l_max_iter:
            gates = (self.l_depth_A > 0, self.l_depth_B > 0)
            self.l_depth_A -= self.l_depth_A_step
            self.l_depth_B -= self.l_depth_B_step
            self.l_cycle_A += self.l_cycle_A_step
            self.l_cycle_B += self.l_cycle_B_step
            self._sync_loss_weights()       # the captured step reads the weights from the device
            if gates != (self.l_depth_A > 0, self.l_depth_B > 0):
                self.reset_graph()          # a term entered / left the loss: the captured control flow is stale

    def calc_test_loss(self):                                       # :307-309
        md = self.opt.max_distance
        with torch.no_grad():
            self.test_depth_dif_A = self._masked_l1(data_to_meters(self.fake_depth_B.detach(), md), data_to_meters(self.real_depth_A, md), self.valid_A)
            self.test_depth_dif_B = self._masked_l1(data_to_meters(self.fake_depth_A.detach(), md), data_to_meters(self.real_depth_B, md), self.valid_B)

    def get_L1_loss(self):                                          # :311-312
        self.calc_test_loss()
        return float(self.test_depth_dif_A)

    def get_L1_loss_syn(self):                                      # :313-314
        self.calc_test_loss()
        return float(self.test_depth_dif_B)
