"""``MainModel`` - the image-guided depth-enhancement training step (``--model main_network_best``),
drop-in for the reference's ``models/main_model.py``.

Same constructor contract (an ``opt`` namespace with the reference's flags), same methods
(``set_input`` / ``forward`` / ``backward_G`` / ``optimize_parameters`` / ``calculate``), same
``loss_*`` / visual attribute names and the same five networks under the same checkpoint names.
What changed is how the step runs on the device (all citations into /root/reference):

* no host round trips inside the step: the rectangle-hole masks are still DRAWN on the host from
  ``np.random`` in the reference's exact call order (main_model.py:257-298) but rasterised on the
  device from a small int32 table; the monitoring scalars (:308-318) stay on the device until
  somebody calls ``float()`` on them; K^-1 / crop go to the device once per ``set_input``;
* the syn and real halves of each network call are batched into one call (every layer is
  per-sample: InstanceNorm / GroupNorm, so the result is unchanged);
* the trainable parameters (Depth_f + Task) live in one flat arena with one flat gradient arena:
  the backward kernels accumulate into it directly, Adam is one kernel over it, and the data
  parallel all-reduce (``parallel.GradBuckets``) works on contiguous slices of it.
"""
import itertools
from types import SimpleNamespace

import numpy as np
import torch

from . import networks, ops, translation_network
from .base_model import BaseModel
from .norms import camera_table

MAX_RECTS = 64          # randint(10, 60) draws at most 59 rectangles (main_model.py:262)


# ------------------------------------------------------------------------------------------------
# loss helpers with the reference's names (main_model.py:15-73)
# ------------------------------------------------------------------------------------------------
def tv_loss(img):
    return ops.tv_loss(img)


def get_smooth_weight(depth, Image, num_scales):
    return ops.smooth_loss(depth, Image, num_scales)


def draw_rects(batch, H, W, stage="train", rng=np.random):
    """The np.random call sequence of ONE of the two rectangle loops (main_model.py:261-267 /
    :282-288): randint(10,n) -> choice(W) -> choice(H) -> randint(W//150, W//8)*binomial(1,p) ->
    randint(H//150, H//8)*binomial(1,p).  Returns (rects int32 [batch, MAX_RECTS, 4], counts int32 [batch])."""
    n = 60 if stage == "train" else 11
    p = 0.90 if stage == "train" else 0
    rects = np.zeros((batch, MAX_RECTS, 4), dtype=np.int32)
    counts = np.zeros((batch,), dtype=np.int32)
    for i in range(batch):
        number = rng.randint(10, n)
        xs = rng.choice(W, number, replace=False)
        ys = rng.choice(H, number, replace=False)
        sizes_x = rng.randint(W // 150, W // 8, number) * rng.binomial(1, p)
        sizes_y = rng.randint(H // 150, H // 8, number) * rng.binomial(1, p)
        counts[i] = number
        rects[i, :number, 0], rects[i, :number, 1] = xs, ys
        rects[i, :number, 2], rects[i, :number, 3] = sizes_x, sizes_y
    return rects, counts


class LazyScalar:
    """A device scalar that only synchronises when converted (``float()``): replaces the four
    per-step ``.cpu().detach().numpy()`` reads of main_model.py:311-318."""

    def __init__(self, fn):
        self._fn = fn

    def __float__(self):
        return float(self._fn())

    def item(self):
        return float(self)


class ParamArena:
    """Flat fp32 storage for the trainable parameters, their gradients and the Adam moments.

    Order = expected gradient-ready order of the backward pass (reverse forward order: Task from its
    last layer to its first, then Depth_f), so contiguous slices complete in order and can be
    all-reduced while the rest of the backward is still running."""

    def __init__(self, nets, device):
        ordered = []
        for net in reversed(nets):                         # nets are given in forward order
            ordered += list(reversed(list(net.parameters())))
        self.params = ordered
        sizes = [p.numel() for p in ordered]
        # 16-byte alignment of every slice keeps float4 access legal in the kernels
        self.offsets, off = [], 0
        for n in sizes:
            self.offsets.append(off)
            off += (n + 3) // 4 * 4
        self.total = off
        self.flat = torch.zeros(off, device=device, dtype=torch.float32)
        self.grad = torch.zeros(off, device=device, dtype=torch.float32)
        self.exp_avg = torch.zeros(off, device=device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(off, device=device, dtype=torch.float32)
        self.index = {}
        for p, o in zip(ordered, self.offsets):
            n = p.numel()
            view = self.flat[o:o + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            gview = self.grad[o:o + n].view(p.shape)
            p.grad = gview
            self.index[p.data_ptr()] = (o, n)
            ops.DIRECT_GRADS[p.data_ptr()] = gview

    def zero_grad(self):
        self.grad.zero_()

    def release(self):
        for p in self.params:
            ops.DIRECT_GRADS.pop(p.data_ptr(), None)


class ArenaAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0) (main_model.py:176) as ONE
    kernel over the parameter arena.  A real ``Optimizer`` so ``get_scheduler`` / LambdaLR work."""

    def __init__(self, arena, lr):
        self.arena = arena
        super().__init__(arena.params, dict(lr=lr, betas=(0.9, 0.999), eps=1e-8))
        self.n_steps = 0
        self.grad_scale = 1.0

    def zero_grad(self, set_to_none=False):
        self.arena.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        self.n_steps += 1
        a = self.arena
        ops.adam_step(a.flat, a.grad, a.exp_avg, a.exp_avg_sq, g["lr"], self.n_steps, g["betas"][0], g["betas"][1],
                      g["eps"], self.grad_scale)
        ops.WEIGHT_EPOCH += 1            # packed 16-bit copies of the trainable weights are now stale


class MainModel(BaseModel):
    @staticmethod
    def modify_commandline_options(parser, is_train=True):         # main_model.py:79-87
        parser.set_defaults(no_dropout=True)
        if is_train:
            parser.add_argument("--lambda_A", type=float, default=10.0)
            parser.add_argument("--lambda_B", type=float, default=10.0)
            parser.add_argument("--lambda_identity", type=float, default=0.5)
        return parser

    def __init__(self, opt):                                        # main_model.py:89-177
        BaseModel.__init__(self, opt)
        self.loss_names = ["task_syn", "holes_syn", "holes_syn_l2", "task_real_by_depth", "task_real_by_image",
                           "syn_mean_diff", "real_mean_diff", "tv_syn_norm", "tv_real_norm", "syn_norms_holes",
                           "tv_syn_norm_old", "tv_real_norm_old", "syn_norms_old"]
        if opt.norm_loss:
            self.loss_names += ["syn_norms"]
        if opt.use_smooth_loss:
            self.loss_names += ["smooth"]
        if getattr(opt, "print_mean", False):
            self.loss_names = ["syn_mean_diff", "real_mean_diff", "mean_of_abs_diff_syn", "mean_of_abs_diff_real",
                               "L1_syn", "L1_real"]
        visual_names_A = ["syn_image", "syn_depth", "syn2real_depth", "syn_mask", "pred_syn_depth",
                          "mask_syn_add_holes", "syn_depth_by_image"]
        visual_names_B = ["real_image", "real_depth", "real_depth_by_image", "pred_real_depth", "real_mask",
                          "mask_real_add_holes"]
        if opt.norm_loss:
            visual_names_A += ["norm_syn", "norm_syn_pred", "norm_syn2real"]
            visual_names_B += ["norm_real", "norm_real_pred"]
        if opt.use_masked:
            visual_names_B += ["depth_masked"]
            visual_names_A += ["syn2real_depth_masked"]
            self.loss_names += ["holes_real", "holes_real_l2"]
        if getattr(opt, "use_edge", False):
            raise NotImplementedError("dsr_b200: --use_edge (CannyFilter is undefined in the reference too)")
        if not opt.use_image_for_trans or getattr(opt, "use_rec_as_real_input", False):
            raise NotImplementedError("dsr_b200: only the --use_image_for_trans path of main_network_best "
                                      "(README.md:70) is built; netG_B_d does not exist in the reference either")
        self.visual_names = visual_names_A + visual_names_B
        self.model_names = ["G_A_d", "I2D_features", "Image2Depth", "Task", "Depth_f"]
        self.border = -0.97

        self.netI2D_features = networks.define_G(3, opt.ImageDepthf_outf, opt.ImageDepthf_basef, opt.ImageDepthf_type,
                                                 opt.norm, not opt.no_dropout, opt.init_type, opt.init_gain,
                                                 self.gpu_ids, opt.replace_transpose)
        self.netImage2Depth = networks.define_G(opt.ImageDepthf_outf, 1, opt.I2D_base, opt.I2D_type, opt.norm,
                                                not opt.no_dropout, opt.init_type, opt.init_gain, self.gpu_ids,
                                                opt.replace_transpose, use_old=False)
        netG_B_opt = SimpleNamespace(ngf_img=32, ngf_depth=32, ngf=64, norm="group", dropout=False,
                                     init_type="normal", gpu_ids=opt.gpu_ids, input_nc_img=3, n_downsampling=2,
                                     use_semantic=False, n_blocks=9, upsampling_type="transpose",
                                     output_nc_depth=1, input_nc_depth=1)             # main_model.py:147
        self.netG_A_d = translation_network.define_Gen(netG_B_opt, input_type="img_depth")
        self.netDepth_f = networks.define_G(2, opt.Depthf_outf, opt.Depthf_basef, opt.Depthf_type, opt.norm,
                                            not opt.no_dropout, opt.init_type, opt.init_gain, self.gpu_ids,
                                            opt.replace_transpose, n_down=opt.Depthf_ndown)
        task_input_features = opt.ImageDepthf_outf + 5 + opt.Depthf_outf
        self.netTask = networks.define_G(task_input_features, 1, opt.Task_basef, opt.Task_type, opt.norm,
                                         not opt.no_dropout, opt.init_type, opt.init_gain, self.gpu_ids,
                                         opt.replace_transpose, n_down=opt.Task_ndown)
        self.loss_L1_syn = 0
        self.loss_L1_real = 0
        self.arena = None
        self.grad_sync = None            # set by parallel.GradBuckets for multi-GPU runs
        if self.isTrain:
            if self.gpu_ids:
                self._build_arena()
            else:   # host-only construction: a plain Adam holder so schedulers / checkpoints still work
                self.optimizer_G = torch.optim.Adam(itertools.chain(self.netDepth_f.parameters(),
                                                                    self.netTask.parameters()), lr=opt.lr)
            self.optimizers.append(self.optimizer_G)

    def _build_arena(self):
        self.arena = ParamArena([self._unwrap(self.netDepth_f), self._unwrap(self.netTask)], self.device)
        self.optimizer_G = ArenaAdam(self.arena, self.opt.lr)

    # ------------------------------------------------------------------------------------------
    def _h2d(self, t):
        t = t if t.dtype == torch.float32 else t.float()
        if t.device.type == "cpu" and self.device.type == "cuda":
            if not t.is_pinned():
                t = t.pin_memory()
            return t.to(self.device, non_blocking=True)
        return t.to(self.device)

    def set_input(self, input):                                     # main_model.py:179-201
        AtoB = self.opt.direction == "AtoB"
        self.syn_image = self._h2d(input["A_i" if AtoB else "B_i"])
        self.real_image = self._h2d(input["B_i" if AtoB else "A_i"])
        self.syn_depth = self._h2d(input["A_d" if AtoB else "B_d"])
        self.real_depth = self._h2d(input["B_d" if AtoB else "A_d"])
        self.image_paths = input["A_paths" if AtoB else "B_paths"]
        self.A_paths = input["A_paths"]
        self.B_paths = input["B_paths"]
        self.K_A, self.K_B = input["K_A"], input["K_B"]
        self.crop_A, self.crop_B = input["crop_A"], input["crop_B"]
        # the fp64 K^-1 / crop table the normal kernels read (norms.py:75-89), uploaded once per batch
        self.cam_A = camera_table(self.K_A, self.crop_A, 0.5, self.device)
        self.cam_B = camera_table(self.K_B, self.crop_B, 0.5, self.device)

    def _upload_rects(self, rects, counts):
        r = torch.from_numpy(rects)
        c = torch.from_numpy(counts)
        if self.device.type == "cuda":
            r, c = r.pin_memory(), c.pin_memory()
        return r.to(self.device, non_blocking=True), c.to(self.device, non_blocking=True)

    def forward(self, stage="train"):                               # main_model.py:204-336
        B, _, H, W = self.real_depth.shape
        self.real_hole_mask, self.real_mask = ops.hole_valid_masks(self.real_depth, self.border)
        _, self.syn_mask = ops.hole_valid_masks(self.syn_depth, self.border)

        with torch.no_grad():                                       # frozen nets (main_model.py:426)
            self.syn2real_depth = self.netG_A_d(self.syn_depth, self.syn_image)
            images = torch.cat([self.syn_image, self.real_image], 0)
            image_features = self.netI2D_features(images)
            depth_by_image = self.netImage2Depth(image_features)
        self.syn_depth_by_image, self.real_depth_by_image = depth_by_image[:B], depth_by_image[B:]

        if not self.opt.use_masked:
            raise NotImplementedError("dsr_b200: --use_masked is required (backward_G of the reference needs gt_mask_syn)")
        # host RNG in the reference's order: real loop first, then syn (main_model.py:257-298)
        rr, rc = draw_rects(B, H, W, stage)
        sr, sc = draw_rects(B, H, W, stage)
        rr, rc = self._upload_rects(rr, rc)
        sr, sc = self._upload_rects(sr, sc)
        self.gt_mask_real, self.depth_masked, self._a_r = ops.rect_holes(self.real_mask, self.real_depth, rr, rc, MAX_RECTS)
        self.gt_mask_syn, self.syn2real_depth_masked, self._a_s = ops.rect_holes(
            self.syn_mask, self.syn2real_depth, sr, sc, MAX_RECTS, extra_border=self.border)

        d_in = ops.cat([torch.cat([self.syn2real_depth_masked, self.depth_masked], 0), depth_by_image])
        feat_depth = self.netDepth_f(d_in)
        pred = self.netTask(ops.LazyCat([image_features, feat_depth, d_in, images]))
        self.pred_syn_depth, self.pred_real_depth = pred[:B], pred[B:]

        n = float(self.syn_depth.numel())
        s_syn = ops.masked_sums(self.syn_depth, self.pred_syn_depth, self.syn_mask)
        s_real = ops.masked_sums(self.real_depth, self.pred_real_depth, self.real_mask)
        self.loss_syn_mean_diff = LazyScalar(lambda: (s_syn[0] - s_syn[1]) / n)
        self.loss_mean_of_abs_diff_syn = LazyScalar(lambda: s_syn[2] / n)
        self.loss_real_mean_diff = LazyScalar(lambda: (s_real[0] - s_real[1]) / n)
        self.loss_mean_of_abs_diff_real = LazyScalar(lambda: s_real[2] / n)
        if getattr(self.opt, "save_all", False) and stage == "test":
            raise NotImplementedError("dsr_b200: PNG export (--save_all) is a 'next' row (SURVEY.md section 8f.4)")

    def backward_G(self, back=True):                                # main_model.py:340-419
        opt = self.opt
        if not opt.norm_loss:
            raise NotImplementedError("dsr_b200: --norm_loss is required (loss_tv_syn_norm is undefined without it "
                                      "in the reference as well)")
        ms, mr = self.syn_mask, self.real_mask
        ps, pr = self.pred_syn_depth, self.pred_real_depth
        # image-space normals x100 (:343-352)
        n_syn = ops.normals_old(self.syn_depth, 100.0)
        n_syn_pred = ops.normals_old(ps, 100.0)
        n_real_pred = ops.normals_old(pr, 100.0)
        self.loss_tv_syn_norm_old = tv_loss(n_syn_pred) * (10 ** -7)
        self.loss_tv_real_norm_old = tv_loss(n_real_pred) * (10 ** -7)
        self.loss_syn_norms_old = ops.masked_l1_l2(n_syn, n_syn_pred, ms)[1]
        a_s = self._a_s                                             # mask_syn_add_holes (:354-357)
        # camera-space normals (:360-372)
        self.norm_syn = ops.normals_new(self.syn_depth, self.cam_A)
        self.norm_syn2real = ops.normals_new(self.syn2real_depth_masked, self.cam_A)
        self.norm_syn_pred = ops.normals_new(ps, self.cam_A)
        self.norm_real = ops.normals_new(self.real_depth, self.cam_B)
        self.norm_real_pred = ops.normals_new(pr, self.cam_B)
        self.loss_tv_syn_norm = tv_loss(self.norm_syn_pred) * (10 ** -7)
        self.loss_tv_real_norm = tv_loss(self.norm_real_pred) * (10 ** -7)
        self.loss_syn_norms = ops.masked_l1_l2(self.norm_syn, self.norm_syn_pred, ms)[0]
        self.loss_syn_norms_holes = ops.masked_l1_l2(self.norm_syn, self.norm_syn_pred, ms, a_s)[0]
        # depth terms (:383-390)
        hs = ops.masked_l1_l2(self.syn_depth, ps, ms, a_s)
        self.loss_holes_syn = hs[0]
        self.loss_holes_syn_l2 = hs[1] * 5
        self.loss_task_syn = ops.masked_l1_l2(self.syn_depth, ps, ms)[0]
        self.loss_task_real_by_depth = ops.masked_l1_l2(self.real_depth, pr, mr)[0]
        self.loss_task_real_by_image = ops.masked_l1_l2(self.real_depth_by_image, pr, self.real_hole_mask)[0]
        self.loss_G = (self.loss_task_syn * opt.w_syn_l1 + self.loss_holes_syn * opt.w_syn_holes
                       + opt.w_syn_holes * self.loss_holes_syn_l2 + self.loss_task_real_by_depth * opt.w_real_l1_d
                       + self.loss_task_real_by_image * opt.w_real_l1_i + self.loss_tv_syn_norm * 1
                       + self.loss_syn_norms_holes * opt.w_syn_norm * 5 + self.loss_tv_real_norm * 1
                       + self.loss_syn_norms_old * opt.w_syn_norm + self.loss_tv_real_norm_old * 1
                       + self.loss_tv_syn_norm_old * 1)             # :393
        hr = ops.masked_l1_l2(self.real_depth, pr, self._a_r)       # mask_real_add_holes (:396-398)
        self.loss_holes_real = hr[0]
        self.loss_holes_real_l2 = hr[1] * 5
        self.loss_G = self.loss_G + self.loss_holes_real * opt.w_real_holes + self.loss_holes_real_l2 * opt.w_real_holes
        self.loss_G = self.loss_G + self.loss_syn_norms * opt.w_syn_norm                      # :404
        if opt.use_smooth_loss:
            self.loss_smooth = get_smooth_weight(pr, self.real_image, 3)                      # :407
            self.loss_G = self.loss_G + self.loss_smooth * opt.w_smooth
        self.loss_G = self.loss_G * opt.scale_G                                               # :417
        if back:
            self.loss_G.backward()

    # visuals the reference overwrites after the losses (:386, :400) - computed only when read
    @property
    def mask_syn_add_holes(self):
        return (self.pred_syn_depth * self.syn_mask * self._a_s).detach()

    @property
    def mask_real_add_holes(self):
        return (self.pred_real_depth * self._a_r).detach()

    def optimize_parameters(self, iters=0, fr=1):                   # main_model.py:422-429
        self.forward()
        self.set_requires_grad([self.netG_A_d, self.netI2D_features, self.netImage2Depth], False)
        self.optimizer_G.zero_grad()
        self.backward_G()
        if self.grad_sync is not None:
            self.grad_sync.finish()
        self.optimizer_G.step()

    def calculate(self, stage="test"):                              # main_model.py:433-436
        self.forward(stage)
        self.backward_G(back=False)
