"""``MainSRModel`` - the image-guided depth super-resolution fine-tune step (``--SR``), drop-in for the
reference's ``models/main_sr_model.py`` (all citations into /root/reference).

Same five networks, optimizer, checkpoint names and method surface as ``MainModel``; the deltas
(SURVEY.md section 8a row A18):

* depth maps, syn images and the real image arrive at HIGH resolution (2h x 2w, h/w = ``opt.crop_size_h/w``);
  ``G_A_d``, ``Depth_f`` and ``Task`` run at HR, ``I2D_features`` / ``Image2Depth`` run on the bicubic x0.5 images
  and their outputs are bicubic x2 up-sampled (main_sr_model.py:279-293);
* ``pred_real_depth`` is the bicubic x0.5 of ``pred_real_depth_hr`` ON the gradient path (:361); the real-side
  losses are evaluated at LR on nearest-down-sampled masks and bicubic-down-sampled depth / image (:394-398);
* rectangle sizes ``// 10`` and p = 0.95 for the real loop (:299-305, :320-326);
* ``syn_norms`` = MSE(norm_syn, norm_real_pred_hr) under syn_mask (:434), ``syn_norms_old`` = un-masked L1 (:410),
  ``tv_real_norm_old`` on the HR prediction (:409), ``task_real_by_image`` against the nearest-down-sampled
  ``syn_depth`` (:452), loss weights x2 / x5 (:455).

Every resize is one launch of ``csrc/resize.cu`` (no ATen interpolate kernels on the path).
"""
import torch

from . import ops
from .main_model import MAX_RECTS, LazyScalar, MainModel, get_smooth_weight, tv_loss
from .norms import camera_table  # noqa: F401  (re-exported for callers that mirror the reference's imports)


class MainSRModel(MainModel):
    RECT_REAL = dict(p_train=0.95, div=10)         # main_sr_model.py:299,304-305
    RECT_SYN = dict(p_train=0.90, div=10)          # main_sr_model.py:320,325-326

    def __init__(self, opt):                                        # main_sr_model.py:97-204
        if getattr(opt, "use_D", False):
            raise NotImplementedError("dsr_b200: --use_D (loss_G_pred is never computed in the reference's "
                                      "MainSRModel either; the flag cannot train there)")
        MainModel.__init__(self, opt)
        self._ones = {}

    def _one_mask(self, like):
        key = tuple(like.shape[:1]) + tuple(like.shape[2:])
        if key not in self._ones:
            self._ones[key] = torch.ones((like.shape[0], 1) + tuple(like.shape[2:]), device=like.device, dtype=torch.float32)
        return self._ones[key]

    def forward(self, stage="train"):                               # main_sr_model.py:228-372
        opt = self.opt
        h, w = opt.crop_size_h, opt.crop_size_w
        B, _, H, W = self.real_depth.shape
        if (H, W) != (2 * h, 2 * w) or tuple(self.syn_depth.shape[2:]) != (H, W):
            raise ValueError(f"MainSRModel expects {2 * h}x{2 * w} inputs (2 x crop_size), got {H}x{W}")
        train = stage == "train"
        self.real_hole_mask, self.real_mask = ops.hole_valid_masks(self.real_depth, self.border)
        _, self.syn_mask = ops.hole_valid_masks(self.syn_depth, self.border)

        with torch.no_grad():                                       # frozen nets (:492)
            self.syn2real_depth = self.netG_A_d(self.syn_depth, self.syn_image)
            self._real_image_lr = ops.bicubic(self.real_image, (h, w))
            if train:
                images_lr = torch.cat([ops.bicubic(self.syn_image, (h, w)), self._real_image_lr], 0)
                images = torch.cat([self.syn_image, self.real_image], 0)
            else:
                images_lr, images = self._real_image_lr, self.real_image
            feats_lr = self.netI2D_features(images_lr)
            dbi_lr = self.netImage2Depth(feats_lr)
            depth_by_image = ops.bicubic(dbi_lr, (H, W))
            image_features = ops.bicubic(feats_lr, (H, W))
        if train:
            self.syn_depth_by_image, self.real_depth_by_image = depth_by_image[:B], depth_by_image[B:]
        else:
            self.real_depth_by_image = depth_by_image

        if not opt.use_masked:
            raise NotImplementedError("dsr_b200: --use_masked is required (backward_G of the reference needs gt_mask_syn)")
        if not self._rects_staged:
            self._stage_rects(B, H, W, stage)
        self._rects_staged = False
        rr, rc, sr, sc = self._rect["dev"]
        self.gt_mask_real, self.depth_masked, self._a_r = ops.rect_holes(self.real_mask, self.real_depth, rr, rc, MAX_RECTS)
        self.gt_mask_syn, self.syn2real_depth_masked, self._a_s = ops.rect_holes(
            self.syn_mask, self.syn2real_depth, sr, sc, MAX_RECTS, extra_border=self.border)

        if train:
            d_in = ops.cat([torch.cat([self.syn2real_depth_masked, self.depth_masked], 0), depth_by_image])
        else:
            d_in = ops.cat([self.depth_masked, depth_by_image])
        feat_depth = self.netDepth_f(d_in)
        pred = self.netTask(ops.LazyCat([image_features, feat_depth, d_in, images]))
        if not train:
            self.pred_real_depth_hr = pred
            if getattr(opt, "save_all", False):                      # main_sr_model.py:376-386: HR prediction, rows [32, H - 32)
                from . import io
                self.saved_files = io.save_predictions(pred, self.B_paths, opt.save_image_folder, 32)
            return
        self.pred_syn_depth, self.pred_real_depth_hr = pred[:B], pred[B:]
        self.pred_real_depth = ops.bicubic(self.pred_real_depth_hr, (h, w))                   # :361, on the gradient path

        n_hr, n_lr = float(self.syn_depth.numel()), float(B * h * w)
        s_syn = ops.masked_sums(self.syn_depth, self.pred_syn_depth, self.syn_mask)
        rd_lr = ops.bicubic(self.real_depth, (h, w))
        rm_bic = ops.bicubic(self.real_mask, (h, w))                # bicubic (not nearest) of the mask, as in :368-372
        s_real = ops.masked_sums(rd_lr, self.pred_real_depth, rm_bic)
        self._real_depth_lr = rd_lr
        self.loss_syn_mean_diff = LazyScalar(lambda: (s_syn[0] - s_syn[1]) / n_hr)
        self.loss_mean_of_abs_diff_syn = LazyScalar(lambda: s_syn[2] / n_hr)
        self.loss_real_mean_diff = LazyScalar(lambda: (s_real[0] - s_real[1]) / n_lr)
        self.loss_mean_of_abs_diff_real = LazyScalar(lambda: s_real[2] / n_lr)

    def backward_G(self, back=True):                                # main_sr_model.py:391-484
        opt = self.opt
        if not opt.norm_loss:
            raise NotImplementedError("dsr_b200: --norm_loss is required (loss_tv_syn_norm is undefined without it "
                                      "in the reference as well)")
        h, w = opt.crop_size_h, opt.crop_size_w
        # the reference overwrites the real-side attributes with their LR versions (:394-398)
        self.real_mask = ops.nearest(self.real_mask, (h, w))
        self.real_hole_mask = ops.nearest(self.real_hole_mask, (h, w))
        self.real_depth = self._real_depth_lr
        self.real_image = self._real_image_lr
        self.real_depth_by_image = ops.bicubic(self.real_depth_by_image, (h, w))
        ms, mr = self.syn_mask, self.real_mask
        ps, pr, pr_hr = self.pred_syn_depth, self.pred_real_depth, self.pred_real_depth_hr
        # image-space normals x100 (:400-410)
        n_syn = ops.normals_old(self.syn_depth, 100.0)
        n_syn_pred = ops.normals_old(ps, 100.0)
        n_real_pred_hr = ops.normals_old(pr_hr, 100.0)
        self.loss_tv_syn_norm_old = tv_loss(n_syn_pred) * (10 ** -7)
        self.loss_tv_real_norm_old = tv_loss(n_real_pred_hr) * (10 ** -7)
        self.loss_syn_norms_old = ops.masked_l1_l2(n_syn, n_syn_pred, self._one_mask(n_syn))[0]
        a_s = self._a_s                                             # mask_syn_add_holes (:412-415)
        # camera-space normals (:422-435)
        self.norm_syn = ops.normals_new(self.syn_depth, self.cam_A)
        self.norm_syn2real = ops.normals_new(self.syn2real_depth_masked, self.cam_A)
        self.norm_syn_pred = ops.normals_new(ps, self.cam_A)
        self.norm_real = ops.normals_new(self.real_depth, self.cam_B)
        self.norm_real_pred = ops.normals_new(pr, self.cam_B)
        self.norm_real_pred_hr = ops.normals_new(pr_hr, self.cam_A)
        self.loss_tv_syn_norm = tv_loss(self.norm_syn_pred) * (10 ** -7)
        self.loss_tv_real_norm = tv_loss(self.norm_real_pred) * (10 ** -7)
        self.loss_syn_norms = ops.masked_l1_l2(self.norm_syn, self.norm_real_pred_hr, ms)[1]
        self.loss_syn_norms_holes = ops.masked_l1_l2(self.norm_syn, self.norm_syn_pred, ms, a_s)[0]
        # depth terms (:445-452)
        hs = ops.masked_l1_l2(self.syn_depth, ps, ms, a_s)
        self.loss_holes_syn = hs[0]
        self.loss_holes_syn_l2 = hs[1] * 5
        self.loss_task_syn = ops.masked_l1_l2(self.syn_depth, ps, ms)[0]
        self.loss_task_real_by_depth = ops.masked_l1_l2(self.real_depth, pr, mr)[0]
        self.loss_task_real_by_image = ops.masked_l1_l2(ops.nearest(self.syn_depth, (h, w)), pr, self.real_hole_mask)[0]
        self.loss_G = (self.loss_task_syn * opt.w_syn_l1 + self.loss_holes_syn * opt.w_syn_holes
                       + opt.w_syn_holes * self.loss_holes_syn_l2 + self.loss_task_real_by_depth * opt.w_real_l1_d
                       + self.loss_task_real_by_image * opt.w_real_l1_i + self.loss_tv_syn_norm * 1
                       + self.loss_syn_norms_holes * opt.w_syn_norm * 5 + self.loss_tv_real_norm * 2
                       + self.loss_syn_norms_old * opt.w_syn_norm * 5 + self.loss_tv_real_norm_old * 2
                       + self.loss_tv_syn_norm_old * 1)             # :455
        self._a_r_lr = ops.nearest(self._a_r, (h, w))               # mask_real_add_holes (:458-459)
        hr = ops.masked_l1_l2(self.real_depth, pr, self._a_r_lr)
        self.loss_holes_real = hr[0]
        self.loss_holes_real_l2 = hr[1] * 5
        self.loss_G = self.loss_G + self.loss_holes_real * opt.w_real_holes + self.loss_holes_real_l2 * opt.w_real_holes
        self.loss_G = self.loss_G + self.loss_syn_norms * opt.w_syn_norm                      # :469
        if opt.use_smooth_loss:
            self.loss_smooth = get_smooth_weight(pr, self.real_image, 3)                      # :472
            self.loss_G = self.loss_G + self.loss_smooth * opt.w_smooth
        self.loss_G = self.loss_G * opt.scale_G                                               # :482
        if back:
            self.loss_G.backward()
            ops.join_side()

    @property
    def mask_real_add_holes(self):                                  # :463
        return (self.pred_real_depth * self._a_r_lr).detach()

    def calculate(self, stage="test"):                              # main_sr_model.py:502-506 (no backward_G)
        self.forward(stage)
