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
import weakref
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


def draw_rects(batch, H, W, stage="train", rng=np.random, p_train=0.90, div=8):
    """The np.random call sequence of ONE of the two rectangle loops (main_model.py:261-267 /
    :282-288): randint(10,n) -> choice(W) -> choice(H) -> randint(W//150, W//div)*binomial(1,p) ->
    randint(H//150, H//div)*binomial(1,p).  Returns (rects int32 [batch, MAX_RECTS, 4], counts int32 [batch]).
    (main_sr_model.py:298-305, :319-326 use div = 10 and p = 0.95 for the real loop.)"""
    n = 60 if stage == "train" else 11
    p = p_train if stage == "train" else 0
    rects = np.zeros((batch, MAX_RECTS, 4), dtype=np.int32)
    counts = np.zeros((batch,), dtype=np.int32)
    for i in range(batch):
        number = rng.randint(10, n)
        xs = rng.choice(W, number, replace=False)
        ys = rng.choice(H, number, replace=False)
        sizes_x = rng.randint(W // 150, W // div, number) * rng.binomial(1, p)
        sizes_y = rng.randint(H // 150, H // div, number) * rng.binomial(1, p)
        counts[i] = number
        rects[i, :number, 0], rects[i, :number, 1] = xs, ys
        rects[i, :number, 2], rects[i, :number, 3] = sizes_x, sizes_y
    return rects, counts


def _scaled(t, w):
    """a loss attribute that is `w * t` (read by logging only): the product is formed when somebody asks for the value"""
    d = t.detach()                 # (a view: no kernel; the closure must not keep the step's autograd graph alive)
    return LazyScalar(lambda: d * w)


def _picked(pair, i, w=1.0):
    """element i of a (L1, MSE) pair, times w, formed on demand"""
    d = pair.detach()
    return LazyScalar(lambda: d[i] * w)


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
        self._registered = []
        for p, o in zip(ordered, self.offsets):
            n = p.numel()
            view = self.flat[o:o + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            gview = self.grad[o:o + n].view(p.shape)
            p.grad = gview
            self.index[p.data_ptr()] = (o, n)
            ops.DIRECT_GRADS[p.data_ptr()] = gview
            self._registered.append((p.data_ptr(), gview))

    def zero_grad(self):
        self.grad.zero_()

    def release(self):
        """unregister the gradient views (called by a finalizer when the owning model is dropped): a dead model's 177 MB
        gradient arena must not stay alive through ops.DIRECT_GRADS, and a later parameter that lands on an old address must
        not inherit a stale gradient target"""
        for ptr, gview in self._registered:
            if ops.DIRECT_GRADS.get(ptr) is gview:
                del ops.DIRECT_GRADS[ptr]
        self._registered = []


class ArenaAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0) (main_model.py:176) as ONE
    kernel over the parameter arena.  A real ``Optimizer`` so ``get_scheduler`` / LambdaLR work."""

    def __init__(self, arena, lr):
        self.arena = arena
        super().__init__(arena.params, dict(lr=lr, betas=(0.9, 0.999), eps=1e-8))
        self.grad_scale = 1.0
        dev = arena.flat.device
        # the step counter and (lr, beta1, beta2, eps) live on the device: the kernel launch never changes, so the
        # step can be replayed as a CUDA graph; the host mirrors are only for bookkeeping / LR schedulers
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.hyper_dev = torch.zeros(4, device=dev, dtype=torch.float64)
        self._hyper_host = torch.zeros(4, dtype=torch.float64).pin_memory() if dev.type == "cuda" else torch.zeros(4, dtype=torch.float64)
        self._hyper_key = None
        self.sync_hyper()

    @property
    def n_steps(self):
        return int(self.step_dev.item())

    def sync_hyper(self):
        """upload (lr, betas, eps) when a scheduler or the user changed them (host-side, outside any graph)"""
        g = self.param_groups[0]
        key = (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]))
        if key != self._hyper_key:
            self._hyper_host.copy_(torch.tensor(key, dtype=torch.float64))
            self.hyper_dev.copy_(self._hyper_host, non_blocking=True)
            self._hyper_key = key

    def zero_grad(self, set_to_none=False):
        self.arena.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
        a = self.arena
        ops.adam_step_dev(a.flat, a.grad, a.exp_avg, a.exp_avg_sq, self.hyper_dev, self.step_dev, self.grad_scale)
        ops.WEIGHT_EPOCH += 1            # packed 16-bit copies of the trainable weights are now stale


class MainModel(BaseModel):
    RECT_REAL = dict(p_train=0.90, div=8)          # main_model.py:260,265-266
    RECT_SYN = dict(p_train=0.90, div=8)           # main_model.py:281,286-287

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
        self.tv_scale = 1.0              # data parallel: the TV terms are batch SUMS (main_model.py:15-19), every other term a
                                         # batch mean - a rank weighs its TV terms by the world size so that the rank-averaged
                                         # gradient equals the gradient of the reference's loss on the concatenated batch
        self.rect_override = None        # (rects_real, counts_real, rects_syn, counts_syn): explicit tables instead of np.random
        # CUDA-graph replay of the whole training step (static input / rectangle buffers, device-side Adam state)
        self.use_graph = bool(getattr(opt, "cuda_graph", False))
        self.graph_warmup = 2            # eager steps before the capture (fills caches, raises smem limits)
        self._graph = None
        self._gstream = None
        self._eager_steps = 0
        self._in = None                  # persistent device inputs
        self._rect = None                # persistent pinned + device rectangle tables
        self._rects_staged = False
        self._side = None                # second stream for the independent frozen chain (forward)
        self._pipe = None                # pipelined graph replay: two input slots, a frozen graph + a training graph per slot
        self._nonfinite = None           # device counter of NaN / Inf values of loss_G (a graph replay cannot raise by itself)
        if self.isTrain:
            if self.gpu_ids:
                self._build_arena()
            else:   # host-only construction: a plain Adam holder so schedulers / checkpoints still work
                self.optimizer_G = torch.optim.Adam(itertools.chain(self.netDepth_f.parameters(),
                                                                    self.netTask.parameters()), lr=opt.lr)
            self.optimizers.append(self.optimizer_G)

    def _build_arena(self):
        self.arena = ParamArena([self._unwrap(self.netDepth_f), self._unwrap(self.netTask)], self.device)
        weakref.finalize(self, ParamArena.release, self.arena)
        self.optimizer_G = ArenaAdam(self.arena, self.opt.lr)

    # ------------------------------------------------------------------------------------------
    def _pinned(self, t):
        t = t if t.dtype == torch.float32 else t.float()
        if t.device.type == "cpu" and self.device.type == "cuda" and not t.is_pinned():
            t = t.pin_memory()
        return t

    def set_input(self, input):                                     # main_model.py:179-201
        AtoB = self.opt.direction == "AtoB"
        src = dict(syn_image=input["A_i" if AtoB else "B_i"], real_image=input["B_i" if AtoB else "A_i"],
                   syn_depth=input["A_d" if AtoB else "B_d"], real_depth=input["B_d" if AtoB else "A_d"])
        self.image_paths = input["A_paths" if AtoB else "B_paths"]
        self.A_paths = input["A_paths"]
        self.B_paths = input["B_paths"]
        self.K_A, self.K_B = input["K_A"], input["K_B"]
        self.crop_A, self.crop_B = input["crop_A"], input["crop_B"]
        # the fp64 K^-1 / crop table the normal kernels read (norms.py:75-89), uploaded once per batch
        cams = dict(cam_A=camera_table(self.K_A, self.crop_A, 0.5), cam_B=camera_table(self.K_B, self.crop_B, 0.5))
        shapes = {k: tuple(v.shape) for k, v in src.items()}
        if self._pipeline_on():
            return self._set_input_pipe(src, cams, shapes)
        if self._in is None or self._in["shapes"] != shapes:
            if self._graph is not None:
                raise RuntimeError("dsr_b200: the captured CUDA graph is bound to the first batch shape; "
                                   "call reset_graph() before changing it")
            self._in = dict(shapes=shapes)
            for k, v in src.items():
                self._in[k] = torch.empty(v.shape, device=self.device, dtype=torch.float32)
            for k, v in cams.items():
                self._in[k] = torch.empty(v.shape, device=self.device, dtype=torch.float64)
        # persistent device tensors (same addresses every step: no allocation, graph-replay safe)
        host = [k for k, v in src.items() if v.device.type == "cpu"] if self.device.type == "cuda" else []
        if host:
            # host inputs go H2D on a COPY stream into a two-slot staging ring, so the transfer of batch i+1 overlaps the
            # compute of batch i (the host runs ahead of the device); the compute stream then moves the slot into the fixed
            # input buffers with one on-device copy per tensor (~2 us each)
            st = self._in.get("stage")
            if st is None:
                st = self._in["stage"] = dict(stream=torch.cuda.Stream(), turn=0, slots=[
                    dict(buf={k: torch.empty_like(self._in[k]) for k in src}, ready=torch.cuda.Event(), consumed=torch.cuda.Event())
                    for _ in range(2)])
                for sl in st["slots"]:
                    sl["consumed"].record()
            sl = st["slots"][st["turn"] % 2]
            st["turn"] += 1
            st["stream"].wait_event(sl["consumed"])
            with torch.cuda.stream(st["stream"]):
                for k in host:
                    sl["buf"][k].copy_(self._pinned(src[k]), non_blocking=True)
                sl["ready"].record()
            torch.cuda.current_stream().wait_event(sl["ready"])
        for k, v in src.items():
            self._in[k].copy_(sl["buf"][k] if k in host else self._pinned(v), non_blocking=True)
            setattr(self, k, self._in[k])
        if host:
            sl["consumed"].record()
        for k, v in cams.items():
            self._in[k].copy_(v.pin_memory() if self.device.type == "cuda" else v, non_blocking=True)
            setattr(self, k, self._in[k])

    def _stage_rects(self, B, H, W, stage):
        """host RNG in the reference's order - real loop first, then syn (main_model.py:257-298) - into persistent
        pinned tables, then one async H2D copy each"""
        if self.rect_override is not None:
            rr, rc, sr, sc = (np.ascontiguousarray(a, dtype=np.int32) for a in self.rect_override)
        else:
            rr, rc = draw_rects(B, H, W, stage, **self.RECT_REAL)
            sr, sc = draw_rects(B, H, W, stage, **self.RECT_SYN)
        cuda = self.device.type == "cuda"
        if self._rect is None or self._rect["B"] != B:
            mk = lambda shape: torch.zeros(shape, dtype=torch.int32).pin_memory() if cuda else torch.zeros(shape, dtype=torch.int32)
            # a small ring of pinned staging tables: the host may run several steps ahead of the device (graph replay),
            # so a slot is rewritten only after the copy that read it has completed
            ring = [dict(host=[mk((B, MAX_RECTS, 4)), mk((B,)), mk((B, MAX_RECTS, 4)), mk((B,))],
                         done=torch.cuda.Event() if cuda else None) for _ in range(4)]
            self._rect = dict(B=B, ring=ring, turn=0)
            self._rect["dev"] = [torch.zeros_like(t, device=self.device) for t in ring[0]["host"]]
        slot = self._rect["ring"][self._rect["turn"] % 4]
        self._rect["turn"] += 1
        if cuda:
            slot["done"].synchronize()
        for h, d, a in zip(slot["host"], self._rect["dev"], (rr, rc, sr, sc)):
            h.copy_(torch.from_numpy(a))
            d.copy_(h, non_blocking=True)
        if cuda:
            slot["done"].record()
        self._rects_staged = True

    def _forward_frozen(self):
        """The frozen networks (main_model.py:232-254, no_grad per :426): G_A_d, I2D_features -> Image2Depth.  They read the
        batch only - not the trainable weights - so a captured training loop runs them for batch i+1 beside the training
        part of batch i (``_optimize_pipe``)."""
        with torch.no_grad():
            # G_A_d and the image branch (I2D_features -> Image2Depth) are independent: they run on two streams (an event
            # fork / join, also inside a captured graph), so the CTAs of one chain fill the partial waves and the small
            # launches of the other
            fork = ops.CONFIG["fork_frozen"] and self.device.type == "cuda"
            if fork:
                if self._side is None:
                    self._side = torch.cuda.Stream()
                cur = torch.cuda.current_stream()
                self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    self.syn2real_depth = self.netG_A_d(self.syn_depth, self.syn_image)
            else:
                self.syn2real_depth = self.netG_A_d(self.syn_depth, self.syn_image)
            self._images = torch.cat([self.syn_image, self.real_image], 0)
            self._image_features = self.netI2D_features(self._images)
            self._depth_by_image = self.netImage2Depth(self._image_features)
            if fork:
                cur.wait_stream(self._side)

    def forward(self, stage="train"):                               # main_model.py:204-336
        self._forward_frozen()
        self._forward_train(stage)

    def _forward_train(self, stage="train"):
        B, _, H, W = self.real_depth.shape
        self.real_hole_mask, self.real_mask = ops.hole_valid_masks(self.real_depth, self.border)
        _, self.syn_mask = ops.hole_valid_masks(self.syn_depth, self.border)
        images, image_features, depth_by_image = self._images, self._image_features, self._depth_by_image
        self.syn_depth_by_image, self.real_depth_by_image = depth_by_image[:B], depth_by_image[B:]

        if not self.opt.use_masked:
            raise NotImplementedError("dsr_b200: --use_masked is required (backward_G of the reference needs gt_mask_syn)")
        if not self._rects_staged:          # (the graph driver stages them before replaying)
            self._stage_rects(B, H, W, stage)
        self._rects_staged = False
        rr, rc, sr, sc = self._rect["dev"]
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
        if getattr(self.opt, "save_all", False) and stage == "test":               # main_model.py:321-333
            from . import io
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("dsr_b200: --save_all writes files: use forward('test'), not forward_test_graph()")
            self.saved_files = io.save_predictions(self.pred_real_depth, self.B_paths, self.opt.save_image_folder, 16)

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
        # The reference multiplies / adds ~40 scalar tensors here (main_model.py:352-417); on the device each of those is a
        # ~2 us launch on the critical path between the forward and the backward pass.  The raw terms come out of the loss
        # kernels as tiny device tensors, ONE launch (ops.loss_sum) forms loss_G from them and one more hands every term its
        # gradient; the scaled per-term attributes (loss_tv_*, loss_holes_*_l2) become lazy products, read only by logging.
        tvw = (10 ** -7) * self.tv_scale
        tv_syn_old, tv_real_old = tv_loss(n_syn_pred), tv_loss(n_real_pred)
        self.loss_tv_syn_norm_old = _scaled(tv_syn_old, tvw)
        self.loss_tv_real_norm_old = _scaled(tv_real_old, tvw)
        norms_old = ops.masked_l1_l2(n_syn, n_syn_pred, ms)
        self.loss_syn_norms_old = _picked(norms_old, 1)
        a_s = self._a_s                                             # mask_syn_add_holes (:354-357)
        # camera-space normals (:360-372)
        self.norm_syn = ops.normals_new(self.syn_depth, self.cam_A)
        self.norm_syn2real = ops.normals_new(self.syn2real_depth_masked, self.cam_A)
        self.norm_syn_pred = ops.normals_new(ps, self.cam_A)
        self.norm_real = ops.normals_new(self.real_depth, self.cam_B)
        self.norm_real_pred = ops.normals_new(pr, self.cam_B)
        tv_syn, tv_real = tv_loss(self.norm_syn_pred), tv_loss(self.norm_real_pred)
        self.loss_tv_syn_norm = _scaled(tv_syn, tvw)
        self.loss_tv_real_norm = _scaled(tv_real, tvw)
        norms = ops.masked_l1_l2(self.norm_syn, self.norm_syn_pred, ms)
        norms_holes = ops.masked_l1_l2(self.norm_syn, self.norm_syn_pred, ms, a_s)
        self.loss_syn_norms = _picked(norms, 0)
        self.loss_syn_norms_holes = _picked(norms_holes, 0)
        # depth terms (:383-390)
        hs = ops.masked_l1_l2(self.syn_depth, ps, ms, a_s)
        self.loss_holes_syn = _picked(hs, 0)
        self.loss_holes_syn_l2 = _picked(hs, 1, 5.0)
        task_syn = ops.masked_l1_l2(self.syn_depth, ps, ms)
        task_real_d = ops.masked_l1_l2(self.real_depth, pr, mr)
        task_real_i = ops.masked_l1_l2(self.real_depth_by_image, pr, self.real_hole_mask)
        self.loss_task_syn = _picked(task_syn, 0)
        self.loss_task_real_by_depth = _picked(task_real_d, 0)
        self.loss_task_real_by_image = _picked(task_real_i, 0)
        hr = ops.masked_l1_l2(self.real_depth, pr, self._a_r)       # mask_real_add_holes (:396-398)
        self.loss_holes_real = _picked(hr, 0)
        self.loss_holes_real_l2 = _picked(hr, 1, 5.0)
        terms = [(task_syn, (opt.w_syn_l1, 0.0)), (hs, (opt.w_syn_holes, 5.0 * opt.w_syn_holes)),
                 (task_real_d, (opt.w_real_l1_d, 0.0)), (task_real_i, (opt.w_real_l1_i, 0.0)), (tv_syn, tvw),
                 (norms_holes, (opt.w_syn_norm * 5, 0.0)), (tv_real, tvw), (norms_old, (0.0, opt.w_syn_norm)),
                 (tv_real_old, tvw), (tv_syn_old, tvw),                                        # :393
                 (hr, (opt.w_real_holes, 5.0 * opt.w_real_holes)),                             # :398-402
                 (norms, (opt.w_syn_norm, 0.0))]                                               # :404
        if opt.use_smooth_loss:
            self.loss_smooth = get_smooth_weight(pr, self.real_image, 3)                      # :407
            terms.append((self.loss_smooth, opt.w_smooth))
        if self._nonfinite is None and self.device.type == "cuda":
            self._nonfinite = torch.zeros(1, device=self.device, dtype=torch.int32)
        self.loss_G = ops.loss_sum(terms, opt.scale_G, self._nonfinite)                       # :417
        if back:
            self.loss_G.backward()
            ops.join_side()          # the weight-gradient stream (ops._on_side) rejoins before anything reads the gradients

    # visuals the reference overwrites after the losses (:386, :400) - computed only when read
    @property
    def mask_syn_add_holes(self):
        return (self.pred_syn_depth * self.syn_mask * self._a_s).detach()

    @property
    def mask_real_add_holes(self):
        return (self.pred_real_depth * self._a_r).detach()

    def _step_body(self):                                           # main_model.py:425-429
        if self.device.type == "cuda":
            ops.zero_pool_reset(self.device)       # one memset for every accumulator of the step
            if self.arena is not None:
                ops.prepack(self.arena.params)     # packed copies of the trainable weights: side stream, beside the frozen nets
        self.forward()
        self._backward_and_step()

    def _backward_and_step(self):
        self.set_requires_grad([self.netG_A_d, self.netI2D_features, self.netImage2Depth], False)
        self.optimizer_G.zero_grad()
        self.backward_G()
        if self.grad_sync is not None:
            self.grad_sync.finish()
        self.optimizer_G.step()

    def _frozen_body(self):
        """captured alone (pipelined mode): the frozen networks with their own accumulator pool"""
        ops.zero_pool_reset(self.device)
        self._forward_frozen()

    def _train_body(self):
        """captured alone (pipelined mode): everything that reads the trainable weights"""
        ops.zero_pool_reset(self.device)
        ops.prepack(self.arena.params)
        self._forward_train("train")
        self._backward_and_step()

    def nonfinite_steps(self):
        """How many training steps so far produced a NaN / Inf ``loss_G`` (0 = healthy).  The loss kernel counts on the
        device - also inside a replayed CUDA graph, where nothing can raise; reading the counter synchronises, so a training
        loop asks every N steps.  ``check_finite()`` raises instead."""
        return 0 if self._nonfinite is None else int(self._nonfinite.item())

    def check_finite(self):
        n = self.nonfinite_steps()
        if n:
            raise FloatingPointError(f"dsr_b200: loss_G was NaN / Inf in {n} training step(s)")

    def reset_graph(self):
        """drop the captured training / inference graphs (a new batch shape, reloaded weights, orderly shutdown)"""
        if self._graph is not None:
            self._graph.reset()
        self._graph, self._eager_steps, self._graph_keep = None, 0, None
        if getattr(self, "_pipe", None) is not None:
            torch.cuda.synchronize()
            for sl in self._pipe["slots"]:
                for g in (sl["gf"], sl["gt"]):
                    if g is not None:
                        g.reset()
            self._pipe = None
        st = getattr(self, "_tgraph", None)
        if st is not None and st.get("graph") is not None:
            st["graph"].reset()
        self._tgraph = None

    def optimize_parameters(self, iters=0, fr=1):                   # main_model.py:422-429
        """One training step.  With ``use_graph`` the first ``graph_warmup`` calls run eagerly, the next one captures
        the whole step (forward, loss stack, backward, gradient all-reduce, Adam) into a CUDA graph, and every later
        call only draws the rectangle tables on the host (reference RNG order) and replays it: ~800 kernel launches
        become one."""
        if not (self.use_graph and self.device.type == "cuda" and self.isTrain):
            return self._step_body()
        if self._pipeline_on() and getattr(self, "_pipe", None) is not None:
            return self._optimize_pipe()
        cur = torch.cuda.current_stream()
        if self._graph is None:
            # Warm-up steps and the capture run on ONE dedicated stream: the autograd engine remembers the stream every
            # node (incl. the parameters' AccumulateGrad nodes) was built on and orders the backward pass across them
            # with events - an event of an un-captured stream inside the capture is an error.
            if self._gstream is None:
                self._gstream = torch.cuda.Stream()
            gs = self._gstream
            gs.wait_stream(cur)
            with torch.cuda.stream(gs):
                if self._eager_steps < self.graph_warmup:
                    self._eager_steps += 1
                    self._step_body()
                    cur.wait_stream(gs)
                    return
                B, _, H, W = self.real_depth.shape
                for k, v in list(vars(self).items()):       # drop the previous step's autograd graph
                    if torch.is_tensor(v) and v.grad_fn is not None:
                        setattr(self, k, v.detach())
                if self._rect is None:
                    self._stage_rects(B, H, W, "train")
                self._rects_staged = True                   # capture records kernels only: the host RNG must not advance
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                from . import _lib
                l0 = _lib.LAUNCHES
                with ops.capturing():
                    with torch.cuda.graph(graph, stream=gs):
                        self._step_body()
                self._graph = graph
                self._graph_keep = ops.packed_weight_refs(self)   # the graph reads these buffers through raw pointers
                self.graph_launches = _lib.LAUNCHES - l0     # library calls recorded into the graph (= per replay)
            cur.wait_stream(gs)
        B, _, H, W = self.real_depth.shape
        self.optimizer_G.sync_hyper()
        self._stage_rects(B, H, W, "train")
        self._rects_staged = False
        self._graph.replay()
        from . import _lib
        _lib.LAUNCHES += self.graph_launches
        ops.WEIGHT_EPOCH += 1                    # eager consumers must re-pack the trainable weights

    # ------------------------------------------------------------------------------------------
    # Pipelined graph replay.  The frozen networks (a third of the step) depend on the batch only, the training part leaves
    # a B200 half idle in places (hundreds of launches that are too small for 148 SMs).  With two input slots and two
    # captured graphs per slot - `gf` = the frozen networks, `gt` = everything that reads the trainable weights -
    # ``set_input(batch i+1)`` replays gf on its own stream while gt of batch i is still running: the ordinary loop
    #     for data in loader: model.set_input(data); model.optimize_parameters()
    # overlaps them without any change, because the host runs ahead of the device.  Every batch still gets exactly one frozen
    # pass and one training pass, in order, with the same arithmetic (tests: replay == eager).  Events order the slots:
    # gf(slot) waits for the gt that last read the slot (`train_done`), gt waits for `frozen_done` of its slot.
    # ------------------------------------------------------------------------------------------
    def _pipeline_on(self):
        return (self.use_graph and self.isTrain and self.device.type == "cuda" and type(self) is MainModel
                and self.arena is not None and getattr(self.opt, "pipeline_frozen", True) and ops.CONFIG.get("pipeline_frozen", True))

    def _set_input_pipe(self, src, cams, shapes):
        P = getattr(self, "_pipe", None)
        if P is None or P["shapes"] != shapes:
            if P is not None and any(sl["gt"] is not None for sl in P["slots"]):
                raise RuntimeError("dsr_b200: the captured CUDA graphs are bound to the first batch shape; "
                                   "call reset_graph() before changing it")
            mk = lambda: dict(inp={k: torch.empty(v.shape, device=self.device, dtype=torch.float32) for k, v in src.items()},
                              cams={k: torch.empty(v.shape, device=self.device, dtype=torch.float64) for k, v in cams.items()},
                              gf=None, gt=None, outs=None, keep=None, has_frozen=False, launches=0,
                              ready=torch.cuda.Event(), frozen_done=torch.cuda.Event(), train_done=torch.cuda.Event())
            P = self._pipe = dict(shapes=shapes, slots=[mk(), mk()], turn=0, cur=0, s_frozen=torch.cuda.Stream())
            for sl in P["slots"]:
                sl["train_done"].record()
        k = P["turn"]
        P["turn"], P["cur"] = k ^ 1, k
        sl, sf = P["slots"][k], P["s_frozen"]
        # the training part that last read this slot (two steps ago) is done.  (Device-resident input tensors must be complete
        # when set_input is called: waiting for the caller's stream here would wait for the training part of the previous
        # batch, which is exactly what this stream is meant to run beside.)
        sf.wait_event(sl["train_done"])
        with torch.cuda.stream(sf):
            for name, v in src.items():
                sl["inp"][name].copy_(self._pinned(v), non_blocking=True)
            for name, v in cams.items():
                sl["cams"][name].copy_(v.pin_memory(), non_blocking=True)
            sl["ready"].record()
            sl["has_frozen"] = sl["gf"] is not None
            if sl["gf"] is not None:
                sl["gf"].replay()
                sl["frozen_done"].record()
        for name in src:
            setattr(self, name, sl["inp"][name])
        for name in cams:
            setattr(self, name, sl["cams"][name])
        if sl["outs"] is not None:
            self.__dict__.update(sl["outs"])         # result attributes now mean THIS slot's tensors
        torch.cuda.current_stream().wait_event(sl["ready"])     # eager consumers (warm-up steps, forward('test')) see the inputs

    def _optimize_pipe(self):
        from . import _lib
        P = self._pipe
        sl, sf = P["slots"][P["cur"]], P["s_frozen"]
        cur = torch.cuda.current_stream()
        B, _, H, W = self.real_depth.shape
        if sl["gt"] is None:
            if self._gstream is None:
                self._gstream = torch.cuda.Stream()
            gs = self._gstream
            gs.wait_stream(cur)
            with torch.cuda.stream(gs):
                if self._eager_steps < self.graph_warmup:
                    self._eager_steps += 1
                    self._step_body()
                    cur.wait_stream(gs)
                    sl["train_done"].record(cur)
                    return
                for k, v in list(vars(self).items()):       # drop the previous step's autograd graph
                    if torch.is_tensor(v) and v.grad_fn is not None:
                        setattr(self, k, v.detach())
                if self._rect is None:
                    self._stage_rects(B, H, W, "train")
                self._rects_staged = True                   # capture records kernels only: the host RNG must not advance
                torch.cuda.synchronize()
                l0 = _lib.LAUNCHES
                gf, gt = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                prev = ops.zero_pool_select("frozen")       # gf of one slot runs beside gt of the other: separate accumulators
                try:
                    self._frozen_body()                     # one eager pass arms the pool: a captured reset clears what the
                    torch.cuda.synchronize()                # pool has handed out BEFORE the capture (its high-water mark)
                    l0 = _lib.LAUNCHES
                    with ops.capturing():
                        with torch.cuda.graph(gf, stream=sf):
                            self._frozen_body()
                finally:
                    ops.zero_pool_select(prev)
                with ops.capturing():
                    with torch.cuda.graph(gt, stream=gs):
                        self._train_body()
                sl["gf"], sl["gt"] = gf, gt
                sl["keep"] = ops.packed_weight_refs(self)   # the graphs read these buffers through raw pointers
                sl["launches"] = _lib.LAUNCHES - l0
                sl["outs"] = {k: v for k, v in vars(self).items() if torch.is_tensor(v) or isinstance(v, LazyScalar)}
                self.graph_launches = sl["launches"]
            cur.wait_stream(gs)
        self.optimizer_G.sync_hyper()
        self._stage_rects(B, H, W, "train")
        self._rects_staged = False
        if not sl["has_frozen"]:                 # this slot's frozen graph did not exist yet when set_input ran
            sf.wait_stream(cur)
            with torch.cuda.stream(sf):
                sl["gf"].replay()
                sl["frozen_done"].record()
            sl["has_frozen"] = True
        cur.wait_event(sl["frozen_done"])
        sl["gt"].replay()
        sl["train_done"].record(cur)
        self.__dict__.update(sl["outs"])
        _lib.LAUNCHES += sl["launches"]
        ops.WEIGHT_EPOCH += 1                    # eager consumers must re-pack the trainable weights

    def calculate(self, stage="test"):                              # main_model.py:433-436
        self.forward(stage)
        self.backward_G(back=False)

    @torch.no_grad()
    def forward_test_graph(self):
        """``forward('test')`` (the inference pass of main.py:109-132) replayed as a CUDA graph: the first two calls run
        eagerly, the third captures, later calls draw the (size-zero) test-stage rectangles on the host in the reference's
        RNG order and replay ~400 launches as one.  Inputs come from ``set_input`` (persistent device buffers); the
        result tensors (``pred_real_depth`` ...) keep their addresses and are overwritten by every replay."""
        if self.device.type != "cuda" or getattr(self, "_pipe", None) is not None:
            return self.forward("test")      # (a pipelined training model alternates its input slots: no fixed graph inputs)
        B, _, H, W = self.real_depth.shape
        key = (B, H, W)
        st = getattr(self, "_tgraph", None)
        if st is None or st["key"] != key:
            st = self._tgraph = dict(key=key, graph=None, eager=0, stream=torch.cuda.Stream())
        cur = torch.cuda.current_stream()
        if st["graph"] is None:
            gs = st["stream"]
            gs.wait_stream(cur)
            with torch.cuda.stream(gs):
                if st["eager"] < 2:
                    st["eager"] += 1
                    ops.zero_pool_reset(self.device)
                    self.forward("test")
                    cur.wait_stream(gs)
                    return
                self._stage_rects(B, H, W, "test")
                self._rects_staged = True
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with ops.capturing():                    # trainable weights are re-packed INSIDE the graph (every replay sees
                    with torch.cuda.graph(graph, stream=gs):   # the current weights, whatever training did in between)
                        ops.zero_pool_reset(self.device)     # the accumulators of the pass are re-zeroed by every replay
                        self.forward("test")
                st["graph"] = graph
                st["outs"] = {k: v for k, v in vars(self).items() if torch.is_tensor(v)}
                st["keep"] = ops.packed_weight_refs(self)    # frozen nets: the graph reads these packed copies by raw pointer
            cur.wait_stream(gs)
            # capture records kernels without running them: fall through and replay, so that THIS call's outputs are
            # computed as well (the rectangle tables of this call are already staged)
        else:
            self._stage_rects(B, H, W, "test")
        self._rects_staged = False
        st["graph"].replay()
        for k, v in st["outs"].items():
            setattr(self, k, v)
