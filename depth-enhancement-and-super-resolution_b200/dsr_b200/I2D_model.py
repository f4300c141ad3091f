"""``I2DModel`` - the Image Guidance Network training step (``--model I2D``: image -> depth), drop-in for the
reference's ``models/I2D_model.py`` (citations into /root/reference).  SURVEY.md section 8f rank 2 / BASELINE
configs[0].

Same constructor contract, methods, ``loss_*`` / visual names and checkpoint names (``Image_f``, ``Task``).  The step
(I2D_model.py:161-250): ``Image_f`` ResNet (3 -> Imagef_outf) and ``Task`` U-Net (-> 1) on the syn and the real image,
masked L1 against the depth maps with mask = (depth >= -0.97), Adam over ``netTask`` ONLY (:143).

Deliberate difference: in the reference ``Image_f`` takes part in autograd although no optimizer owns it - its ``.grad``
fields accumulate forever and are never read or zeroed.  Here ``Image_f`` runs under
``no_grad``: identical losses, predictions and weight trajectory, one third of the work.  ``--use_D`` raises
(``netD_depth`` is commented out in the reference's constructor, I2D_model.py:120-122, so the flag cannot run there).
"""
import itertools

import torch

from . import networks, ops
from .base_model import BaseModel, GraphStepMixin
from .main_model import ArenaAdam, ParamArena


class I2DModel(GraphStepMixin, BaseModel):
    @staticmethod
    def modify_commandline_options(parser, is_train=True):         # I2D_model.py:68-76
        parser.set_defaults(no_dropout=True)
        if is_train:
            parser.add_argument("--lambda_A", type=float, default=10.0)
            parser.add_argument("--lambda_B", type=float, default=10.0)
            parser.add_argument("--lambda_identity", type=float, default=0.5)
        return parser

    def __init__(self, opt):                                        # I2D_model.py:78-148
        BaseModel.__init__(self, opt)
        if getattr(opt, "use_D", False):
            raise NotImplementedError("dsr_b200: --use_D (netD_depth is never constructed in the reference's I2DModel)")
        self.loss_names = ["task_syn", "task_real"]
        if opt.norm_loss:
            self.loss_names += ["syn_norms"]
        visual_names_A = ["syn_image", "syn_depth", "pred_syn_depth"]
        visual_names_B = ["real_image", "real_depth", "pred_real_depth"]
        self.model_names = ["Image_f", "Task"]
        if opt.norm_loss:
            visual_names_A += ["norm_syn", "norm_syn_pred"]
            visual_names_B += ["norm_real", "norm_real_pred"]
        self.visual_names = visual_names_A + visual_names_B
        self.netImage_f = networks.define_G(3, opt.Imagef_outf, opt.Imagef_basef, opt.Imagef_type, opt.norm,
                                            not opt.no_dropout, opt.init_type, opt.init_gain, self.gpu_ids,
                                            opt.replace_transpose, n_down=opt.Imagef_ndown)
        self.netTask = networks.define_G(opt.Imagef_outf, 1, opt.Task_basef, opt.Task_type, opt.norm,
                                         not opt.no_dropout, opt.init_type, opt.init_gain, self.gpu_ids,
                                         opt.replace_transpose, n_down=opt.Task_ndown)
        self.loss_L1_syn = 0
        self.loss_L1_real = 0
        self.loss_D_depth = 0
        self.arena = None
        self.grad_sync = None
        self._in = None
        self._graph_init(opt)        # CUDA-graph replay of the whole step (persistent inputs, device-side Adam state)
        if self.isTrain:
            if self.gpu_ids:
                self.arena = ParamArena([self._unwrap(self.netTask)], self.device)
                self.optimizer_G = ArenaAdam(self.arena, opt.lr)
            else:
                self.optimizer_G = torch.optim.Adam(itertools.chain(self.netTask.parameters()), lr=opt.lr)
            self.optimizers.append(self.optimizer_G)

    def set_input(self, input):                                     # I2D_model.py:152-161
        AtoB = self.opt.direction == "AtoB"
        src = dict(syn_image=input["A_i" if AtoB else "B_i"], real_image=input["B_i" if AtoB else "A_i"],
                   syn_depth=input["A_d" if AtoB else "B_d"], real_depth=input["B_d" if AtoB else "A_d"])
        self.image_paths = input["A_paths" if AtoB else "B_paths"]
        self.A_paths, self.B_paths = input.get("A_paths"), input.get("B_paths")
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

    def forward(self, stage="train"):                               # I2D_model.py:163-183
        """``stage``: the reference's ``--save_all`` branch tests an undefined name ``stage`` (I2D_model.py:171 - a NameError as
        soon as the flag is set); here it is a parameter with the meaning it has in MainModel.forward (main_model.py:321)."""
        B = self.syn_image.shape[0]
        images = torch.cat([self.syn_image, self.real_image], 0)
        with torch.no_grad():                                       # see the module docstring
            features = self.netImage_f(images)
        self.features_syn, self.features_real = features[:B], features[B:]
        pred = self.netTask(features)
        self.pred_syn_depth, self.pred_real_depth = pred[:B], pred[B:]
        if getattr(self.opt, "save_all", False) and stage == "test":               # I2D_model.py:171-182
            from . import io
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("dsr_b200: --save_all writes files: it cannot run inside a captured CUDA graph")
            # uint16(clip((pred + 1) / 2, 0, 1) * 5100) of rows [16, H - 16) into f'{save_image_folder}{basename(B_path)}.png'
            self.saved_files = io.save_predictions(self.pred_real_depth, self.B_paths, self.opt.save_image_folder, 16)

    def backward_G(self, back=True):                                # I2D_model.py:212-235
        opt = self.opt
        if opt.norm_loss:                                           # computed, reported, NOT part of loss_G (:217,:230)
            with torch.no_grad():
                self.norm_syn = ops.normals_old(self.syn_depth, 1.0)
                self.norm_syn_pred = ops.normals_old(self.pred_syn_depth.detach(), 1.0)
                self.norm_real = ops.normals_old(self.real_depth, 1.0)
                self.norm_real_pred = ops.normals_old(self.pred_real_depth.detach(), 1.0)
                one = torch.ones((self.norm_syn.shape[0], 1) + tuple(self.norm_syn.shape[2:]), device=self.device)
                self.loss_syn_norms = ops.masked_l1_l2(self.norm_syn, self.norm_syn_pred, one)[0]
        mask_syn = ops.below_mask(self.syn_depth, -0.97)
        self.loss_task_syn = ops.masked_l1_l2(self.syn_depth, self.pred_syn_depth, mask_syn)[0]
        mask_real = ops.below_mask(self.real_depth, -0.97)
        self.loss_task_real = ops.masked_l1_l2(self.real_depth, self.pred_real_depth, mask_real)[0]
        self.loss_G = self.loss_task_syn * opt.w_syn_l1 + self.loss_task_real * opt.w_real_l1
        self.loss_G = self.loss_G * opt.scale_G
        if back:
            self.loss_G.backward()
            ops.join_side()

    def optimize_parameters(self, iters=0, fr=700):                 # I2D_model.py:237-250
        self._graph_optimize()

    def _step_body(self):
        if self.device.type == "cuda":
            ops.zero_pool_reset(self.device)
        self.forward()
        self.optimizer_G.zero_grad()
        self.backward_G()
        if self.grad_sync is not None:
            self.grad_sync.finish()
        self.optimizer_G.step()

    def calculate(self, stage="train"):                             # I2D_model.py:253-257
        self.forward()
        self.backward_G(back=False)
