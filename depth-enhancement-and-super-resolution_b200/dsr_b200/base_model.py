"""Model facade base class - drop-in for the reference's ``models/base_model.py``.

Keeps the public surface the caller (``main.py:54-132``) uses: ``setup``, ``eval`` / ``_train``,
``update_learning_rate``, ``get_current_losses`` / ``get_current_visuals``, ``save_networks`` /
``load_networks`` (one ``{tag}_net_{name}.pth`` per network = ``OrderedDict[str -> fp32 CPU tensor]``,
keys without ``module.``; base_model.py:150-166, :182-237), ``set_requires_grad``.
Differences, all deliberate: one process drives ONE device (``gpu_ids[0]``); a missing or unreadable
checkpoint is reported (and raises with ``strict=True``) instead of being silently swallowed
(base_model.py:235-237); saving does not bounce the whole net D2H and back (base_model.py:163-164).
"""
import os
from abc import ABC, abstractmethod
from collections import OrderedDict

import torch

from . import networks


class BaseModel(ABC):
    def __init__(self, opt):                                        # base_model.py:18-44
        self.opt = opt
        self.gpu_ids = opt.gpu_ids
        self.isTrain = opt.isTrain
        # gpu_ids == [] builds the parameter containers on the host (checkpoint conversion, key
        # inspection); there is no CPU compute path - forward() then raises from the first op.
        self.device = torch.device("cuda:{}".format(self.gpu_ids[0])) if self.gpu_ids else torch.device("cpu")
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        self.loss_names = []
        self.model_names = []
        self.visual_names = []
        self.optimizers = []
        self.image_paths = []
        self.metric = 0

    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    @abstractmethod
    def set_input(self, input):
        pass

    @abstractmethod
    def forward(self):
        pass

    @abstractmethod
    def optimize_parameters(self):
        pass

    def setup(self, opt):                                           # base_model.py:78-89
        if self.isTrain:
            self.schedulers = [networks.get_scheduler(optimizer, opt) for optimizer in self.optimizers]
        if not self.isTrain or opt.continue_train:
            load_suffix = "iter_%d" % opt.load_iter if opt.load_iter > 0 else opt.epoch
            self.load_networks(load_suffix)
        self.print_networks(opt.verbose)

    def _nets(self):
        for name in self.model_names:
            if isinstance(name, str):
                yield name, getattr(self, "net" + name)

    def eval(self):                                                 # base_model.py:91-96
        for _, net in self._nets():
            net.eval()

    def _train(self):                                               # base_model.py:98-103
        for _, net in self._nets():
            net.train()

    def test(self):
        with torch.no_grad():
            self.forward()
            self.compute_visuals()

    def compute_visuals(self):
        pass

    def get_image_paths(self):
        return self.image_paths

    def update_learning_rate(self):                                 # base_model.py:121-130
        for scheduler in self.schedulers:
            if self.opt.lr_policy == "plateau":
                scheduler.step(self.metric)
            else:
                scheduler.step()
        lr = self.optimizers[0].param_groups[0]["lr"]
        print("learning rate = %.7f" % lr)

    def get_current_visuals(self):                                  # base_model.py:132-138
        visual_ret = OrderedDict()
        for name in self.visual_names:
            if isinstance(name, str):
                visual_ret[name] = getattr(self, name)
        return visual_ret

    def get_current_losses(self):                                   # base_model.py:140-148
        errors_ret = OrderedDict()
        for name in self.loss_names:
            if isinstance(name, str):
                errors_ret[name] = float(getattr(self, "loss_" + name))
        return errors_ret

    def loss_async(self, name="loss_G"):
        """Enqueue the device->host copy of a loss scalar into pinned memory and return a zero-argument callable that waits
        for THAT copy and returns the float: a training loop reads step i's loss while step i+1 is already running instead
        of draining the device every step (``float(model.loss_G)`` does)."""
        t = getattr(self, name)
        if not torch.is_tensor(t) or not t.is_cuda:
            return lambda: float(t)
        ring = self.__dict__.setdefault("_loss_ring", dict(turn=0, slots=[(torch.zeros(1).pin_memory(), torch.cuda.Event()) for _ in range(4)]))
        host, ev = ring["slots"][ring["turn"] % 4]
        ring["turn"] += 1
        host.copy_(t.detach().reshape(1), non_blocking=True)
        ev.record()

        def read():
            ev.synchronize()
            return float(host[0])
        return read

    @staticmethod
    def _unwrap(net):
        return net.module if hasattr(net, "module") else net

    def save_networks(self, epoch):                                 # base_model.py:150-166
        os.makedirs(self.save_dir, exist_ok=True)
        for name, net in self._nets():
            save_path = os.path.join(self.save_dir, "%s_net_%s.pth" % (epoch, name))
            sd = OrderedDict((k, v.detach().to("cpu", torch.float32).contiguous())
                             for k, v in self._unwrap(net).state_dict().items())
            torch.save(sd, save_path)

    def load_networks(self, epoch, strict=False):                   # base_model.py:182-237
        if hasattr(self, "reset_graph"):
            self.reset_graph()      # a captured graph reads packed copies of the OLD weights through raw pointers: it goes
        from . import ops
        ops.WEIGHT_EPOCH += 1
        for name, net in self._nets():
            load_filename = "%s_net_%s.pth" % (epoch, name)
            load_path = os.path.join(self.save_dir, load_filename)
            net = self._unwrap(net)
            print("loading the model from %s" % load_path)
            try:
                state_dict = torch.load(load_path, map_location="cpu")
                if load_filename == "latest_net_G_A_d.pth" and ("netG_B" in list(state_dict.keys())):
                    state_dict = state_dict["netG_B"]               # base_model.py:204-205
                if hasattr(state_dict, "_metadata"):
                    del state_dict._metadata
                net_dict = net.state_dict()
                keep = {k: v for k, v in state_dict.items() if (k in net_dict) and (v.shape == net_dict[k].shape)}
                skipped = [k for k in net_dict if k not in keep]
                net_dict.update(keep)                               # shape-filtered partial load (:225)
                net.load_state_dict(net_dict)
                if skipped:
                    print("  [dsr_b200] %s: %d tensors kept their current values (missing / shape mismatch): %s"
                          % (load_filename, len(skipped), ", ".join(skipped[:4]) + (" ..." if len(skipped) > 4 else "")))
            except Exception as e:
                if strict:
                    raise
                print("  [dsr_b200] WARNING: could not load %s (%s: %s) - weights left as they are"
                      % (load_path, type(e).__name__, e))

    def print_networks(self, verbose):                              # base_model.py:239-255
        print("---------- Networks initialized -------------")
        for name, net in self._nets():
            num_params = sum(p.numel() for p in net.parameters())
            if verbose:
                print(net)
            print("[Network %s] Total number of parameters : %.3f M" % (name, num_params / 1e6))
        print("-----------------------------------------------")

    def set_requires_grad(self, nets, requires_grad=False):        # base_model.py:257-268
        if not isinstance(nets, list):
            nets = [nets]
        for net in nets:
            if net is not None:
                for param in net.parameters():
                    param.requires_grad = requires_grad


class GraphStepMixin:
    """CUDA-graph replay of a whole training step for models whose step has no host-side randomness: the first
    ``graph_warmup`` calls run ``_step_body`` eagerly, the next one captures it, later calls replay it.  Needs persistent
    input buffers (``set_input`` copies into fixed device tensors) and device-side optimizer state (``ArenaAdam``).
    (``MainModel`` has its own variant that also stages the rectangle tables drawn on the host.)"""
    use_graph = False
    graph_warmup = 2

    def _graph_init(self, opt):
        self.use_graph = bool(getattr(opt, "cuda_graph", False))
        self._graph, self._gstream, self._eager_steps, self.graph_launches = None, None, 0, 0

    def reset_graph(self):
        if getattr(self, "_graph", None) is not None:
            self._graph.reset()
        self._graph, self._eager_steps, self._graph_keep = None, 0, None

    def _graph_optimize(self):
        from . import _lib, ops
        if not (self.use_graph and self.device.type == "cuda" and self.isTrain):
            return self._step_body()
        cur = torch.cuda.current_stream()
        if self._graph is None:
            if self._gstream is None:
                self._gstream = torch.cuda.Stream()
            gs = self._gstream
            gs.wait_stream(cur)
            with torch.cuda.stream(gs):     # warm-up and capture on ONE stream (autograd remembers each node's stream)
                if self._eager_steps < self.graph_warmup:
                    self._eager_steps += 1
                    self._step_body()
                    cur.wait_stream(gs)
                    return
                for k, v in list(vars(self).items()):
                    if torch.is_tensor(v) and v.grad_fn is not None:
                        setattr(self, k, v.detach())
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                l0 = _lib.LAUNCHES
                with ops.capturing():
                    with torch.cuda.graph(graph, stream=gs):
                        self._step_body()
                self._graph, self.graph_launches = graph, _lib.LAUNCHES - l0
                self._graph_keep = ops.packed_weight_refs(self)     # the graph reads these buffers through raw pointers
            cur.wait_stream(gs)
        for o in self.optimizers:
            if hasattr(o, "sync_hyper"):
                o.sync_hyper()
        self._graph.replay()
        _lib.LAUNCHES += self.graph_launches
        ops.WEIGHT_EPOCH += 1
