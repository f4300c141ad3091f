"""The UNMODIFIED reference (byte-compiled into ``oracle/_ref`` by oracle/build_ref.py) driven on the CPU through its
own public API: ``TrainOptions().parse()`` with the README flag set -> ``MainModel(opt)`` -> ``set_input`` ->
``optimize_parameters`` (models/main_model.py:179-201, :422-429; recipe of SURVEY.md Appendix D).

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by ``bench.py --impl reference`` and the ``cpu_baseline`` leg
(``kind == "reference"``)."""
import os
import sys
import types

from . import build_ref

FLAGS_MAIN = ["--w_syn_l1", "15", "--w_real_l1_d", "40", "--norm_loss", "--w_syn_norm", "2", "--use_smooth_loss", "--w_smooth", "1",
              "--w_syn_holes", "800", "--w_real_holes", "1600", "--lr", "0.0001"]                       # README.md:70
FLAGS_SR = ["--w_syn_l1", "15", "--w_real_l1_d", "90", "--norm_loss", "--w_syn_norm", "3", "--use_smooth_loss", "--w_smooth", "1",
            "--w_syn_holes", "1600", "--w_real_holes", "1600", "--lr", "0.00002", "--SR"]                # README.md:86


def available():
    return build_ref.available()


class _RefFinder:
    """meta-path finder for the byte-compiled reference tree: <pkg>/<module>.pyb, packages as <pkg>/__init__.pyb"""

    @staticmethod
    def find_spec(fullname, path=None, target=None):
        import importlib.machinery
        import importlib.util
        if fullname.split(".")[0] not in build_ref.PACKAGES:
            return None
        base = os.path.join(build_ref.OUT, *fullname.split("."))
        if os.path.exists(os.path.join(base, "__init__.pyb")):
            f = os.path.join(base, "__init__.pyb")
            return importlib.util.spec_from_file_location(fullname, f, loader=importlib.machinery.SourcelessFileLoader(fullname, f),
                                                          submodule_search_locations=[base])
        if os.path.exists(base + ".pyb"):
            return importlib.util.spec_from_file_location(fullname, base + ".pyb",
                                                          loader=importlib.machinery.SourcelessFileLoader(fullname, base + ".pyb"))
        return None


def _import_reference():
    if not any(f is _RefFinder for f in sys.meta_path):
        sys.meta_path.insert(0, _RefFinder)
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))      # only used by the --save_all branch (main_model.py:11)


def make_model(B, H, W, sr=False, seed=0):
    """the reference's MainModel / MainSRModel on --gpu_ids -1 with the seeded reference initialisation"""
    import numpy as np
    import torch
    _import_reference()
    argv = sys.argv
    sys.argv = ["main.py", "--gpu_ids", "-1", "--image_and_depth", "--custom_pathes", "--use_image_for_trans", "--use_masked",
                "--use_scannet", "--model", "main_network_best", "--batch_size", str(B), "--name", "ref_live", "--do_train",
                "--model_type", "main", "--checkpoints_dir", "/tmp/dsr_ref_live", "--crop_size_h", str(H), "--crop_size_w", str(W)] \
        + (FLAGS_SR if sr else FLAGS_MAIN)
    try:
        from options.train_options import TrainOptions
        opt = TrainOptions().parse()
    finally:
        sys.argv = argv
    torch.manual_seed(seed)
    np.random.seed(seed)
    if sr:
        from models import translation_network
        orig = translation_network.init_net                             # main_sr_model.py:166 hard-codes gpu_ids=[0,1,2,3]
        translation_network.init_net = lambda net, init_type="normal", init_gain="relu", gpu_ids=[], param=None: \
            orig(net, init_type, init_gain, [], param)
        from models.main_sr_model import MainSRModel as Model
    else:
        from models.main_model import MainModel as Model
    model = Model(opt)
    model.setup(opt)
    model._train()
    return model


def time_steps(B, H, W, batch, steps, warmup, sr=False):
    """-> seconds per set_input + optimize_parameters of the live reference on this host's cores"""
    import time
    import numpy as np
    model = make_model(B, H, W, sr=sr)
    np.random.seed(0)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.set_input(batch)
        model.optimize_parameters(i, 1)
        ts.append(time.perf_counter() - t0)
    ts = ts[warmup:]
    return sum(ts) / len(ts), float(model.loss_G)
