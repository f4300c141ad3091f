"""torch.autograd.Function wrappers over the C-ABI library (one Function per op of the hot path).

PyTorch is used for device memory, streams and the autograd graph only; every kernel that runs is
ours (csrc/*.cu through include/dsr_b200.h).  Activations are logical NCHW tensors whose memory is
NHWC (torch channels_last); loss-stack tensors (depth, normals, masks) are plain NCHW planes exactly
as in the reference.  There is no CPU path: a non-CUDA tensor raises.
"""
import torch
from torch.autograd import Function

from . import _lib

PAD_ZERO, PAD_REFLECT, PAD_REPLICATE = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
PAD_MODES = {"zeros": PAD_ZERO, "zero": PAD_ZERO, "reflect": PAD_REFLECT, "replicate": PAD_REPLICATE}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t, dtype=torch.float32):
    """device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("dsr_b200 ops run on CUDA tensors only (no CPU fallback); got a %s tensor" % t.device)
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError("internal error: non-contiguous tensor handed to the C-ABI")
    return t.data_ptr()


def _call(name, *args):
    _lib.call(name, *args, _stream())
    _lib.PROFILE_META = None


# ------------------------------------------------------------------------------------------------
# layout helpers (raw, no autograd)
# ------------------------------------------------------------------------------------------------
def nhwc(x):
    """logical NCHW fp32 tensor -> contiguous (N,H,W,C) tensor sharing memory when possible."""
    if x.dim() != 4:
        raise ValueError("expected a 4-D NCHW tensor")
    if not x.is_cuda:
        raise RuntimeError("dsr_b200 ops run on CUDA tensors only (no CPU fallback); got a %s tensor" % x.device)
    if x.dtype != torch.float32:
        raise TypeError(f"expected float32, got {x.dtype}")
    xp = x.permute(0, 2, 3, 1)
    if xp.is_contiguous():
        return xp
    N, C, H, W = x.shape
    xc = x if x.is_contiguous() else x.contiguous()
    y = torch.empty((N, H, W, C), device=x.device, dtype=torch.float32)
    _call("dsr_nchw_to_nhwc", _p(xc), _p(y), N, C, H * W)
    return y


def nchw(y):
    """contiguous (N,H,W,C) tensor -> logical NCHW view (channels_last strides)."""
    return y.permute(0, 3, 1, 2)


def planes(x):
    """logical NCHW tensor -> NCHW-contiguous tensor (the reference's layout for the loss stack)."""
    if x.is_contiguous():
        return x
    N, C, H, W = x.shape
    xp = x.permute(0, 2, 3, 1)
    if xp.is_contiguous() and x.is_cuda and x.dtype == torch.float32:
        y = torch.empty((N, C, H, W), device=x.device, dtype=torch.float32)
        _call("dsr_nhwc_to_nchw", _p(xp), _p(y), N, C, H * W)
        return y
    return x.contiguous()


class _ZeroPool:
    """Pre-zeroed fp64 scratch for the step's accumulators (norm statistics, IN-backward sums, bias-gradient sums,
    loss reductions): ~200 tiny buffers per step come out of one arena that is cleared by ONE memset per step
    (``zero_pool_reset``, called by the model at the top of the step) instead of ~200 fill kernels.
    Pools are named: two captured graphs that may run at the same time (the frozen networks of the next batch beside the
    training part of the current one, ``MainModel`` pipelining) must not share accumulators, so each selects its own
    (``zero_pool_select``)."""
    CAP = 1 << 22            # doubles (32 MB)

    def __init__(self):
        self.buf, self.used, self.high = {}, {}, {}
        self.name = "step"

    def _key(self, device):
        return (device.type, device.index, self.name)

    def take(self, n, device):
        key = self._key(device)
        if key not in self.buf:
            return torch.zeros(n, device=device, dtype=torch.float64)       # pool not armed on this device
        n8 = (n + 7) // 8 * 8
        u = self.used[key]
        if u + n8 > self.CAP:
            return torch.zeros(n, device=device, dtype=torch.float64)
        self.used[key] = u + n8
        self.high[key] = max(self.high.get(key, 0), u + n8)
        return self.buf[key][u:u + n]

    def reset(self, device):
        key = self._key(device)
        if key not in self.buf:
            self.buf[key] = torch.zeros(self.CAP, device=device, dtype=torch.float64)
            self.used[key] = 0
            self.high[key] = 0
            return
        # everything the pool has EVER handed out on this device is cleared: a reset captured into a CUDA graph has a fixed
        # length, and the pass that preceded the capture may have used less of the pool than the captured step does
        u = self.high.get(key, 0)
        if u:
            self.buf[key][:u].zero_()
        self.used[key] = 0


_ZERO_POOL = _ZeroPool()


def zero_pool_reset(device):
    """call once at the top of a step: re-zeroes what the previous step handed out"""
    _ZERO_POOL.reset(torch.device(device))
    _PREMADE.clear()


def zero_pool_select(name="step"):
    """-> the previous name.  Accumulators taken from now on come out of the pool `name` (see _ZeroPool)."""
    prev, _ZERO_POOL.name = _ZERO_POOL.name, name
    return prev


def _zeros_f64(n, device):
    return _ZERO_POOL.take(n, torch.device(device))


# ------------------------------------------------------------------------------------------------
# convolutions
# ------------------------------------------------------------------------------------------------
def _pack(weight, kdim):
    D0, D1, R, S = weight.shape
    w = weight.detach()
    w = w if w.is_contiguous() else w.contiguous()
    out = torch.empty(D0 * D1 * R * S, device=w.device, dtype=torch.float32)
    _call("dsr_pack_weight", _p(w), D0, D1, R, S, kdim, _p(out))
    return out


# Parameters registered here (data_ptr -> fp32 gradient view inside the gradient arena) get their
# gradients ACCUMULATED in place by the backward kernels instead of being returned to autograd
# (no AccumulateGrad add kernels, one flat buffer for Adam and for the NCCL buckets).  The arena
# owner zeroes it once per step.  GRAD_READY, if set, is called with the parameter's data_ptr as
# soon as its gradient kernel is enqueued - the hook the bucketed all-reduce overlaps on.
DIRECT_GRADS = {}
GRAD_READY = None


def _grad_ready(param):
    if GRAD_READY is not None:
        GRAD_READY(param.data_ptr())


# Weight / bias gradients of arena parameters are off the critical path of the backward pass (nothing reads them before the
# all-reduce / Adam), while the data-gradient chain is a long sequence of dependent launches, many of them far too small to
# fill 148 SMs.  They therefore run on a SECOND stream: forked behind the operand preparation of dY on the main stream,
# joined by `join_side()` right after `loss.backward()`.  Tensors allocated on the main stream and read on the side stream
# are kept alive until the join (`_keep`), so the caching allocator cannot hand their memory to a later main-stream
# allocation while the side stream still reads it.  Also legal inside a CUDA-graph capture (event fork / join).
_SIDE = {"stream": {}, "forked": False, "keep": []}


def _use_side(*params):
    """side stream only for gradients accumulated in place in the arena (a gradient RETURNED to autograd must be ready on
    the main stream)"""
    return CONFIG["side_wgrad"] and all(p is None or p.data_ptr() in DIRECT_GRADS for p in params)


class _on_side:
    def __init__(self, enabled, device):
        self.enabled, self.device, self.ctx = enabled, device, None

    def __enter__(self):
        if not self.enabled:
            return self
        key = self.device.index
        side = _SIDE["stream"].get(key)
        if side is None:
            side = _SIDE["stream"][key] = torch.cuda.Stream(self.device)
        _SIDE["main"] = torch.cuda.current_stream()
        side.wait_stream(_SIDE["main"])
        self.ctx = torch.cuda.stream(side)
        self.ctx.__enter__()
        _SIDE["forked"] = True
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _keep(*tensors):
    if _SIDE["forked"]:
        _SIDE["keep"].extend(t for t in tensors if t is not None)


def order_after_gradient_writers():
    """the current stream waits for every stream that may have written gradients (called before a bucket's all-reduce is
    enqueued: a bucket can hold gradients written on the main AND on the weight-gradient stream)"""
    if not _SIDE["forked"]:
        return
    cur = torch.cuda.current_stream()
    for s in list(_SIDE["stream"].values()) + [_SIDE.get("main")]:
        if s is not None and s != cur:
            cur.wait_stream(s)


def join_side():
    """main stream waits for the weight-gradient stream (call after loss.backward(), before anything reads the gradients)"""
    if _SIDE["forked"]:
        cur = torch.cuda.current_stream()
        for side in _SIDE["stream"].values():
            cur.wait_stream(side)
        _SIDE["forked"] = False
        _SIDE["prepack_pending"] = False
    _SIDE["keep"] = []


def _bias_grad(dy_nhwc, Co, bias, gP=None):
    """bias gradient = per-channel sum of dY: taken for free by the operand prep of dY when there was one (gP.csum),
    by one channel_sums pass otherwise"""
    reps = 0
    if gP is not None and gP.csum is not None:
        sums, stride, reps = gP.csum, 1, gP.csum.numel() // Co        # [reps][Co] replica rows (dsr_tc_prep's csum_reps)
    else:
        sums, stride = _zeros_f64(Co * 2, dy_nhwc.device), 2
        _call("dsr_channel_sums", _p(dy_nhwc), 1, dy_nhwc.numel() // Co, Co, _p(sums, torch.float64))
    tgt = DIRECT_GRADS.get(bias.data_ptr())
    gb = tgt if tgt is not None else torch.empty(Co, device=dy_nhwc.device, dtype=torch.float32)
    if reps > 1:
        _call("dsr_sum_reps_f64_f32", _p(sums, torch.float64), reps, Co, _p(gb), int(tgt is not None))
    else:
        _call("dsr_cvt_f64_f32", _p(sums, torch.float64), stride, _p(gb), Co, 1.0, int(tgt is not None))
    if tgt is not None:
        _grad_ready(bias)
        return None
    return gb


def _weight_grad(dwk, weight, kdim):
    D0, D1, R, S = weight.shape
    tgt = DIRECT_GRADS.get(weight.data_ptr())
    if tgt is not None:
        _call("dsr_unpack_weight", _p(dwk), D0, D1, R, S, kdim, _p(tgt), 1)
        _grad_ready(weight)
        return None
    gw = torch.empty(weight.shape, device=dwk.device, dtype=torch.float32)
    _call("dsr_unpack_weight", _p(dwk), D0, D1, R, S, kdim, _p(gw), 0)
    return gw


# ------------------------------------------------------------------------------------------------
# tcgen05 engine (csrc/conv_tc.cu): operand preparation + weight packing + implicit GEMM
# ------------------------------------------------------------------------------------------------
CONFIG = {
    "s2_border": True,   # replicate / reflect stride-2 convs: tcgen05 data gradient + CUDA-core border term
    "engine": "tc",      # "tc": tcgen05 implicit GEMM wherever the layer shape allows; "simt": fp32 CUDA cores only
    "passes": 3,         # 1 = single 16-bit pass, 2 = activations hi+lo, 3 = activations and weights hi+lo (parity mode)
    "dtype": "f16",      # forward operand format: "f16" (11-bit significand, hi+lo = 22 bits) or "bf16" (8 / 16 bits: measured
                         # worst per-tensor gradient cosine 0.9989 < the 0.999 gate, so not the default)
    "bwd_dtype": "bf16", # backward (gradient) operand format: gradients span too many octaves for unscaled f16
    "tc_backward": True, # run dgrad / wgrad on the tcgen05 path as well
    "wgrad_passes": 1,   # weight-gradient GEMM: 1 = one 16-bit pass, 2 = dY hi+lo, 3 = dY and x hi+lo.  Measured
                         # (scripts/precision_probe.py, golden step): the worst per-tensor gradient cosine is 0.999276 / 0.999304 /
                         # 0.999280 for 1 / 2 / 3 passes - it is set by the dY that reaches the layer, not by the rounding of the
                         # weight-gradient operands (a sum over >= 32 K pixels averages 2^-9 operand noise away), so one pass
    "split_k": -1,       # -1 = automatic split-K for tiny-M layers
    "tc_halo": True,     # second-generation GEMM (halo-resident A patches, persistent CTAs) wherever it applies
    "tc_cm": True,       # third-generation channel-major GEMM for Cout >= 128 layers
    "reuse_fwd_operand": True,  # wgrad reads the forward pass's arranged operand (when formats match) instead of re-preparing x
    "out1": True,        # one-output-channel heads on CUDA cores (bandwidth-bound reductions) instead of N = 16 MMAs
    "tc_first_layers": True,
    "tc_compact_first": True,   # ... reading the 8-pixel K blocks from a compact 8-channel operand (no 8x expansion)   # 7x7 first layers (1..8 input channels) on the tensor path (8-pixel K blocks)
    "halo_min_tiles": 90,    # fewest output tiles for the persistent halo kernels (below: first generation with split-K).  Was 120;
                             # measured r4j: the six 96-tile U-Net layers (64 -> 128 at 6 x 64^2, 128 -> 256 at 12 x 32^2, 256 -> 512 at
                             # 12 x 16^2) run 0.354 -> 0.235 ms on the channel-major kernel although a third of the SMs idle
    "fwd_passes": 0,     # forward GEMMs of the trainable nets: 0 = `passes`, 2 = activations hi+lo x weights hi only
    "big_hw": 1024,      # layers with >= this many output pixels per sample may run fewer MMA passes (0 = off):
    "big_fwd_passes": 0, #   forward GEMMs (0 = no override).  Measured on the golden step (scripts/precision_probe.py, r74):
                         #   2-pass forwards at >= 32^2 / 64^2 / 128^2 give worst gradient cosine 0.972 / 0.953 / 0.973 - fail
    "big_bwd_passes": 1, #   data-gradient GEMMs: ONE bf16 pass.  Measured: worst per-tensor cosine 0.999310 with 1 pass at
                         #   >= 32^2 (0.999263 with 1 pass everywhere) against 0.999277 with 3 - like the weight gradient, the
                         #   gate margin is set by the dY that reaches the layer, not by the operand rounding.  Layers below
                         #   32^2 (latency-bound, no time to win) keep `passes`
    "frozen_passes": 0,  # forward GEMMs of the frozen nets (run under no_grad): 0 = as fwd_passes / passes
    "prepack_in_use_order": True,   # ... in the order the step first uses them (not in the arena's gradient-ready order)
    "prepack": True,     # packed 16-bit copies of the trainable weights are re-made on the side stream at the top of the step
    "fwd_bf16_copy": True,   # forward operand prep of trainable layers also writes the bf16 copy the weight gradient reads
    "fused_cat": True,   # Conv2d over a LazyCat: the operand preparation reads the parts (no materialised concatenation)
    "side_wgrad": True,  # weight / bias gradients of arena parameters on a second stream beside the data-gradient chain
    "fork_frozen": True, # MainModel.forward: G_A_d on a second stream beside I2D_features -> Image2Depth
    "wgrad_kernel": 2,   # 2 = csrc/wgrad_tc2.cu (8 column blocks per CTA, single pass); 1 = first-generation kernel (conv_tc.cu)
    "csum_reps": 8,      # replica rows of the bias-gradient sums dsr_tc_prep takes on the way (fp64 atomics spread over 8 addresses
                         # per channel, which lets the launch keep the wide grid of the sum-free form); 1 = one row, narrow grid
    "fuse_bwd_prep": True,   # InstanceNorm backward also writes the dY operand of the convolution in front of the norm layer
                             # (dsr_tc_prep_in_bwd: one pass instead of dsr_in_bwd_apply + dsr_tc_prep), from the second step on:
                             # the operand a convolution's backward asks for is recorded on its weight the first time
    "fuse_skip_grad": True,  # residual blocks: the skip connection's gradient joins the first convolution's padding adjoint
                             # (dsr_pad2d_bwd_pitch_add) instead of autograd's separate add kernel
    "fuse_norm_prep": True,  # the closing norm (+ skip add) of a residual block also writes the next conv's operand (dsr_tc_prep_norm_res)
    "fold_finalize": True,   # dsr_norm_finalize folded into its first consumer (dsr_tc_prep_fin / dsr_norm_apply_fwd_fin)
    "wgrad_slabs": False,  # True: weight-gradient K splits store their own slabs (dsr_tc_wgrad2p) and the unpack sums them in a
                           # fixed order: bit-reproducible weight gradients, no atomics.  Measured on B200 (r2z): +0.35 ms per step
                           # (the unpack reads up to 48 slabs: 0.50 -> 1.43 ms; the GEMMs themselves gain 3 %), so off by default
    "dgrad_quad": True,  # ... FOUR adjacent pixels per row for Cin = 32: 128 GEMM rows = one M tile of the channel-major kernel
    "dgrad_pair": True,  # stride-1 data gradients with <= 32 output channels: two adjacent pixels per GEMM row (MMA N = 64, not 32)
}
WEIGHT_EPOCH = 0         # bumped by the optimizer: invalidates packed copies of trainable weights
_LAYOUT_NORMAL, _LAYOUT_PAIR, _LAYOUT_S2D = 0, 1, 2
_W_CONV, _W_CONV_PAIR, _W_CONV_S2D, _W_CONVT_PH, _W_CONV_DGRAD, _W_CONV_DGRAD_PAIR = 0, 1, 2, 3, 4, 5


_FWD = {"trainable": False, "hw": 0}     # set by conv2d / conv_transpose2d / cat_conv2d: is the running layer on the autograd tape?


def _passes(dtype):
    """MMA passes of a forward GEMM (dtype None = the forward operand format) or of a data-gradient GEMM"""
    if CONFIG["big_hw"] and _FWD["hw"] >= CONFIG["big_hw"]:
        n = CONFIG["big_fwd_passes"] if dtype is None else CONFIG["big_bwd_passes"]
        if n:
            return min(n, CONFIG["passes"])
    if dtype is None and not _FWD["trainable"] and CONFIG["frozen_passes"]:
        return min(CONFIG["frozen_passes"], CONFIG["passes"])       # frozen nets run under no_grad (main_model.py:426)
    if dtype is None and _FWD["trainable"] and CONFIG["fwd_passes"]:
        return min(CONFIG["fwd_passes"], CONFIG["passes"])
    return CONFIG["passes"]


_CAPTURE = {"id": 0, "active": False}


class capturing:
    """``with ops.capturing():`` around a CUDA-graph capture.  Packed 16-bit copies of TRAINABLE weights made before the
    capture are not trusted inside it: the pack kernels are recorded into the graph, so every replay re-packs from the
    current fp32 weights (a graph that captured a cache hit would read stale - or freed - buffers after the next
    optimizer step).  Frozen weights keep their cached copies; the graph owner holds references to them
    (``packed_weight_refs``) and drops the graph when they are reloaded."""

    def __enter__(self):
        _CAPTURE["id"] += 1
        _CAPTURE["active"] = True
        return self

    def __exit__(self, *exc):
        _CAPTURE["active"] = False
        return False


def _pack_cache(weight):
    """-> the per-parameter cache of packed copies, emptied when the parameter changed (in-place update: ``_version``; arena
    optimizer: WEIGHT_EPOCH; a capture in progress: copies of trainable weights made outside it)"""
    trainable = weight.data_ptr() in DIRECT_GRADS
    epoch = WEIGHT_EPOCH if trainable else -1
    cap = _CAPTURE["id"] if (trainable and _CAPTURE["active"]) else -1
    stamp = (weight._version, epoch, weight.data_ptr(), cap)
    cache = getattr(weight, "_dsr_pack", None)
    if cache is None or cache.get("stamp") != stamp:
        cache = {"stamp": stamp}
        try:
            weight._dsr_pack = cache
        except AttributeError:
            pass
    return cache


def packed_weight_refs(model):
    """strong references to every packed weight copy currently cached on the model's parameters"""
    keep = []
    for name in getattr(model, "model_names", []):
        net = getattr(model, "net" + name, None)
        if net is None:
            continue
        for p in net.parameters():
            c = getattr(p, "_dsr_pack", None)
            if c:
                keep.append([v for k, v in c.items() if k not in ("stamp", "_ev")])
    return keep


W_SCALE = 64.0          # power-of-two weight scale of the f16 path: keeps the low half of N(0, 0.02)-sized weights normal


def _tc_fmt(dtype=None):
    f16 = (dtype or CONFIG["dtype"]) == "f16"
    return int(f16), (1.0 / W_SCALE if f16 else 1.0)


def _rup(v, m):
    return (v + m - 1) // m * m


def _int_array(vals):
    import ctypes
    return (ctypes.c_int * len(vals))(*vals)


def tc_conv_plan(kind, Ci, Co, R, S, stride, pad, opad, H, W, min_ci=16):
    """Which arranged-operand layout serves this layer on the tcgen05 path (None -> CUDA-core path)."""
    if CONFIG["engine"] != "tc" or R != S:
        return None
    if kind == "conv" and stride == 1 and Ci <= 8 and R >= 5 and CONFIG["tc_first_layers"]:
        # 1..8-channel first layers: 8 horizontally adjacent pixels x 8 channels form one 64-wide K block per kernel row.
        # "compact": the forward GEMM reads that block straight out of an 8-channel arranged tensor (16-byte pixel rows,
        # overlapping MMA rows) instead of an 8x expanded copy; the expanded PAIR form still serves the weight gradient.
        plan = dict(layout=_LAYOUT_PAIR, variant=_W_CONV_PAIR, Cp=8, Ca=64, T=R * ((S + 7) // 8))
        if S <= 8 and CONFIG["tc_compact_first"]:
            plan["compact"] = True
        return plan
    if kind == "conv" and stride == 1 and R * S <= 64 and Ci >= 32:
        if Ci == 32:
            return dict(layout=_LAYOUT_PAIR, variant=_W_CONV_PAIR, Cp=32, Ca=64, T=R * ((S + 1) // 2))
        return dict(layout=_LAYOUT_NORMAL, variant=_W_CONV, Cp=_rup(Ci, 8), Ca=_rup(Ci, 64), T=R * S)
    if kind == "conv" and stride == 2 and R in (3, 4) and pad == 1 and H % 2 == 0 and W % 2 == 0 and Ci >= min_ci:
        Cp = _rup(Ci, 16)
        return dict(layout=_LAYOUT_S2D, variant=_W_CONV_S2D, Cp=Cp, Ca=4 * Cp, T=4)
    if kind == "convT" and stride == 2 and pad == 1 and ((R == 4 and opad == 0) or (R == 3 and opad == 1)) and Ci >= 32:
        return dict(layout=_LAYOUT_NORMAL, variant=_W_CONVT_PH, Cp=_rup(Ci, 8), Ca=_rup(Ci, 64), T=4)
    return None


_PACK_SEQ = [0]


def _remember_pack(weight, key, recipe):
    """the packed copies a parameter needs per step, so that `prepack` can make all of them ahead of time on the side stream"""
    if weight.data_ptr() in DIRECT_GRADS:
        try:
            rec = weight.__dict__.setdefault("_dsr_recipes", {})
            rec[key] = recipe
            seq = weight.__dict__.setdefault("_dsr_recipe_seq", {})
            if key not in seq:                       # position of the copy's FIRST use in the step (forward layers first, in
                _PACK_SEQ[0] += 1                    # forward order, then the data-gradient variants in backward order)
                seq[key] = _PACK_SEQ[0]
        except AttributeError:
            pass


def prepack(params):
    """Re-make every packed 16-bit weight copy the previous step used (forward AND data-gradient variants of the trainable
    nets), on the side stream: the ~60 small gather kernels overlap the frozen networks at the top of the step instead of
    sitting on the critical path of the trainable forward / backward passes.  The first consumer joins the stream
    (`_tc_weights` -> `_prepack_join`)."""
    if not CONFIG["prepack"]:
        return
    todo = [(p, r, getattr(p, "_dsr_recipe_seq", {}).get(k, 0)) for p in params for k, r in getattr(p, "_dsr_recipes", {}).items()]
    if not todo:
        return
    if CONFIG["prepack_in_use_order"]:
        # the arena lists the parameters in gradient-ready order - the REVERSE of the forward pass: packed in that order, the
        # copy the first trainable layer waits for was the last of ~60 launches (0.8 ms of serial small kernels at the head of
        # the training graph, timeline r3e).  In order of first use the forward pass follows right behind the packing chain.
        todo.sort(key=lambda t: t[2])
    todo = [(p, r) for p, r, _ in todo]
    with _on_side(True, todo[0][0].device):
        side = torch.cuda.current_stream()
        for w, (kind, plan, Co, phase, pad, dtype, npass) in todo:
            if kind == "phases":
                _tc_weights_phases(w, plan, Co, pad, dtype, npass=npass)
            else:
                _tc_weights(w, plan, Co, phase, pad, dtype, npass=npass)
            # one event per packed copy: its first consumer waits for THIS copy only (inside a captured graph an event is
            # just a dependency edge), so the trainable forward pass starts while the later layers are still being packed
            ev = torch.cuda.Event()
            ev.record(side)
            _pack_cache(w)["_ev"] = ev          # (a later copy of the same weight overwrites it: same stream, so it covers both)
    _SIDE["prepack_pending"] = True


def _prepack_join(weight):
    """the consumer's stream waits for the packed copies of THIS weight (made on the side stream by `prepack`)"""
    if _SIDE.get("prepack_pending") and weight.data_ptr() in DIRECT_GRADS:
        c = getattr(weight, "_dsr_pack", None)
        ev = c.get("_ev") if c else None
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        elif torch.cuda.current_stream() == _SIDE.get("main"):     # no event (copy made elsewhere): join the whole stream
            _SIDE["prepack_pending"] = False
            join_side()


def _tc_weights(weight, plan, Co, phase=(0, 0), pad=0, dtype=None, npass=None):
    """bf16 hi/lo packed copy of a parameter, cached ON the parameter object until it changes
    (in-place updates bump ``_version``; the arena optimizer bumps WEIGHT_EPOCH)."""
    prepacking = npass is not None
    if not prepacking:
        _prepack_join(weight)
        npass = _passes(dtype)
    f16 = (dtype or CONFIG["dtype"]) == "f16"
    key = (plan["variant"], plan["Ca"], phase, pad, npass >= 3, f16)
    cache = _pack_cache(weight)
    hit = cache.get(key)
    if hit is not None:
        return hit
    _remember_pack(weight, key, ("conv", dict(plan), Co, phase, pad, dtype, npass))
    D0, D1, R, S = weight.shape
    w = weight.detach()
    w = w if w.is_contiguous() else w.contiguous()
    K = plan["T"] * plan["Ca"]
    whi = torch.empty((Co, K), device=w.device, dtype=torch.bfloat16)
    wlo = torch.empty((Co, K), device=w.device, dtype=torch.bfloat16) if npass >= 3 else None
    _call("dsr_tc_pack_weight", _p(w), D0, D1, R, S, plan["variant"], plan["Cp"], phase[0], phase[1], pad, Co,
          plan["T"], plan["Ca"], _p(whi, torch.bfloat16), _p(wlo, torch.bfloat16), int(f16), W_SCALE if f16 else 1.0)
    cache[key] = (whi, wlo)
    return whi, wlo


# Operands made by the PRODUCER of an activation (the fused norm + skip-add + preparation pass at the end of a residual
# block): data_ptr -> (the fp32 tensor itself - held so that its address cannot be handed to another tensor while the entry
# lives -, cache key, (ahi, alo, Ha, Wa), bf16 plane or None, bias-gradient sums or None).  Claimed once by the _Prepared of the
# consumer, dropped at the top of every step.  The backward pass uses the same table for the dY operands the fused
# InstanceNorm backward makes (_in_bwd).
_PREMADE = {}
# What the backward pass of a convolution asks of its dY the first time (weight data_ptr -> dict(key = the _Prepared cache key,
# need_lo, csum, shape)): recorded by _Prepared.get, read by _in_bwd in the NEXT backward pass (the producer of a dY runs before
# its consumer, so the request of the previous step is the only one it can know).  A stale or foreign entry costs one unused
# operand at worst: the consumer looks its operand up by key and makes its own when the key differs.
_DY_SPEC = {}


def operand_hint(kind, x_shape, weight, stride, pad, pad_mode=PAD_ZERO, opad=0):
    """What the convolution (kind 'conv' / 'convT') that will read an (N, C, H, W) activation asks of its operand
    preparation: dict(plan, pad, mode, bf16) - or None when that layer does not take a plain NORMAL / S2D operand."""
    if not CONFIG["fuse_norm_prep"] or CONFIG["engine"] != "tc":
        return None
    N, C, H, W = x_shape
    pad_mode = PAD_MODES[pad_mode] if isinstance(pad_mode, str) else pad_mode
    trains = bool(weight.requires_grad) and torch.is_grad_enabled()
    if kind == "conv":
        Co, Ci, R, S = weight.shape
        if Ci != C or (Co == 1 and stride == 1 and CONFIG["out1"]):
            return None
        plan = tc_conv_plan("conv", Ci, Co, R, S, stride, pad, 0, H, W)
    else:
        Ci, Co, R, S = weight.shape
        if Ci != C or (Co == 1 and CONFIG["out1"]):
            return None
        plan = tc_conv_plan("convT", Ci, Co, R, S, stride, pad, opad, H, W)
        pad, pad_mode = 1, PAD_ZERO                       # _tc_convT_fwd: zero halo of 1
    if plan is None or plan.get("compact") or plan["layout"] not in (_LAYOUT_NORMAL, _LAYOUT_S2D):
        return None
    cg = plan["Ca"] // 8
    if C % 8 or cg > 256 or cg & (cg - 1):
        return None
    return dict(plan=plan, pad=pad, mode=pad_mode, bf16=_bwd_copy_wanted(trains))


def _norm_res_prep(xh, sums, groups, gamma, beta, eps, act, rh, hint):
    """one launch: y = act(norm(x)) + res (fp32) AND the arranged operand of the next convolution, registered in _PREMADE.
    -> (y, prm)"""
    N, H, W, C = xh.shape
    plan, pad, mode = hint["plan"], hint["pad"], hint["mode"]
    Hq, Wq = H + 2 * pad, W + 2 * pad
    Ha, Wa = ((Hq + 1) // 2, (Wq + 1) // 2) if plan["layout"] == _LAYOUT_S2D else (Hq, Wq)
    Ca, dt = plan["Ca"], CONFIG["dtype"]
    prm = torch.empty(3 * N * C, device=xh.device, dtype=torch.float32)
    y = torch.empty_like(xh)
    ahi = torch.empty((N, Ha, Wa, Ca), device=xh.device, dtype=torch.bfloat16)
    alo = torch.empty((N, Ha, Wa, Ca), device=xh.device, dtype=torch.bfloat16) if CONFIG["passes"] >= 2 else None
    abf = torch.empty((N, Ha, Wa, Ca), device=xh.device, dtype=torch.bfloat16) if (hint["bf16"] and dt == "f16") else None
    if _lib.PROFILE is not None:
        _lib.PROFILE_META = dict(macs=0, shape=(N, H, W, C, Ca, plan["layout"], pad, 1, int(alo is not None)))
    _call("dsr_tc_prep_norm_res", _p(xh), N, H, W, C, _p(sums, torch.float64), groups, _p(gamma), _p(beta), eps, _p(prm), act,
          _p(rh), _p(y), pad, mode, plan["layout"], plan["Cp"], _p(ahi, torch.bfloat16), _p(alo, torch.bfloat16),
          _p(abf, torch.bfloat16), Ha, Wa, Ca, int(dt == "f16"))
    key = (plan["layout"], plan["Cp"], plan["Ca"], pad, mode, dt, plan.get("Wa", 0))
    if len(_PREMADE) > 64:
        _PREMADE.clear()
    _PREMADE[y.data_ptr()] = (y, key, (ahi, alo, Ha, Wa), abf, None)
    return y, prm


class _Prepared:
    """Arranged 16-bit copies of ONE fp32 NHWC tensor, made at most once per (layout, padding, format):
    the backward pass feeds the same dY (or x) to the data-gradient and the weight-gradient GEMMs."""

    def __init__(self, xh, prm=None, act=ACT_NONE, slope=0.0, want_csum=False, rec=None):
        self.xh, self.made = xh, {}
        self.rec = rec                                        # the parameter whose backward this dY belongs to: its FIRST operand
                                                              # request is recorded there for the producer of the next step's dY
        self.shape, self.device = xh.shape, xh.device
        self.prm, self.act, self.slope = prm, act, slope      # fused prologue: norm-apply + activation on the way in
        self.want_csum, self.csum = want_csum, None           # per-channel sums of xh (bias gradient), taken by the first
                                                              # prep that writes every element exactly once
        if prm is None and act == ACT_NONE and _PREMADE:
            pre = _PREMADE.pop(xh.data_ptr(), None)           # operand already made by the producer of xh (_norm_res_prep)
            if pre is not None and pre[0].shape == xh.shape:
                _, key, hit, abf, csum = pre
                self.made[key] = hit
                if abf is not None:
                    self.made[key[:5] + ("bf16",) + key[6:]] = (abf, None, hit[2], hit[3])
                if csum is not None and want_csum:
                    self.csum = csum

    def get(self, plan, pad, pad_mode, dtype=None, need_lo=True, also_bf16=False):
        """need_lo=False: the caller reads the high plane only (single-pass weight gradient) - skip writing the low one.
        also_bf16: the same pass also writes a plain bf16 copy in the same arranged layout (the operand the weight-gradient
        GEMM of the backward pass needs - both its operands must have ONE 16-bit format), found later under dtype 'bf16'."""
        dt = dtype or CONFIG["dtype"]
        key = (plan["layout"], plan["Cp"], plan["Ca"], pad, pad_mode, dt, plan.get("Wa", 0))
        if self.rec is not None:
            rec, self.rec = self.rec, None
            ok = CONFIG["fuse_bwd_prep"] and self.prm is None and self.act == ACT_NONE and (pad == 0 or pad_mode == PAD_ZERO) \
                and plan["layout"] in (_LAYOUT_NORMAL, _LAYOUT_S2D)
            _DY_SPEC[rec.data_ptr()] = dict(key=key, need_lo=bool(need_lo), csum=bool(self.want_csum), shape=tuple(self.shape)) if ok else None
        hit = self.made.get(key)
        if hit is not None and need_lo and hit[1] is None and CONFIG["passes"] >= 2:
            hit = None                                    # made without its low plane earlier: make it again in full
        if hit is None:
            csum = None
            if self.want_csum and self.csum is None and plan["layout"] != _LAYOUT_PAIR and (pad == 0 or pad_mode == PAD_ZERO):
                csum = self.csum = _zeros_f64(self.shape[3] * CONFIG["csum_reps"], self.device)
            kbf = key[:5] + ("bf16",) + key[6:]
            also_bf16 = also_bf16 and dt == "f16" and kbf not in self.made
            hit = self._make(plan, pad, pad_mode, dtype, csum, need_lo, also_bf16)
            if also_bf16:
                self.made[kbf] = (hit[4], None, hit[2], hit[3])
            hit = self.made[key] = hit[:4]
        return hit

    def _make(self, plan, pad, pad_mode, dtype, csum, need_lo, also_bf16):
        return _tc_prep(self.xh, plan, pad, pad_mode, self.prm, self.act, self.slope, dtype=dtype, csum=csum, need_lo=need_lo,
                        also_bf16=also_bf16)

    def any_normal(self, Ca, dtype):
        """an already-made NORMAL-layout zero-padded copy -> (ahi, alo, Ha, Wa, pad) or None"""
        for (layout, _, ca, pad, mode, dt, _wa), v in self.made.items():
            if layout == _LAYOUT_NORMAL and ca == Ca and dt == dtype and (mode == PAD_ZERO or pad == 0):
                return v + (pad,)
        return None


class _PreparedCat(_Prepared):
    """_Prepared over torch.cat(parts, dim=1) of NHWC tensors that is never materialised: the preparation kernel reads the
    parts side by side (dsr_tc_prep_cat; main_model.py:305-306, networks.py:629)."""

    def __init__(self, parts):
        if len(parts) > 4:
            raise ValueError("at most four concatenated sources")
        N, H, W, _ = parts[0].shape
        self.parts = list(parts)
        self.xh, self.made = None, {}
        self.shape, self.device = (N, H, W, sum(p.shape[3] for p in parts)), parts[0].device
        self.prm, self.act, self.slope, self.want_csum, self.csum = None, ACT_NONE, 0.0, False, None
        self.rec = None

    def _make(self, plan, pad, pad_mode, dtype, csum, need_lo, also_bf16):
        N, H, W, C = self.shape
        Hq, Wq = H + 2 * pad, W + 2 * pad
        Ha, Wa = ((Hq + 1) // 2, (Wq + 1) // 2) if plan["layout"] == _LAYOUT_S2D else (Hq, Wq)
        Ca = plan["Ca"]
        ahi = torch.empty((N, Ha, Wa, Ca), device=self.device, dtype=torch.bfloat16)
        alo = torch.empty((N, Ha, Wa, Ca), device=self.device, dtype=torch.bfloat16) if (CONFIG["passes"] >= 2 and need_lo) else None
        abf = torch.empty((N, Ha, Wa, Ca), device=self.device, dtype=torch.bfloat16) if also_bf16 else None
        src = []
        for i in range(4):
            src += [_p(self.parts[i]), self.parts[i].shape[3]] if i < len(self.parts) else [None, 0]
        if _lib.PROFILE is not None:
            _lib.PROFILE_META = dict(macs=0, shape=(N, H, W, C, Ca, plan["layout"], pad, 0, int(alo is not None)))
        _call("dsr_tc_prep_cat", *src, N, H, W, pad, pad_mode, plan["layout"], plan["Cp"], _p(ahi, torch.bfloat16),
              _p(alo, torch.bfloat16), _p(abf, torch.bfloat16), Ha, Wa, Ca, int((dtype or CONFIG["dtype"]) == "f16"))
        return ahi, alo, Ha, Wa, abf


def _tc_weights_phases(weight, plan, Co, pad, dtype=None, npass=None):
    """the four output-phase weight matrices of a stride-2 transposed conv, stacked along rows ([4*Co][T*Ca]) so one GEMM
    launch serves all phases; cached on the parameter like _tc_weights"""
    if npass is None:
        _prepack_join(weight)
        npass = _passes(dtype)
    f16 = (dtype or CONFIG["dtype"]) == "f16"
    key = ("phases", plan["Ca"], pad, npass >= 3, f16)
    cache = _pack_cache(weight)
    hit = cache.get(key)
    if hit is not None:
        return hit
    _remember_pack(weight, key, ("phases", dict(plan), Co, None, pad, dtype, npass))
    D0, D1, R, S = weight.shape
    w = weight.detach()
    w = w if w.is_contiguous() else w.contiguous()
    K = plan["T"] * plan["Ca"]
    whi = torch.empty((4 * Co, K), device=w.device, dtype=torch.bfloat16)
    wlo = torch.empty((4 * Co, K), device=w.device, dtype=torch.bfloat16) if npass >= 3 else None
    # phase -1: one launch packs the four phase matrices into the stacked rows (grid.y = phase)
    _call("dsr_tc_pack_weight", _p(w), D0, D1, R, S, plan["variant"], plan["Cp"], -1, -1, pad, Co, plan["T"], plan["Ca"],
          _p(whi, torch.bfloat16), _p(wlo, torch.bfloat16) if wlo is not None else None, int(f16), W_SCALE if f16 else 1.0)
    cache[key] = (whi, wlo)
    return whi, wlo


def _tc_prep(xh, plan, pad, pad_mode, prm=None, act=ACT_NONE, slope=0.0, dtype=None, csum=None, need_lo=True, also_bf16=False):
    """fp32 NHWC -> arranged bf16 hi(+lo) operand (+ the plain bf16 copy as a fifth element with also_bf16)."""
    if isinstance(xh, _Prepared):
        return xh.get(plan, pad, pad_mode, dtype, need_lo, also_bf16)
    N, H, W, C = xh.shape
    Hq, Wq = H + 2 * pad, W + 2 * pad
    if plan["layout"] == _LAYOUT_S2D:
        Ha, Wa = (Hq + 1) // 2, (Wq + 1) // 2
    else:
        Ha, Wa = Hq, max(Wq, plan.get("Wa", 0))        # plan['Wa']: wider rows, zero-filled beyond the padded image
    Ca = plan["Ca"]
    ahi = torch.empty((N, Ha, Wa, Ca), device=xh.device, dtype=torch.bfloat16)
    alo = torch.empty((N, Ha, Wa, Ca), device=xh.device, dtype=torch.bfloat16) if (CONFIG["passes"] >= 2 and need_lo) else None
    if _lib.PROFILE is not None:
        _lib.PROFILE_META = dict(macs=0, shape=(N, H, W, C, Ca, plan["layout"], pad, int(prm is not None), int(alo is not None)))
    abf = torch.empty((N, Ha, Wa, Ca), device=xh.device, dtype=torch.bfloat16) if also_bf16 else None
    fin = _take_fin(prm)
    if fin is not None:
        sums, _, _, P, groups, gamma, beta, eps = fin    # statistics not finalised yet: this launch does it (and writes prm)
        _call("dsr_tc_prep_fin", _p(xh), N, H, W, C, _p(sums, torch.float64), groups, _p(gamma), _p(beta), eps, _p(prm),
              act, slope, pad, pad_mode, plan["layout"], plan["Cp"],
              _p(ahi, torch.bfloat16), _p(alo, torch.bfloat16), _p(abf, torch.bfloat16), Ha, Wa, Ca,
              int((dtype or CONFIG["dtype"]) == "f16"), _p(csum, torch.float64), (csum.numel() // C) if csum is not None else 1)
    else:
        _call("dsr_tc_prep", _p(xh), N, H, W, C, _p(prm), act, slope, pad, pad_mode, plan["layout"], plan["Cp"],
              _p(ahi, torch.bfloat16), _p(alo, torch.bfloat16), _p(abf, torch.bfloat16), Ha, Wa, Ca,
              int((dtype or CONFIG["dtype"]) == "f16"), _p(csum, torch.float64), (csum.numel() // C) if csum is not None else 1)
    return (ahi, alo, Ha, Wa, abf) if also_bf16 else (ahi, alo, Ha, Wa)


def _tc_kernel_for(N, Ht, Wt, Co, ds, nphase=1):
    """which GEMM kernel serves this output grid: 3 = channel-major (csrc/conv_tc3.cu, Cout >= 128), 2 = halo
    pixel-major (csrc/conv_tc2.cu), 1 = first generation with split-K (few tiles, long K; forced split-K)."""
    if not CONFIG["tc_halo"] or CONFIG["split_k"] > 1 or Ht * Wt < 128 or Wt < 8 or max(ds) > 8:
        return 1
    if Co >= 128 and Ht >= 16 and CONFIG["tc_cm"]:
        tiles = nphase * N * ((Ht + 31) // 32) * ((Wt + 7) // 8) * ((Co + 127) // 128)
        return 3 if tiles >= CONFIG["halo_min_tiles"] else 1
    bn = 256 if Co >= 256 else 128
    tiles = nphase * N * ((Ht + 15) // 16) * ((Wt + 7) // 8) * ((Co + bn - 1) // bn)
    return 2 if tiles >= CONFIG["halo_min_tiles"] else 1


def _tc_gemm(ahi, alo, N, Ha, Wa, Ca, whi, wlo, Co, T, dr, ds, aoh, aow, Ht, Wt, bias, y, Ho, Wo, os_, ph, pw, act_out,
             dtype, split_k, stats=None, nphase=1, a_mode=0):
    """one GEMM launch on the best kernel for the shape; returns True when `stats` was filled by the epilogue"""
    which = 2 if a_mode else (CONFIG.get("force_kernel") or _tc_kernel_for(N, Ht, Wt, Co, list(ds), nphase))
    if _lib.PROFILE_META is not None:
        _lib.PROFILE_META["passes"] = _passes(dtype)            # bench.py: MMAs actually issued per product
    if which == 1:
        _call("dsr_tc_gemm", _p(ahi, torch.bfloat16), _p(alo, torch.bfloat16), N, Ha, Wa, Ca,
              _p(whi, torch.bfloat16), _p(wlo, torch.bfloat16), Co, T, dr, ds, aoh, aow, Ht, Wt, _p(bias), _p(y),
              Ho, Wo, os_, ph, pw, nphase, act_out, _passes(dtype), split_k, *_tc_fmt(dtype))
        return False
    extra = (a_mode,) if which == 2 else ()
    _call("dsr_tc_gemm3" if which == 3 else "dsr_tc_gemm2", _p(ahi, torch.bfloat16), _p(alo, torch.bfloat16), N, Ha, Wa, Ca,
          _p(whi, torch.bfloat16), _p(wlo, torch.bfloat16), Co, T, dr, ds, aoh, aow, Ht, Wt, _p(bias), _p(y),
          Ho, Wo, os_, ph, pw, nphase, act_out, _passes(dtype), *_tc_fmt(dtype), _p(stats, torch.float64), *extra)
    return stats is not None


def _bwd_copy_wanted(want):
    """forward operand prep also emits the bf16 copy for the weight-gradient GEMM (single-pass bf16 weight gradients only)"""
    return bool(want) and CONFIG["reuse_fwd_operand"] and CONFIG["tc_backward"] and CONFIG["dtype"] == "f16" and \
        CONFIG["bwd_dtype"] == "bf16" and min(CONFIG["wgrad_passes"], CONFIG["passes"]) == 1 and CONFIG["fwd_bf16_copy"]


def _tc_conv_fwd(xh, weight, bias, plan, stride, pad, pad_mode, act_out, Ho, Wo, dtype=None, Co=None, macs=None,
                 stats=None, bwd_copy=False):
    """stride-1 / stride-2 convolution of `xh` with a Conv2d-layout weight on the tcgen05 path.
    Also serves as the dgrad of ConvTranspose2d (its weight read as a Conv2d weight) and, with
    plan['variant'] == _W_CONV_DGRAD, as the dgrad of a stride-1 Conv2d (flipped / transposed taps)."""
    _FWD["hw"] = Ho * Wo
    N, H, W, Ci = xh.shape
    R, S = weight.shape[2], weight.shape[3]
    if Co is None:
        Co = weight.shape[0]
    whi, wlo = _tc_weights(weight, plan, Co, dtype=dtype)
    dr, ds = _tc_taps(plan, R, S)
    a_mode, a_plan = 0, plan
    if plan.get("compact") and Ho * Wo >= 128 and Wo >= 8:
        a_mode, a_plan = 1, dict(layout=_LAYOUT_NORMAL, Cp=8, Ca=8)       # 8-channel operand, same weights / taps
    ahi, alo, Ha, Wa = _tc_prep(xh, a_plan, pad, pad_mode, dtype=dtype, need_lo=_passes(dtype) >= 2,
                                also_bf16=_bwd_copy_wanted(bwd_copy) and a_mode == 0 and isinstance(xh, _Prepared))[:4]
    y = torch.empty((N, Ho, Wo, Co), device=xh.device, dtype=torch.float32)
    _lib.PROFILE_META = dict(macs=macs if macs is not None else N * Ho * Wo * Co * Ci * R * S, shape=(N, H, W, Ci, Co, R, stride))
    filled = _tc_gemm(ahi, alo, N, Ha, Wa, a_plan["Ca"], whi, wlo, Co, plan["T"], _int_array(dr), _int_array(ds), 0, 0, Ho, Wo,
                      bias, y, Ho, Wo, 1, 0, 0, act_out, dtype, CONFIG["split_k"], stats, a_mode=a_mode)
    if stats is not None:
        stats.filled = filled
    return y


def _tc_convT_fwd(xh, weight, bias, plan, pad, act_out, Ho, Wo, dtype=None, stats=None, bwd_copy=False):
    """stride-2 transposed convolution (4 output phases) of `xh` with a ConvTranspose2d-layout weight;
    also the dgrad of a stride-2 Conv2d (whose (Cout, Cin, R, S) weight IS a ConvTranspose2d weight
    from Cout to Cin channels)."""
    _FWD["hw"] = Ho * Wo
    N, H, W, Ci = xh.shape
    _, Co, R, S = weight.shape
    ahi, alo, Ha, Wa = _tc_prep(xh, plan, 1, PAD_ZERO, dtype=dtype, need_lo=_passes(dtype) >= 2,     # zero halo of 1
                                also_bf16=_bwd_copy_wanted(bwd_copy) and isinstance(xh, _Prepared))[:4]
    y = torch.empty((N, Ho, Wo, Co), device=xh.device, dtype=torch.float32)
    dr, ds = _int_array([0, 0, 1, 1]), _int_array([0, 1, 0, 1])
    Ht, Wt = (Ho + 1) // 2, (Wo + 1) // 2
    if Ho % 2 or Wo % 2:
        raise ValueError("transposed convolution phases need an even output size")
    whi, wlo = _tc_weights_phases(weight, plan, Co, pad, dtype)
    _lib.PROFILE_META = dict(macs=N * H * W * Co * Ci * R * S, shape=(N, H, W, Ci, Co, R, -2))
    filled = _tc_gemm(ahi, alo, N, Ha, Wa, plan["Ca"], whi, wlo, Co, 4, dr, ds, 0, 0, Ht, Wt, bias, y, Ho, Wo, 2, 0, 0,
                      act_out, dtype, CONFIG["split_k"] if act_out == ACT_NONE else 1, stats, nphase=4)
    if stats is not None:
        stats.filled = filled
    return y


def _tc_dgrad_group(gP, weight, padq, Hout, Wout, dt, g=2):
    """Stride-1 data gradient of a Conv2d with few input channels (the 7x7 32 -> 128 head of the ResNets), written as a
    GEMM whose rows are GROUPS of g horizontally adjacent output pixels.  g = 2: N = 2 * Cin = 64 runs the pixel-major
    tensor pipe at full rate where N = 32 runs at half, for (S/2 + 1) * 2 / S = 8/7 of the MACs.  g = 4: 4 * Cin = 128 GEMM
    rows are exactly one M tile of the channel-major kernel (csrc/conv_tc3.cu: N = 256 pixel groups per MMA, 95 % of the
    tensor floor against the ~105-cycle floor of the pixel-major M = 128 x N = 64 MMAs) for 12/7 of the MACs.
    The zero-padded dY (N, Ha, Wa, Co) is read through the view (N, Ha, Wa/g, g*Co) - no copy - and the output through
    (N, Hout, Wc/g, g*Cin), Wc = Wout rounded up to whole groups.  -> (dX of shape (N, Hout, Wc, Cin), Wc)."""
    _FWD["hw"] = Hout * Wout
    N, Ho, Wo, Co = gP.shape
    _, Ci, R, S = weight.shape
    Sg = (g + S - 2) // g + 1
    T = R * Sg
    Wc = _rup(Wout, g)
    Gout = Wc // g
    a_plan = dict(layout=_LAYOUT_NORMAL, Cp=Co, Ca=Co, Wa=g * (Gout + Sg - 1))      # >= Wo + 2 * padq, zero-filled beyond
    ahi, alo, Ha, Wa = _tc_prep(gP, a_plan, padq, PAD_ZERO, dtype=dt, need_lo=_passes(dt) >= 2)
    w_plan = dict(variant=_W_CONV_DGRAD_PAIR, Cp=Co, Ca=g * Co, T=T)
    whi, wlo = _tc_weights(weight, w_plan, g * Ci, dtype=dt)
    dr, ds = [t // Sg for t in range(T)], [t % Sg for t in range(T)]
    y = torch.empty((N, Hout, Wc, Ci), device=gP.device, dtype=torch.float32)
    _lib.PROFILE_META = dict(macs=N * Hout * Wout * Co * Ci * R * S, shape=(N, Hout, Wout, Co, Ci, R, 1))
    _tc_gemm(ahi, alo, N, Ha, Wa // g, g * Co, whi, wlo, g * Ci, T, _int_array(dr), _int_array(ds), 0, 0, Hout, Gout,
             None, y, Hout, Gout, 1, 0, 0, ACT_NONE, dt, CONFIG["split_k"])
    return y, Wc


def _tc_conv_dgrad(g, weight, stride, pad, pad_mode, H, W, add=None):
    """dL/dx of a Conv2d on the tcgen05 path (None when the shape is not covered).  g: (N,Ho,Wo,Co).
    add = [t]: a second gradient of x, (N, H, W, Ci), to be added; the route that can do so in its own last pass does and
    sets add[0] = None."""
    if not CONFIG["tc_backward"]:
        return None
    N, Ho, Wo, Co = g.shape
    _, Ci, R, S = weight.shape
    dt = CONFIG["bwd_dtype"]
    if stride == 1:
        plan = tc_conv_plan("conv", Co, Ci, R, S, 1, 0, 0, Ho, Wo)
        if plan is None or plan["layout"] != _LAYOUT_NORMAL:
            return None
        plan = dict(plan, variant=_W_CONV_DGRAD)
        macs = N * Ho * Wo * Co * Ci * R * S
        direct = pad_mode == PAD_ZERO or pad == 0      # gradient w.r.t. the un-padded input directly
        Wout = W if direct else W + 2 * pad
        grp = 0                                        # pixels per GEMM row of the grouped form (0: plain)
        if CONFIG["dgrad_pair"] and Ci <= 32 and Ci % 8 == 0 and Co % 64 == 0 and isinstance(g, _Prepared) and S >= 3:
            if (CONFIG["dgrad_quad"] and Ci == 32 and R * ((S + 2) // 4 + 1) <= 64 and Wout >= 32 and H >= 16
                    and (Wout % 4 == 0 or not direct)):
                grp = 4
            elif W % 2 == 0 and Wo % 2 == 0 and R * (S // 2 + 1) <= 64 and (W + 2 * pad) // 2 >= 8:
                grp = 2
        if direct:
            if grp:
                return _tc_dgrad_group(g, weight, R - 1 - pad, H, W, dt, grp)[0]
            return _tc_conv_fwd(g, weight, None, plan, 1, R - 1 - pad, PAD_ZERO, ACT_NONE, H, W, dtype=dt, Co=Ci, macs=macs)
        Wc = Wout
        if grp:
            gxp, Wc = _tc_dgrad_group(g, weight, R - 1, H + 2 * pad, Wout, dt, grp)
        else:
            gxp = _tc_conv_fwd(g, weight, None, plan, 1, R - 1, PAD_ZERO, ACT_NONE, H + 2 * pad, W + 2 * pad, dtype=dt, Co=Ci, macs=macs)
        gx = torch.empty((N, H, W, Ci), device=g.device, dtype=torch.float32)
        if add is not None and add[0] is not None and add[0].shape == gx.shape:
            _call("dsr_pad2d_bwd_pitch_add", _p(gxp), _p(add[0]), _p(gx), N, H, W, Ci, pad, pad_mode, Wc)
            add[0] = None
        else:
            _call("dsr_pad2d_bwd_pitch", _p(gxp), _p(gx), N, H, W, Ci, pad, pad_mode, Wc)
        return gx
    if stride == 2 and (pad_mode == PAD_ZERO or (pad == 1 and CONFIG["s2_border"])):
        opad = H - ((Ho - 1) * 2 - 2 * pad + R)
        plan = tc_conv_plan("convT", Co, Ci, R, S, 2, pad, opad, Ho, Wo)
        if plan is None or W - ((Wo - 1) * 2 - 2 * pad + S) != opad:
            return None
        gx = _tc_convT_fwd(g, weight, None, plan, pad, ACT_NONE, H, W, dtype=dt)
        if pad_mode != PAD_ZERO:
            # replicate / reflect padding: the interior of the padded-input gradient IS the zero-padding data gradient;
            # its one-pixel frame is folded onto the border pixels by a small CUDA-core kernel (csrc/conv_out1.cu)
            w = weight.detach()
            gh = g.xh if isinstance(g, _Prepared) else g
            _call("dsr_conv_s2_border_dgrad", _p(gh), N, Ho, Wo, Co, _p(w if w.is_contiguous() else w.contiguous()), Ci, R, S,
                  pad_mode, _p(gx), H, W)
        return gx
    return None


def _tc_convT_dgrad(g, weight, stride, pad, H, W):
    """dL/dx of a ConvTranspose2d = stride-2 Conv2d of g with the same weight read as (Cout=Ci, Cin=Co, R, S)."""
    if not CONFIG["tc_backward"]:
        return None
    N, Ho, Wo, Co = g.shape
    Ci, _, R, S = weight.shape
    plan = tc_conv_plan("conv", Co, Ci, R, S, stride, pad, 0, Ho, Wo)
    if plan is None or (Ho + 2 * pad - R) // stride + 1 != H or (Wo + 2 * pad - S) // stride + 1 != W:
        return None
    return _tc_conv_fwd(g, weight, None, plan, stride, pad, PAD_ZERO, ACT_NONE, H, W, dtype=CONFIG["bwd_dtype"], Co=Ci)


def _tc_wgrad_m_operand(M, Cm_real):
    """the zero-padded NORMAL 16-bit copy of the weight-gradient GEMM's row operand (made by the data gradient when one ran)"""
    if not CONFIG["tc_backward"] or CONFIG["engine"] != "tc":
        return
    dt, npass = CONFIG["bwd_dtype"], min(CONFIG["wgrad_passes"], CONFIG["passes"])
    Cm = _rup(Cm_real, 64)
    if M.any_normal(Cm, dt) is None:
        M.get(dict(layout=_LAYOUT_NORMAL, Cp=_rup(Cm_real, 8), Ca=Cm), 0, PAD_ZERO, dt, need_lo=npass >= 2)


def _tc_wgrad(M, Cm_real, A, a_plan, a_pad, a_pad_mode, dr, ds, Hb, Wb, weight, variant):
    """weight gradient GEMM over the base grid (Hb x Wb x N):  rows = channels of M (prepared NORMAL, zero pad),
    columns = arranged channels of A x taps; result unpacked / accumulated into the parameter's gradient."""
    dt, npass = CONFIG["bwd_dtype"], min(CONFIG["wgrad_passes"], CONFIG["passes"])
    Cm = _rup(Cm_real, 64)
    # both operands must have ONE 16-bit format (tcgen05 kind::f16 with A = bf16, B = f16 is an illegal instruction -
    # measured); with bf16 forward operands the copies the forward pass made are found here and reused
    got = M.any_normal(Cm, dt)
    if got is None:
        m_plan = dict(layout=_LAYOUT_NORMAL, Cp=_rup(Cm_real, 8), Ca=Cm)
        got = M.get(m_plan, 0, PAD_ZERO, dt, need_lo=npass >= 2) + (0,)
    mhi, mlo, Hm, Wm, mpad = got
    if npass >= 2 and mlo is None:
        npass = 1
    ahi, alo, Ha, Wa = A.get(a_plan, a_pad, a_pad_mode, dt, need_lo=npass >= 3)
    if npass >= 3 and alo is None:
        npass = 2
    N = M.shape[0]
    T, Ca = a_plan["T"], a_plan["Ca"]
    D0, D1, R, S = weight.shape
    f16 = 3 if dt == "f16" else 0                                        # bit 0: format of M, bit 1: format of A
    _lib.PROFILE_META = dict(macs=N * Hb * Wb * D0 * D1 * R * S, shape=(N, Hb, Wb, Cm_real, Ca, T, 0), passes=npass)
    nsplit = 1
    if CONFIG["wgrad_kernel"] == 2 and npass == 1:
        if CONFIG["wgrad_slabs"]:
            # every K split stores its own slab; the unpack below sums them in a fixed order (no memset, no atomics)
            nsplit = _lib.load().dsr_tc_wgrad2_splits(N, Hb, Wb, Cm_real, Ca, T, CONFIG["split_k"])
            if nsplit < 1:
                raise RuntimeError("dsr_tc_wgrad2_splits: bad weight-gradient shape")
        dwp = torch.empty((nsplit, Cm_real, T * Ca), device=M.device, dtype=torch.float32)
        _call("dsr_tc_wgrad2p", _p(mhi, torch.bfloat16), N, Hm, Wm, Cm, Cm_real, mpad, mpad, _p(ahi, torch.bfloat16), Ha, Wa, Ca, T,
              _int_array(dr), _int_array(ds), 0, 0, Hb, Wb, _p(dwp), f16, 1.0, nsplit if CONFIG["wgrad_slabs"] else CONFIG["split_k"],
              int(CONFIG["wgrad_slabs"]))
    else:
        dwp = torch.empty((Cm_real, T * Ca), device=M.device, dtype=torch.float32)
        _call("dsr_tc_wgrad", _p(mhi, torch.bfloat16), _p(mlo, torch.bfloat16), N, Hm, Wm, Cm, Cm_real, mpad, mpad,
              _p(ahi, torch.bfloat16), _p(alo, torch.bfloat16), Ha, Wa, Ca, T, _int_array(dr), _int_array(ds), 0, 0,
              Hb, Wb, _p(dwp), npass, f16, 1.0, -1)
    _keep(mhi, mlo, ahi, alo, M.xh, A.xh, *getattr(M, "parts", ()), *getattr(A, "parts", ()))
    tgt = DIRECT_GRADS.get(weight.data_ptr())
    if tgt is not None:
        _call("dsr_tc_unpack_wgrad_splits", _p(dwp), nsplit, D0, D1, R, S, variant, a_plan["Cp"], T, Ca, _p(tgt), 1)
        _keep(dwp)
        _grad_ready(weight)
        return None
    gw = torch.empty(weight.shape, device=M.device, dtype=torch.float32)
    _call("dsr_tc_unpack_wgrad_splits", _p(dwp), nsplit, D0, D1, R, S, variant, a_plan["Cp"], T, Ca, _p(gw), 0)
    _keep(dwp)
    return gw


def _tc_taps(plan, R, S):
    if plan["layout"] == _LAYOUT_PAIR:
        g = plan["Ca"] // plan["Cp"]
        Sg = (S + g - 1) // g
        return [t // Sg for t in range(plan["T"])], [g * (t % Sg) for t in range(plan["T"])]
    if plan["layout"] == _LAYOUT_S2D:
        return [0, 0, 1, 1], [0, 1, 0, 1]
    return [t // S for t in range(R * S)], [t % S for t in range(R * S)]


def _tc_conv_wgrad(xP, gP, weight, stride, pad, pad_mode):
    """dL/dW of a Conv2d on the tcgen05 path -> (handled, grad or None).  xP / gP: _Prepared x (N,H,W,Ci), dY."""
    if not CONFIG["tc_backward"]:
        return False, None
    N, H, W, Ci = xP.shape
    _, Ho, Wo, Co = gP.shape
    _, _, R, S = weight.shape
    plan = tc_conv_plan("conv", Ci, Co, R, S, stride, pad, 0, H, W)
    if plan is None:
        return False, None
    dr, ds = _tc_taps(plan, R, S)
    return True, _tc_wgrad(gP, Co, xP, plan, pad, pad_mode, dr, ds, Ho, Wo, weight, plan["variant"])


def _tc_convT_wgrad(xP, gP, weight, stride, pad):
    """dL/dW of a stride-2 ConvTranspose2d = the weight gradient of the adjoint Conv2d (dY -> x) whose
    (Cout=Ci, Cin=Co, R, S) weight is this very parameter."""
    if not CONFIG["tc_backward"]:
        return False, None
    N, H, W, Ci = xP.shape
    _, Ho, Wo, Co = gP.shape
    _, _, R, S = weight.shape
    plan = tc_conv_plan("conv", Co, Ci, R, S, stride, pad, 0, Ho, Wo, min_ci=1)
    if plan is None or plan["layout"] != _LAYOUT_S2D or (Ho + 2 * pad - R) // stride + 1 != H or \
            (Wo + 2 * pad - S) // stride + 1 != W:
        return False, None
    dr, ds = _tc_taps(plan, R, S)
    return True, _tc_wgrad(xP, Ci, gP, plan, pad, PAD_ZERO, dr, ds, H, W, weight, _W_CONV_S2D)


class Prologue:
    """What sits between the producer of a conv's input and the conv itself, folded into the conv's operand
    preparation instead of running as separate passes: a normalisation layer (InstanceNorm2d(affine=False) when
    groups == 0, GroupNorm(groups, C, affine) otherwise; `stats` = per-(n, c) sums of the RAW input, usually taken by
    the producing GEMM's epilogue) and / or an activation (ReLU / LeakyReLU)."""

    def __init__(self, stats=None, norm=False, eps=1e-5, groups=0, gamma=None, beta=None, act=ACT_NONE, slope=0.0):
        self.stats, self.norm, self.eps, self.groups = stats, norm, eps, groups
        self.gamma, self.beta, self.act, self.slope = gamma, beta, act, slope

    def params(self, xh, lazy=False):
        """lazy: the caller's next launch is an operand preparation that can finalise the statistics itself"""
        if not self.norm:
            return None
        g = self.gamma.detach().contiguous() if self.gamma is not None else None
        b = self.beta.detach().contiguous() if self.beta is not None else None
        return _norm_params(xh, self.groups, g, b, self.eps, self.stats, lazy=lazy)


def _in_bwd(xh, g, prm, act, src=None):
    """InstanceNorm2d(affine=False) [+ReLU] backward: gradient w.r.t. the raw input xh given g, both (N, H, W, C).
    src = the weight of the convolution that produced xh (conv2d / conv_transpose2d leave it on the statistics they hand to
    the norm layer).  When that convolution's backward has recorded the dY operand it asks for (`_Prepared.rec`), the apply
    pass writes it as well - dsr_tc_prep_in_bwd - and registers it in _PREMADE for the _Prepared that convolution makes of
    the returned gradient; otherwise dsr_in_bwd_apply alone."""
    N, H, W, C = xh.shape
    sums2 = _zeros_f64(N * C * 2, g.device)
    _call("dsr_in_bwd_sums", _p(xh), _p(g), _p(prm), N, H * W, C, act, _p(sums2, torch.float64))
    gx = torch.empty_like(xh)
    spec = _DY_SPEC.get(src.data_ptr()) if (src is not None and CONFIG["fuse_bwd_prep"] and CONFIG["engine"] == "tc") else None
    if spec is not None and spec["shape"] == tuple(xh.shape) and g.shape == xh.shape and act in (ACT_NONE, ACT_RELU):
        layout, Cp, Ca, pad, _mode, dt, wa_min = key = spec["key"]
        cg = Ca // 8
        fits = dt == CONFIG["bwd_dtype"] and C % 8 == 0 and Cp >= C and cg <= 256 and cg & (cg - 1) == 0 and \
            (4 * _rup(C, 4) + C) * 4 <= 48 * 1024 and (Ca >= Cp if layout == _LAYOUT_NORMAL else Ca == 4 * Cp)
        if fits:
            Hq, Wq = H + 2 * pad, W + 2 * pad
            Ha, Wa = ((Hq + 1) // 2, (Wq + 1) // 2) if layout == _LAYOUT_S2D else (Hq, max(Wq, wa_min))
            ahi = torch.empty((N, Ha, Wa, Ca), device=xh.device, dtype=torch.bfloat16)
            alo = torch.empty((N, Ha, Wa, Ca), device=xh.device, dtype=torch.bfloat16) if (CONFIG["passes"] >= 2 and spec["need_lo"]) else None
            csum = _zeros_f64(C * CONFIG["csum_reps"], xh.device) if spec["csum"] else None
            if _lib.PROFILE is not None:
                _lib.PROFILE_META = dict(macs=0, shape=(N, H, W, C, Ca, layout, pad, 2, int(alo is not None)))
            _call("dsr_tc_prep_in_bwd", _p(xh), _p(g), _p(prm), _p(sums2, torch.float64), _p(gx), N, H, W, C, act, pad, layout, Cp,
                  _p(ahi, torch.bfloat16), _p(alo, torch.bfloat16), Ha, Wa, Ca, int(dt == "f16"), _p(csum, torch.float64),
                  CONFIG["csum_reps"] if csum is not None else 1)
            if len(_PREMADE) > 64:
                _PREMADE.clear()
            _PREMADE[gx.data_ptr()] = (gx, key, (ahi, alo, Ha, Wa), None, csum)
            return gx
    _call("dsr_in_bwd_apply", _p(xh), _p(g), _p(prm), _p(sums2, torch.float64), _p(gx), N, H * W, C, act)
    return gx


def _prologue_bwd(pro, prm, xh, gz):
    """gradient w.r.t. the RAW conv input given the gradient w.r.t. the normalised / activated operand"""
    if pro is None:
        return gz
    N, H, W, C = xh.shape
    if pro.norm:
        if pro.groups != 0:
            raise NotImplementedError("dsr_b200: GroupNorm backward is not on the main_network_best hot path")
        return _in_bwd(xh, gz, prm, pro.act, getattr(pro.stats, "src", None))
    if pro.act != ACT_NONE:
        gx = torch.empty_like(xh)
        _call("dsr_act_bwd", _p(xh), _p(gz), _p(gx), gz.numel(), pro.act, pro.slope)
        return gx
    return gz


class _Conv2d(Function):
    """nn.Conv2d with zeros / reflect / replicate padding.  networks.py:378-379,385,413-414,453,544;
    translation_network.py:472,478,495,563."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, pad_mode, act_out, want_stats, pro, skip_out=False):
        """skip_out: x is also returned as a third output (an alias) - a residual block routes its skip connection through
        it so that BOTH gradients of x arrive in this node's backward and are added inside its last pass."""
        xh = nhwc(x)
        N, H, W, Ci = xh.shape
        Co, Ci2, R, S = weight.shape
        if Ci2 != Ci:
            raise ValueError(f"conv2d: input has {Ci} channels, weight expects {Ci2}")
        Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
        b = bias.detach() if bias is not None else None
        stats = _new_stats(N, Co, x.device) if want_stats else None
        plan = tc_conv_plan("conv", Ci, Co, R, S, stride, pad, 0, H, W)
        prm = None
        out1 = Co == 1 and stride == 1 and R <= 9 and S <= 9 and CONFIG["out1"] and (pro is None or plan is not None)
        if pro is not None:
            if plan is None:
                raise RuntimeError("internal error: a fused prologue needs the tcgen05 path (see conv_fusable)")
            prm = pro.params(xh, lazy=not out1)
        ctx.xP = None
        if out1:
            # depth head: a bandwidth-bound reduction, on CUDA cores straight from the fp32 activation (csrc/conv_out1.cu)
            y = torch.empty((N, Ho, Wo, 1), device=x.device, dtype=torch.float32)
            w = weight.detach()
            _call("dsr_conv_out1", _p(xh), N, H, W, Ci, _p(prm), pro.act if pro else ACT_NONE, pro.slope if pro else 0.0,
                  _p(w if w.is_contiguous() else w.contiguous()), _p(b), R, S, pad, pad_mode, 0, act_out, _p(y))
        elif plan is not None:
            xin = _Prepared(xh) if pro is None else _Prepared(xh, prm, pro.act, pro.slope)
            y = _tc_conv_fwd(xin, weight, b, plan, stride, pad, pad_mode, act_out, Ho, Wo, stats=stats,
                             bwd_copy=ctx.needs_input_grad[1])
            if CONFIG["reuse_fwd_operand"] and ctx.needs_input_grad[1]:
                ctx.xP = xin
        else:
            xp, p = _explicit_pad(xh, pad, pad_mode)
            y = torch.empty((N, Ho, Wo, Co), device=x.device, dtype=torch.float32)
            wk = _pack(weight, 1)
            _lib.PROFILE_META = dict(macs=0, shape=("conv_simt", N, H, W, Ci, Co, R, stride))
            _call("dsr_conv_simt", _p(xp), _p(wk), _p(b), _p(y), N, xp.shape[1], xp.shape[2], Ci, Ho, Wo, Co, R, S,
                  stride, p, 0, act_out)
        ctx.cfg = (stride, pad, pad_mode, act_out, bias is not None)
        ctx.bias_ref, ctx.pro = bias, pro
        ctx.save_for_backward(xh, weight, y if act_out == ACT_TANH else None, _prm_ready(prm))
        return _with_stats(ctx, y, stats) + ((x,) if skip_out else (None,))

    @staticmethod
    def backward(ctx, gy, _gstats=None, gskip=None):
        xh, weight, y, prm = ctx.saved_tensors
        pro = ctx.pro
        stride, pad, pad_mode, act_out, has_bias = ctx.cfg
        N, H, W, Ci = xh.shape
        Co, _, R, S = weight.shape
        if gy is None:                                          # only the skip alias was used
            return gskip, None, None, None, None, None, None, None, None, None
        add = [nhwc(gskip) if (gskip is not None and ctx.needs_input_grad[0]) else None]    # consumed by whoever adds it
        g = nhwc(gy)
        _, Ho, Wo, _ = g.shape
        if act_out == ACT_TANH:
            g2 = torch.empty_like(g)
            _call("dsr_act_bwd", _p(y), _p(g), _p(g2), g.numel(), ACT_TANH, 0.0)
            g = g2
        gP = _Prepared(g, want_csum=has_bias, rec=weight)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gxp = _tc_conv_dgrad(gP, weight, stride, pad, pad_mode, H, W, add=add if pro is None else None)
            if gxp is None and stride == 1 and CONFIG["out1"] and R <= 9 and S <= 9 and (
                    (Co == 1 and Ci % 16 == 0 and R * S * Ci * 4 <= 180 * 1024) or Ci == 1):
                # one-channel layers (csrc/conv_out1.cu): a conv with ONE output channel has an outer-product data gradient;
                # a conv with ONE input channel has a many-to-one convolution (flipped taps) as its data gradient
                explicit = pad > 0 and pad_mode != PAD_ZERO
                Hx, Wx = (H + 2 * pad, W + 2 * pad) if explicit else (H, W)
                w = weight.detach()
                gxp = torch.empty((N, Hx, Wx, Ci), device=g.device, dtype=torch.float32)
                if Co == 1:
                    _call("dsr_conv1_dgrad", _p(g), N, Ho, Wo, _p(w if w.is_contiguous() else w.contiguous()), Ci, R, S,
                          0 if explicit else pad, _p(gxp), Hx, Wx)
                else:
                    wf = torch.flip(w[:, 0], dims=(1, 2)).contiguous()                  # [Co][R][S], taps flipped
                    _call("dsr_conv_out1", _p(g), N, Ho, Wo, Co, None, ACT_NONE, 0.0, _p(wf), None, R, S,
                          R - 1 - (0 if explicit else pad), PAD_ZERO, 0, ACT_NONE, _p(gxp))
                if explicit:
                    gxh = torch.empty_like(xh)
                    _call("dsr_pad2d_bwd", _p(gxp), _p(gxh), N, H, W, Ci, pad, pad_mode)
                    gxp = gxh
            if gxp is None:
                xp, p = _explicit_pad(xh, pad, pad_mode)            # recomputed, not stored
                Hp, Wp = xp.shape[1], xp.shape[2]
                wk = _pack(weight, 0)                               # [(r,s,co)][ci]
                gxp = torch.empty((N, Hp, Wp, Ci), device=g.device, dtype=torch.float32)
                _lib.PROFILE_META = dict(macs=0, shape=("conv_simt", N, H, W, Ci, Co, R, stride))
                _call("dsr_conv_simt", _p(g), _p(wk), None, _p(gxp), N, Ho, Wo, Co, Hp, Wp, Ci, R, S, stride, p, 1, ACT_NONE)
                if xp is not xh:
                    gxh = torch.empty_like(xh)
                    _call("dsr_pad2d_bwd", _p(gxp), _p(gxh), N, H, W, Ci, pad, pad_mode)
                    gxp = gxh
            gx = nchw(_prologue_bwd(pro, prm, xh, gxp))
            if add[0] is not None:
                gx = gx + gskip                                 # (a route without the fused add)
        need_w, need_b = ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]
        if need_w and gx is None:
            _tc_wgrad_m_operand(gP, Co)           # no data gradient ran: make the dY operand on the main stream
        with _on_side((need_w or need_b) and _use_side(weight if need_w else None, ctx.bias_ref if need_b else None), g.device):
            if need_w:
                xP = ctx.xP or (_Prepared(xh) if pro is None else _Prepared(xh, prm, pro.act, pro.slope))
                done, gw = _tc_conv_wgrad(xP, gP, weight, stride, pad, pad_mode)
                if not done:
                    xp, p = _explicit_pad(xh, pad, pad_mode)
                    dwk = torch.empty(R * S * Ci * Co, device=g.device, dtype=torch.float32)
                    _lib.PROFILE_META = dict(macs=0, shape=("wgrad_simt", N, H, W, Ci, Co, R, stride))
                    _call("dsr_wgrad_simt", _p(xp), _p(g), _p(dwk), N, xp.shape[1], xp.shape[2], Ci, Ho, Wo, Co, R, S,
                          stride, p)
                    gw = _weight_grad(dwk, weight, 1)
                    _keep(xp, xh)
            if need_b:
                gb = _bias_grad(g, Co, ctx.bias_ref, gP)
            _keep(g, gP.csum, prm)
        return gx, gw, gb, None, None, None, None, None, None, None


def _new_stats(N, C, device):
    """per-(n, c) (sum, sum of squares) accumulator the GEMM epilogue adds into (InstanceNorm / GroupNorm statistics)"""
    st = _zeros_f64(N * C * 2, device)              # the step's pre-zeroed pool: no fill kernel per layer (137 per step before)
    st.filled = False
    return st


def _with_stats(ctx, y, stats):
    """Function outputs: the activation and (non-differentiable) its channel statistics, taken by the GEMM epilogue
    when the kernel supports it and by one channel_sums pass otherwise."""
    if stats is None:
        return nchw(y), None
    if not stats.filled:
        N, H, W, C = y.shape
        _call("dsr_channel_sums", _p(y), N, H * W, C, _p(stats, torch.float64))
    ctx.mark_non_differentiable(stats)
    return nchw(y), stats


def _explicit_pad(xh, pad, pad_mode):
    """materialise a non-zero padding mode for the CUDA-core path; zero padding stays implicit."""
    if pad == 0 or pad_mode == PAD_ZERO:
        return xh, pad
    N, H, W, C = xh.shape
    xp = torch.empty((N, H + 2 * pad, W + 2 * pad, C), device=xh.device, dtype=torch.float32)
    _call("dsr_pad2d_fwd", _p(xh), _p(xp), N, H, W, C, pad, pad_mode)
    return xp, 0


class _ConvTranspose2d(Function):
    """nn.ConvTranspose2d.  networks.py:406,553,605,612; translation_network.py:508."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, opad, act_out, want_stats, pro):
        xh = nhwc(x)
        N, H, W, Ci = xh.shape
        Ci2, Co, R, S = weight.shape
        if Ci2 != Ci:
            raise ValueError(f"conv_transpose2d: input has {Ci} channels, weight expects {Ci2}")
        Ho, Wo = (H - 1) * stride - 2 * pad + R + opad, (W - 1) * stride - 2 * pad + S + opad
        b = bias.detach() if bias is not None else None
        stats = _new_stats(N, Co, x.device) if want_stats else None
        plan = tc_conv_plan("convT", Ci, Co, R, S, stride, pad, opad, H, W)
        prm = None
        out1 = Co == 1 and R == 4 and S == 4 and stride == 2 and pad == 1 and opad == 0 and CONFIG["out1"] and \
            (pro is None or plan is not None)
        if pro is not None:
            if plan is None:
                raise RuntimeError("internal error: a fused prologue needs the tcgen05 path (see conv_fusable)")
            prm = pro.params(xh, lazy=not out1)
        ctx.xP = None
        if out1:
            y = torch.empty((N, Ho, Wo, 1), device=x.device, dtype=torch.float32)
            w = weight.detach()
            _call("dsr_conv_out1", _p(xh), N, H, W, Ci, _p(prm), pro.act if pro else ACT_NONE, pro.slope if pro else 0.0,
                  _p(w if w.is_contiguous() else w.contiguous()), _p(b), 4, 4, 1, PAD_ZERO, 1, act_out, _p(y))
        elif plan is not None:
            xin = _Prepared(xh) if pro is None else _Prepared(xh, prm, pro.act, pro.slope)
            y = _tc_convT_fwd(xin, weight, b, plan, pad, act_out, Ho, Wo, stats=stats, bwd_copy=ctx.needs_input_grad[1])
            if CONFIG["reuse_fwd_operand"] and ctx.needs_input_grad[1]:
                ctx.xP = xin        # its zero-haloed copy of x is the M operand of the weight-gradient GEMM
        else:
            y = torch.empty((N, Ho, Wo, Co), device=x.device, dtype=torch.float32)
            wk = _pack(weight, 0)                       # [(r,s,ci)][co]
            _lib.PROFILE_META = dict(macs=0, shape=("conv_simt", N, H, W, Ci, Co, R, stride))
            _call("dsr_conv_simt", _p(xh), _p(wk), _p(b), _p(y), N, H, W, Ci, Ho, Wo, Co, R, S, stride, pad, 1, act_out)
        ctx.cfg = (stride, pad, act_out, bias is not None)
        ctx.bias_ref, ctx.pro = bias, pro
        ctx.save_for_backward(xh, weight, y if act_out == ACT_TANH else None, _prm_ready(prm))
        return _with_stats(ctx, y, stats)

    @staticmethod
    def backward(ctx, gy, _gstats=None):
        xh, weight, y, prm = ctx.saved_tensors
        pro = ctx.pro
        stride, pad, act_out, has_bias = ctx.cfg
        N, H, W, Ci = xh.shape
        _, Co, R, S = weight.shape
        g = nhwc(gy)
        _, Ho, Wo, _ = g.shape
        if act_out == ACT_TANH:
            g2 = torch.empty_like(g)
            _call("dsr_act_bwd", _p(y), _p(g), _p(g2), g.numel(), ACT_TANH, 0.0)
            g = g2
        gP = _Prepared(g, want_csum=has_bias, rec=weight)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gxh = _tc_convT_dgrad(gP, weight, stride, pad, H, W)
            if gxh is None:
                wk = _pack(weight, 1)                   # [(r,s,co)][ci]
                gxh = torch.empty((N, H, W, Ci), device=g.device, dtype=torch.float32)
                _lib.PROFILE_META = dict(macs=0, shape=("conv_simt", N, H, W, Ci, Co, R, stride))
                _call("dsr_conv_simt", _p(g), _p(wk), None, _p(gxh), N, Ho, Wo, Co, H, W, Ci, R, S, stride, pad, 0, ACT_NONE)
            gx = nchw(_prologue_bwd(pro, prm, xh, gxh))
        need_w, need_b = ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]
        with _on_side((need_w or need_b) and _use_side(weight if need_w else None, ctx.bias_ref if need_b else None), g.device):
            if need_w:
                xP = ctx.xP or (_Prepared(xh) if pro is None else _Prepared(xh, prm, pro.act, pro.slope))
                done, gw = _tc_convT_wgrad(xP, gP, weight, stride, pad)
                if not done:
                    dwk = torch.empty(R * S * Co * Ci, device=g.device, dtype=torch.float32)
                    _lib.PROFILE_META = dict(macs=0, shape=("wgrad_simt", N, H, W, Ci, Co, R, stride))
                    _call("dsr_wgrad_simt", _p(g), _p(xh), _p(dwk), N, Ho, Wo, Co, H, W, Ci, R, S, stride, pad)
                    gw = _weight_grad(dwk, weight, 1)
            if need_b:
                gb = _bias_grad(g, Co, ctx.bias_ref, gP)
            _keep(g, gP.csum, prm, xh)
        return gx, gw, gb, None, None, None, None, None, None


class LazyCat:
    """torch.cat(parts, dim=1) that has not been materialised yet: a Conv2d that receives one runs `cat_conv2d`,
    which concatenates inside the op and - in the backward pass - computes the data gradient only for the channel
    range of the parts that need one (main_model.py:305-306: of the 261 Task input channels only the 128 of
    Depth_f's features do)."""

    def __init__(self, parts):
        self.parts = list(parts)
        self.shape = (parts[0].shape[0], sum(p.shape[1] for p in parts)) + tuple(parts[0].shape[2:])


class _CatConv2d(Function):
    """nn.Conv2d applied to torch.cat(xs, 1).  networks.py:544 on main_model.py:305-306."""

    @staticmethod
    def forward(ctx, weight, bias, stride, pad, pad_mode, act_out, *xs):
        hs = [nhwc(x) for x in xs]
        N, H, W, _ = hs[0].shape
        Cs = [h.shape[3] for h in hs]
        Ci = sum(Cs)
        Co, Ci2, R, S = weight.shape
        if Ci2 != Ci:
            raise ValueError(f"conv2d: input has {Ci} channels, weight expects {Ci2}")
        plan = tc_conv_plan("conv", Ci, Co, R, S, stride, pad, 0, H, W)
        if plan is None:
            raise RuntimeError("internal error: cat_conv2d needs the tcgen05 path")
        Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
        b = bias.detach() if bias is not None else None
        if len(hs) <= 4 and CONFIG["fused_cat"]:
            xin = _PreparedCat(hs)                      # the concatenation is never written: the operand prep reads the parts
        else:
            xh = torch.empty((N, H, W, Ci), device=hs[0].device, dtype=torch.float32)
            off = 0
            for h, c in zip(hs, Cs):
                _call("dsr_copy_channels", _p(h), c, 0, _p(xh), Ci, off, c, N * H * W, 0)
                off += c
            xin = _Prepared(xh)
            hs = [xh]
        y = _tc_conv_fwd(xin, weight, b, plan, stride, pad, pad_mode, act_out, Ho, Wo, bwd_copy=ctx.needs_input_grad[0])
        ctx.xP = xin if (CONFIG["reuse_fwd_operand"] and ctx.needs_input_grad[0]) else None
        ctx.cfg = (stride, pad, pad_mode, act_out, bias is not None, Cs, (N, H, W, Ci))
        ctx.bias_ref = bias
        ctx.n_src = len(hs)
        ctx.save_for_backward(weight, y if act_out == ACT_TANH else None, *hs)
        return nchw(y)

    @staticmethod
    def backward(ctx, gy):
        weight, y, *hs = ctx.saved_tensors
        stride, pad, pad_mode, act_out, has_bias, Cs, (N, H, W, Ci) = ctx.cfg
        Co, _, R, S = weight.shape
        g = nhwc(gy)
        if act_out == ACT_TANH:
            g2 = torch.empty_like(g)
            _call("dsr_act_bwd", _p(y), _p(g), _p(g2), g.numel(), ACT_TANH, 0.0)
            g = g2
        gP = _Prepared(g, want_csum=has_bias, rec=weight)
        needs = ctx.needs_input_grad[6:]
        offs = [sum(Cs[:i]) for i in range(len(Cs))]
        gxs = [None] * len(Cs)
        idx = [i for i, nd in enumerate(needs) if nd]
        if idx:
            c0, c1 = offs[idx[0]], offs[idx[-1]] + Cs[idx[-1]]
            w_slice = weight.detach()[:, c0:c1].contiguous()
            gxr = _tc_conv_dgrad(gP, w_slice, stride, pad, pad_mode, H, W)
            if gxr is None:
                raise RuntimeError("internal error: cat_conv2d data gradient is not covered by the tcgen05 path")
            for i in idx:
                if offs[i] == c0 and Cs[i] == c1 - c0:
                    gxs[i] = nchw(gxr)
                else:
                    gi = torch.empty((N, H, W, Cs[i]), device=g.device, dtype=torch.float32)
                    _call("dsr_copy_channels", _p(gxr), c1 - c0, offs[i] - c0, _p(gi), Cs[i], 0, Cs[i], N * H * W, 0)
                    gxs[i] = nchw(gi)
        gw = gb = None
        need_w, need_b = ctx.needs_input_grad[0], has_bias and ctx.needs_input_grad[1]
        if need_w and not idx:
            _tc_wgrad_m_operand(gP, Co)
        with _on_side((need_w or need_b) and _use_side(weight if need_w else None, ctx.bias_ref if need_b else None), g.device):
            if need_w:
                xP = ctx.xP or (_PreparedCat(hs) if len(hs) == len(Cs) and len(Cs) > 1 else _Prepared(hs[0]))
                done, gw = _tc_conv_wgrad(xP, gP, weight, stride, pad, pad_mode)
                if not done:
                    raise RuntimeError("internal error: cat_conv2d weight gradient is not covered by the tcgen05 path")
            if need_b:
                gb = _bias_grad(g, Co, ctx.bias_ref, gP)
            _keep(g, gP.csum, *hs)
        return (gw, gb, None, None, None, None) + tuple(gxs)


def cat_conv2d(parts, weight, bias=None, stride=1, padding=0, act_out=ACT_NONE, pad_mode=PAD_ZERO):
    pad_mode = PAD_MODES[pad_mode] if isinstance(pad_mode, str) else pad_mode
    _FWD["trainable"] = torch.is_grad_enabled() and (weight.requires_grad or any(p.requires_grad for p in parts))
    return _CatConv2d.apply(weight, bias, stride, padding, pad_mode, act_out, *parts)


def conv_fusable(kind, x, weight, stride, padding, output_padding=0):
    """True when this layer runs on the tcgen05 path, i.e. can take a fused Prologue"""
    N, C, H, W = x.shape
    if kind == "conv":
        Co, Ci, R, S = weight.shape
        return tc_conv_plan("conv", Ci, Co, R, S, stride, padding, 0, H, W) is not None
    Ci, Co, R, S = weight.shape
    return tc_conv_plan("convT", Ci, Co, R, S, stride, padding, output_padding, H, W) is not None


def conv2d(x, weight, bias=None, stride=1, padding=0, act_out=ACT_NONE, pad_mode=PAD_ZERO, want_stats=False, pro=None,
           skip_out=False):
    """-> y, or (y, stats) with want_stats: stats = float64 [N*Cout*2] per-(n, c) (sum, sum of squares) of y for the
    normalisation layer that follows (instance_norm / group_norm take it through their `stats` argument)."""
    pad_mode = PAD_MODES[pad_mode] if isinstance(pad_mode, str) else pad_mode
    _FWD["trainable"] = torch.is_grad_enabled() and (weight.requires_grad or x.requires_grad)
    y, st, sk = _Conv2d.apply(x, weight, bias, stride, padding, pad_mode, act_out, want_stats, pro, skip_out)
    if st is not None:
        st.src = weight                                   # the norm layer that takes these statistics: see _in_bwd
    if skip_out:                                          # (..., alias of x: see _Conv2d.forward)
        return (y, st, sk) if want_stats else (y, sk)
    return (y, st) if want_stats else y


def conv_transpose2d(x, weight, bias=None, stride=1, padding=0, output_padding=0, act_out=ACT_NONE, want_stats=False,
                     pro=None):
    _FWD["trainable"] = torch.is_grad_enabled() and (weight.requires_grad or x.requires_grad)
    y, st = _ConvTranspose2d.apply(x, weight, bias, stride, padding, output_padding, act_out, want_stats, pro)
    if st is not None:
        st.src = weight
    return (y, st) if want_stats else y


class _Pad2d(Function):
    @staticmethod
    def forward(ctx, x, pad, mode):
        xh = nhwc(x)
        N, H, W, C = xh.shape
        y = torch.empty((N, H + 2 * pad, W + 2 * pad, C), device=x.device, dtype=torch.float32)
        _call("dsr_pad2d_fwd", _p(xh), _p(y), N, H, W, C, pad, mode)
        ctx.cfg = (N, H, W, C, pad, mode)
        return nchw(y)

    @staticmethod
    def backward(ctx, gy):
        N, H, W, C, pad, mode = ctx.cfg
        g = nhwc(gy)
        gx = torch.empty((N, H, W, C), device=g.device, dtype=torch.float32)
        _call("dsr_pad2d_bwd", _p(g), _p(gx), N, H, W, C, pad, mode)
        return nchw(gx), None, None


def pad2d(x, pad, mode):
    return _Pad2d.apply(x, pad, PAD_MODES[mode] if isinstance(mode, str) else mode)


# ------------------------------------------------------------------------------------------------
# normalisation / activation / concat
# ------------------------------------------------------------------------------------------------
def _norm_params(xh, groups, gamma, beta, eps, sums=None, lazy=False):
    """(mean, scale, shift) per (n, c) from the channel sums.  lazy (and CONFIG['fold_finalize']): the finalisation is left
    to the first consumer - dsr_tc_prep_fin / dsr_norm_apply_fwd_fin derive the constants inside their own launch and write
    `prm` for the backward pass; the pending work rides on the tensor as `prm._fin` until _take_fin / _prm_ready claims it."""
    N, H, W, C = xh.shape
    if sums is None:
        sums = _zeros_f64(N * C * 2, xh.device)
        _call("dsr_channel_sums", _p(xh), N, H * W, C, _p(sums, torch.float64))
    elif sums.numel() != N * C * 2:
        raise ValueError("norm statistics do not match the activation shape")
    prm = torch.empty(3 * N * C, device=xh.device, dtype=torch.float32)
    if lazy and CONFIG["fold_finalize"]:
        prm._fin = (sums, N, C, H * W, groups, gamma, beta, eps)
        return prm
    _call("dsr_norm_finalize", _p(sums, torch.float64), N, C, H * W, groups, _p(gamma), _p(beta), eps, _p(prm))
    return prm


def _take_fin(prm):
    """claim the pending finalisation of `prm` (None when there is none): the caller's launch must perform it"""
    fin = getattr(prm, "_fin", None) if prm is not None else None
    if fin is not None:
        prm._fin = None
    return fin


def _prm_ready(prm):
    """make sure `prm` holds finalised constants (a consumer without the folded route, or nobody claimed the work)"""
    fin = _take_fin(prm)
    if fin is not None:
        sums, N, C, P, groups, gamma, beta, eps = fin
        _call("dsr_norm_finalize", _p(sums, torch.float64), N, C, P, groups, _p(gamma), _p(beta), eps, _p(prm))
    return prm


def _norm_apply_fwd(xh, prm, rh, y, act):
    """y = act((x - mean) * scale + shift) (+ residual); finalises the statistics in the same launch when they are pending"""
    N, H, W, C = xh.shape
    fin = _take_fin(prm)
    if fin is not None:
        sums, _, _, P, groups, gamma, beta, eps = fin
        _call("dsr_norm_apply_fwd_fin", _p(xh), _p(sums, torch.float64), groups, _p(gamma), _p(beta), eps, _p(prm), _p(rh), _p(y),
              N, H * W, C, act)
    else:
        _call("dsr_norm_apply_fwd", _p(xh), _p(prm), _p(rh), _p(y), N, H * W, C, act)


class _InstanceNorm(Function):
    """InstanceNorm2d(affine=False) [+ ReLU] [+ residual add].  networks.py:30, :380-381, :480."""

    @staticmethod
    def forward(ctx, x, eps, act, residual, stats, hint=None):
        xh = nhwc(x)
        N, H, W, C = xh.shape
        rh = nhwc(residual) if residual is not None else None
        if hint is not None and stats is not None and stats.numel() == N * C * 2 and act in (ACT_NONE, ACT_RELU):
            y, prm = _norm_res_prep(xh, stats, 0, None, None, eps, act, rh, hint)
        else:
            prm = _norm_params(xh, 0, None, None, eps, stats, lazy=True)
            y = torch.empty_like(xh)
            _norm_apply_fwd(xh, prm, rh, y, act)
        ctx.act = act
        ctx.has_res = residual is not None
        ctx.src = getattr(stats, "src", None)               # weight of the convolution that produced x (see _in_bwd)
        ctx.save_for_backward(xh, prm)
        return nchw(y)

    @staticmethod
    def backward(ctx, gy):
        xh, prm = ctx.saved_tensors
        g = nhwc(gy)
        gx = None
        if ctx.needs_input_grad[0]:
            gx = nchw(_in_bwd(xh, g, prm, ctx.act, ctx.src))
        gres = gy if (ctx.has_res and ctx.needs_input_grad[3]) else None
        return gx, None, None, gres, None, None


def instance_norm(x, eps=1e-5, act=ACT_NONE, residual=None, stats=None, hint=None):
    """hint (ops.operand_hint of the convolution that reads the result): the same launch also writes that operand"""
    return _InstanceNorm.apply(x, eps, act, residual, stats, hint)


def _param_grad(vals, param):
    """gradient of a 1-D parameter: accumulated into the gradient arena when the parameter lives there, returned otherwise"""
    tgt = DIRECT_GRADS.get(param.data_ptr())
    if tgt is not None:
        tgt.add_(vals)
        _grad_ready(param)
        return None
    return vals


class _GroupNorm(Function):
    """GroupNorm(groups, C, affine=True) [+ReLU] [+residual] with its backward (csrc/nn.cu gn_bwd_*): the generators of the
    translation block train through it.  translation_network.py:46, :472-483, :563-574."""

    @staticmethod
    def forward(ctx, x, weight, bias, groups, eps, act, residual, stats, hint=None):
        xh = nhwc(x)
        N, H, W, C = xh.shape
        w, b = weight.detach().contiguous(), bias.detach().contiguous()
        if stats is None:
            stats = _zeros_f64(N * C * 2, xh.device)
            _call("dsr_channel_sums", _p(xh), N, H * W, C, _p(stats, torch.float64))
        rh = nhwc(residual) if residual is not None else None
        if hint is not None and stats.numel() == N * C * 2 and act in (ACT_NONE, ACT_RELU):
            y, _ = _norm_res_prep(xh, stats, groups, w, b, eps, act, rh, hint)
        else:
            prm = _norm_params(xh, groups, w, b, eps, stats, lazy=True)
            y = torch.empty_like(xh)
            _norm_apply_fwd(xh, prm, rh, y, act)
        if any(ctx.needs_input_grad):
            hat = torch.empty(3 * N * C, device=xh.device, dtype=torch.float32)         # (mean, rstd, 0): gamma = beta = NULL
            _call("dsr_norm_finalize", _p(stats, torch.float64), N, C, H * W, groups, None, None, eps, _p(hat))
            ctx.save_for_backward(xh, hat, weight, bias)
        ctx.cfg = (groups, act, residual is not None)
        return nchw(y)

    @staticmethod
    def backward(ctx, gy):
        xh, hat, weight, bias = ctx.saved_tensors
        groups, act, has_res = ctx.cfg
        N, H, W, C = xh.shape
        g = nhwc(gy)
        w, b = weight.detach().contiguous(), bias.detach().contiguous()
        sums2 = _zeros_f64(N * C * 2, g.device)
        _call("dsr_gn_bwd_sums", _p(xh), _p(g), _p(hat), _p(w), _p(b), N, H * W, C, act, _p(sums2, torch.float64))
        coef = torch.empty(N * C * 2, device=g.device, dtype=torch.float32)
        dgamma = torch.empty(C, device=g.device, dtype=torch.float32)
        dbeta = torch.empty(C, device=g.device, dtype=torch.float32)
        _call("dsr_gn_bwd_finalize", _p(sums2, torch.float64), _p(w), N, C, groups, H * W, _p(coef), _p(dgamma), _p(dbeta), 0)
        gx = None
        if ctx.needs_input_grad[0]:
            gxh = torch.empty_like(xh)
            _call("dsr_gn_bwd_apply", _p(xh), _p(g), _p(hat), _p(w), _p(b), _p(coef), _p(gxh), N, H * W, C, act)
            gx = nchw(gxh)
        gw = _param_grad(dgamma, weight) if ctx.needs_input_grad[1] else None
        gb = _param_grad(dbeta, bias) if ctx.needs_input_grad[2] else None
        gres = gy if (has_res and ctx.needs_input_grad[6]) else None
        return gx, gw, gb, None, None, None, gres, None, None


def group_norm(x, groups, weight, bias, eps=1e-5, act=ACT_NONE, residual=None, stats=None, hint=None):
    """GroupNorm(groups, C, affine=True) [+ReLU] [+residual].  translation_network.py:46."""
    return _GroupNorm.apply(x, weight, bias, groups, eps, act, residual, stats, hint)


class _Act(Function):
    @staticmethod
    def forward(ctx, x, kind, slope):
        xh = nhwc(x)
        y = torch.empty_like(xh)
        _call("dsr_act_fwd", _p(xh), _p(y), xh.numel(), kind, slope)
        ctx.cfg = (kind, slope)
        ctx.save_for_backward(y if kind == ACT_TANH else xh)
        return nchw(y)

    @staticmethod
    def backward(ctx, gy):
        (ref,) = ctx.saved_tensors
        kind, slope = ctx.cfg
        g = nhwc(gy)
        gx = torch.empty_like(g)
        _call("dsr_act_bwd", _p(ref), _p(g), _p(gx), g.numel(), kind, slope)
        return nchw(gx), None, None


def relu(x):
    return _Act.apply(x, ACT_RELU, 0.0)


def leaky_relu(x, slope=0.2):
    return _Act.apply(x, ACT_LRELU, slope)


def tanh(x):
    return _Act.apply(x, ACT_TANH, 0.0)


class _Cat(Function):
    """torch.cat(dim=1) of NHWC-backed tensors.  networks.py:629, main_model.py:302-306."""

    @staticmethod
    def forward(ctx, *xs):
        hs = [nhwc(x) for x in xs]
        N, H, W, _ = hs[0].shape
        Cs = [h.shape[3] for h in hs]
        y = torch.empty((N, H, W, sum(Cs)), device=hs[0].device, dtype=torch.float32)
        off = 0
        for h, c in zip(hs, Cs):
            _call("dsr_copy_channels", _p(h), c, 0, _p(y), sum(Cs), off, c, N * H * W, 0)
            off += c
        ctx.Cs = Cs
        return nchw(y)

    @staticmethod
    def backward(ctx, gy):
        g = nhwc(gy)
        N, H, W, Ct = g.shape
        outs, off = [], 0
        for i, c in enumerate(ctx.Cs):
            if ctx.needs_input_grad[i]:
                gx = torch.empty((N, H, W, c), device=g.device, dtype=torch.float32)
                _call("dsr_copy_channels", _p(g), Ct, off, _p(gx), c, 0, c, N * H * W, 0)
                outs.append(nchw(gx))
            else:
                outs.append(None)
            off += c
        return tuple(outs)


def cat(xs):
    return _Cat.apply(*xs)


# ------------------------------------------------------------------------------------------------
# loss stack (NCHW planes)
# ------------------------------------------------------------------------------------------------
def hole_valid_masks(depth, border=-0.97):
    """-> (hole, valid) float32 {0,1}.  main_model.py:208-230."""
    d = planes(depth.detach())
    B, C, H, W = d.shape
    hole, valid = torch.empty_like(d), torch.empty_like(d)
    _call("dsr_hole_valid_masks", _p(d), B * C, H, W, border, _p(hole), _p(valid))
    return hole, valid


def below_mask(depth, thr):
    """-> float32 {0,1}: 0 where depth < thr.  I2D_model.py:223,226."""
    d = planes(depth.detach())
    out = torch.empty_like(d)
    _call("dsr_below_mask", _p(d), d.numel(), float(thr), _p(out))
    return out


def rect_holes(valid, depth, rects_dev, counts_dev, max_rects, extra_border=float("-inf")):
    """-> (gt_mask uint8, masked depth, extra-hole mask).  main_model.py:257-298, :354-357, :396."""
    v, d = planes(valid), planes(depth.detach())
    B, _, H, W = d.shape
    gt = torch.empty((B, 1, H, W), device=d.device, dtype=torch.uint8)
    masked, extra = torch.empty_like(d), torch.empty_like(d)
    _call("dsr_rect_holes", _p(v), _p(d), _p(rects_dev, torch.int32), _p(counts_dev, torch.int32), max_rects,
          B, H, W, extra_border, _p(gt, torch.uint8), _p(masked), _p(extra))
    return gt, masked, extra


class _NormalsOld(Function):
    @staticmethod
    def forward(ctx, depth, scale):
        d = planes(depth)
        B, _, H, W = d.shape
        out = torch.empty((B, 3, H, W), device=d.device, dtype=torch.float32)
        _call("dsr_normals_old_fwd", _p(d), B, H, W, scale, _p(out))
        ctx.scale = scale
        ctx.save_for_backward(d)
        return out

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        B, _, H, W = d.shape
        gd = torch.empty_like(d)
        _call("dsr_normals_old_bwd", _p(d), _p(planes(g)), B, H, W, ctx.scale, _p(gd))
        return gd, None


class _NormalsNew(Function):
    @staticmethod
    def forward(ctx, depth, cams):
        d = planes(depth)
        B, _, H, W = d.shape
        out = torch.empty((B, 3, H, W), device=d.device, dtype=torch.float32)
        _call("dsr_normals_new_fwd", _p(d), _p(cams, torch.float64), B, H, W, _p(out))
        ctx.save_for_backward(d, cams)
        return out

    @staticmethod
    def backward(ctx, g):
        d, cams = ctx.saved_tensors
        B, _, H, W = d.shape
        gd = torch.empty_like(d)
        _call("dsr_normals_new_bwd", _p(d), _p(planes(g)), _p(cams, torch.float64), B, H, W, _p(gd))
        return gd, None


def normals_old(depth, scale=1.0):
    if depth.dtype != torch.float32:
        raise TypeError("Input shold be torch.float32")      # norms.py:205-208
    return _NormalsOld.apply(depth, float(scale))


def normals_new(depth, cams):
    return _NormalsNew.apply(depth, cams)


class _TV(Function):
    @staticmethod
    def forward(ctx, x):
        xp = planes(x)
        B, C, H, W = xp.shape
        acc = _zeros_f64(1, xp.device)
        _call("dsr_tv_fwd", _p(xp), B * C, H, W, _p(acc, torch.float64))
        out = torch.empty((), device=xp.device, dtype=torch.float32)
        _call("dsr_cvt_f64_f32", _p(acc, torch.float64), 1, _p(out), 1, 1.0, 0)
        ctx.save_for_backward(xp)
        return out

    @staticmethod
    def backward(ctx, g):
        (xp,) = ctx.saved_tensors
        B, C, H, W = xp.shape
        gx = torch.empty_like(xp)
        _call("dsr_tv_bwd", _p(xp), B * C, H, W, _p(g.contiguous()), 1.0, _p(gx))
        return gx


def tv_loss(x):
    return _TV.apply(x)


class _MaskedDiff(Function):
    """(mean |a*m - b*m|, mean (a*m - b*m)^2) over all elements; gradient flows to b only."""

    @staticmethod
    def forward(ctx, a, b, m1, m2):
        a, b, m1 = planes(a.detach()), planes(b), planes(m1)
        m2 = planes(m2) if m2 is not None else None
        B, C, H, W = b.shape
        acc = _zeros_f64(2, b.device)
        _call("dsr_masked_diff_fwd", _p(a), _p(b), _p(m1), _p(m2), B, C, H * W, _p(acc, torch.float64))
        out = torch.empty(2, device=b.device, dtype=torch.float32)
        _call("dsr_cvt_f64_f32", _p(acc, torch.float64), 1, _p(out), 2, 1.0 / b.numel(), 0)
        ctx.save_for_backward(a, b, m1, m2)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b, m1, m2 = ctx.saved_tensors
        B, C, H, W = b.shape
        gb = torch.empty_like(b)
        inv = 1.0 / b.numel()
        g = g.contiguous()
        _call("dsr_masked_diff_bwd", _p(a), _p(b), _p(m1), _p(m2), B, C, H * W, _p(g), _p(g) + 4, inv, inv, _p(gb))
        return None, gb, None, None


def masked_l1_l2(a, b, m1, m2=None):
    """-> float32[2] = (L1 mean, MSE mean) of (a*m1*m2, b*m1*m2)."""
    return _MaskedDiff.apply(a, b, m1, m2)


def masked_sums(d, p, m):
    """-> float64[3] = (sum d*m, sum p*m, sum |d*m - p*m|).  main_model.py:308-318."""
    d, p, m = planes(d.detach()), planes(p.detach()), planes(m)
    out = torch.zeros(3, device=d.device, dtype=torch.float64)      # read lazily by the caller: not from the per-step pool
    _call("dsr_masked_sums", _p(d), _p(p), _p(m), d.numel(), _p(out, torch.float64))
    return out


class _Smooth(Function):
    """get_smooth_weight(depth, image, 3).  main_model.py:22-73."""

    @staticmethod
    def forward(ctx, depth, image, num_scales):
        d, im = planes(depth), planes(image.detach())
        B, _, H, W = d.shape
        C = im.shape[1]
        levels = []                       # coarsest first, like scale_pyramid(...).reverse()
        for e in range(num_scales - 1, -1, -1):
            if e == 0:
                levels.append((d, im, H, W))
            else:
                nh, nw = H // 2 ** e, W // 2 ** e
                dl = torch.empty((B, 1, nh, nw), device=d.device, dtype=torch.float32)
                il = torch.empty((B, C, nh, nw), device=d.device, dtype=torch.float32)
                _call("dsr_bilinear_ac_fwd", _p(d), B, H, W, nh, nw, _p(dl))
                _call("dsr_bilinear_ac_fwd", _p(im), B * C, H, W, nh, nw, _p(il))
                levels.append((dl, il, nh, nw))
        acc = _zeros_f64(2 * num_scales, d.device)
        out = torch.zeros((), device=d.device, dtype=torch.float32)
        coefs = []
        for i, (dl, il, h, w) in enumerate(levels):
            _call("dsr_smooth_level_fwd", _p(dl), _p(il), B, C, h, w, _p(acc[2 * i:], torch.float64))
            cx = 1.0 / (B * (h - 1) * w) / 2 ** i
            cy = 1.0 / (B * h * (w - 1)) / 2 ** i
            coefs.append((cx, cy))
            _call("dsr_cvt_f64_f32", _p(acc[2 * i:], torch.float64), 1, _p(out), 1, cx, 1)
            _call("dsr_cvt_f64_f32", _p(acc[2 * i + 1:], torch.float64), 1, _p(out), 1, cy, 1)
        ctx.levels, ctx.coefs, ctx.shape = levels, coefs, (B, C, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        B, C, H, W = ctx.shape
        g = g.contiguous()
        d_full = ctx.levels[-1][0]
        gd = torch.empty_like(d_full)
        cx, cy = ctx.coefs[-1]
        _call("dsr_smooth_level_bwd", _p(d_full), _p(ctx.levels[-1][1]), B, C, H, W, _p(g), cx, cy, _p(gd), 0)
        for (dl, il, h, w), (cx, cy) in zip(ctx.levels[:-1], ctx.coefs[:-1]):
            gl = torch.empty_like(dl)
            _call("dsr_smooth_level_bwd", _p(dl), _p(il), B, C, h, w, _p(g), cx, cy, _p(gl), 0)
            _call("dsr_bilinear_ac_bwd", _p(gl), B, H, W, h, w, _p(gd))
        ctx.levels = None
        return gd, None, None


def smooth_loss(depth, image, num_scales=3):
    return _Smooth.apply(depth, image, num_scales)


class _Bicubic(Function):
    """F.interpolate(x, size=(Ho, Wo), mode='bicubic') (align_corners=False).  main_sr_model.py:279-293, :361, :396-398.
    Keeps the input's memory layout (NHWC-backed activations stay NHWC, NCHW planes stay planes)."""

    @staticmethod
    def forward(ctx, x, Ho, Wo):
        B, C, H, W = x.shape
        xp = x.permute(0, 2, 3, 1)
        if C > 1 and xp.is_contiguous():                       # channels-last activation
            y = torch.empty((B, Ho, Wo, C), device=x.device, dtype=torch.float32)
            _call("dsr_bicubic_fwd", _p(xp), B, H, W, C, Ho, Wo, _p(y))
            ctx.cfg = (B, H, W, C, Ho, Wo, True)
            return nchw(y)
        xc = planes(x)
        y = torch.empty((B, C, Ho, Wo), device=x.device, dtype=torch.float32)
        _call("dsr_bicubic_fwd", _p(xc), B * C, H, W, 1, Ho, Wo, _p(y))
        ctx.cfg = (B, H, W, C, Ho, Wo, False)
        return y

    @staticmethod
    def backward(ctx, gy):
        B, H, W, C, Ho, Wo, cl = ctx.cfg
        if cl:
            g = nhwc(gy)
            gx = torch.zeros((B, H, W, C), device=gy.device, dtype=torch.float32)
            _call("dsr_bicubic_bwd", _p(g), B, H, W, C, Ho, Wo, _p(gx))
            return nchw(gx), None, None
        g = planes(gy)
        gx = torch.zeros((B, C, H, W), device=gy.device, dtype=torch.float32)
        _call("dsr_bicubic_bwd", _p(g), B * C, H, W, 1, Ho, Wo, _p(gx))
        return gx, None, None


def bicubic(x, size):
    return _Bicubic.apply(x, int(size[0]), int(size[1]))


def nearest(x, size):
    """F.interpolate(x, size, mode='nearest') of NCHW planes (masks / depth; no gradient).
    main_sr_model.py:394-395, :452, :459."""
    xc = planes(x.detach())
    B, C, H, W = xc.shape
    y = torch.empty((B, C, int(size[0]), int(size[1])), device=xc.device, dtype=torch.float32)
    _call("dsr_nearest_fwd", _p(xc), B * C, H, W, 1, int(size[0]), int(size[1]), _p(y))
    return y


class _FovNormals(Function):
    """translation_network.SurfaceNormals (models/translation_network.py:329-360): field-of-view normals."""

    @staticmethod
    def forward(ctx, depth):
        d = planes(depth)
        B, _, H, W = d.shape
        out = torch.empty((B, 3, H, W), device=d.device, dtype=torch.float32)
        _call("dsr_fov_normals_fwd", _p(d), B, H, W, _p(out))
        ctx.save_for_backward(d)
        return out

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        B, _, H, W = d.shape
        gd = torch.zeros_like(d)
        _call("dsr_fov_normals_bwd", _p(d), _p(planes(g)), B, H, W, _p(gd))
        return gd


def fov_normals(depth):
    return _FovNormals.apply(depth)


class _CosSim(Function):
    """CosSimLoss (models/translation_network.py:310-316): mean over pixels of 1 - cos(x, y) along dim 1; gradient to x."""

    @staticmethod
    def forward(ctx, x, y):
        xp, yp = planes(x), planes(y.detach())
        B, C, H, W = xp.shape
        acc = _zeros_f64(1, xp.device)
        _call("dsr_cos_sim_fwd", _p(xp), _p(yp), B, C, H * W, _p(acc, torch.float64))
        out = torch.empty((), device=xp.device, dtype=torch.float32)
        _call("dsr_cvt_f64_f32", _p(acc, torch.float64), 1, _p(out), 1, 1.0 / (B * H * W), 0)
        ctx.save_for_backward(xp, yp)
        return out

    @staticmethod
    def backward(ctx, g):
        xp, yp = ctx.saved_tensors
        B, C, H, W = xp.shape
        gx = torch.empty_like(xp)
        _call("dsr_cos_sim_bwd", _p(xp), _p(yp), B, C, H * W, _p(g.contiguous()), 1.0 / (B * H * W), _p(gx))
        return gx, None


def cos_sim_loss(x, y):
    return _CosSim.apply(x, y)


class _CosSimMasked(Function):
    """sum over pixels of mask * (1 - cos(x, y)) along dim 1 (MaskedCosSimLoss, models/translation_network.py:320-327, without
    its denominator); gradient to x."""

    @staticmethod
    def forward(ctx, x, y, mask):
        xp, yp, mp = planes(x), planes(y.detach()), planes(mask.detach())
        B, C, H, W = xp.shape
        acc = _zeros_f64(1, xp.device)
        _call("dsr_cos_sim_masked_fwd", _p(xp), _p(yp), _p(mp), B, C, H * W, _p(acc, torch.float64))
        out = torch.empty((), device=xp.device, dtype=torch.float32)
        _call("dsr_cvt_f64_f32", _p(acc, torch.float64), 1, _p(out), 1, 1.0, 0)
        ctx.save_for_backward(xp, yp, mp)
        return out

    @staticmethod
    def backward(ctx, g):
        xp, yp, mp = ctx.saved_tensors
        B, C, H, W = xp.shape
        gx = torch.empty_like(xp)
        _call("dsr_cos_sim_masked_bwd", _p(xp), _p(yp), _p(mp), B, C, H * W, _p(g.contiguous()), 1.0, _p(gx))
        return gx, None, None


def cos_sim_masked_sum(x, y, mask):
    return _CosSimMasked.apply(x, y, mask)


class _MaskedMeanDif(Function):
    """MaskedMeanDif (models/translation_network.py:288-293): mean over samples of |sum (y - x) mask / (sum mask + 1e-6)|;
    gradient to x (one channel)."""

    @staticmethod
    def forward(ctx, x, y, mask):
        xp, yp, mp = planes(x), planes(y.detach()), planes(mask.detach())
        B, C, H, W = xp.shape
        if C != 1:
            raise ValueError("masked_mean_dif: one-channel depth maps only")
        sums = _zeros_f64(2 * B, xp.device)
        out = torch.empty((), device=xp.device, dtype=torch.float32)
        _call("dsr_masked_mean_dif_fwd", _p(xp), _p(yp), _p(mp), B, H * W, _p(sums, torch.float64), _p(out))
        ctx.save_for_backward(mp, sums)
        ctx.shape = (B, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        mp, sums = ctx.saved_tensors
        B, H, W = ctx.shape
        gx = torch.empty_like(mp)
        _call("dsr_masked_mean_dif_bwd", _p(mp), _p(sums, torch.float64), B, H * W, _p(g.contiguous()), _p(gx))
        return gx, None, None


def masked_mean_dif(x, y, mask):
    return _MaskedMeanDif.apply(x, y, mask)


class _LossSum(Function):
    """scale * sum_k sum_j w[k][j] * term_k[j] over tiny device tensors (1..4 elements each) in one launch; one more launch
    hands every term its gradient.  main_model.py:393-417."""

    @staticmethod
    def forward(ctx, scale, weights, flag, *terms):
        import ctypes
        n = len(terms)
        ts = [t.detach().contiguous().view(-1) for t in terms]
        ptrs = (ctypes.c_void_p * n)(*[_p(t) for t in ts])
        counts = _int_array([t.numel() for t in ts])
        flat = []
        for w, t in zip(weights, ts):
            w = list(w) if isinstance(w, (tuple, list)) else [w]
            if len(w) != t.numel() or not 1 <= len(w) <= 4:
                raise ValueError("loss_sum: one weight per element, 1..4 elements per term")
            flat += [float(x) for x in w] + [0.0] * (4 - len(w))
        wts = (ctypes.c_float * len(flat))(*flat)
        out = torch.empty((), device=ts[0].device, dtype=torch.float32)
        _call("dsr_loss_sum_fwd", ptrs, counts, wts, n, float(scale), _p(out), _p(flag, torch.int32))
        ctx.cfg = (float(scale), wts, n, [tuple(t.shape) for t in terms])
        return out

    @staticmethod
    def backward(ctx, g):
        scale, wts, n, shapes = ctx.cfg
        grads = torch.empty(4 * n, device=g.device, dtype=torch.float32)
        _call("dsr_loss_sum_bwd", _p(g.contiguous()), wts, n, scale, _p(grads))
        outs = []
        for k, shp in enumerate(shapes):
            m = 1
            for d in shp:
                m *= d
            outs.append(grads[4 * k:4 * k + m].view(shp) if ctx.needs_input_grad[3 + k] else None)
        return (None, None, None) + tuple(outs)


def loss_sum(terms, scale=1.0, nonfinite=None):
    """terms: [(tensor, weight or tuple of per-element weights), ...] -> scale * sum of the weighted elements (0-dim).
    `nonfinite`: optional int32[1] device counter, incremented whenever the result is NaN / Inf."""
    return _LossSum.apply(scale, tuple(w for _, w in terms), nonfinite, *[t for t, _ in terms])


def ssim(a, b):
    """Mean SSIM (11x11 Gaussian, sigma 1.5).  pytorch_ssim/__init__.py:17-37.  Forward only."""
    a, b = planes(a.detach()), planes(b.detach())
    B, C, H, W = a.shape
    acc = _zeros_f64(1, a.device)
    _call("dsr_ssim_fwd", _p(a), _p(b), B * C, H, W, _p(acc, torch.float64), None)
    return (acc / a.numel()).to(torch.float32)[0]


def adam_step(p, g, m, v, lr, step, b1=0.9, b2=0.999, eps=1e-8, grad_scale=1.0):
    _call("dsr_adam_step", _p(p), _p(g), _p(m), _p(v), p.numel(), lr, b1, b2, eps, step, grad_scale)


def adam_step_dev(p, g, m, v, hyper, step, grad_scale=1.0):
    """Adam with device-resident step counter (int32[1], incremented here) and hyper-parameters (float64[4])."""
    _call("dsr_adam_step_dev", _p(p), _p(g), _p(m), _p(v), p.numel(), _p(hyper, torch.float64), _p(step, torch.int32),
          grad_scale)
