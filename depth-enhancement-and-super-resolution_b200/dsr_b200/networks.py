"""Generator constructors of the hot path - drop-in for the reference's ``models/networks.py``.

Same public surface (``define_G``, ``ResnetGenerator``, ``UnetGenerator``, ``init_net``,
``init_weights``, ``get_norm_layer``, ``get_scheduler``), same module tree, hence the same
``state_dict`` keys / shapes and - because every layer subclasses the torch
layer the reference instantiates - the same RNG consumption, so ``torch.manual_seed(s)`` followed by
the same constructor calls yields bit-identical initial weights.  Only ``forward`` differs: every
layer runs through the hand-written sm_100a kernels of ``ops`` (no ATen compute, no CPU fallback).
Reference: /root/reference/models/networks.py (file:line cited per symbol).
"""
import functools

import torch
import torch.nn as nn
from torch.nn import init
from torch.optim import lr_scheduler

from . import ops


# ------------------------------------------------------------------------------------------------
# layers: torch parameter containers with our forward
# ------------------------------------------------------------------------------------------------
class Identity(nn.Module):                        # networks.py:13-15
    def forward(self, x):
        return x


class Conv2d(nn.Conv2d):
    """nn.Conv2d with zeros / reflect / replicate ``padding_mode`` (networks.py:379; translation_network.py:472)."""

    def forward(self, x, act_out=ops.ACT_NONE, pre_pad=None, want_stats=False, pro=None, skip_out=False):
        """pre_pad = (pad, mode) of an nn.ReflectionPad2d / ReplicationPad2d module placed right before
        this conv: it is folded into the conv's operand preparation instead of materialising a padded copy.
        want_stats: also return the per-(n, c) statistics of the output for the norm layer that follows.
        skip_out: also return an alias of x (last element) for a residual block's skip connection (ops._Conv2d)."""
        if self.dilation != (1, 1) or self.groups != 1:
            raise NotImplementedError("dsr_b200.Conv2d: dilation/groups are not on the hot path")
        p, mode = self.padding[0], self.padding_mode
        if pre_pad is not None:
            if p != 0:
                raise NotImplementedError("dsr_b200.Conv2d: explicit pad module followed by a padded conv")
            p, mode = pre_pad
        if isinstance(x, ops.LazyCat):
            if want_stats or pro is not None or skip_out:
                raise NotImplementedError("dsr_b200.Conv2d: a LazyCat input takes no fused prologue / statistics")
            if ops.conv_fusable("conv", x, self.weight, self.stride[0], p):
                return ops.cat_conv2d(x.parts, self.weight, self.bias, self.stride[0], p, act_out, pad_mode=mode)
            x = ops.cat(x.parts)
        return ops.conv2d(x, self.weight, self.bias, self.stride[0], p, act_out, pad_mode=mode, want_stats=want_stats,
                          pro=pro, skip_out=skip_out)

    def fusable(self, x, pre_pad=None):
        p = pre_pad[0] if pre_pad is not None else self.padding[0]
        return ops.conv_fusable("conv", x, self.weight, self.stride[0], p)


class ConvTranspose2d(nn.ConvTranspose2d):        # networks.py:406, :553
    def forward(self, x, act_out=ops.ACT_NONE, want_stats=False, pro=None):
        return ops.conv_transpose2d(x, self.weight, self.bias, self.stride[0], self.padding[0],
                                    self.output_padding[0], act_out, want_stats=want_stats, pro=pro)

    def fusable(self, x, pre_pad=None):
        return ops.conv_fusable("convT", x, self.weight, self.stride[0], self.padding[0], self.output_padding[0])


class ReflectionPad2d(nn.ReflectionPad2d):        # networks.py:378
    def forward(self, x):
        return ops.pad2d(x, self.padding[0], "reflect")


class ReplicationPad2d(nn.ReplicationPad2d):
    def forward(self, x):
        return ops.pad2d(x, self.padding[0], "replicate")


class InstanceNorm2d(nn.InstanceNorm2d):          # networks.py:30 (affine=False, no running stats)
    def forward(self, x, act=ops.ACT_NONE, residual=None, stats=None, hint=None):
        if self.affine or self.track_running_stats:
            raise NotImplementedError("dsr_b200.InstanceNorm2d: only affine=False, track_running_stats=False")
        return ops.instance_norm(x, self.eps, act, residual, stats, hint)


class GroupNorm(nn.GroupNorm):                    # translation_network.py:46
    def forward(self, x, act=ops.ACT_NONE, residual=None, stats=None, hint=None):
        return ops.group_norm(x, self.num_groups, self.weight, self.bias, self.eps, act, residual, stats, hint)


class ReLU(nn.ReLU):
    def forward(self, x):
        return ops.relu(x)


class LeakyReLU(nn.LeakyReLU):
    def forward(self, x):
        return ops.leaky_relu(x, self.negative_slope)


class Tanh(nn.Tanh):
    def forward(self, x):
        return ops.tanh(x)


def _conv_like(m):
    """the Conv2d / ConvTranspose2d a module list entry runs (translation_network.ConvTranspose wraps one)"""
    if isinstance(m, (Conv2d, ConvTranspose2d)):
        return m
    inner = getattr(m, "transposeconv", None)
    return inner if isinstance(inner, ConvTranspose2d) else None


def next_operand_hint(mods, i, x):
    """ops.operand_hint of the convolution that will read the activation `x` when `mods[i:]` run next (a residual block's first
    conv, or the pad + conv / transposed conv that follows the block stack); None when there is no such plain consumer."""
    pads = (ReflectionPad2d, ReplicationPad2d)
    if i >= len(mods):
        return None
    m = mods[i]
    inner = getattr(m, "conv_block", None)
    if inner is not None:                                  # a residual block: its own first layers
        return next_operand_hint(list(inner), 0, x)
    pad_mod = None
    if isinstance(m, pads) and i + 1 < len(mods) and isinstance(mods[i + 1], Conv2d) and mods[i + 1].padding[0] == 0:
        pad_mod, m = m, mods[i + 1]
    conv = _conv_like(m)
    if conv is None or conv.dilation != (1, 1) or conv.groups != 1:
        return None
    if isinstance(conv, Conv2d):
        p, mode = conv.padding[0], conv.padding_mode
        if pad_mod is not None:
            p, mode = pad_mod.padding[0], "reflect" if isinstance(pad_mod, ReflectionPad2d) else "replicate"
        return ops.operand_hint("conv", x.shape, conv.weight, conv.stride[0], p, mode)
    return ops.operand_hint("convT", x.shape, conv.weight, conv.stride[0], conv.padding[0], opad=conv.output_padding[0])


def run_fused(mods, x, tail_stats=False, skip_first=False):
    """Run a module list as fused units  [norm] [ReLU | LeakyReLU] [pad module] conv [Tanh]:
      * the norm-apply, the activation and the padding are folded into the conv's operand preparation
        (ops.Prologue) whenever the conv runs on the tcgen05 path - the normalised tensor is never written;
      * Tanh goes into the GEMM epilogue;
      * when a norm layer follows, the GEMM epilogue also takes its statistics and hands them over.
    Anything else runs module by module (norm + ReLU still share one pass).  tail_stats: the list ends with a conv
    whose norm layer is applied by the caller (residual blocks) -> returns (x, stats).  skip_first: when the list opens with a
    fused [pad] Conv2d, that conv also hands out an alias of the input for the caller's skip connection (ops._Conv2d.forward)
    -> returns (x, stats, alias or None)."""
    norms, pads = (InstanceNorm2d, GroupNorm), (ReflectionPad2d, ReplicationPad2d)
    n, i, stats = len(mods), 0, None          # stats = statistics of x, meaningful only while mods[i] is a norm layer
    skip = None
    while i < n:
        j, norm, act, pad = i, None, None, None
        if isinstance(mods[j], norms):
            norm, j = mods[j], j + 1
        if j < n and isinstance(mods[j], (ReLU, LeakyReLU)):
            act, j = mods[j], j + 1
        if j + 1 < n and isinstance(mods[j], pads) and isinstance(mods[j + 1], Conv2d) and mods[j + 1].padding[0] == 0:
            pad, j = mods[j], j + 1
        conv = _conv_like(mods[j]) if j < n else None
        pre_pad = (pad.padding[0], "reflect" if isinstance(pad, ReflectionPad2d) else "replicate") if pad is not None else None
        plain = norm is None and act is None
        # a GroupNorm that trains (its gamma / beta need gradients) runs as its own op: the fused prologue is forward-only
        gn_trains = isinstance(norm, GroupNorm) and torch.is_grad_enabled() and (norm.weight.requires_grad or x.requires_grad)
        if conv is not None and not gn_trains and (plain or conv.fusable(x, pre_pad)):
            kw = {}
            if pre_pad is not None:
                kw["pre_pad"] = pre_pad
            if not plain:
                a, slope = ops.ACT_NONE, 0.0
                if isinstance(act, LeakyReLU):
                    a, slope = ops.ACT_LRELU, act.negative_slope
                elif act is not None:
                    a = ops.ACT_RELU
                if isinstance(norm, GroupNorm):
                    kw["pro"] = ops.Prologue(stats, True, norm.eps, norm.num_groups, norm.weight, norm.bias, a, slope)
                elif norm is not None:
                    if norm.affine or norm.track_running_stats:
                        raise NotImplementedError("dsr_b200.InstanceNorm2d: only affine=False, track_running_stats=False")
                    kw["pro"] = ops.Prologue(stats, True, norm.eps, 0, None, None, a, slope)
                else:
                    kw["pro"] = ops.Prologue(act=a, slope=slope)
            after = mods[j + 1] if j + 1 < n else None
            stats = None
            want_skip = skip_first and i == 0 and plain and isinstance(conv, Conv2d) and not isinstance(x, ops.LazyCat)
            if want_skip:
                kw["skip_out"] = True
            if isinstance(after, Tanh):
                out = conv(x, act_out=ops.ACT_TANH, **kw)
                i = j + 2
            elif isinstance(after, norms) or (after is None and tail_stats):
                out = conv(x, want_stats=True, **kw)
                if want_skip:
                    out, skip = out[:2], out[2]
                x, stats = out
                i = j + 1
                continue
            else:
                out = conv(x, **kw)
                i = j + 1
            x, skip = out if want_skip else (out, skip)
            continue
        m = mods[i]
        if isinstance(m, norms):
            # (a stand-alone norm in front of a residual block: the same pass also writes the block's first operand)
            if i + 1 < n and isinstance(mods[i + 1], ReLU):
                x = m(x, act=ops.ACT_RELU, stats=stats, hint=next_operand_hint(mods, i + 2, x) if stats is not None else None)
                i += 2
            else:
                x = m(x, stats=stats, hint=next_operand_hint(mods, i + 1, x) if stats is not None else None)
                i += 1
        elif hasattr(m, "conv_block") and hasattr(m, "forward_hinted"):
            x = m.forward_hinted(x, next_operand_hint(mods, i + 1, x))      # residual block: its closing norm feeds mods[i + 1]
            i += 1
        else:
            x = m(x)
            i += 1
        stats = None
    if skip_first:
        return x, stats, skip
    return (x, stats) if tail_stats else x


class FusedSequential(nn.Sequential):
    """nn.Sequential with the same indices (=> same state_dict keys), executed by ``run_fused``."""

    def forward(self, x):
        return run_fused(list(self), x)


# ------------------------------------------------------------------------------------------------
# helpers with the reference's signatures
# ------------------------------------------------------------------------------------------------
def get_norm_layer(norm_type="instance"):         # networks.py:18-37
    if norm_type == "instance":
        return functools.partial(InstanceNorm2d, affine=False, track_running_stats=False)
    if norm_type == "none":
        return lambda x: Identity()
    if norm_type in ("batch", "group"):
        # 'batch' is never used by the main path (default --norm instance); 'group' crashes in the
        # reference itself for define_G nets.
        raise NotImplementedError("normalization layer [%s] is not on the dsr_b200 hot path" % norm_type)
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


def get_scheduler(optimizer, opt):                # networks.py:40-66
    if opt.lr_policy == "linear":
        def lambda_rule(epoch):
            return 1.0 - max(0, epoch + opt.epoch_count - opt.n_epochs) / float(opt.n_epochs_decay + 1)
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda_rule)
    if opt.lr_policy == "step":
        return lr_scheduler.StepLR(optimizer, step_size=opt.lr_decay_iters, gamma=0.1)
    if opt.lr_policy == "plateau":
        return lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=0.2, threshold=0.01, patience=5)
    if opt.lr_policy == "cosine":
        return lr_scheduler.CosineAnnealingLR(optimizer, T_max=opt.n_epochs, eta_min=0)
    return NotImplementedError("learning rate policy [%s] is not implemented", opt.lr_policy)


def init_weights(net, init_type="normal", init_gain=0.02):     # networks.py:69-100
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
            if init_type == "normal":
                init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if hasattr(m, "bias") and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif classname.find("BatchNorm2d") != -1:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)

    print("initialize network with %s" % init_type)
    net.apply(init_func)


class DeviceModule(nn.Module):
    """What ``init_net`` returns for non-empty gpu_ids: exposes ``.module`` like the reference's
    ``torch.nn.DataParallel`` wrapper (networks.py:113-116, base_model.py:163,194) but runs the net on
    the ONE device of this process - data parallelism is one process per GPU (``dsr_b200.parallel``),
    never a single-process scatter/gather."""

    def __init__(self, module, device):
        super().__init__()
        self.module = module
        self.device = torch.device(device)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def init_net(net, init_type="normal", init_gain=0.02, gpu_ids=[]):     # networks.py:103-118
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        dev = torch.device("cuda", gpu_ids[0])
        net.to(dev)
        net = DeviceModule(net, dev)
    init_weights(net, init_type, init_gain=init_gain)
    return net


def define_G(input_nc, output_nc, ngf, netG, norm="batch", use_dropout=False, init_type="normal", init_gain=0.02,
             gpu_ids=[], replace_transpose=False, n_down=2, use_sr=False, use_old=False):   # networks.py:121-163
    norm_layer = get_norm_layer(norm_type=norm)
    if netG == "resnet_9blocks":
        net = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout, n_blocks=9,
                              replace_transpose=replace_transpose, n_downsampling=n_down)
    elif netG == "resnet_6blocks":
        net = ResnetGenerator(input_nc, output_nc, ngf, norm_layer=norm_layer, use_dropout=use_dropout, n_blocks=6,
                              replace_transpose=replace_transpose, n_downsampling=n_down)
    elif netG == "unet_128":
        net = UnetGenerator(input_nc, output_nc, 7, ngf, norm_layer=norm_layer, use_dropout=use_dropout,
                            use_sr=use_sr, use_old=use_old)
    elif netG == "unet_256":
        net = UnetGenerator(input_nc, output_nc, 8, ngf, norm_layer=norm_layer, use_dropout=use_dropout,
                            use_sr=use_sr, use_old=use_old)
    else:
        raise NotImplementedError("Generator model name [%s] is not recognized" % netG)
    return init_net(net, init_type, init_gain, gpu_ids)


def _uses_bias(norm_layer):
    if type(norm_layer) == functools.partial:
        return norm_layer.func in (InstanceNorm2d, nn.InstanceNorm2d)
    return norm_layer in (InstanceNorm2d, nn.InstanceNorm2d)


# ------------------------------------------------------------------------------------------------
# ResNet generator (I2D_features, Depth_f)   networks.py:353-481
# ------------------------------------------------------------------------------------------------
class ResnetGenerator(nn.Module):
    def __init__(self, input_nc, output_nc, ngf=64, norm_layer=InstanceNorm2d, use_dropout=False, n_blocks=6,
                 padding_type="reflect", replace_transpose=False, n_downsampling=2):
        assert n_blocks >= 0
        super().__init__()
        if use_dropout or replace_transpose:
            raise NotImplementedError("dsr_b200: use_dropout / replace_transpose are not on the hot path")
        use_bias = _uses_bias(norm_layer)
        model = [ReflectionPad2d(3), Conv2d(input_nc, ngf, kernel_size=7, padding=0, bias=use_bias),
                 norm_layer(ngf), ReLU(True)]
        for i in range(n_downsampling):
            mult = 2 ** i
            model += [Conv2d(ngf * mult, ngf * mult * 2, kernel_size=3, stride=2, padding=1, bias=use_bias),
                      norm_layer(ngf * mult * 2), ReLU(True)]
        mult = 2 ** n_downsampling
        for i in range(n_blocks):
            model += [ResnetBlock(ngf * mult, padding_type=padding_type, norm_layer=norm_layer,
                                  use_dropout=use_dropout, use_bias=use_bias)]
        for i in range(n_downsampling):
            mult = 2 ** (n_downsampling - i)
            model += [ConvTranspose2d(ngf * mult, int(ngf * mult / 2), kernel_size=3, stride=2, padding=1,
                                      output_padding=1, bias=use_bias),
                      norm_layer(int(ngf * mult / 2)), ReLU(True)]
        model += [ReflectionPad2d(3)]
        model += [Conv2d(ngf, output_nc, kernel_size=7, padding=0)]
        model += [Tanh()]
        self.model = FusedSequential(*model)

    def forward(self, input):
        return self.model(input)


class ResnetBlock(nn.Module):                     # networks.py:424-481
    def __init__(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        super().__init__()
        self.conv_block = self.build_conv_block(dim, padding_type, norm_layer, use_dropout, use_bias)

    def build_conv_block(self, dim, padding_type, norm_layer, use_dropout, use_bias):
        pads = {"reflect": ReflectionPad2d, "replicate": ReplicationPad2d}
        conv_block = []
        p = 0
        if padding_type in pads:
            conv_block += [pads[padding_type](1)]
        elif padding_type == "zero":
            p = 1
        else:
            raise NotImplementedError("padding [%s] is not implemented" % padding_type)
        conv_block += [Conv2d(dim, dim, kernel_size=3, padding=p, bias=use_bias), norm_layer(dim), ReLU(True)]
        if use_dropout:
            raise NotImplementedError("dsr_b200: dropout is not on the hot path")
        if padding_type in pads:
            conv_block += [pads[padding_type](1)]
        conv_block += [Conv2d(dim, dim, kernel_size=3, padding=p, bias=use_bias), norm_layer(dim)]
        return FusedSequential(*conv_block)

    def forward(self, x):
        return self.forward_hinted(x, None)

    def forward_hinted(self, x, hint):
        """hint = ops.operand_hint of the layer that reads this block's output: the closing norm + skip add then also writes
        that layer's arranged operand (one pass instead of two)"""
        mods = list(self.conv_block)
        if isinstance(mods[-1], (InstanceNorm2d, GroupNorm)):      # skip add fused into the norm pass
            if ops.CONFIG["fuse_skip_grad"] and torch.is_grad_enabled() and x.requires_grad:
                # the skip connection leaves through the first conv's node: both gradients of x meet in ITS backward
                y, stats, skip = run_fused(mods[:-1], x, tail_stats=True, skip_first=True)
                return mods[-1](y, residual=x if skip is None else skip, stats=stats, hint=hint)
            y, stats = run_fused(mods[:-1], x, tail_stats=True)
            return mods[-1](y, residual=x, stats=stats, hint=hint)  # networks.py:480
        return x + self.conv_block(x)


# ------------------------------------------------------------------------------------------------
# U-Net generator (Image2Depth, Task)   networks.py:484-629
# ------------------------------------------------------------------------------------------------
class UnetGenerator(nn.Module):
    def __init__(self, input_nc, output_nc, num_downs, ngf=64, norm_layer=InstanceNorm2d, use_dropout=False,
                 use_sr=False, use_old=False):
        super().__init__()
        if use_sr or use_old:
            raise NotImplementedError("dsr_b200: use_sr / use_old U-Net variants are not on the hot path")
        blk = UnetSkipConnectionBlock(ngf * 8, ngf * 8, input_nc=None, submodule=None, norm_layer=norm_layer,
                                      innermost=True)
        for i in range(num_downs - 5):
            blk = UnetSkipConnectionBlock(ngf * 8, ngf * 8, input_nc=None, submodule=blk, norm_layer=norm_layer,
                                          use_dropout=use_dropout)
        blk = UnetSkipConnectionBlock(ngf * 4, ngf * 8, input_nc=None, submodule=blk, norm_layer=norm_layer)
        blk = UnetSkipConnectionBlock(ngf * 2, ngf * 4, input_nc=None, submodule=blk, norm_layer=norm_layer)
        blk = UnetSkipConnectionBlock(ngf, ngf * 2, input_nc=None, submodule=blk, norm_layer=norm_layer)
        self.model = UnetSkipConnectionBlock(output_nc, ngf, input_nc=input_nc, submodule=blk, outermost=True,
                                             norm_layer=norm_layer)

    def forward(self, input):
        return self.model(input)


class UnetSkipConnectionBlock(nn.Module):
    def __init__(self, outer_nc, inner_nc, input_nc=None, submodule=None, outermost=False, innermost=False,
                 norm_layer=InstanceNorm2d, use_dropout=False):
        super().__init__()
        self.outermost = outermost
        use_bias = _uses_bias(norm_layer)
        if input_nc is None:
            input_nc = outer_nc
        downconv = Conv2d(input_nc, inner_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
        downrelu = LeakyReLU(0.2, False)
        downnorm = norm_layer(inner_nc)
        uprelu = ReLU(True)
        upnorm = norm_layer(outer_nc)
        if outermost:
            upconv = ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1)
            model = [downconv] + [submodule] + [uprelu, upconv, Tanh()]
        elif innermost:
            upconv = ConvTranspose2d(inner_nc, outer_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
            model = [downrelu, downconv] + [uprelu, upconv, upnorm]
        else:
            upconv = ConvTranspose2d(inner_nc * 2, outer_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
            if use_dropout:
                raise NotImplementedError("dsr_b200: dropout is not on the hot path")
            model = [downrelu, downconv, downnorm] + [submodule] + [uprelu, upconv, upnorm]
        self.model = FusedSequential(*model)

    def forward(self, x):
        if self.outermost:
            return self.model(x)
        return ops.cat([x, self.model(x)])        # networks.py:629
