"""Translation generator (``define_Gen`` / ``Generator``: G_A_d, frozen on the hot path) and the PatchGAN
discriminator (``define_D`` / ``NLayerDiscriminator``) - drop-in for those parts of the reference's
``models/translation_network.py``.

Same module trees => same ``state_dict`` keys (``enc_img.model.*``, ``enc_depth.model.*``,
``bottlenec.model.*.conv_block.*``, ``dec_depth.model.*``, discriminator ``model.N.*``).
The main training step never back-propagates into G_A_d (main_model.py:426) and runs it under ``no_grad`` with the
norm layers folded into the conv prologues; with gradients enabled the GroupNorm layers run as their own op
(``ops.group_norm``, which has a backward), so both families train - the building blocks of the translation_block
row (the ``TranslationModel`` step itself is not built yet).
Reference: /root/reference/models/translation_network.py (file:line cited per symbol).
"""
import functools

import torch
import torch.nn as nn
from torch.nn import init

from . import ops
from .networks import (next_operand_hint, Conv2d, ConvTranspose2d as _ConvT2d, DeviceModule, FusedSequential, GroupNorm, Identity,
                       InstanceNorm2d, LeakyReLU, ReLU, Tanh, run_fused)


def get_norm_layer(norm_type="instance"):         # translation_network.py:34-52
    if norm_type == "group":
        return lambda n_ch: GroupNorm(num_groups=8, num_channels=n_ch, affine=True)
    if norm_type == "instance":
        return functools.partial(InstanceNorm2d, affine=False, track_running_stats=False)
    if norm_type == "none":
        return lambda x: Identity()
    if norm_type == "batch":
        raise NotImplementedError("normalization layer [batch] is not on the dsr_b200 hot path")
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


def get_upsampling(upsampling_type="transpose"):  # translation_network.py:21-32
    if upsampling_type == "transpose":
        return functools.partial(ConvTranspose)
    raise NotImplementedError("upsample layer [%s] is not on the dsr_b200 hot path" % upsampling_type)


def init_weights(net, init_type="normal", init_gain="relu", param=None):   # translation_network.py:82-116
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
            if init_type == "normal":
                init.normal_(m.weight.data, mean=0.0, std=0.02)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=init.calculate_gain(init_gain, param))
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(m.weight.data, gain=init.calculate_gain(init_gain, param))
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if hasattr(m, "bias") and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif hasattr(m, "weight") and (m.weight is not None) and (classname.find("Norm") != -1):
            init.normal_(m.weight.data, mean=1.0, std=0.02)
            init.constant_(m.bias.data, 0.0)

    print("initialize network with %s" % init_type)
    net.apply(init_func)


def init_net(net, init_type="normal", init_gain="relu", gpu_ids=[], param=None):   # translation_network.py:119-133
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        dev = torch.device("cuda", gpu_ids[0])
        net = DeviceModule(net, dev).to(dev)
    init_weights(net=net, init_type=init_type, init_gain=init_gain, param=param)
    return net


class Encoder(nn.Module):                         # translation_network.py:466-483
    def __init__(self, input_nc, base_nc, norm_layer, use_bias, opt):
        super().__init__()
        model = [Conv2d(input_nc, base_nc, kernel_size=7, stride=1, padding=3, dilation=1, padding_mode="replicate",
                        bias=use_bias), norm_layer(base_nc), ReLU(True)]
        for i in range(opt.n_downsampling):
            mult = 2 ** i
            model += [Conv2d(base_nc * mult, base_nc * mult * 2, kernel_size=4, stride=2, padding=1, dilation=1,
                             padding_mode="replicate", bias=use_bias), norm_layer(base_nc * mult * 2), ReLU(True)]
        self.model = FusedSequential(*model)

    def forward(self, x):
        return self.model(x)


class ConvTranspose(nn.Module):                   # translation_network.py:505-510
    def __init__(self, in_chanels, out_chanels, use_bias, opt):
        super().__init__()
        self.transposeconv = _ConvT2d(in_chanels, out_chanels, kernel_size=4, stride=2, padding=1,
                                      output_padding=0, dilation=1, padding_mode="zeros", bias=use_bias)

    def forward(self, x):
        return self.transposeconv(x)


class Decoder(nn.Module):                         # translation_network.py:485-503
    def __init__(self, base_nc, output_nc, norm_layer, use_bias, up_layer, opt, output="depth"):
        super().__init__()
        model = []
        for i in range(opt.n_downsampling):
            mult = 2 ** (opt.n_downsampling - i)
            model += [up_layer(mult * base_nc, int(base_nc * mult / 2), use_bias=use_bias, opt=opt),
                      norm_layer(int(base_nc * mult / 2)), ReLU(True)]
        model += [Conv2d(base_nc, output_nc, kernel_size=7, stride=1, padding=3, dilation=1,
                         padding_mode="replicate", bias=True)]
        if output == "depth":
            assert output_nc == 1, "only 1 chanels for depth"
            model += [Tanh()]
        else:
            assert output == "semantic"
        self.model = FusedSequential(*model)

    def forward(self, x):
        return self.model(x)


class ResnetBlock(nn.Module):                     # translation_network.py:554-575
    def __init__(self, dim, dilation, norm_layer, use_bias, opt):
        super().__init__()
        if dilation != 1 or opt.dropout:
            raise NotImplementedError("dsr_b200: dilated / dropout bottleneck blocks are not on the hot path")
        self.conv_block = FusedSequential(
            Conv2d(dim, dim, kernel_size=3, stride=1, padding=1, dilation=1, padding_mode="replicate", bias=use_bias),
            norm_layer(dim), ReLU(True),
            Conv2d(dim, dim, kernel_size=3, padding=1, dilation=1, padding_mode="replicate", bias=use_bias),
            norm_layer(dim))

    def forward(self, x):
        return self.forward_hinted(x, None)

    def forward_hinted(self, x, hint):
        mods = list(self.conv_block)
        y, stats = run_fused(mods[:-1], x, tail_stats=True)
        return mods[-1](y, residual=x, stats=stats, hint=hint)     # translation_network.py:574


class ResnetBottlenec(nn.Module):                 # translation_network.py:533-552
    def __init__(self, base_nc, n_blocks, norm_layer, use_bias, opt, use_dilation=False):
        super().__init__()
        if use_dilation:
            raise NotImplementedError("dsr_b200: dilated bottleneck is not on the hot path")
        mult = 2 ** opt.n_downsampling
        self.model = nn.Sequential(*[ResnetBlock(dim=base_nc * mult, dilation=1, norm_layer=norm_layer,
                                                 use_bias=use_bias, opt=opt) for _ in range(n_blocks)])

    def forward(self, depth, img=None):
        x = ops.cat([depth, img]) if img is not None else depth     # translation_network.py:549
        blocks = list(self.model)
        for k, blk in enumerate(blocks):                             # each block's closing norm also prepares the next block's operand
            x = blk.forward_hinted(x, next_operand_hint(blocks, k + 1, x))
        return x


def define_Gen(opt, input_type, out_type="depth"):                 # translation_network.py:577-585
    use_bias = opt.norm == "instance"
    if (input_type == "img" and out_type == "feature") or (input_type == "feature" and out_type == "depth"):
        raise NotImplementedError("dsr_b200: GeneratorI_F / GeneratorF_D are not on the hot path")
    net = Generator(opt, input_type, use_bias)
    return init_net(net=net, init_type=opt.init_type, init_gain="relu", gpu_ids=opt.gpu_ids)


class Generator(nn.Module):                       # translation_network.py:612-662
    def __init__(self, opt, input_type, use_bias):
        super().__init__()
        self.input_type = input_type
        self.opt = opt
        norm_layer = get_norm_layer(norm_type=opt.norm)
        up_layer = get_upsampling(upsampling_type=opt.upsampling_type)
        if getattr(opt, "use_semantic", False):
            raise NotImplementedError("dsr_b200: the semantic decoder is not on the hot path")
        if input_type == "img_depth":
            base_nc = opt.ngf_img + opt.ngf_depth
            self.enc_img = Encoder(opt.input_nc_img, opt.ngf_img, norm_layer, use_bias, opt)
            self.enc_depth = Encoder(opt.input_nc_depth, opt.ngf_depth, norm_layer, use_bias, opt)
        elif input_type == "depth":
            base_nc = opt.ngf_depth * 2
            self.enc_depth = Encoder(opt.input_nc_depth, base_nc, norm_layer, use_bias, opt)
        else:
            raise NotImplementedError("Specify input type")
        self.bottlenec = ResnetBottlenec(base_nc, opt.n_blocks, norm_layer, use_bias, opt)
        self.dec_depth = Decoder(base_nc, opt.output_nc_depth, norm_layer, use_bias, up_layer, opt, output="depth")

    def forward(self, depth, img=None, return_logits=False):
        if self.input_type == "img_depth":
            img = self.enc_img(img)
            depth = self.enc_depth(depth)
            return self.dec_depth(self.bottlenec(depth, img))
        depth = self.enc_depth(depth)
        return self.dec_depth(self.bottlenec(depth))


class NLayerDiscriminator(nn.Module):             # translation_network.py:735-776
    """PatchGAN discriminator: conv k4 s2 + LeakyReLU(0.2), (n_layers - 1) x [conv k4 s2, norm, LeakyReLU], conv k4 s1,
    norm, LeakyReLU, conv k4 s1 -> 1 channel prediction map."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=None, use_bias=False):
        super().__init__()
        norm_layer = norm_layer or (lambda c: Identity())
        kw, padw = 4, 1
        sequence = [Conv2d(input_nc, ndf, kernel_size=kw, stride=2, padding=padw, bias=True), LeakyReLU(0.2, True)]
        nf_mult = 1
        for n in range(1, n_layers):
            nf_mult_prev, nf_mult = nf_mult, min(2 ** n, 8)
            sequence += [Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=2, padding=padw, bias=use_bias),
                         norm_layer(ndf * nf_mult), LeakyReLU(0.2, True)]
        nf_mult_prev, nf_mult = nf_mult, min(2 ** n_layers, 8)
        sequence += [Conv2d(ndf * nf_mult_prev, ndf * nf_mult, kernel_size=kw, stride=1, padding=padw, bias=use_bias),
                     norm_layer(ndf * nf_mult), LeakyReLU(0.2, True)]
        sequence += [Conv2d(ndf * nf_mult, 1, kernel_size=kw, stride=1, padding=padw, bias=True)]
        self.model = nn.Sequential(*sequence)      # plain Sequential: an Identity norm sits between conv and activation

    def forward(self, input):
        return self.model(input)


def define_D(opt, input_type="depth"):            # translation_network.py:666-724
    try:
        input_nc = {"depth": 1, "normal": 3, "depth_normal": 4}[input_type]
    except KeyError:
        raise NotImplementedError("Input for discriminator [%s] is not recognized" % input_type)
    norm_layer = get_norm_layer(norm_type=opt.norm_d)
    use_bias = opt.norm_d == "instance"
    if opt.netD == "basic":
        net = NLayerDiscriminator(input_nc, opt.ndf, n_layers=3, norm_layer=norm_layer, use_bias=use_bias)
    elif opt.netD == "n_layers":
        net = NLayerDiscriminator(input_nc, opt.ndf, opt.n_layers_D, norm_layer=norm_layer, use_bias=use_bias)
    elif opt.netD in ("pixel", "Gu"):
        raise NotImplementedError("dsr_b200: discriminator [%s] is not built (README.md:51 trains with n_layers)" % opt.netD)
    else:
        raise NotImplementedError("Discriminator model name [%s] is not recognized" % opt.netD)
    if getattr(opt, "use_spnorm", False):
        raise NotImplementedError("dsr_b200: spectral normalisation of the discriminator is not built")
    return init_net(net=net, init_type=opt.init_type, init_gain="leaky_relu", gpu_ids=opt.gpu_ids, param=0.2)
