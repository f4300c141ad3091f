"""Oracle (CPU) functional restatement of the five networks of the hot path.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  The arithmetic of every layer lives in a
third-party dependency of the reference, PyTorch (``requirements.txt:19`` pins torch==1.6.0; this
image has 2.11.0 - no semantic change for these ops), so the restatement
calls ``torch.nn.functional`` on CPU in fp32 and is anchored on the reference's own call sites.
Weights come in as a plain ``state_dict`` (``OrderedDict[str, Tensor]``) in the reference's
checkpoint layout.
"""
import torch
import torch.nn.functional as F


def _in(x, eps=1e-5):
    """InstanceNorm2d(affine=False, track_running_stats=False)  (models/networks.py:30)."""
    return F.instance_norm(x, eps=eps)


def _gn(x, w, b):
    """GroupNorm(8, C, affine=True)  (models/translation_network.py:46)."""
    return F.group_norm(x, 8, w, b, eps=1e-5)


def resnet_generator(sd, x, n_blocks=6, n_down=2):
    """networks.ResnetGenerator.forward (models/networks.py:353-421), blocks :424-481.

    Sequential indices: 0 ReflPad3, 1 Conv7, 2 IN, 3 ReLU; then per down-sampling (Conv3 s2 p1,
    IN, ReLU); n_blocks ResnetBlocks; per up-sampling (ConvT3 s2 p1 op1, IN, ReLU); ReflPad3,
    Conv7, Tanh.
    """
    i = 1
    x = F.conv2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), sd[f"model.{i}.weight"], sd[f"model.{i}.bias"])
    x = F.relu(_in(x))
    i = 4
    for _ in range(n_down):
        x = F.relu(_in(F.conv2d(x, sd[f"model.{i}.weight"], sd[f"model.{i}.bias"], stride=2, padding=1)))
        i += 3
    for _ in range(n_blocks):
        p = f"model.{i}.conv_block."
        y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), sd[p + "1.weight"], sd[p + "1.bias"])
        y = F.relu(_in(y))
        y = F.conv2d(F.pad(y, (1, 1, 1, 1), mode="reflect"), sd[p + "5.weight"], sd[p + "5.bias"])
        x = x + _in(y)                                                  # networks.py:480
        i += 1
    for _ in range(n_down):
        x = F.conv_transpose2d(x, sd[f"model.{i}.weight"], sd[f"model.{i}.bias"], stride=2,
                               padding=1, output_padding=1)
        x = F.relu(_in(x))
        i += 3
    i += 1
    x = F.conv2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), sd[f"model.{i}.weight"], sd[f"model.{i}.bias"])
    return torch.tanh(x)


def unet_generator(sd, x, num_downs=7):
    """networks.UnetGenerator.forward (models/networks.py:484-513) with the recursive
    UnetSkipConnectionBlock (:516-629).  Down path pre-activation is a NON in-place
    LeakyReLU(0.2) (:546), so each skip carries the un-activated tensor (:629)."""

    def block(prefix, x, depth):
        outermost = depth == 0
        innermost = depth == num_downs - 1
        if outermost:
            d = F.conv2d(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"], stride=2, padding=1)
            s = block(prefix + "1.model.", d, depth + 1)
            u = F.conv_transpose2d(F.relu(s), sd[prefix + "3.weight"], sd[prefix + "3.bias"],
                                   stride=2, padding=1)
            return torch.tanh(u)
        if innermost:
            d = F.conv2d(F.leaky_relu(x, 0.2), sd[prefix + "1.weight"], sd[prefix + "1.bias"],
                         stride=2, padding=1)
            u = F.conv_transpose2d(F.relu(d), sd[prefix + "3.weight"], sd[prefix + "3.bias"],
                                   stride=2, padding=1)
            return torch.cat([x, _in(u)], 1)
        d = F.conv2d(F.leaky_relu(x, 0.2), sd[prefix + "1.weight"], sd[prefix + "1.bias"],
                     stride=2, padding=1)
        s = block(prefix + "3.model.", _in(d), depth + 1)
        u = F.conv_transpose2d(F.relu(s), sd[prefix + "5.weight"], sd[prefix + "5.bias"],
                               stride=2, padding=1)
        return torch.cat([x, _in(u)], 1)

    return block("model.model.", x, 0)


def _rconv(x, w, b, k, stride):
    """Conv2d(padding_mode='replicate') as used in translation_network.py:472-478."""
    p = (k - 1) // 2 if stride == 1 else 1
    return F.conv2d(F.pad(x, (p, p, p, p), mode="replicate"), w, b, stride=stride)


def translation_generator(sd, depth, img=None, n_blocks=9, n_down=2):
    """translation_network.Generator('img_depth').forward (models/translation_network.py:641-649):
    Encoder x2 (:466-483), ResnetBottlenec on cat(depth, img) (:533-575), Decoder (:485-510).
    img None = Generator('depth') (:650-653): one 64-channel depth encoder, no image branch."""

    def enc(p, x):
        x = F.relu(_gn(_rconv(x, sd[p + "0.weight"], None, 7, 1), sd[p + "1.weight"], sd[p + "1.bias"]))
        j = 3
        for _ in range(n_down):
            x = _rconv(x, sd[p + f"{j}.weight"], None, 4, 2)
            x = F.relu(_gn(x, sd[p + f"{j + 1}.weight"], sd[p + f"{j + 1}.bias"]))
            j += 3
        return x

    fd = enc("enc_depth.model.", depth)
    x = torch.cat((fd, enc("enc_img.model.", img)), dim=1) if img is not None else fd    # translation_network.py:549
    for i in range(n_blocks):
        p = f"bottlenec.model.{i}.conv_block."
        y = F.relu(_gn(_rconv(x, sd[p + "0.weight"], None, 3, 1), sd[p + "1.weight"], sd[p + "1.bias"]))
        y = _gn(_rconv(y, sd[p + "3.weight"], None, 3, 1), sd[p + "4.weight"], sd[p + "4.bias"])
        x = x + y
    j = 0
    for _ in range(n_down):
        p = f"dec_depth.model.{j}."
        x = F.conv_transpose2d(x, sd[p + "transposeconv.weight"], None, stride=2, padding=1)
        x = F.relu(_gn(x, sd[f"dec_depth.model.{j + 1}.weight"], sd[f"dec_depth.model.{j + 1}.bias"]))
        j += 3
    x = _rconv(x, sd[f"dec_depth.model.{j}.weight"], sd[f"dec_depth.model.{j}.bias"], 7, 1)
    return torch.tanh(x)


def nlayer_discriminator(sd, x, n_layers=3):
    """translation_network.NLayerDiscriminator with norm_d='none' (models/translation_network.py:735-776): conv k4 s2 + bias,
    LeakyReLU(0.2); (n_layers - 1) x [conv k4 s2, LeakyReLU]; conv k4 s1, LeakyReLU; conv k4 s1 + bias -> 1 channel.
    Sequential indices: 0 | 2, 5, ... (conv, Identity, LeakyReLU) | last."""
    x = F.leaky_relu(F.conv2d(x, sd["model.0.weight"], sd["model.0.bias"], stride=2, padding=1), 0.2)
    j = 2
    for _ in range(1, n_layers):
        x = F.leaky_relu(F.conv2d(x, sd[f"model.{j}.weight"], sd.get(f"model.{j}.bias"), stride=2, padding=1), 0.2)
        j += 3
    x = F.leaky_relu(F.conv2d(x, sd[f"model.{j}.weight"], sd.get(f"model.{j}.bias"), stride=1, padding=1), 0.2)
    j += 3
    return F.conv2d(x, sd[f"model.{j}.weight"], sd[f"model.{j}.bias"], stride=1, padding=1)


def gan_block_step(sd_g, sd_d, depth, img, real):
    """The compute of BASELINE configs[4] (generator + PatchGAN discriminator, forward + backward) with the LSGAN terms of
    models/translation_model.py: loss_G = 0.5 * MSE(D(G(depth, img)), 1) (:214), back-propagated through D into G;
    loss_D = 0.5 * (MSE(D(real), 1) + MSE(D(fake.detach()), 0)) (:199-205), back-propagated into D.
    sd_g / sd_d: state dicts whose tensors require grad.  -> dict(fake, pred_fake, loss_G, loss_D, grads_g, grads_d)."""
    fake = translation_generator(sd_g, depth, img)
    pred_fake = nlayer_discriminator(sd_d, fake)
    loss_G = 0.5 * F.mse_loss(pred_fake, torch.ones_like(pred_fake))
    gg = torch.autograd.grad(loss_G, [v for v in sd_g.values()], allow_unused=True)
    pred_real = nlayer_discriminator(sd_d, real)
    pred_fake_d = nlayer_discriminator(sd_d, fake.detach())
    loss_D = 0.5 * (F.mse_loss(pred_real, torch.ones_like(pred_real)) + F.mse_loss(pred_fake_d, torch.zeros_like(pred_fake_d)))
    gd = torch.autograd.grad(loss_D, [v for v in sd_d.values()])
    return dict(fake=fake.detach(), pred_fake=pred_fake.detach(), loss_G=float(loss_G), loss_D=float(loss_D),
                grads_g=dict(zip(sd_g.keys(), gg)), grads_d=dict(zip(sd_d.keys(), gd)))
