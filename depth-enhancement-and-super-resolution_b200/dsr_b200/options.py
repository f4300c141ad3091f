"""Option namespace with the reference's flag names and defaults for the main training step.

The reference builds ``opt`` with its own argparse stack (options/base_options.py:20-61,
options/train_options.py:10-170), which stays usable as-is: ``MainModel`` only reads attributes.
This helper exists so the step can be driven without that stack (bench, tests, smoke); every
default below is the reference's default, and ``main_flags()`` applies README.md:70's training flags.
"""
from types import SimpleNamespace

_DEFAULTS = dict(
    # base_options.py
    name="experiment_name", gpu_ids=[0], checkpoints_dir="./checkpoints", model="main_network_best",
    model_type="main", input_nc=3, output_nc=3, ngf=64, ndf=64, norm="instance", init_type="normal",
    init_gain=0.02, no_dropout=True, direction="AtoB", batch_size=1, preprocess="resize_and_crop",
    epoch="latest", load_iter=0, verbose=False, suffix="",
    # train_options.py
    continue_train=False, epoch_count=1, phase="train", n_epochs=100, n_epochs_decay=100, beta1=0.5, lr=0.0002,
    gan_mode="lsgan", lr_policy="linear", lr_decay_iters=50, replace_transpose=False, print_mean=False,
    save_all=False, save_image_folder="./results/", SR=False, Depthf_ndown=2, Task_ndown=2, Depthf_basef=32, Task_basef=64, Depthf_outf=128,
    Depthf_type="resnet_6blocks", Task_type="unet_128", use_rec_as_real_input=False, use_image_for_trans=False,
    norm_loss=False, use_smooth_loss=False, w_syn_norm=0.0, w_syn_l1=1.0, w_syn_holes=2.0, w_real_holes=5.0,
    w_real_l1_d=1.0, w_real_l1_i=0.1, w_smooth=0.1, ImageDepthf_outf=128, ImageDepthf_basef=32,
    ImageDepthf_type="resnet_6blocks", I2D_base=64, I2D_type="unet_128", scale_G=1.0, use_edge=False,
    use_masked=False, crop_size_h=384, crop_size_w=512, lambda_identity=0.5, isTrain=True,
    Imagef_ndown=2, Imagef_basef=32, Imagef_outf=16, Imagef_type="resnet_6blocks", w_real_l1=0.1, use_D=False, pool_size=50,
)


def default_opt(**overrides):
    d = dict(_DEFAULTS)
    d.update(overrides)
    return SimpleNamespace(**d)


def i2d_flags(**overrides):
    """README.md:28 - the published Image Guidance Network (``--model I2D``) training command."""
    d = dict(model="I2D", model_type="I2D", w_real_l1=1.0, w_syn_l1=1.0, lr=0.0002, Imagef_outf=128, Imagef_basef=32,
             norm_loss=True, batch_size=12, crop_size_h=256, crop_size_w=256)
    d.update(overrides)
    return default_opt(**d)


def translation_flags(**overrides):
    """README.md:51 - the published translation_block training command (normal init for the parity fixtures) with the
    defaults of TranslationModel.modify_commandline_options (translation_model.py:14-43)."""
    d = dict(model="translation_block", model_type="translation", netD="n_layers", n_layers_D=3, norm_d="none", ndf=64, lr=0.0002,
             beta1=0.5, max_distance=5100, batch_size=6, crop_size_h=256, crop_size_w=256, use_spnorm=False,
             l_cycle_A_begin=10.0, l_cycle_A_end=10.0, l_cycle_B_begin=5.0, l_cycle_B_end=5.0, l_identity=1.0, l_normal=1.0,
             l_depth_A_begin=5.0, l_depth_A_end=0.0, l_depth_B_begin=5.0, l_depth_B_end=0.0, l_mean_A=0.0, l_mean_B=0.0, l_tv_A=0.0,
             l_max_iter=5000, l_num_iter=5000, num_iter_gen=3, num_iter_dis=1, no_idt_A=True, use_cycle_A=False, use_cycle_B=True,
             disc_for_normals=True, disc_for_depth=True, inp_B="img_depth", w_decay_G=0.0001)
    d.update(overrides)
    return default_opt(**d)


def main_flags(**overrides):
    """README.md:70 - the published main_network_best training command."""
    d = dict(use_image_for_trans=True, w_syn_l1=15.0, w_real_l1_d=40.0, norm_loss=True, w_syn_norm=2.0,
             use_smooth_loss=True, w_smooth=1.0, w_syn_holes=800.0, w_real_holes=1600.0, use_masked=True,
             lr=0.0001, batch_size=6, crop_size_h=256, crop_size_w=256)
    d.update(overrides)
    return default_opt(**d)
