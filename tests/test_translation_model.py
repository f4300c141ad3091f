"""``TranslationModel`` (SURVEY.md section 8f rank 3 / BASELINE configs[4]): the oracle against golden vectors of one
optimize_parameters call of the live reference (CPU), and the CUDA model against golden + oracle (GPU), main-step gates."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import ref_translation
from tests_proj import proj_vec
from util import cosine, load_golden, rel_l2

NETS = ["G_A", "G_B", "D_A_depth", "D_B_depth", "D_A_normal", "D_B_normal"]
DISCS = ["D_A_depth", "D_A_normal", "D_B_depth", "D_B_normal"]


def translation_batch(B, H, W, seed=4):          # == tests/golden/make_golden.py:translation_batch
    g = torch.Generator().manual_seed(seed)
    d = lambda: torch.rand(B, 1, H, W, generator=g) * 1.6 - 0.7
    A_d, B_d = d(), d()
    A_d[:, :, 5:9, 10:20] = -1.0
    return dict(A_name=["a"] * B, B_name=["b"] * B, A_img=torch.rand(B, 3, H, W, generator=g) * 2 - 1, A_depth=A_d,
                B_img=torch.rand(B, 3, H, W, generator=g) * 2 - 1, B_depth=B_d)


def _model(gpu_ids=(), **flags):
    from dsr_b200 import options, translation_model
    opt = options.translation_flags(gpu_ids=[], batch_size=1, crop_size_h=64, crop_size_w=64, num_iter_gen=2, name="t",
                                    checkpoints_dir="/tmp/dsr_ck", **flags)
    torch.manual_seed(0)
    host = translation_model.TranslationModel(opt)
    sds = {n: {k: v.detach().clone() for k, v in getattr(host, "net" + n).state_dict().items()} for n in NETS}
    if not gpu_ids:
        return host, sds
    opt.gpu_ids = list(gpu_ids)
    dev = translation_model.TranslationModel(opt)
    for n, sd in sds.items():
        dev._unwrap(getattr(dev, "net" + n)).load_state_dict(sd)
    return dev, sds


def test_translation_oracle_matches_reference():
    g = load_golden("translation_step_b1_64.npz")
    host, sds = _model()
    for n in NETS:                                   # same seed + constructor order => the golden run's weights
        assert list(sds[n].keys()) == list(g["wkeys/" + n]), n
        a = float(sum(v.double().abs().sum() for v in sds[n].values()))
        assert abs(a - float(g["wsum/" + n][0])) <= 1e-9 * a, n
    orc = ref_translation.OracleTranslationStep(sds, num_iter_gen=2)
    out = orc.step(translation_batch(1, 64, 64))
    f = out["first"]
    for k in ("fake_depth_B", "fake_depth_A", "rec_depth_B", "idt_B", "fake_norm_B", "real_norm_A"):
        assert rel_l2(f["tensors"][k], g["s0/" + k]) <= 2e-5, k
    for k, v in f["losses"].items():
        assert abs(v - float(g["s0/loss/" + k])) <= 2e-5 * abs(float(g["s0/loss/" + k])), (k, v)
    gi = 0
    for name in ("G_A", "G_B"):
        for n in sds[name]:
            ref_norm, ref_proj = g[f"s0/g/{name}/{n}"]
            gr = f["grads"][(name, n)].double().flatten()
            assert abs(float(gr.norm()) - ref_norm) <= 2e-3 * ref_norm, (name, n)
            assert abs(float(gr @ proj_vec(gr.numel(), 4000 + gi)) - ref_proj) <= 2e-3 * ref_norm, (name, n)
            gi += 1
    for k in DISCS + ["G_A", "G_B", "cycle_B", "cycle_n_B", "idt_B", "depth_range_A", "depth_range_B"]:
        assert abs(out["losses"][k] - float(g["end/loss/" + k])) <= 2e-3 * abs(float(g["end/loss/" + k])), k
    for n in NETS:                                   # weights after the whole call: both Adam variants (weight decay on G)
        v = torch.cat([t.detach().double().flatten() for t in orc.sd[n].values()])
        ref_norm, ref_proj = g["end/w/" + n]
        assert abs(float(v.norm()) - ref_norm) <= 1e-6 * ref_norm and abs(float(v @ proj_vec(v.numel(), 6000)) - ref_proj) <= 1e-4 * ref_norm, n


def test_translation_oracle_optional_loss_terms_match_reference():
    """use_cycle_A + l_mean_A / l_mean_B + l_tv_A (translation_model.py:222-249; MaskedCosSimLoss with its 1e+6 denominator,
    MaskedMeanDif, TV_norm on the first two normal components): the oracle against one optimize_parameters call of the live
    reference run with those flags (tests/golden/make_golden.py translation_flags)."""
    g = load_golden("translation_flags_b1_64.npz")
    host, sds = _model()
    for n in NETS:
        a = float(sum(v.double().abs().sum() for v in sds[n].values()))
        assert abs(a - float(g["wsum/" + n][0])) <= 1e-9 * a, n
    orc = ref_translation.OracleTranslationStep(sds, num_iter_gen=2, use_cycle_A=True, l_cycle_A=10.0, l_mean_A=0.5, l_mean_B=0.7,
                                                l_tv_A=2.0)
    out = orc.step(translation_batch(1, 64, 64))
    f = out["first"]
    assert rel_l2(f["tensors"]["rec_depth_A"], g["s0/rec_depth_A"]) <= 2e-5
    for k in ("cycle_A", "cycle_n_A", "mean_dif_A", "mean_dif_B", "tv_norm_A", "G", "cycle_B", "depth_range_A"):
        assert abs(f["losses"][k] - float(g["s0/loss/" + k])) <= 2e-5 * abs(float(g["s0/loss/" + k])), (k, f["losses"][k])
    gi = 0
    for name in ("G_A", "G_B"):
        for n in sds[name]:
            ref_norm, ref_proj = g[f"s0/g/{name}/{n}"]
            gr = f["grads"][(name, n)].double().flatten()
            assert abs(float(gr.norm()) - ref_norm) <= 2e-3 * ref_norm, (name, n)
            assert abs(float(gr @ proj_vec(gr.numel(), 4000 + gi)) - ref_proj) <= 2e-3 * ref_norm, (name, n)
            gi += 1
    for k in DISCS + ["cycle_A", "cycle_n_A", "mean_dif_A", "mean_dif_B", "tv_norm_A"]:
        assert abs(out["losses"][k] - float(g["end/loss/" + k])) <= 2e-3 * abs(float(g["end/loss/" + k])), k
    for n in NETS:
        v = torch.cat([t.detach().double().flatten() for t in orc.sd[n].values()])
        ref_norm, ref_proj = g["end/w/" + n]
        assert abs(float(v.norm()) - ref_norm) <= 1e-6 * ref_norm and abs(float(v @ proj_vec(v.numel(), 6000)) - ref_proj) <= 1e-4 * ref_norm, n


@pytest.mark.gpu
def test_fov_normals_and_cos_sim_ops(built_lib):
    from dsr_b200 import ops
    g = torch.Generator().manual_seed(8)
    for (H, W) in ((40, 56), (64, 64), (2, 3)):
        d = (torch.rand(2, 1, H, W, generator=g) * 1.6 - 0.7).requires_grad_(True)
        go = torch.randn(2, 3, H, W, generator=g)
        ref = ref_translation.fov_normals(d)
        (ref * go).sum().backward()
        dc = d.detach().cuda().requires_grad_(True)
        out = ops.fov_normals(dc)
        (out * go.cuda()).sum().backward()
        assert float((out.detach().cpu() - ref.detach()).abs().max()) <= 2e-5          # unit normals, fp32 cross products
        assert rel_l2(dc.grad.cpu(), d.grad) <= 2e-3
    x = torch.randn(2, 3, 17, 23, generator=g).requires_grad_(True)
    y = torch.randn(2, 3, 17, 23, generator=g)
    ref = ref_translation.cos_sim_loss(x, y)
    ref.backward()
    xc = x.detach().cuda().requires_grad_(True)
    out = ops.cos_sim_loss(xc, y.cuda())
    out.backward()
    assert abs(float(out) - float(ref)) <= 1e-6 and rel_l2(xc.grad.cpu(), x.grad) <= 1e-5


@pytest.mark.gpu
def test_translation_step_matches_reference_golden_and_oracle(built_lib):
    g = load_golden("translation_step_b1_64.npz")
    model, sds = _model(gpu_ids=[0])
    batch = translation_batch(1, 64, 64)
    orc = ref_translation.OracleTranslationStep(sds, num_iter_gen=2)
    ref = orc.step(batch)
    # first generator iteration by hand (forward + backward_G) to compare gradients before Adam moves the weights
    model.set_input(batch)
    model.set_requires_grad(model.disc, False)
    model.forward()
    model.optimizer_G.zero_grad()
    model.backward_G()
    for k in ("fake_depth_B", "fake_depth_A", "rec_depth_B", "idt_B"):
        assert rel_l2(getattr(model, k).detach().cpu(), g["s0/" + k]) <= 1e-2, k                     # the gate
        assert rel_l2(getattr(model, k).detach().cpu(), ref["first"]["tensors"][k]) <= 2e-3, k
    for k in ("fake_norm_B", "real_norm_A"):
        assert rel_l2(getattr(model, k).detach().cpu(), g["s0/" + k]) <= 1e-2, k
    for k in ("G_A", "G_B", "cycle_B", "cycle_n_B", "idt_B", "depth_range_A", "depth_range_B"):
        v, want = float(getattr(model, "loss_" + k)), float(g["s0/loss/" + k])
        assert abs(v - want) <= 1e-3 * abs(want), (k, v, want)
    assert abs(float(model.loss_G) - float(g["s0/loss/G"])) <= 1e-3 * float(g["s0/loss/G"])
    fa, fb = [], []
    for name in ("G_A", "G_B"):
        for n, prm in model._unwrap(getattr(model, "net" + name)).named_parameters():
            gr = ref["first"]["grads"].get((name, n))
            if gr is None or float(gr.norm()) == 0.0:
                continue
            c = cosine(prm.grad.detach().cpu(), gr)
            assert c >= 0.999, (name, n, c)
            fa.append(prm.grad.detach().cpu().flatten()); fb.append(gr.flatten())
    assert cosine(torch.cat(fa), torch.cat(fb)) >= 0.999
    model.set_requires_grad(model.disc, True)
    # the whole call from the same initial weights on a fresh model: losses at the end and the weights after both updates
    model2, _ = _model(gpu_ids=[0])
    model2.set_input(batch)
    model2.optimize_parameters(0)
    for k in DISCS + ["G_A", "G_B", "cycle_B", "cycle_n_B", "idt_B", "depth_range_A", "depth_range_B"]:
        v, want = float(getattr(model2, "loss_" + k)), float(g["end/loss/" + k])
        assert abs(v - want) <= 1e-2 * abs(want), (k, v, want)        # behind one Adam step: gate-level tolerance (as in the main step)
    for n in NETS:
        net = model2._unwrap(getattr(model2, "net" + n))
        v = torch.cat([t.detach().double().flatten().cpu() for t in net.state_dict().values()])
        ref_norm, ref_proj = g["end/w/" + n]
        assert abs(float(v.norm()) - ref_norm) <= 1e-5 * ref_norm, n
    losses = model2.get_current_losses()
    assert all(np.isfinite(v) for v in losses.values()) and set(model2.loss_names) == set(losses)


def test_translation_oracle_tv_term_matches_reference():
    g = load_golden("translation_tv_b1_64.npz")
    host, sds = _model(l_tv_A=2.0)
    assert "tv_norm_A" in host.loss_names
    orc = ref_translation.OracleTranslationStep(sds, num_iter_gen=2, l_tv_A=2.0)
    out = orc.step(translation_batch(1, 64, 64))
    for k in ("tv_norm_A", "G", "G_A", "cycle_B"):
        assert abs(out["first"]["losses"][k] - float(g["s0/loss/" + k])) <= 2e-5 * abs(float(g["s0/loss/" + k])), k
    assert abs(out["losses"]["tv_norm_A"] - float(g["end/loss/tv_norm_A"])) <= 2e-3 * abs(float(g["end/loss/tv_norm_A"]))
    for n in NETS:
        v = torch.cat([t.detach().double().flatten() for t in orc.sd[n].values()])
        ref_norm, ref_proj = g["end/w/" + n]
        assert abs(float(v.norm()) - ref_norm) <= 1e-6 * ref_norm, n


@pytest.mark.gpu
def test_translation_step_with_tv_norm_term(built_lib):
    """--l_tv_A 2.0 (translation_model.py:247-249): the CUDA model against the live-reference golden and the oracle - first
    generator iteration (losses, gradients), then the whole call (end losses, weights after both Adam updates)."""
    g = load_golden("translation_tv_b1_64.npz")
    model, sds = _model(gpu_ids=[0], l_tv_A=2.0)
    batch = translation_batch(1, 64, 64)
    ref = ref_translation.OracleTranslationStep(sds, num_iter_gen=2, l_tv_A=2.0).step(batch)
    model.set_input(batch)
    model.set_requires_grad(model.disc, False)
    model.forward()
    model.optimizer_G.zero_grad()
    model.backward_G()
    for k in ("tv_norm_A", "G_A", "G_B", "cycle_B", "cycle_n_B", "idt_B", "depth_range_A", "depth_range_B"):
        v, want = float(getattr(model, "loss_" + k)), float(g["s0/loss/" + k])
        assert abs(v - want) <= 1e-3 * abs(want), (k, v, want)
    assert abs(float(model.loss_G) - float(g["s0/loss/G"])) <= 1e-3 * float(g["s0/loss/G"])
    fa, fb = [], []
    for n, prm in model._unwrap(model.netG_A).named_parameters():          # the TV term reaches G_A only
        gr = ref["first"]["grads"].get(("G_A", n))
        if gr is None or float(gr.norm()) == 0.0:
            continue
        assert cosine(prm.grad.detach().cpu(), gr) >= 0.999, n
        fa.append(prm.grad.detach().cpu().flatten()); fb.append(gr.flatten())
    assert cosine(torch.cat(fa), torch.cat(fb)) >= 0.999
    model.set_requires_grad(model.disc, True)
    model2, _ = _model(gpu_ids=[0], l_tv_A=2.0)
    model2.set_input(batch)
    model2.optimize_parameters(0)
    v, want = float(model2.loss_tv_norm_A), float(g["end/loss/tv_norm_A"])
    assert abs(v - want) <= 1e-2 * abs(want)
    for n in ("G_A", "G_B"):
        net = model2._unwrap(getattr(model2, "net" + n))
        w = torch.cat([t.detach().double().flatten().cpu() for t in net.state_dict().values()])
        assert abs(float(w.norm()) - g["end/w/" + n][0]) <= 1e-5 * g["end/w/" + n][0], n
    assert set(model2.loss_names) == set(model2.get_current_losses())


@pytest.mark.gpu
def test_graph_replay_follows_the_loss_weight_schedule(built_lib):
    """update_loss_weight (translation_model.py:300-305) after the step was captured: the replayed graph multiplies by the
    device-resident weights, so it tracks the schedule like the eager step does"""
    import numpy as np
    import torch
    from dsr_b200 import ops, options, translation_model
    from oracle import ref_step

    def make(graph):
        torch.manual_seed(0)
        opt = options.translation_flags(gpu_ids=[0], batch_size=1, crop_size_h=64, crop_size_w=64, name="t", checkpoints_dir="/tmp/dsr_ck",
                                        cuda_graph=graph, l_max_iter=0, l_num_iter=4, lr=0.0)     # lr 0: the weights stay put,
        # so the loss terms change through the schedule only
        return translation_model.TranslationModel(opt)

    b = ref_step.synthetic_batch(1, 64, 64, seed=1, depth_kind="smooth")
    batch = dict(A_name=b["A_paths"], B_name=b["B_paths"], A_img=b["A_i"], A_depth=b["A_d"], B_img=b["B_i"], B_depth=b["B_d"])
    out = {}
    for graph in (True, False):
        m = make(graph)
        vals, captured_mid = [], False
        for it in range(6):
            m.set_input(batch)
            m.optimize_parameters(it, 1)
            captured_mid = captured_mid or m._graph is not None     # (the graph is dropped again when a term leaves the loss)
            vals.append((float(m.loss_depth_range_A), float(m.loss_cycle_B)) if m.l_depth_A > 0 else (0.0, float(m.loss_cycle_B)))
            m.update_loss_weight(it + 1)                 # l_max_iter = 0: the weights move after every step
        out[graph] = (vals, m.l_depth_A, captured_mid)
    (vg, lg, captured), (ve, le, _) = out[True], out[False]
    assert captured and lg == le
    assert ve[0][0] > 0 and ve[3][0] < ve[0][0]          # the depth-range weight decays 5 -> 0 over l_num_iter = 4 updates
    for it, (a, e) in enumerate(zip(vg, ve)):
        assert abs(a[0] - e[0]) <= 1e-3 * max(abs(e[0]), 1e-3) and abs(a[1] - e[1]) <= 1e-3 * max(abs(e[1]), 1e-3), (vg, ve)
        if it < 4:                                        # weight after `it` updates = 5 - 1.25 it: the replay follows it
            assert abs(a[0] / vg[0][0] - (5 - 1.25 * it) / 5) <= 1e-3, (it, vg)


FLAG_SETS = {
    "flags": (dict(use_cycle_A=True, l_mean_A=0.5, l_mean_B=0.7, l_tv_A=2.0), dict(use_cycle_A=True, l_cycle_A=10.0, l_mean_A=0.5, l_mean_B=0.7, l_tv_A=2.0),
              "translation_flags_b1_64.npz", ("cycle_A", "cycle_n_A", "mean_dif_A", "mean_dif_B", "tv_norm_A")),
    "inpB": (dict(use_cycle_A=True, inp_B="depth"), dict(use_cycle_A=True, l_cycle_A=10.0, inp_B="depth"),
             "translation_inpB_b1_64.npz", ("cycle_A", "cycle_n_A")),
}


def test_translation_oracle_depth_only_G_B_matches_reference():
    """--inp_B depth --use_cycle_A (translation_model.py:146-147, :167-168, :185-186): the oracle against the live reference"""
    model_flags, orc_flags, golden, extra = FLAG_SETS["inpB"]
    g = load_golden(golden)
    host, sds = _model(**model_flags)
    assert "enc_img.model.0.weight" not in sds["G_B"] and sds["G_B"]["enc_depth.model.0.weight"].shape[0] == 64
    for n in NETS:
        a = float(sum(v.double().abs().sum() for v in sds[n].values()))
        assert abs(a - float(g["wsum/" + n][0])) <= 1e-9 * a, n
    orc = ref_translation.OracleTranslationStep(sds, num_iter_gen=2, **orc_flags)
    out = orc.step(translation_batch(1, 64, 64))
    f = out["first"]
    assert rel_l2(f["tensors"]["rec_depth_A"], g["s0/rec_depth_A"]) <= 2e-5
    for k in extra + ("G", "G_A", "G_B", "cycle_B", "idt_B", "depth_range_B"):
        assert abs(f["losses"][k] - float(g["s0/loss/" + k])) <= 2e-5 * abs(float(g["s0/loss/" + k])), (k, f["losses"][k])
    for n in NETS:
        v = torch.cat([t.detach().double().flatten() for t in orc.sd[n].values()])
        ref_norm, ref_proj = g["end/w/" + n]
        assert abs(float(v.norm()) - ref_norm) <= 1e-6 * ref_norm, n


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["flags", "inpB"])
def test_translation_step_optional_loss_terms_on_gpu(built_lib, which):
    """use_cycle_A (MaskedL1 + MaskedCosSim with its 1e+6 denominator), l_mean_A / l_mean_B (MaskedMeanDif), l_tv_A and the
    depth-only G_B (translation_model.py:146-171, :222-249): the CUDA model against the live-reference golden and the oracle -
    first generator iteration (losses, gradients), then the whole call (end losses, weights after both Adam updates)."""
    model_flags, orc_flags, golden, extra = FLAG_SETS[which]
    g = load_golden(golden)
    model, sds = _model(gpu_ids=[0], **model_flags)
    batch = translation_batch(1, 64, 64)
    ref = ref_translation.OracleTranslationStep(sds, num_iter_gen=2, **orc_flags).step(batch)
    model.set_input(batch)
    model.set_requires_grad(model.disc, False)
    model.forward()
    model.optimizer_G.zero_grad()
    model.backward_G()
    assert rel_l2(model.rec_depth_A.detach().cpu(), g["s0/rec_depth_A"]) <= 1e-2
    for k in extra + ("G_A", "G_B", "cycle_B", "cycle_n_B", "idt_B", "depth_range_A", "depth_range_B"):
        v, want = float(getattr(model, "loss_" + k)), float(g["s0/loss/" + k])
        assert abs(v - want) <= 1e-3 * abs(want), (k, v, want)
    assert abs(float(model.loss_G) - float(g["s0/loss/G"])) <= 1e-3 * float(g["s0/loss/G"])
    fa, fb = [], []
    for name in ("G_A", "G_B"):
        for n, prm in model._unwrap(getattr(model, "net" + name)).named_parameters():
            gr = ref["first"]["grads"].get((name, n))
            if gr is None or float(gr.norm()) == 0.0:
                continue
            c = cosine(prm.grad.detach().cpu(), gr)
            assert c >= 0.999, (name, n, c)
            fa.append(prm.grad.detach().cpu().flatten()); fb.append(gr.flatten())
    assert cosine(torch.cat(fa), torch.cat(fb)) >= 0.999
    model.set_requires_grad(model.disc, True)
    model2, _ = _model(gpu_ids=[0], **model_flags)
    model2.set_input(batch)
    model2.optimize_parameters(0)
    for k in extra:
        v, want = float(getattr(model2, "loss_" + k)), float(g["end/loss/" + k])
        assert abs(v - want) <= 1e-2 * abs(want), (k, v, want)
    for n in NETS:
        net = model2._unwrap(getattr(model2, "net" + n))
        w = torch.cat([t.detach().double().flatten().cpu() for t in net.state_dict().values()])
        assert abs(float(w.norm()) - g["end/w/" + n][0]) <= 1e-5 * g["end/w/" + n][0], n
    assert set(model2.loss_names) == set(model2.get_current_losses())
    assert "rec_depth_A" in model2.get_current_visuals()
