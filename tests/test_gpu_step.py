"""GPU parity of the networks and of the whole optimize_parameters step, through the public API
(define_G / define_Gen / MainModel.set_input / optimize_parameters), against (a) the golden vectors
recorded from the live reference and (b) the oracle on the same seeded inputs and weights.
Gates (BASELINE.json north_star): input-derived masks bit-exact; pred rel-L2 <= 1e-2; each loss within
1e-3 relative; parameter gradients cosine >= 0.999."""
import numpy as np
import pytest
import torch

from oracle import ref_nets, ref_step
from util import (build_host_model, cosine, grad_is_informative, load_golden, rehome, rel_l2, state_dicts)

pytestmark = pytest.mark.gpu

# per-network forward tolerance in the default precision mode (f16 hi+lo operands = 22-bit significands, 3 MMAs per product)
NET_TOL = 1e-4


@pytest.fixture(scope="module")
def model_and_oracle(built_lib):
    host = build_host_model(2, 128, 128)
    sds = state_dicts(host)
    dev = rehome(host, host.opt, [0])
    return dev, sds


def test_networks_forward_match_oracle(built_lib):
    from dsr_b200 import networks, translation_network
    from types import SimpleNamespace
    g = torch.Generator().manual_seed(5)
    torch.manual_seed(3)
    res = networks.define_G(3, 128, 32, "resnet_6blocks", "instance", False, "normal", 0.02, [])
    x = torch.rand(2, 3, 64, 48, generator=g) * 2 - 1
    ref = ref_nets.resnet_generator(res.state_dict(), x)
    with torch.no_grad():
        out = res.cuda()(x.cuda())
    assert rel_l2(out.cpu(), ref) <= NET_TOL
    unet = networks.define_G(128, 1, 64, "unet_128", "instance", False, "normal", 0.02, [])
    f = torch.rand(1, 128, 128, 256, generator=g) * 2 - 1
    ref = ref_nets.unet_generator(unet.state_dict(), f)
    with torch.no_grad():
        out = unet.cuda()(f.cuda())
    assert rel_l2(out.cpu(), ref) <= NET_TOL
    o = SimpleNamespace(ngf_img=32, ngf_depth=32, ngf=64, norm="group", dropout=False, init_type="normal", gpu_ids=[],
                        input_nc_img=3, n_downsampling=2, use_semantic=False, n_blocks=9, upsampling_type="transpose",
                        output_nc_depth=1, input_nc_depth=1)
    gen = translation_network.define_Gen(o, input_type="img_depth")
    d = torch.rand(2, 1, 32, 48, generator=g) * 2 - 1
    im = torch.rand(2, 3, 32, 48, generator=g) * 2 - 1
    ref = ref_nets.translation_generator(gen.state_dict(), d, im)
    with torch.no_grad():
        out = gen.cuda()(d.cuda(), im.cuda())
    assert rel_l2(out.cpu(), ref) <= NET_TOL


def _check_step(model, out_losses, ref_losses, tol):
    for k, ref in ref_losses.items():
        if k == "G":
            v = float(model.loss_G)
        elif k.startswith("mean_of_abs"):
            v = float(getattr(model, "loss_" + k))
        else:
            v = out_losses[k]
        assert abs(v - ref) <= tol * max(abs(ref), 1e-3), (k, v, ref)


def test_step_matches_reference_golden_and_oracle(model_and_oracle):
    model, sds = model_and_oracle
    g = load_golden("step_b2_128.npz")
    batch = ref_step.synthetic_batch(2, 128, 128, seed=1, depth_kind="smooth")
    orc = ref_step.OracleStep(sds, lr=1e-4)
    model._train()
    np.random.seed(0)
    rng_state = np.random.get_state()
    for it in range(2):
        np.random.set_state(rng_state)
        ref = orc.step(batch)
        np.random.set_state(rng_state)
        model.set_input(batch)
        model.optimize_parameters(it, 1)
        rng_state = np.random.get_state()
        p = f"s{it}/"
        if it == 0:                                                   # integer / index work: bit-exact
            for k in ("syn_mask", "real_mask", "real_hole_mask"):
                assert np.array_equal(getattr(model, k).cpu().numpy().astype(np.uint8), g[p + k]), k
            assert np.array_equal(model.gt_mask_real.cpu().numpy(), g[p + "gt_mask_real"])
            assert np.array_equal(model.gt_mask_syn.cpu().numpy(), g[p + "gt_mask_syn"])
        for k in ("syn2real_depth", "syn_depth_by_image", "real_depth_by_image", "pred_syn_depth", "pred_real_depth"):
            assert rel_l2(getattr(model, k).detach().cpu(), g[p + k]) <= 1e-2, (k, it)          # the gate
            # step 2 sits behind one Adam update, which turns the sign of every rounding-noise gradient
            # into a full +-lr move (m/sqrt(v) = +-1 at t=1): only the gate-level tolerance is meaningful there
            assert rel_l2(getattr(model, k).detach().cpu(), ref["tensors"][k].detach()) <= (2e-3 if it == 0 else 1e-2), (k, it)
        losses = model.get_current_losses()
        gold = {k[len(p) + 5:]: float(g[k]) for k in g.files if k.startswith(p + "loss/")}
        _check_step(model, losses, gold, 1e-3 if it == 0 else 1e-2)
        if it == 0:
            _check_step(model, losses, ref["losses"], 1e-3)
            # gradients (read from the arena views before Adam consumed them? Adam does not modify grads)
            flat_a, flat_b, worst = [], [], 1.0
            for net in ("Depth_f", "Task"):
                params = dict(model._unwrap(getattr(model, "net" + net)).named_parameters())
                for n, gr in ((n, ref["grads"][(net, n)]) for n in orc.sd[net]):
                    mine = params[n].grad.detach().cpu()
                    if grad_is_informative(net, n):
                        c = cosine(mine, gr)
                        worst = min(worst, c)
                        assert c >= 0.999, (net, n, c)
                        flat_a.append(mine.flatten()); flat_b.append(gr.flatten())
                        if f"{p}gfull/{net}/{n}" in g.files:
                            assert cosine(mine, g[f"{p}gfull/{net}/{n}"]) >= 0.999, (net, n)
                    else:
                        assert float(mine.norm()) <= 1e-2 * max(float(torch.cat(flat_b).norm()) if flat_b else 1.0, 1e-3)
            assert cosine(torch.cat(flat_a), torch.cat(flat_b)) >= 0.999
    # weights after two Adam steps
    for net in ("Depth_f", "Task"):
        params = dict(model._unwrap(getattr(model, "net" + net)).named_parameters())
        a = torch.cat([params[n].detach().cpu().flatten() for n in orc.sd[net]])
        b = torch.cat([orc.sd[net][n].detach().flatten() for n in orc.sd[net]])
        assert rel_l2(a, b) <= 5e-3


def test_cuda_graph_replay_matches_eager(built_lib):
    """One replay of the captured step (static buffers, device-side Adam state, host RNG outside the graph) against one
    eager step FROM THE SAME STATE.  (Whole trajectories cannot be compared: fp32 atomics make two eager runs differ by
    ~1e-5 after one Adam step, and the first Adam steps move every weight by +-lr whatever the gradient size, so that
    noise grows several-fold per step.)"""
    from dsr_b200 import ops
    host = build_host_model(2, 128, 128)
    batches = [ref_step.synthetic_batch(2, 128, 128, seed=s, depth_kind="smooth") for s in (1, 2)]
    m = rehome(host, host.opt, [0])
    m.use_graph = True
    m._train()
    np.random.seed(7)
    for it in range(4):                                 # 2 eager warm-up steps, capture + replay, one more replay
        m.set_input(batches[it % 2])
        m.optimize_parameters(it, 1)
    assert (m._graph is not None or m._pipe is not None) and m.optimizer_G.n_steps == 4
    state = (m.arena.flat, m.arena.exp_avg, m.arena.exp_avg_sq, m.optimizer_G.step_dev)
    snap = [t.clone() for t in state]
    out = []
    for use_graph in (True, False):
        for t, s0 in zip(state, snap):
            t.copy_(s0)
        ops.WEIGHT_EPOCH += 1
        m.use_graph = use_graph
        np.random.seed(11)
        m.set_input(batches[0])
        m.optimize_parameters(9, 1)
        out.append((float(m.loss_G), m.pred_real_depth.detach().clone(), m.arena.grad.clone(), m.arena.flat.clone()))
    (la, pa, ga, wa), (lb, pb, gb, wb) = out
    assert abs(la - lb) <= 1e-5 * abs(lb), (la, lb)
    assert rel_l2(pa.cpu(), pb.cpu()) <= 1e-5
    # two EAGER runs from the same state already differ by 1 - cos = 2e-6 .. 5e-6 (scripts/noise_probe.py: atomics sum in
    # another order every run and the nets amplify that ~1e4-fold), with occasional 5e-5 outliers: gate at 1e-4
    assert cosine(ga.cpu(), gb.cpu()) >= 0.9999
    assert float((wa - wb).abs().max()) <= 2.1e-4          # at most a sign flip of a noise-level gradient: 2 * lr
    assert m.optimizer_G.n_steps == 5


def test_pipelined_replay_matches_single_graph_replay(built_lib):
    """The pipelined loop (frozen networks of batch i+1 replayed beside the training part of batch i, two input slots) against
    the single-graph loop on the same sequence of DIFFERENT batches: same losses step by step (first step to fp32 noise, later
    steps within the growth of that noise through Adam), and the result attributes always belong to the batch just trained."""
    from dsr_b200 import ops
    host = build_host_model(2, 128, 128)
    batches = [ref_step.synthetic_batch(2, 128, 128, seed=s, depth_kind="smooth") for s in (1, 2, 3)]
    runs = []
    for pipe in (True, False):
        ops.CONFIG["pipeline_frozen"] = pipe
        try:
            m = rehome(host, host.opt, [0])
            m.use_graph = True
            m._train()
            np.random.seed(7)
            losses = []
            for it in range(7):                          # 2 eager steps, 2 capturing steps (one per slot), 3 pure replays
                m.set_input(batches[it % 3])
                m.optimize_parameters(it, 1)
                losses.append(m.loss_G.detach().clone())     # device-side copy, no synchronisation: the host keeps running ahead
            assert (m._pipe is not None) == pipe
            # the attributes of the last step belong to the last batch's slot
            last = (float(m.loss_G), m.pred_real_depth.detach().clone(), m.real_depth.detach().clone())
            runs.append((last, m.optimizer_G.n_steps, [float(x) for x in losses]))
            m.reset_graph()
        finally:
            ops.CONFIG["pipeline_frozen"] = True
    (la, pa, da), na, seq_a = runs[0]
    (lb, pb, db), nb, seq_b = runs[1]
    assert na == nb == 7
    assert abs(seq_a[0] - seq_b[0]) <= 1e-5 * abs(seq_b[0])
    for x, y in zip(seq_a, seq_b):
        assert abs(x - y) <= 5e-3 * abs(y), (seq_a, seq_b)
    assert torch.equal(da.cpu(), batches[6 % 3]["B_d"]) and torch.equal(db.cpu(), batches[6 % 3]["B_d"])
    assert abs(la - lb) <= 5e-3 * abs(lb), (la, lb)              # seven Adam steps of amplified atomics noise
    # (the prediction itself is not compared after seven steps: at this initialisation it is a noise-level map whose relative
    # L2 between two EAGER runs is already ~0.1 by then; test_cuda_graph_replay_matches_eager pins one pipelined replay
    # against one eager step from the same state to 1e-5)
    assert pa.shape == pb.shape


def test_pipelined_replay_survives_irregular_call_patterns(built_lib):
    """set_input twice before a step (a skipped batch), a step repeated on the same batch, an eager validation pass in the middle
    of training and a graph reset: the pipelined loop must keep training on the batch of the LAST set_input"""
    host = build_host_model(1, 128, 128)
    m = rehome(host, host.opt, [0])
    m.use_graph = True
    m._train()
    b = [ref_step.synthetic_batch(1, 128, 128, seed=s, depth_kind="smooth") for s in (1, 2, 3)]
    np.random.seed(5)
    for it in range(5):                                  # warm-up, both slots captured, one pure replay
        m.set_input(b[it % 3])
        m.optimize_parameters(it, 1)
    assert m._pipe is not None and all(sl["gt"] is not None for sl in m._pipe["slots"])
    m.set_input(b[0])
    m.set_input(b[1])                                    # b[0] is skipped
    m.optimize_parameters(5, 1)
    assert torch.equal(m.real_depth.cpu(), b[1]["B_d"]) and np.isfinite(float(m.loss_G))
    l1 = float(m.loss_G)
    m.optimize_parameters(6, 1)                          # the same batch again (weights moved: a different loss)
    assert torch.equal(m.real_depth.cpu(), b[1]["B_d"]) and np.isfinite(float(m.loss_G)) and float(m.loss_G) != l1
    with torch.no_grad():                                # validation between two training steps
        m.set_input(b[2])
        m.forward("test")
        pred_eager = m.pred_real_depth.detach().clone()
        m.forward_test_graph()                           # (falls back to the eager pass on a pipelined model)
        assert rel_l2(m.pred_real_depth.cpu(), pred_eager.cpu()) <= 1e-5
    m.set_input(b[0])
    m.optimize_parameters(7, 1)
    assert torch.equal(m.real_depth.cpu(), b[0]["B_d"]) and m.optimizer_G.n_steps == 8
    m.reset_graph()
    assert m._pipe is None
    m.set_input(b[1])
    m.optimize_parameters(8, 1)                          # starts over: eager warm-up steps, new graphs later
    assert m.optimizer_G.n_steps == 9 and m.nonfinite_steps() == 0


def test_nonfinite_loss_is_counted_on_the_device(built_lib):
    """a NaN in the batch must not pass silently through a (replayed) step: the loss kernel counts non-finite loss_G values"""
    host = build_host_model(1, 128, 128)
    m = rehome(host, host.opt, [0])
    m._train()
    good = ref_step.synthetic_batch(1, 128, 128, seed=3, depth_kind="smooth")
    bad = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in good.items()}
    bad["B_d"][0, 0, 5, 7] = float("nan")
    np.random.seed(3)
    m.set_input(good)
    m.optimize_parameters(0, 1)
    assert m.nonfinite_steps() == 0
    m.check_finite()
    m.set_input(bad)
    m.optimize_parameters(1, 1)
    assert m.nonfinite_steps() == 1
    with pytest.raises(FloatingPointError):
        m.check_finite()


def test_calculate_eval_mode_and_visuals(model_and_oracle):
    model, _ = model_and_oracle
    batch = ref_step.synthetic_batch(2, 128, 128, seed=2, depth_kind="noise")
    model.eval()
    with torch.no_grad():
        model.set_input(batch)
        model.calculate("test")
    vis = model.get_current_visuals()
    for k in model.visual_names:
        assert k in vis and torch.is_tensor(vis[k]) and vis[k].shape[0] == 2, k
    # test stage: p = 0 -> no artificial holes: depth_masked == real_depth
    assert torch.equal(model.depth_masked.cpu(), batch["B_d"])
    assert all(np.isfinite(v) for v in model.get_current_losses().values())
    model._train()


def test_inference_graph_replay_matches_eager(model_and_oracle):
    """forward('test') replayed as a CUDA graph (forward_test_graph) == the eager pass, for two different inputs"""
    model, _ = model_and_oracle
    model.eval()
    batches = [ref_step.synthetic_batch(2, 128, 128, seed=s, depth_kind="noise") for s in (4, 5)]
    with torch.no_grad():
        for i in range(5):                              # 2 eager, capture, 2 replays
            model.set_input(batches[i % 2])
            model.forward_test_graph()
        assert model._tgraph["graph"] is not None
        for b in batches:
            model.set_input(b)
            model.forward_test_graph()
            got = {k: getattr(model, k).detach().clone() for k in ("pred_real_depth", "pred_syn_depth", "syn2real_depth", "depth_masked")}
            model.set_input(b)
            model.forward("test")
            for k, v in got.items():
                assert rel_l2(v.cpu(), getattr(model, k).detach().cpu()) <= 1e-5, k
    model._train()


def test_checkpoint_roundtrip(model_and_oracle, tmp_path):
    model, sds = model_and_oracle
    model.save_dir = str(tmp_path)
    model.save_networks("latest")
    for name in model.model_names:
        sd = torch.load(str(tmp_path / f"latest_net_{name}.pth"), map_location="cpu")
        assert list(sd.keys()) == list(sds[name].keys())
        assert all(v.dtype == torch.float32 and v.device.type == "cpu" for v in sd.values())
    before = {n: p.detach().clone() for n, p in model.netTask.named_parameters()}
    with torch.no_grad():
        for p in model.netTask.parameters():
            p.add_(1.0)
    model.load_networks("latest", strict=True)
    for n, p in model.netTask.named_parameters():
        assert torch.equal(p.detach(), before[n])
    assert model.arena is not None and all(p.data_ptr() in model.arena.index for p in model.arena.params)


def test_inference_graph_capture_call_returns_computed_outputs(built_lib):
    """the call that CAPTURES forward('test') must also run it: its outputs equal the eager pass (they used to be
    uninitialised graph-pool memory until the next call)"""
    host = build_host_model(1, 128, 128)
    m = rehome(host, host.opt, [0])
    m.eval()
    batches = [ref_step.synthetic_batch(1, 128, 128, seed=s, depth_kind="noise") for s in (21, 22, 23)]
    with torch.no_grad():
        for i in range(3):                              # eager, eager, capture (+ replay)
            m.set_input(batches[i])
            m.forward_test_graph()
        assert m._tgraph["graph"] is not None
        got = {k: getattr(m, k).detach().clone() for k in ("pred_real_depth", "pred_syn_depth", "syn2real_depth", "depth_masked")}
        m.set_input(batches[2])
        m.forward("test")
        for k, v in got.items():
            assert rel_l2(v.cpu(), getattr(m, k).detach().cpu()) <= 1e-5, k


def test_inference_graph_sees_weights_trained_in_between(built_lib):
    """training steps between two replays of the inference graph: the trainable nets are re-packed INSIDE the graph, so the
    replay uses the current weights (not the packed copies cached at capture time, which the optimizer step invalidates
    and the allocator may hand to somebody else)"""
    host = build_host_model(1, 128, 128)
    m = rehome(host, host.opt, [0])
    b = ref_step.synthetic_batch(1, 128, 128, seed=31, depth_kind="smooth")
    np.random.seed(3)
    m.eval()
    with torch.no_grad():
        for i in range(4):
            m.set_input(b)
            m.forward_test_graph()
    before = m.pred_real_depth.detach().clone()
    m._train()
    m.optimizer_G.param_groups[0]["lr"] = 1e-2          # a visible update
    for it in range(2):
        m.set_input(b)
        m.optimize_parameters(it, 1)
    junk = [torch.full((1 << 20,), 7.0, device="cuda") for _ in range(64)]      # recycle whatever the allocator freed
    m.eval()
    with torch.no_grad():
        m.set_input(b)
        m.forward_test_graph()
        replay = m.pred_real_depth.detach().clone()
        m.set_input(b)
        m.forward("test")
        eager = m.pred_real_depth.detach().clone()
    del junk
    assert rel_l2(replay.cpu(), eager.cpu()) <= 1e-5
    assert rel_l2(before.cpu(), eager.cpu()) >= 1e-3    # the weights did move
