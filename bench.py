#!/usr/bin/env python
"""bench.py - pair-samples/s of the main_network_best training step (MainModel.optimize_parameters)
on synthetic RGB-D crops.  Contract: one JSON line on stdout from rank 0 (see DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]
N > 1 is launched by the driver through torch.distributed.run (one rank per GPU, NCCL).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "depth-enhancement-and-super-resolution_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {   # BASELINE.json configs[1] / configs[2]
    "c2": dict(B=6, H=256, W=256, name="main_network_best training step, batch 6 per GPU, 256x256 crops"),
    "c3": dict(B=3, H=512, W=640, name="main_network_best full-size 640x480 (fed as 512x640), batch 3 per GPU"),
    # BASELINE.json configs[3] at the reference's native x2 (main_sr_model.py; README.md:86 uses batch 1): H, W = LR crop
    # (the U-Net-128 of Image2Depth needs LR sides that are multiples of 128, so 256x320 -> 512x640 cannot run in the reference)
    "c4": dict(B=1, H=512, W=640, sr=True, name="main_sr_model x2 depth super-resolution step (LR 512x640 -> HR 1024x1280, README.md:86), batch 1 per GPU"),
    # BASELINE.json configs[0] (the reference's own CPU-runnable case), on the GPU: I2D_model.py, README.md:28 flags
    "c1": dict(B=2, H=256, W=256, i2d=True, name="I2D Image Guidance Network training step, batch 2, 256x256"),
    # BASELINE.json configs[4]: TranslationModel.optimize_parameters (3 generator iterations + 1 discriminator update), README.md:51
    "c5": dict(B=6, H=256, W=256, tr=True, name="translation_block TranslationModel.optimize_parameters (3 G iterations + 1 D update), batch 6, 256x256"),
    # the two network families alone (dsr_b200/translation_blocks.py): generator + PatchGAN discriminator, LSGAN G step + D step
    "c5b": dict(B=6, H=256, W=256, gan=True, name="generator + n_layers PatchGAN discriminator, LSGAN G step + D step, batch 6, 256x256"),
    "tiny": dict(B=1, H=128, W=128, name="debug"),
}
FLOP_PER_PAIR_256 = 606.2e9      # 2*174.68 + 4*64.20 GMAC-pairs
# SR step: G_A_d, Depth_f (fwd + bwd), Task (fwd + bwd) run on HR pixels, I2D_features and Image2Depth on LR pixels
FLOP_SR_HR_256 = 2e9 * (50.55 + 3 * (43.63 + 20.57))
FLOP_SR_LR_256 = 2e9 * (43.84 + 16.11)


def step_flops(wl):
    px = wl["H"] * wl["W"] / 65536.0
    if wl.get("tr"):         # per generator iteration: 5 generator forwards + 4 backward (dgrad + wgrad) = 13 passes of 50.55 GMAC,
        # 4 discriminator forwards + 4 data gradients (3.1 GMAC each); discriminator update: 8 forwards + 8 dgrad + 8 wgrad
        return wl["B"] * 2e9 * (3 * (13 * 50.55 + 8 * 3.10) + 24 * 3.10) * px
    if wl.get("gan"):        # generator fwd + dgrad + wgrad (50.55 GMAC each), discriminator: 3 forwards, 3 data-gradient and 2 weight-
        return wl["B"] * 2e9 * (3 * 50.55 + 8 * 3.10) * px          # gradient passes of 3.10 GMAC (k4 convs 1-64-128-256-512-1)
    if wl.get("i2d"):        # Image_f forward (21.92 GMAC) + Task U-Net 128->1 forward + dgrad + wgrad (8.05 GMAC each) per image, 2 images per pair
        return wl["B"] * 2 * 2e9 * (21.92 + 3 * 8.05) * px
    if wl.get("sr"):
        return wl["B"] * (FLOP_SR_HR_256 * 4 * px + FLOP_SR_LR_256 * px)
    return wl["B"] * FLOP_PER_PAIR_256 * px


def run_config(wl, n_gpus):
    """the `config` object of the JSON line - identical for both arms (`--impl ours` / `--impl reference`)"""
    return dict(workload=wl["name"], crop=[wl["H"], wl["W"]], batch_per_gpu=wl["B"], parallelism=f"dp{n_gpus}",
                l2_policy="inputs+activations per step (>1 GB) exceed the 126 MB L2; no explicit flush",
                algorithmic_tflop_per_step=step_flops(wl) / 1e12)


def make_model(wl, gpu_ids, graph, name="bench"):
    from dsr_b200 import I2D_model, main_model, main_sr_model, options
    if wl.get("tr"):
        from dsr_b200 import translation_model
        return translation_model.TranslationModel(options.translation_flags(gpu_ids=gpu_ids, batch_size=wl["B"], crop_size_h=wl["H"],
                                                                            crop_size_w=wl["W"], name=name, checkpoints_dir="/tmp/dsr_bench",
                                                                            cuda_graph=bool(graph)))
    if wl.get("gan"):
        from dsr_b200 import translation_blocks
        return translation_blocks.GanBlockStep(options.default_opt(gpu_ids=gpu_ids, batch_size=wl["B"], crop_size_h=wl["H"],
                                                                   crop_size_w=wl["W"], name=name, checkpoints_dir="/tmp/dsr_bench",
                                                                   lr=0.0002, netD="n_layers", n_layers_D=3, norm_d="none", ndf=64,
                                                                   cuda_graph=bool(graph)))
    if wl.get("i2d"):
        return I2D_model.I2DModel(options.i2d_flags(gpu_ids=gpu_ids, batch_size=wl["B"], crop_size_h=wl["H"], crop_size_w=wl["W"],
                                                    name=name, checkpoints_dir="/tmp/dsr_bench", cuda_graph=bool(graph)))
    kw = dict(gpu_ids=gpu_ids, batch_size=wl["B"], crop_size_h=wl["H"], crop_size_w=wl["W"], name=name,
              checkpoints_dir="/tmp/dsr_bench", cuda_graph=bool(graph))
    if wl.get("sr"):       # README.md:86
        kw.update(w_real_l1_d=90.0, w_syn_norm=3.0, w_syn_holes=1600.0, w_real_holes=1600.0, lr=0.00002, SR=True)
        return main_sr_model.MainSRModel(options.main_flags(**kw))
    return main_model.MainModel(options.main_flags(**kw))


def make_batch(wl, seed):
    from dsr_b200.synthetic import synthetic_batch, synthetic_sr_batch
    if wl.get("tr"):
        b = synthetic_batch(wl["B"], wl["H"], wl["W"], seed=seed, depth_kind="smooth")
        return dict(A_name=b["A_paths"], B_name=b["B_paths"], A_img=b["A_i"], A_depth=b["A_d"], B_img=b["B_i"], B_depth=b["B_d"])
    if wl.get("sr"):
        return synthetic_sr_batch(wl["B"], wl["H"], wl["W"], seed=seed, depth_kind="smooth")
    return synthetic_batch(wl["B"], wl["H"], wl["W"], seed=seed, depth_kind="smooth")


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0), "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        import statistics
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.samples))


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the newest committed `ncu --set full` summary under profiles/
    (scripts/ncu_summary.py traffic ...); None when no capture is committed."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*traffic*.json")), reverse=True):
        try:
            d = json.load(open(path)).get(kernel)
        except Exception:
            d = None
        if d:
            return dict(bytes_per_launch=d["dram_bytes_per_launch"], captured_launches=d["captured_launches"],
                        source=os.path.relpath(path, ROOT))
    return None


def cpu_gan_baseline(sds, batch, steps, warmup):
    """oracle/ref_nets.gan_block_step (generator + discriminator forward / backward, torch CPU fp32) on the host cores"""
    import torch
    from oracle import ref_nets
    sd_g = {k: v.detach().clone().float().requires_grad_(True) for k, v in sds["G_A"].items()}
    sd_d = {k: v.detach().clone().float().requires_grad_(True) for k, v in sds["D_A_depth"].items()}
    ts = []
    for _ in range(warmup + steps):
        t0 = time.perf_counter()
        ref_nets.gan_block_step(sd_g, sd_d, batch["A_d"], batch["A_i"], batch["B_d"])
        ts.append(time.perf_counter() - t0)
    ts = ts[warmup:]
    B = batch["A_d"].shape[0]
    return dict(s_per_step=sum(ts) / len(ts), value=B * len(ts) / sum(ts), cores=torch.get_num_threads(), host_cpus=os.cpu_count(),
                kind="port", what="oracle/ref_nets.gan_block_step, torch CPU fp32")


def cpu_baseline(B, H, W, steps=2, warmup=1, sds=None, sr=False, i2d=False, gan=False, tr=False):
    """The reference's CPU path (--gpu_ids -1) on this box's host cores, all of them: the UNMODIFIED reference staged in
    oracle/_ref (oracle/build_ref.py; kind "reference") for the main / SR steps when present, the oracle port otherwise."""
    import contextlib
    import numpy as np
    import torch
    from oracle import ref_live, ref_step
    # torch.distributed.run exports OMP_NUM_THREADS=1: give the CPU arm every host core whatever launched us
    torch.set_num_threads(os.cpu_count() or 1)
    wl = dict(B=B, H=H, W=W, sr=sr, i2d=i2d, gan=gan, tr=tr)
    if ref_live.available() and not (i2d or gan or tr):
        with contextlib.redirect_stdout(sys.stderr):          # the reference prints its option table and network summary
            s_per_step, _ = ref_live.time_steps(B, H, W, make_batch(wl, 1), steps, warmup, sr=sr)
        return dict(s_per_step=s_per_step, value=B / s_per_step, cores=torch.get_num_threads(), host_cpus=os.cpu_count(),
                    kind="reference", what="unmodified reference (oracle/_ref, models/main%s_model.py) on --gpu_ids -1" % ("_sr" if sr else ""))
    if sds is None:
        torch.manual_seed(0)
        host = make_model(wl, [], False, name="cpu")
        sds = {n: getattr(host, "net" + n).state_dict() for n in host.model_names}
    if tr:
        from oracle import ref_translation
        orc = ref_translation.OracleTranslationStep(sds)
        batch = make_batch(wl, 1)
        ts = []
        for _ in range(warmup + steps):
            t0 = time.perf_counter()
            orc.step(batch)
            ts.append(time.perf_counter() - t0)
        ts = ts[warmup:]
        return dict(s_per_step=sum(ts) / len(ts), value=B * len(ts) / sum(ts), cores=torch.get_num_threads(), host_cpus=os.cpu_count(),
                    kind="port", what="oracle/ref_translation.py, torch CPU fp32")
    if gan:
        return cpu_gan_baseline(sds, make_batch(wl, 1), steps, warmup)
    if i2d:
        orc = ref_step.OracleI2DStep(sds, lr=2e-4)
    else:
        orc = ref_step.OracleSRStep(sds, (H, W), lr=2e-5) if sr else ref_step.OracleStep(sds, lr=1e-4)
    batch = make_batch(wl, 1)
    np.random.seed(0)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc.step(batch)
        ts.append(time.perf_counter() - t0)
    ts = ts[warmup:]
    return dict(s_per_step=sum(ts) / len(ts), value=B * len(ts) / sum(ts), cores=torch.get_num_threads(),
                host_cpus=os.cpu_count(), kind="port", what="oracle/ref_step.py, torch CPU fp32")


def torch_gpu_bar(wl, steps=5, warmup=3, device="cuda:0"):
    """--torch-gpu-bar (SURVEY.md section 8d "GPU reference bar"): the five nets of the step through torch / cuDNN eager on the
    SAME GPU - frozen G_A_d, I2D_features, Image2Depth forward, Depth_f and Task forward + backward (oracle/ref_nets.py, the
    functional restatement of the reference's nn.Modules over reference-layout state_dicts; none of this repo's kernels).
    No rectangle holes, no loss stack (an L1 on the two predictions stands in), no optimizer: a LOWER bound of the reference's
    GPU step.  A reported baseline like `cpu_baseline`, not the thing measured or shipped."""
    import torch
    from oracle import ref_nets
    torch.manual_seed(0)
    host = make_model(wl, [], False, name="cpu")
    sds = {n: {k: v.detach().float().to(device) for k, v in getattr(host, "net" + n).state_dict().items()} for n in host.model_names}
    for n in ("Depth_f", "Task"):
        for v in sds[n].values():
            v.requires_grad_(True)
    b = make_batch(wl, 1)
    si, ri, sd_, rd = (b[k].float().to(device) for k in ("A_i", "B_i", "A_d", "B_d"))
    out = {}
    for tag, tf32 in (("fp32", False), ("tf32_conv", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True

        def step():
            for n in ("Depth_f", "Task"):
                for v in sds[n].values():
                    v.grad = None
            with torch.no_grad():
                s2r = ref_nets.translation_generator(sds["G_A_d"], sd_, si)
                f_s, f_r = ref_nets.resnet_generator(sds["I2D_features"], si), ref_nets.resnet_generator(sds["I2D_features"], ri)
                dbi_s, dbi_r = ref_nets.unet_generator(sds["Image2Depth"], f_s), ref_nets.unet_generator(sds["Image2Depth"], f_r)
            in_s, in_r = torch.cat([s2r, dbi_s], 1), torch.cat([rd, dbi_r], 1)
            fd_s, fd_r = ref_nets.resnet_generator(sds["Depth_f"], in_s), ref_nets.resnet_generator(sds["Depth_f"], in_r)
            p_s = ref_nets.unet_generator(sds["Task"], torch.cat([f_s, fd_s, in_s, si], 1))
            p_r = ref_nets.unet_generator(sds["Task"], torch.cat([f_r, fd_r, in_r, ri], 1))
            ((p_s - sd_).abs().mean() + (p_r - rd).abs().mean()).backward()

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        out[tag] = dict(ms_per_step=e0.elapsed_time(e1) / steps, pair_samples_per_s=wl["B"] * steps / (e0.elapsed_time(e1) * 1e-3))
    out["note"] = ("torch %s / cuDNN eager on the same GPU, batch %d at %dx%d: five nets forward, Depth_f + Task backward, L1 stand-in loss; "
                   "no rectangle holes / loss stack / optimizer (lower bound of the reference's GPU step); fp32 = "
                   "cudnn.allow_tf32 False, tf32_conv = torch's default" % (torch.__version__, wl["B"], wl["H"], wl["W"]))
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the workload's own batch, all host threads
    (the unmodified reference from oracle/_ref when staged, else the oracle port)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    B = wl["B"]
    r = cpu_baseline(B, wl["H"], wl["W"], steps=args.steps, warmup=args.warmup, sr=bool(wl.get("sr")), i2d=bool(wl.get("i2d")), gan=bool(wl.get("gan")), tr=bool(wl.get("tr")))
    line = dict(impl="reference", metric="RGB-D train pair-samples/sec (main net)", value=r["value"], unit="pair-samples/s",
                n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * r["s_per_step"],
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=run_config(wl, args.gpus),
                cpu_baseline=dict(value=r["value"], unit="pair-samples/s", cores=r["cores"], kind=r["kind"],
                                  sample=f"{args.steps} steps of batch {B} at {wl['H']}x{wl['W']} after {args.warmup} warm-up ({r['what']}, {r['host_cpus']} host CPUs)"),
                e2e=dict(value=r["value"], unit="pair-samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def inference_ms_per_frame(local, iters=20):
    """BASELINE.json's second number: ms per 640x480 frame (fed as 512x640, batch 1) of the enhancement forward
    (`calculate('test')` minus PNG writing, main_model.py:433-436 -> forward only here), device-timed, host inputs."""
    import numpy as np
    import torch
    wl = dict(B=1, H=512, W=640)
    torch.manual_seed(0)
    m = make_model(wl, [local], False, name="infer")
    m.eval()
    b = make_batch(wl, 5)
    for k in ("A_i", "B_i", "A_d", "B_d"):
        b[k] = b[k].pin_memory()
    np.random.seed(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for i in range(3 + iters):
            if i == 3:
                torch.cuda.synchronize()
                e0.record()
            m.set_input(b)
            m.forward("test")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        # the same pass replayed as a CUDA graph (MainModel.forward_test_graph)
        for i in range(4 + iters):
            if i == 4:
                torch.cuda.synchronize()
                e0.record()
            m.set_input(b)
            m.forward_test_graph()
        e1.record()
        torch.cuda.synchronize()
        ms_g = e0.elapsed_time(e1) / iters
    return dict(ms_per_frame=ms_g, frames_per_s=1e3 / ms_g, ms_per_frame_eager=ms, shape=[512, 640], batch=1,
                note="set_input (H2D from pinned host) + forward('test'): G_A_d + I2D_features + Image2Depth + Depth_f + Task on "
                     "the [syn; real] pair; ms_per_frame = CUDA-graph replay (forward_test_graph), ms_per_frame_eager = eager launches")


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from dsr_b200 import _lib, ops, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    _lib.load()
    ops.CONFIG.update(engine=args.engine, passes=args.passes, dtype=args.dtype)
    for kv in filter(None, args.cfg.split(",")):            # ablations: --cfg side_wgrad=0,fork_frozen=0
        k, v = kv.split("=")
        ops.CONFIG[k] = type(ops.CONFIG[k])(int(v)) if isinstance(ops.CONFIG[k], (bool, int)) else v
    wl = WORKLOADS[args.workload]
    B, H, W = wl["B"], wl["H"], wl["W"]
    torch.manual_seed(0)
    model = make_model(wl, [local], args.graph)
    model._train()
    sync, dp_check = None, None
    if world > 1:
        parallel.broadcast_weights(model)
        sync = parallel.GradBuckets(model, overlap=bool(args.dp_overlap))
        if not any(wl.get(k) for k in ("i2d", "gan", "tr", "sr")):
            # numerical check of the data-parallel step before anything is timed or captured: rank-averaged gradients of one
            # sharded step == gradients of ONE process on the concatenated batch (every rank generates every shard)
            from dsr_b200 import main_model
            Bc = min(B, 2)
            cb, cr = [], []
            for r in range(world):
                full = make_batch(dict(wl, B=Bc), 101 + r)
                cb.append(full)
                rng = np.random.RandomState(500 + r)
                rr, rc = main_model.draw_rects(Bc, H, W, "train", rng=rng)
                sr_, sc = main_model.draw_rects(Bc, H, W, "train", rng=rng)
                cr.append((rr, rc, sr_, sc))
            dp_check = parallel.dp_self_check(model, sync, cb, cr)
            dp_check["sample"] = f"one step, {Bc} pairs per rank at {H}x{W}; cos / rel_l2 = worst over ranks"
    # a few distinct host batches in pinned memory (per-rank seeds: each rank draws its own shard)
    host_batches = []
    for i in range(2):
        b = make_batch(wl, 1 + 17 * rank + i)
        for k in b:
            if torch.is_tensor(b[k]) and b[k].dtype == torch.float32:
                b[k] = b[k].pin_memory()
        host_batches.append(b)
    dev_batches = [{k: (v.cuda() if torch.is_tensor(v) and v.dtype == torch.float32 else v) for k, v in b.items()}
                   for b in host_batches]
    np.random.seed(1234 + rank)
    in_keys = ("A_img", "A_depth", "B_img", "B_depth") if wl.get("tr") else (("A_i", "A_d", "B_d") if wl.get("gan") else ("A_i", "B_i", "A_d", "B_d"))
    h2d = sum(host_batches[0][k].numel() * 4 for k in in_keys)
    if not any(wl.get(k) for k in ("i2d", "gan", "tr")):
        h2d += 2 * B * 11 * 8 + 2 * B * (64 * 4 + 1) * 4          # camera tables + rectangle tables

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batches, steps, read_loss):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        pending = None
        for i in range(steps):
            model.set_input(batches[i % len(batches)])
            model.optimize_parameters(i, 1)
            if read_loss:                             # D2H read of EVERY step's result: the copy of step i is enqueued behind
                nxt = model.loss_async("loss_G") if hasattr(model, "loss_async") else (lambda m=model: float(m.loss_G))
                if pending is not None:               # step i, the host reads it (and checks it) while step i+1 runs
                    v = pending()
                    if v != v or v in (float("inf"), float("-inf")):
                        raise FloatingPointError(f"loss_G is not finite at step {i - 1}")
                pending = nxt
        if pending is not None:
            pending()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    timed(dev_batches, args.warmup + (model.graph_warmup + 1 if model.use_graph else 0), False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.LAUNCHES
    ms = timed(dev_batches, args.steps, False)                 # inputs resident in HBM
    launches = (_lib.LAUNCHES - l0) // max(args.steps, 1)
    timed(host_batches, 2, True)                               # untimed: the staging ring / pinned loss slots are created here
    ms_e2e = timed(host_batches, args.steps, True)             # pinned host inputs, H2D + loss D2H inside
    sampler.stop_flag = True

    # per-kernel pass: CUDA events around every library call of ONE more step (same stream)
    if args.kernel_table and rank == 0 and world == 1:
        # in-situ kernel durations (CUPTI through torch.profiler): warm caches and real overlap, unlike the serialised
        # cold-cache ncu pass; used for the time breakdown only, never for the headline numbers
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as tp:
            for i in range(3):
                model.set_input(dev_batches[i % len(dev_batches)])
                model.optimize_parameters(i, 1)
            torch.cuda.synchronize()
        agg = {}
        for ev in tp.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                d = agg.setdefault(ev.name[:90], [0, 0.0])
                d[0] += 1; d[1] += ev.device_time
        evs = sorted((ev for ev in tp.events() if ev.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
        third = evs[2 * len(evs) // 3:]                                 # the launches of (about) the last step, in order
        json.dump(dict(per_kernel=sorted(([k, v[0] / 3.0, v[1] / 3.0] for k, v in agg.items()), key=lambda r: -r[2]),
                       last_step=[[ev.name[:60], ev.device_time] for ev in third],
                       timeline=[[ev.name[:40], ev.time_range.start - third[0].time_range.start, ev.device_time,
                                  getattr(ev, "device_resource_id", -1)] for ev in third],
                       # all three steps (pipelined replay: the frozen networks of step i+1 run beside the training part of step i)
                       timeline_all=[[ev.name[:40], ev.time_range.start - evs[0].time_range.start, ev.device_time,
                                      getattr(ev, "device_resource_id", -1)] for ev in evs]), open(args.kernel_table, "w"))
    prof = None
    # every rank runs this extra eager step (its gradient all-reduce is a collective); rank 0 brackets each call
    if rank == 0:
        _lib.PROFILE = []
    model.set_input(dev_batches[0])
    # The eager step is host-bound (~35 us of Python per call): with an idle GPU every bracket [event, kernel, event] would
    # also contain the host's gap between recording the first event and launching the kernel.  A ~100 ms device-side spin
    # ahead of the step lets the host enqueue the whole step first, so the brackets run back to back on the device and
    # measure kernel durations only.
    torch.cuda._sleep(int(0.10 * 1.9e9))
    was = {k: ops.CONFIG.get(k) for k in ("fork_frozen", "side_wgrad")}
    ops.CONFIG.update(fork_frozen=False, side_wgrad=False)    # one stream: a bracket must contain ONE kernel, not overlapping chains
    model._step_body()                        # eager launches, so every library call can be bracketed by events
    ops.CONFIG.update(was)
    torch.cuda.synchronize()
    if rank == 0:
        rec, _lib.PROFILE = _lib.PROFILE, None
        if args.layer_table:
            rows = [dict(call=name, ms=round(a.elapsed_time(b), 4), **({"shape": list(meta["shape"]), "gmacs": round(meta["macs"] / 1e9, 3)} if meta else {}))
                    for name, a, b, meta in rec]
            json.dump(rows, open(args.layer_table, "w"))
        by = {}
        for name, a, b, meta in rec:
            d = by.setdefault(name, dict(ms=0.0, n=0, macs=0, xmacs=0))
            d["ms"] += a.elapsed_time(b); d["n"] += 1
            if meta:
                d["macs"] += meta["macs"]
                d["xmacs"] += meta["macs"] * meta.get("passes", 0)          # MMAs issued: passes differ per launch
        prof = by
    line = None
    if rank == 0:
        pk, pk_src = peaks()
        total_ms = sum(d["ms"] for d in prof.values())
        top = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]
        mult = {1: 1, 2: 2, 3: 3}[args.passes]
        wmult = min(mult, int(ops.CONFIG["wgrad_passes"]))          # the weight-gradient GEMM issues fewer passes
        passes_of = dict(dsr_tc_gemm=mult, dsr_tc_gemm2=mult, dsr_tc_gemm3=mult, dsr_tc_wgrad=wmult)
        gemm_calls = {k: d for k, d in prof.items() if k in ("dsr_tc_gemm", "dsr_tc_gemm2", "dsr_tc_gemm3", "dsr_tc_wgrad") and d["n"]}
        kernel_of = dict(dsr_tc_gemm="conv_tc_kernel", dsr_tc_gemm2="conv_tc2_kernel", dsr_tc_gemm3="conv_tc3_kernel",
                         dsr_tc_wgrad="wgrad_tc_kernel")
        # MMAs issued per kernel family: recorded per launch (forward 3 passes, data gradients 1 or 3, weight gradients 1);
        # launches without a record fall back to the configured pass count
        xm = lambda k: gemm_calls[k].get("xmacs") or gemm_calls[k]["macs"] * passes_of[k]
        if gemm_calls:
            dom = max(gemm_calls, key=lambda k: gemm_calls[k]["ms"])           # the dominant kernel of the step
            tc = gemm_calls[dom]
            achieved = 2.0 * tc["macs"] / (tc["ms"] * 1e-3) / 1e12
            fam_ms = sum(d["ms"] for d in gemm_calls.values())
            fam_macs = sum(d["macs"] for d in gemm_calls.values())
            tr = ncu_traffic(kernel_of[dom])
            roof = dict(bound="tensor", kernel=f"{kernel_of[dom]} ({dom})", achieved=achieved, peak=pk["bf16_tflops_sustained"],
                        unit="TFLOP/s", frac=achieved / pk["bf16_tflops_sustained"],
                        traffic=tr["bytes_per_launch"] if tr else None, traffic_source=tr,
                        peak_source=pk_src + " (sustained: timed inside a long step)",
                        launches_per_step=tc["n"], avg_launch_us=1e3 * tc["ms"] / tc["n"], share_of_step=tc["ms"] / total_ms,
                        mma_passes=round(xm(dom) / max(tc["macs"], 1), 3), executed_tflops=2.0 * xm(dom) / (tc["ms"] * 1e-3) / 1e12,
                        executed_frac=2.0 * xm(dom) / (tc["ms"] * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
                        all_tcgen05_gemms=dict(share_of_step=fam_ms / total_ms, achieved=2.0 * fam_macs / (fam_ms * 1e-3) / 1e12,
                                               executed_tflops=2.0 * sum(xm(k) for k in gemm_calls) / (fam_ms * 1e-3) / 1e12,
                                               mma_passes={kernel_of[k]: round(xm(k) / max(d["macs"], 1), 3) for k, d in gemm_calls.items()},
                                               ms={kernel_of[k]: round(d["ms"], 3) for k, d in gemm_calls.items()}),
                        note="achieved = algorithmic conv FLOPs of the layers this kernel served / their summed launch time (CUDA "
                             "events on the launching stream around every library call of one eager step enqueued behind a device-side spin, so "
                             "the brackets contain no host gaps); each product is issued as "
                             "`mma_passes` 16-bit MMAs on average (3 for forward GEMMs - the hi/lo operand split the parity gates need - "
                             "1 for weight gradients and for data gradients of layers >= 32^2), so the tensor pipe executes "
                             "`executed_tflops`; traffic = dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the launches "
                             "captured by ncu --set full (profiles/, a previous run of the same command on the same workload)")
        else:
            k, d = top[0]
            roof = dict(bound="hbm", kernel=k, achieved=None, peak=pk["hbm_gbs"], unit="GB/s", frac=None, traffic=None,
                        share_of_step=d["ms"] / total_ms)
        flop_step = step_flops(wl)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            r = cpu_baseline(B, H, W, steps=2, warmup=1, sr=bool(wl.get("sr")), i2d=bool(wl.get("i2d")), gan=bool(wl.get("gan")), tr=bool(wl.get("tr")))
            cpu = dict(value=r["value"], unit="pair-samples/s", cores=r["cores"], kind=r["kind"],
                       sample=f"2 steps of batch {B} at {H}x{W} after 1 warm-up ({r['what']}, {r['host_cpus']} host CPUs)")
        hbm = None
        if args.stencils and world == 1:
            from dsr_b200 import stencil_bench
            tab = stencil_bench.run(batch=96, iters=10, peak=pk["hbm_gbs"], device=local)
            rows = [r for r in tab["rows"] if not r["kernel"].startswith("ssim")]
            worst = min(rows, key=lambda r: r["frac"])
            ssim_row = [r for r in tab["rows"] if r["kernel"].startswith("ssim")]
            hbm = dict(bound="hbm", peak=pk["hbm_gbs"], unit="GB/s", peak_source=pk_src,
                       worst=dict(kernel=worst["kernel"], achieved=worst["gbs"], frac=worst["frac"]),
                       median_frac=sorted(r["frac"] for r in rows)[len(rows) // 2],
                       at_or_above_70pct=sum(r["frac"] >= 0.70 for r in rows), kernels=len(rows),
                       table={r["kernel"]: [r["gbs"], r["frac"]] for r in rows},
                       ssim=(dict(achieved=ssim_row[0]["gbs"], frac_of_hbm=ssim_row[0]["frac"],
                                  note="fp32-pipe bound (five 11x11 separable Gaussian windows, 127 FMA/pixel): its ceiling is ~35% of the HBM rate")
                             if ssim_row else None),
                       shape=tab["shape"], note="achieved = algorithmic bytes (SURVEY.md 8d) / mean launch time, CUDA events on the launching "
                       "stream, 10 back-to-back launches after 3 warm-ups, in this process; 96 planes of 512x640 (126 MB per fp32 plane set, "
                       "beyond the L2) - the C3 frame size, batched so that nothing is served from cache")
        gpu_bar = None
        if args.torch_gpu_bar and world == 1 and not any(wl.get(k) for k in ("sr", "i2d", "gan", "tr")):
            try:
                gpu_bar = torch_gpu_bar(wl, device=f"cuda:{local}")
            except Exception as e:                      # a baseline that cannot run must not take the measurement down
                gpu_bar = dict(error=f"{type(e).__name__}: {e}"[:300])
        extras = None
        if args.inference and not any(wl.get(k) for k in ("sr", "i2d", "gan", "tr")) and world == 1:
            extras = inference_ms_per_frame(local)
        line = dict(metric="RGB-D train pair-samples/sec (main net)", value=world * B * args.steps / (ms * 1e-3),
                    unit="pair-samples/s", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps,
                    higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype=(f"{args.dtype} x{args.passes} operands (weight gradients: {ops.CONFIG['bwd_dtype']} x{wmult}"
                           + (f"; data gradients of layers >= {ops.CONFIG['big_hw']} px: {ops.CONFIG['bwd_dtype']} x{ops.CONFIG['big_bwd_passes']}"
                              if ops.CONFIG["big_hw"] and ops.CONFIG["big_bwd_passes"] else "") + "), f32 accumulate"
                           if args.engine == "tc" else "f32"),
                    data="synthetic",
                    config=run_config(wl, world), engine=args.engine, cuda_graph=bool(model.use_graph),
                    e2e=dict(value=world * B * args.steps / (ms_e2e * 1e-3), unit="pair-samples/s", h2d_bytes_per_step=h2d,
                             d2h_bytes_per_step=4, ms_per_step=ms_e2e / args.steps),
                    gpu_launches=launches, clocks=sampler.summary(), roofline=roof, roofline_hbm=hbm, cpu_baseline=cpu, dp_check=dp_check,
                    step_tflops=flop_step / (ms / args.steps * 1e-3) / 1e12, inference_640x480=extras,
                    kernel_times_ms={k: round(d["ms"], 3) for k, d in top})
        if gpu_bar is not None:
            line["torch_eager_gpu"] = gpu_bar
        print(json.dumps(line), flush=True)
    if world > 1:
        # orderly exit: the captured graph holds the communicator's collectives, so it is released first, then the device
        # is drained and the process group destroyed (parallel.shutdown)
        sys.stdout.flush()
        parallel.shutdown([model])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=list(WORKLOADS))
    ap.add_argument("--engine", default="tc", choices=["tc", "simt"])
    ap.add_argument("--passes", type=int, default=3, choices=[1, 2, 3])
    ap.add_argument("--dtype", default="f16", choices=["f16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-gpu-bar", action="store_true",
                    help="also time the five nets through torch / cuDNN eager on the same GPU (SURVEY.md 8d 'GPU reference bar'; off by default)")
    ap.add_argument("--layer-table", default="", help="write every library call of one step (name, ms, shape, GMACs) to this JSON file")
    ap.add_argument("--kernel-table", default="", help="write the in-situ per-kernel device times of 3 steps (torch.profiler / CUPTI, warm caches) to this JSON file")
    ap.add_argument("--inference", type=int, default=1, help="1 = also time the 640x480 inference forward (ms/frame)")
    ap.add_argument("--cfg", default="", help="comma-separated ops.CONFIG overrides (ablations), e.g. side_wgrad=0,fork_frozen=0")
    ap.add_argument("--dp-overlap", type=int, default=1, help="0 = one in-order all-reduce of the gradient arena instead of overlapped buckets")
    ap.add_argument("--stencils", type=int, default=1, help="1 = also measure the HBM roofline table of the stencil / reduction kernels (roofline_hbm)")
    ap.add_argument("--graph", type=int, default=1, help="1 = replay the training step as a CUDA graph (default), 0 = eager launches")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
