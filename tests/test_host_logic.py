"""CPU-only checks: the C-ABI library loads and exports every symbol include/dsr_b200.h declares
(no compute), the host RNG reproduces the reference's rectangle stream, and the product refuses to
compute without CUDA."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import ref_ops


def test_library_exports_every_declared_symbol(built_lib):
    from dsr_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 35
    lib = ctypes.CDLL(built_lib)
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in include/dsr_b200.h but not exported"
    assert _lib.load().dsr_version() == 100


def test_header_cites_reference_call_sites():
    from dsr_b200 import _lib
    src = open(_lib.HEADER_PATH).read()
    for needle in ("models/main_model.py:208-230", "models/norms.py:103-108", "models/networks.py", "models/main_model.py:176"):
        assert needle in src


@pytest.mark.parametrize("stage", ["train", "test"])
def test_rect_rng_stream_matches_oracle(stage):
    from dsr_b200 import main_model
    for H, W in ((128, 128), (256, 256), (512, 640)):
        np.random.seed(5)
        a = ref_ops.draw_rects(4, H, W, stage)
        a2 = ref_ops.draw_rects(4, H, W, stage)
        np.random.seed(5)
        r, c = main_model.draw_rects(4, H, W, stage)
        r2, c2 = main_model.draw_rects(4, H, W, stage)
        for i in range(4):
            assert np.array_equal(r[i, :c[i]], a[i]) and np.array_equal(r2[i, :c2[i]], a2[i])
        assert c.max() < main_model.MAX_RECTS


def test_no_cpu_fallback(built_lib):
    from dsr_b200 import networks, ops
    net = networks.define_G(2, 128, 32, "resnet_6blocks", "instance", False, "normal", 0.02, [])
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 2, 16, 16))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.normals_old(torch.zeros(1, 1, 8, 8))


def test_unknown_names_raise_like_reference():
    from dsr_b200 import networks
    with pytest.raises(NotImplementedError):
        networks.define_G(3, 1, 64, "no_such_net", "instance")
    with pytest.raises(NotImplementedError):
        networks.get_norm_layer("nope")


def test_scheduler_and_optimizer_contract():
    from util import build_host_model
    m = build_host_model(1, 128, 128)
    m.setup(m.opt)
    assert len(m.optimizers) == 1 and m.optimizers[0].param_groups[0]["lr"] == pytest.approx(1e-4)
    m.update_learning_rate()
    n_train = sum(p.numel() for p in m.netDepth_f.parameters()) + sum(p.numel() for p in m.netTask.parameters())
    assert n_train == 2159616 + 42085761


def test_mma_pass_policy_is_pure_host_logic():
    """ops._passes: forward GEMMs keep the parity mode's three passes; data-gradient GEMMs (a backward operand format is
    passed) of layers with >= big_hw output pixels run big_bwd_passes; the knobs switch the policy off."""
    from dsr_b200 import ops
    old, old_fwd = dict(ops.CONFIG), dict(ops._FWD)
    try:
        assert ops.CONFIG["passes"] == 3 and ops.CONFIG["big_hw"] == 1024 and ops.CONFIG["big_bwd_passes"] == 1
        ops._FWD.update(hw=64 * 64, trainable=True)
        assert ops._passes(None) == 3 and ops._passes("bf16") == 1
        ops._FWD.update(hw=16 * 16)
        assert ops._passes(None) == 3 and ops._passes("bf16") == 3
        ops.CONFIG.update(big_hw=0)
        ops._FWD.update(hw=256 * 256)
        assert ops._passes("bf16") == 3
        ops.CONFIG.update(big_hw=1024, big_fwd_passes=2, passes=1)
        assert ops._passes(None) == 1 and ops._passes("bf16") == 1          # never more than the global pass count
    finally:
        ops.CONFIG.clear(); ops.CONFIG.update(old)
        ops._FWD.clear(); ops._FWD.update(old_fwd)


def test_ring_planner_is_callable_without_a_gpu(built_lib):
    """dsr_smooth_ring_suits is host arithmetic: big plane sets suit the row ring, training-crop sizes and ragged widths do not"""
    from dsr_b200 import _lib
    lib = _lib.load()
    assert lib.dsr_smooth_ring_suits(96, 3, 512, 640) == 1
    assert lib.dsr_smooth_ring_suits(6, 3, 256, 256) == 0           # 393 K pixels: the register kernels
    assert lib.dsr_smooth_ring_suits(96, 3, 512, 642) == 0          # rows must be 16-byte multiples
    assert lib.dsr_smooth_ring_suits(64, 3, 1024, 8192) == 0        # one row of all planes must fit a 40 KB group


def test_package_synthetic_batches_equal_the_oracle_generator():
    """bench.py's product leg draws its inputs from dsr_b200.synthetic (it must not import oracle/): same tensors"""
    from dsr_b200 import synthetic
    from oracle import ref_step
    for kind in ("smooth", "noise"):
        a, b = synthetic.synthetic_batch(2, 32, 48, seed=4, depth_kind=kind), ref_step.synthetic_batch(2, 32, 48, seed=4, depth_kind=kind)
        assert a.keys() == b.keys()
        assert all(torch.equal(a[k], b[k]) if torch.is_tensor(a[k]) else a[k] == b[k] for k in a)
    a, b = synthetic.synthetic_sr_batch(1, 16, 24, seed=2), ref_step.synthetic_sr_batch(1, 16, 24, seed=2)
    assert all(torch.equal(a[k], b[k]) if torch.is_tensor(a[k]) else a[k] == b[k] for k in a)


def test_bench_product_leg_does_not_import_the_oracle():
    import ast
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    allowed = {"cpu_baseline", "cpu_gan_baseline", "torch_gpu_bar", "run_reference"}        # the CPU / reference legs
    for fn in tree.body:
        if isinstance(fn, ast.FunctionDef) and fn.name not in allowed:
            for node in ast.walk(fn):
                if isinstance(node, ast.ImportFrom) and (node.module or "").startswith("oracle"):
                    raise AssertionError(f"bench.py:{fn.name} imports {node.module}")
    pkg = os.path.join(root, "depth-enhancement-and-super-resolution_b200", "dsr_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py") and f != "selfcheck.py":
            src = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


def test_zero_pool_clears_its_high_water_mark():
    """a reset captured into a CUDA graph has a fixed length: it must cover everything the pool ever handed out"""
    from dsr_b200 import ops
    pool = ops._ZeroPool()
    dev = torch.device("cpu")
    pool.reset(dev)
    a = pool.take(100, dev); a += 1
    b = pool.take(50, dev); b += 1
    pool.reset(dev)                  # zeroes [:152]
    c = pool.take(10, dev); c += 1   # a shorter pass ...
    pool.reset(dev)                  # ... must still clear the whole region handed out before
    key = pool._key(dev)
    assert pool.high[key] >= 152
    assert float(pool.buf[key][:200].abs().sum()) == 0.0
    # named pools (two captured graphs that may run at the same time must not share accumulators)
    pool.name = "frozen"
    pool.reset(dev)
    d = pool.take(8, dev); d += 1
    assert pool._key(dev) != key and float(pool.buf[key][:200].abs().sum()) == 0.0
    pool.name = "step"
    assert float(pool.take(8, dev).abs().sum()) == 0.0
