"""Depth metrics evaluator (dsr_b200.metrics, csrc/metrics.cu; new_metrics.py of the reference): the numpy oracle against
golden values of the LIVE reference functions (CPU), and the CUDA batch evaluator against both (GPU)."""
import numpy as np
import pytest

from oracle import ref_metrics
from util import load_golden

TOL = 1e-9          # float64 sums in both implementations; the orders of summation differ


def _cases():
    g = load_golden("metrics.npz")
    names = [str(n) for n in g["names"]]
    cases = []
    i = 0
    while f"c{i}/pred" in g.files:
        cases.append((g[f"c{i}/pred"], g[f"c{i}/target"], g[f"c{i}/input"], dict(zip(names, g[f"c{i}/values"]))))
        i += 1
    return cases, g["K"], names


def test_metrics_oracle_matches_reference():
    cases, K, names = _cases()
    for pred, target, inp, want in cases:
        got = ref_metrics.calc_metrics(pred, target, inp, K, 5100)
        for n in names:
            assert abs(got[n] - want[n]) <= TOL * max(abs(want[n]), 1.0), (n, got[n], want[n])


@pytest.mark.gpu
def test_metrics_gpu_matches_reference_and_oracle(built_lib):
    from dsr_b200 import metrics
    cases, K, names = _cases()
    for pred, target, inp, want in cases:
        got = metrics.calc_metrics(pred[None], target[None], inp[None], K, 5100, names)
        for n in names:
            assert abs(float(got[n][0]) - want[n]) <= 1e-8 * max(abs(want[n]), 1.0), (n, float(got[n][0]), want[n])
    # a batch of 3 with an all-valid image (mae_h undefined -> nan) and a 2x target (SR evaluation, new_metrics.py:217-218)
    pred, target, inp, _ = cases[0]
    full = target.copy(); full[full < 50] = 900.0
    batch_p = np.stack([pred, pred, pred]); batch_t = np.stack([target, full, target]); batch_i = np.stack([inp, full, inp])
    got = metrics.calc_metrics(batch_p, batch_t, batch_i, K, 5100, names)
    assert np.isnan(got["mae_h"][1]) and np.isnan(got["rmse_h"][1]) and not np.isnan(got["mae_h"][0])
    ref1 = ref_metrics.calc_metrics(pred, full, full, K, 5100)
    for n in ("mae", "rmse", "psnr", "ssim", "mae_d", "rmse_d", "mse_v"):
        assert abs(float(got[n][1]) - ref1[n]) <= 1e-8 * max(abs(ref1[n]), 1.0), n
    mean = metrics.mean_over_images(got)
    assert abs(mean["mae_h"] - float(got["mae_h"][0])) <= 1e-12            # nan-aware mean (:246-248)
    big = np.repeat(np.repeat(target, 2, 0), 2, 1)
    got2 = metrics.calc_metrics(pred[None], big[None], inp[None], K, 5100, ["mae", "rmse"])
    assert abs(float(got2["mae"][0]) - float(got["mae"][0])) <= 1e-12
