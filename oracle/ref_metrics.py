"""Oracle (numpy, float64) restatement of the reference's depth metrics, new_metrics.py:17-201.  TEST INFRASTRUCTURE ONLY.
Pinned against outputs of the live reference functions (tests/golden/metrics.npz, tests/golden/make_golden.py metrics)."""
import numpy as np

HOLES_THRESHOLD = 50      # new_metrics.py:14


def _points(depth, K, shift=0.5):
    """depth_to_absolute_coordinates, 'orthogonal' (:49-66)."""
    h, w = depth.shape
    v, u = np.meshgrid(np.arange(h, dtype=np.float64) + shift, np.arange(w, dtype=np.float64) + shift, indexing="ij")
    pts = np.einsum("lk,kij->lij", np.linalg.inv(K), np.stack([u, v, np.ones_like(v)]))
    pts = pts / pts[2:3]
    return pts * depth[None]


def _normals(c):
    """coords_to_normals (:17-47): forward differences, last column / row replicated, normalize(eps=1e-12)."""
    du = np.concatenate([c[:, :, 1:] - c[:, :, :-1], (c[:, :, 1:] - c[:, :, :-1])[:, :, -1:]], axis=2)
    dv = np.concatenate([c[:, 1:, :] - c[:, :-1, :], (c[:, 1:, :] - c[:, :-1, :])[:, -1:, :]], axis=1)
    n = np.stack([dv[1] * du[2] - du[1] * dv[2], dv[2] * du[0] - du[2] * dv[0], dv[0] * du[1] - du[0] * dv[1]])
    return n / np.maximum(np.sqrt((n * n).sum(0, keepdims=True)), 1e-12)


def _ssim_valid(a, b):
    """_ssim (:81-113): 11x11 Gaussian sigma 1.5, 'valid' mode, L = 1."""
    x, y = np.mgrid[-5:6, -5:6]
    g = np.exp(-((x ** 2 + y ** 2) / (2.0 * 1.5 ** 2)))
    g = g / g.sum()
    H, W = a.shape

    def blur(img):
        out = np.zeros((H - 10, W - 10))
        for r in range(11):
            for c in range(11):
                out += g[r, c] * img[r:r + H - 10, c:c + W - 10]
        return out
    m1, m2 = blur(a), blur(b)
    s1, s2, s12 = blur(a * a) - m1 * m1, blur(b * b) - m2 * m2, blur(a * b) - m1 * m2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return float(np.mean(((2 * m1 * m2 + C1) * (2 * s12 + C2)) / ((m1 * m1 + m2 * m2 + C1) * (s1 + s2 + C2))))


def calc_metrics(pred, target, input_orig, K, max_depth=5100):
    """calc_metrics_for_path without the file reads (:205-232) + calc_metrics (:193-201) for ONE image."""
    pred = np.asarray(pred, dtype=np.float64).clip(0, max_depth)
    target = np.asarray(target, dtype=np.float64).clip(0, max_depth)
    hole, thole = np.asarray(input_orig, dtype=np.float64) < HOLES_THRESHOLD, target < HOLES_THRESHOLD
    out = {}
    d = pred[~thole] - target[~thole]
    out["mae"], out["rmse"] = np.mean(np.abs(d)), np.sqrt(np.mean(d ** 2))
    out["psnr"] = 20.0 * np.log10(1) - 10 * np.log10(np.mean((d / max_depth) ** 2))
    m = ~thole & hole
    out["mae_h"] = np.mean(np.abs(pred[m] - target[m])) if m.any() else np.nan
    out["rmse_h"] = np.sqrt(np.mean((pred[m] - target[m]) ** 2)) if m.any() else np.nan
    u = hole | thole
    out["mae_d"] = np.mean(np.abs(pred[~u] - target[~u])) if not u.all() else np.nan
    out["rmse_d"] = np.sqrt(np.mean((pred[~u] - target[~u]) ** 2)) if not u.all() else np.nan
    out["ssim"] = _ssim_valid(~thole * pred / max_depth, ~thole * target / max_depth)
    if K is not None:
        nt, np_ = _normals(_points(target, K)), _normals(_points(pred, K))
        dm = thole.copy()
        dm[:, 1:] |= thole[:, :-1]; dm[:, :-1] |= thole[:, 1:]; dm[1:, :] |= thole[:-1, :]; dm[:-1, :] |= thole[1:, :]
        keep = np.broadcast_to(~dm, np_.shape)
        out["mse_v"] = np.mean((nt[keep] - np_[keep]) ** 2)
    return out
