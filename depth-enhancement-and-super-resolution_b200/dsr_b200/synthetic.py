"""Synthetic RGB-D batches of SURVEY.md section 8(d) / Appendix D, in the dataset's dict layout
(data/my_main_dataset.py:195, data/my_naive_sr_dataset.py:190-207): what bench.py and the self-checks feed the step.
(The oracle keeps its own statement of the same generator; tests/test_host_logic.py pins the two to each other.)"""
import torch


def synthetic_sr_batch(B, h, w, seed=1, depth_kind="smooth"):
    """HR (2h x 2w) synthetic batch with the K / crop conventions of data/my_naive_sr_dataset.py:190-207:
    K_A scaled by [[2,1,2],[1,2,2],[1,1,1]], crop_A = HR extent, crop_B = LR extent."""
    b = synthetic_batch(B, 2 * h, 2 * w, seed=seed, depth_kind=depth_kind)
    scale = torch.tensor([[2.0, 1, 2], [1, 2, 2], [1, 1, 1]], dtype=torch.float64)
    b["K_A"] = b["K_A"] * scale
    b["crop_B"] = torch.tensor([[0, h, 0, w]] * B)
    return b


def synthetic_batch(B, H, W, seed=1, depth_kind="noise"):
    """Synthetic RGB-D batch of SURVEY.md section 8(d) / Appendix D (CPU tensors, dataset dict keys of
    data/my_main_dataset.py:195)."""
    g = torch.Generator().manual_seed(seed)

    def depth():
        if depth_kind == "noise":
            d = torch.rand(B, 1, H, W, generator=g) * 1.6 - 0.8
            d[torch.rand(B, 1, H, W, generator=g) < 0.05] = -1.0
            return d
        yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
        c = torch.rand(B, 5, generator=g) - 0.5
        d = (c[:, 0, None, None] * xx + c[:, 1, None, None] * yy + 0.5 * c[:, 2, None, None]
             + 0.2 * torch.sin(3.0 * xx * (1 + c[:, 3, None, None]) + 2.0 * yy * (1 + c[:, 4, None, None])))
        d = d.clamp(-0.9, 0.9)[:, None].contiguous()
        for b in range(B):
            for _ in range(int(torch.randint(3, 9, (1,), generator=g))):
                y0 = int(torch.randint(0, H - 8, (1,), generator=g)); x0 = int(torch.randint(0, W - 8, (1,), generator=g))
                hh = int(torch.randint(2, max(3, H // 8), (1,), generator=g)); ww = int(torch.randint(2, max(3, W // 8), (1,), generator=g))
                d[b, 0, y0:y0 + hh, x0:x0 + ww] = -1.0
        return d

    A_i = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    B_i = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    A_d, B_d = depth(), depth()
    K = torch.tensor([[577.87, 0, 319.5], [0, 577.87, 239.5], [0, 0, 1]], dtype=torch.float64).repeat(B, 1, 1)
    crop = torch.tensor([[0, H, 0, W]] * B)
    return dict(A_i=A_i, B_i=B_i, A_d=A_d, B_d=B_d, A_paths=["a"] * B, B_paths=["b"] * B,
                K_A=K, K_B=K.clone(), crop_A=crop, crop_B=crop.clone())
