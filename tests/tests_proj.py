import torch


def proj_vec(n, seed):
    """Fixed pseudo-random projection vector (same as tests/golden/make_golden.py:proj_vec)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, generator=g, dtype=torch.float64)
