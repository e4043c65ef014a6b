"""CPU: the index / weight formulas csrc/frontend.cu uses for F.interpolate (nearest: min(floor(dst * fp32(in / out)), in - 1);
bicubic: align_corners = False, A = -0.75, border-clamped taps), restated in numpy exactly as the kernels compute them, against
torch over many random (non-integer-ratio) size pairs - the GPU test compares the kernels themselves at a few shapes."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def nearest_src(dst, in_size, out_size):
    scale = np.float32(in_size) / np.float32(out_size)
    s = np.floor(dst.astype(np.float32) * scale).astype(np.int64)
    return np.minimum(s, in_size - 1)


def cubic_coeffs(t):
    A = np.float32(-0.75)
    c1 = lambda x: ((A + 2) * x - (A + 3)) * x * x + 1
    c2 = lambda x: ((A * x - 5 * A) * x + 8 * A) * x - 4 * A
    return np.stack([c2(t + 1), c1(t), c1(1 - t), c2(2 - t)], -1).astype(np.float32)


def bicubic(x, out):
    Hs, Ws = x.shape
    Ho, Wo = out
    def axis(n_in, n_out):
        r = np.float32(n_in) / np.float32(n_out) * (np.arange(n_out, dtype=np.float32) + np.float32(0.5)) - np.float32(0.5)
        f = np.floor(r)
        idx = np.clip(f.astype(np.int64)[:, None] - 1 + np.arange(4)[None], 0, n_in - 1)
        return idx, cubic_coeffs((r - f).astype(np.float32))
    iy, cy = axis(Hs, Ho)
    ix, cx = axis(Ws, Wo)
    rows = (x[:, ix] * cx[None]).sum(-1)                    # [Hs, Wo]
    return (rows[iy] * cy[:, :, None]).sum(1)               # [Ho, Wo]


@pytest.mark.parametrize("seed", range(6))
def test_nearest_matches_torch(seed):
    g = np.random.default_rng(seed)
    for _ in range(40):
        n_in, n_out = int(g.integers(1, 700)), int(g.integers(1, 700))
        x = torch.arange(n_in, dtype=torch.float32).view(1, 1, 1, n_in)
        ref = F.interpolate(x, (1, n_out)).view(-1).numpy().astype(np.int64)          # the source index torch picked
        assert np.array_equal(nearest_src(np.arange(n_out), n_in, n_out), ref), (n_in, n_out)


@pytest.mark.parametrize("seed", range(4))
def test_bicubic_matches_torch(seed):
    g = np.random.default_rng(100 + seed)
    for _ in range(12):
        Hs, Ws, Ho, Wo = (int(v) for v in g.integers(2, 140, 4))
        x = g.standard_normal((Hs, Ws)).astype(np.float32)
        ref = F.interpolate(torch.from_numpy(x)[None, None], (Ho, Wo), mode="bicubic")[0, 0].numpy()
        assert np.abs(bicubic(x, (Ho, Wo)) - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), (Hs, Ws, Ho, Wo)
