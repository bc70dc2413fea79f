"""Synthetic workload of BASELINE.json configs[2]: 640x480 depth pairs of a smooth surface.

Workload generator for bench.py and the tools, not part of the registration path.  depth_pair() is the NumPy
definition (seed = 1234 + pair index); both arms of the benchmark register these very frames for the pairs the
parity gate compares, and tests/test_host_logic.py checks it against the test oracle's generator bit for bit.
depth_batch_device() draws the same surface model with torch's generator on the GPU for the bulk of a 1024-pair
batch (NumPy needs ~20 ms per pair).
"""
import numpy as np

H, W = 480, 640
SEED_BASE = 1234


def _surface(u, v):
    return 2000.0 + 600.0 * np.sin(u / 97.0) + 400.0 * np.cos(v / 61.0)


def depth_pair(pair_index, h=H, w=W, seed_base=SEED_BASE):
    """(frame a, frame b) uint16: a = surface + N(0, 2) rounded, 10 % of the pixels dropped to 0; b = the surface
    shifted by integers du, dv in [-8, 8] and dz in [-20, 20] with fresh noise and dropout.  Draw order: (du, dv), dz,
    noise a, dropout a, noise b, dropout b."""
    rng = np.random.default_rng(seed_base + pair_index)
    u = np.arange(w, dtype=np.float64)[None, :]
    v = np.arange(h, dtype=np.float64)[:, None]
    du, dv = rng.integers(-8, 9, size=2)
    dz = rng.integers(-20, 21)
    frames = []
    for uu, vv, z in ((u, v, None), (u + du, v + dv, dz)):
        f = _surface(uu, vv) if z is None else _surface(uu, vv) + z
        f = np.rint(f + rng.normal(0.0, 2.0, size=(h, w)))
        f[rng.random((h, w)) < 0.10] = 0
        frames.append(f.astype(np.uint16))
    return frames[0], frames[1]


def depth_pairs(first, count, h=H, w=W):
    """Stacked uint16 [count, h, w] x 2 for pairs first .. first + count - 1."""
    a = np.empty((count, h, w), np.uint16)
    b = np.empty((count, h, w), np.uint16)
    for k in range(count):
        a[k], b[k] = depth_pair(first + k, h, w)
    return a, b


def depth_batch_device(n_pairs, seed, device, h=H, w=W):
    """The same surface model drawn on the device: uint16 [n_pairs, h, w] x 2."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(SEED_BASE + seed)
    u = torch.arange(w, device=device, dtype=torch.float32)[None, None, :]
    v = torch.arange(h, device=device, dtype=torch.float32)[None, :, None]
    a = torch.empty((n_pairs, h, w), dtype=torch.uint16, device=device)
    b = torch.empty((n_pairs, h, w), dtype=torch.uint16, device=device)
    for s in range(0, n_pairs, 64):
        e = min(n_pairs, s + 64)
        k = e - s
        du = torch.randint(-8, 9, (k, 1, 1), generator=g, device=device).float()
        dv = torch.randint(-8, 9, (k, 1, 1), generator=g, device=device).float()
        dz = torch.randint(-20, 21, (k, 1, 1), generator=g, device=device).float()
        fa = 2000.0 + 600.0 * torch.sin(u / 97.0) + 400.0 * torch.cos(v / 61.0) + torch.randn((k, h, w), generator=g, device=device) * 2.0
        fb = (2000.0 + 600.0 * torch.sin((u + du) / 97.0) + 400.0 * torch.cos((v + dv) / 61.0) + dz
              + torch.randn((k, h, w), generator=g, device=device) * 2.0)
        fa = torch.round(fa)
        fb = torch.round(fb)
        fa[torch.rand((k, h, w), generator=g, device=device) < 0.10] = 0
        fb[torch.rand((k, h, w), generator=g, device=device) < 0.10] = 0
        a[s:e] = fa.to(torch.int32).to(torch.uint16)
        b[s:e] = fb.to(torch.int32).to(torch.uint16)
    return a, b
