"""K1 parity: CUDA back-projection vs the reference's outputs (golden) and the oracle."""
import numpy as np
import pytest

from conftest import load_golden, sub
from oracle import ref_port as O

pytestmark = pytest.mark.gpu


def test_golden_bit_exact():
    U = sub("utils.utils")
    g = load_golden("backproject")
    small = g["small"]
    for got, key in [
        (U.depth_to_voxel_ld(small), "small_ld"),
        (U.depth_to_voxel_ld(small, 0.5), "small_ld_half"),
        (U.depth_to_voxel(small), "small_px"),
        (U.depth_to_voxel(small, 0.1), "small_px_tenth"),
        (U.depth_to_voxel(small, 1.), "small_px_t3d"),
        (U.depth_to_voxel(g["fimg"]), "fimg_px"),
        (U.depth_to_voxel(g["fimg"], 2.5), "fimg_px_scaled"),
        (U.depth_to_voxel_ld(g["crop"]), "crop_ld"),
    ]:
        want = g[key]
        assert got.dtype == want.dtype, key
        assert got.shape == want.shape, key
        assert np.array_equal(got, want), key


def test_full_frame_vs_oracle_and_batch():
    R = sub("reconstruct")
    frames = np.stack([O.synth_depth_pair(k)[0] for k in range(3)] + [O.synth_depth_pair(0)[1]])
    frames[1, :, :] = 0                      # an all-empty frame
    frames[2, 100:200, :] = 40000            # int16 wrap: negative z, kept
    got = R.depth_images_to_3d_pts_ld(frames)
    want = O.depth_images_to_3d_pts_ld(frames)
    assert len(got) == len(want) == 4
    for a, b in zip(got, want):
        assert a.shape == b.shape and a.dtype == b.dtype
        assert np.array_equal(a, b)
    assert got[1].shape == (0, 3)
    got = R.depth_images_to_3d_pts(frames, scale=0.1)
    want = O.depth_images_to_3d_pts(frames, scale=0.1)
    for a, b in zip(got, want):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    got = R.depth_images_to_3d_pts(list(frames))      # scale defaults to 1. (float)
    want = O.depth_images_to_3d_pts(list(frames))
    for a, b in zip(got, want):
        assert a.dtype == b.dtype and np.array_equal(a, b)


def test_ragged_sizes_and_layouts():
    import torch
    ops = sub("ops")
    rng = np.random.default_rng(5)
    for (h, w) in [(1, 1), (3, 7), (33, 61), (480, 640), (31, 2049)]:
        img = rng.integers(0, 5000, size=(2, h, w)).astype(np.uint16)
        img[rng.random(img.shape) < 0.3] = 0
        d = torch.from_numpy(img).cuda()
        pts, off = ops.backproject(d, layout=ops.LAYOUT_XYZW)
        off = off.cpu().numpy()
        for f in range(2):
            want = O.depth_to_voxel_ld(img[f])
            got = pts[off[f]:off[f + 1]].cpu().numpy()
            assert np.array_equal(got[:, :3], want)
            # w lane carries the point's index inside its frame
            assert np.array_equal(got[:, 3].view(np.int64) & 0xffffffff, np.arange(len(want)))
        p32, off32 = ops.backproject(d, layout=ops.LAYOUT_XYZW_F32)
        got32 = p32[:int(off32[-1])].cpu().numpy()
        want32 = np.concatenate([O.depth_to_voxel_ld(img[f]) for f in range(2)]).astype(np.float32)
        assert np.array_equal(got32[:, :3], want32)


def test_fast_division_is_exact_over_the_whole_uint16_domain():
    """K1 divides by the reference's fx = fy = 525 with a reciprocal multiply and one fused correction instead of the
    IEEE division routine.  That is bit-exact only if it returns the correctly rounded quotient for every numerator the
    domain can produce: (u - 319.5) * d for u in 0..639 and (v - 239.5) * d for v in 0..479, d in 0..65535.  Enumerate
    all of them against NumPy's division."""
    import torch
    ops = sub("ops")
    u = np.arange(640, dtype=np.int64)[None, None, :]
    v = np.arange(480, dtype=np.int64)[None, :, None]
    plans = [(137, lambda f: (f * 480 + v) + 0 * u),          # every d for every column u
             (103, lambda f: (f * 640 + u) + 0 * v)]          # every d for every row v
    for n_frames, pattern in plans:
        for first in range(0, n_frames, 16):
            f = np.arange(first, min(n_frames, first + 16), dtype=np.int64)[:, None, None]
            depth = (pattern(f) & 0xffff).astype(np.uint16)
            pts, off = ops.backproject(torch.from_numpy(depth).cuda(), mode=ops.BP_PINHOLE, layout=ops.LAYOUT_XYZ)
            got, off = pts.cpu().numpy(), off.cpu().numpy()
            for k in range(depth.shape[0]):
                want = O.depth_to_voxel_ld(depth[k])
                assert off[k + 1] - off[k] == want.shape[0]
                assert np.array_equal(got[off[k]:off[k + 1]], want), (first + k)


def test_integer_scale_and_wide_integer_depth_follow_numpy():
    """depth_to_voxel / depth_to_voxel_ld with an INTEGER scale other than 1 (int16 arithmetic that wraps, rows whose wrapped z
    is zero dropped) and with integer depth images outside the uint16 range (x / y from the full value, z from its low 16
    bits): the drop-ins must return what the reference's NumPy expressions return (oracle = those expressions)."""
    import warnings
    utils = sub("utils.utils")
    rng = np.random.default_rng(9)
    img = rng.integers(0, 3000, size=(37, 53)).astype(np.uint16)
    img[rng.random(img.shape) < 0.1] = 0
    img[0, :4] = [256, 512, 16384, 32768]            # x 256 wraps to 0 in int16: those rows disappear
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for scale in (3, 256, -7):
            for fn_name in ("depth_to_voxel", "depth_to_voxel_ld"):
                want = getattr(O, fn_name)(img, scale)
                got = getattr(utils, fn_name)(img, scale)
                assert got.dtype == want.dtype and got.shape == want.shape, (fn_name, scale, got.shape, want.shape)
                assert np.array_equal(got, want), (fn_name, scale)
        wide = rng.integers(-70000, 200000, size=(29, 31)).astype(np.int32)
        wide[::3, ::2] = 0
        for fn_name in ("depth_to_voxel", "depth_to_voxel_ld"):
            want = getattr(O, fn_name)(wide)
            got = getattr(utils, fn_name)(wide)
            assert got.shape == want.shape and np.array_equal(got, want.astype(got.dtype)), fn_name
