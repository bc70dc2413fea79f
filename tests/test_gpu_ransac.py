"""K4 parity: 2-D homography RANSAC scoring (reference-exact) and 3-D hypothesis scoring."""
import warnings

import numpy as np
import pytest

from conftest import load_golden, sub
from oracle import ref_port as O

pytestmark = pytest.mark.gpu


class KP:
    def __init__(self, x, y):
        self.pt = (float(x), float(y))


def test_q9_pieces_golden():
    q9 = sub("utils.homography_utils.q9")
    g = load_golden("ransac2d")
    Hh, Ww = g["img_shape"]
    orig = g["xy1"][:, ::-1].copy()
    prime = g["xy2"][:, ::-1].copy()
    binned = q9.make_binned_array((Hh, Ww), prime)
    assert len(binned) == Hh and len(binned[0]) == Ww
    for k in range(3):
        tp = q9.get_transformed_points(orig, g["Hs"][k])
        np.testing.assert_allclose(tp, g["tp%d" % k], rtol=1e-13, atol=1e-11)
        m, s = q9.get_ssd_v2(orig, g["tp%d" % k], prime, binned)
        assert np.array_equal(m, g["m%d" % k])
        np.testing.assert_allclose(s, g["s%d" % k], rtol=1e-12, atol=0)


def test_scoring_counts_match_oracle_for_many_hypotheses():
    ops = sub("ops")
    g = load_golden("ransac2d")
    shape = tuple(g["img_shape"])
    orig = g["xy1"][:, ::-1].copy()
    prime = g["xy2"][:, ::-1].copy()
    np.random.seed(11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Hs, _ = O.ransac_hypotheses(g["xy1"], g["xy2"], g["bf"], n_iter=60)
        binned = O.make_binned_array(shape, prime)
        want, want_m = [], []
        for H in Hs:
            tp = O.q9_get_transformed_points(orig, H)
            m, s = O.get_ssd_v2(orig, tp, prime, binned)
            want.append(len(m))
            want_m.append(m)
    counts, flags, mj, ms = ops.ransac2d_score(Hs, orig, prime, shape, want_matches=True)
    counts = counts.cpu().numpy()
    assert np.array_equal(counts, np.array(want)), "inlier counts must match exactly"
    mj = mj.cpu().numpy()
    for h in (0, int(np.argmax(want)), 59):
        i = np.nonzero(mj[h] >= 0)[0]
        got = np.stack((i, mj[h][i]), axis=1) if len(i) else np.zeros((0, 2), int)
        ref = want_m[h] if len(want_m[h]) else np.zeros((0, 2), int)
        assert np.array_equal(got, ref)


def test_ransac_loop_golden_seeded():
    q9 = sub("utils.homography_utils.q9")
    g = load_golden("ransac2d")
    kp1 = [KP(*p) for p in g["xy1"]]
    kp2 = [KP(*p) for p in g["xy2"]]
    img = np.zeros(tuple(g["img_shape"]), np.uint8)
    np.random.seed(7)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        H, fm = q9.ransac_loop(img, img, kp1, kp2, g["bf"])
    after = np.random.randint(1 << 30)
    assert np.array_equal(fm, g["final_matches"])
    np.testing.assert_allclose(H, g["H_final"], rtol=1e-9, atol=1e-12)
    # the global RNG must have been consumed exactly as the reference consumes it
    np.random.seed(7)
    for _ in range(100):
        np.random.randint(len(g["bf"]), size=4)
    assert after == np.random.randint(1 << 30)


def test_binned_array_wrap_and_errors():
    q9 = sub("utils.homography_utils.q9")
    kp2 = np.array([[0.4, 0.7], [30.2, 0.2], [10.5, 20.5]])        # first two wrap to the last row/col
    tp = np.array([[118.5, 158.2], [29.0, 158.9], [10.0, 20.0], [np.nan, 1.0]])
    binned = q9.make_binned_array((120, 160), kp2)
    with pytest.raises(ValueError):
        q9.get_ssd_v2(tp, tp, kp2, binned)
    m, s = q9.get_ssd_v2(tp[:3], tp[:3], kp2, binned)
    mo, so = O.get_ssd_v2(tp[:3], tp[:3], kp2, O.make_binned_array((120, 160), kp2))
    assert np.array_equal(m, mo)
    with pytest.raises(IndexError):
        q9.make_binned_array((120, 160), np.array([[130.0, 5.0]]))
    # a homography with w == 0 for some keypoint: the reference dies in int(); so do we
    H = np.array([[1.0, 0, 0], [0, 1, 0], [0, 0, 0]])
    with pytest.raises((ValueError, OverflowError)):
        sub("ops")
        counts, flags, _, _ = sub("ops").ransac2d_score(H[None], tp[:3], kp2, (120, 160))
        q9._raise_for_flags(flags.cpu().numpy())


def test_complex_homography_matches_numpy():
    """numpy.linalg.eig can hand q9 a complex H; only the final assignment drops the imaginary part."""
    q9 = sub("utils.homography_utils.q9")
    g = load_golden("ransac2d")
    Hc = g["dlt_H"]
    k = int(np.argmax(np.abs(np.imag(Hc)).reshape(len(Hc), -1).max(axis=1)))
    assert np.abs(np.imag(Hc[k])).max() > 0
    orig = g["xy1"][:, ::-1].copy()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = O.q9_get_transformed_points(orig, Hc[k])
    got = q9.get_transformed_points(orig, Hc[k])
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("kind", ["rigid", "affine"])
def test_ransac3d_counts(kind):
    ops = sub("ops")
    import torch
    rng = np.random.default_rng(11)
    n = 5000
    P = rng.uniform(-1000, 1000, (n, 3)).astype(np.float32).astype(np.float64)
    ang = 0.2
    T = np.eye(4)
    T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    T[:3, 3] = [15, -8, 5]
    Q = O.get_transformed_points(P, T) + rng.normal(size=(n, 3))
    out = rng.random(n) < 0.4
    Q[out] = rng.uniform(-1000, 1000, (out.sum(), 3))
    samples = np.random.RandomState(12345).randint(n, size=(300, 3 if kind == "rigid" else 4))
    samples[:5] = np.nonzero(~out)[0][:5 * samples.shape[1]].reshape(5, -1)     # some all-inlier samples
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        hyps = O.rigid_hypotheses(P, Q, samples) if kind == "rigid" else O.affine_hypotheses(P, Q, samples)
    hyps = np.concatenate((hyps, T[None]))
    want = O.score_hypotheses_3d(hyps, P, Q)
    got = ops.ransac3d_score(torch.from_numpy(hyps).cuda(), ops.pack_xyzw(torch.from_numpy(P).cuda()),
                             ops.pack_xyzw(torch.from_numpy(Q).cuda())).cpu().numpy()
    assert want[-1] > 0.4 * n
    assert np.array_equal(got, want)


def test_ransac3d_ragged_sizes():
    ops = sub("ops")
    import torch
    rng = np.random.default_rng(3)
    for n, nh in [(1, 1), (511, 3), (513, 129), (4097, 1000)]:
        P = rng.normal(size=(n, 3)) * 3
        Q = P + rng.normal(size=(n, 3)) * 2
        hyps = np.tile(np.eye(4), (nh, 1, 1))
        hyps[:, :3, 3] = rng.normal(size=(nh, 3))
        want = O.score_hypotheses_3d(hyps, P, Q)
        got = ops.ransac3d_score(torch.from_numpy(hyps).cuda(), ops.pack_xyzw(torch.from_numpy(P).cuda()),
                                 ops.pack_xyzw(torch.from_numpy(Q).cuda())).cpu().numpy()
        assert np.array_equal(got, want), (n, nh)


@pytest.mark.parametrize("kind", ["rigid", "affine"])
def test_device_hypotheses_match_the_reference_estimators(kind):
    """K6: hypotheses generated on the device from minimal samples vs the oracle's per-sample calls of the
    reference estimators (r3d.rigid_transform_3D / homo_rigid_transform_3d), and the full device-resident
    sweep (generate + score) vs host-generated hypotheses scored by the oracle."""
    import contextlib
    import io
    import torch
    ops = sub("ops")
    rng = np.random.default_rng(21)
    n = 20000
    P = rng.uniform(-1000, 1000, (n, 3)).astype(np.float32).astype(np.float64)
    ang = 0.25
    T = np.eye(4)
    T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    T[:3, 3] = [12, -7, 30]
    Q = O.get_transformed_points(P, T) + rng.normal(size=(n, 3))
    k = 3 if kind == "rigid" else 4
    samples = np.random.RandomState(5).randint(n, size=(3000, k)).astype(np.int32)
    samples[7, 1] = samples[7, 0]                      # a repeated index: degenerate
    Pd = ops.pack_xyzw(torch.from_numpy(P).cuda())
    Qd = ops.pack_xyzw(torch.from_numpy(Q).cuda())
    Td, status = ops.hypotheses_3d(Pd, Qd, torch.from_numpy(samples).cuda(), kind)
    Th, status = Td.cpu().numpy(), status.cpu().numpy()
    assert status[7] == 1 and status.sum() < 30
    ok = np.nonzero(status == 0)[0][:600]
    with contextlib.redirect_stdout(io.StringIO()):
        want = O.rigid_hypotheses(P, Q, samples[ok]) if kind == "rigid" else O.affine_hypotheses(P, Q, samples[ok])
    if kind == "rigid":
        np.testing.assert_allclose(Th[ok], want, rtol=0, atol=1e-7)
        assert np.allclose(np.linalg.det(Th[ok][:, :3, :3]), 1.0, atol=1e-12)
    else:
        # exactly determined systems: GE vs LAPACK lstsq differ by conditioning x eps
        scale = np.maximum(np.abs(want).max(axis=(1, 2), keepdims=True), 1.0)
        assert np.median(np.abs(Th[ok] - want) / scale) < 1e-10
        assert (np.abs(Th[ok] - want) / scale).max() < 1e-5
    # device-resident sweep: counts of the device hypotheses = oracle counts of the same matrices
    counts = ops.ransac3d_score(Td, Pd, Qd).cpu().numpy()
    assert np.array_equal(counts[ok], O.score_hypotheses_3d(Th[ok], P, Q))


def test_k4b_against_the_reference_golden():
    """The pin of K4B (BASELINE.json configs[4] at 1/10 size, tests/golden/make_golden_k4b.py): 2 000 rigid and 2 000 affine
    hypotheses x 5 000 correspondences.  Hypotheses, residuals and the < 10 threshold were computed by the reference's OWN
    rigid_transform_3D / homo_rigid_transform_3d / get_transformed_points; the scoring kernel must return the same inlier
    counts for the reference's hypotheses, and the hypothesis kernels (K6) the same matrices from the same samples."""
    import torch
    from conftest import load_golden
    ops = sub("ops")
    g = load_golden("k4b")
    P, Q = g["P"].astype(np.float64), g["Q"]
    Pd = ops.pack_xyzw(torch.from_numpy(P).cuda())
    Qd = ops.pack_xyzw(torch.from_numpy(Q).cuda())
    star = ops.ransac3d_score(torch.from_numpy(g["T_star"][None].copy()).cuda(), Pd, Qd).cpu().numpy()
    assert int(star[0]) == int(g["count_star"])
    for kind, T, want, samples in (("rigid", g["T_rigid"], g["counts_rigid"], g["samples3"]),
                                   ("affine", g["T_affine"], g["counts_affine"], g["samples4"])):
        fin = np.isfinite(T).all(axis=(1, 2))
        got = ops.ransac3d_score(torch.from_numpy(np.ascontiguousarray(T[fin])).cuda(), Pd, Qd).cpu().numpy()
        assert np.array_equal(got, want[fin]), "%s: %d of %d inlier counts differ from the reference's" % (
            kind, (got != want[fin]).sum(), fin.sum())
        Td, status = ops.hypotheses_3d(Pd, Qd, torch.from_numpy(samples).cuda(), kind)
        Th, ok = Td.cpu().numpy(), (status.cpu().numpy() == 0) & fin
        assert ok.sum() > 1950
        scale = np.maximum(np.abs(T[ok]).max(axis=(1, 2), keepdims=True), 1.0)
        err = np.abs(Th[ok] - T[ok]) / scale
        assert np.median(err) < 1e-10 and err.max() < 1e-5, (kind, np.median(err), err.max())
        # device-generated hypotheses near T* find (almost) the same consensus set; the bulk scores 0-2 inliers either way
        dev_counts = ops.ransac3d_score(Td, Pd, Qd).cpu().numpy()
        assert (np.abs(dev_counts[ok] - want[ok]) <= np.maximum(2, want[ok] // 50)).all()


def test_ransac3d_fp32_filter_equals_the_fp64_kernel():
    """r3d_ransac3d_score's opt-in kernel (R3D_OPT_K4B_FP32) filters in fp32 and decides in fp64 inside the rounding band; its
    counts must be bit for bit those of the default all-fp64 kernel -- also when MOST residuals sit at the threshold (hypotheses
    within 1e-3 of the true motion, threshold = the median residual), far from the origin, and for scaled / degenerate
    hypotheses."""
    import torch
    ops = sub("ops")
    L = sub("_capi").lib()
    rng = np.random.default_rng(4)
    n = 30000
    for offset, spread in ((0.0, 1000.0), (5.0e4, 300.0), (0.0, 3.0)):
        P = rng.uniform(-spread, spread, (n, 3)) + offset
        ang = 0.1
        T = np.eye(4)
        T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
        T[:3, 3] = [3.0, -2.0, 1.0]
        Q = O.get_transformed_points(P, T) + rng.normal(size=(n, 3)) * np.sqrt(10.0 / 3.0)      # E[r2] = 10: the threshold
        hyps = np.tile(T, (600, 1, 1))
        hyps[:, :3, 3] += rng.normal(size=(600, 3)) * 1e-3
        hyps[:200, :3, :3] += rng.normal(size=(200, 3, 3)) * 1e-6
        hyps[590:] = np.eye(4) * np.array([1, 1, 1, 1.0])[None, :, None] * rng.uniform(0.5, 2.0, size=(10, 1, 1))
        hyps[599] = 0.0
        Pd, Qd = ops.pack_xyzw(torch.from_numpy(P).cuda()), ops.pack_xyzw(torch.from_numpy(Q).cuda())
        Hd = torch.from_numpy(hyps).cuda()
        for thresh in (10.0, 2.5, 400.0):
            want = ops.ransac3d_score(Hd, Pd, Qd, thresh=thresh).cpu().numpy()
            try:
                assert L.r3d_set_option(8, 1.0) == 0
                got = ops.ransac3d_score(Hd, Pd, Qd, thresh=thresh).cpu().numpy()
            finally:
                L.r3d_set_option(8, 0.0)
            assert np.array_equal(got, want), (offset, spread, thresh, int((got != want).sum()))
            if thresh == 10.0:
                assert 0.2 * n < want[0] < 0.9 * n      # the threshold cuts through the bulk of the residuals
