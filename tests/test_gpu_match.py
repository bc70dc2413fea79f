"""K5 parity: cross-checked brute-force descriptor matching vs the reference run's recorded OpenCV output,
vs OpenCV itself, and vs the oracle."""
import numpy as np
import pytest

from conftest import load_golden, sub
from oracle import ref_port as O

pytestmark = pytest.mark.gpu


def _pairs(matches):
    return np.array([(m.queryIdx, m.trainIdx) for m in matches]).reshape(-1, 2), np.array([m.distance for m in matches], np.float32)


def test_sofa_descriptors_golden():
    t3d = sub("utils.transformation3d")
    g = load_golden("sofa_pair")
    m = t3d.match_descriptors(g["des1"].astype(np.float32), g["des2"].astype(np.float32))
    pairs, dist = _pairs(m)
    assert np.array_equal(pairs, g["bf"])
    assert np.array_equal(dist, g["bf_dist"])            # integer-valued descriptors: bit-exact distances
    assert all(hasattr(x, "queryIdx") and hasattr(x, "trainIdx") and hasattr(x, "distance") for x in m)


@pytest.mark.parametrize("n1,n2,dim,kind", [
    (1, 1, 128, "int"), (1, 40, 128, "int"), (37, 1, 128, "int"), (300, 257, 128, "int"), (1100, 1200, 128, "int"),
    (200, 333, 128, "ties"), (64, 64, 128, "dup"), (150, 90, 64, "int"), (90, 150, 33, "int"), (50, 60, 7, "ties"),
    (120, 80, 128, "float"),
])
def test_against_opencv_and_oracle(n1, n2, dim, kind):
    cv2 = pytest.importorskip("cv2")
    t3d = sub("utils.transformation3d")
    rng = np.random.default_rng(n1 * 31 + n2 + dim)
    if kind == "float":
        d1 = rng.normal(size=(n1, dim)).astype(np.float32)
        d2 = rng.normal(size=(n2, dim)).astype(np.float32)
    else:
        d1 = rng.integers(0, 256, (n1, dim)).astype(np.float32)
        d2 = rng.integers(0, 256, (n2, dim)).astype(np.float32)
    if kind == "ties":          # low-entropy descriptors: exact distance ties in both directions
        d1[:, 3:] = 0
        d2[:, 3:] = 0
        d1[:, :3] = rng.integers(0, 3, (n1, 3))
        d2[:, :3] = rng.integers(0, 3, (n2, 3))
    if kind == "dup":           # identical sets: every row is its own mutual neighbour, duplicates resolve to the first
        d2 = d1.copy()
        d2[10:20] = d2[0:10]
    pairs, dist = _pairs(t3d.match_descriptors(d1, d2))
    q, t, d = O.match_descriptors(d1, d2)
    want_pairs, want_dist = _pairs(cv2.BFMatcher(normType=cv2.NORM_L2, crossCheck=True).match(d1, d2))
    assert np.array_equal(np.stack((q, t), axis=1), want_pairs), "oracle vs OpenCV"
    assert np.array_equal(pairs, want_pairs), "CUDA vs OpenCV"
    if kind == "float":
        np.testing.assert_allclose(dist, want_dist, rtol=2e-6)      # summation order differs; indices still agree
    else:
        assert np.array_equal(dist, want_dist)


def test_argument_errors():
    t3d = sub("utils.transformation3d")
    assert t3d.match_descriptors(np.zeros((0, 128), np.float32), np.zeros((5, 128), np.float32)) == []
    with pytest.raises(ValueError):
        t3d.match_descriptors(np.zeros((4, 128), np.float32), np.zeros((5, 64), np.float32))


@pytest.mark.parametrize("n1,n2,dim,kind,ratio,cross", [
    (1, 1, 128, "int", 0.8, True), (40, 1, 128, "int", 0.8, True), (1, 40, 128, "int", 0.8, True),
    (300, 257, 128, "near", 0.8, True), (1100, 1200, 128, "near", 0.8, True), (200, 333, 128, "ties", 0.8, True),
    (64, 64, 128, "dup", 0.8, True), (150, 90, 64, "near", 0.6, False), (90, 150, 33, "near", 0.95, True),
    (120, 80, 128, "near", 1.0, True),
])
def test_ratio_test_against_skimage_restatement(n1, n2, dim, kind, ratio, cross):
    """q8.py:90: match_descriptors(des1, des2, max_ratio=0.8, cross_check=True)."""
    q8 = sub("utils.homography_utils.q8")
    rng = np.random.default_rng(n1 * 17 + n2 * 3 + dim)
    d1 = rng.integers(0, 256, (n1, dim)).astype(np.float32)
    d2 = rng.integers(0, 256, (n2, dim)).astype(np.float32)
    if kind == "near":          # a third of the train set are noisy copies of queries, so the ratio test has survivors
        m = min(n1, n2) // 3
        d2[:m] = np.clip(d1[rng.permutation(n1)[:m]] + rng.integers(-12, 13, (m, dim)), 0, 255)
    if kind == "ties":
        d1[:, 3:] = 0
        d2[:, 3:] = 0
        d1[:, :3] = rng.integers(0, 3, (n1, 3))
        d2[:, :3] = rng.integers(0, 3, (n2, 3))
    if kind == "dup":           # exact duplicates: best distance 0; a duplicated train row makes the second best 0 -> eps
        d2 = d1.copy()
        d2[10:20] = d2[0:10]
    got = q8.match_descriptors(d1, d2, max_ratio=ratio, cross_check=cross)
    want = O.match_descriptors_skimage(d1, d2, max_ratio=ratio, cross_check=cross)
    assert got.dtype == np.int64 and got.shape[1] == 2
    assert np.array_equal(got, want)
    if kind == "near" and ratio < 1.0 and cross:
        assert 0 < len(got) < n1


def test_mathching_skimage_window():
    q8 = sub("utils.homography_utils.q8")

    class KP:
        def __init__(self, x, y):
            self.pt = (x, y)

    rng = np.random.default_rng(5)
    d1 = rng.integers(0, 256, (200, 128)).astype(np.float32)
    d2 = np.clip(d1[rng.permutation(200)] + rng.integers(-5, 6, (200, 128)), 0, 255).astype(np.float32)
    kp1 = [KP(float(x), float(y)) for x, y in rng.uniform(0, 400, (200, 2))]
    full = q8.mathching_skimage(None, kp1, d1, None, None, d2)
    assert np.array_equal(full, O.match_descriptors_skimage(d1, d2, 0.8, True)) and len(full) > 100
    win = q8.mathching_skimage(None, kp1, d1, None, None, d2, x0=200, y0=200, window=100)
    want = [m for m in full if abs(kp1[m[0]].pt[0] - 200) <= 100 and abs(kp1[m[0]].pt[1] - 200) <= 100]
    assert len(win) == len(want) and all(np.array_equal(a, b) for a, b in zip(win, want))
