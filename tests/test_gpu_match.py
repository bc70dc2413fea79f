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
