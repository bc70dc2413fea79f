"""K3 / K3c parity: Kabsch, rigid_transform_3D, affine fit, get_transformed_points."""
import numpy as np
import pytest

from conftest import load_golden, sub
from oracle import ref_port as O
from parity import assert_transform_parity

pytestmark = pytest.mark.gpu


def test_kabsch_golden():
    icp = sub("utils.icp")
    r3d = sub("utils.r3d")
    g = load_golden("kabsch")
    ext = O.scene_extent(g["A"])
    assert_transform_parity(icp.best_fit_transform(g["A"], g["B"]), g["T"], ext, "best_fit", 1e-10, 1e-12)
    assert_transform_parity(icp.best_fit_transform(g["A"], g["B_reflect"]), g["T_reflect"], ext, "reflect", 1e-10, 1e-12)
    for a, b, rk, tk in [("A", "B", "R1", "t1"), ("A", "B_reflect", "R2", "t2"), ("Ai", "Bi", "R3", "t3")]:
        R, t = r3d.rigid_transform_3D(g[a].T, g[b].T)
        assert R.shape == (3, 3) and t.shape == (3, 1) and R.dtype == np.float64
        np.testing.assert_allclose(R, g[rk], atol=1e-10)
        np.testing.assert_allclose(t, g[tk], atol=1e-8)
    with pytest.raises(Exception):
        r3d.rigid_transform_3D(g["A"], g["B"])
    with pytest.raises(AssertionError):
        r3d.rigid_transform_3D(g["A"].T, g["B"].T[:2])


def test_kabsch_random_properties():
    icp = sub("utils.icp")
    rng = np.random.default_rng(11)
    for trial in range(25):
        n = int(rng.integers(3, 4000))
        A = rng.normal(size=(n, 3)) * rng.uniform(0.1, 1000) + rng.normal(size=3) * 5000
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        w, x, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        t = rng.normal(size=3) * 100
        B = A @ R.T + t + rng.normal(size=A.shape) * rng.choice([0.0, 1e-3, 1.0])
        T = icp.best_fit_transform(A, B)
        Tr = O.best_fit_transform(A, B)
        assert_transform_parity(T, Tr, O.scene_extent(A), "trial %d" % trial, 1e-8, 1e-10)
        assert abs(np.linalg.det(T[:3, :3]) - 1) < 1e-10
        np.testing.assert_allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-10)


def test_kabsch_degenerate_sets():
    """Planar and collinear clouds: the rotation is still a proper rotation and still maps the
    centroids; the in-plane part must match the oracle (the free part is arbitrary in LAPACK too)."""
    icp = sub("utils.icp")
    rng = np.random.default_rng(3)
    A = rng.normal(size=(200, 3)) * 50
    A[:, 2] = 7.0                                   # planar
    ang = 0.4
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    B = A @ Rz.T + [1, 2, 3]
    T = icp.best_fit_transform(A, B)
    np.testing.assert_allclose(T, O.best_fit_transform(A, B), atol=1e-9)
    L = np.outer(np.linspace(-5, 5, 50), [1.0, 2.0, -1.0])     # collinear
    T = icp.best_fit_transform(L, L + [3, 0, 0])
    assert abs(np.linalg.det(T[:3, :3]) - 1) < 1e-9
    np.testing.assert_allclose((T[:3, :3] @ L.T).T + T[:3, 3], L + [3, 0, 0], atol=1e-8)


def test_large_best_fit():
    icp = sub("utils.icp")
    a, b = O.synth_depth_pair(2)
    A = O.depth_to_voxel_ld(a)
    rng = np.random.default_rng(0)
    B = A + rng.normal(size=A.shape) + [3, -2, 9]
    assert_transform_parity(icp.best_fit_transform(A, B), O.best_fit_transform(A, B), O.scene_extent(A), "276k", 1e-9, 1e-10)


def test_affine_fit_and_transform():
    t3d = sub("utils.transformation3d")
    g = load_golden("affine")
    h = t3d.homo_rigid_transform_3d(g["p"], g["q"])
    assert h.shape == (4, 4) and h.dtype == np.float64
    np.testing.assert_allclose(h, g["h"], rtol=1e-8, atol=1e-7)
    np.testing.assert_allclose(t3d.get_transformed_points(g["cloud"], g["h"]), g["cloud_h"], rtol=1e-13, atol=1e-9)
    rng = np.random.default_rng(9)
    for k in (4, 5, 50, 2000):
        p = rng.integers(0, 5000, size=(k, 3))
        M = np.eye(4)
        M[:3] += rng.normal(size=(3, 4)) * 0.05
        q = np.rint(np.concatenate((p, np.ones((k, 1))), axis=1) @ M.T)[:, :3].astype(np.int64)
        np.testing.assert_allclose(t3d.homo_rigid_transform_3d(p, q), O.homo_rigid_transform_3d(p, q), rtol=1e-6, atol=1e-5)
