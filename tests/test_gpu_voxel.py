"""K7 parity: voxel down-sampling and the merge loop of chain_transformation (reconstruct.py:201-220) vs the NumPy
restatement of Open3D's published algorithm (oracle/ref_port.py; Open3D itself is absent: parity unpinned against it).
Bit-exact: integer voxel indices, sequential in-order sums, one IEEE division per coordinate."""
import numpy as np
import pytest

from conftest import sub
from oracle import ref_port as O

pytestmark = pytest.mark.gpu


def _cloud(kind, n, rng):
    if kind == "uniform":
        return rng.uniform(-300.0, 700.0, (n, 3))
    if kind == "surface":          # a depth-like sheet: many points per voxel column
        u, v = rng.uniform(0, 640, n), rng.uniform(0, 480, n)
        return np.column_stack(((u - 319.5) * 3.8, (v - 239.5) * 3.8, 2000 + 600 * np.sin(u / 97) + rng.normal(0, 2, n)))
    if kind == "integers":         # points exactly on voxel boundaries and exact duplicates
        return rng.integers(-40, 40, (n, 3)).astype(np.float64)
    if kind == "wide":             # 19-bit voxel coordinates: three radix passes per axis
        p = rng.uniform(0.0, 1.0e6, (n, 3))
        p[: n // 2] = p[n // 2: 2 * (n // 2)] + rng.uniform(-0.4, 0.4, (n // 2, 3))
        return p
    if kind == "line":             # one axis only varies: single-pass sorts on the others
        return np.column_stack((rng.uniform(0, 5000, n), np.full(n, 3.25), np.full(n, -7.5)))
    raise ValueError(kind)


@pytest.mark.parametrize("kind,n,voxel", [
    ("uniform", 1, 2.0), ("uniform", 2, 2.0), ("uniform", 1000, 2.0), ("uniform", 5000, 50.0), ("uniform", 3000, 1e5),
    ("surface", 300_000, 2.0), ("surface", 100_000, 5.0), ("surface", 50_000, 0.5),
    ("integers", 20_000, 2.0), ("integers", 20_000, 1.0), ("integers", 5_000, 0.25),
    ("wide", 40_000, 2.0), ("line", 10_000, 2.0),
])
def test_voxel_down_sample_bit_exact(kind, n, voxel):
    rec = sub("reconstruct")
    pts = _cloud(kind, n, np.random.default_rng(n + int(voxel * 8)))
    got = rec.voxel_down_sample(pts, voxel)
    want = O.voxel_down_sample(pts, voxel)
    assert got.shape == want.shape and got.dtype == np.float64
    assert np.array_equal(got, want)


def test_voxel_down_sample_full_size_properties():
    """2 M points (configs[3] scale): bit-exact against the oracle, voxel count = distinct voxel triples, and the
    count-weighted sum of the outputs returns the sum of the inputs."""
    rec = sub("reconstruct")
    rng = np.random.default_rng(7)
    pts = _cloud("surface", 2_000_000, rng)
    got = rec.voxel_down_sample(pts, 2.0)
    minb = pts.min(axis=0) - 1.0
    vox = np.floor((pts - minb) / 2.0).astype(np.int64)
    key = (vox[:, 2] << 42) | (vox[:, 1] << 21) | vox[:, 0]
    uniq, counts = np.unique(key, return_counts=True)
    assert len(got) == len(uniq)
    np.testing.assert_allclose((got * counts[:, None]).sum(axis=0), pts.sum(axis=0), rtol=1e-12)
    gv = np.floor((got - minb) / 2.0).astype(np.int64)          # a mean stays inside its (convex) voxel ...
    gk = (gv[:, 2] << 42) | (gv[:, 1] << 21) | gv[:, 0]
    assert (gk == uniq).mean() > 0.999                           # ... up to rounding on a boundary; order is (z, y, x)
    assert np.array_equal(got, O.voxel_down_sample(pts, 2.0))


def test_voxel_errors_and_empty():
    rec = sub("reconstruct")
    assert rec.voxel_down_sample(np.zeros((0, 3)), 2.0).shape == (0, 3)
    with pytest.raises(RuntimeError, match="voxel_size <= 0"):
        rec.voxel_down_sample(np.zeros((4, 3)), 0.0)
    big = np.array([[0.0, 0.0, 0.0], [1.0e10, 0.0, 0.0]])
    with pytest.raises(RuntimeError, match="too small"):
        rec.voxel_down_sample(big, 1.0)
    with pytest.raises(RuntimeError, match="too small"):
        O.voxel_down_sample(big, 1.0)
    bad = np.zeros((10, 3))
    bad[3, 1] = np.nan
    with pytest.raises(ValueError):
        rec.voxel_down_sample(bad, 2.0)
    with pytest.raises(ValueError):
        rec.voxel_down_sample(np.zeros((4, 2)), 2.0)


def test_merge_point_clouds_exact_transforms():
    """Transforms whose arithmetic is exact (signed permutations + integer translations on quarter-integer points), so
    the device's and NumPy's transformed clouds are identical and the whole merge loop must be bit-exact."""
    rec = sub("reconstruct")
    rng = np.random.default_rng(11)
    clouds = [np.round(_cloud("surface", 60_000, rng) * 4) / 4 for _ in range(4)]
    Ts = []
    for k in range(3):
        T = np.eye(4)
        perm = rng.permutation(3)
        T[:3, :3] = np.eye(3)[perm] * rng.choice([-1.0, 1.0], 3)[:, None]
        T[:3, 3] = rng.integers(-50, 50, 3)
        Ts.append(T)
    got = rec.merge_point_clouds(clouds, Ts, voxel_size=2)
    want = O.merge_point_clouds(clouds, Ts, voxel_size=2)
    assert np.array_equal(got, want)
    assert len(got) < sum(len(c) for c in clouds)


def test_merge_point_clouds_general_transforms():
    """General rigid transforms: the transformed coordinates agree to rounding (fp64 products summed in the same order,
    FMA contraction aside), so a point within an ulp of a voxel face may change voxel: compare the merged clouds as
    sets with that allowance."""
    rec = sub("reconstruct")
    rng = np.random.default_rng(12)
    clouds = [_cloud("surface", 40_000, rng) for _ in range(3)]
    Ts = []
    for k in range(2):
        a = rng.uniform(-0.05, 0.05)
        T = np.eye(4)
        T[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        T[:3, 3] = rng.uniform(-20, 20, 3)
        Ts.append(T)
    got = rec.merge_point_clouds(clouds, Ts, voxel_size=2)
    want = O.merge_point_clouds(clouds, Ts, voxel_size=2)
    assert abs(len(got) - len(want)) <= 4
    if len(got) == len(want):
        assert np.mean(np.all(np.abs(got - want) < 1e-9, axis=1)) > 0.999
    with pytest.raises(ValueError):
        rec.merge_point_clouds(clouds, Ts[:1])
    P = np.eye(4)
    P[3, 0] = 0.1
    with pytest.raises(ValueError):
        rec.merge_point_clouds(clouds, [P, P])
