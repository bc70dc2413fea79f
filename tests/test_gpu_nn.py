"""K2 parity: exact grid nearest neighbour vs KD-tree (reference call) and brute force."""
import numpy as np
import pytest

from conftest import load_golden, sub
from oracle import ref_port as O
from parity import assert_nn_parity

pytestmark = pytest.mark.gpu


def test_golden():
    icp = sub("utils.icp")
    g = load_golden("nearest_neighbor")
    d, i = icp.nearest_neighbor(g["src"], g["dst"])
    assert d.dtype == np.float64 and i.dtype == np.int64 and d.shape == i.shape == (1500,)
    assert np.array_equal(i, g["idx"])
    assert_nn_parity(g["src"], g["dst"], i, d, g["idx"], g["dist"], "random")
    d, i = icp.nearest_neighbor(g["lat_src"], g["lat_dst"])
    nties = assert_nn_parity(g["lat_src"], g["lat_dst"], i, d, g["lat_idx"], g["lat_dist"], "lattice")
    assert nties > 0
    # lowest-index tie-break == brute force, bit for bit
    db, ib = O.nearest_neighbor_brute(g["lat_src"], g["lat_dst"])
    assert np.array_equal(i, ib)


@pytest.mark.parametrize("n,m,kind", [
    (1, 1, "gauss"), (5, 1, "gauss"), (1, 7, "gauss"), (300, 2, "gauss"), (4000, 3500, "gauss"),
    (3000, 3000, "uniform"), (2500, 2600, "surface"), (3000, 2000, "far"), (2000, 2000, "dup"),
    (2000, 1500, "line"), (1500, 1500, "identical"),
])
def test_against_brute_force(n, m, kind):
    icp = sub("utils.icp")
    rng = np.random.default_rng(n * 7 + m)
    if kind == "gauss":
        src, dst = rng.normal(size=(n, 3)) * 100, rng.normal(size=(m, 3)) * 100
    elif kind == "uniform":
        src, dst = rng.uniform(-1000, 1000, size=(n, 3)), rng.uniform(-1000, 1000, size=(m, 3))
    elif kind == "surface":
        uv = rng.uniform(-500, 500, size=(m, 2))
        dst = np.column_stack((uv, 2000 + 100 * np.sin(uv[:, 0] / 50)))
        uv = rng.uniform(-500, 500, size=(n, 2))
        src = np.column_stack((uv, 2010 + 100 * np.sin(uv[:, 0] / 50)))
    elif kind == "far":     # queries far outside the destination's bounding box, some enormous
        dst = rng.normal(size=(m, 3)) * 10
        src = rng.normal(size=(n, 3)) * 10 + rng.choice([0, 500, -1e5, 1e7], size=(n, 1))
    elif kind == "dup":     # many duplicate points: exact ties everywhere
        base = rng.integers(0, 6, size=(40, 3)).astype(float)
        dst = base[rng.integers(0, 40, m)]
        src = base[rng.integers(0, 40, n)] + rng.choice([0.0, 0.5], size=(n, 1))
    elif kind == "line":    # degenerate extent: a cloud on a line
        dst = np.outer(rng.uniform(0, 100, m), [1.0, 0.0, 0.0])
        src = rng.normal(size=(n, 3)) * 30
    else:                   # all destination points identical
        dst = np.tile([[3.0, 4.0, 5.0]], (m, 1))
        src = rng.normal(size=(n, 3))
    d, i = icp.nearest_neighbor(src, dst)
    db, ib = O.nearest_neighbor_brute(src, dst)
    assert np.array_equal(i, ib), kind
    np.testing.assert_allclose(d, db, rtol=1e-14, atol=0)


def test_full_density_frame_vs_kdtree():
    """276k x 276k: the reference's KD-tree call on back-projected synthetic frames."""
    icp = sub("utils.icp")
    a, b = O.synth_depth_pair(1)
    A, B = O.depth_to_voxel_ld(a), O.depth_to_voxel_ld(b)
    d, i = icp.nearest_neighbor(A, B)
    dr, ir = O.nearest_neighbor(A, B)
    assert_nn_parity(A, B, i, d, ir, dr, "full frame")


def test_explicit_cell_sizes():
    """The answer must not depend on the grid resolution."""
    ops = sub("ops")
    C = sub("_capi")
    rng = np.random.default_rng(1)
    src, dst = rng.normal(size=(3000, 3)) * 50, rng.normal(size=(3000, 3)) * 50
    db, ib = O.nearest_neighbor_brute(src, dst)
    s = ops.pack_xyzw(C.to_device(src))
    d = ops.pack_xyzw(C.to_device(dst))
    for cell in (0.05, 1.0, 7.0, 60.0, 5000.0):
        idx, dist = ops.nn_batch(s, ops._offsets([3000]), d, ops._offsets([3000]), 1, 3000, 3000, cell_size=cell)
        assert np.array_equal(idx.cpu().numpy(), ib), cell


def test_batched_pairs_ragged():
    ops = sub("ops")
    C = sub("_capi")
    rng = np.random.default_rng(2)
    ns, ms = [100, 0, 2500, 1, 700], [300, 50, 0, 900, 1]
    srcs = [rng.normal(size=(n, 3)) * 20 for n in ns]
    dsts = [rng.normal(size=(m, 3)) * 20 for m in ms]
    s = ops.pack_xyzw(C.to_device(np.concatenate(srcs)))
    d = ops.pack_xyzw(C.to_device(np.concatenate(dsts)))
    idx, dist = ops.nn_batch(s, ops._offsets(ns), d, ops._offsets(ms), 5, max(ns), max(ms))
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    so = np.concatenate(([0], np.cumsum(ns)))
    for p in range(5):
        if ns[p] == 0 or ms[p] == 0:
            continue          # empty pair: outputs unspecified (the reference raises)
        db, ib = O.nearest_neighbor_brute(srcs[p], dsts[p])
        assert np.array_equal(idx[so[p]:so[p + 1]], ib), p
        np.testing.assert_allclose(dist[so[p]:so[p + 1]], db, rtol=1e-14)


def test_fuzz_against_brute_force():
    """Many small random configurations -- scales from 1e-3 to 1e6, anisotropic and degenerate extents, clusters,
    queries inside, outside and far from the destination cloud, tiny and lopsided sizes -- every answer must be
    brute force's, bit for bit (index and squared distance)."""
    icp = sub("utils.icp")
    rng = np.random.default_rng(2024)
    for case in range(120):
        n = int(rng.choice([1, 2, 3, 31, 32, 33, 100, 500, 1500, 3000]))
        m = int(rng.choice([1, 2, 5, 64, 300, 1000, 2500]))
        scale = 10.0 ** rng.uniform(-3, 6)
        aniso = 10.0 ** rng.uniform(-3, 0, size=3) if rng.random() < 0.4 else np.ones(3)
        kind = rng.integers(0, 5)
        if kind == 0:        # gaussian blob
            dst = rng.normal(size=(m, 3))
        elif kind == 1:      # surface z = f(x, y)
            xy = rng.uniform(-1, 1, size=(m, 2))
            dst = np.column_stack((xy, 0.3 * np.sin(3 * xy[:, 0]) + 0.2 * xy[:, 1] ** 2))
        elif kind == 2:      # a few tight clusters far apart
            centres = rng.normal(size=(4, 3)) * 5
            dst = centres[rng.integers(0, 4, m)] + rng.normal(size=(m, 3)) * 0.01
        elif kind == 3:      # lattice: exact ties
            dst = rng.integers(0, 5, size=(m, 3)).astype(float)
        else:                # uniform cube
            dst = rng.uniform(-1, 1, size=(m, 3))
        where = rng.integers(0, 4)
        if where == 0:       # queries drawn like the destination, slightly moved
            src = dst[rng.integers(0, m, n)] + rng.normal(size=(n, 3)) * 0.02
        elif where == 1:     # same region
            src = rng.uniform(-1.2, 1.2, size=(n, 3))
        elif where == 2:     # shifted well outside
            src = rng.normal(size=(n, 3)) + rng.choice([-1, 1], size=3) * 20
        else:                # a mixture with a few enormous outliers
            src = rng.uniform(-1, 1, size=(n, 3))
            src[rng.random(n) < 0.05] *= 1e4
        if kind == 3:
            src = np.round(src * 2) / 2
        dst = dst * aniso * scale
        src = src * aniso * scale
        d, i = icp.nearest_neighbor(src, dst)
        db, ib = O.nearest_neighbor_brute(src, dst)
        assert np.array_equal(i, ib), (case, n, m, kind, where, scale)
        np.testing.assert_allclose(d, db, rtol=1e-14, atol=0, err_msg=str(case))
