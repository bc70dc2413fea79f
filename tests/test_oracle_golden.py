"""The oracle (oracle/ref_port.py) against vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import contextlib
import io

import numpy as np
import pytest

from conftest import load_golden
from oracle import ref_port as O


def test_backproject_bit_exact():
    g = load_golden("backproject")
    small = g["small"]
    for got, key in [
        (O.depth_to_voxel_ld(small), "small_ld"),
        (O.depth_to_voxel_ld(small, 0.5), "small_ld_half"),
        (O.depth_to_voxel(small), "small_px"),
        (O.depth_to_voxel(small, 0.1), "small_px_tenth"),
        (O.depth_to_voxel(small, 1.), "small_px_t3d"),
        (O.depth_to_voxel(g["fimg"]), "fimg_px"),
        (O.depth_to_voxel(g["fimg"], 2.5), "fimg_px_scaled"),
        (O.depth_to_voxel_ld(g["crop"]), "crop_ld"),
    ]:
        want = g[key]
        assert got.dtype == want.dtype, key
        assert got.shape == want.shape, key
        assert np.array_equal(got, want), key
    # int16 wrap-around rows are kept with negative z
    assert (O.depth_to_voxel_ld(small)[:, 2] < 0).sum() == 3


def test_nearest_neighbor_kdtree_and_brute():
    g = load_golden("nearest_neighbor")
    d, i = O.nearest_neighbor(g["src"], g["dst"])
    assert np.array_equal(i, g["idx"]) and np.array_equal(d, g["dist"])
    db, ib = O.nearest_neighbor_brute(g["src"], g["dst"])
    assert np.array_equal(ib, g["idx"])
    np.testing.assert_allclose(db, g["dist"], rtol=1e-14)
    # lattice data: indices may differ on exact ties, distances may not
    dl, il = O.nearest_neighbor_brute(g["lat_src"], g["lat_dst"])
    np.testing.assert_allclose(dl, g["lat_dist"], rtol=1e-14, atol=0)
    tie = il != g["lat_idx"]
    assert tie.sum() > 0, "fixture is meant to contain exact ties"
    d_ref_choice = np.linalg.norm(g["lat_src"][tie] - g["lat_dst"][g["lat_idx"][tie]], axis=1)
    np.testing.assert_allclose(d_ref_choice, dl[tie], rtol=1e-14)
    assert (il[tie] < g["lat_idx"][tie]).all(), "brute force breaks ties towards the lowest index"


def test_kabsch():
    g = load_golden("kabsch")
    assert np.allclose(O.best_fit_transform(g["A"], g["B"]), g["T"], rtol=0, atol=1e-12)
    assert np.allclose(O.best_fit_transform(g["A"], g["B_reflect"]), g["T_reflect"], rtol=0, atol=1e-12)
    for a, b, rk, tk in [("A", "B", "R1", "t1"), ("A", "B_reflect", "R2", "t2"), ("Ai", "Bi", "R3", "t3")]:
        R, t = O.rigid_transform_3D(g[a].T, g[b].T)
        assert t.shape == (3, 1)
        assert np.allclose(R, g[rk], atol=1e-12) and np.allclose(t, g[tk], atol=1e-9)
    with pytest.raises(Exception):
        O.rigid_transform_3D(g["A"], g["B"])          # (K,3): not 3xN
    with pytest.raises(AssertionError):
        O.rigid_transform_3D(g["A"].T, g["B"][:2].T[:2])


@pytest.mark.parametrize("tag,kw", [
    ("full30", dict(max_iterations=30, tolerance=0.0, partial_set_size=1.0)),
    ("part30", dict(max_iterations=30, tolerance=0.0, partial_set_size=0.7)),
    ("tol", dict(max_iterations=50, tolerance=0.001, partial_set_size=0.7)),
    ("default", dict()),
])
def test_icp(tag, kw):
    g = load_golden("icp")
    ca, cb = O.depth_to_voxel_ld(_embed(g["frame_a"])), O.depth_to_voxel_ld(_embed(g["frame_b"]))
    assert np.array_equal(ca, g["cloud_a"]) and np.array_equal(cb, g["cloud_b"])
    T, dist, it = O.icp(ca, cb, **kw)
    assert it == int(g[tag + "_iters"])
    np.testing.assert_allclose(T, g[tag + "_T"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(np.sort(dist), g[tag + "_dist_sorted"], rtol=1e-10, atol=1e-10)


def test_icp_init_pose():
    g = load_golden("icp")
    for tag, init, n in [("init", g["init"], 10), ("aff", g["init_aff"], 6)]:
        T, dist, it = O.icp(g["cloud_a"], g["cloud_b"], init_pose=init, max_iterations=n, tolerance=0.0)
        assert it == int(g[tag + "_iters"])
        np.testing.assert_allclose(T, g[tag + "_T"], rtol=0, atol=1e-10)


def _embed(small):
    f = np.zeros((480, 640), np.uint16)
    f[200:260, 280:360] = small
    return f


def test_synth_pair_is_deterministic():
    g = load_golden("icp")
    a, b = O.synth_depth_pair(0, h=60, w=80)
    assert np.array_equal(a, g["frame_a"]) and np.array_equal(b, g["frame_b"])


def test_affine_fit():
    g = load_golden("affine")
    h = O.homo_rigid_transform_3d(g["p"], g["q"])
    np.testing.assert_allclose(h, g["h"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(O.homo_rigid_transform_3d_scipy(g["p"], g["q"]), g["h"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(O.get_transformed_points(g["cloud"], g["h"]), g["cloud_h"])


def test_ransac2d_pieces():
    g = load_golden("ransac2d")
    Hh, Ww = g["img_shape"]
    orig = g["xy1"][:, ::-1]
    prime = g["xy2"][:, ::-1]
    binned = O.make_binned_array((Hh, Ww), prime)
    # CSR form must list the same (cell, item) pairs in the same order as the reference's
    # list-of-lists walked row-major
    cells = np.repeat(np.arange(Hh * Ww), np.diff(binned["start"]))
    assert np.array_equal(cells, g["binned_cells"]) and np.array_equal(binned["items"], g["binned_items"])
    assert (g["binned_cells"] // Ww == Hh - 1).any(), "fixture exercises the negative-index wrap"
    for k in range(3):
        tp = O.q9_get_transformed_points(orig, g["Hs"][k])
        assert np.array_equal(tp, g["tp%d" % k])
        m, s = O.get_ssd_v2(orig, tp, prime, binned)
        assert np.array_equal(m, g["m%d" % k])
        assert np.array_equal(s, g["s%d" % k])
    assert len(g["m0"]) > 100


def test_ransac2d_loop_seeded():
    g = load_golden("ransac2d")
    np.random.seed(7)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Hs, picks = O.ransac_hypotheses(g["xy1"], g["xy2"], g["bf"])
    assert np.array_equal(picks, g["picks"])
    assert np.array_equal(Hs, g["dlt_H"])
    np.random.seed(7)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        H, fm, counts = O.ransac_loop_pts(tuple(g["img_shape"]), g["xy1"], g["xy2"], g["bf"])
    assert np.array_equal(fm, g["final_matches"])
    np.testing.assert_allclose(H, g["H_final"], rtol=1e-9, atol=1e-12)
    assert counts.max() > 100


def test_chain_prefix():
    rng = np.random.default_rng(0)
    Ts = [np.eye(4) + rng.normal(size=(4, 4)) * 0.01 for _ in range(5)]
    P = O.chain_prefix(Ts)
    assert np.allclose(P[3], Ts[3] @ Ts[2] @ Ts[1] @ Ts[0])


def test_sofa_pair_reference_run():
    """The reference's own register_imgs on examples/sofa_rgbd frames 1-2 (tests/golden/make_golden_sofa.py):
    the oracle must reproduce its RANSAC stage (same seeded draws) and its ICP stage."""
    import warnings
    g = load_golden("sofa_pair")
    np.random.seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        H, fm, counts = O.ransac_loop_pts(tuple(g["img_shape"]), g["kp1"], g["kp2"], g["bf"])
    assert np.array_equal(fm, g["matches"])
    np.testing.assert_allclose(H, g["H"], rtol=1e-9, atol=1e-12)
    tol, max_it, partial = g["icp_kw"]
    T, dist, it = O.icp(g["icp_A"], g["icp_B"], tolerance=float(tol), max_iterations=int(max_it), partial_set_size=float(partial))
    assert it == int(g["icp_iters"])
    np.testing.assert_allclose(T, g["icp_T"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(np.sort(dist), g["icp_dist_sorted"], rtol=0, atol=1e-7)


def test_match_descriptors_vs_cv2():
    """K5 oracle: the recorded cv2.BFMatcher(crossCheck=True) output of the reference run on the sofa frames, and
    OpenCV itself (a third-party dependency of the reference, present in this image) on random sets with ties."""
    g = load_golden("sofa_pair")
    q, t, d = O.match_descriptors(g["des1"].astype(np.float32), g["des2"].astype(np.float32))
    assert np.array_equal(np.stack((q, t), axis=1), g["bf"])
    assert np.array_equal(d, g["bf_dist"])
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(4)
    for trial in range(4):
        n1, n2 = rng.integers(20, 300, 2)
        d1 = rng.integers(0, 256, (n1, 128)).astype(np.float32)
        d2 = rng.integers(0, 256, (n2, 128)).astype(np.float32)
        if trial >= 2:      # duplicates and low-entropy descriptors: exact ties in both directions
            d2[: n2 // 3] = d1[rng.integers(0, n1, n2 // 3)]
            d1[:, 6:] = 0
            d2[:, 6:] = 0
        m = cv2.BFMatcher(normType=cv2.NORM_L2, crossCheck=True).match(d1, d2)
        q, t, d = O.match_descriptors(d1, d2)
        assert np.array_equal(np.stack((q, t), axis=1), np.array([(x.queryIdx, x.trainIdx) for x in m]).reshape(-1, 2))
        assert np.array_equal(d, np.array([x.distance for x in m], dtype=np.float32))


def test_voxel_down_sample_hand_example():
    """Open3D's published algorithm on a case small enough to do by hand: min = (0,0,0), voxel 2 -> min_bound = -1;
    voxel(p) = floor((p + 1) / 2)."""
    pts = np.array([[0.0, 0.0, 0.0], [0.9, 0.5, 0.2], [1.0, 0.0, 0.0], [2.5, 0.0, 0.0], [0.0, 0.0, 3.0], [0.5, 0.5, 0.5]])
    out = O.voxel_down_sample(pts, 2.0)
    # voxels: (0,0,0) <- points 0, 1, 5; (1,0,0) <- points 2, 3; (0,0,2) <- point 4; output in (z, y, x) order
    want = np.array([[(0.0 + 0.9 + 0.5) / 3, (0.0 + 0.5 + 0.5) / 3, (0.0 + 0.2 + 0.5) / 3], [(1.0 + 2.5) / 2, 0.0, 0.0], [0.0, 0.0, 3.0]])
    assert np.array_equal(out, want)


def test_skimage_match_restatement_hand_example():
    """q8.py:90 semantics on a 3x3 case: mutual nearest neighbours, then best / second-best < max_ratio."""
    d1 = np.array([[0.0, 0.0], [10.0, 0.0], [0.0, 10.0]], dtype=np.float32)
    d2 = np.array([[1.0, 0.0], [10.0, 1.0], [5.0, 7.0]], dtype=np.float32)
    # distances: row0 = [1, 10.05, 8.60]; row1 = [9, 1, 8.60]; row2 = [10.05, 13.45, 5.83]
    assert np.array_equal(O.match_descriptors_skimage(d1, d2, max_ratio=1.0, cross_check=True), [[0, 0], [1, 1], [2, 2]])
    # ratios: 1/8.60 = 0.116, 1/8.60 = 0.116, 5.83/10.05 = 0.580
    assert np.array_equal(O.match_descriptors_skimage(d1, d2, max_ratio=0.5, cross_check=True), [[0, 0], [1, 1]])
    assert np.array_equal(O.match_descriptors_skimage(d1, d2, max_ratio=0.8, cross_check=True), [[0, 0], [1, 1], [2, 2]])


def test_kitchen_pair_reference_run():
    """BASELINE.json configs[1] (`--mode rigid3d`), one pair: the loop body of the reference's rigid3d_proc
    (reconstruct.py:264-278) run by tests/golden/make_golden_kitchen.py on examples/kitchen_rgbd frames 1-2.  The oracle
    must reproduce the matching stage, the seeded RANSAC stage and the rigid fit on the matched 3-D keypoints."""
    import warnings
    g = load_golden("kitchen_pair")
    bf = O.match_descriptors_skimage(g["des1"].astype(np.float32), g["des2"].astype(np.float32), max_ratio=0.8, cross_check=True)
    assert np.array_equal(bf, g["bf"])
    np.random.seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        H, fm, counts = O.ransac_loop_pts(tuple(g["img_shape"]), g["kp1"], g["kp2"], g["bf"])
    assert np.array_equal(fm, g["matches"])
    np.testing.assert_allclose(H, g["H"], rtol=1e-9, atol=1e-12)
    R, t = O.rigid_transform_3D(g["kps3d_1"][fm[:, 0]].T, g["kps3d_2"][fm[:, 1]].T)
    np.testing.assert_allclose(R, g["R"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(t, g["t"], rtol=0, atol=1e-9)


def _digest(a):
    import hashlib
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], dtype=np.uint64)[0]


def test_k4b_golden_pins_the_3d_hypothesis_scoring():
    """BASELINE.json configs[4] at 1/10 size: tests/golden/make_golden_k4b.py ran the composition SURVEY.md section 8(c)
    defines with the reference's own rigid_transform_3D / homo_rigid_transform_3d / get_transformed_points and the < 10
    threshold of q9.py:122.  The oracle must give the same hypotheses and the same inlier counts."""
    g = load_golden("k4b")
    P, Q = g["P"].astype(np.float64), g["Q"]
    with contextlib.redirect_stdout(io.StringIO()):
        Tr = O.rigid_hypotheses(P, Q, g["samples3"][:300])
        Ta = O.affine_hypotheses(P, Q, g["samples4"][:300])
    np.testing.assert_allclose(Tr, g["T_rigid"][:300], rtol=0, atol=1e-9)
    ok = np.isfinite(g["T_affine"][:300]).all(axis=(1, 2))
    scale = np.maximum(np.abs(g["T_affine"][:300][ok]).max(axis=(1, 2), keepdims=True), 1.0)
    assert (np.abs(Ta[ok] - g["T_affine"][:300][ok]) / scale).max() < 1e-9
    assert int(O.score_hypotheses_3d(g["T_star"][None], P, Q)[0]) == int(g["count_star"]) == 2940
    for T, want in ((g["T_rigid"], g["counts_rigid"]), (g["T_affine"], g["counts_affine"])):
        fin = np.isfinite(T).all(axis=(1, 2))
        assert fin.sum() > 1990
        assert np.array_equal(O.score_hypotheses_3d(T[fin], P, Q), want[fin])
        assert want[fin].max() > 2000 and (want[fin] > 100).sum() >= 15


def test_kitchen_config1_both_pairs():
    """BASELINE.json configs[1] IN FULL (tests/golden/make_golden_configs.py): SIFT keypoints, the loop body of
    rigid3d_proc (reconstruct.py:260-283) for both pairs of examples/kitchen_rgbd, one np.random.seed(0) before the loop."""
    import warnings
    cv2 = pytest.importorskip("cv2")
    g = load_golden("kitchen_config1")
    gray, depth = g["gray"], g["depth"]
    sift = cv2.xfeatures2d.SIFT_create() if hasattr(cv2, "xfeatures2d") else cv2.SIFT_create()
    kps, des = zip(*[sift.detectAndCompute(img, None) for img in gray])
    assert [len(k) for k in kps] == list(g["n_kp"])
    np_kps = [np.array([(k.pt[0], k.pt[1]) for k in kp], dtype=int) for kp in kps]            # reconstruct.py:31
    kps_3d = [np.array([(p[0], p[1], depth[i][p[1], p[0]]) for p in np_kps[i]]) for i in range(len(depth))]   # :52-59
    np.random.seed(0)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        for i in range(1, len(gray)):
            bf = O.match_descriptors_skimage(des[i], des[i - 1], max_ratio=0.8, cross_check=True)
            assert len(bf) == g["n_bf"][i - 1]
            H, fm = O.ransac_loop(gray[i], gray[i - 1], kps[i], kps[i - 1], bf)
            assert len(fm) == g["n_inliers"][i - 1] and _digest(np.array(fm, dtype=np.int64)) == g["match_hash"][i - 1]
            R, t = O.rigid_transform_3D(kps_3d[i][fm[:, 0]].T, kps_3d[i - 1][fm[:, 1]].T)
            np.testing.assert_allclose(R, g["transforms"][i - 1][:3, :3], rtol=0, atol=1e-12)
            np.testing.assert_allclose(t[:, 0], g["transforms"][i - 1][:3, 3], rtol=0, atol=1e-9)
    assert _digest(np.random.get_state()[1]) == g["rng_end"][0]


def test_sofa_config0_all_seven_pairs():
    """BASELINE.json configs[0] IN FULL (tests/golden/make_golden_configs.py): the pair loop of trans3d_proc
    (reconstruct.py:319-329) on the 8 frames of examples/sofa_rgbd, filter_pts_frac 0.1, partial_set_frac 0.7, the global RNG
    flowing from pair to pair.  The oracle reproduces the global stage of all 7 pairs (inlier sets by hash, the RNG end
    state) and -- to keep the CPU suite short -- the ICP stage of the two pairs the reference converged fastest on."""
    import warnings
    pytest.importorskip("cv2")
    g = load_golden("sofa_config0")
    gray, depth = g["gray"], g["depth"]
    clouds = [O.depth_to_voxel_ld(d) for d in depth]
    assert [len(c) for c in clouds] == list(g["n_cloud"])
    icp_pairs = set(np.argsort(g["icp_iters"])[:2])
    np.random.seed(0)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        for i in range(1, len(clouds)):
            rec = {}
            out = O.register_imgs(gray[i], gray[i - 1], depth[i], depth[i - 1], clouds[i], clouds[i - 1], filter_pts_frac=0.1,
                                  partial_set_frac=0.7, run_icp=(i - 1) in icp_pairs, record=rec)
            assert len(rec["matches"]) == g["n_inliers"][i - 1]
            assert _digest(np.array(rec["matches"], dtype=np.int64)) == g["match_hash"][i - 1]
            assert rec["n_icp_src"] == g["n_icp_src"][i - 1]
            if (i - 1) in icp_pairs:
                assert rec["icp_iters"] == g["icp_iters"][i - 1]
                np.testing.assert_allclose(out, g["final_h"][i - 1], rtol=0, atol=1e-8)
    assert _digest(np.random.get_state()[1]) == g["rng_end"][0]
