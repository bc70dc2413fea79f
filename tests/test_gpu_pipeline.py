"""Fused batched registration (K1 -> grid -> ICP) vs the per-pair drop-ins and the oracle."""
import numpy as np
import pytest

from conftest import sub
from oracle import ref_port as O
from parity import assert_transform_parity

pytestmark = pytest.mark.gpu


def _pairs(n, h=480, w=640, first=20):
    fa, fb = [], []
    for k in range(n):
        a, b = O.synth_depth_pair(first + k, h=h, w=w)
        fa.append(a)
        fb.append(b)
    return np.stack(fa), np.stack(fb)


def test_device_and_host_paths_agree_with_drop_ins():
    import torch
    pipeline = sub("pipeline")
    icp = sub("utils.icp")
    fa, fb = _pairs(5, h=120, w=160)
    big_a = np.zeros((5, 480, 640), np.uint16)
    big_b = np.zeros((5, 480, 640), np.uint16)
    big_a[:, 180:300, 240:400] = fa
    big_b[:, 180:300, 240:400] = fb
    big_a[3] = 0                                  # an empty source frame in the middle of the batch
    kw = dict(max_iterations=8, tolerance=0.0, partial_set_size=1.0, pairs_per_chunk=2, n_streams=2)
    dev = pipeline.register_depth_pairs(torch.from_numpy(big_a).cuda(), torch.from_numpy(big_b).cuda(), **kw)
    host = pipeline.register_depth_pairs(big_a, big_b, **kw)
    Td = dev["T"].cpu().numpy()
    assert np.array_equal(Td, host["T"])
    assert np.array_equal(dev["iters"].cpu().numpy(), host["iters"])
    assert np.array_equal(dev["n_src"].cpu().numpy(), host["n_src"])
    for p in (0, 1, 2, 4):
        A, B = O.depth_to_voxel_ld(big_a[p]), O.depth_to_voxel_ld(big_b[p])
        assert host["n_src"][p] == len(A) and host["n_dst"][p] == len(B)
        T, dist, it = icp.icp(A, B, max_iterations=8, tolerance=0.0)
        assert np.array_equal(T, Td[p]), "batched and single-pair paths must be bitwise identical"
        Tr, _, itr = O.icp(A, B, max_iterations=8, tolerance=0.0)
        assert it == itr == int(host["iters"][p])
        # vs the KD-tree oracle: exact distance ties on quantised depth resolve differently, so the
        # north_star tolerance applies (1e-4), not bitwise equality
        assert_transform_parity(Td[p], Tr, O.scene_extent(B), "pair %d" % p)
    assert host["n_src"][3] == 0


def test_full_size_pairs_partial():
    pipeline = sub("pipeline")
    fa, fb = _pairs(3)
    out = pipeline.register_depth_pairs(fa, fb, max_iterations=4, tolerance=0.0, partial_set_size=0.7, pairs_per_chunk=3,
                                        n_streams=1)
    for p in range(3):
        A, B = O.depth_to_voxel_ld(fa[p]), O.depth_to_voxel_ld(fb[p])
        Tr, _, itr = O.icp(A, B, max_iterations=4, tolerance=0.0, partial_set_size=0.7)
        assert int(out["iters"][p]) == itr
        assert_transform_parity(out["T"][p], Tr, O.scene_extent(B), "pair %d" % p)


def test_chunking_does_not_change_results():
    pipeline = sub("pipeline")
    fa, fb = _pairs(7, h=96, w=128)
    big_a = np.zeros((7, 480, 640), np.uint16)
    big_b = np.zeros((7, 480, 640), np.uint16)
    big_a[:, 192:288, 256:384] = fa
    big_b[:, 192:288, 256:384] = fb
    ref = None
    for ppc, ns in [(1, 1), (3, 2), (7, 1), (16, 3)]:
        out = pipeline.register_depth_pairs(big_a, big_b, max_iterations=6, tolerance=1e-3, pairs_per_chunk=ppc, n_streams=ns)
        if ref is None:
            ref = out
        else:
            assert np.array_equal(out["T"], ref["T"]) and np.array_equal(out["iters"], ref["iters"])


def test_chain_matches_reference_prefix_product():
    pipeline = sub("pipeline")
    rec = sub("reconstruct")
    rng = np.random.default_rng(0)
    Ts = np.stack([np.eye(4) + rng.normal(size=(4, 4)) * 0.01 for _ in range(6)])
    want = O.chain_prefix(list(Ts))
    assert np.array_equal(pipeline.chain(Ts), np.stack(want))
    assert all(np.array_equal(a, b) for a, b in zip(rec.chain_prefix(list(Ts)), want))


def test_sofa_pair_against_the_reference_run():
    """BASELINE.json configs[0], one pair: the CUDA drop-ins against the recorded run of the unmodified reference on
    examples/sofa_rgbd frames 1-2 (real SIFT keypoints, real depth): RANSAC inlier set identical, homography equal,
    ICP (partial_set_size 0.7, 100 iterations, 13 k points) within north_star's tolerance."""
    import warnings
    from conftest import load_golden
    from parity import assert_transform_parity
    q9 = sub("utils.homography_utils.q9")
    icp = sub("utils.icp")
    g = load_golden("sofa_pair")

    class KP:
        def __init__(self, p):
            self.pt = (float(p[0]), float(p[1]))

    img = np.zeros(tuple(g["img_shape"]), np.uint8)
    np.random.seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        H, fm = q9.ransac_loop(img, img, [KP(p) for p in g["kp1"]], [KP(p) for p in g["kp2"]], g["bf"])
    assert np.array_equal(fm, g["matches"]), "RANSAC inlier set differs from the reference's"
    np.testing.assert_allclose(H, g["H"], rtol=1e-9, atol=1e-12)

    tol, max_it, partial = g["icp_kw"]
    T, dist, it = icp.icp(g["icp_A"], g["icp_B"], tolerance=float(tol), max_iterations=int(max_it), partial_set_size=float(partial))
    assert it == int(g["icp_iters"])
    assert dist.shape == g["icp_dist_sorted"].shape
    assert_transform_parity(T, g["icp_T"], O.scene_extent(g["icp_B"]), "sofa pair")
    np.testing.assert_allclose(np.sort(dist), g["icp_dist_sorted"], rtol=0, atol=1e-4 * O.scene_extent(g["icp_B"]))


def test_register_imgs_end_to_end_on_the_sofa_frames():
    """The whole drop-in chain on real frames -- cv2 SIFT, K5 matching, K4A RANSAC, K3c affine fit, transform, ICP
    (K2 + K3) and the reference's h.t composition -- against the 4x4 the unmodified reference's register_imgs
    returned for the same call (tests/golden/make_golden_sofa.py), same global-RNG seed."""
    import time
    import warnings
    pytest.importorskip("cv2")
    from conftest import load_golden
    t3d = sub("utils.transformation3d")
    utils = sub("utils.utils")
    g = load_golden("sofa_pair")
    f = load_golden("sofa_frames")
    gray, depth = f["gray"], f["depth"]
    clouds = [utils.depth_to_voxel_ld(d) for d in depth]
    assert [c.shape[0] for c in clouds] == list(g["n_cloud"])
    np.random.seed(0)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        p1, p2, h = t3d.register_imgs(gray[1], gray[0], depth[1], depth[0], img1_pts=clouds[1], img2_pts=clouds[0],
                                      filter_pts_frac=0.05, partial_set_frac=0.7)
    print("register_imgs (sofa frames 2 -> 1): %.3f s" % (time.perf_counter() - t0))
    assert p1 is clouds[1] and p2 is clouds[0]
    assert_transform_parity(h, g["final_h"], O.scene_extent(clouds[0]), "register_imgs")
    # the RNG was consumed exactly as the reference consumes it (100 x randint(4) + two choice() calls)
    after = np.random.randint(1 << 30)
    np.random.seed(0)
    for _ in range(100):
        np.random.randint(len(g["bf"]), size=4)
    np.random.choice(clouds[1].shape[0], int(clouds[1].shape[0] * 0.05), replace=False)
    np.random.choice(clouds[0].shape[0], int(clouds[0].shape[0] * 0.05), replace=False)
    assert after == np.random.randint(1 << 30)


def test_kitchen_pair_against_the_reference_run():
    """BASELINE.json configs[1] (`reconstruct_rgbd.py --mode rigid3d`), one pair: the CUDA drop-ins against the recorded
    run of the reference's rigid3d_proc loop body (reconstruct.py:264-278) on examples/kitchen_rgbd frames 1-2:
    q8.mathching_skimage (K5 + ratio test) -> q9.ransac_loop (K4A) -> r3d.rigid_transform_3D (K3).  Match lists and
    the RANSAC inlier set are identical, the homography and the rigid transform agree to rounding."""
    import contextlib
    import io
    import warnings
    from conftest import load_golden
    q8 = sub("utils.homography_utils.q8")
    q9 = sub("utils.homography_utils.q9")
    r3d = sub("utils.r3d")
    g = load_golden("kitchen_pair")

    class KP:
        def __init__(self, p):
            self.pt = (float(p[0]), float(p[1]))

    kp1, kp2 = [KP(p) for p in g["kp1"]], [KP(p) for p in g["kp2"]]
    img = np.zeros(tuple(g["img_shape"]), np.uint8)
    bf = q8.mathching_skimage(img, kp1, g["des1"].astype(np.float32), img, kp2, g["des2"].astype(np.float32), False)
    assert np.array_equal(bf, g["bf"]), "ratio-test matches differ from the reference run's"
    np.random.seed(0)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        H, fm = q9.ransac_loop(img, img, kp1, kp2, bf)
        assert np.array_equal(fm, g["matches"]), "RANSAC inlier set differs from the reference's"
        np.testing.assert_allclose(H, g["H"], rtol=1e-9, atol=1e-12)
        m1 = [g["kps3d_1"][m[0]] for m in fm]                      # reconstruct.py:271-278
        m2 = [g["kps3d_2"][m[1]] for m in fm]
        R, t = r3d.rigid_transform_3D(np.array(m1).T, np.array(m2).T)
    assert R.shape == (3, 3) and t.shape == (3, 1)
    np.testing.assert_allclose(R, g["R"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(t, g["t"], rtol=0, atol=1e-8)


def test_register_consecutive_pairs_batches_the_icp_stage():
    """The pair loop of trans3d_proc (reconstruct.py:314-329) with the ICP refinements of all pairs in one batched call:
    every 4x4 must equal, bit for bit, what register_imgs returns pair by pair with the same RNG seed, and the global
    RNG must end in the same state.  Frames: sofa 1, 2, 1 (two pairs: 2 -> 1 and 1 -> 2)."""
    import time
    import warnings
    pytest.importorskip("cv2")
    from conftest import load_golden
    t3d = sub("utils.transformation3d")
    utils = sub("utils.utils")
    rec = sub("reconstruct")
    g = load_golden("sofa_pair")
    f = load_golden("sofa_frames")
    order = [0, 1, 0]
    gray = [f["gray"][k] for k in order]
    depth = [f["depth"][k] for k in order]
    clouds = [utils.depth_to_voxel_ld(d) for d in depth]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.random.seed(0)
        t0 = time.perf_counter()
        seq = []
        for i in range(1, 3):
            _, _, h = t3d.register_imgs(gray[i], gray[i - 1], depth[i], depth[i - 1], img1_pts=clouds[i], img2_pts=clouds[i - 1],
                                        filter_pts_frac=0.05, partial_set_frac=0.7)
            seq.append(h)
        t_seq = time.perf_counter() - t0
        state_seq = np.random.randint(1 << 30)
        np.random.seed(0)
        t0 = time.perf_counter()
        bat = rec.register_consecutive_pairs(clouds, gray, depth, filter_pts_frac=0.05, partial_set_frac=0.7)
        t_bat = time.perf_counter() - t0
        state_bat = np.random.randint(1 << 30)
    print("two pairs: sequential %.3f s, batched ICP stage %.3f s" % (t_seq, t_bat))
    assert len(bat) == 2 and state_bat == state_seq
    for a, b in zip(seq, bat):
        assert np.array_equal(a, b)
    assert_transform_parity(bat[0], g["final_h"], O.scene_extent(clouds[0]), "first pair vs the reference run")
    assert rec.register_consecutive_pairs(clouds[:1], gray[:1], depth[:1]) == []


def _digest(a):
    import hashlib
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], dtype=np.uint64)[0]


def test_sofa_config0_all_seven_pairs():
    """BASELINE.json configs[0] IN FULL: `reconstruct_rgbd.py --mode 3dhomo` on examples/sofa_rgbd -- the pair loop of
    trans3d_proc (reconstruct.py:319-329) over all 8 frames with the reference's own settings (filter_pts_frac 0.1,
    partial_set_frac 0.7, one seed before the loop).  tests/golden/make_golden_configs.py recorded the unmodified
    reference's 7 returned 4x4s, RANSAC inlier lists (hashed), ICP iteration counts and the final state of the global
    RNG; the drop-in register_consecutive_pairs must reproduce all of them."""
    import time
    import warnings
    pytest.importorskip("cv2")
    from conftest import load_golden
    t3d = sub("utils.transformation3d")
    ops = sub("ops")
    rec = sub("reconstruct")
    g = load_golden("sofa_config0")
    gray, depth = list(g["gray"]), list(g["depth"])
    clouds = rec.depth_images_to_3d_pts_ld(depth)
    assert [c.shape[0] for c in clouds] == list(g["n_cloud"])
    seen = {"matches": [], "iters": None, "n_src": None}
    real_ransac, real_icp_batch = t3d.ransac_loop, ops.icp_batch

    def ransac_spy(*a, **kw):
        H, m = real_ransac(*a, **kw)
        seen["matches"].append(np.array(m, dtype=np.int64))
        return H, m

    def icp_spy(s, so, *a, **kw):
        r = real_icp_batch(s, so, *a, **kw)
        seen["iters"] = r["iters"].cpu().numpy()
        seen["n_src"] = np.diff(so.cpu().numpy())
        return r

    t3d.ransac_loop, ops.icp_batch = ransac_spy, icp_spy
    try:
        np.random.seed(0)
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            hs = rec.register_consecutive_pairs(clouds, gray, depth, filter_pts_frac=0.1, partial_set_frac=0.7)
        print("configs[0], 7 pairs: %.2f s" % (time.perf_counter() - t0))
    finally:
        t3d.ransac_loop, ops.icp_batch = real_ransac, real_icp_batch
    assert _digest(np.random.get_state()[1]) == g["rng_end"][0], "global RNG not consumed as the reference consumes it"
    assert len(hs) == 7
    assert [len(m) for m in seen["matches"]] == list(g["n_inliers"])
    assert [_digest(m) for m in seen["matches"]] == list(g["match_hash"]), "RANSAC inlier lists differ from the reference's"
    assert list(seen["n_src"]) == list(g["n_icp_src"])
    assert list(seen["iters"]) == list(g["icp_iters"]), (seen["iters"], g["icp_iters"])
    for k in range(7):
        assert_transform_parity(hs[k], g["final_h"][k], O.scene_extent(clouds[k]), "sofa pair %d" % (k + 1))


def test_kitchen_config1_both_pairs():
    """BASELINE.json configs[1] IN FULL: the loop body of rigid3d_proc (reconstruct.py:260-283) on both pairs of
    examples/kitchen_rgbd through the drop-ins q8.mathching_skimage (K5) -> q9.ransac_loop (K4A) -> r3d.rigid_transform_3D
    (K3), against the recorded run of the reference (tests/golden/make_golden_configs.py)."""
    import contextlib
    import io
    import warnings
    cv2 = pytest.importorskip("cv2")
    from conftest import load_golden
    q8 = sub("utils.homography_utils.q8")
    q9 = sub("utils.homography_utils.q9")
    r3d = sub("utils.r3d")
    g = load_golden("kitchen_config1")
    gray, depth = g["gray"], g["depth"]
    sift = cv2.xfeatures2d.SIFT_create() if hasattr(cv2, "xfeatures2d") else cv2.SIFT_create()
    kps, des = zip(*[sift.detectAndCompute(img, None) for img in gray])
    assert [len(k) for k in kps] == list(g["n_kp"])
    np_kps = [np.array([(k.pt[0], k.pt[1]) for k in kp], dtype=int) for kp in kps]
    kps_3d = [[(p[0], p[1], depth[i][p[1], p[0]]) for p in np_kps[i]] for i in range(len(depth))]
    np.random.seed(0)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        for i in range(1, len(gray)):
            bf = q8.mathching_skimage(gray[i], kps[i], des[i], gray[i - 1], kps[i - 1], des[i - 1], False)
            assert len(bf) == g["n_bf"][i - 1]
            H, fm = q9.ransac_loop(gray[i], gray[i - 1], kps[i], kps[i - 1], bf)
            assert len(fm) == g["n_inliers"][i - 1]
            assert _digest(np.array(fm, dtype=np.int64)) == g["match_hash"][i - 1], "RANSAC inlier list differs"
            m1 = [kps_3d[i][m[0]] for m in fm]
            m2 = [kps_3d[i - 1][m[1]] for m in fm]
            R, t = r3d.rigid_transform_3D(np.array(m1).T, np.array(m2).T)
            np.testing.assert_allclose(R, g["transforms"][i - 1][:3, :3], rtol=0, atol=1e-10)
            np.testing.assert_allclose(t[:, 0], g["transforms"][i - 1][:3, 3], rtol=0, atol=1e-8)
    assert _digest(np.random.get_state()[1]) == g["rng_end"][0]
