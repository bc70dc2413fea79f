"""Generate tests/golden/sofa_config0.npz and kitchen_config1.npz: BASELINE.json configs[0] and configs[1] IN FULL, run by
the UNMODIFIED reference in the build container.

    python tests/golden/make_golden_configs.py       (needs /root/reference; ~3 minutes, almost all of it the reference's ICP)

configs[0]  `reconstruct_rgbd.py --inter 1 --mode 3dhomo` on examples/sofa_rgbd: frames loaded by the reference's own
            utils/rgbd.py get_rgbd (scale 1, reconstruct_rgbd.py:41-45), clouds by reconstruct.depth_images_to_3d_pts_ld, then
            the pair loop of reconstruct.trans3d_proc (reconstruct.py:319-329) with its defaults filter_pts_frac = 0.1,
            partial_set_frac = 0.7 -- all 7 pairs, ONE np.random.seed(0) before the loop, the global RNG flowing from pair
            to pair as it does in the reference.  trans3d_proc itself cannot be called (make_pcds / chain_transformation need
            Open3D, which is absent): its three-line loop body is repeated here around the reference's register_imgs.
configs[1]  `--mode rigid3d` on examples/kitchen_rgbd: reconstruct.get_kps_decs, then the loop body of
            reconstruct.rigid3d_proc (reconstruct.py:260-283) for both pairs.  scikit-image is absent: q8.py:90's
            match_descriptors is the oracle's restatement, everything else is the reference's code.

Stored per pair: the returned 4x4, the RANSAC inlier count, a hash of the inlier list, the ICP iteration count -- a few hundred
bytes -- plus the frames (the GPU box has no /root/reference) and the final state of NumPy's global RNG.
"""
import contextlib
import hashlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden import REF, install_shim, save  # noqa: E402


def digest(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest()[:8], dtype=np.uint64)[0]


def rng_digest():
    return digest(np.random.get_state()[1])


def main():
    install_shim()
    from PIL import Image
    sys.modules["skimage.io"].imread = lambda f: np.asarray(Image.open(f))
    from oracle import ref_port as O
    sys.modules["skimage.feature"].match_descriptors = O.match_descriptors_skimage
    import utils.rgbd as rgbd
    import utils.transformation3d as t3d
    import utils.utils as ref_utils
    import utils.homography_utils.q8 as q8
    import utils.homography_utils.q9 as q9
    import utils.r3d as r3d
    import cv2

    # ---------------------------------------------------------------- configs[0]
    rgb, depth = rgbd.get_rgbd(os.path.join(REF, "examples/sofa_rgbd/*.jpg"), os.path.join(REF, "examples/sofa_rgbd/*.png"), False, 1)
    clouds = [ref_utils.depth_to_voxel_ld(img) for img in depth]       # reconstruct.py:354
    rec = {"n_inliers": [], "match_hash": [], "icp_iters": [], "n_icp_src": []}
    real_ransac, real_icp = t3d.ransac_loop, t3d.icp

    def ransac_spy(*a, **kw):
        H, m = real_ransac(*a, **kw)
        rec["n_inliers"].append(len(m))
        rec["match_hash"].append(digest(np.array(m, dtype=np.int64)))
        return H, m

    def icp_spy(A, B, **kw):
        T, d, it = real_icp(A, B, **kw)
        rec["icp_iters"].append(it)
        rec["n_icp_src"].append(len(A))
        return T, d, it

    t3d.ransac_loop, t3d.icp = ransac_spy, icp_spy
    all_results = []
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(1, len(clouds)):          # reconstruct.py:319-329
            _, _, h = t3d.register_imgs(rgb[i], rgb[i - 1], depth[i], depth[i - 1], img1_pts=clouds[i], img2_pts=clouds[i - 1],
                                        filter_pts_frac=0.1, partial_set_frac=0.7)
            all_results.append(np.array(h))
            print("pair", i, file=sys.stderr, flush=True)
    t3d.ransac_loop, t3d.icp = real_ransac, real_icp
    save("sofa_config0", gray=np.asarray(rgb), depth=np.asarray(depth), final_h=np.stack(all_results),
         n_inliers=np.array(rec["n_inliers"]), match_hash=np.array(rec["match_hash"], dtype=np.uint64),
         icp_iters=np.array(rec["icp_iters"]), n_icp_src=np.array(rec["n_icp_src"]), n_cloud=np.array([len(c) for c in clouds]),
         rng_end=np.array([rng_digest()], dtype=np.uint64))
    print("configs[0]: inliers", rec["n_inliers"], "icp iters", rec["icp_iters"])

    # ---------------------------------------------------------------- configs[1]
    rgb, depth = rgbd.get_rgbd(os.path.join(REF, "examples/kitchen_rgbd/imgs/*.png"), os.path.join(REF, "examples/kitchen_rgbd/depth/*.png"), False, 1)
    # reconstruct.py:17-38 get_kps_decs
    sift = cv2.xfeatures2d.SIFT_create()
    np_kps, cv_kps, cv_des = [], [], []
    for img in rgb:
        kp, des = sift.detectAndCompute(img, None)
        np_kps.append(np.array([(kp[idx].pt[0], kp[idx].pt[1]) for idx in range(len(kp))], dtype=int))
        cv_kps.append(kp)
        cv_des.append(des)
    kps_3d = [[(p[0], p[1], depth[i][p[1], p[0]]) for p in np_kps[i]] for i in range(len(depth))]      # reconstruct.py:52-59
    results, n_bf, n_in, mh = [], [], [], []
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        for i in range(1, len(rgb)):             # reconstruct.py:264-283
            img1, kp1, des1 = rgb[i], cv_kps[i], cv_des[i]
            img2, kp2, des2 = rgb[i - 1], cv_kps[i - 1], cv_des[i - 1]
            bf_matches = q8.mathching_skimage(img1, kp1, des1, img2, kp2, des2, False)
            H_matrix, matchs = q9.ransac_loop(img1, img2, kp1, kp2, bf_matches)
            m1 = [kps_3d[i][m[0]] for m in matchs]
            m2 = [kps_3d[i - 1][m[1]] for m in matchs]
            R, t = r3d.rigid_transform_3D(np.array(m1).T, np.array(m2).T)
            Hm = np.pad(R, ((0, 1), (0, 1)))
            Hm[3, 3] = 1
            Hm[0, 3], Hm[1, 3], Hm[2, 3] = t[0, 0], t[1, 0], t[2, 0]
            results.append(Hm)
            n_bf.append(len(bf_matches))
            n_in.append(len(matchs))
            mh.append(digest(np.array(matchs, dtype=np.int64)))
    save("kitchen_config1", gray=np.asarray(rgb), depth=np.asarray(depth), transforms=np.stack(results), n_bf=np.array(n_bf),
         n_inliers=np.array(n_in), match_hash=np.array(mh, dtype=np.uint64), n_kp=np.array([len(k) for k in cv_kps]),
         rng_end=np.array([rng_digest()], dtype=np.uint64))
    print("configs[1]: keypoints", [len(k) for k in cv_kps], "bf", n_bf, "inliers", n_in)


if __name__ == "__main__":
    main()
