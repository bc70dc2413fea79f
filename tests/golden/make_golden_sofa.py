"""Generate tests/golden/sofa_pair.npz: the UNMODIFIED reference's register_imgs (utils/transformation3d.py:217-289)
on the first two frames of examples/sofa_rgbd -- BASELINE.json configs[0] for one pair -- with the inputs and outputs of
its two hot stages recorded: q9.ransac_loop (SIFT keypoints, cv2 brute-force matches) and icp.icp (subsampled clouds).

    python tests/golden/make_golden_sofa.py       (build container only: needs /root/reference)

filter_pts_frac is 0.05 here (reconstruct.py:297 uses 0.1) to keep the fixture near half a megabyte; everything else is
the reference's call sequence, np.random.seed(0) first.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, install_shim, save  # noqa: E402


def main():
    install_shim()
    import cv2
    from PIL import Image
    import utils.transformation3d as t3d
    import utils.utils as ref_utils

    # utils/rgbd.py:31: the "RGB" images are read as greyscale
    rgb = [cv2.imread(os.path.join(REF, "examples/sofa_rgbd/File%d.jpg" % k), cv2.IMREAD_GRAYSCALE) for k in (1, 2)]
    depth = [np.asarray(Image.open(os.path.join(REF, "examples/sofa_rgbd/File%d.png" % k))) for k in (1, 2)]
    clouds = [ref_utils.depth_to_voxel_ld(d) for d in depth]

    rec = {}
    real_ransac, real_icp = t3d.ransac_loop, t3d.icp

    def ransac_spy(img1, img2, kp1, kp2, matches, **kw):
        rec["kp1"] = np.array([k.pt for k in kp1])
        rec["kp2"] = np.array([k.pt for k in kp2])
        rec["bf"] = np.array(matches)
        rec["img_shape"] = np.array(img2.shape[:2])
        rec["rng_before_ransac"] = np.random.get_state()[1].copy()
        H, m = real_ransac(img1, img2, kp1, kp2, matches, **kw)
        rec["H"], rec["matches"] = np.array(H), np.array(m)
        return H, m

    def icp_spy(A, B, **kw):
        rec["icp_A"], rec["icp_B"] = np.array(A), np.array(B)
        rec["icp_kw"] = np.array([kw["tolerance"], kw["max_iterations"], kw["partial_set_size"]])
        T, d, it = real_icp(A, B, **kw)
        rec["icp_T"], rec["icp_dist_sorted"], rec["icp_iters"] = np.array(T), np.sort(d), np.array(it)
        return T, d, it

    # the descriptor matching stage (utils/transformation3d.py:27-32): same SIFT call, descriptors recorded
    sift = cv2.xfeatures2d.SIFT_create()
    _, des1 = sift.detectAndCompute(rgb[1], mask=None)
    _, des2 = sift.detectAndCompute(rgb[0], mask=None)
    assert np.array_equal(des1, np.rint(des1)) and des1.min() >= 0 and des1.max() <= 255      # integer-valued float32
    bf = cv2.BFMatcher(normType=cv2.NORM_L2, crossCheck=True).match(des1, des2)
    rec["des1"], rec["des2"] = des1.astype(np.uint8), des2.astype(np.uint8)
    rec["bf_dist"] = np.array([m.distance for m in bf], dtype=np.float32)
    bf_pairs = np.array([(m.queryIdx, m.trainIdx) for m in bf])

    t3d.ransac_loop, t3d.icp = ransac_spy, icp_spy
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        # reconstruct.py:322: register image i onto image i-1
        _, _, h = t3d.register_imgs(rgb[1], rgb[0], depth[1], depth[0], img1_pts=clouds[1], img2_pts=clouds[0],
                                    filter_pts_frac=0.05, partial_set_frac=0.7)
    t3d.ransac_loop, t3d.icp = real_ransac, real_icp
    rec.pop("rng_before_ransac")
    assert np.array_equal(rec["bf"], bf_pairs), "register_imgs matched other descriptors than the recorded ones"
    # the 3-D keypoint triples and the affine fit between the two stages (utils/transformation3d.py:261-275)
    save("sofa_pair", final_h=np.array(h), n_cloud=np.array([c.shape[0] for c in clouds]), **rec)
    # the frames themselves (greyscale as utils/rgbd.py:31 reads them, 16-bit depth), as arrays: inputs of the
    # end-to-end drop-in test of register_imgs
    save("sofa_frames", gray=np.stack(rgb), depth=np.stack(depth))
    print({k: getattr(v, "shape", v) for k, v in rec.items()})
    print("iters", rec["icp_iters"], "matches", rec["matches"].shape)


if __name__ == "__main__":
    main()
