"""Generate tests/golden/kitchen_pair.npz: the loop body of the UNMODIFIED reference's rigid3d_proc (reconstruct.py:264-282)
on the first pair of examples/kitchen_rgbd -- BASELINE.json configs[1] (`reconstruct_rgbd.py --mode rigid3d`) for one pair:

    q8.mathching_skimage -> q9.ransac_loop -> r3d.rigid_transform_3D on the matched 3-D keypoints

    python tests/golden/make_golden_kitchen.py       (build container only: needs /root/reference)

scikit-image is not installed: skimage.feature.match_descriptors, which q8.py:90 calls, is replaced by the restatement of
its published algorithm in oracle/ref_port.py ("any deterministic stand-in used by both sides").
Everything else is the reference's own code, np.random.seed(0) first.  Images are read as utils/rgbd.py:25-31 reads them
(greyscale; depth * scale with scale = 1, reconstruct_rgbd.py:41-45).
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden import REF, install_shim, save  # noqa: E402


def main():
    install_shim()
    from oracle import ref_port as O
    sys.modules["skimage.feature"].match_descriptors = O.match_descriptors_skimage
    import cv2
    from PIL import Image
    import utils.homography_utils.q8 as q8
    import utils.homography_utils.q9 as q9
    import utils.r3d as r3d

    names = [("kitchen1.png", "kitchen1.png"), ("kitchen2.png", "kitchen2.png")]
    gray = [cv2.imread(os.path.join(REF, "examples/kitchen_rgbd/imgs", a), cv2.IMREAD_GRAYSCALE) for a, _ in names]
    depth = [np.asarray(Image.open(os.path.join(REF, "examples/kitchen_rgbd/depth", b))) * 1 for _, b in names]

    # reconstruct.py:17-38 get_kps_decs
    sift = cv2.xfeatures2d.SIFT_create()
    np_kps, cv_kps, cv_des = [], [], []
    for img in gray:
        kp, des = sift.detectAndCompute(img, None)
        np_kps.append(np.array([(kp[idx].pt[0], kp[idx].pt[1]) for idx in range(len(kp))], dtype=int))
        cv_kps.append(kp)
        cv_des.append(des)
    assert all(np.array_equal(d, np.rint(d)) and d.min() >= 0 and d.max() <= 255 for d in cv_des)
    # reconstruct.py:52-59 make_3d_kps_depth_img
    kps_3d = [[(p[0], p[1], depth[i][p[1], p[0]]) for p in np_kps[i]] for i in range(2)]

    # reconstruct.py:264-278, i = 1
    i = 1
    img1, kp1, des1 = gray[i], cv_kps[i], cv_des[i]
    img2, kp2, des2 = gray[i - 1], cv_kps[i - 1], cv_des[i - 1]
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        bf_matches = q8.mathching_skimage(img1, kp1, des1, img2, kp2, des2, False)
        H_matrix, matchs = q9.ransac_loop(img1, img2, kp1, kp2, bf_matches)
        m1 = [kps_3d[i][m[0]] for m in matchs]
        m2 = [kps_3d[i - 1][m[1]] for m in matchs]
        R, t = r3d.rigid_transform_3D(np.array(m1).T, np.array(m2).T)
    save("kitchen_pair", des1=des1.astype(np.uint8), des2=des2.astype(np.uint8),
         kp1=np.array([k.pt for k in kp1]), kp2=np.array([k.pt for k in kp2]), img_shape=np.array(img2.shape[:2]),
         bf=np.array(bf_matches), H=np.array(H_matrix), matches=np.array(matchs),
         kps3d_1=np.array(kps_3d[i], dtype=np.int64), kps3d_2=np.array(kps_3d[i - 1], dtype=np.int64), R=np.array(R), t=np.array(t))
    print("keypoints", len(kp1), len(kp2), "bf", len(bf_matches), "ransac matches", len(matchs))
    print(R, t.ravel())


if __name__ == "__main__":
    main()
