"""Where the time of one register_imgs call goes (sofa frames of tests/golden): python tools/time_register_imgs.py"""
import importlib
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

t3d = importlib.import_module("2dto3dreconstruction_b200.utils.transformation3d")
utils = importlib.import_module("2dto3dreconstruction_b200.utils.utils")
icp = importlib.import_module("2dto3dreconstruction_b200.utils.icp")
q9 = importlib.import_module("2dto3dreconstruction_b200.utils.homography_utils.q9")
f = np.load(os.path.join(ROOT, "tests", "golden", "sofa_frames.npz"))
g = np.load(os.path.join(ROOT, "tests", "golden", "sofa_pair.npz"))
gray, depth = f["gray"], f["depth"]


def timed(label, fn, reps=3):
    best = 1e9
    out = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("%-34s %8.2f ms" % (label, best * 1e3), flush=True)
    return out


warnings.simplefilter("ignore")
clouds = timed("depth_to_voxel_ld x2", lambda: [utils.depth_to_voxel_ld(d) for d in depth])
import cv2  # noqa: E402
sift = cv2.SIFT_create()
kd = timed("cv2 SIFT x2 (CPU, out of scope)", lambda: [sift.detectAndCompute(im, None) for im in (gray[1], gray[0])])
m = timed("match_descriptors (K5)", lambda: t3d.match_descriptors(kd[0][1], kd[1][1]))
matches = np.array([(x.queryIdx, x.trainIdx) for x in m])


def ransac():
    np.random.seed(0)
    return q9.ransac_loop(gray[1], gray[0], kd[0][0], kd[1][0], matches)


timed("ransac_loop (host DLT x100 + K4A)", ransac)
A, B = g["icp_A"], g["icp_B"]
timed("icp 13k x 13k, 100 it, partial 0.7", lambda: icp.icp(A, B, tolerance=0.001, max_iterations=100, partial_set_size=0.7))
timed("icp 13k x 13k, 100 it, partial 1.0", lambda: icp.icp(A, B, tolerance=0.0, max_iterations=100, partial_set_size=1.0))


def whole():
    np.random.seed(0)
    return t3d.register_imgs(gray[1], gray[0], depth[1], depth[0], img1_pts=clouds[1], img2_pts=clouds[0], filter_pts_frac=0.05,
                             partial_set_frac=0.7)


timed("register_imgs (whole)", whole)
