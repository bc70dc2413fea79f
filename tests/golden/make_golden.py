"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference's hot-path modules import matplotlib / skimage / open3d at module level
for plotting and I/O only; those are stubbed here so the arithmetic runs untouched.
Every fixture stores the exact inputs next to the reference's outputs, so the tests
need nothing but the .npz files.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def install_shim():
    import cv2
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.io", "skimage.feature", "open3d"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    sys.modules["skimage"].feature = sys.modules["skimage.feature"]
    sys.modules["skimage.feature"].match_descriptors = None
    sys.modules["skimage.feature"].plot_matches = None
    if not hasattr(cv2, "xfeatures2d"):
        cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=cv2.SIFT_create)
    sys.path.insert(0, REF)


class KP:
    def __init__(self, x, y):
        self.pt = (float(x), float(y))


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    install_shim()
    import utils.icp as ref_icp
    import utils.r3d as ref_r3d
    import utils.utils as ref_utils
    import utils.transformation3d as ref_t3d
    import utils.homography_utils.q9 as ref_q9
    from oracle.ref_port import synth_depth_pair

    # ---- 1. back-projection -------------------------------------------------------
    rng = np.random.default_rng(101)
    small = rng.integers(0, 9000, size=(37, 53)).astype(np.uint16)
    small[rng.random(small.shape) < 0.25] = 0
    small[3, 4] = 40000      # int16 wrap-around: negative z, not dropped
    small[5, 6] = 32768
    small[7, 8] = 65535
    fimg = (rng.random((24, 32)) * 10).astype(np.float32)   # DenseDepth-style float depth
    fimg[fimg < 1.5] = 0.3                                  # truncates to 0 -> dropped
    from PIL import Image
    real = np.asarray(Image.open(os.path.join(REF, "examples/sofa_rgbd/File1.png")))
    crop = np.ascontiguousarray(real[100:220, 200:360])
    save("backproject",
         small=small,
         small_ld=ref_utils.depth_to_voxel_ld(small),
         small_ld_half=ref_utils.depth_to_voxel_ld(small, 0.5),
         small_px=ref_utils.depth_to_voxel(small),
         small_px_tenth=ref_utils.depth_to_voxel(small, 0.1),
         small_px_t3d=ref_t3d.depth_to_voxel(small, 1.),
         fimg=fimg,
         fimg_px=ref_utils.depth_to_voxel(fimg),
         fimg_px_scaled=ref_utils.depth_to_voxel(fimg, 2.5),
         crop=crop,
         crop_ld=ref_utils.depth_to_voxel_ld(crop),
         real_shape=np.array(real.shape),
         real_count=np.array(ref_utils.depth_to_voxel_ld(real).shape[0]),
         real_sum=ref_utils.depth_to_voxel_ld(real).sum(axis=0))

    # ---- 2. nearest neighbour --------------------------------------------------------
    rng = np.random.default_rng(202)
    src = rng.normal(size=(1500, 3)) * 50
    dst = rng.normal(size=(1400, 3)) * 50
    d, i = ref_icp.nearest_neighbor(src, dst)
    # lattice-like clouds (quantised depth): exact distance ties are real
    lat_dst = rng.integers(0, 12, size=(900, 3)).astype(np.float64) * 4.0
    lat_src = rng.integers(0, 24, size=(800, 3)).astype(np.float64) * 2.0
    dl, il = ref_icp.nearest_neighbor(lat_src, lat_dst)
    save("nearest_neighbor", src=src, dst=dst, dist=d, idx=i,
         lat_src=lat_src, lat_dst=lat_dst, lat_dist=dl, lat_idx=il)

    # ---- 3. Kabsch -------------------------------------------------------------------
    rng = np.random.default_rng(303)
    A = rng.normal(size=(400, 3)) * 30 + np.array([100., -50., 2000.])
    ang = 0.3
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]])
    B = A @ Rz.T + np.array([5., -3., 12.]) + rng.normal(size=A.shape) * 0.5
    B_reflect = A * np.array([1., 1., -1.]) + rng.normal(size=A.shape) * 0.1   # forces det<0 branch
    Ai = rng.integers(0, 640, size=(60, 3))
    Bi = Ai + rng.integers(-3, 4, size=(60, 3))
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        R1, t1 = ref_r3d.rigid_transform_3D(A.T, B.T)
        R2, t2 = ref_r3d.rigid_transform_3D(A.T, B_reflect.T)
        R3, t3 = ref_r3d.rigid_transform_3D(Ai.T, Bi.T)
    save("kabsch", A=A, B=B, B_reflect=B_reflect, Ai=Ai, Bi=Bi,
         T=ref_icp.best_fit_transform(A, B),
         T_reflect=ref_icp.best_fit_transform(A, B_reflect),
         R1=R1, t1=t1, R2=R2, t2=t2, R3=R3, t3=t3)

    # ---- 4. ICP ----------------------------------------------------------------------
    da, db = synth_depth_pair(0, h=60, w=80)
    # embed the small frames in the middle of a 480x640 image so the hard-coded intrinsics
    # give a sensible cloud
    fa = np.zeros((480, 640), np.uint16)
    fb = np.zeros((480, 640), np.uint16)
    fa[200:260, 280:360] = da
    fb[200:260, 280:360] = db
    ca = ref_utils.depth_to_voxel_ld(fa)
    cb = ref_utils.depth_to_voxel_ld(fb)
    out = {}
    for tag, kw in {
        "full30": dict(max_iterations=30, tolerance=0.0, partial_set_size=1.0),
        "part30": dict(max_iterations=30, tolerance=0.0, partial_set_size=0.7),
        "tol": dict(max_iterations=50, tolerance=0.001, partial_set_size=0.7),
        "default": dict(),
    }.items():
        T, dist, it = ref_icp.icp(ca, cb, **kw)
        out[tag + "_T"] = T
        out[tag + "_dist_sorted"] = np.sort(dist)
        out[tag + "_iters"] = np.array(it)
    init = np.eye(4)
    init[:3, 3] = [3., -2., 10.]
    T, dist, it = ref_icp.icp(ca, cb, init_pose=init, max_iterations=10, tolerance=0.0)
    out["init_T"], out["init_dist_sorted"], out["init_iters"] = T, np.sort(dist), np.array(it)
    # a non-rigid init pose: the final best_fit_transform(A, src) is then a projection
    init_aff = init.copy()
    init_aff[:3, :3] = np.diag([1.02, 0.97, 1.0])
    T, dist, it = ref_icp.icp(ca, cb, init_pose=init_aff, max_iterations=6, tolerance=0.0)
    out["aff_T"], out["aff_dist_sorted"], out["aff_iters"] = T, np.sort(dist), np.array(it)
    save("icp", frame_a=da, frame_b=db, init=init, init_aff=init_aff, cloud_a=ca, cloud_b=cb, **out)

    # ---- 5. 12-dof affine fit ----------------------------------------------------------
    rng = np.random.default_rng(505)
    p = np.stack((rng.integers(0, 640, 300), rng.integers(0, 480, 300), rng.integers(500, 8000, 300)), axis=1)
    M = np.array([[0.98, 0.02, 0.001, 4.], [-0.02, 0.99, 0.002, -6.], [0.5, -0.3, 1.01, 20.]])
    q = np.rint(np.concatenate((p, np.ones((300, 1))), axis=1) @ M.T).astype(np.int64)
    h = ref_t3d.homo_rigid_transform_3d(p, q)
    cloud = rng.normal(size=(500, 3)) * 100
    save("affine", p=p, q=q, h=h, cloud=cloud, cloud_h=ref_t3d.get_transformed_points(cloud, h))

    # ---- 6./7. 2-D homography RANSAC ---------------------------------------------------
    rng = np.random.default_rng(606)
    Hh, Ww = 120, 160
    n1 = 320
    xy1 = np.stack((rng.uniform(2, Ww - 3, n1), rng.uniform(2, Hh - 3, n1)), axis=1)   # (x, y)
    Htrue = np.array([[1.01, 0.015, 2.5], [-0.012, 0.99, -1.5], [1e-5, -2e-5, 1.0]])
    ph = np.concatenate((xy1, np.ones((n1, 1))), axis=1) @ Htrue.T
    xy2_all = ph[:, :2] / ph[:, 2:3] + rng.normal(size=(n1, 2)) * 0.4
    inside = (xy2_all[:, 0] > 1) & (xy2_all[:, 0] < Ww - 2) & (xy2_all[:, 1] > 1) & (xy2_all[:, 1] < Hh - 2)
    keep2 = np.where(inside)[0][:260]
    extra = np.stack((rng.uniform(0, Ww - 1, 60), rng.uniform(0, Hh - 1, 60)), axis=1)  # some with coord < 1
    extra[0] = [0.4, 0.7]         # exercises the [-1] negative-index wrap in make_binned_array
    extra[1] = [30.2, 0.2]
    xy2 = np.concatenate((xy2_all[keep2], extra), axis=0)
    bf = np.stack((keep2, np.arange(len(keep2))), axis=1)
    wrong = rng.choice(len(bf), 60, replace=False)
    bf[wrong, 1] = rng.integers(0, len(xy2), 60)                                         # outlier matches
    kp1 = [KP(*p_) for p_ in xy1]
    kp2 = [KP(*p_) for p_ in xy2]
    img1 = np.zeros((Hh, Ww), np.uint8)
    img2 = np.zeros((Hh, Ww), np.uint8)

    orig = np.array([(k.pt[1], k.pt[0]) for k in kp1], dtype=float)
    prime = np.array([(k.pt[1], k.pt[0]) for k in kp2], dtype=float)
    with contextlib.redirect_stdout(io.StringIO()):
        binned = ref_q9.make_binned_array(img2.shape, prime)
    flat_cells, flat_items = [], []
    for y in range(Hh):
        for x in range(Ww):
            for j in binned[y][x]:
                flat_cells.append(y * Ww + x)
                flat_items.append(j)
    Hs = np.stack([Htrue, np.eye(3), Htrue + rng.normal(size=(3, 3)) * 1e-3])
    tps, ms, ss = [], [], []
    for Hm in Hs:
        tp = ref_q9.get_transformed_points(orig, Hm)
        m, s = ref_q9.get_ssd_v2(orig, tp, prime, binned)
        tps.append(tp)
        ms.append(m)
        ss.append(s)
    np.random.seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        Hfinal, final_matches = ref_q9.ransac_loop(img1, img2, kp1, kp2, bf)
    # hypotheses the reference drew (re-derived with the same seed) and their DLTs
    np.random.seed(7)
    picks = np.stack([np.random.randint(len(bf), size=4) for _ in range(100)])
    dlt_H = []
    for pk in picks:
        base = [ref_q9.Match(kp1[bf[t][0]].pt[0], kp1[bf[t][0]].pt[1], kp2[bf[t][1]].pt[0], kp2[bf[t][1]].pt[1])
                for t in pk]
        dlt_H.append(ref_q9.solve_homography_LS_matrix(ref_q9.make_homography_LS_matrix(base)))
    save("ransac2d", img_shape=np.array([Hh, Ww]), xy1=xy1, xy2=xy2, bf=bf,
         binned_cells=np.array(flat_cells), binned_items=np.array(flat_items),
         Hs=Hs, tp0=tps[0], tp1=tps[1], tp2=tps[2],
         m0=ms[0], m1=ms[1], m2=ms[2], s0=ss[0], s1=ss[1], s2=ss[2],
         picks=picks, dlt_H=np.stack(dlt_H),
         H_final=Hfinal, final_matches=final_matches)


if __name__ == "__main__":
    main()
