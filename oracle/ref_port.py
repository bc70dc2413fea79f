"""CPU oracle for the registration hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in NumPy, what the reference's Python does on the
registration path (back-projection, ICP, Kabsch, 12-dof affine fit, the 2-D
homography RANSAC scoring).  It exists so that the CUDA path can be checked on
machines where ``/root/reference`` is absent (the GPU box).  Nothing under
``2dto3dreconstruction_b200/`` imports it; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (through a stub shim for matplotlib / skimage /
open3d, which the arithmetic never touches), runs it on seeded inputs and stores
inputs + outputs in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks every function below against those vectors.  ``score_hypotheses_3d`` (K4B)
has no single reference function behind it: it is the composition of reference
primitives SURVEY.md section 8(c) defines, and it is pinned by ``tests/golden/k4b.npz``,
which ``tests/golden/make_golden_k4b.py`` produced by running that composition with the
reference's OWN functions (``rigid_transform_3D``, ``homo_rigid_transform_3d``,
``get_transformed_points``) imported from ``/root/reference``.

Third-party arithmetic the reference delegates to (not vendored under
``/root/reference``): scikit-learn ``NearestNeighbors`` (pinned 0.20.3,
requirements.txt:95), NumPy/LAPACK ``svd``/``det``/``eig`` (numpy 1.19.0,
requirements.txt:54) and SciPy ``lsq_linear`` (scipy 1.4.1, requirements.txt:97).
The mathematical functions are: exact Euclidean 1-NN, the singular value
decomposition, and the unbounded linear least-squares minimiser.  This port calls
the same libraries at the versions present in the image (they are part of the
image on the GPU box as well) and, for the NN search, also provides a brute-force
restatement (``nearest_neighbor_brute``) which is what "exact" means.

All ``file:line`` citations are into ``/root/reference``.
"""
from __future__ import annotations

import math
from collections import namedtuple

import numpy as np

# --------------------------------------------------------------------------------------
# Back-projection  (utils/utils.py:21-76, duplicate at utils/transformation3d.py:81-105)
# --------------------------------------------------------------------------------------

#: intrinsics hard-coded at utils/utils.py:63-66
FX, FY, CX, CY = 525, 525, 319.5, 239.5


def _pixel_grid(img):
    h, w = img.shape[0], img.shape[1]
    uu, vv = np.meshgrid(np.arange(w), np.arange(h))  # row-major: v outer, u inner
    return uu, vv


def depth_to_voxel(img, scale=1):
    """Pixel-space cloud (u, v, int16(d)*scale), zero rows dropped (utils/utils.py:21-45)."""
    uu, vv = _pixel_grid(img)
    z = img.astype(np.int16) * scale
    cloud = np.stack((uu, vv, z), axis=2).reshape(-1, 3)
    return cloud[cloud[:, 2] != 0]


def depth_to_voxel_ld(img, scale=1):
    """Pinhole cloud ((u-cx)*d/fx, (v-cy)*d/fy, int16(d)*scale) (utils/utils.py:48-76).

    The operation order ``((u - cx) * d) / fx`` of utils/utils.py:72 is kept because
    the double-precision results are compared bit for bit.
    """
    uu, vv = _pixel_grid(img)
    x = (uu - CX) * img / FX
    y = (vv - CY) * img / FY
    z = img.astype(np.int16) * scale
    cloud = np.stack((x, y, z), axis=2).reshape(-1, 3)
    return cloud[cloud[:, 2] != 0]


def depth_images_to_3d_pts(depth_images, scale=1.0):
    """reconstruct.py:336-344."""
    return [depth_to_voxel(f, scale) for f in depth_images]


def depth_images_to_3d_pts_ld(depth_images):
    """reconstruct.py:347-354."""
    return [depth_to_voxel_ld(f) for f in depth_images]


# --------------------------------------------------------------------------------------
# Kabsch / Arun  (utils/icp.py:22-55, utils/r3d.py:22-60)
# --------------------------------------------------------------------------------------

def _kabsch_rotation(H):
    """R = V U^T with the reflection fix on the last right-singular vector
    (utils/icp.py:39-45, utils/r3d.py:49-56)."""
    U, _, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt[2, :] *= -1
        R = Vt.T @ U.T
    return R


def best_fit_transform(A, B):
    """(m,3),(m,3) -> 4x4 mapping A onto B (utils/icp.py:22-55)."""
    ca = A.mean(axis=0)
    cb = B.mean(axis=0)
    H = (A - ca).T @ (B - cb)
    R = _kabsch_rotation(H)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = cb - R @ ca
    return T


def rigid_transform_3D(A, B):
    """(3,K),(3,K) -> R (3,3), t (3,1) (utils/r3d.py:22-60); same exceptions."""
    assert len(A) == len(B)
    rows, cols = A.shape
    if rows != 3:
        raise Exception("matrix A is not 3xN, it is {}x{}".format(rows, cols))
    rows, cols = B.shape
    if rows != 3:
        raise Exception("matrix B is not 3xN, it is {}x{}".format(rows, cols))
    ca = A.mean(axis=1).reshape(3, 1)
    cb = B.mean(axis=1).reshape(3, 1)
    H = (A - ca) @ (B - cb).T
    R = _kabsch_rotation(H)
    t = -R @ ca + cb
    return R, t


# --------------------------------------------------------------------------------------
# Nearest neighbour  (utils/icp.py:58-71)
# --------------------------------------------------------------------------------------

def nearest_neighbor(src, dst):
    """The reference's call: sklearn NearestNeighbors(n_neighbors=1) (utils/icp.py:68-70).
    float64, algorithm 'auto' -> KD-tree, tie order = tree traversal order."""
    from sklearn.neighbors import NearestNeighbors
    nn = NearestNeighbors(n_neighbors=1)
    nn.fit(dst)
    d, i = nn.kneighbors(src, return_distance=True)
    return d.ravel(), i.ravel()


def nearest_neighbor_brute(src, dst, chunk=2048):
    """Definition of the exact answer: argmin_j |src_i - dst_j|^2 in float64 in the
    direct (a-b)^2 form, lowest j on ties.  O(N*M); use for small cases."""
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.ascontiguousarray(dst, dtype=np.float64)
    n = src.shape[0]
    idx = np.empty(n, dtype=np.int64)
    d2 = np.empty(n, dtype=np.float64)
    for s in range(0, n, chunk):
        q = src[s:s + chunk]
        dx = q[:, None, 0] - dst[None, :, 0]
        dy = q[:, None, 1] - dst[None, :, 1]
        dz = q[:, None, 2] - dst[None, :, 2]
        dd = dx * dx + dy * dy + dz * dz
        j = dd.argmin(axis=1)  # first minimum = lowest index
        idx[s:s + chunk] = j
        d2[s:s + chunk] = dd[np.arange(q.shape[0]), j]
    return np.sqrt(d2), idx


# --------------------------------------------------------------------------------------
# ICP  (utils/icp.py:74-133)
# --------------------------------------------------------------------------------------

def icp(A, B, init_pose=None, max_iterations=20, tolerance=0.001, partial_set_size=1.0,
        nn=nearest_neighbor, trace=None):
    """Point-to-point ICP with top-fraction matching (utils/icp.py:74-133).

    ``nn`` lets tests swap the KD-tree for the brute-force definition; ``trace``
    (a list) receives (indices, distances, T) per iteration for step-wise parity.
    Returns (T, distances_of_last_iteration, iters) with the reference's iteration
    count convention: starts at 1, equals max_iterations+1 when never converged.
    """
    n = A.shape[0]
    src = np.vstack((A.T, np.ones((1, n))))
    dst = np.vstack((B.T, np.ones((1, B.shape[0]))))
    if init_pose is not None:
        src = init_pose @ src

    prev_error = 0
    distances = None
    iters = 1
    while iters <= max_iterations:
        distances, indices = nn(src[:3].T, dst[:3].T)
        cutoff = int(src.shape[1] * partial_set_size)          # utils/icp.py:109
        keep = np.argpartition(distances, cutoff - 1)[:cutoff]  # utils/icp.py:110
        distances = distances[keep]
        indices = indices[keep]
        T = best_fit_transform(src[:3, keep].T, dst[:3, indices].T)
        src = T @ src
        if trace is not None:
            trace.append((keep.copy(), indices.copy(), distances.copy(), T.copy()))
        mean_error = distances.mean()
        if abs(prev_error - mean_error) < tolerance:
            break
        prev_error = mean_error
        iters += 1

    T = best_fit_transform(A, src[:3].T)  # utils/icp.py:131
    return T, distances, iters


# --------------------------------------------------------------------------------------
# 12-dof affine ("3-D homography") fit  (utils/transformation3d.py:133-187)
# --------------------------------------------------------------------------------------

def make_homography_ls_matrix(keypoints):
    """(K,3) -> (3K,12) float32 block design matrix (utils/transformation3d.py:133-147)."""
    kp = np.asarray(keypoints)
    k = kp.shape[0]
    hom = np.concatenate((kp, np.ones((k, 1), dtype=kp.dtype)), axis=1)
    a = np.zeros((3 * k, 12), dtype=np.result_type(kp.dtype, np.int64))
    for r in range(3):
        a[r::3, 4 * r:4 * r + 4] = hom
    return a.astype(np.float32)


def homo_rigid_transform_3d(img1_kp, img2_kp):
    """LS affine p -> q.  The reference goes through scipy.optimize.lsq_linear with no
    bounds (utils/transformation3d.py:163-180), which returns numpy.linalg.lstsq's
    minimum-norm solution directly; that is what is called here."""
    A = make_homography_ls_matrix(img1_kp)
    b = np.asarray(img2_kp).flatten()
    x = np.linalg.lstsq(A, b, rcond=-1)[0]
    return np.concatenate((x.reshape(3, 4), np.array([[0, 0, 0, 1]])), axis=0)


def homo_rigid_transform_3d_scipy(img1_kp, img2_kp):
    """Same fit through the reference's literal call path (scipy lsq_linear)."""
    from scipy import optimize
    A = make_homography_ls_matrix(img1_kp)
    b = np.asarray(img2_kp).flatten()
    x = optimize.lsq_linear(A, b)['x']
    return np.concatenate((x.reshape(3, 4), np.array([[0, 0, 0, 1]])), axis=0)


def get_transformed_points(points, h):
    """[P 1] h^T, xyz kept, no perspective divide (utils/transformation3d.py:183-187)."""
    hom = np.concatenate((points, np.ones((points.shape[0], 1))), axis=1)
    return (hom @ h.T)[:, :3]


# --------------------------------------------------------------------------------------
# 2-D homography RANSAC  (utils/homography_utils/q9.py)
# --------------------------------------------------------------------------------------

Match = namedtuple('Match', 'x y x_prime y_prime')  # q9.py:17


def dlt_design(matches):
    """Two rows per correspondence (q9.py:19-38)."""
    rows = []
    for m in matches:
        rows.append([m.x, m.y, 1, 0, 0, 0, -m.x_prime * m.x, -m.x_prime * m.y, -m.x_prime])
        rows.append([0, 0, 0, m.x, m.y, 1, -m.y_prime * m.x, -m.y_prime * m.y, -m.y_prime])
    return np.array(rows, dtype=float)


def dlt_solve(A):
    """Eigenvector of A^T A with the smallest eigenvalue, as a 3x3 (q9.py:40-50).
    numpy.linalg.eig (non-symmetric driver) is kept, as the reference uses it."""
    w, v = np.linalg.eig(A.T @ A)
    return v[:, np.argmin(w)].reshape(3, 3)


def q9_get_transformed_points(keypoints, H):
    """(row, col) keypoints through H acting on (col, row, 1), divided by w, returned as
    (row, col) (q9.py:52-69).  Vectorised; the per-point product is a 3-term dot."""
    kp = np.asarray(keypoints, dtype=float)
    out = np.zeros(kp.shape)
    for i in range(kp.shape[0]):
        v = np.array([[kp[i, 1], kp[i, 0], 1.0]])
        p = H @ v.T
        q = p[:2, :] / p[2, :]
        out[i] = [q[1, 0], q[0, 0]]
    return out


def make_binned_array(img_shape, keypoints):
    """Cell lists of keypoint indices, cell = [int(r)-1][int(c)-1] with Python's negative
    index wrap-around (q9.py:129-144).  Returned in CSR form (cell_start, items) plus the
    grid shape; cell id = row * w + col, items inside a cell in insertion order."""
    h, w = img_shape[0], img_shape[1]
    cells = np.empty(len(keypoints), dtype=np.int64)
    for i, pt in enumerate(keypoints):
        r = int(pt[0]) - 1
        c = int(pt[1]) - 1
        if r < -h or r >= h or c < -w or c >= w:
            raise IndexError("list index out of range")
        cells[i] = (r % h) * w + (c % w)
    order = np.argsort(cells, kind="stable")
    counts = np.bincount(cells, minlength=h * w)
    start = np.zeros(h * w + 1, dtype=np.int64)
    np.cumsum(counts, out=start[1:])
    return {"h": h, "w": w, "start": start, "items": order.astype(np.int64)}


def _py_int(x):
    """Python int(float): truncation toward zero; raises on nan/inf like the reference."""
    return int(x)


def get_ssd_v2(original_points, transformed_points, keypoints_prime, binned):
    """Greedy-unique windowed nearest keypoint test (q9.py:92-127).

    For every transformed point, in order: gather the keypoints of image 2 lying in
    the cell window rows [int(r)-6, int(r)+5) x cols [int(c)-6, int(c)+5) (clipped to
    the grid), in row-major cell order then insertion order; take the first minimum of
    the squared distance; accept when it is < 10 and that keypoint was not accepted
    before.  Returns (matches (K,2) int, ssd (K,))."""
    h, w, start, items = binned["h"], binned["w"], binned["start"], binned["items"]
    kp2 = np.asarray(keypoints_prime, dtype=float)
    used = set()
    matches, ssd = [], []
    for i in range(len(transformed_points)):
        pt = transformed_points[i]
        r0 = _py_int(pt[0])
        c0 = _py_int(pt[1])
        cand = []
        clo, chi = max(0, c0 - 6), min(c0 + 5, w)
        if chi > clo:
            for y in range(max(0, r0 - 6), min(r0 + 5, h)):
                s, e = start[y * w + clo], start[y * w + chi]
                if e > s:
                    cand.append(items[s:e])
        if not cand:
            continue
        cand = np.concatenate(cand)
        diff = kp2[cand] - pt
        vals = diff[:, 0] ** 2 + diff[:, 1] ** 2
        k = int(np.argmin(vals))
        j = int(cand[k])
        if vals[k] < 10 and j not in used:
            matches.append((i, j))
            ssd.append(vals[k])
            used.add(j)
    return np.array(matches), np.array(ssd)


def ransac_hypotheses(kp1_pts, kp2_pts, bf_matches, n_iter=100):
    """The sampling half of q9.py:169-182: ``n_iter`` draws of
    ``np.random.randint(len(bf_matches), size=4)`` from the GLOBAL NumPy RNG and the
    4-point DLT of each.  kp*_pts are (x, y) as in cv2.KeyPoint.pt.  Returns
    (H (n_iter,3,3), sample indices (n_iter,4))."""
    Hs = []
    samples = np.empty((n_iter, 4), dtype=np.int64)
    for r in range(n_iter):
        pick = np.random.randint(len(bf_matches), size=4)
        base = []
        for t in range(4):
            i1, i2 = bf_matches[pick[t]]
            x, y = kp1_pts[i1]
            xp, yp = kp2_pts[i2]
            base.append(Match(x, y, xp, yp))
        Hs.append(dlt_solve(dlt_design(base)))
        samples[r] = pick
    # numpy.linalg.eig can return a complex eigenvector for degenerate samples (repeated
    # index); the reference then carries a complex H into q9.get_transformed_points, where
    # only the final assignment into a float array drops the imaginary part (q9.py:67).
    return np.stack(Hs), samples


def ransac_loop_pts(img2_shape, kp1_pts, kp2_pts, bf_matches, n_iter=100):
    """q9.ransac_loop (q9.py:146-210) on plain (x, y) arrays instead of cv2.KeyPoint
    objects.  Returns (H, final_matches, counts_per_hypothesis)."""
    kp1_pts = np.asarray(kp1_pts, dtype=float)
    kp2_pts = np.asarray(kp2_pts, dtype=float)
    orig = kp1_pts[:, ::-1].copy()    # (row, col)  q9.py:162
    prime = kp2_pts[:, ::-1].copy()   # q9.py:163
    binned = make_binned_array(img2_shape, prime)

    best_n, best_matches = 0, []
    counts = []
    Hs, _ = ransac_hypotheses(kp1_pts, kp2_pts, bf_matches, n_iter)
    for r in range(n_iter):
        tp = q9_get_transformed_points(orig, Hs[r])
        m, _ = get_ssd_v2(orig, tp, prime, binned)
        counts.append(len(m))
        if len(m) > best_n:          # strict '>': the earliest best hypothesis wins
            best_n, best_matches = len(m), m

    final = [Match(orig[a][1], orig[a][0], prime[b][1], prime[b][0]) for a, b in best_matches[:best_n]]
    H = dlt_solve(dlt_design(final))
    tp = q9_get_transformed_points(orig, H)
    final_matches, _ = get_ssd_v2(orig, tp, prime, binned)
    return H, final_matches, np.array(counts)


class _KP:
    """Stand-in for cv2.KeyPoint: only ``.pt`` is read by q9.ransac_loop."""
    __slots__ = ("pt",)

    def __init__(self, x, y):
        self.pt = (float(x), float(y))


def ransac_loop(img1, img2, kp1, kp2, bf_matches, plot=False, x0=None, y0=None, window=None):
    """Signature-compatible q9.ransac_loop (q9.py:146).  kp1/kp2: objects with ``.pt``."""
    k1 = np.array([k.pt for k in kp1], dtype=float).reshape(-1, 2)
    k2 = np.array([k.pt for k in kp2], dtype=float).reshape(-1, 2)
    H, fm, _ = ransac_loop_pts(img2.shape, k1, k2, bf_matches)
    return H, fm


# --------------------------------------------------------------------------------------
# register_imgs (utils/transformation3d.py:217-289): the glue of BASELINE.json configs[0].
# --------------------------------------------------------------------------------------

def register_imgs(img1, img2, depth1, depth2, pts1, pts2, filter_pts_frac=1., partial_set_frac=1., run_icp=True, record=None):
    """utils/transformation3d.py:240-289 out of the oracle's own pieces: SIFT + cross-checked brute-force matching
    (OpenCV, the reference's third-party dependency, :27-32), ransac_loop (:243), rounded int16 keypoints and their
    depths with zero-depth pairs dropped (:252-261), the affine fit (:267), the transformed cloud (:268), the two
    np.random.choice sub-samples from the global RNG (:271-272), ICP (:275-278) and h.t (:289).
    run_icp=False stops before ICP (which draws no random numbers) and returns h; `record` (a dict) receives the
    intermediate results."""
    import cv2
    sift = cv2.xfeatures2d.SIFT_create() if hasattr(cv2, "xfeatures2d") else cv2.SIFT_create()
    kp1, des1 = sift.detectAndCompute(img1, mask=None)
    kp2, des2 = sift.detectAndCompute(img2, mask=None)
    bf = cv2.BFMatcher(normType=cv2.NORM_L2, crossCheck=True).match(des1, des2)
    bf = np.array([(m.queryIdx, m.trainIdx) for m in bf])
    _, matches = ransac_loop(img1, img2, kp1, kp2, bf)
    p1 = np.rint(np.array([kp1[m[0]].pt for m in matches])).astype(np.int16)
    p2 = np.rint(np.array([kp2[m[1]].pt for m in matches])).astype(np.int16)
    k1 = [(p[0], p[1], depth1[p[1], p[0]]) for p in p1]
    k2 = [(p[0], p[1], depth2[p[1], p[0]]) for p in p2]
    kept = [z for z in zip(k1, k2) if z[0][2] != 0 and z[1][2] != 0]
    k1, k2 = (np.array(x) for x in zip(*kept))
    h = homo_rigid_transform_3d(k1, k2)
    moved = get_transformed_points(pts1, h)
    i1 = np.random.choice(moved.shape[0], int(moved.shape[0] * filter_pts_frac), replace=False)
    i2 = np.random.choice(pts2.shape[0], int(pts2.shape[0] * filter_pts_frac), replace=False)
    if record is not None:
        record.update(matches=np.array(matches), h=h, n_icp_src=len(i1))
    if not run_icp:
        return h
    t, _, it = icp(moved[i1], pts2[i2], tolerance=0.001, max_iterations=100, partial_set_size=partial_set_frac)
    if record is not None:
        record.update(icp_iters=it, icp_T=t)
    return np.dot(h, t)


# --------------------------------------------------------------------------------------
# K4B: 3-D hypothesis scoring (BASELINE.json configs[4]).  The reference has no single function for it;
# SURVEY.md section 8(c) defines it as a composition of reference primitives: residual of
# get_transformed_points (utils/transformation3d.py:183-187) against Q, inlier iff the
# squared residual is below the reference's SSD threshold (q9.py:122), counted in fp64.
# Pinned by tests/golden/k4b.npz: the same composition run with the reference's own functions.
# --------------------------------------------------------------------------------------

def score_hypotheses_3d(T, P, Q, thresh=10.0, chunk=64):
    T = np.asarray(T, dtype=np.float64).reshape(-1, 4, 4)
    P = np.asarray(P, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    counts = np.empty(T.shape[0], dtype=np.int64)
    for s in range(0, T.shape[0], chunk):
        Tc = T[s:s + chunk]
        # x' = R p + t evaluated coordinate by coordinate, left to right
        X = (Tc[:, None, :3, 0] * P[None, :, 0:1] + Tc[:, None, :3, 1] * P[None, :, 1:2]
             + Tc[:, None, :3, 2] * P[None, :, 2:3] + Tc[:, None, :3, 3])
        D = X - Q[None]
        r2 = D[..., 0] * D[..., 0] + D[..., 1] * D[..., 1] + D[..., 2] * D[..., 2]
        counts[s:s + chunk] = (r2 < thresh).sum(axis=1)
    return counts


def rigid_hypotheses(P, Q, samples):
    """3-point Kabsch hypotheses via rigid_transform_3D (utils/r3d.py:22) -> (n,4,4)."""
    out = np.zeros((len(samples), 4, 4))
    for k, s in enumerate(samples):
        R, t = rigid_transform_3D(P[s].T, Q[s].T)
        out[k, :3, :3] = R
        out[k, :3, 3] = t[:, 0]
        out[k, 3, 3] = 1
    return out


def affine_hypotheses(P, Q, samples):
    """4-point affine hypotheses via homo_rigid_transform_3d -> (n,4,4)."""
    return np.stack([homo_rigid_transform_3d(P[s], Q[s]) for s in samples])


# --------------------------------------------------------------------------------------
# K5: cross-checked brute-force descriptor matching.
# The reference calls cv2.BFMatcher(normType=cv2.NORM_L2, crossCheck=True).match(des1, des2)
# (utils/transformation3d.py:31-32); OpenCV is a third-party dependency (opencv-contrib-python 3.4.2.16,
# requirements.txt:59) that is not vendored under /root/reference.  Its published behaviour, restated: a
# query i and a train j are matched iff j = argmin_j |d1_i - d2_j| and i = argmin_i |d1_i - d2_j| (first
# minimum on ties), the distance is the float32 L2 norm, and the matches come ordered by query index.
# Pinned against the OpenCV in this image (4.13) by tests/test_oracle_golden.py::test_match_descriptors_vs_cv2;
# releases before 3.4.9 / 4.2 applied a weaker, one-directional cross-check.
# --------------------------------------------------------------------------------------

def match_descriptors(des1, des2):
    """Returns (query_idx int64 [K], train_idx int64 [K], dist float32 [K])."""
    d1 = np.asarray(des1, dtype=np.float32)
    d2 = np.asarray(des2, dtype=np.float32)
    D2 = np.empty((d1.shape[0], d2.shape[0]), dtype=np.float32)
    for s in range(0, d1.shape[0], 256):
        diff = d1[s:s + 256, None, :] - d2[None, :, :]
        D2[s:s + 256] = (diff * diff).sum(axis=2, dtype=np.float32)
    q2t = D2.argmin(axis=1)
    t2q = D2.argmin(axis=0)
    q = np.nonzero(t2q[q2t] == np.arange(d1.shape[0]))[0]
    return q, q2t[q], np.sqrt(D2[q, q2t[q]]).astype(np.float32)


def match_descriptors_skimage(des1, des2, max_ratio=0.8, cross_check=True):
    """skimage.feature.match_descriptors(des1, des2, max_ratio=..., cross_check=...) as the reference calls it at
    utils/homography_utils/q8.py:90 (scikit-image 0.17.2, requirements.txt:94 -- a third-party dependency that is neither
    vendored nor installed here, so this restates its published algorithm, skimage/feature/match.py; PARITY UNPINNED
    against the package itself).  Returns the (K, 2) int64 array of (index1, index2)."""
    d1 = np.asarray(des1, dtype=np.float64)
    d2 = np.asarray(des2, dtype=np.float64)
    distances = np.empty((d1.shape[0], d2.shape[0]))
    for s in range(0, d1.shape[0], 256):                     # = scipy cdist(..., 'euclidean')
        diff = d1[s:s + 256, None, :] - d2[None, :, :]
        distances[s:s + 256] = np.sqrt((diff * diff).sum(axis=2))
    indices1 = np.arange(d1.shape[0])
    indices2 = np.argmin(distances, axis=1)
    if cross_check:
        matches1 = np.argmin(distances, axis=0)
        mask = indices1 == matches1[indices2]
        indices1, indices2 = indices1[mask], indices2[mask]
    if max_ratio < 1.0:
        best = distances[indices1, indices2]
        distances[indices1, indices2] = np.inf
        second_idx = np.argmin(distances[indices1], axis=1)
        second = distances[indices1, second_idx]
        second[second == 0] = np.finfo(np.double).eps
        mask = best / second < max_ratio
        indices1, indices2 = indices1[mask], indices2[mask]
    return np.column_stack((indices1, indices2)).astype(np.int64)


# --------------------------------------------------------------------------------------
# Chaining prefix product  (reconstruct.py:201-204)
# --------------------------------------------------------------------------------------

def chain_prefix(transformations):
    out = [np.asarray(transformations[0])]
    for t in transformations[1:]:
        out.append(np.dot(t, out[-1]))
    return out


# --------------------------------------------------------------------------------------
# Merge + voxel down-sample  (reconstruct.py:201-220).  open3d 0.9.0 (requirements.txt) is a third-party
# dependency that is neither vendored nor installed here: PointCloud::Transform and PointCloud::VoxelDownSample
# are restated from their published source -- PARITY UNPINNED against the package itself.  Open3D emits the
# voxels in unordered_map order (unspecified); this restatement emits them in ascending (z, y, x) voxel order.
# --------------------------------------------------------------------------------------

def voxel_down_sample(points, voxel_size):
    pts = np.asarray(points, dtype=np.float64)
    if voxel_size <= 0.0:
        raise RuntimeError("[VoxelDownSample] voxel_size <= 0.")
    if len(pts) == 0:
        return pts.copy()
    min_bound = pts.min(axis=0) - voxel_size * 0.5
    max_bound = pts.max(axis=0) + voxel_size * 0.5
    if voxel_size * np.iinfo(np.int32).max < (max_bound - min_bound).max():
        raise RuntimeError("[VoxelDownSample] voxel_size is too small.")
    vox = np.floor((pts - min_bound) / voxel_size).astype(np.int64)
    order = np.lexsort((vox[:, 0], vox[:, 1], vox[:, 2]))               # stable: index order inside a voxel
    sv = vox[order]
    head = np.ones(len(pts), dtype=bool)
    head[1:] = np.any(sv[1:] != sv[:-1], axis=1)
    vid = np.cumsum(head) - 1
    sums = np.zeros((int(vid[-1]) + 1, 3))
    np.add.at(sums, vid, pts[order])                                    # unbuffered: sequential accumulation in index order
    counts = np.bincount(vid).astype(np.float64)
    return sums / counts[:, None]


def merge_point_clouds(point_clouds, transformations, voxel_size=2):
    pre = chain_prefix(transformations) if len(transformations) else []
    combined = np.asarray(point_clouds[0], dtype=np.float64)
    for i in range(1, len(point_clouds)):
        moved = get_transformed_points(np.asarray(point_clouds[i], dtype=np.float64), pre[i - 1])
        combined = voxel_down_sample(np.vstack((combined, moved)), voxel_size)
    return combined


# --------------------------------------------------------------------------------------
# Synthetic workloads.  Shared by tests and bench.py.
# --------------------------------------------------------------------------------------

def synth_depth_pair(pair_index, h=480, w=640, seed_base=1234):
    """One C3 depth pair: frame a = smooth surface + N(0,2), 10% dropout; frame b = the
    clean surface shifted by integer (du,dv) in [-8,8]^2 and dz in [-20,20], fresh noise
    and dropout.  uint16."""
    rng = np.random.default_rng(seed_base + pair_index)
    u = np.arange(w)[None, :].astype(np.float64)
    v = np.arange(h)[:, None].astype(np.float64)

    def clean(uu, vv):
        return 2000.0 + 600.0 * np.sin(uu / 97.0) + 400.0 * np.cos(vv / 61.0)

    du, dv = rng.integers(-8, 9, size=2)
    dz = rng.integers(-20, 21)
    a = np.rint(clean(u, v) + rng.normal(0.0, 2.0, size=(h, w)))
    a[rng.random((h, w)) < 0.10] = 0
    b = np.rint(clean(u + du, v + dv) + dz + rng.normal(0.0, 2.0, size=(h, w)))
    b[rng.random((h, w)) < 0.10] = 0
    return a.astype(np.uint16), b.astype(np.uint16)


def scene_extent(points):
    p = np.asarray(points)
    return float(np.linalg.norm(p.max(axis=0) - p.min(axis=0)))


def nearest_neighbor_brute_c(src, dst, threads=None):
    """Same definition as nearest_neighbor_brute, through oracle/nn_brute.c for sizes where the
    NumPy distance matrix would not fit; the queries are split over host threads (ctypes releases
    the GIL).  Returns (distances, indices)."""
    import ctypes
    import os
    from concurrent.futures import ThreadPoolExecutor
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libnn_brute.so")
    lib = ctypes.CDLL(path)
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.ascontiguousarray(dst, dtype=np.float64)
    n, m = src.shape[0], dst.shape[0]
    idx = np.empty(n, dtype=np.int64)
    d2 = np.empty(n, dtype=np.float64)
    threads = threads or os.cpu_count() or 1
    step = max(1, (n + threads - 1) // threads)

    def work(s):
        e = min(n, s + step)
        lib.nn_brute(ctypes.c_void_p(src.ctypes.data + s * 24), ctypes.c_int64(e - s),
                     dst.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(m),
                     ctypes.c_void_p(idx.ctypes.data + s * 8), ctypes.c_void_p(d2.ctypes.data + s * 8))

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(work, range(0, n, step)))
    return np.sqrt(d2), idx
