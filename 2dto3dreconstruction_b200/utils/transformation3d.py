"""Drop-in for the registration functions of the reference's utils/transformation3d.py.

    homo_rigid_transform_3d(img1_kp, img2_kp)     utils/transformation3d.py:163-180  -> K3c kernel
    get_transformed_points(points, h)             utils/transformation3d.py:183-187  -> transform kernel
    register_imgs(...)                            utils/transformation3d.py:217-289  (glue)
    depth_to_voxel(img, scale=1.)                 utils/transformation3d.py:81-105   (same as utils.utils)

register_imgs keeps the reference's control flow, its global-RNG consumption (two np.random.choice
calls, in the same order) and its quirks (pixel-space keypoints, h.t composition order); only the
arithmetic moved to the device.  SIFT detection stays on OpenCV (out of scope); the brute-force
cross-checked descriptor matching that follows it runs in the K5 kernel (csrc/match.cu).
"""
import numpy as np

from .. import _capi as C
from .. import ops
from .homography_utils.q9 import ransac_loop
from .icp import icp
from .utils import depth_to_voxel as _depth_to_voxel


def depth_to_voxel(img, scale=1.):
    """utils/transformation3d.py:81-105 (duplicate of utils/utils.py:21-45)."""
    return _depth_to_voxel(img, scale)


def homo_rigid_transform_3d(img1_kp, img2_kp):
    """Least-squares 12-dof affine mapping img1_kp -> img2_kp as a 4x4 with last row 0 0 0 1.
    The reference solves the (3K,12) float32 system with lsq_linear/lstsq; the kernel solves the
    equivalent centred normal equations in fp64.  For a rank-deficient keypoint set (coplanar depth such as
    a flat wall, fewer than 4 matches) the normal equations are singular and the reference returns lstsq's
    minimum-norm solution: that degenerate case alone is solved the reference's way on the host (a 12-column
    system of a few hundred flops), so the drop-in returns what the reference returns instead of raising."""
    p = C.to_device(np.asarray(img1_kp), np.float64)
    q = C.to_device(np.asarray(img2_kp), np.float64)
    if p.dim() != 2 or p.shape[1] != 3 or q.shape != p.shape:
        raise ValueError("expected two (K,3) arrays")
    T, status = ops.affine_fit(p, q)
    if int(status.item()) != 0:
        return _min_norm_affine(np.asarray(img1_kp), np.asarray(img2_kp))
    return T.cpu().numpy()


def _min_norm_affine(p, q):
    """The degenerate case of homo_rigid_transform_3d: scipy's unbounded lsq_linear is np.linalg.lstsq(A, b, rcond=-1)
    on the (3K,12) float32 design matrix of utils/transformation3d.py:133-147 (rows [p 1] in the x, y and z column
    blocks) and b = img2_kp.flatten() (:163-180)."""
    K = p.shape[0]
    ph = np.concatenate((p, np.ones((K, 1))), axis=1)
    A = np.zeros((3 * K, 12), dtype=np.float32)
    A[0::3, 0:4] = ph
    A[1::3, 4:8] = ph
    A[2::3, 8:12] = ph
    x = np.linalg.lstsq(A, q.flatten(), rcond=-1)[0]
    return np.concatenate((x.reshape((3, 4)), np.array([[0, 0, 0, 1]])), axis=0)


def get_transformed_points(points, h):
    """[P 1] . h^T, xyz kept, no perspective divide (utils/transformation3d.py:183-187)."""
    pts = C.to_device(np.asarray(points), np.float64)
    T = C.to_device(np.asarray(h), np.float64)
    return ops.transform_points(pts, T).cpu().numpy()


class DMatch:
    """The three fields of cv2.DMatch the reference reads (utils/transformation3d.py:40, :254)."""
    __slots__ = ("queryIdx", "trainIdx", "distance", "imgIdx")

    def __init__(self, q, t, d):
        self.queryIdx, self.trainIdx, self.distance, self.imgIdx = int(q), int(t), float(d), 0


def match_descriptors(des1, des2):
    """cv2.BFMatcher(normType=cv2.NORM_L2, crossCheck=True).match(des1, des2) (utils/transformation3d.py:31-32):
    mutual nearest neighbours in L2, first minimum on ties, ordered by query index."""
    d1 = np.ascontiguousarray(des1, dtype=np.float32)
    d2 = np.ascontiguousarray(des2, dtype=np.float32)
    if d1.ndim != 2 or d2.ndim != 2 or d1.shape[1] != d2.shape[1]:
        raise ValueError("expected two (n, dim) descriptor arrays of equal dim")
    if d1.shape[0] == 0 or d2.shape[0] == 0:
        return []
    train, dist = ops.match_descriptors(C.to_device(d1), C.to_device(d2), cross_check=True)
    train, dist = train.cpu().numpy(), dist.cpu().numpy()
    return [DMatch(i, train[i], dist[i]) for i in np.nonzero(train >= 0)[0]]


def generate_keypoints_and_match(img1, img2):
    """SIFT keypoints (OpenCV; detection is outside the accelerated path) + cross-checked brute-force
    matches on the device (utils/transformation3d.py:15-34)."""
    import cv2
    sift = cv2.xfeatures2d.SIFT_create() if hasattr(cv2, "xfeatures2d") else cv2.SIFT_create()
    kp1, des1 = sift.detectAndCompute(img1, mask=None)
    kp2, des2 = sift.detectAndCompute(img2, mask=None)
    return kp1, kp2, match_descriptors(des1, des2)


def get_key_points_from_matches(kp1, kp2, matches):
    """Rounded int16 (x, y) of the matched keypoints (utils/transformation3d.py:66-78)."""
    p1 = np.array([kp1[m[0]].pt for m in matches])
    p2 = np.array([kp2[m[1]].pt for m in matches])
    return np.rint(p1).astype(np.int16), np.rint(p2).astype(np.int16)


def register_imgs(img1_rgb, img2_rgb, img1_depth, img2_depth, scale=1., filter_pts_frac=1., partial_set_frac=1.,
                  img1_pts=None, img2_pts=None, plot=False):
    """Global (keypoints -> RANSAC -> affine fit) then local (ICP) registration of image 1 onto
    image 2 (utils/transformation3d.py:217-289).  Returns (img1_pts, img2_pts, 4x4)."""
    if plot:
        raise NotImplementedError("plotting is outside the accelerated path")
    if img1_pts is None:
        img1_pts = depth_to_voxel(img1_depth, scale=scale)
    if img2_pts is None:
        img2_pts = depth_to_voxel(img2_depth, scale=scale)

    h, src, dst = global_registration(img1_rgb, img2_rgb, img1_depth, img2_depth, img1_pts, img2_pts, filter_pts_frac)
    t, _, _ = icp(src, dst, tolerance=0.001, max_iterations=100, partial_set_size=partial_set_frac)
    return img1_pts, img2_pts, np.dot(h, t)


def global_registration(img1_rgb, img2_rgb, img1_depth, img2_depth, img1_pts, img2_pts, filter_pts_frac):
    """Everything of register_imgs before its ICP call (utils/transformation3d.py:240-281): SIFT + matching, RANSAC,
    affine fit on the matched 3-D keypoints, and the two random sub-samples.  Consumes NumPy's global RNG exactly as
    the reference does.  Returns (h, sub-sampled transformed cloud 1, sub-sampled cloud 2): the ICP inputs."""
    kp1, kp2, matches = generate_keypoints_and_match(img1_rgb, img2_rgb)
    matches = np.array([(m.queryIdx, m.trainIdx) for m in matches])
    _, matches = ransac_loop(img1_rgb, img2_rgb, kp1, kp2, matches)

    kp1, kp2 = get_key_points_from_matches(kp1, kp2, matches)
    img1_kp = [(p[0], p[1], img1_depth[p[1], p[0]]) for p in kp1]
    img2_kp = [(p[0], p[1], img2_depth[p[1], p[0]]) for p in kp2]
    zipped = [z for z in zip(img1_kp, img2_kp) if z[0][2] != 0 and z[1][2] != 0]
    img1_kp, img2_kp = zip(*zipped)
    img1_kp = np.array(img1_kp)
    img2_kp = np.array(img2_kp)

    h = homo_rigid_transform_3d(img1_kp, img2_kp)
    img2_new = get_transformed_points(img1_pts, h)

    # same two draws from the global NumPy RNG, same order (utils/transformation3d.py:279-280)
    pts_idx1 = np.random.choice(img2_new.shape[0], int(img2_new.shape[0] * filter_pts_frac), replace=False)
    pts_idx2 = np.random.choice(img2_pts.shape[0], int(img2_pts.shape[0] * filter_pts_frac), replace=False)
    return h, img2_new[pts_idx1], img2_pts[pts_idx2]
