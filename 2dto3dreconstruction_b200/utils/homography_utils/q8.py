"""Drop-in for the matching step of utils/homography_utils/q8.py (config 2, `--mode rigid3d`):

    mathching_skimage(img1, kp1, des1, img2, kp2, des2, plot, x0, y0, window)   q8.py:70-104
    match_descriptors(descriptors1, descriptors2, max_ratio, cross_check)       skimage.feature.match_descriptors
                                                                                 as called at q8.py:90

scikit-image (0.17.2, requirements.txt:94) is a third-party dependency of the reference and is not vendored; the
function is restated from its published algorithm: Euclidean cdist in float64, first-minimum argmin per row, mutual
check, then best / second-best < max_ratio with the best entry masked out (oracle/ref_port.py match_descriptors_skimage).
The distance matrix, both argmins, the mutual check and the ratio test run on the device (csrc/match.cu, K5); plotting
is outside the path (plot=True raises).
"""
import numpy as np

from ... import _capi as C
from ... import ops


def match_descriptors(descriptors1, descriptors2, metric=None, p=2, max_distance=np.inf, cross_check=True, max_ratio=1.0):
    """(K, 2) int64 array of matched (index in descriptors1, index in descriptors2), ascending in the first column."""
    if metric not in (None, "euclidean") or p != 2:
        raise NotImplementedError("only the Euclidean metric the reference uses (q8.py:90) is implemented")
    d1 = np.ascontiguousarray(descriptors1, dtype=np.float32)
    d2 = np.ascontiguousarray(descriptors2, dtype=np.float32)
    if d1.ndim != 2 or d2.ndim != 2 or d1.shape[1] != d2.shape[1]:
        raise ValueError("Descriptor length must equal.")          # skimage's message
    if d1.shape[0] == 0 or d2.shape[0] == 0:
        return np.empty((0, 2), dtype=np.int64)
    train, dist = ops.match_descriptors(C.to_device(d1), C.to_device(d2), cross_check=bool(cross_check), max_ratio=float(max_ratio))
    train, dist = train.cpu().numpy(), dist.cpu().numpy()
    keep = train >= 0
    if max_distance < np.inf:
        keep &= dist.astype(np.float64) < max_distance
    idx1 = np.nonzero(keep)[0]
    return np.column_stack((idx1.astype(np.int64), train[idx1].astype(np.int64)))


def mathching_skimage(img1, kp1, des1, img2, kp2, des2, plot=False, x0=None, y0=None, window=None):
    """Brute-force matches with ratio test and cross-check, optionally restricted to a window around (x0, y0) of the
    first image (q8.py:70-104).  Returns the (K, 2) array, or a list of its rows when the window filter is used."""
    if plot:
        raise NotImplementedError("plotting is outside the accelerated path")
    matches = match_descriptors(des1, des2, max_ratio=0.8, cross_check=True)
    if y0 is None or x0 is None or window is None:
        return matches
    print(x0, y0, window)
    pts = np.array([(kp1[i].pt[1], kp1[i].pt[0]) for i in range(len(kp1))], dtype=float)
    out = []
    for m in matches:
        y, x = pts[m[0]]
        if (x > x0 + window or x < x0 - window) or (y > y0 + window or y < y0 - window):
            continue
        out.append(m)
    return out
