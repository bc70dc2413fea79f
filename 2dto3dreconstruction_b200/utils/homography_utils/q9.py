"""Drop-in for utils/homography_utils/q9.py: 2-D homography RANSAC.

    ransac_loop(img1, img2, kp1, kp2, bf_matches, plot, x0, y0, window)   q9.py:146-210
    get_ssd_v2(orginal_points, transformed_points, keypoints_prime, binned_array_prime)   q9.py:92-127
    get_transformed_points(keypoints, H_matrix)     q9.py:52-69
    make_binned_array(img_shape, keypoints)         q9.py:129-144
    make_homography_LS_matrix / solve_homography_LS_matrix / Match   q9.py:17-50 (host, unchanged maths)

What runs where: hypothesis SAMPLING and the 4-point DLT stay on the host, in NumPy, in the
reference's order, because they consume NumPy's global RNG and numpy.linalg.eig's null-vector
choice.  All 100 hypotheses are then SCORED by one kernel launch
(csrc/ransac.cu), which is where the reference spends its time (q9.py:107-125).
"""
from collections import namedtuple

import numpy as np

from ... import _capi as C
from ... import ops

Match = namedtuple('Match', 'x y x_prime y_prime')


def LS_array_from_xi_eq(m):
    return np.array([m.x, m.y, 1, 0, 0, 0, -m.x_prime * m.x, -m.x_prime * m.y, -m.x_prime], dtype=float)


def LS_array_from_yi_eq(m):
    return np.array([0, 0, 0, m.x, m.y, 1, -m.y_prime * m.x, -m.y_prime * m.y, -m.y_prime], dtype=float)


def make_homography_LS_matrix(matches):
    """(2K,9) DLT system, two rows per correspondence (q9.py:25-38)."""
    rows = []
    for m in matches:
        rows.append(LS_array_from_xi_eq(m))
        rows.append(LS_array_from_yi_eq(m))
    return np.array(rows, dtype=float)


def solve_homography_LS_matrix(A_matrix):
    """Eigenvector of A^T A for the smallest eigenvalue, as 3x3 (q9.py:40-50).  numpy.linalg.eig is
    kept (not eigh): its choice of vector on degenerate samples is part of the reference's behaviour."""
    eigen_values, eigen_vectors = np.linalg.eig(np.dot(np.transpose(A_matrix), A_matrix))
    return eigen_vectors[:, np.argmin(eigen_values)].reshape(3, 3)


def _split_complex(H):
    H = np.asarray(H)
    re = C.to_device(np.real(H).reshape(-1, 9), np.float64)
    im = None
    if np.iscomplexobj(H) and np.any(np.imag(H) != 0):
        im = C.to_device(np.imag(H).reshape(-1, 9), np.float64)
    return re, im


def get_transformed_points(keypoints, H_matrix):
    """(row, col) keypoints through H acting on (col, row, 1), divided by w (q9.py:52-69)."""
    C.require_device()
    t = C.torch()
    kp = C.to_device(np.asarray(keypoints), np.float64)
    re, im = _split_complex(H_matrix)
    out = t.empty_like(kp)
    C.check(C.lib().r3d_homography_points(C.ptr(re), C.ptr(im), C.ptr(kp), kp.shape[0], C.ptr(out), C.stream_ptr()),
            "r3d_homography_points")
    return out.cpu().numpy()


class BinnedArray:
    """What make_binned_array returns here: the grid SHAPE plus the keypoints.  The cell lists
    themselves are built on the device inside the scoring call (with the reference's
    [int(r)-1][int(c)-1] negative-index wrap); len() and len(x[0]) work like on the reference's
    list-of-lists so that get_ssd_v2's shape reads keep working."""

    def __init__(self, img_shape, keypoints):
        self.h, self.w = int(img_shape[0]), int(img_shape[1])
        self.keypoints = np.asarray(keypoints, dtype=float)

    def __len__(self):
        return self.h

    def __getitem__(self, y):
        return _Row(self.w)


class _Row:
    def __init__(self, w):
        self.w = w

    def __len__(self):
        return self.w


def make_binned_array(img_shape, keypoints):
    """q9.py:129-144.  Raises IndexError, like the reference's list indexing, when a keypoint falls
    outside the grid."""
    kp = np.asarray(keypoints, dtype=float)
    h, w = img_shape
    if kp.size:
        r = np.trunc(kp[:, 0]) - 1
        c = np.trunc(kp[:, 1]) - 1
        if ((r < -h) | (r >= h) | (c < -w) | (c >= w)).any():
            raise IndexError("list index out of range")
    return BinnedArray((h, w), kp)


def _raise_for_flags(flags):
    f = int(np.bitwise_or.reduce(flags)) if len(flags) else 0
    if f & 4:
        raise IndexError("list index out of range")
    if f & 1:
        raise ValueError("cannot convert float NaN to integer")
    if f & 2:
        raise OverflowError("cannot convert float infinity to integer")


def _score(H, orig, prime, shape, want_matches):
    counts, flags, mj, ms = ops.ransac2d_score(H, orig, prime, shape, want_matches=want_matches)
    _raise_for_flags(flags.cpu().numpy())
    return counts.cpu().numpy(), (mj.cpu().numpy() if mj is not None else None), (ms.cpu().numpy() if ms is not None else None)


def _matches_from(mj_row, ms_row):
    i = np.nonzero(mj_row >= 0)[0]
    if i.size == 0:
        return np.array([]), np.array([])
    return np.stack((i, mj_row[i]), axis=1).astype(np.int64), ms_row[i]


def get_ssd_v2(orginal_points, transformed_points, keypoints_prime, binned_array_prime):
    """Windowed greedy-unique inlier test (q9.py:92-127).  Returns (matches (K,2), ssd (K,)).
    ``binned_array_prime`` supplies the grid shape (a BinnedArray from make_binned_array here, or
    the reference's list-of-lists); the cells are rebuilt on the device from keypoints_prime."""
    h = len(binned_array_prime)
    w = len(binned_array_prime[0])
    tp = np.asarray(transformed_points, dtype=float)
    if np.isnan(tp).any():
        raise ValueError("cannot convert float NaN to integer")
    if np.isinf(tp).any():
        raise OverflowError("cannot convert float infinity to integer")
    # identity homography: (1*c + 0*r) + 0 and / 1 are exact, so the kernel scores tp itself
    _, mj, ms = _score(np.eye(3)[None], tp, np.asarray(keypoints_prime, dtype=float), (h, w), True)
    return _matches_from(mj[0], ms[0])


def ransac_loop(img1, img2, kp1, kp2, bf_matches, plot=False, x0=None, y0=None, window=None):
    """100 x (4 random matches -> DLT -> score), refit on the best inlier set, rescore
    (q9.py:146-210).  Returns (H_matrix (3,3), final_matches (K,2))."""
    if plot:
        raise NotImplementedError("plotting is outside the accelerated path")
    orginal_points = np.array([(kp1[idx].pt[1], kp1[idx].pt[0]) for idx in range(len(kp1))], dtype=float)
    keypoints_prime = np.array([(kp2[idx].pt[1], kp2[idx].pt[0]) for idx in range(len(kp2))], dtype=float)
    shape = (img2.shape[0], img2.shape[1])
    make_binned_array(shape, keypoints_prime)      # same IndexError as the reference, same place

    # the reference draws and scores in lock step; scoring consumes no random numbers, so drawing
    # all 100 samples first leaves the global RNG stream identical (q9.py:172)
    hyps, pairings = [], []
    for r in range(100):
        match_idxs = np.random.randint(len(bf_matches), size=4)
        base = []
        for i in range(4):
            kp1_idx, kp2_idx = bf_matches[match_idxs[i]]
            xi, yi = kp1[kp1_idx].pt
            xi_prime, yi_prime = kp2[kp2_idx].pt
            base.append(Match(xi, yi, xi_prime, yi_prime))
        hyps.append(solve_homography_LS_matrix(make_homography_LS_matrix(base)))
        pairings.append(base)

    counts, mj, ms = _score(np.stack(hyps), orginal_points, keypoints_prime, shape, True)
    # strict '>' against a running best that starts at 0: the first hypothesis with the maximum
    # count wins, and nothing wins if every count is 0 (q9.py:188)
    best = int(np.argmax(counts))
    number_of_inliers = int(counts[best])
    best_matches = _matches_from(mj[best], ms[best])[0] if number_of_inliers > 0 else []

    final_inlier_matches = []
    for i in range(number_of_inliers):
        cur_kp1 = orginal_points[best_matches[i][0]]
        cur_kp2 = keypoints_prime[best_matches[i][1]]
        final_inlier_matches.append(Match(cur_kp1[1], cur_kp1[0], cur_kp2[1], cur_kp2[0]))

    H_matrix = solve_homography_LS_matrix(make_homography_LS_matrix(final_inlier_matches))
    _, mj, ms = _score(np.asarray(H_matrix)[None], orginal_points, keypoints_prime, shape, True)
    final_matches, _ = _matches_from(mj[0], ms[0])
    return H_matrix, final_matches
