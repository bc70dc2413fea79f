"""Drop-in for utils/icp.py (utils/icp.py:22-133): best_fit_transform, nearest_neighbor, icp.

NumPy arrays in, NumPy arrays out, same shapes and dtypes as the reference.  The work runs in
csrc/grid.cu + csrc/icp.cu: a sparse-block voxel grid over B is built once (the reference rebuilds
its KD-tree every iteration, utils/icp.py:68-69), and the whole iteration loop stays on the device.
"""
import numpy as np

from .. import _capi as C
from .. import ops


def best_fit_transform(A, B):
    """(m,3),(m,3) -> (4,4) least-squares rigid transform mapping A onto B (utils/icp.py:22-55)."""
    a = C.to_device(A, np.float64)
    b = C.to_device(B, np.float64)
    return ops.kabsch(a, b, layout="n3").cpu().numpy()


def _cloud(x):
    return ops.pack_xyzw(C.to_device(x, np.float64))


def nearest_neighbor(src, dst):
    """Exact Euclidean nearest neighbour in dst of every src row (utils/icp.py:58-71).
    Returns (distances float64 (N,), indices int64 (N,)); ties resolve to the lowest index."""
    src = np.asarray(src)
    dst = np.asarray(dst)
    s = _cloud(src)
    d = _cloud(dst)
    idx, dist = ops.nn_batch(s, ops._offsets([src.shape[0]]), d, ops._offsets([dst.shape[0]]), 1, max(src.shape[0], 1),
                             max(dst.shape[0], 1))
    return dist.cpu().numpy(), idx.cpu().numpy().astype(np.int64)


def icp(A, B, init_pose=None, max_iterations=20, tolerance=0.001, partial_set_size=1.):
    """Iterative Closest Point (utils/icp.py:74-133).

    Returns (T (4,4), distances (cutoff,), iters): T maps A onto B; distances are the kept
    nearest-neighbour distances of the last executed iteration (in source order; the reference's
    order is argpartition's, which is unspecified); iters follows the reference's convention
    (1-based, max_iterations + 1 when the tolerance was never met)."""
    A = np.asarray(A)
    B = np.asarray(B)
    if max_iterations > 0 and A.shape[0] > 0 and int(A.shape[0] * partial_set_size) == 0:
        # utils/icp.py:109-117 with cutoff 0: best_fit_transform of two EMPTY sets -> mean of nothing -> SVD of NaN.
        # The reference fails loudly there; so does the drop-in (the kernels report such a pair with iters = 0).
        raise np.linalg.LinAlgError("SVD did not converge")
    s = _cloud(A)
    d = _cloud(B)
    pose = None
    if init_pose is not None:
        pose = C.to_device(np.asarray(init_pose, dtype=np.float64).reshape(1, 16))
    r = ops.icp_batch(s, ops._offsets([A.shape[0]]), d, ops._offsets([B.shape[0]]), 1, max(A.shape[0], 1),
                      max(B.shape[0], 1), init_pose=pose, max_iterations=max_iterations, tolerance=tolerance,
                      partial_set_size=partial_set_size, want_last=max_iterations > 0)
    T = r["T"][0].cpu().numpy()
    iters = int(r["iters"][0].item())
    distances = None
    if max_iterations > 0:
        keep = r["keep"].bool()
        distances = r["dist"][keep].cpu().numpy()
    return T, distances, iters
