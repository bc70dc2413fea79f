"""Drop-in for utils/r3d.py: rigid_transform_3D(A, B) (utils/r3d.py:22-60).

Expects 3xN matrices, returns (R (3,3), t (3,1)); raises the same exceptions on bad shapes and
prints the same message when the reflection branch is taken.  The moments and the 3x3 SVD/Kabsch
solve run in csrc/fit.cu (fp64).
"""
import numpy as np

from .. import _capi as C
from .. import ops


def rigid_transform_3D(A, B):
    assert len(A) == len(B)
    A = np.asarray(A)
    B = np.asarray(B)
    num_rows, num_cols = A.shape
    if num_rows != 3:
        raise Exception("matrix A is not 3xN, it is {}x{}".format(num_rows, num_cols))
    num_rows, num_cols = B.shape
    if num_rows != 3:
        raise Exception("matrix B is not 3xN, it is {}x{}".format(num_rows, num_cols))
    a = C.to_device(A, np.float64)
    b = C.to_device(B, np.float64)
    T = ops.kabsch(a, b, layout="3n").cpu().numpy()
    R = T[:3, :3].copy()
    t = T[:3, 3].reshape(3, 1).copy()
    return R, t
