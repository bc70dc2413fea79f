"""Drop-in for the back-projection functions of the reference's utils/utils.py.

    depth_to_voxel(img, scale=1)      utils/utils.py:21-45   (duplicate: utils/transformation3d.py:81-105)
    depth_to_voxel_ld(img, scale=1)   utils/utils.py:48-76

Same arguments, same return: an (N,3) array of the points with non-zero depth in row-major pixel
order, float64 -- or int64 for depth_to_voxel with an integer scale, as NumPy's type promotion gives
in the reference.  The arithmetic runs in the K1 kernel (csrc/backproject.cu); values are bit-exact.
"""
import numpy as np

from .. import _capi as C
from .. import ops


def _device_depth(img):
    """One frame as a CUDA [1,H,W] tensor in a dtype the kernel reads natively."""
    a = np.asarray(img)
    if a.ndim != 2:
        a = a.reshape(a.shape[0], a.shape[1])
    if a.dtype == np.uint16 or a.dtype == np.float32 or a.dtype == np.float64:
        a = np.ascontiguousarray(a)
    elif a.dtype == np.uint8:
        a = a.astype(np.uint16)
    elif np.issubdtype(a.dtype, np.integer):
        # astype(int16) keeps the low 16 bits of any integer type while x / y use the full value.  In-range images go through
        # the u16 path (same bits); anything else that fits int32 through the float64 path, whose z is (int16)(int32)d -- the
        # same low 16 bits -- and whose x / y are exact in double
        if a.size and (a.min() < 0 or a.max() > 65535):
            if a.min() < -2**31 or a.max() > 2**31 - 1:
                raise NotImplementedError("integer depth outside the int32 range is not supported by the CUDA path")
            a = a.astype(np.float64)
        else:
            a = a.astype(np.uint16)
    else:
        a = a.astype(np.float64)
    C.require_device()
    return C.torch().from_numpy(a[None]).cuda()


def _run(img, scale, mode):
    d = _device_depth(img)
    int_scale = isinstance(scale, (int, np.integer)) and not isinstance(scale, bool) and int(scale) != 1
    pts, off = ops.backproject(d, mode=mode, zscale=1.0 if int_scale else float(scale), layout=ops.LAYOUT_XYZ)
    n = int(off[1].item())
    out = pts[:n].cpu().numpy()
    if int_scale:
        # `img.astype(np.int16) * scale` with an INTEGER scale is integer arithmetic in NumPy: the product keeps the int16
        # dtype and wraps (or NumPy refuses a scale that does not fit), and rows whose wrapped z is 0 are dropped
        # (utils/utils.py:40-43).  The kernel delivered z = int16(depth); NumPy itself does the rest, so whatever the installed
        # NumPy's promotion rules are, they are the reference's.
        z = out[:, 2].astype(np.int16) * scale
        out[:, 2] = z
        out = out[z != 0]
    return out


def depth_to_voxel(img, scale=1):
    """Pixel-space cloud (u, v, int16(depth) * scale), zero rows dropped (utils/utils.py:21-45)."""
    out = _run(img, scale, ops.BP_PIXEL)
    if isinstance(scale, (int, np.integer)):
        return out.astype(np.int64)     # the reference stacks int64 pixel grids with an int16 depth
    return out


def depth_to_voxel_ld(img, scale=1):
    """Pinhole cloud with the reference's fixed intrinsics (utils/utils.py:48-76)."""
    return _run(img, scale, ops.BP_PINHOLE)
