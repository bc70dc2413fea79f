"""ctypes binding of the C ABI in include/r3d_b200.h.

This module is the only place that touches the shared library.  It fails loudly when the
library is missing (there is no CPU fallback anywhere in this package): importing it is fine
on a machine without a GPU -- that is how the symbol test runs -- but every compute call checks
for an sm_100 device first and raises otherwise.

PyTorch is used for device memory, streams and pinned host buffers only.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# R3D_LIB_PATH selects another build of the same library (tools/ use the diagnostic twin built with R3D_NN_STATS=1)
LIB_PATH = os.environ.get("R3D_LIB_PATH") or os.path.join(_HERE, "lib", "libr3d_b200.so")

c_void_p, c_int, c_size_t, c_double, c_int64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_double,
                                                ctypes.c_int64)


class IcpParams(ctypes.Structure):
    """r3d_icp_params (include/r3d_b200.h)."""
    _fields_ = [("max_iterations", ctypes.c_int32), ("reserved", ctypes.c_int32), ("tolerance", c_double),
                ("partial_set_size", c_double), ("cell_size", c_double)]


#: r3d_exchange_fn (include/r3d_b200.h): int (*)(void* ctx, void* device_buf, size_t n_doubles, void* stream)
EXCHANGE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_void_p, c_size_t, c_void_p)

#: every symbol include/r3d_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "r3d_version": (c_int, []),
    "r3d_status_string": (ctypes.c_char_p, [c_int]),
    "r3d_last_cuda_error": (ctypes.c_char_p, []),
    "r3d_check_device": (c_int, []),
    "r3d_backproject_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "r3d_backproject": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_double, c_double, c_double, c_double,
                                c_double, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "r3d_kabsch_workspace_bytes": (c_size_t, [c_int64]),
    "r3d_kabsch": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_size_t,
                           c_void_p]),
    "r3d_affine_fit_workspace_bytes": (c_size_t, [c_int64]),
    "r3d_affine_fit": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "r3d_transform_points": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "r3d_xyz_to_xyzw": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "r3d_nn_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "r3d_nn_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p,
                             c_void_p, c_size_t, c_void_p]),
    "r3d_icp_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "r3d_icp_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                              ctypes.POINTER(IcpParams), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_void_p]),
    "r3d_icp_batch_sharded": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                      ctypes.POINTER(IcpParams), c_int, c_int, EXCHANGE_FN, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
    "r3d_ransac2d_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "r3d_ransac2d_score": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "r3d_homography_points": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "r3d_ransac3d_score": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_double, c_void_p, c_void_p]),
    "r3d_rigid_hypotheses": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "r3d_affine_hypotheses": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "r3d_match_workspace_bytes": (c_size_t, [c_int, c_int]),
    "r3d_match_descriptors": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                      c_void_p]),
    "r3d_match_descriptors_ratio": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                            c_size_t, c_void_p]),
    "r3d_voxel_downsample_workspace_bytes": (c_size_t, [c_int64]),
    "r3d_voxel_downsample": (c_int, [c_void_p, c_int64, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "r3d_register_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "r3d_register_depth_pairs": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double, c_double, c_double,
                                         ctypes.POINTER(IcpParams), c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_size_t, c_void_p]),
    "r3d_register_depth_pairs_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double, c_double,
                                              c_double, ctypes.POINTER(IcpParams), c_int, c_int, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "r3d_set_option": (c_int, [c_int, c_double]),
    "r3d_launch_count": (ctypes.c_uint64, []),
    "r3d_nn_timing_enable": (None, [c_int]),
    "r3d_nn_timing_read": (c_int, [ctypes.POINTER(c_double), ctypes.POINTER(c_int64), ctypes.POINTER(c_int64), c_int]),
    "r3d_nn_stats_read": (c_int, [ctypes.POINTER(ctypes.c_uint64), c_int]),
    "r3d_nn_stats_read_ext": (c_int, [ctypes.POINTER(ctypes.c_uint64), c_int]),
}

_lib = None


class R3DError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library with typed signatures."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise R3DError(
                "CUDA extension %s is missing: build it with `python __graft_entry__.py build` "
                "(there is no CPU fallback)" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError if the library does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(status, what=""):
    if status != 0:
        L = lib()
        msg = L.r3d_status_string(status).decode()
        if status == -3:
            msg += " (" + L.r3d_last_cuda_error().decode() + ")"
        raise R3DError("%s failed: %s" % (what or "r3d call", msg))


_torch = None
_device_ok = False


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_device():
    """Raise unless a B200-class (sm_100) CUDA device is usable.  No CPU path exists."""
    global _device_ok
    if _device_ok:
        return
    t = torch()
    if not t.cuda.is_available():
        raise R3DError("no CUDA device: the registration kernels are sm_100a only and have no CPU fallback")
    t.cuda.init()
    t.zeros(1, device="cuda")          # make sure a context exists before the library's first call
    check(lib().r3d_check_device(), "r3d_check_device")
    _device_ok = True


def ptr(tensor):
    return c_void_p(tensor.data_ptr()) if tensor is not None else c_void_p(0)


def stream_ptr():
    return c_void_p(torch().cuda.current_stream().cuda_stream)


def workspace(nbytes):
    t = torch()
    return t.empty(int(nbytes) + 256, dtype=t.uint8, device="cuda")


def to_device(array, dtype=None):
    """NumPy -> CUDA tensor (one H2D copy of a contiguous array)."""
    require_device()
    t = torch()
    a = np.ascontiguousarray(array, dtype=dtype)
    return t.from_numpy(a).cuda()


def launch_count():
    return int(lib().r3d_launch_count())
