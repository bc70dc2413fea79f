"""Drop-in for the registration-path helpers of the reference's reconstruct.py.

    depth_images_to_3d_pts(depth_images, scale=1.)   reconstruct.py:336-344
    depth_images_to_3d_pts_ld(depth_images)          reconstruct.py:347-354
    chain_prefix(transformations)                    reconstruct.py:201-204 (prefix product only)
    register_consecutive_pairs(point_clouds, rgb_images, depth_images, filter_pts_frac=0.1, partial_set_frac=0.7)
                                                     reconstruct.py:314-329, the pair loop of trans3d_proc, with the ICP stage
                                                     of ALL pairs in one batched call
    voxel_down_sample(points, voxel_size)            open3d PointCloud.voxel_down_sample as used at reconstruct.py:220
    merge_point_clouds(point_clouds, transformations, voxel_size=2)
                                                     reconstruct.py:201-220, the merge loop of chain_transformation on
                                                     plain (N,3) arrays (the Open3D objects, normals and meshing that
                                                     follow it are outside the path)

The reference maps depth_to_voxel over the frames in a Python list comprehension; here all frames
go through ONE K1 launch and the result is split into the same list of (N_f,3) arrays.
"""
import numpy as np

from . import _capi as C
from . import ops


def _stack(depth_images):
    frames = [np.asarray(f) for f in depth_images]
    arr = np.stack(frames, axis=0) if not isinstance(depth_images, np.ndarray) else np.asarray(depth_images)
    if arr.ndim == 4 and arr.shape[-1] == 1:
        arr = arr[..., 0]
    if arr.dtype == np.uint8:
        arr = arr.astype(np.uint16)
    elif arr.dtype not in (np.uint16, np.float32, np.float64):
        if np.issubdtype(arr.dtype, np.integer) and (arr.size == 0 or (arr.min() >= 0 and arr.max() <= 65535)):
            arr = arr.astype(np.uint16)
        else:
            arr = arr.astype(np.float64)
    return np.ascontiguousarray(arr)


def _batched(depth_images, mode, scale):
    C.require_device()
    arr = _stack(depth_images)
    d = C.torch().from_numpy(arr).cuda()
    pts, off = ops.backproject(d, mode=mode, zscale=float(scale), layout=ops.LAYOUT_XYZ)
    off_h = off.cpu().numpy()
    host = pts[:int(off_h[-1])].cpu().numpy()
    return [host[off_h[f]:off_h[f + 1]] for f in range(arr.shape[0])]


def depth_images_to_3d_pts(depth_images, scale=1.):
    """reconstruct.py:336-344."""
    if isinstance(scale, (int, np.integer)) and not isinstance(scale, bool) and int(scale) != 1:
        # integer arithmetic that wraps in int16 (utils/utils.py:40): the per-frame drop-in reproduces NumPy's rules
        from .utils.utils import depth_to_voxel
        return [depth_to_voxel(f, scale) for f in depth_images]
    out = _batched(depth_images, ops.BP_PIXEL, scale)
    if isinstance(scale, (int, np.integer)):
        out = [o.astype(np.int64) for o in out]
    return out


def depth_images_to_3d_pts_ld(depth_images):
    """reconstruct.py:347-354."""
    return _batched(depth_images, ops.BP_PINHOLE, 1.0)


def chain_prefix(transformations):
    """P_i = T_i . P_{i-1} (reconstruct.py:201-204): host side, after the pair transforms have
    been gathered; a handful of 4x4 products."""
    out = [np.asarray(transformations[0])]
    for t in transformations[1:]:
        out.append(np.dot(t, out[-1]))
    return out


def voxel_down_sample(points, voxel_size):
    """open3d.geometry.PointCloud.voxel_down_sample(voxel_size) on an (N,3) array (K7, csrc/voxel.cu): one point per
    occupied voxel, the mean of its points; voxels in ascending (z, y, x) order (Open3D's order is unspecified)."""
    C.require_device()
    t = C.torch()
    pts = np.ascontiguousarray(points, dtype=np.float64)
    if pts.ndim != 2 or pts.shape[1] != 3:
        raise ValueError("expected an (N,3) array")
    return ops.voxel_down_sample(t.from_numpy(pts).cuda(), float(voxel_size)).cpu().numpy()


def merge_point_clouds(point_clouds, transformations, voxel_size=2):
    """The merge loop of chain_transformation (reconstruct.py:201-220): transformations[i] maps cloud i+1 (after the
    prefix product) onto cloud 0's frame; every cloud is transformed, appended to the running cloud, and the running
    cloud is voxel-down-sampled -- all on the device, one count read back per merge.  Returns the merged (N,3) array."""
    C.require_device()
    t = C.torch()
    if len(transformations) != len(point_clouds) - 1:
        raise ValueError("need one transformation per consecutive pair of clouds")
    pre = chain_prefix(transformations) if len(transformations) else []
    combined = t.from_numpy(np.ascontiguousarray(point_clouds[0], dtype=np.float64)).cuda()
    for i in range(1, len(point_clouds)):
        T = np.ascontiguousarray(pre[i - 1], dtype=np.float64)
        if T.shape != (4, 4) or not np.array_equal(T[3], [0.0, 0.0, 0.0, 1.0]):
            raise ValueError("transformations must be 4x4 with last row [0,0,0,1]")      # Open3D would divide by w
        cloud = t.from_numpy(np.ascontiguousarray(point_clouds[i], dtype=np.float64)).cuda()
        moved = ops.transform_points(cloud, t.from_numpy(T).cuda())
        combined = ops.voxel_down_sample(t.cat((combined, moved), dim=0), float(voxel_size))
    return combined.cpu().numpy()


def register_consecutive_pairs(point_clouds, rgb_images, depth_images, filter_pts_frac=0.1, partial_set_frac=0.7):
    """The pair loop of trans3d_proc (reconstruct.py:314-329): register image i onto image i-1 for every consecutive
    pair and return the list of 4x4s (`all_results`).  The global stage of each pair (SIFT, matching, RANSAC, affine
    fit, sub-sampling) runs pair after pair in the reference's order, so NumPy's global RNG is consumed exactly as by
    the reference's loop; the ICP refinements, which draw no random numbers, are then run for all pairs in ONE
    r3d_icp_batch call (a single 13-26 k point pair cannot fill the GPU: tools/batch_small.py).  Every 4x4 is
    bit-identical to what utils.transformation3d.register_imgs returns for that pair."""
    from .utils import transformation3d as t3d
    C.require_device()
    hs, srcs, dsts = [], [], []
    for i in range(1, len(point_clouds)):
        h, src, dst = t3d.global_registration(rgb_images[i], rgb_images[i - 1], depth_images[i], depth_images[i - 1],
                                              point_clouds[i], point_clouds[i - 1], filter_pts_frac)
        hs.append(h)
        srcs.append(np.ascontiguousarray(src, dtype=np.float64))
        dsts.append(np.ascontiguousarray(dst, dtype=np.float64))
    if not hs:
        return []
    P = len(hs)
    s = ops.pack_xyzw(C.to_device(np.concatenate(srcs, axis=0)))
    d = ops.pack_xyzw(C.to_device(np.concatenate(dsts, axis=0)))
    ns, nd = [len(x) for x in srcs], [len(x) for x in dsts]
    r = ops.icp_batch(s, ops._offsets(ns), d, ops._offsets(nd), P, max(ns), max(nd), max_iterations=100, tolerance=0.001,
                      partial_set_size=partial_set_frac)
    T = r["T"].cpu().numpy()
    return [np.dot(hs[k], T[k]) for k in range(P)]
