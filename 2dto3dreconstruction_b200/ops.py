"""Device-level operators: thin typed wrappers over the C ABI working on CUDA tensors.

The NumPy-in / NumPy-out drop-ins in ``utils/`` are built on these; ``pipeline.py`` uses them for
the batched device-resident path.  Nothing here computes on the CPU.
"""
import ctypes

import numpy as np

from . import _capi as C

BP_PIXEL, BP_PINHOLE = 0, 1
LAYOUT_XYZ, LAYOUT_XYZW, LAYOUT_XYZW_F32 = 0, 1, 2
_DEPTH_DTYPES = {"torch.uint16": 0, "torch.float32": 1, "torch.float64": 2}

#: intrinsics hard-coded by the reference (utils/utils.py:63-66)
FX, FY, CX, CY = 525.0, 525.0, 319.5, 239.5


def backproject(depth, mode=BP_PINHOLE, zscale=1.0, layout=LAYOUT_XYZ, fx=FX, fy=FY, cx=CX, cy=CY):
    """depth: CUDA tensor [F,H,W] (uint16 / float32 / float64).  Returns (points, offsets):
    points is [F*H*W, 3|4] (capacity; rows past offsets[-1] are unspecified), offsets int32 [F+1]."""
    C.require_device()
    t = C.torch()
    assert depth.is_cuda and depth.is_contiguous() and depth.dim() == 3
    F, H, W = depth.shape
    code = _DEPTH_DTYPES[str(depth.dtype)]
    if layout == LAYOUT_XYZ:
        pts = t.empty((F * H * W, 3), dtype=t.float64, device="cuda")
    elif layout == LAYOUT_XYZW:
        pts = t.empty((F * H * W, 4), dtype=t.float64, device="cuda")
    else:
        pts = t.empty((F * H * W, 4), dtype=t.float32, device="cuda")
    off = t.empty(F + 1, dtype=t.int32, device="cuda")
    L = C.lib()
    nws = L.r3d_backproject_workspace_bytes(F, H, W)
    ws = C.workspace(nws)
    C.check(L.r3d_backproject(C.ptr(depth), code, F, H, W, mode, fx, fy, cx, cy, float(zscale), layout, C.ptr(pts),
                              C.ptr(off), C.ptr(ws), nws, C.stream_ptr()), "r3d_backproject")
    return pts, off


def kabsch(a, b, layout="n3"):
    """Rigid fit a -> b on given correspondences.  a, b: CUDA float64 tensors, (n,3) for layout
    'n3' or (3,n) for layout '3n'.  Returns a CUDA (4,4) float64 tensor."""
    C.require_device()
    t = C.torch()
    assert a.is_cuda and b.is_cuda and a.dtype == t.float64 and b.dtype == t.float64
    a = a.contiguous()
    b = b.contiguous()
    if layout == "n3":
        n = a.shape[0]
        sa = (3, 1)
    else:
        n = a.shape[1]
        sa = (1, n)
    T = t.empty((4, 4), dtype=t.float64, device="cuda")
    L = C.lib()
    nws = L.r3d_kabsch_workspace_bytes(n)
    ws = C.workspace(nws)
    C.check(L.r3d_kabsch(C.ptr(a), sa[0], sa[1], C.ptr(b), sa[0], sa[1], n, C.ptr(T), C.ptr(ws), nws, C.stream_ptr()),
            "r3d_kabsch")
    return T


def affine_fit(p, q):
    """12-dof LS affine p -> q; p, q CUDA float64 (n,3).  Returns ((4,4) tensor, status tensor)."""
    C.require_device()
    t = C.torch()
    p = p.contiguous()
    q = q.contiguous()
    n = p.shape[0]
    T = t.empty((4, 4), dtype=t.float64, device="cuda")
    status = t.zeros(1, dtype=t.int32, device="cuda")
    L = C.lib()
    nws = L.r3d_affine_fit_workspace_bytes(n)
    ws = C.workspace(nws)
    C.check(L.r3d_affine_fit(C.ptr(p), C.ptr(q), n, C.ptr(T), C.ptr(status), C.ptr(ws), nws, C.stream_ptr()),
            "r3d_affine_fit")
    return T, status


def transform_points(points, T):
    """[P 1] T^T on a CUDA (n,3) float64 tensor; T CUDA (4,4) float64."""
    C.require_device()
    t = C.torch()
    points = points.contiguous()
    out = t.empty_like(points)
    C.check(C.lib().r3d_transform_points(C.ptr(points), points.shape[0], C.ptr(T.contiguous()), C.ptr(out),
                                         C.stream_ptr()), "r3d_transform_points")
    return out


def voxel_down_sample(xyz, voxel_size):
    """K7: open3d voxel_down_sample on a CUDA (n,3) float64 tensor -> (m,3) tensor, voxels in ascending (z, y, x) order."""
    C.require_device()
    t = C.torch()
    xyz = xyz.contiguous()
    assert xyz.dtype == t.float64 and xyz.dim() == 2 and xyz.shape[1] == 3
    n = xyz.shape[0]
    if n == 0:
        return xyz.clone()
    if not (voxel_size > 0.0):
        raise RuntimeError("[VoxelDownSample] voxel_size <= 0.")         # Open3D's message
    out = t.empty_like(xyz)
    meta = t.zeros(2, dtype=t.int32, device="cuda")                      # n_out, status
    L = C.lib()
    nws = L.r3d_voxel_downsample_workspace_bytes(n)
    ws = C.workspace(nws)
    C.check(L.r3d_voxel_downsample(C.ptr(xyz), n, float(voxel_size), C.ptr(out), C.ptr(meta[0:]), C.ptr(meta[1:]), C.ptr(ws), nws,
                                   C.stream_ptr()), "r3d_voxel_downsample")
    m, status = (int(v) for v in meta.cpu())
    if status == 1:
        raise RuntimeError("[VoxelDownSample] voxel_size is too small.")     # Open3D's message
    if status == 2:
        raise ValueError("voxel_down_sample: non-finite coordinate")
    return out[:m]


def pack_xyzw(xyz):
    """(n,3) float64 CUDA tensor -> (n,4) xyzw tensor (32-byte aligned rows)."""
    C.require_device()
    t = C.torch()
    xyz = xyz.contiguous()
    out = t.empty((xyz.shape[0], 4), dtype=t.float64, device="cuda")
    C.check(C.lib().r3d_xyz_to_xyzw(C.ptr(xyz), xyz.shape[0], C.ptr(out), C.stream_ptr()), "r3d_xyz_to_xyzw")
    return out


def _offsets(sizes):
    t = C.torch()
    off = np.zeros(len(sizes) + 1, dtype=np.int32)
    np.cumsum(sizes, out=off[1:])
    return t.from_numpy(off).cuda()


def nn_batch(src, src_off, dst, dst_off, n_pairs, max_src, max_dst, cell_size=0.0):
    """Exact NN for a batch of pairs.  src/dst: xyzw CUDA tensors; offsets int32 CUDA tensors.
    Returns (idx int32 [total_src], dist float64 [total_src])."""
    C.require_device()
    t = C.torch()
    total = src.shape[0]
    idx = t.empty(total, dtype=t.int32, device="cuda")
    dist = t.empty(total, dtype=t.float64, device="cuda")
    L = C.lib()
    nws = L.r3d_nn_workspace_bytes(n_pairs, max_src, max_dst)
    ws = C.workspace(nws)
    C.check(L.r3d_nn_batch(C.ptr(src), C.ptr(src_off), C.ptr(dst), C.ptr(dst_off), n_pairs, max_src, max_dst,
                           float(cell_size), C.ptr(idx), C.ptr(dist), C.ptr(ws), nws, C.stream_ptr()), "r3d_nn_batch")
    return idx, dist


def icp_batch(src, src_off, dst, dst_off, n_pairs, max_src, max_dst, init_pose=None, max_iterations=20,
              tolerance=0.001, partial_set_size=1.0, cell_size=0.0, want_last=False, workspace=None):
    """Batched ICP (utils/icp.py:74-133 per pair).  Returns dict of CUDA tensors:
    T [P,4,4], iters int32 [P], mean_err [P] and, with want_last, idx / dist / keep per source point."""
    C.require_device()
    t = C.torch()
    T = t.empty((n_pairs, 4, 4), dtype=t.float64, device="cuda")
    iters = t.empty(n_pairs, dtype=t.int32, device="cuda")
    err = t.empty(n_pairs, dtype=t.float64, device="cuda")
    total = src.shape[0]
    idx = dist = keep = None
    if want_last:
        idx = t.empty(total, dtype=t.int32, device="cuda")
        dist = t.empty(total, dtype=t.float64, device="cuda")
        keep = t.empty(total, dtype=t.uint8, device="cuda")
    prm = C.IcpParams(int(max_iterations), 0, float(tolerance), float(partial_set_size), float(cell_size))
    L = C.lib()
    nws = L.r3d_icp_workspace_bytes(n_pairs, max_src, max_dst)
    ws = workspace if workspace is not None and workspace.numel() >= nws else C.workspace(nws)
    C.check(L.r3d_icp_batch(C.ptr(src), C.ptr(src_off), C.ptr(dst), C.ptr(dst_off), n_pairs, max_src, max_dst,
                            C.ptr(init_pose), ctypes.byref(prm), C.ptr(T), C.ptr(iters), C.ptr(err), C.ptr(idx),
                            C.ptr(dist), C.ptr(keep), C.ptr(ws), ws.numel(), C.stream_ptr()), "r3d_icp_batch")
    return {"T": T, "iters": iters, "mean_err": err, "idx": idx, "dist": dist, "keep": keep}


def ransac2d_score(H, kp1, kp2, img_shape, want_matches=False):
    """K4A.  H: (n_hyp,3,3) array (real or complex), kp1/kp2: (n,2) (row, col) float64 arrays.
    Returns (counts int32 [n_hyp], flags int32 [n_hyp], match_j or None, match_ssd or None) as CUDA
    tensors."""
    C.require_device()
    t = C.torch()
    H = np.asarray(H)
    n_hyp = H.shape[0]
    Hre = C.to_device(np.real(H).reshape(n_hyp, 9), np.float64)
    Him = None
    if np.iscomplexobj(H) and np.any(np.imag(H) != 0):
        Him = C.to_device(np.imag(H).reshape(n_hyp, 9), np.float64)
    k1 = C.to_device(kp1, np.float64)
    k2 = C.to_device(kp2, np.float64)
    n1, n2 = k1.shape[0], k2.shape[0]
    h, w = int(img_shape[0]), int(img_shape[1])
    counts = t.empty(n_hyp, dtype=t.int32, device="cuda")
    flags = t.empty(n_hyp, dtype=t.int32, device="cuda")
    mj = ms = None
    if want_matches:
        mj = t.empty((n_hyp, n1), dtype=t.int32, device="cuda")
        ms = t.empty((n_hyp, n1), dtype=t.float64, device="cuda")
    L = C.lib()
    nws = L.r3d_ransac2d_workspace_bytes(n_hyp, n1, n2, h, w)
    ws = C.workspace(nws)
    C.check(L.r3d_ransac2d_score(C.ptr(Hre), C.ptr(Him), n_hyp, C.ptr(k1), n1, C.ptr(k2), n2, h, w, C.ptr(counts),
                                 C.ptr(mj), C.ptr(ms), C.ptr(flags), C.ptr(ws), nws, C.stream_ptr()),
            "r3d_ransac2d_score")
    return counts, flags, mj, ms


def ransac3d_score(T, P, Q, thresh=10.0):
    """K4B.  T: CUDA (n_hyp,4,4) float64; P, Q: CUDA xyzw float64 (n,4).  Returns int32 counts."""
    C.require_device()
    t = C.torch()
    T = T.contiguous()
    n_hyp = T.shape[0]
    counts = t.empty(n_hyp, dtype=t.int32, device="cuda")
    C.check(C.lib().r3d_ransac3d_score(C.ptr(T), n_hyp, C.ptr(P), C.ptr(Q), P.shape[0], float(thresh), C.ptr(counts),
                                       C.stream_ptr()), "r3d_ransac3d_score")
    return counts


def match_descriptors(des1, des2, cross_check=True, max_ratio=1.0):
    """K5: brute-force L2 matching of float32 descriptor sets (CUDA tensors [n1,dim], [n2,dim]).
    max_ratio < 1 adds skimage's best / second-best ratio test (q8.py:90).
    Returns (train_idx int32 [n1] with -1 for unmatched queries, dist float32 [n1])."""
    C.require_device()
    t = C.torch()
    des1 = des1.contiguous()
    des2 = des2.contiguous()
    assert des1.dtype == t.float32 and des2.dtype == t.float32 and des1.dim() == 2 and des2.dim() == 2
    assert des1.shape[1] == des2.shape[1]
    n1, n2, dim = des1.shape[0], des2.shape[0], des1.shape[1]
    train = t.empty(n1, dtype=t.int32, device="cuda")
    dist = t.empty(n1, dtype=t.float32, device="cuda")
    L = C.lib()
    nws = L.r3d_match_workspace_bytes(n1, n2)
    ws = C.workspace(nws)
    C.check(L.r3d_match_descriptors_ratio(C.ptr(des1), n1, C.ptr(des2), n2, dim, 1 if cross_check else 0, float(max_ratio),
                                          C.ptr(train), C.ptr(dist), C.ptr(ws), nws, C.stream_ptr()), "r3d_match_descriptors_ratio")
    return train, dist


def hypotheses_3d(P, Q, samples, kind="rigid"):
    """K6: one 4x4 hypothesis per minimal sample.  P, Q: CUDA xyzw float64 (n,4); samples: CUDA int32 [n_hyp, 3]
    (kind 'rigid', utils/r3d.py:22-60 on the sample) or [n_hyp, 4] (kind 'affine',
    utils/transformation3d.py:163-180 on the sample).  Returns (T [n_hyp,4,4] float64, status int32 [n_hyp])."""
    C.require_device()
    t = C.torch()
    samples = samples.contiguous()
    k = 3 if kind == "rigid" else 4
    assert samples.dtype == t.int32 and samples.dim() == 2 and samples.shape[1] == k
    n_hyp = samples.shape[0]
    T = t.empty((n_hyp, 4, 4), dtype=t.float64, device="cuda")
    status = t.empty(n_hyp, dtype=t.int32, device="cuda")
    fn = C.lib().r3d_rigid_hypotheses if kind == "rigid" else C.lib().r3d_affine_hypotheses
    C.check(fn(C.ptr(P), C.ptr(Q), P.shape[0], C.ptr(samples), n_hyp, C.ptr(T), C.ptr(status), C.stream_ptr()),
            "r3d_%s_hypotheses" % kind)
    return T, status
