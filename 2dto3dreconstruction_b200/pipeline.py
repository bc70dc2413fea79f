"""Batched, device-resident registration of many depth-frame pairs, and its multi-GPU sharding.

The reference registers consecutive frames in a Python loop (reconstruct.py:264, :319) and the
pairs are independent until the prefix product of chain_transformation (reconstruct.py:201-204).
Here a whole batch of pairs goes through K1 -> grid build -> ICP in chunks that share every kernel
launch (csrc/pipeline.cu).  Across GPUs the pairs are split into contiguous shards, one process
per GPU, no communication during the iterations; the only exchange is one all-gather of the
(B/G, 4, 4) float64 transforms (128 B per pair) before the host-side chaining.
"""
import ctypes

import numpy as np

from . import _capi as C
from . import ops

_ws_cache = {}


def _workspace(n_pairs, ppc, h, w, n_streams):
    key = (n_pairs, ppc, h, w, n_streams, C.torch().cuda.current_device())
    ws = _ws_cache.get(key)
    if ws is None:
        _ws_cache.clear()                      # one live workspace at a time (it is GBs)
        nbytes = C.lib().r3d_register_workspace_bytes(n_pairs, ppc, h, w, n_streams)
        ws = C.workspace(nbytes)
        _ws_cache[key] = ws
    return ws


def pinned_like(shape, dtype):
    """Pinned host tensor (for asynchronous H2D/D2H in the host-input path)."""
    t = C.torch()
    return t.empty(shape, dtype=dtype, pin_memory=True)


def register_depth_pairs(depth_src, depth_dst, max_iterations=30, tolerance=0.0, partial_set_size=1.0,
                         pairs_per_chunk=16, n_streams=3, cell_size=0.0, fx=ops.FX, fy=ops.FY, cx=ops.CX, cy=ops.CY):
    """Register depth_src[p] onto depth_dst[p] for every pair p: pinhole back-projection of both
    frames (utils/utils.py:48-76), then ICP (utils/icp.py:74-133) of the source cloud onto the
    destination cloud.

    depth_src / depth_dst: uint16 [B,H,W]; either CUDA tensors (device path, asynchronous, results
    are CUDA tensors) or host arrays / CPU tensors (host path: chunks are staged to the device with
    copies overlapped with compute, results come back as NumPy arrays; pinned memory makes the copies
    truly asynchronous).

    Returns dict: T [B,4,4] float64, iters int32 [B] (reference convention), mean_err [B],
    n_src / n_dst int32 [B] (cloud sizes)."""
    C.require_device()
    t = C.torch()
    L = C.lib()
    prm = C.IcpParams(int(max_iterations), 0, float(tolerance), float(partial_set_size), float(cell_size))
    on_device = hasattr(depth_src, "is_cuda") and depth_src.is_cuda
    if on_device:
        assert depth_dst.is_cuda and depth_src.dtype == t.uint16 and depth_dst.dtype == t.uint16
        assert depth_src.is_contiguous() and depth_dst.is_contiguous() and depth_src.shape == depth_dst.shape
        B, H, W = depth_src.shape
        ws = _workspace(B, pairs_per_chunk, H, W, n_streams)
        out = {
            "T": t.empty((B, 4, 4), dtype=t.float64, device="cuda"),
            "iters": t.empty(B, dtype=t.int32, device="cuda"),
            "mean_err": t.empty(B, dtype=t.float64, device="cuda"),
            "n_src": t.empty(B, dtype=t.int32, device="cuda"),
            "n_dst": t.empty(B, dtype=t.int32, device="cuda"),
        }
        C.check(L.r3d_register_depth_pairs(C.ptr(depth_src), C.ptr(depth_dst), B, H, W, fx, fy, cx, cy, ctypes.byref(prm),
                                           pairs_per_chunk, n_streams, C.ptr(out["T"]), C.ptr(out["iters"]),
                                           C.ptr(out["mean_err"]), C.ptr(out["n_src"]), C.ptr(out["n_dst"]), C.ptr(ws),
                                           ws.numel(), C.stream_ptr()), "r3d_register_depth_pairs")
        return out

    # host path
    def as_host_tensor(x):
        if hasattr(x, "is_cuda"):
            assert x.dtype == t.uint16 and x.is_contiguous()
            return x
        a = np.ascontiguousarray(x)
        assert a.dtype == np.uint16
        return t.from_numpy(a)

    hs, hd = as_host_tensor(depth_src), as_host_tensor(depth_dst)
    assert hs.shape == hd.shape and hs.dim() == 3
    B, H, W = hs.shape
    ws = _workspace(B, pairs_per_chunk, H, W, n_streams)
    T = pinned_like((B, 4, 4), t.float64)
    iters = pinned_like((B,), t.int32)
    err = pinned_like((B,), t.float64)
    n_src = pinned_like((B,), t.int32)
    n_dst = pinned_like((B,), t.int32)
    C.check(L.r3d_register_depth_pairs_host(C.ptr(hs), C.ptr(hd), B, H, W, fx, fy, cx, cy, ctypes.byref(prm),
                                            pairs_per_chunk, n_streams, C.ptr(T), C.ptr(iters), C.ptr(err), C.ptr(n_src),
                                            C.ptr(n_dst), C.ptr(ws), ws.numel(), C.stream_ptr()),
            "r3d_register_depth_pairs_host")
    return {"T": T.numpy(), "iters": iters.numpy(), "mean_err": err.numpy(), "n_src": n_src.numpy(), "n_dst": n_dst.numpy()}


# ---------------------------------------------------------------------------------------------
# multi-GPU: contiguous shards, one all-gather of the transforms, host-side chaining
# ---------------------------------------------------------------------------------------------

def shard_range(n_items, rank, world_size):
    """Contiguous, balanced partition: the first (n % world) ranks get one extra item."""
    base, extra = divmod(int(n_items), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_transforms(local_T, n_total, group=None):
    """All-gather the per-shard (n_local,4,4) float64 transforms into the global (n_total,4,4)
    tensor, in pair order, on every rank.  Shards may differ by one pair, so every rank pads to the
    largest shard (one fixed-size all_gather: NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    t = C.torch()
    world = dist.get_world_size(group)
    if world == 1:
        return local_T
    per = (n_total + world - 1) // world
    pad = t.zeros((per, 4, 4), dtype=t.float64, device=local_T.device)
    pad[:local_T.shape[0]] = local_T
    bucket = [t.empty_like(pad) for _ in range(world)]
    dist.all_gather(bucket, pad, group=group)
    parts = []
    for r in range(world):
        s, e = shard_range(n_total, r, world)
        parts.append(bucket[r][:e - s])
    return t.cat(parts, dim=0)


def chain(transforms):
    """Prefix product P_i = T_i . P_{i-1} over gathered pair transforms (reconstruct.py:201-204)."""
    T = np.asarray(transforms.cpu() if hasattr(transforms, "cpu") else transforms)
    out = [T[0]]
    for k in range(1, T.shape[0]):
        out.append(np.dot(T[k], out[-1]))
    return np.stack(out)


def register_depth_pairs_sharded(depth_src, depth_dst, group=None, **kw):
    """Data-parallel registration: every rank passes the GLOBAL batch (or any object sliceable along
    dim 0), registers its contiguous shard on its own GPU, and gets the global (B,4,4) transforms
    back after one all-gather.  Requires an initialised torch.distributed process group."""
    import torch.distributed as dist
    t = C.torch()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B = depth_src.shape[0]
    s, e = shard_range(B, rank, world)
    if e > s:
        local = register_depth_pairs(depth_src[s:e], depth_dst[s:e], **kw)
        T_local = local["T"] if hasattr(local["T"], "is_cuda") else t.from_numpy(local["T"])
    else:
        # fewer pairs than ranks: this rank has nothing to register but must still take part in the collective
        local = None
        T_local = t.zeros((0, 4, 4), dtype=t.float64)
    if dist.get_backend(group) == "nccl" and not T_local.is_cuda:
        T_local = T_local.cuda()
    T_all = gather_transforms(T_local, B, group)
    return {"T": T_all, "local": local, "shard": (s, e)}


def make_exchange(buffer, group=None, stage_through_host=None):
    """The r3d_exchange_fn of r3d_icp_batch_sharded for a workspace held in the tensor `buffer`: sums the
    n_doubles float64 at device_buf (a pointer INTO buffer) over the ranks of `group`, in place, with
    torch.distributed.all_reduce on the current stream (NCCL for CUDA tensors).  With a backend that cannot reduce
    CUDA tensors (gloo without CUDA support) the slice is staged through the host.  Returns (ctypes callback,
    python function); keep the callback alive while the library may call it."""
    import torch.distributed as dist
    t = C.torch()
    base = buffer.data_ptr()
    if stage_through_host is None:
        stage_through_host = buffer.is_cuda and dist.get_backend(group) != "nccl"

    def exchange(ctx, device_buf, n_doubles, stream):
        try:
            off = int(device_buf) - base
            if off < 0 or off % 8 or off + 8 * int(n_doubles) > buffer.numel():
                raise ValueError("exchange buffer is not inside the workspace tensor")
            view = buffer[off:off + 8 * int(n_doubles)].view(t.float64)
            if stage_through_host:
                host = view.cpu()
                dist.all_reduce(host, op=dist.ReduceOp.SUM, group=group)
                view.copy_(host)
            else:
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception as e:      # an exception must not unwind through the C caller
            import sys
            print("r3d exchange failed: %r" % (e,), file=sys.stderr)
            return 1

    return C.EXCHANGE_FN(exchange), exchange


def icp_pair_sharded(A, B, init_pose=None, max_iterations=20, tolerance=0.001, group=None):
    """utils/icp.py:74-133 for ONE large pair with the queries sharded over the ranks of `group` (SURVEY.md section 8f
    N4; BASELINE.json configs[3]).  Every rank passes the same (N,3) / (M,3) float64 arrays and gets the same result,
    bit-identical to utils.icp.icp on one GPU: (T (4,4), mean distance of the last iteration, iters).  partial_set_size
    is 1.0 (include/r3d_b200.h).  One all-reduce of the virtual-warp moment rows (256 KB - 1 MB) per iteration."""
    import torch.distributed as dist
    C.require_device()
    t = C.torch()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    A = np.ascontiguousarray(A, dtype=np.float64)
    B = np.ascontiguousarray(B, dtype=np.float64)
    assert A.ndim == 2 and A.shape[1] == 3 and B.ndim == 2 and B.shape[1] == 3
    src = ops.pack_xyzw(C.to_device(A))
    dst = ops.pack_xyzw(C.to_device(B))
    so, do = ops._offsets([len(A)]), ops._offsets([len(B)])
    pose = None if init_pose is None else C.to_device(np.ascontiguousarray(init_pose, dtype=np.float64).reshape(1, 16))
    T = t.empty((1, 4, 4), dtype=t.float64, device="cuda")
    iters = t.empty(1, dtype=t.int32, device="cuda")
    err = t.empty(1, dtype=t.float64, device="cuda")
    prm = C.IcpParams(int(max_iterations), 0, float(tolerance), 1.0, 0.0)
    L = C.lib()
    nws = L.r3d_icp_workspace_bytes(1, len(A), len(B))
    ws = C.workspace(nws)
    cb, _ = make_exchange(ws, group)
    C.check(L.r3d_icp_batch_sharded(C.ptr(src), C.ptr(so), C.ptr(dst), C.ptr(do), 1, len(A), len(B), C.ptr(pose),
                                    ctypes.byref(prm), rank, world, cb, None, C.ptr(T), C.ptr(iters), C.ptr(err), C.ptr(ws),
                                    ws.numel(), C.stream_ptr()), "r3d_icp_batch_sharded")
    t.cuda.synchronize()
    return T[0].cpu().numpy(), float(err.item()), int(iters.item())
