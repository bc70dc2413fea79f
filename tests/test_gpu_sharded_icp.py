"""Query-sharded ICP of one pair: `world` processes, each searching every world-th virtual
warp of the queries and exchanging the moment rows once per iteration, must return the single-GPU result BIT FOR BIT.
The driver's GPU box has one device, so the processes share cuda:0 and exchange over gloo (staged through the host);
on a multi-GPU box the same code runs one process per GPU over NCCL (tools/bench_sharded_pair.py)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, sub
from oracle import ref_port as O

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pair(n_rows):
    a, b = O.synth_depth_pair(3)
    fa = np.zeros_like(a)
    fb = np.zeros_like(b)
    fa[:n_rows], fb[:n_rows] = a[:n_rows], b[:n_rows]
    return O.depth_to_voxel_ld(fa), O.depth_to_voxel_ld(fb)


def _worker(rank, world, port, n_rows, iters, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pipeline = importlib.import_module("2dto3dreconstruction_b200.pipeline")
    A, B = _pair(n_rows)
    T, err, it = pipeline.icp_pair_sharded(A, B, max_iterations=iters, tolerance=0.0)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), T=T, err=err, it=it)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_rows,iters", [(2, 480, 12), (3, 100, 8)])      # 276 k points (32 queries per task) and 58 k
def test_sharded_pair_is_bit_identical_to_one_gpu(tmp_path, world, n_rows, iters):
    import torch.multiprocessing as mp
    icp = sub("utils.icp")
    A, B = _pair(n_rows)
    T1, d1, it1 = icp.icp(A, B, max_iterations=iters, tolerance=0.0)
    mp.spawn(_worker, args=(world, _free_port(), n_rows, iters, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        g = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        assert np.array_equal(g["T"], T1), "rank %d of %d differs from the single-GPU transform" % (r, world)
        assert int(g["it"]) == it1
        assert float(g["err"]) == float(d1.mean()) or abs(float(g["err"]) - d1.mean()) < 1e-12 * d1.mean()


def test_sharded_entry_point_argument_checks():
    import ctypes
    C = sub("_capi")
    ops = sub("ops")
    L = C.lib()
    A, B = _pair(40)
    s, d = ops.pack_xyzw(C.to_device(A)), ops.pack_xyzw(C.to_device(B))
    so, do = ops._offsets([len(A)]), ops._offsets([len(B)])
    t = C.torch()
    T = t.empty((1, 16), dtype=t.float64, device="cuda")
    ws = C.workspace(L.r3d_icp_workspace_bytes(1, len(A), len(B)))
    null_cb = C.EXCHANGE_FN(0)

    def call(prm, rank, world, cb, n_pairs=1):
        return L.r3d_icp_batch_sharded(C.ptr(s), C.ptr(so), C.ptr(d), C.ptr(do), n_pairs, len(A), len(B), None, ctypes.byref(prm), rank,
                                       world, cb, None, C.ptr(T), None, None, C.ptr(ws), ws.numel(), C.stream_ptr())

    ok = C.IcpParams(3, 0, 0.0, 1.0, 0.0)
    assert call(ok, 0, 1, null_cb) == 0                               # world 1 needs no exchange
    assert call(ok, 0, 2, null_cb) != 0                               # world > 1 without an exchange function
    assert call(ok, 2, 2, null_cb) != 0                               # rank out of range
    assert call(C.IcpParams(3, 0, 0.0, 0.7, 0.0), 0, 1, null_cb) != 0   # top-fraction selection is not sharded
    assert call(ok, 0, 1, null_cb, n_pairs=2) != 0                    # one pair only
    t.cuda.synchronize()
