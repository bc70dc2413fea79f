"""CPU-only host logic: shard partition, all-gather of transforms over gloo (world_size 2 and 3),
chaining, drop-in argument handling that never reaches the device."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, sub
from oracle import ref_port as O


def test_shard_range_partitions_exactly():
    pipeline = sub("pipeline")
    for n in (0, 1, 7, 8, 1024, 1031):
        for world in (1, 2, 3, 8):
            spans = [pipeline.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pipeline = importlib.import_module("2dto3dreconstruction_b200.pipeline")
    rng = np.random.default_rng(123)
    all_T = rng.normal(size=(n_total, 4, 4))           # every rank can regenerate the global truth
    s, e = pipeline.shard_range(n_total, rank, world)
    local = torch.from_numpy(all_T[s:e].copy())
    gathered = pipeline.gather_transforms(local, n_total)
    ok = np.array_equal(gathered.numpy(), all_T)
    chained = pipeline.chain(gathered)
    ok = ok and np.array_equal(chained, np.stack(O.chain_prefix(list(all_T))))
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
        f.write("ok" if ok else "bad")
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 10), (2, 7), (3, 8)])
def test_gather_transforms_gloo(tmp_path, world, n_total):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"


def _sharded_worker(rank, world, port, n_total, out_dir):
    """register_depth_pairs_sharded with the device call stubbed out: what is tested is the sharding and the
    collective, in particular that a rank whose shard is EMPTY (fewer pairs than ranks) still enters the all-gather."""
    sys.path.insert(0, ROOT)
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pipeline = importlib.import_module("2dto3dreconstruction_b200.pipeline")
    truth = np.random.default_rng(9).normal(size=(n_total, 4, 4))
    calls = []

    def fake_register(src, dst, **kw):
        assert src.shape[0] > 0, "the device call must not be made for an empty shard"
        calls.append(src.shape[0])
        return {"T": truth[src[:, 0, 0].astype(int)].copy()}

    pipeline.register_depth_pairs = fake_register
    frames = np.zeros((n_total, 2, 2), np.uint16)
    frames[:, 0, 0] = np.arange(n_total)                  # the pair index rides in the frame
    out = pipeline.register_depth_pairs_sharded(frames, frames)
    s, e = out["shard"]
    ok = np.array_equal(out["T"].numpy(), truth) and (calls == ([e - s] if e > s else [])) and ((out["local"] is None) == (e == s))
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
        f.write("ok" if ok else "bad")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(3, 2), (2, 1), (2, 5)])
def test_sharded_registration_with_empty_shards_gloo(tmp_path, world, n_total):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_sharded_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"


def test_drop_in_argument_errors_do_not_need_a_gpu():
    r3d = sub("utils.r3d")
    A = np.zeros((5, 3))
    with pytest.raises(Exception) as ei:
        r3d.rigid_transform_3D(A, A)
    assert "not 3xN" in str(ei.value)
    with pytest.raises(AssertionError):
        r3d.rigid_transform_3D(np.zeros((3, 5)), np.zeros((2, 5)))
    q9 = sub("utils.homography_utils.q9")
    with pytest.raises(IndexError):
        q9.make_binned_array((10, 10), np.array([[50.0, 2.0]]))
    with pytest.raises(NotImplementedError):      # integer depth beyond int32: neither the u16 nor the float64 path reproduces astype(int16)
        sub("utils.utils").depth_to_voxel(np.full((4, 4), 2**40, np.int64))


def test_q9_host_side_dlt_matches_oracle():
    q9 = sub("utils.homography_utils.q9")
    rng = np.random.default_rng(5)
    ms = [q9.Match(*rng.uniform(0, 100, 4)) for _ in range(6)]
    A = q9.make_homography_LS_matrix(ms)
    assert np.array_equal(A, O.dlt_design([O.Match(*m) for m in ms]))
    assert np.array_equal(q9.solve_homography_LS_matrix(A), O.dlt_solve(A))


def _exchange_worker(rank, world, port, out_dir):
    """The r3d_exchange_fn built by pipeline.make_exchange, called the way the library calls it (through the ctypes
    callback, with a pointer into the workspace tensor), on CPU tensors over gloo: every rank owns the rows
    r = rank (mod world) of a row table, the others are zero, and after the exchange all ranks hold the full table
    bit for bit -- the property the query-sharded ICP relies on."""
    sys.path.insert(0, ROOT)
    import ctypes
    import importlib
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pipeline = importlib.import_module("2dto3dreconstruction_b200.pipeline")
    rows = 2048
    full = torch.from_numpy(np.random.default_rng(5).normal(size=(rows, 16)) * 1e6)       # every rank regenerates the truth
    ws = torch.zeros(4096 + rows * 16 * 8, dtype=torch.uint8)                              # the "workspace": table at byte 1024
    table = ws[1024:1024 + rows * 16 * 8].view(torch.float64).view(rows, 16)
    table[rank::world] = full[rank::world]
    cb, fn = pipeline.make_exchange(ws, None)
    rc = cb(None, ctypes.c_void_p(ws.data_ptr() + 1024), rows * 16, None)
    ok = rc == 0 and np.array_equal(table.numpy(), full.numpy()) and not ws[:1024].any() and not ws[1024 + rows * 16 * 8:].any()
    # a pointer outside the buffer must be reported as a failure, not raise through the C caller
    bad = cb(None, ctypes.c_void_p(ws.data_ptr() + ws.numel()), 16, None)
    ok = ok and bad != 0
    with open(os.path.join(out_dir, "rank%d" % rank), "w") as f:
        f.write("ok" if ok else "bad")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_icp_exchange_callback_gloo(tmp_path, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_exchange_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(os.path.join(str(tmp_path), "rank%d" % r)).read() == "ok"


def test_bench_workload_generator_equals_the_oracle_generator():
    """bench.py's inputs (2dto3dreconstruction_b200/synth.py) are the SURVEY.md 8(d) NumPy frames the oracle generates:
    both arms of the benchmark, and its parity gate, register the same pairs."""
    from oracle import ref_port as O
    synth = sub("synth")
    for k, (h, w) in [(0, (480, 640)), (7, (480, 640)), (3, (96, 128))]:
        a, b = synth.depth_pair(k, h, w)
        a2, b2 = O.synth_depth_pair(k, h=h, w=w)
        assert a.dtype == np.uint16 and np.array_equal(a, a2) and np.array_equal(b, b2)
    fa, fb = synth.depth_pairs(5, 2, 48, 64)
    assert np.array_equal(fa[1], O.synth_depth_pair(6, h=48, w=64)[0])


def test_bench_helpers():
    """bench.py's host-side pieces: both arms print the same `config` object, the parity numbers are north_star's two
    quantities, and the reference arm and the CUDA arm agree on how many pairs they share."""
    import argparse
    import bench
    args = argparse.Namespace(pairs=1024, partial=1.0, pairs_per_chunk=32, streams=3, scaling="weak")
    c1, c2 = bench.make_config(args, 1), bench.make_config(args, 1)
    assert c1 == c2 and c1["pairs_per_gpu"] == 1024 and c1["global_pairs"] == 1024 and "configs[2]" in c1["workload"]
    c8 = bench.make_config(args, 8)
    assert c8["pairs_per_gpu"] == 1024 and c8["global_pairs"] == 8192
    args.scaling = "strong"
    cs = bench.make_config(args, 8)
    assert cs["pairs_per_gpu"] == 128 and cs["global_pairs"] == 1024
    T = np.eye(4)
    T2 = np.eye(4)
    T2[:3, 3] = [3.0, 0.0, 4.0]
    T2[0, 0] = 1.0 + 1e-5
    dR, dt = bench.transform_gap(T2, T, 1000.0)
    assert abs(dR - 1e-5) < 1e-12 and abs(dt - 5.0e-3) < 1e-12
    assert 1 <= bench.n_shared_pairs(1024) <= 64 and bench.n_shared_pairs(3) <= 3
