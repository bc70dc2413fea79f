"""configs[3] (one ~2 M x 2 M pair, 30 iterations) with the queries sharded over the GPUs of a box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/bench_sharded_pair.py

One process per GPU over NCCL.  Prints the time of pipeline.icp_pair_sharded's device work (CUDA events, max over ranks)
and checks that every rank returns the same transform.  With N = 1 it times the plain single-GPU path.
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
pipeline = importlib.import_module("2dto3dreconstruction_b200.pipeline")
C = importlib.import_module("2dto3dreconstruction_b200._capi")
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
hh, ww = 1226, 1632
uu = torch.arange(ww, dtype=torch.float64, device=dev)[None, :].expand(hh, ww)
vv = torch.arange(hh, dtype=torch.float64, device=dev)[:, None].expand(hh, ww)
dd = torch.round(2000 + 600 * torch.sin(uu / 97) + 400 * torch.cos(vv / 61) + 2 * torch.randn((hh, ww), dtype=torch.float64, device=dev, generator=g))
keep = torch.rand((hh, ww), device=dev, generator=g) >= 0.1
dd, uu, vv = dd[keep], uu[keep], vv[keep]
A = torch.stack(((uu - 319.5) * dd / 525.0, (vv - 239.5) * dd / 525.0, dd), dim=1).contiguous()
ang = np.deg2rad(2.0)
Rz = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=torch.float64, device=dev)
B = (A @ Rz.T + torch.tensor([15.0, -8.0, 5.0], dtype=torch.float64, device=dev)
     + 0.5 * torch.randn(A.shape, dtype=torch.float64, device=dev, generator=g)).contiguous()
A_h, B_h = A.cpu().numpy(), B.cpu().numpy()          # same seed on every rank: identical clouds
best = None
for rep in range(4):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if world > 1:
        T, err, it = pipeline.icp_pair_sharded(A_h, B_h, max_iterations=30, tolerance=0.0)
    else:
        icp = importlib.import_module("2dto3dreconstruction_b200.utils.icp")
        T, d, it = icp.icp(A_h, B_h, max_iterations=30, tolerance=0.0)
        err = float(d.mean())
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rep > 0:
        best = float(ms) if best is None else min(best, float(ms))
Tt = torch.from_numpy(T).to(dev)
same = True
if world > 1:
    lst = [torch.empty_like(Tt) for _ in range(world)]
    dist.all_gather(lst, Tt)
    same = all(torch.equal(x, lst[0]) for x in lst)
if rank == 0:
    print("gpus %d: %d points, 30 iterations: %.2f ms per registration incl. H2D of both clouds (best of 3, max over ranks); "
          "mean_err %.6f; identical on all ranks: %s; T[0] = %s" % (world, len(A_h), best, err, same, np.array2string(T[0], precision=12).replace("\n", " ")), flush=True)
if world > 1:
    dist.destroy_process_group()
