"""Throughput of a BATCH of small pairs (the sofa pair of tests/golden replicated): python tools/batch_small.py [copies] [iters]
Small pairs use few queries per task (csrc/icp.cu queries_per_task): good for the latency of one pair; this shows what it
costs or gains when many such pairs share a launch."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

C = importlib.import_module("2dto3dreconstruction_b200._capi")
ops = importlib.import_module("2dto3dreconstruction_b200.ops")
copies = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
g = np.load(os.path.join(ROOT, "tests", "golden", "sofa_pair.npz"))
A, B = g["icp_A"], g["icp_B"]
for P in (1, copies):
    s = ops.pack_xyzw(C.to_device(np.tile(A, (P, 1))))
    d = ops.pack_xyzw(C.to_device(np.tile(B, (P, 1))))
    so, do = ops._offsets([len(A)] * P), ops._offsets([len(B)] * P)
    for partial in (1.0, 0.7):
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = ops.icp_batch(s, so, d, do, P, len(A), len(B), max_iterations=iters, tolerance=0.0, partial_set_size=partial)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        same = bool((r["T"] == r["T"][0]).all())
        print("%s: %2d pairs x %d points, %d iterations, partial %.1f: %.2f ms = %.2f ms per pair; all pairs identical: %s" % (
            os.environ.get("R3D_LIB_PATH", "default").split("/")[-1], P, len(A), iters, partial, best * 1e3, best * 1e3 / P, same), flush=True)
