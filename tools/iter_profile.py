"""Per-iteration profile of the nn kernel on one chunk of synthetic pairs: launch time, rounds per task and
lanes skipped per task for every ICP iteration.

    R3D_LIB_PATH=.../libr3d_b200_stats.so python tools/iter_profile.py [pairs] [iters] [partial]

ICP is deterministic, so the run with max_iterations = k repeats the first k-1 iterations of the run before it;
the per-iteration figures are the differences of the cumulative event times / counters of successive runs.
"""
import ctypes
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
partial = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
pipeline = importlib.import_module("2dto3dreconstruction_b200.pipeline")
C = importlib.import_module("2dto3dreconstruction_b200._capi")
L = C.lib()
a, b = bench.synth_batch_torch(pairs, 0, torch.device("cuda", 0))
stats_on = L.r3d_nn_stats_read(None, 1) == 0
prev_ms, prev = 0.0, [0] * 32
impl = int(os.environ.get("IMPL", "1"))
L.r3d_set_option(3, float(impl))
if os.environ.get("SLACK"):      # "mult,maxfrac"
    mult, frac = (float(x) for x in os.environ["SLACK"].split(","))
    L.r3d_set_option(4, mult), L.r3d_set_option(5, frac)
L.r3d_set_option(7, float(os.environ.get("GROUP", "12")))
per_lane = 1.0
print("iter  launch_us  rounds/task  searching lanes/task  candidates/lane  | per task: group-searched  passes  steps  sent on to the walk (lanes) | mean_err")
for k in range(1, iters + 1):
    kw = dict(max_iterations=k, tolerance=0.0, partial_set_size=partial, pairs_per_chunk=pairs, n_streams=1)
    best = None
    for rep in range(2):
        if stats_on:
            L.r3d_nn_stats_read(None, 1)
        L.r3d_nn_timing_enable(1)
        out = pipeline.register_depth_pairs(a, b, **kw)
        torch.cuda.synchronize()
        ms, n, q = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
        L.r3d_nn_timing_read(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(q), 1)
        L.r3d_nn_timing_enable(0)
        best = ms.value if best is None else min(best, ms.value)
    st = (ctypes.c_uint64 * 32)()
    if stats_on:
        L.r3d_nn_stats_read_ext(st, 1)
    cur = [int(x) for x in st]
    d = [c - p for c, p in zip(cur, prev)]
    tasks = max(d[0], 1)
    print("%4d  %9.1f  %11.2f  %20.2f  %15.1f  | %6.3f  %6.3f  %6.3f  %6.3f | %.6f" % (
        k, 1e3 * (best - prev_ms), d[1] / tasks, 32 - d[3] / tasks, d[2] / tasks / per_lane, d[16] / tasks, d[17] / tasks,
        d[19] / tasks, d[18] / tasks, float(out["mean_err"][0])), flush=True)
    prev_ms, prev = best, cur
