"""Timing sweep of the nn-search variants on one chunk of synthetic pairs (device-resident inputs).

    python tools/sweep_nn.py [pairs] [iters] [partial]

For every (impl, group size, cell factor) it registers the same chunk twice (warm-up + timed), reads
the library's per-launch event timing of the nn kernel, and checks that the transforms agree with
the first configuration to 1e-9 (every variant is an exact search; only summation order may differ
between group sizes).
"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
partial = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
configs = os.environ.get("SWEEP", "2:32:2.0,1:32:2.0,2:32:1.5,2:32:2.5,2:32:3.0")

pipeline = importlib.import_module("2dto3dreconstruction_b200.pipeline")
C = importlib.import_module("2dto3dreconstruction_b200._capi")
L = C.lib()
a, b = bench.synth_batch_torch(pairs, 0, torch.device("cuda", 0))
if os.environ.get("SWEEP_CROP"):      # "w,h": keep a centred window of the frames only (smaller clouds, same point spacing)
    cw, ch = (int(x) for x in os.environ["SWEEP_CROP"].split(","))
    def crop(x):      # through NumPy: torch has no uint16 arithmetic on the device
        h = x.cpu().numpy().copy()
        m = np.zeros(h.shape[1:], dtype=bool)
        m[(480 - ch) // 2:(480 + ch) // 2, (640 - cw) // 2:(640 + cw) // 2] = True
        h[:, ~m] = 0
        return torch.from_numpy(h).cuda()
    a, b = crop(a), crop(b)
kw = dict(max_iterations=iters, tolerance=0.0, partial_set_size=partial, pairs_per_chunk=pairs, n_streams=1)
ref = None
for cfg in configs.split(","):
    parts = cfg.split(":")
    impl, group, factor = parts[:3]
    mult, maxfrac = (parts[3], parts[4]) if len(parts) >= 5 else ("8", "0.25")
    cache = parts[5] if len(parts) >= 6 else "12"
    assert L.r3d_set_option(9, float(parts[6]) if len(parts) >= 7 else 0.0) == 0
    assert L.r3d_set_option(7, float(cache)) == 0
    assert L.r3d_set_option(4, float(mult)) == 0
    assert L.r3d_set_option(5, float(maxfrac)) == 0
    assert L.r3d_set_option(3, float(impl)) == 0
    assert L.r3d_set_option(2, float(factor)) == 0
    out = pipeline.register_depth_pairs(a, b, **kw)
    torch.cuda.synchronize()
    stats_on = L.r3d_nn_stats_read(None, 1) == 0      # only the diagnostic build (R3D_LIB_PATH=..._stats.so) has counters
    L.r3d_nn_timing_enable(1)
    t0 = time.perf_counter()
    out = pipeline.register_depth_pairs(a, b, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ms, n, q = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
    L.r3d_nn_timing_read(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(q), 1)
    L.r3d_nn_timing_enable(0)
    T = out["T"].cpu()
    if ref is None:
        ref = T
    dev = float((T - ref).abs().max())
    nq = int(out["n_src"].sum())
    print("clamp %s | impl %s G %2s factor %s slack %sx<=%s group<=%s: chunk %.3f ms  %.1f pairs/s | nn kernel %.3f ms over %d launches, avg %.1f us, "
          "%.2f G corr/s in kernel | max|T - T0| %.2e err0 %.6f" % (
              parts[6] if len(parts) >= 7 else "0", impl, group, factor, mult, maxfrac, cache, dt * 1e3, pairs / dt, ms.value, n.value, 1e3 * ms.value / max(n.value, 1),
              nq * n.value / max(ms.value, 1e-9) / 1e6, dev, float(out["mean_err"][0])), flush=True)
    if stats_on:
        sx = (ctypes.c_uint64 * 32)()
        L.r3d_nn_stats_read_ext(sx, 1)
        st = list(sx)[:8]
        tasks = max(int(st[0]), 1)
        ev = st[2] / tasks
        print("    per task of 32 queries: rounds %.2f, candidates/lane %.1f, lanes skipped %.2f, lookup rounds / staging passes %.2f, "
              "segments / block slots %.1f, non-empty / staged %.1f | violated skip bounds: %d" % (
                  st[1] / tasks, ev, st[3] / tasks, st[4] / tasks, st[5] / tasks, st[6] / tasks, st[7]), flush=True)
        print("    group search: %.3f of the tasks, %.2f passes and %.2f candidate steps per task, %.2f queries per task sent on to the union walk" % (
            sx[16] / tasks, sx[17] / tasks, sx[19] / tasks, sx[18] / tasks), flush=True)
        assert st[7] == 0, "a skipped lane saw a point closer than its seed"
    assert dev < 1e-9, "variants disagree"
