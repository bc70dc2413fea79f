"""Short driver for ncu: registers one chunk of synthetic pairs (warm-up + one measured call).

    python tools/profile_icp.py [pairs] [partial] [iters] [group] [cell factor] [impl]
"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 8
partial = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 30
group = int(sys.argv[4]) if len(sys.argv) > 4 else 32
factor = float(sys.argv[5]) if len(sys.argv) > 5 else 2.0
impl = int(sys.argv[6]) if len(sys.argv) > 6 else 1
warm = int(os.environ.get("WARM", "1"))

import bench  # noqa: E402

pipeline = importlib.import_module("2dto3dreconstruction_b200.pipeline")
C = importlib.import_module("2dto3dreconstruction_b200._capi")
C.lib().r3d_set_option(2, factor)
C.lib().r3d_set_option(3, float(impl))
if os.environ.get("SLACK"):      # "mult,maxfrac"
    mult, frac = (float(x) for x in os.environ["SLACK"].split(","))
    C.lib().r3d_set_option(4, mult), C.lib().r3d_set_option(5, frac)
C.lib().r3d_set_option(7, float(os.environ.get("GROUP", "12")))
a, b = bench.synth_batch_torch(pairs, 0, torch.device("cuda", 0))
kw = dict(max_iterations=iters, tolerance=0.0, partial_set_size=partial, pairs_per_chunk=pairs, n_streams=1)
l0 = C.launch_count()
for _ in range(warm):
    out = pipeline.register_depth_pairs(a, b, **kw)
torch.cuda.synchronize()
l1 = C.launch_count()
t0 = time.perf_counter()
out = pipeline.register_depth_pairs(a, b, **kw)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("impl", impl, "group", group, "factor", factor, "launches per call:", (l1 - l0) // max(warm, 1), "time %.3f ms" % (dt * 1e3),
      "pairs/s %.1f" % (pairs / dt), "n_src", out["n_src"][:2].tolist(), "err", out["mean_err"][:2].tolist())
