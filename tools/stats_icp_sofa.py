"""Search counters of the nn kernel on the sofa pair of tests/golden (real, non-uniform data).
R3D_LIB_PATH=.../libr3d_b200_stats.so python tools/stats_icp_sofa.py [cell factor] [iterations] [partial]"""
import ctypes
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

C = importlib.import_module("2dto3dreconstruction_b200._capi")
icp = importlib.import_module("2dto3dreconstruction_b200.utils.icp")
L = C.lib()
factor = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
partial = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
L.r3d_set_option(2, factor)
g = np.load(os.path.join(ROOT, "tests", "golden", "sofa_pair.npz"))
A, B = g["icp_A"], g["icp_B"]
print("A", A.shape, "extent", np.ptp(A, axis=0), "B extent", np.ptp(B, axis=0))
icp.icp(A, B, tolerance=0.0, max_iterations=3, partial_set_size=partial)
torch.cuda.synchronize()
has_stats = L.r3d_nn_stats_read(None, 1) == 0
L.r3d_nn_timing_enable(1)
t0 = time.perf_counter()
T, d, it = icp.icp(A, B, tolerance=0.0, max_iterations=iters, partial_set_size=partial)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
ms, n, q = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
L.r3d_nn_timing_read(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(q), 1)
L.r3d_nn_timing_enable(0)
print("factor %.2f: icp %.2f ms for %d iterations; nn kernel %.2f ms over %d launches (%.1f us each); mean dist %.3f" % (
    factor, dt * 1e3, iters, ms.value, n.value, 1e3 * ms.value / max(n.value, 1), d.mean()))
if has_stats:
    st = (ctypes.c_uint64 * 32)()
    L.r3d_nn_stats_read_ext(st, 1)
    tasks = max(int(st[0]), 1)
    print("per task: rounds %.2f, candidates/lane %.1f, lanes skipped %.2f, lookup rounds %.2f, segments %.1f, non-empty %.1f" % (
        st[1] / tasks, st[2] / tasks, st[3] / tasks, st[4] / tasks, st[5] / tasks, st[6] / tasks))
    print("far searches/task %.3f, sweep steps/far %.1f, scan steps/far %.1f, block-granular walks/task %.3f | task cycles: mean %.0f, max %d" % (
        st[8] / tasks, st[9] / max(int(st[8]), 1), st[10] / max(int(st[8]), 1), st[11] / tasks, st[12] / tasks, st[13]))
    print("group search: tasks %.3f of all, passes/task %.2f, candidate steps/task %.2f, queries sent on to the union walk per task %.2f" % (
        st[16] / tasks, st[17] / tasks, st[19] / tasks, st[18] / tasks))
