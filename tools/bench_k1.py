"""K1 only (64 frames 640x480): python tools/bench_k1.py  -- the K1 block of tools/bench_kernels.py, for build-knob sweeps."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402

C = importlib.import_module("2dto3dreconstruction_b200._capi")
L = C.lib()
C.require_device()
dev = torch.device("cuda", 0)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
F = 64
a, _ = bench.synth_batch_torch(F, 0, dev)
pts = torch.empty((F * 480 * 640, 4), dtype=torch.float64, device=dev)
off = torch.empty(F + 1, dtype=torch.int32, device=dev)
nws = L.r3d_backproject_workspace_bytes(F, 480, 640)
ws = C.workspace(nws)
st = C.stream_ptr()
ts = []
for k in range(9):
    flush_buf.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    C.check(L.r3d_backproject(C.ptr(a), 0, F, 480, 640, 1, 525.0, 525.0, 319.5, 239.5, 1.0, 1, C.ptr(pts), C.ptr(off), C.ptr(ws), nws, st))
    e1.record()
    torch.cuda.synchronize()
    if k >= 2:
        ts.append(e0.elapsed_time(e1))
nvalid = int(off[-1].item())
algo = 2 * F * 480 * 640 + 32 * nvalid
ms = float(np.median(ts))
print("K1 %s: %.4f ms, %.0f GB/s algorithmic" % (os.environ.get("R3D_LIB_PATH", "default").split("/")[-1], ms, algo / ms / 1e6))
