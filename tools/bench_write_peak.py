"""Write-only and copy bandwidth of the device (for K1's roofline: K1 writes 15x what it reads).

    python tools/bench_write_peak.py
"""
import json

import torch

dev = torch.device("cuda", 0)
n = 600 << 20
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


t_fill = timed(lambda: a.zero_())
t_copy = timed(lambda: b.copy_(a))
print(json.dumps({"bytes": n, "fill_ms": t_fill, "write_only_gbs": n / t_fill / 1e6, "copy_ms": t_copy,
                  "copy_gbs_read_plus_write": 2 * n / t_copy / 1e6}))
