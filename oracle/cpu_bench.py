"""CPU timing of the oracle port (the reference's algorithm on host cores).  TEST/BENCH
INFRASTRUCTURE ONLY -- used by bench.py's `cpu_baseline` leg and by `bench.py --impl reference`.

One worker process per core; each registers one full-density depth pair the way the reference
does (utils/utils.py:48-76 back-projection of both frames, then utils/icp.py:74-133 with the
scikit-learn KD-tree it rebuilds every iteration).  A full 30-iteration pair costs ~14 s per core;
with sample_iters < full_iters the timed sample runs `sample_iters` ICP iterations per pair and the
per-pair time is scaled to `full_iters` (per-iteration cost is constant: one tree build + one query
+ one Kabsch; bench.py --impl reference measured scaled / full = 0.988).
"""
import os
import time


def _one_pair(args):
    import numpy as np  # noqa: F401  (spawned worker: import here)
    from oracle import ref_port as O
    frame_a, frame_b, sample_iters, partial = args
    t0 = time.perf_counter()
    A = O.depth_to_voxel_ld(frame_a)
    B = O.depth_to_voxel_ld(frame_b)
    t1 = time.perf_counter()
    T, dist, iters = O.icp(A, B, max_iterations=sample_iters, tolerance=0.0, partial_set_size=partial)
    t2 = time.perf_counter()
    return {"t_backproject": t1 - t0, "t_icp": t2 - t1, "n_src": int(A.shape[0]), "T": T, "iters": iters,
            "extent": O.scene_extent(B)}


def time_pairs(frames_a, frames_b, sample_iters=2, full_iters=30, partial=1.0, cores=None):
    """frames_a/b: uint16 [P,H,W] host arrays, P >= number of worker processes wanted.
    Returns dict with pairs/s and correspondences/s extrapolated to `full_iters` iterations."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    n = min(cores, len(frames_a))
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:
        os.environ[k] = "1"          # one core per worker: the reference is single-threaded per pair
    try:
        ctx = mp.get_context("spawn")
        with ctx.Pool(n) as pool:
            pool.map(_noop, range(n))                      # start-up (imports) outside the timed region
            t0 = time.perf_counter()
            res = pool.map(_one_pair, [(frames_a[i], frames_b[i], sample_iters, partial) for i in range(n)])
            wall = time.perf_counter() - t0
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    # per-pair time at full_iters: back-projection once, ICP scaled by iteration count
    per_pair = [r["t_backproject"] + r["t_icp"] * (full_iters / float(sample_iters)) for r in res]
    mean_pair_s = sum(per_pair) / len(per_pair)
    pairs_per_s = n / mean_pair_s                       # n independent workers in parallel
    corr = sum(r["n_src"] for r in res) / len(res) * full_iters
    return {
        "pairs_per_s": pairs_per_s,
        "corr_per_s": pairs_per_s * corr,
        "cores": n,
        "wall_s": wall,
        "sample": ("%d full-density 640x480 pairs, one per core, %d of %d ICP iterations each (KD-tree rebuilt per "
                   "iteration as in utils/icp.py:68-69)" % (n, sample_iters, full_iters))
                  + (", per-pair time scaled to %d iterations" % full_iters if sample_iters != full_iters else ", nothing scaled"),
        "results": res,
    }


def _noop(_):
    import numpy  # noqa: F401
    import sklearn.neighbors  # noqa: F401
    from oracle import ref_port  # noqa: F401
    return 0
