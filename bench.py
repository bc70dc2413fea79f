#!/usr/bin/env python
"""Benchmark of the registration hot path (BASELINE.json: ICP frame-pairs/s and correspondences/s
on synthetic 640x480 depth pairs; configs[2]: batch of 1024 pairs, back-projection + 30-iteration
ICP, sharded over 1/2/4/8 GPUs).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" registers every pair of the rank's batch once: K1 back-projection of both frames, grid
build, 30 ICP iterations (tolerance 0, partial_set_size 1.0), one all-gather of the 4x4s when N > 1.
Weak scaling: every GPU holds `--pairs` (default 1024) pairs; `value` counts all ranks' pairs.  The same
line also carries `strong` (ONE 1024-pair batch split over the N ranks), `secondary` (the other
configurations of BASELINE.json, measured in the same run) and `parity`: the first pairs of the batch are the
NumPy-generated frames of SURVEY.md section 8(d), the CPU arm registers exactly those, and the two sets of
transforms are compared with north_star's tolerances -- the program exits non-zero when they differ.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 480, 640
ICP_ITERS = 30
# 32 B source point + 32 B destination point (sorted cloud read once per launch, M ~ N)
# + 4 B previous neighbour position read + 4 B new position written (SURVEY.md 8(d)'s 40 B is for float4 points)
ALGO_BYTES_PER_CORR = 72.0
FALLBACK_HBM_GBS = 6650.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("R3D_BENCH_PAIRS", "1024")), help="pairs per GPU")
    ap.add_argument("--pairs-per-chunk", type=int, default=int(os.environ.get("R3D_BENCH_PPC", "32")))
    ap.add_argument("--streams", type=int, default=int(os.environ.get("R3D_BENCH_STREAMS", "3")))
    ap.add_argument("--partial", type=float, default=1.0)
    ap.add_argument("--cpu-sample-iters", type=int, default=None,
                    help="ICP iterations per pair of the bounded CPU sample (~0.45 s per iteration per core); default: 16 for the "
                         "cpu_baseline leg of the CUDA arm, all 30 for --impl reference (no extrapolation)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="also skips the parity gate (it needs the CPU transforms)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configurations (partial 0.7, configs[3], configs[4], K1)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: --pairs is the GLOBAL batch, split over the ranks, and is what `value` reports")
    return ap.parse_args()


def make_config(args, world):
    """The `config` object of the JSON line -- the same for both arms (the reference arm registers a bounded sample of it)."""
    strong = args.scaling == "strong"
    per_gpu = args.pairs // world if strong else args.pairs
    return {"workload": "synthetic 640x480 depth-pair batch: pinhole back-projection + %d-iteration ICP (tolerance 0, "
                        "partial_set_size %.2f), BASELINE.json configs[2]" % (ICP_ITERS, args.partial),
            "pairs_per_gpu": per_gpu, "global_pairs": per_gpu * world, "icp_iterations": ICP_ITERS,
            "inputs": "SURVEY.md 8(d) surface model; pair p < host cores (max 64) of rank 0 = np.random.default_rng(1234 + p) "
                      "(2dto3dreconstruction_b200/synth.py), the frames both arms register and the parity gate compares",
            "pairs_per_chunk": args.pairs_per_chunk, "streams": args.streams,
            "parallelism": "pairs sharded x%d, one all-gather of 4x4s" % world,
            "l2": "inputs %.2f GB per step per GPU (> 126 MB L2); no flush needed" % (2 * per_gpu * H * W * 2 / 1e9)}


def n_shared_pairs(limit):
    return max(1, min(os.cpu_count() or 1, 64, limit))


def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def synth_batch_torch(n_pairs, seed, device):
    """The SURVEY.md section 8(d) surface model drawn on the device (2dto3dreconstruction_b200/synth.py)."""
    import importlib
    return importlib.import_module("2dto3dreconstruction_b200.synth").depth_batch_device(n_pairs, seed, device)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu), "-f", self.path], stdout=subprocess.DEVNULL,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
            os.unlink(self.path)
            sm = sorted(float(r[1]) for r in rows if len(r) >= 9)
            if sm:
                out["sm_mhz"] = sm[len(sm) // 2]
                out["sm_max_mhz"] = float(rows[0][2])
                out["samples"] = len(sm)
                out["power_w_max"] = max(float(r[3]) for r in rows if len(r) >= 9)
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for k, nm in enumerate(names):
                    if any("Active" == r[5 + k].strip() for r in rows if len(r) >= 9):
                        out["reasons"].append(nm)
        except Exception:
            pass
        return out


def transform_gap(T, T_ref, extent):
    """north_star's two numbers: Frobenius norm of the rotation difference, translation difference / scene extent."""
    import numpy as np
    T, T_ref = np.asarray(T), np.asarray(T_ref)
    return float(np.linalg.norm(T[:3, :3] - T_ref[:3, :3])), float(np.linalg.norm(T[:3, 3] - T_ref[:3, 3]) / extent)


def run_reference(args):
    """The reference's own CPU path (oracle port: NumPy + scikit-learn KD-tree, as utils/icp.py
    calls them) on all host cores, same metric and config.  Rank 0 only.  A step registers one pair per core -- the
    very frames the CUDA arm's parity gate compares -- for all 30 iterations (~15 s per step).  With `--cpu-sample-iters k`
    a step runs k of the 30 iterations and `value` scales the per-pair time; one untimed full pass then validates the
    scaling (`full_run`; measured 2026-10: scaled / full = 0.988)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import importlib
    from oracle import cpu_bench
    synth = importlib.import_module("2dto3dreconstruction_b200.synth")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n = n_shared_pairs(args.pairs)
    fa, fb = synth.depth_pairs(0, n)
    full = None
    if args.cpu_sample_iters != ICP_ITERS:
        full = cpu_bench.time_pairs(fa, fb, sample_iters=ICP_ITERS, full_iters=ICP_ITERS, partial=args.partial, cores=n)
    vals, corr, walls = [], [], []
    for s in range(args.warmup + args.steps):
        r = cpu_bench.time_pairs(fa, fb, sample_iters=args.cpu_sample_iters, full_iters=ICP_ITERS, partial=args.partial,
                                 cores=n)
        if s >= args.warmup:
            vals.append(r["pairs_per_s"])
            corr.append(r["corr_per_s"])
            walls.append(r["wall_s"])
    v = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "icp_frame_pairs_per_sec", "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sum(walls) / len(walls), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "correspondences_per_sec": sum(corr) / len(corr),
        "config": make_config(args, world),
        "ms_per_step_note": "measured wall time of one step = %d pairs x %d of %d iterations; `value` = pairs / (per-pair time scaled "
                            "to %d iterations)" % (n, args.cpu_sample_iters, ICP_ITERS, ICP_ITERS),
        "full_run": None if full is None else {
            "what": "the same %d pairs registered once for all %d iterations (untimed by the step clock)" % (n, ICP_ITERS),
            "pairs_per_s": full["pairs_per_s"], "wall_s": full["wall_s"], "scaled_over_full": v / full["pairs_per_s"]},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        if args.cpu_sample_iters is None:
            args.cpu_sample_iters = ICP_ITERS
        run_reference(args)
        return
    if args.cpu_sample_iters is None:
        args.cpu_sample_iters = 16
    if args.warmup < 3:
        args.warmup = 3

    import importlib
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its "NCCL version ..." / NCCL_DEBUG lines to stdout by default: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    pkg = importlib.import_module("2dto3dreconstruction_b200")
    pipeline = importlib.import_module("2dto3dreconstruction_b200.pipeline")
    C = pkg._capi
    C.require_device()
    L = C.lib()

    synth = importlib.import_module("2dto3dreconstruction_b200.synth")
    strong = args.scaling == "strong"
    if strong and args.pairs % world:
        raise SystemExit("--scaling strong needs --pairs divisible by the number of ranks")
    P = args.pairs // world if strong else args.pairs
    depth_a, depth_b = synth_batch_torch(P, seed=rank, device=dev)
    n_shared = n_shared_pairs(P) if rank == 0 else 0
    if n_shared:
        # the pairs both arms register (SURVEY.md 8(d) NumPy generator); the rest of the batch is drawn on the device
        np_a, np_b = synth.depth_pairs(0, n_shared)
        depth_a[:n_shared] = torch.from_numpy(np_a).to(dev)
        depth_b[:n_shared] = torch.from_numpy(np_b).to(dev)
    torch.cuda.synchronize()
    kw = dict(max_iterations=ICP_ITERS, tolerance=0.0, partial_set_size=args.partial,
              pairs_per_chunk=args.pairs_per_chunk, n_streams=args.streams)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        out = pipeline.register_depth_pairs(depth_a, depth_b, **kw)
        if world > 1:
            out["T_all"] = pipeline.gather_transforms(out["T"], P * world)
        return out

    # ---------------- device-resident timing (`value`) ----------------
    for _ in range(args.warmup):
        out = step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = C.launch_count()
    L.r3d_nn_timing_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step_device()
    e1.record()
    barrier()
    L.r3d_nn_timing_enable(0)
    ms_total = e0.elapsed_time(e1)
    launches = C.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    import ctypes
    nn_ms, nn_launches, nn_q = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
    C.check(L.r3d_nn_timing_read(ctypes.byref(nn_ms), ctypes.byref(nn_launches), ctypes.byref(nn_q), 1))

    # the same kernel timed alone on the GPU (one stream, one chunk in flight): the clean per-launch duration
    L.r3d_nn_timing_enable(1)
    solo_pairs = min(P, 2 * args.pairs_per_chunk)
    pipeline.register_depth_pairs(depth_a[:solo_pairs], depth_b[:solo_pairs], max_iterations=ICP_ITERS, tolerance=0.0,
                                  partial_set_size=args.partial, pairs_per_chunk=args.pairs_per_chunk, n_streams=1)
    torch.cuda.synchronize()
    L.r3d_nn_timing_enable(0)
    solo_ms, solo_launches = ctypes.c_double(), ctypes.c_int64()
    C.check(L.r3d_nn_timing_read(ctypes.byref(solo_ms), ctypes.byref(solo_launches), ctypes.byref(nn_q), 1))

    n_src = out["n_src"].cpu().numpy().astype(np.int64)
    iters = out["iters"].cpu().numpy().astype(np.int64)
    executed = np.minimum(iters, ICP_ITERS)               # iterations actually run per pair
    corr_per_step = int((n_src * executed).sum())

    # ---------------- end-to-end timing (`e2e`): host buffers in, host results out ----------------
    host_a = torch.empty((P, H, W), dtype=torch.uint16, pin_memory=True)
    host_b = torch.empty((P, H, W), dtype=torch.uint16, pin_memory=True)
    host_a.copy_(depth_a)
    host_b.copy_(depth_b)
    torch.cuda.synchronize()

    def step_host():
        r = pipeline.register_depth_pairs(host_a, host_b, **kw)       # synchronous: results are on the host
        if world > 1:
            r["T_all"] = pipeline.gather_transforms(torch.from_numpy(r["T"]).to(dev), P * world)
        return r

    for _ in range(2):
        rh = step_host()
    barrier()
    t0 = time.perf_counter()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        rh = step_host()
    g1.record()
    barrier()
    e2e_ms = g0.elapsed_time(g1)
    e2e_wall_ms = (time.perf_counter() - t0) * 1000.0
    parity_dev_host = bool(np.array_equal(rh["T"], out["T"].cpu().numpy()))

    # ---------------- strong scaling beside the weak line: ONE `--pairs` batch split over the ranks ----------------
    strong_ms = None
    if world > 1 and not strong and args.pairs % world == 0:
        Ps = args.pairs // world
        sa, sb = depth_a[:Ps], depth_b[:Ps]

        def step_strong():
            r = pipeline.register_depth_pairs(sa, sb, **kw)
            r["T_all"] = pipeline.gather_transforms(r["T"], Ps * world)
            return r

        for _ in range(3):
            step_strong()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            step_strong()
        s1.record()
        barrier()
        strong_ms = s0.elapsed_time(s1)

    # max over ranks
    if world > 1:
        tt = torch.tensor([ms_total, e2e_ms, strong_ms or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, strong_ms = float(tt[0]), float(tt[1]), (float(tt[2]) if strong_ms else None)
        cc = torch.tensor([corr_per_step, launches], dtype=torch.int64, device=dev)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        corr_total, launches_total = int(cc[0]), int(cc[1])
    else:
        corr_total, launches_total = corr_per_step, launches

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = ms_total / args.steps
    pairs_total = P * world
    value = pairs_total / (ms_per_step / 1000.0)
    e2e_value = pairs_total / (e2e_ms / args.steps / 1000.0)
    peak, peak_src = peak_hbm()
    nn_avg_ms = nn_ms.value / max(nn_launches.value, 1)
    # rank 0's nn launches processed rank 0's correspondences.  The kernel is bound by instruction issue, not by HBM
    # (DESIGN.md section 4): the primary fraction is warp instructions issued / issue slots, with the warp instructions per
    # correspondence taken from the committed ncu capture of all 30 launches of a chunk (profiles/nn_kernel_traffic.json);
    # the HBM figures (algorithmic bytes / measured copy peak) are kept beside it.
    corr_in_kernel_per_s = corr_per_step * args.steps / max(nn_ms.value / 1000.0, 1e-12)
    solo_corr_per_s = float(n_src[:solo_pairs].sum()) * ICP_ITERS / max(solo_ms.value / 1000.0, 1e-12)
    hbm_achieved = ALGO_BYTES_PER_CORR * corr_in_kernel_per_s / 1e9
    tf = {}
    try:
        tf = json.load(open(os.path.join(ROOT, "profiles", "nn_kernel_traffic.json")))
    except Exception:
        pass
    instr_per_corr = tf.get("warp_instructions_per_correspondence")
    sm_hz = 1.0e6 * ((clocks or {}).get("sm_mhz") or 1965.0)
    issue_peak = 592.0 * sm_hz      # one warp instruction per clock per SM sub-partition, 4 x 148 of them
    roofline = {
        "bound": "issue" if instr_per_corr else "hbm",
        "bound_note": "instruction issue slots of the 592 SM sub-partitions (non-tensor kernel: K = 3, fp64); the contract's "
                      "hbm figures are hbm_achieved / hbm_peak / hbm_frac",
        "kernel": "nn_query_kernel<1,1,1> (K2 exact grid NN: warp-union walk + group search, fused K3 moments)",
        "achieved": (instr_per_corr * solo_corr_per_s / 1e9) if instr_per_corr else hbm_achieved,
        "peak": issue_peak / 1e9 if instr_per_corr else peak,
        "unit": "G warp-instructions/s" if instr_per_corr else "GB/s",
        "traffic": None, "peak_source": "592 SMSPs x SM clock under load (%.0f MHz)" % (sm_hz / 1e6) if instr_per_corr else peak_src,
        "warp_instructions_per_correspondence": instr_per_corr,
        "hbm_achieved": hbm_achieved, "hbm_peak": peak, "hbm_frac": hbm_achieved / peak, "hbm_peak_source": peak_src,
        "hbm_achieved_solo": ALGO_BYTES_PER_CORR * solo_corr_per_s / 1e9,
        "algorithmic_bytes_per_correspondence": ALGO_BYTES_PER_CORR, "launches_timed": int(nn_launches.value),
        "avg_launch_ms": nn_avg_ms, "kernel_share_of_step": nn_ms.value / (ms_total * args.streams),
        "share_note": "sum of the kernel's event-bracketed time / (timed region x %d streams): launches of different streams "
                      "overlap, so in-step durations include waiting for the other streams' kernels" % args.streams,
        "avg_launch_ms_solo": solo_ms.value / max(solo_launches.value, 1),
        "correspondences_per_sec_in_kernel": corr_in_kernel_per_s,
        "correspondences_per_sec_in_kernel_solo": solo_corr_per_s,
        "note": "achieved = instructions per correspondence (ncu, whole chunk) x correspondences/s of the kernel timed alone "
                "(one stream); in the step itself three streams share the issue slots",
    }
    roofline["frac"] = roofline["achieved"] / roofline["peak"]
    if "dram_bytes_per_correspondence" in tf:
        corr_per_launch = corr_per_step * args.steps / max(int(nn_launches.value), 1)
        roofline["traffic"] = tf["dram_bytes_per_correspondence"] * corr_per_launch
        roofline["traffic_bytes_per_correspondence"] = tf["dram_bytes_per_correspondence"]
        roofline["traffic_source"] = tf.get("source")

    cpu_baseline, parity = None, None
    if not args.no_cpu_baseline:
        from oracle import cpu_bench
        n = n_shared
        r = cpu_bench.time_pairs(np_a, np_b, sample_iters=args.cpu_sample_iters, full_iters=ICP_ITERS, partial=args.partial, cores=n)
        cpu_baseline = {"value": r["pairs_per_s"], "unit": "pairs/s", "cores": r["cores"], "kind": "port",
                        "sample": r["sample"], "correspondences_per_sec": r["corr_per_s"], "wall_s": r["wall_s"]}
        # parity gate: the CUDA path on the very frames the CPU arm just registered, stopped after the
        # same number of iterations, against the CPU transforms; north_star: |dR|_F <= 1e-4, |dt| <= 1e-4 x scene extent
        g = pipeline.register_depth_pairs(depth_a[:n], depth_b[:n], max_iterations=args.cpu_sample_iters, tolerance=0.0,
                                          partial_set_size=args.partial, pairs_per_chunk=args.pairs_per_chunk, n_streams=1)
        Tg, itg, ng = g["T"].cpu().numpy(), g["iters"].cpu().numpy(), g["n_src"].cpu().numpy()
        gaps = [transform_gap(Tg[k], r["results"][k]["T"], r["results"][k]["extent"]) for k in range(n)]
        parity = {"pairs": n, "iterations": args.cpu_sample_iters, "max_dR": max(x[0] for x in gaps),
                  "max_dt_over_extent": max(x[1] for x in gaps), "tolerance": 1e-4,
                  "iters_equal": bool(all(int(itg[k]) == int(r["results"][k]["iters"]) for k in range(n))),
                  "cloud_sizes_equal": bool(all(int(ng[k]) == int(r["results"][k]["n_src"]) for k in range(n))),
                  "against": "oracle port (NumPy + scikit-learn KD-tree) on the same frames, same iteration count"}
        parity["ok"] = bool(parity["max_dR"] <= 1e-4 and parity["max_dt_over_extent"] <= 1e-4 and parity["iters_equal"]
                            and parity["cloud_sizes_equal"] and parity_dev_host)

    secondary = None
    if world == 1 and not args.no_secondary:
        # the other configurations of BASELINE.json, same run, same process (tools/bench_kernels.py: CUDA events, median of 5)
        del depth_a, depth_b, host_a, host_b
        pipeline._ws_cache.clear()
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_kernels
        k = bench_kernels.run(["k1", "icp", "k4b", "configs3"], icp_partials=(args.partial, 0.7) if args.partial != 0.7 else (0.7,))
        secondary = {
            "partial_set_size_0.7": k.get("K2K3_icp_30it_partial_0.7"),
            "partial_set_size_%.1f_same_8_pairs" % args.partial: k.get("K2K3_icp_30it_partial_%.1f" % args.partial),
            "configs[3]": k.get("configs3_single_pair_2M"),
            "configs[4]": k.get("K4B_ransac3d_sweep"),
            "K1_backproject": k.get("K1_backproject"),
        }

    line = {
        "metric": "icp_frame_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "correspondences_per_sec": corr_total / (ms_per_step / 1000.0),
        "config": make_config(args, world),
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * P * H * W * 2) * world,
                "d2h_bytes_per_step": int(P * (128 + 4 + 8 + 4 + 4)) * world, "ms_per_step": e2e_ms / args.steps,
                "wall_ms_per_step_rank0": e2e_wall_ms / args.steps,
                "api": "pipeline.register_depth_pairs(host uint16 frames) -> r3d_register_depth_pairs_host",
                "device_vs_host_results_identical": parity_dev_host},
        "gpu_launches": launches_total,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "parity": parity,
        "secondary": secondary,
    }
    if strong_ms is not None:
        line["strong"] = {"global_pairs": args.pairs, "pairs_per_gpu": args.pairs // world, "ms_per_step": strong_ms / args.steps,
                          "value": args.pairs / (strong_ms / args.steps / 1000.0), "unit": "pairs/s",
                          "note": "one %d-pair batch split over the %d ranks, device-resident, all-gather included, max over ranks" % (
                              args.pairs, world)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        print("PARITY GATE FAILED: %r" % (parity,), file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()
