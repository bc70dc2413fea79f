"""Per-kernel measurements for every row of SURVEY.md section 8 (secondary to bench.py's headline line).

    python tools/bench_kernels.py [section ...] > gpurun_out/kernels.json        sections: k1 nn_once icp fit k4b k4a k5 k7 configs3
    bench.py imports run() for the secondary keys of its JSON line (k1, icp at partial_set_size 0.7, k4b, configs3).

Each entry: what was run, CUDA-event time (median of 5 after 2 warm-ups, inputs larger than L2 or the L2
flushed in between), the algorithmic work, and the fraction of the bound named in DESIGN.md section 3.
Calls go through the C ABI directly with pre-allocated buffers, so the timed region is the library call only.
"""
import ctypes
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def run(sections=None, icp_partials=(1.0, 0.7)):
    """Run the named sections (all by default) and return the dict of measurements."""
    want = set(sections or SECTIONS)
    need_pairs = bool(want & {"nn_once", "icp"})
    C = importlib.import_module("2dto3dreconstruction_b200._capi")
    ops = importlib.import_module("2dto3dreconstruction_b200.ops")
    L = C.lib()
    C.require_device()
    dev = torch.device("cuda", 0)
    HBM, HBM_SRC = bench.peak_hbm()
    FP64_FLOPS = 148 * 64 * 2 * 1.965e9          # nominal non-tensor fp64, DESIGN.md section 3
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, reps=5, warm=2, flush=True):
        reps, warm = int(os.environ.get("BK_REPS", reps)), int(os.environ.get("BK_WARM", warm))      # 1 / 0 under ncu
        ts = []
        for k in range(warm + reps):
            if flush:
                flush_buf.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if k >= warm:
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    out = {"hbm_peak_gbs": HBM, "hbm_peak_source": HBM_SRC, "fp64_peak_tflops_nominal": FP64_FLOPS / 1e12}

    st = C.stream_ptr()
    g = torch.Generator(device=dev).manual_seed(1)
    tms, tn, tq = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()

    # ---- K1 back-projection: 64 frames 640x480 uint16 -> xyzw fp64 ---------------------------------------------
    if "k1" in want:
        F = 64
        a, _ = bench.synth_batch_torch(F, 0, dev)
        pts = torch.empty((F * 480 * 640, 4), dtype=torch.float64, device=dev)
        off = torch.empty(F + 1, dtype=torch.int32, device=dev)
        nws = L.r3d_backproject_workspace_bytes(F, 480, 640)
        ws = C.workspace(nws)

        def k1():
            C.check(L.r3d_backproject(C.ptr(a), 0, F, 480, 640, 1, 525.0, 525.0, 319.5, 239.5, 1.0, 1, C.ptr(pts), C.ptr(off), C.ptr(ws),
                                      nws, st))

        ms = timed(k1)
        nvalid = int(off[-1].item())
        algo = 2 * F * 480 * 640 + 32 * nvalid
        # K1 writes 15x what it reads, and HBM writes are slower than the read + write mix of a copy: the write-only rate of
        # the device (a fill of as many bytes as K1 stores) is the roofline that applies
        fill = torch.empty(32 * nvalid, dtype=torch.uint8, device=dev)
        ms_fill = timed(lambda: fill.zero_())
        out["K1_backproject"] = {"what": "%d frames 640x480 u16 -> xyzw fp64, pinhole (one pass: ticketed 4096-pixel tiles, chained scan with "
                                         "decoupled look-back)" % F, "ms": ms,
                                 "frames_per_s": F / ms * 1e3, "algorithmic_bytes": algo, "achieved_gbs": algo / ms / 1e6,
                                 "frac_of_hbm": algo / ms / 1e6 / HBM, "stored_gbs": 32 * nvalid / ms / 1e6,
                                 "write_only_peak_gbs": 32 * nvalid / ms_fill / 1e6, "frac_of_write_only_peak": ms_fill / ms,
                                 "note": "bound = HBM WRITE bandwidth: a fill (torch zero_) of the same number of bytes takes %.3f ms; "
                                         "frac_of_hbm is against the measured copy peak (reads + writes)" % ms_fill}

    # ---- grid build + one unseeded query (r3d_nn_batch) on 8 pairs ----------------------------------------------
    if need_pairs:
        P = 8
        fa, fb = bench.synth_batch_torch(P, 0, dev)
        sp = torch.empty((P * 480 * 640, 4), dtype=torch.float64, device=dev)
        dp = torch.empty((P * 480 * 640, 4), dtype=torch.float64, device=dev)
        so = torch.empty(P + 1, dtype=torch.int32, device=dev)
        do = torch.empty(P + 1, dtype=torch.int32, device=dev)
        nws2 = L.r3d_backproject_workspace_bytes(P, 480, 640)
        ws2 = C.workspace(nws2)
        C.check(L.r3d_backproject(C.ptr(fa), 0, P, 480, 640, 1, 525.0, 525.0, 319.5, 239.5, 1.0, 1, C.ptr(sp), C.ptr(so), C.ptr(ws2), nws2, st))
        C.check(L.r3d_backproject(C.ptr(fb), 0, P, 480, 640, 1, 525.0, 525.0, 319.5, 239.5, 1.0, 1, C.ptr(dp), C.ptr(do), C.ptr(ws2), nws2, st))
        torch.cuda.synchronize()
        nq = int(so[-1].item())
        hw = 480 * 640
        nnws = L.r3d_icp_workspace_bytes(P, hw, hw)
        wsn = C.workspace(nnws)
        idx = torch.empty(nq, dtype=torch.int32, device=dev)
        dist = torch.empty(nq, dtype=torch.float64, device=dev)

        def nn_once():
            C.check(L.r3d_nn_batch(C.ptr(sp), C.ptr(so), C.ptr(dp), C.ptr(do), P, hw, hw, 0.0, C.ptr(idx), C.ptr(dist), C.ptr(wsn), nnws, st))

        L.r3d_nn_timing_enable(1)
        ms = timed(nn_once)
        L.r3d_nn_timing_read(ctypes.byref(tms), ctypes.byref(tn), ctypes.byref(tq), 1)
        L.r3d_nn_timing_enable(0)
        q_ms = tms.value / max(tn.value, 1)
        out["K2_nn_batch_unseeded"] = {"what": "nearest_neighbor for 8 pairs x ~276k: grid build + source sort + one unseeded query + export",
                                       "ms": ms, "query_kernel_ms": q_ms, "build_sort_export_ms": ms - q_ms,
                                       "correspondences_per_s": nq / ms * 1e3, "query_correspondences_per_s": nq / q_ms * 1e3}

    # ---- ICP loop: 30 iterations on the same 8 pairs --------------------------------------------------------------
    if "icp" in want:
        T = torch.empty((P, 16), dtype=torch.float64, device=dev)
        it = torch.empty(P, dtype=torch.int32, device=dev)
        err = torch.empty(P, dtype=torch.float64, device=dev)
        for partial in icp_partials:
            prm = C.IcpParams(30, 0, 0.0, partial, 0.0)

            def icp30():
                C.check(L.r3d_icp_batch(C.ptr(sp), C.ptr(so), C.ptr(dp), C.ptr(do), P, hw, hw, None, ctypes.byref(prm), C.ptr(T), C.ptr(it),
                                        C.ptr(err), None, None, None, C.ptr(wsn), nnws, st))

            L.r3d_nn_timing_enable(1)
            ms = timed(icp30)
            L.r3d_nn_timing_read(ctypes.byref(tms), ctypes.byref(tn), ctypes.byref(tq), 1)
            L.r3d_nn_timing_enable(0)
            out["K2K3_icp_30it_partial_%.1f" % partial] = {
                "what": "r3d_icp_batch, 8 pairs x ~276k, 30 iterations, tolerance 0, partial_set_size %.1f" % partial, "ms": ms,
                "pairs_per_s": P / ms * 1e3, "correspondences_per_s": 30 * nq / ms * 1e3,
                "nn_kernel_ms_per_launch": tms.value / max(tn.value, 1), "nn_kernel_share": tms.value / (ms * 7),
                "mean_err": float(err.mean().item())}

    # ---- K3 Kabsch / K3c affine on 2 M given correspondences -------------------------------------------------------
    if "fit" in want:
        n = 2_000_000
        A = torch.randn((n, 3), dtype=torch.float64, device=dev, generator=g) * 500
        B = A + torch.randn((n, 3), dtype=torch.float64, device=dev, generator=g)
        kws = L.r3d_kabsch_workspace_bytes(n)
        wsk = C.workspace(kws)
        T1 = torch.empty(16, dtype=torch.float64, device=dev)

        def kab():
            C.check(L.r3d_kabsch(C.ptr(A), 3, 1, C.ptr(B), 3, 1, n, C.ptr(T1), C.ptr(wsk), kws, st))

        ms = timed(kab)
        out["K3_kabsch"] = {"what": "best_fit_transform on 2 M pairs (packed (n,3) fp64)", "ms": ms, "algorithmic_bytes": 48 * n,
                            "achieved_gbs": 48 * n / ms / 1e6, "frac_of_hbm": 48 * n / ms / 1e6 / HBM}
        aws = L.r3d_affine_fit_workspace_bytes(n)
        wsa = C.workspace(aws)
        stat = torch.empty(1, dtype=torch.int32, device=dev)

        def aff():
            C.check(L.r3d_affine_fit(C.ptr(A), C.ptr(B), n, C.ptr(T1), C.ptr(stat), C.ptr(wsa), aws, st))

        ms = timed(aff)
        out["K3c_affine_fit"] = {"what": "homo_rigid_transform_3d on 2 M pairs", "ms": ms, "algorithmic_bytes": 48 * n,
                                 "achieved_gbs": 48 * n / ms / 1e6, "frac_of_hbm": 48 * n / ms / 1e6 / HBM}

    # ---- K4B: 100k hypotheses x 50k correspondences (configs[4]), inputs of SURVEY.md 8(d) C5 ------------------------------
    if "k4b" in want:
        nh, npnt = 100_000, 50_000
        rng5 = np.random.default_rng(11)
        P5 = rng5.uniform(-1000, 1000, (npnt, 3)).astype(np.float32).astype(np.float64)
        ang5 = np.deg2rad(2.0)
        R5 = np.array([[np.cos(ang5), -np.sin(ang5), 0], [np.sin(ang5), np.cos(ang5), 0], [0, 0, 1]])
        Q5 = P5 @ R5.T + np.array([15.0, -8.0, 5.0]) + rng5.normal(size=(npnt, 3))
        outl = rng5.random(npnt) < 0.4
        Q5[outl] = rng5.uniform(-1000, 1000, (int(outl.sum()), 3))
        Pp = ops.pack_xyzw(torch.from_numpy(P5).to(dev))
        Qp = ops.pack_xyzw(torch.from_numpy(Q5).to(dev))
        rs5 = np.random.RandomState(12345)
        entry = {"what": "SURVEY 8(d) C5: 50k correspondences (60 % inliers of a rigid motion + N(0,1), 40 % outliers), 100k hypotheses from "
                         "random minimal samples generated on the device (K6), threshold 10", "fp64_instructions_per_evaluation": 16}
        for kind, k in (("rigid", 3), ("affine", 4)):
            samples = torch.from_numpy(rs5.randint(npnt, size=(nh, k)).astype(np.int32)).to(dev)
            Hy, status = ops.hypotheses_3d(Pp, Qp, samples, kind)
            Hy = Hy.contiguous()
            cnt = torch.empty(nh, dtype=torch.int32, device=dev)

            def k4b():
                C.check(L.r3d_ransac3d_score(C.ptr(Hy), nh, C.ptr(Pp), C.ptr(Qp), npnt, 10.0, C.ptr(cnt), st))

            ms = timed(k4b, flush=False)
            ref_counts = cnt.clone()
            L.r3d_set_option(8, 1.0)
            ms32 = timed(k4b, flush=False)
            L.r3d_set_option(8, 0.0)
            ev = nh * npnt
            entry[kind] = {"ms": ms, "evaluations_per_s": ev / ms * 1e3, "frac_of_fp64_pipe": 16 * ev / 32 / ms * 1e3 / (592 * 1.965e9 / 2),
                           "ms_fp32_filter_kernel": ms32, "evaluations_per_s_fp32_filter_kernel": ev / ms32 * 1e3,
                           "counts_equal": bool(torch.equal(ref_counts, cnt)), "best_inlier_count": int(ref_counts.max().item()),
                           "degenerate_samples": int((status != 0).sum().item())}
        entry["ms"] = entry["rigid"]["ms"]
        entry["evaluations_per_s"] = entry["rigid"]["evaluations_per_s"]
        entry["frac_of_fp64_pipe"] = entry["rigid"]["frac_of_fp64_pipe"]
        entry["note"] = ("default kernel: all fp64, 9 DFMA + 3 DADD + 3 (r2) + 1 compare per evaluation, pipe peak = 16 lanes/clk/SMSP x 592 SMSPs; "
                         "opt-in kernel (R3D_OPT_K4B_FP32): fp32 filter, two hypotheses per lane, the fp64 expression redone inside the rounding "
                         "band -- identical counts, measured slower (profiles/r02_k4b_fp32.md)")
        out["K4B_ransac3d_sweep"] = entry

    # ---- K4A: reference-exact 2-D scoring at its native size (100 hypotheses x 1100 keypoints) ---------------------
    if "k4a" in want:
        rng = np.random.default_rng(3)
        n1 = n2 = 1100
        kp1 = np.column_stack((rng.uniform(0, 479, n1), rng.uniform(0, 639, n1)))
        kp2 = kp1 + rng.normal(0, 1.0, kp1.shape)
        Hs = np.tile(np.eye(3), (100, 1, 1)) + rng.normal(0, 1e-4, (100, 3, 3))
        Hd = C.to_device(Hs.reshape(100, 9))
        k1d, k2d = C.to_device(kp1), C.to_device(kp2)
        c2 = torch.empty(100, dtype=torch.int32, device=dev)
        fl = torch.empty(100, dtype=torch.int32, device=dev)
        w2 = L.r3d_ransac2d_workspace_bytes(100, n1, n2, 480, 640)
        ws2d = C.workspace(w2)

        def k4a():
            C.check(L.r3d_ransac2d_score(C.ptr(Hd), None, 100, C.ptr(k1d), n1, C.ptr(k2d), n2, 480, 640, C.ptr(c2), None, None, C.ptr(fl),
                                         C.ptr(ws2d), w2, st))

        ms = timed(k4a, flush=False)
        out["K4A_ransac2d"] = {"what": "q9.ransac_loop scoring body: 100 hypotheses x 1100 keypoints, 480x640 bins", "ms": ms,
                               "hypotheses_per_s": 100 / ms * 1e3, "note": "the reference needs ~65 ms per hypothesis (SURVEY.md section 6)"}

    # ---- K5: descriptor matching at the kitchen size (1100 x 1200 SIFT descriptors), with and without the ratio test ------
    if "k5" in want:
        d1 = torch.randint(0, 256, (1100, 128), device=dev, generator=g).float()
        d2 = torch.randint(0, 256, (1200, 128), device=dev, generator=g).float()
        mws = L.r3d_match_workspace_bytes(1100, 1200)
        wsm = C.workspace(mws)
        tr = torch.empty(1100, dtype=torch.int32, device=dev)
        di = torch.empty(1100, dtype=torch.float32, device=dev)

        def k5():
            C.check(L.r3d_match_descriptors_ratio(C.ptr(d1), 1100, C.ptr(d2), 1200, 128, 1, 0.8, C.ptr(tr), C.ptr(di), C.ptr(wsm), mws, st))

        ms = timed(k5, flush=False)
        out["K5_match_descriptors"] = {"what": "1100 x 1200 x 128 float32, both directions + cross-check + ratio test", "ms": ms,
                                       "gflops": 2 * 3 * 1100 * 1200 * 128 / ms / 1e6}

    # ---- K7: voxel down-sampling of a 2 M point merged cloud (voxel 2) ------------------------------------------------------
    if "k7" in want:
        nv = 2_000_000
        u = torch.rand(nv, dtype=torch.float64, device=dev, generator=g) * 640
        v = torch.rand(nv, dtype=torch.float64, device=dev, generator=g) * 480
        cloud = torch.stack(((u - 319.5) * 3.8, (v - 239.5) * 3.8, 2000 + 600 * torch.sin(u / 97)), dim=1).contiguous()
        vws = L.r3d_voxel_downsample_workspace_bytes(nv)
        wsv = C.workspace(vws)
        vout = torch.empty_like(cloud)
        meta = torch.zeros(2, dtype=torch.int32, device=dev)

        def k7():
            C.check(L.r3d_voxel_downsample(C.ptr(cloud), nv, 2.0, C.ptr(vout), C.ptr(meta[0:]), C.ptr(meta[1:]), C.ptr(wsv), vws, st))

        ms = timed(k7)
        nvox = int(meta[0].item())
        passes = 6          # 2 radix passes per axis at this extent
        algo = nv * (24 + 12) + passes * nv * 16 + nv * (4 + 12 + 24) + nvox * 24
        out["K7_voxel_downsample"] = {"what": "2 M points, voxel 2 -> %d voxels (bbox, coords, 3 stable sorts, heads, scan, reduce)" % nvox,
                                      "ms": ms, "points_per_s": nv / ms * 1e3, "algorithmic_bytes": algo, "achieved_gbs": algo / ms / 1e6,
                                      "frac_of_hbm": algo / ms / 1e6 / HBM}

    # ---- configs[3]: one 2 M x 2 M pair, 30 iterations (single GPU; SURVEY 8e: replicas only) --------------------------------
    if "configs3" in want:
        # SURVEY 8d C4 built on the device (same surface model as tests/test_gpu_full_size.py, torch RNG): a 1632x1226 depth
        # frame back-projected with the _ld intrinsics, dst = rigid motion (2 degrees about z, t = (15,-8,5)) + N(0, 0.5)
        hh, ww = 1226, 1632
        uu = torch.arange(ww, dtype=torch.float64, device=dev)[None, :].expand(hh, ww)
        vv = torch.arange(hh, dtype=torch.float64, device=dev)[:, None].expand(hh, ww)
        dd = torch.round(2000 + 600 * torch.sin(uu / 97) + 400 * torch.cos(vv / 61) + 2 * torch.randn((hh, ww), dtype=torch.float64, device=dev, generator=g))
        keep = torch.rand((hh, ww), device=dev, generator=g) >= 0.1
        dd, uu, vv = dd[keep], uu[keep], vv[keep]
        A2 = torch.stack(((uu - 319.5) * dd / 525.0, (vv - 239.5) * dd / 525.0, dd), dim=1).contiguous()
        ang = np.deg2rad(2.0)
        Rz = torch.tensor([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=torch.float64, device=dev)
        B2 = (A2 @ Rz.T + torch.tensor([15.0, -8.0, 5.0], dtype=torch.float64, device=dev)
              + 0.5 * torch.randn(A2.shape, dtype=torch.float64, device=dev, generator=g)).contiguous()
        a4 = ops.pack_xyzw(A2)
        b4 = ops.pack_xyzw(B2)
        oa = torch.tensor([0, len(A2)], dtype=torch.int32, device=dev)
        ob = torch.tensor([0, len(B2)], dtype=torch.int32, device=dev)
        bws = L.r3d_icp_workspace_bytes(1, len(A2), len(B2))
        wsb = C.workspace(bws)
        prm = C.IcpParams(30, 0, 0.0, 1.0, 0.0)
        Tb = torch.empty((1, 16), dtype=torch.float64, device=dev)
        itb = torch.empty(1, dtype=torch.int32, device=dev)
        eb = torch.empty(1, dtype=torch.float64, device=dev)

        def big_icp():
            C.check(L.r3d_icp_batch(C.ptr(a4), C.ptr(oa), C.ptr(b4), C.ptr(ob), 1, len(A2), len(B2), None, ctypes.byref(prm), C.ptr(Tb),
                                    C.ptr(itb), C.ptr(eb), None, None, None, C.ptr(wsb), bws, st))

        L.r3d_nn_timing_enable(1)
        ms = timed(big_icp)
        L.r3d_nn_timing_read(ctypes.byref(tms), ctypes.byref(tn), ctypes.byref(tq), 1)
        L.r3d_nn_timing_enable(0)
        out["configs3_single_pair_2M"] = {"what": "r3d_icp_batch on one %d x %d pair, 30 iterations, tolerance 0 (grid build + source sort + loop)" % (len(A2), len(B2)),
                                          "ms": ms, "correspondences_per_s": 30 * len(A2) / ms * 1e3,
                                          "nn_kernel_ms_per_launch": tms.value / max(tn.value, 1), "mean_err": float(eb.item())}

    return out


SECTIONS = ["k1", "nn_once", "icp", "fit", "k4b", "k4a", "k5", "k7", "configs3"]

if __name__ == "__main__":
    print(json.dumps(run(sys.argv[1:] or None), indent=1))
