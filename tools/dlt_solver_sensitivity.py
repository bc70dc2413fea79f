"""Is ransac_loop's result sensitive to WHICH eigen-solver takes the DLT null vector?   

    python tools/dlt_solver_sensitivity.py [seeds] [eigh|jacobi] > gpurun_out/dlt_sensitivity.json

q9.py:40-50 takes the homography of a 4-point sample as the eigenvector of the smallest eigenvalue of the unnormalised 9x9 matrix
A^T A, from numpy.linalg.eig (LAPACK dgeev on a matrix that is symmetric only up to rounding).  A device DLT would use another
solver (a batched symmetric Jacobi), whose null vector differs from dgeev's in the low digits.  This tool measures what that does to
the OUTPUT of ransac_loop -- the inlier list of the best of 100 hypotheses after the refit -- on the reference's own keypoints
(tests/golden/sofa_pair.npz, kitchen_pair.npz), over many seeds of NumPy's global RNG: the loop of the drop-in
(utils/homography_utils/q9.py::ransac_loop: same draws, same scoring kernel K4A) run with numpy.linalg.eig, with numpy.linalg.eigh
on the symmetrised matrix, and with a cyclic Jacobi sweep in fp64 (what a kernel would do).
"""
import importlib
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

q9 = importlib.import_module("2dto3dreconstruction_b200.utils.homography_utils.q9")
seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
only = sys.argv[2] if len(sys.argv) > 2 else None      # "eigh" / "jacobi": that solver alone, on all the seeds


def solve_eig(AtA):
    w, v = np.linalg.eig(AtA)
    return v[:, np.argmin(w)].reshape(3, 3)


def solve_eigh(AtA):
    w, v = np.linalg.eigh(0.5 * (AtA + AtA.T))
    return v[:, 0].reshape(3, 3)


def solve_jacobi(AtA, sweeps=30):
    """Cyclic Jacobi eigenvalue iteration on the symmetrised matrix, fp64, rotations in the fixed (p, q) order a kernel would use."""
    a = 0.5 * (AtA + AtA.T)
    n = a.shape[0]
    v = np.eye(n)
    for _ in range(sweeps):
        off = 0.0
        for p in range(n - 1):
            for q in range(p + 1, n):
                apq = a[p, q]
                if apq == 0.0:
                    continue
                off = max(off, abs(apq) / (np.sqrt(abs(a[p, p] * a[q, q])) + 1e-300))
                theta = (a[q, q] - a[p, p]) / (2.0 * apq)
                t = (1.0 if theta >= 0 else -1.0) / (abs(theta) + np.sqrt(theta * theta + 1.0))
                c = 1.0 / np.sqrt(t * t + 1.0)
                s = t * c
                ap, aq = a[p, :].copy(), a[q, :].copy()
                a[p, :], a[q, :] = c * ap - s * aq, s * ap + c * aq
                ap, aq = a[:, p].copy(), a[:, q].copy()
                a[:, p], a[:, q] = c * ap - s * aq, s * ap + c * aq
                vp, vq = v[:, p].copy(), v[:, q].copy()
                v[:, p], v[:, q] = c * vp - s * vq, s * vp + c * vq
        if off < 1e-17:
            break
    return v[:, int(np.argmin(np.diag(a)))].reshape(3, 3)


def ransac(kp1, kp2, bf, shape, solve):
    """utils/homography_utils/q9.py::ransac_loop with a pluggable null-vector solver; returns (best count, final matches)."""
    orig = np.array([(p[1], p[0]) for p in kp1], dtype=float)
    prime = np.array([(p[1], p[0]) for p in kp2], dtype=float)
    hyps = []
    for _ in range(100):
        idx = np.random.randint(len(bf), size=4)
        base = [q9.Match(kp1[bf[i][0]][0], kp1[bf[i][0]][1], kp2[bf[i][1]][0], kp2[bf[i][1]][1]) for i in idx]
        A = q9.make_homography_LS_matrix(base)
        hyps.append(solve(A.T @ A))
    counts, mj, ms = q9._score(np.stack(hyps), orig, prime, shape, True)
    best = int(np.argmax(counts))
    n_in = int(counts[best])
    bm = q9._matches_from(mj[best], ms[best])[0] if n_in > 0 else []
    fin = [q9.Match(orig[bm[i][0]][1], orig[bm[i][0]][0], prime[bm[i][1]][1], prime[bm[i][1]][0]) for i in range(n_in)]
    A = q9.make_homography_LS_matrix(fin)
    H = solve(A.T @ A)
    _, mj, ms = q9._score(np.asarray(H)[None], orig, prime, shape, True)
    return n_in, best, q9._matches_from(mj[0], ms[0])[0]


out = {"seeds": seeds, "what": __doc__.split("\n")[0]}
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    for name in ("sofa_pair", "kitchen_pair"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        kp1 = [(float(p[0]), float(p[1])) for p in g["kp1"]]
        kp2 = [(float(p[0]), float(p[1])) for p in g["kp2"]]
        bf = [tuple(int(x) for x in m) for m in g["bf"]]
        shape = tuple(int(x) for x in g["img_shape"][:2])
        res = {"keypoints": [len(kp1), len(kp2)], "matches": len(bf)}
        for alt_name, alt, n_seeds in (("eigh", solve_eigh, seeds), ("jacobi", solve_jacobi, seeds if only else max(seeds // 5, 1))):
            if only and alt_name != only:
                continue
            diff_final = diff_best = diff_count = diff_raise = ref_raises = 0
            worst = 0

            def attempt(solver, seed):
                # a degenerate sample (randint draws WITH replacement) can give a homography that sends a keypoint to infinity:
                # the reference's int() raises there, and so does the drop-in -- which sample does it depends on the solver
                np.random.seed(seed)
                try:
                    return ransac(kp1, kp2, bf, shape, solver)
                except (OverflowError, ValueError, IndexError) as exc:
                    return type(exc).__name__

            for seed in range(n_seeds):
                r0, r1 = attempt(solve_eig, seed), attempt(alt, seed)
                if isinstance(r0, str) or isinstance(r1, str):
                    ref_raises += isinstance(r0, str)
                    diff_raise += r0 != r1 if (isinstance(r0, str) and isinstance(r1, str)) else 1
                    continue
                (c0, b0, m0), (c1, b1, m1) = r0, r1
                diff_count += c0 != c1
                diff_best += b0 != b1
                same = m0.shape == m1.shape and np.array_equal(m0, m1)
                diff_final += not same
                if not same:
                    worst = max(worst, len(set(map(tuple, m0)) ^ set(map(tuple, m1))))
            res[alt_name] = {"seeds": n_seeds, "reference_solver_raises": int(ref_raises), "raises_differently": int(diff_raise),
                             "final_inlier_list_differs": int(diff_final), "best_hypothesis_differs": int(diff_best),
                             "best_inlier_count_differs": int(diff_count), "largest_symmetric_difference": int(worst)}
        out[name] = res
print(json.dumps(out, indent=1))
