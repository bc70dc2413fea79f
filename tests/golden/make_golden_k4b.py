"""Generate tests/golden/k4b.npz: the pin of K4B (3-D hypothesis scoring, BASELINE.json configs[4]).

    python tests/golden/make_golden_k4b.py        (build container only: needs /root/reference)

The reference has no 3-D hypothesis loop, so there is no function to call for the whole thing; what
SURVEY.md section 8(c) defines is a COMPOSITION OF THE REFERENCE'S OWN PRIMITIVES, and that composition is what runs here,
every piece imported unmodified from /root/reference:

    hypothesis   utils/r3d.py:22 rigid_transform_3D on a 3-point sample, or
                 utils/transformation3d.py:163 homo_rigid_transform_3d on a 4-point sample
    residual     || utils/transformation3d.py:183 get_transformed_points(P, T) - Q ||^2   (per point, fp64)
    inlier       residual < 10                                       (the reference's SSD threshold, q9.py:122)

Inputs follow SURVEY.md section 8(d) at 1/10 of the sweep's size: 5 000 points uniform in [-1000, 1000]^3 stored as float32
values (so every implementation sees identical numbers), Q = T* P + N(0, 1) for 60 %, uniform outliers for 40 % (seed 11),
2 000 rigid and 2 000 affine samples from np.random.RandomState(12345).randint.  Stored: P, Q, the samples, the reference's
hypotheses (they also pin the hypothesis kernels, K6) and the inlier counts.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import install_shim, save  # noqa: E402

N_PTS, N_HYP = 5000, 2000


def main():
    install_shim()
    import utils.r3d as r3d
    import utils.transformation3d as t3d

    rng = np.random.default_rng(11)
    P = rng.uniform(-1000, 1000, (N_PTS, 3)).astype(np.float32).astype(np.float64)
    ang = np.deg2rad(2.0)
    Tstar = np.eye(4)
    Tstar[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]
    Tstar[:3, 3] = [15, -8, 5]
    Q = t3d.get_transformed_points(P, Tstar) + rng.normal(size=(N_PTS, 3))
    outlier = rng.random(N_PTS) < 0.4
    Q[outlier] = rng.uniform(-1000, 1000, (int(outlier.sum()), 3))
    rs = np.random.RandomState(12345)
    samples3 = rs.randint(N_PTS, size=(N_HYP, 3))
    samples4 = rs.randint(N_PTS, size=(N_HYP, 4))
    inl = np.nonzero(~outlier)[0]
    samples3[:20] = inl[:60].reshape(20, 3)       # some all-inlier samples: hypotheses near T* with thousands of inliers
    samples4[:20] = inl[60:140].reshape(20, 4)

    def count(T):
        r2 = ((t3d.get_transformed_points(P, T) - Q) ** 2).sum(axis=1)
        return int((r2 < 10).sum())

    T_rigid = np.zeros((N_HYP, 4, 4))
    T_affine = np.zeros((N_HYP, 4, 4))
    with contextlib.redirect_stdout(io.StringIO()):       # rigid_transform_3D prints when it corrects a reflection
        for k in range(N_HYP):
            R, t = r3d.rigid_transform_3D(P[samples3[k]].T, Q[samples3[k]].T)
            T_rigid[k, :3, :3], T_rigid[k, :3, 3], T_rigid[k, 3, 3] = R, t[:, 0], 1.0
            T_affine[k] = t3d.homo_rigid_transform_3d(P[samples4[k]], Q[samples4[k]])
    ok_r = np.isfinite(T_rigid).all(axis=(1, 2))
    ok_a = np.isfinite(T_affine).all(axis=(1, 2))
    counts_rigid = np.array([count(T) if ok else -1 for T, ok in zip(T_rigid, ok_r)])
    counts_affine = np.array([count(T) if ok else -1 for T, ok in zip(T_affine, ok_a)])
    save("k4b", P=P.astype(np.float32), Q=Q, T_star=Tstar, count_star=np.array(count(Tstar)), samples3=samples3.astype(np.int32),
         samples4=samples4.astype(np.int32), T_rigid=T_rigid, T_affine=T_affine, counts_rigid=counts_rigid, counts_affine=counts_affine)
    print("T*: %d inliers; rigid: max %d, %d hypotheses with > 100; affine: max %d, %d with > 100; non-finite %d / %d" % (
        count(Tstar), counts_rigid.max(), (counts_rigid > 100).sum(), counts_affine.max(), (counts_affine > 100).sum(),
        (~ok_r).sum(), (~ok_a).sum()))


if __name__ == "__main__":
    main()
