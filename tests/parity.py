"""Comparators implementing the tolerances BASELINE.json's north_star states."""
import numpy as np


def assert_nn_parity(src, dst, idx, dist, idx_ref, dist_ref, what=""):
    """NN indices must be identical, except where the two choices are a distance tie within 1e-6
    relative; distances must agree to 1e-6 relative everywhere (1e-12 expected in practice)."""
    idx = np.asarray(idx).astype(np.int64)
    idx_ref = np.asarray(idx_ref).astype(np.int64)
    assert idx.shape == idx_ref.shape, what
    assert (idx >= 0).all() and (idx < len(dst)).all(), what
    diff = np.nonzero(idx != idx_ref)[0]
    if diff.size:
        d_mine = np.linalg.norm(src[diff] - dst[idx[diff]], axis=1)
        d_ref = np.linalg.norm(src[diff] - dst[idx_ref[diff]], axis=1)
        rel = np.abs(d_mine - d_ref) / np.maximum(d_ref, 1e-300)
        bad = (rel > 1e-6) & (np.abs(d_mine - d_ref) > 1e-12)
        assert not bad.any(), "%s: %d index mismatches that are not ties (worst rel %g)" % (what, bad.sum(), rel.max())
    np.testing.assert_allclose(dist, dist_ref, rtol=1e-6, atol=1e-12, err_msg=what)
    return diff.size


def assert_transform_parity(T, T_ref, extent, what="", rot_tol=1e-4, trans_tol=1e-4):
    """Frobenius norm of the rotation difference <= 1e-4; translation within 1e-4 x scene extent."""
    T = np.asarray(T)
    T_ref = np.asarray(T_ref)
    dr = np.linalg.norm(T[:3, :3] - T_ref[:3, :3])
    dt = np.linalg.norm(T[:3, 3] - T_ref[:3, 3])
    assert dr <= rot_tol, "%s: rotation differs by %g (Frobenius)" % (what, dr)
    assert dt <= trans_tol * extent, "%s: translation differs by %g (extent %g)" % (what, dt, extent)
    assert np.allclose(T[3], [0, 0, 0, 1]), what
    return dr, dt
