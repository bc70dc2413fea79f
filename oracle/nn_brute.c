/* Brute-force exact nearest neighbour in float64 -- TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Restates what utils/icp.py:58-71 nearest_neighbor computes (the Euclidean 1-NN that
 * scikit-learn's KD-tree returns) in its defining form: argmin_j |s_i - d_j|^2, evaluated as
 * (a-b)^2 sums in double precision, lowest j on exact ties.  Used by tests at sizes where the
 * NumPy version in ref_port.py would need too much memory (callers split the queries over threads).  Build: `make -C oracle`.
 */
#include <stddef.h>
#include <stdint.h>

void nn_brute(const double* src, int64_t n, const double* dst, int64_t m, int64_t* idx, double* d2out) {
    for (int64_t i = 0; i < n; i++) {
        const double qx = src[3 * i], qy = src[3 * i + 1], qz = src[3 * i + 2];
        double best = 1.0 / 0.0;
        int64_t bj = -1;
        for (int64_t j = 0; j < m; j++) {
            const double dx = qx - dst[3 * j], dy = qy - dst[3 * j + 1], dz = qz - dst[3 * j + 2];
            const double d = dx * dx + dy * dy + dz * dz;
            if (d < best) { best = d; bj = j; }
        }
        idx[i] = bj;
        d2out[i] = best;
    }
}
