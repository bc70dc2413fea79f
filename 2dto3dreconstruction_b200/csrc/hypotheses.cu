// K6: minimal-sample hypothesis generation on the device, one hypothesis per thread.
//   r3d_rigid_hypotheses   3 correspondences -> rigid 4x4: utils/r3d.py:22-60 rigid_transform_3D on the sample
//                          (centroids, H = Am Bm^T, R = V U^T with the reflection fix, t = cB - R cA)
//   r3d_affine_hypotheses  4 correspondences -> affine 4x4: utils/transformation3d.py:163-180
//                          homo_rigid_transform_3d on the sample; four points determine the 12 unknowns, so the
//                          least-squares solution is the solution of the 4x4 system [p 1] X = q (three right-hand
//                          sides), solved here by Gaussian elimination with partial pivoting in fp64; p is rounded
//                          through float32 first, as the reference's float32 design matrix does (:146).
// The sample indices come from the host (the reference's sampling uses NumPy's global RNG and stays in Python).
// status: 0 ok; 1 = degenerate sample (repeated / collinear points for the rigid fit, coplanar for the affine fit):
// the reference's LAPACK calls return one of infinitely many solutions there, which is not reproduced.
#include "common.cuh"

namespace r3d {

__global__ void __launch_bounds__(128)
rigid_hyp_kernel(const Point4* __restrict__ P, const Point4* __restrict__ Q, int n_pts, const int32_t* __restrict__ samples,
                 int n_hyp, double* __restrict__ T_out, int32_t* __restrict__ status) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hyp) return;
    double a[3][3], b[3][3];
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        int i = samples[(size_t)h * 3 + k];
        if (i < 0 || i >= n_pts) { bad = true; i = 0; }
        const Point4 p = load_point(P + i), q = load_point(Q + i);
        a[k][0] = p.x; a[k][1] = p.y; a[k][2] = p.z;
        b[k][0] = q.x; b[k][1] = q.y; b[k][2] = q.z;
    }
    double ca[3], cb[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        ca[d] = (a[0][d] + a[1][d] + a[2][d]) / 3.0;
        cb[d] = (b[0][d] + b[1][d] + b[2][d]) / 3.0;
    }
    double H[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 3; k++) s += (a[k][i] - ca[i]) * (b[k][j] - cb[j]);
            H[i * 3 + j] = s;
        }
    double R[9];
    kabsch_rotation(H, R);
    // rank of the sample: the centred source triangle must have two independent edges (and so must the target)
    auto tri_degenerate = [](const double (&p)[3][3]) {
        const double e1[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
        const double e2[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
        const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
        const double n1 = e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2], n2 = e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2];
        return !(cx * cx + cy * cy + cz * cz > 1e-20 * n1 * n2) || !(n1 > 0.0) || !(n2 > 0.0);
    };
    bad = bad || tri_degenerate(a) || tri_degenerate(b);
    double* T = T_out + (size_t)h * 16;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        T[i * 4 + 0] = R[i * 3 + 0];
        T[i * 4 + 1] = R[i * 3 + 1];
        T[i * 4 + 2] = R[i * 3 + 2];
        T[i * 4 + 3] = cb[i] - (R[i * 3 + 0] * ca[0] + R[i * 3 + 1] * ca[1] + R[i * 3 + 2] * ca[2]);
    }
    T[12] = 0.0; T[13] = 0.0; T[14] = 0.0; T[15] = 1.0;
    if (status) status[h] = bad ? 1 : 0;
}

__global__ void __launch_bounds__(128)
affine_hyp_kernel(const Point4* __restrict__ P, const Point4* __restrict__ Q, int n_pts, const int32_t* __restrict__ samples,
                  int n_hyp, double* __restrict__ T_out, int32_t* __restrict__ status) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_hyp) return;
    // augmented system [px py pz 1 | qx qy qz], four rows
    double M[4][7];
    bool bad = false;
    double scale = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int i = samples[(size_t)h * 4 + k];
        if (i < 0 || i >= n_pts) { bad = true; i = 0; }
        const Point4 p = load_point(P + i), q = load_point(Q + i);
        M[k][0] = (double)(float)p.x; M[k][1] = (double)(float)p.y; M[k][2] = (double)(float)p.z; M[k][3] = 1.0;
        M[k][4] = q.x; M[k][5] = q.y; M[k][6] = q.z;
        scale = fmax(scale, fmax(fabs(M[k][0]), fmax(fabs(M[k][1]), fabs(M[k][2]))));
    }
    scale = fmax(scale, 1.0);
    double minpiv = CUDART_INF;
    // Gaussian elimination with partial pivoting; fully unrolled so that M stays in registers
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int piv = c;
        double best = fabs(M[c][c]);
#pragma unroll
        for (int r = c + 1; r < 4; r++) {
            const double v = fabs(M[r][c]);
            if (v > best) { best = v; piv = r; }
        }
#pragma unroll
        for (int r = c + 1; r < 4; r++) {
            if (piv == r) {
#pragma unroll
                for (int j = 0; j < 7; j++) { const double t = M[c][j]; M[c][j] = M[r][j]; M[r][j] = t; }
            }
        }
        const double p = M[c][c];
        minpiv = fmin(minpiv, fabs(p) / (c < 3 ? scale : 1.0));
        const double inv = (p != 0.0) ? 1.0 / p : 0.0;
#pragma unroll
        for (int r = c + 1; r < 4; r++) {
            const double f = M[r][c] * inv;
#pragma unroll
            for (int j = c; j < 7; j++) M[r][j] -= f * M[c][j];
        }
    }
    bad = bad || !(minpiv > 1e-9);
    // back substitution for the three right-hand sides: X[c][d] = coefficient of [px py pz 1][c] in output row d
    double X[4][3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
#pragma unroll
        for (int c = 3; c >= 0; c--) {
            double s = M[c][4 + d];
#pragma unroll
            for (int j = c + 1; j < 4; j++) s -= M[c][j] * X[j][d];
            X[c][d] = (M[c][c] != 0.0) ? s / M[c][c] : 0.0;
        }
    }
    double* T = T_out + (size_t)h * 16;
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
        for (int c = 0; c < 4; c++) T[d * 4 + c] = X[c][d];
    T[12] = 0.0; T[13] = 0.0; T[14] = 0.0; T[15] = 1.0;
    if (status) status[h] = bad ? 1 : 0;
}

}  // namespace r3d

using namespace r3d;

extern "C" int r3d_rigid_hypotheses(const double* P, const double* Q, int n_pts, const int32_t* samples, int n_hyp,
                                    double* T_out, int32_t* status_out, void* stream) {
    if (!P || !Q || !samples || !T_out || n_pts <= 0 || n_hyp <= 0) return R3D_ERR_INVALID;
    if ((((uintptr_t)P) & 31) || (((uintptr_t)Q) & 31)) return R3D_ERR_INVALID;      // xyzw records are read with 256-bit loads
    R3D_LAUNCH(rigid_hyp_kernel, (n_hyp + 127) / 128, 128, 0, (cudaStream_t)stream, (const Point4*)P, (const Point4*)Q, n_pts,
               samples, n_hyp, T_out, status_out);
    return R3D_OK;
}

extern "C" int r3d_affine_hypotheses(const double* P, const double* Q, int n_pts, const int32_t* samples, int n_hyp,
                                     double* T_out, int32_t* status_out, void* stream) {
    if (!P || !Q || !samples || !T_out || n_pts <= 0 || n_hyp <= 0) return R3D_ERR_INVALID;
    if ((((uintptr_t)P) & 31) || (((uintptr_t)Q) & 31)) return R3D_ERR_INVALID;      // xyzw records are read with 256-bit loads
    R3D_LAUNCH(affine_hyp_kernel, (n_hyp + 127) / 128, 128, 0, (cudaStream_t)stream, (const Point4*)P, (const Point4*)Q, n_pts,
               samples, n_hyp, T_out, status_out);
    return R3D_OK;
}
