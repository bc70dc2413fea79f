// K3 / K3c: closed-form fits on given correspondences, and applying a 4x4 to a cloud.
//   r3d_kabsch            utils/icp.py:22-55 best_fit_transform, utils/r3d.py:22-60 rigid_transform_3D
//   r3d_affine_fit        utils/transformation3d.py:133-180 homo_rigid_transform_3d
//   r3d_transform_points  utils/transformation3d.py:183-187 get_transformed_points
// All arithmetic in fp64.  Reductions are two-level (per-CTA partials in a fixed order, then one
// CTA), so results do not depend on scheduling.
#include "common.cuh"

namespace r3d {

constexpr int FIT_THREADS = 256;
constexpr int FIT_MAX_BLOCKS = 592;   // 4 CTAs per SM

// Block-wide sum of NV doubles per thread -> result valid in thread 0's smem row `out[NV]`.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* NV * 8 */, double* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        double s = warp_sum(v[k]);
        if (lane == 0) smem[k * 8 + warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < FIT_THREADS / 32; w++) s += smem[threadIdx.x * 8 + w];
        out[threadIdx.x] = s;
    }
}

// ---- Kabsch on given pairs --------------------------------------------------------------
__global__ void __launch_bounds__(FIT_THREADS)
kabsch_partial(const double* __restrict__ a, int64_t aps, int64_t ads, const double* __restrict__ b, int64_t bps,
               int64_t bds, int64_t n, double* __restrict__ partial) {
    // shift by the first pair: keeps the raw second moments small relative to the centred ones
    const double sa0 = a[0], sa1 = a[ads], sa2 = a[2 * ads];
    const double sb0 = b[0], sb1 = b[bds], sb2 = b[2 * bds];
    double m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m[k] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * FIT_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * FIT_THREADS) {
        const double* pa = a + i * aps;
        const double* pb = b + i * bps;
        double ax = pa[0] - sa0, ay = pa[ads] - sa1, az = pa[2 * ads] - sa2;
        double bx = pb[0] - sb0, by = pb[bds] - sb1, bz = pb[2 * bds] - sb2;
        m[0] += ax; m[1] += ay; m[2] += az;
        m[3] += bx; m[4] += by; m[5] += bz;
        m[6] += ax * bx; m[7] += ax * by; m[8] += ax * bz;
        m[9] += ay * bx; m[10] += ay * by; m[11] += ay * bz;
        m[12] += az * bx; m[13] += az * by; m[14] += az * bz;
    }
    __shared__ double smem[16 * 8];
    block_sum<16>(m, smem, partial + (size_t)blockIdx.x * 16);
}

__global__ void __launch_bounds__(32)
kabsch_final(const double* __restrict__ partial, int nblocks, const double* __restrict__ a, int64_t ads,
             const double* __restrict__ b, int64_t bds, int64_t n, double* __restrict__ T_out) {
    const int lane = threadIdx.x;
    double v = 0.0;   // lanes 0..15 each own one moment; fixed summation order
    if (lane < 16)
        for (int k = 0; k < nblocks; k++) v += partial[(size_t)k * 16 + lane];
    double m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m[k] = __shfl_sync(0xffffffffu, v, k);
    if (lane == 0) {
        // separate shifts for a and b: fold them into centroids by hand
        const double sa[3] = {a[0], a[ads], a[2 * ads]};
        const double sb[3] = {b[0], b[bds], b[2 * bds]};
        const double dn = (double)n;
        double ca[3] = {m[0] / dn, m[1] / dn, m[2] / dn};
        double cb[3] = {m[3] / dn, m[4] / dn, m[5] / dn};
        double H[9], R[9];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) H[i * 3 + j] = m[6 + i * 3 + j] - dn * ca[i] * cb[j];
        kabsch_rotation(H, R);
        for (int i = 0; i < 3; i++) { ca[i] += sa[i]; cb[i] += sb[i]; }
        for (int i = 0; i < 3; i++) {
            T_out[i * 4 + 0] = R[i * 3 + 0];
            T_out[i * 4 + 1] = R[i * 3 + 1];
            T_out[i * 4 + 2] = R[i * 3 + 2];
            T_out[i * 4 + 3] = cb[i] - (R[i * 3 + 0] * ca[0] + R[i * 3 + 1] * ca[1] + R[i * 3 + 2] * ca[2]);
        }
        T_out[12] = 0; T_out[13] = 0; T_out[14] = 0; T_out[15] = 1;
    }
}

// ---- 12-dof affine least squares -----------------------------------------------------------
// The reference's (3K,12) design matrix is block diagonal in three copies of [p 1]
// (utils/transformation3d.py:133-147), so the normal equations are one shared 4x4 Gram matrix
// and a 4x3 right-hand side.  With p centred the Gram matrix is diag(C, K): M = (sum q pc^T) C^-1,
// t = mean(q) - M mean(p).  Moments: sum p (3), sum q (3), sum pc pc^T (6), sum q pc^T (9) with
// pc, q shifted by the first point (two-pass centring is folded in algebraically afterwards).
__global__ void __launch_bounds__(FIT_THREADS)
affine_partial(const double* __restrict__ p, const double* __restrict__ q, int64_t n, double* __restrict__ partial) {
    const double sp0 = (double)(float)p[0], sp1 = (double)(float)p[1], sp2 = (double)(float)p[2];
    const double sq0 = q[0], sq1 = q[1], sq2 = q[2];
    double m[21];
#pragma unroll
    for (int k = 0; k < 21; k++) m[k] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * FIT_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * FIT_THREADS) {
        // the design matrix is float32 in the reference: p goes through a float rounding
        double px = (double)(float)p[i * 3 + 0] - sp0, py = (double)(float)p[i * 3 + 1] - sp1,
               pz = (double)(float)p[i * 3 + 2] - sp2;
        double qx = q[i * 3 + 0] - sq0, qy = q[i * 3 + 1] - sq1, qz = q[i * 3 + 2] - sq2;
        m[0] += px; m[1] += py; m[2] += pz;
        m[3] += qx; m[4] += qy; m[5] += qz;
        m[6] += px * px; m[7] += px * py; m[8] += px * pz; m[9] += py * py; m[10] += py * pz; m[11] += pz * pz;
        m[12] += qx * px; m[13] += qx * py; m[14] += qx * pz;
        m[15] += qy * px; m[16] += qy * py; m[17] += qy * pz;
        m[18] += qz * px; m[19] += qz * py; m[20] += qz * pz;
    }
    __shared__ double smem[21 * 8];
    block_sum<21>(m, smem, partial + (size_t)blockIdx.x * 21);
}

__global__ void __launch_bounds__(32)
affine_final(const double* __restrict__ partial, int nblocks, const double* __restrict__ p, const double* __restrict__ q,
             int64_t n, double* __restrict__ T_out, int32_t* __restrict__ status_out) {
    const int lane = threadIdx.x;
    double v = 0.0;
    if (lane < 21)
        for (int k = 0; k < nblocks; k++) v += partial[(size_t)k * 21 + lane];
    double m[21];
#pragma unroll
    for (int k = 0; k < 21; k++) m[k] = __shfl_sync(0xffffffffu, v, k);
    if (lane != 0) return;
    const double dn = (double)n;
    const double sp[3] = {(double)(float)p[0], (double)(float)p[1], (double)(float)p[2]};
    const double sq[3] = {q[0], q[1], q[2]};
    double mp[3] = {m[0] / dn, m[1] / dn, m[2] / dn};   // centroids in shifted coordinates
    double mq[3] = {m[3] / dn, m[4] / dn, m[5] / dn};
    // centred second moments
    double C[3][3], B[3][3];
    const double raw[3][3] = {{m[6], m[7], m[8]}, {m[7], m[9], m[10]}, {m[8], m[10], m[11]}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            C[i][j] = raw[i][j] - dn * mp[i] * mp[j];
            B[i][j] = m[12 + i * 3 + j] - dn * mq[i] * mp[j];   // sum (q - mq)_i (p - mp)_j
        }
    // inverse of the SPD 3x3 C via the adjugate (fp64; C is well scaled after centring)
    double c00 = C[1][1] * C[2][2] - C[1][2] * C[2][1];
    double c01 = C[0][2] * C[2][1] - C[0][1] * C[2][2];
    double c02 = C[0][1] * C[1][2] - C[0][2] * C[1][1];
    double det = C[0][0] * c00 + C[1][0] * c01 + C[2][0] * c02;
    double scale = fabs(C[0][0]) * fabs(C[1][1]) * fabs(C[2][2]);
    int bad = !(fabs(det) > 1e-13 * scale) || !isfinite(det);
    double inv[3][3];
    inv[0][0] = c00 / det; inv[0][1] = c01 / det; inv[0][2] = c02 / det;
    inv[1][0] = inv[0][1];
    inv[1][1] = (C[0][0] * C[2][2] - C[0][2] * C[2][0]) / det;
    inv[1][2] = (C[0][2] * C[1][0] - C[0][0] * C[1][2]) / det;
    inv[2][0] = inv[0][2];
    inv[2][1] = inv[1][2];
    inv[2][2] = (C[0][0] * C[1][1] - C[0][1] * C[1][0]) / det;
    double M[3][3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M[i][j] = B[i][0] * inv[0][j] + B[i][1] * inv[1][j] + B[i][2] * inv[2][j];
    for (int i = 0; i < 3; i++) {
        double cp0 = mp[0] + sp[0], cp1 = mp[1] + sp[1], cp2 = mp[2] + sp[2];
        T_out[i * 4 + 0] = M[i][0];
        T_out[i * 4 + 1] = M[i][1];
        T_out[i * 4 + 2] = M[i][2];
        T_out[i * 4 + 3] = (mq[i] + sq[i]) - (M[i][0] * cp0 + M[i][1] * cp1 + M[i][2] * cp2);
    }
    T_out[12] = 0; T_out[13] = 0; T_out[14] = 0; T_out[15] = 1;
    if (status_out) *status_out = bad;
}

__global__ void __launch_bounds__(256)
transform_points_kernel(const double* in, int64_t n, const double* __restrict__ T, double* out) {
    __shared__ double t[12];
    if (threadIdx.x < 12) t[threadIdx.x] = T[threadIdx.x];
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double x = in[i * 3], y = in[i * 3 + 1], z = in[i * 3 + 2];
        // same accumulation order as a row of [P 1] . h^T
        out[i * 3 + 0] = ((x * t[0] + y * t[1]) + z * t[2]) + t[3];
        out[i * 3 + 1] = ((x * t[4] + y * t[5]) + z * t[6]) + t[7];
        out[i * 3 + 2] = ((x * t[8] + y * t[9]) + z * t[10]) + t[11];
    }
}

__global__ void __launch_bounds__(256)
pack_xyzw_kernel(const double* __restrict__ xyz, int64_t n, Point4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        store_point(out + i, xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2], (int32_t)i);
}

static int fit_blocks(int64_t n) {
    int64_t b = (n + FIT_THREADS - 1) / FIT_THREADS;
    if (b < 1) b = 1;
    if (b > FIT_MAX_BLOCKS) b = FIT_MAX_BLOCKS;
    return (int)b;
}

}  // namespace r3d

extern "C" size_t r3d_kabsch_workspace_bytes(int64_t n) {
    return r3d::align_up((size_t)r3d::fit_blocks(n) * 16 * sizeof(double), 256) + 256;
}

extern "C" int r3d_kabsch(const double* a, int64_t aps, int64_t ads, const double* b, int64_t bps, int64_t bds,
                          int64_t n, double* T_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!a || !b || !T_out || n <= 0) return R3D_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    r3d::Arena ar(workspace, workspace_bytes);
    const int nb = r3d::fit_blocks(n);
    double* partial = ar.take<double>((size_t)nb * 16);
    if (!ar.ok) return R3D_ERR_WORKSPACE;
    R3D_LAUNCH(r3d::kabsch_partial, nb, r3d::FIT_THREADS, 0, st, a, aps, ads, b, bps, bds, n, partial);
    R3D_LAUNCH(r3d::kabsch_final, 1, 32, 0, st, partial, nb, a, ads, b, bds, n, T_out);
    return R3D_OK;
}

extern "C" size_t r3d_affine_fit_workspace_bytes(int64_t n) {
    return r3d::align_up((size_t)r3d::fit_blocks(n) * 21 * sizeof(double), 256) + 256;
}

extern "C" int r3d_affine_fit(const double* p, const double* q, int64_t n, double* T_out, int32_t* status_out,
                              void* workspace, size_t workspace_bytes, void* stream) {
    if (!p || !q || !T_out || n <= 0) return R3D_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    r3d::Arena ar(workspace, workspace_bytes);
    const int nb = r3d::fit_blocks(n);
    double* partial = ar.take<double>((size_t)nb * 21);
    if (!ar.ok) return R3D_ERR_WORKSPACE;
    R3D_LAUNCH(r3d::affine_partial, nb, r3d::FIT_THREADS, 0, st, p, q, n, partial);
    R3D_LAUNCH(r3d::affine_final, 1, 32, 0, st, partial, nb, p, q, n, T_out, status_out);
    return R3D_OK;
}

extern "C" int r3d_transform_points(const double* points, int64_t n, const double* T, double* out, void* stream) {
    if (!points || !T || !out || n < 0) return R3D_ERR_INVALID;
    if (n == 0) return R3D_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t nb = (n + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    R3D_LAUNCH(r3d::transform_points_kernel, (int)nb, 256, 0, st, points, n, T, out);
    return R3D_OK;
}

extern "C" int r3d_xyz_to_xyzw(const double* xyz, int64_t n, double* xyzw, void* stream) {
    if (!xyz || !xyzw || n < 0 || n > 0x7fffffffll) return R3D_ERR_INVALID;
    if (((uintptr_t)xyzw) & 31) return R3D_ERR_INVALID;
    if (n == 0) return R3D_OK;
    int64_t nb = (n + 255) / 256;
    if (nb > 148 * 8) nb = 148 * 8;
    R3D_LAUNCH(r3d::pack_xyzw_kernel, (int)nb, 256, 0, (cudaStream_t)stream, xyz, n, (r3d::Point4*)xyzw);
    return R3D_OK;
}
