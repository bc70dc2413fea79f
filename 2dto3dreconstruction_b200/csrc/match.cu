// K5: brute-force descriptor matching with cross-check -- the step right before
// RANSAC.  Replaces cv2.BFMatcher(normType=cv2.NORM_L2, crossCheck=True).match(des1, des2) as the reference
// calls it (utils/transformation3d.py:31-32): query i and train j are matched iff j is the nearest train
// descriptor of i AND i is the nearest query descriptor of j, first minimum (lowest index) on ties in both
// directions; the distance is sqrt(sum (a-b)^2) in float32.
//
// SIFT descriptors are float32 holding integers 0..255, so every squared distance is an integer below 2^24 and
// exact in float32 whatever the summation order: indices AND distances are bit-identical to OpenCV's.  For
// general float descriptors the indices agree wherever the minima are not within rounding of each other.
//
// r3d_match_descriptors_ratio adds the second-best ratio test of skimage.feature.match_descriptors(des1, des2,
// max_ratio=0.8, cross_check=True) as the reference calls it at utils/homography_utils/q8.py:90 (scikit-image 0.17.2,
// requirements.txt:94, not vendored): after the mutual check a query is kept iff best / second_best < max_ratio,
// where second_best is the smallest distance from the query to any OTHER train descriptor (0 replaced by DBL_EPSILON),
// the distances being float64 square roots -- of exact integers for SIFT, so the test is bit-identical too.
//
// Sizes are small (hundreds to a few thousand descriptors of 128 floats, < 1 GFLOP): one launch of a
// shared-memory tiled argmin kernel (both directions, blockIdx.y) and one cross-check pass.  No tensor
// cores: the whole problem is tens of microseconds of CUDA-core work and has to be exact.
#include "common.cuh"

namespace r3d {

constexpr int MT_THREADS = 128;
constexpr int MT_XR = 8;                                  // X rows per thread (their running minima live in registers)
constexpr int MT_X_TILE = (MT_THREADS / 32) * MT_XR;      // 32 X rows per CTA
constexpr int MT_Y_TILE = 32;                             // Y rows per step: one per lane
constexpr int MT_C_TILE = 32;                             // descriptor columns staged per step

// For every row of X: argmin over the rows of Y of the squared L2 distance (first minimum on ties), and the second
// smallest distance.  blockIdx.y selects the direction: 0: X = A, Y = B; 1: X = B, Y = A.
// Round 2 layout: a LANE owns one Y row of the current 32-row tile and a WARP owns eight X rows, so a thread accumulates
// eight full distances in registers -- per staged column one conflict-free LDS of its Y value, two broadcast LDS.128 of
// the eight X values, eight subtract + FMA pairs -- and the only cross-lane step is one lexicographic reduction per X row
// at the very end (round 1 reduced every one of the nx x ny distances over the 32 lanes: 0.91 ms for 1100 x 1200).
__global__ void __launch_bounds__(MT_THREADS)
bf_argmin_kernel(const float* __restrict__ A, int na, const float* __restrict__ B, int nb, int dim, int dim_pad,
                 int32_t* __restrict__ idx_ab, float* __restrict__ d2_ab, int32_t* __restrict__ idx_ba,
                 float* __restrict__ d2_ba, float* __restrict__ second_ab) {
    extern __shared__ __align__(16) float sm[];
    const bool rev = blockIdx.y == 1;
    const float* X = rev ? B : A;
    const float* Y = rev ? A : B;
    const int nx = rev ? nb : na, ny = rev ? na : nb;
    int32_t* idx_out = rev ? idx_ba : idx_ab;
    float* d2_out = rev ? d2_ba : d2_ab;
    const int x0 = blockIdx.x * MT_X_TILE;
    if (x0 >= nx) return;
    float* sX = sm;                                       // [dim_pad][MT_X_TILE]  column-major: the 8 rows of a warp are contiguous
    float* sY = sm + (size_t)dim_pad * MT_X_TILE;         // [MT_C_TILE][MT_Y_TILE + 1]  one column tile of the Y rows, transposed
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = threadIdx.x; e < MT_X_TILE * dim_pad; e += MT_THREADS) {
        const int r = e / dim_pad, c = e - r * dim_pad;      // (coalesced global reads along the descriptor)
        sX[(size_t)c * MT_X_TILE + r] = (x0 + r < nx && c < dim) ? X[(size_t)(x0 + r) * dim + c] : 0.0f;
    }
    const float inf = __int_as_float(0x7f800000);
    float best[MT_XR], second[MT_XR];      // per lane: over the Y rows this lane has owned so far
    int bidx[MT_XR];
#pragma unroll
    for (int k = 0; k < MT_XR; k++) { best[k] = second[k] = inf; bidx[k] = 0x7fffffff; }
    for (int y0 = 0; y0 < ny; y0 += MT_Y_TILE) {
        float acc[MT_XR];
#pragma unroll
        for (int k = 0; k < MT_XR; k++) acc[k] = 0.0f;
        for (int c0 = 0; c0 < dim_pad; c0 += MT_C_TILE) {
            __syncthreads();      // (also orders the sX fill before its first use)
            for (int e = threadIdx.x; e < MT_Y_TILE * MT_C_TILE; e += MT_THREADS) {
                const int r = e / MT_C_TILE, c = e - r * MT_C_TILE;
                sY[c * (MT_Y_TILE + 1) + r] = (y0 + r < ny && c0 + c < dim) ? Y[(size_t)(y0 + r) * dim + c0 + c] : 0.0f;
            }
            __syncthreads();
#pragma unroll 8
            for (int c = 0; c < MT_C_TILE; c++) {
                const float yv = sY[c * (MT_Y_TILE + 1) + lane];
                const float4 xa = *reinterpret_cast<const float4*>(sX + (size_t)(c0 + c) * MT_X_TILE + warp * MT_XR);
                const float4 xb = *reinterpret_cast<const float4*>(sX + (size_t)(c0 + c) * MT_X_TILE + warp * MT_XR + 4);
                float d;
                d = xa.x - yv; acc[0] = fmaf(d, d, acc[0]);
                d = xa.y - yv; acc[1] = fmaf(d, d, acc[1]);
                d = xa.z - yv; acc[2] = fmaf(d, d, acc[2]);
                d = xa.w - yv; acc[3] = fmaf(d, d, acc[3]);
                d = xb.x - yv; acc[4] = fmaf(d, d, acc[4]);
                d = xb.y - yv; acc[5] = fmaf(d, d, acc[5]);
                d = xb.z - yv; acc[6] = fmaf(d, d, acc[6]);
                d = xb.w - yv; acc[7] = fmaf(d, d, acc[7]);
            }
        }
        const int y = y0 + lane;
        if (y < ny) {
#pragma unroll
            for (int k = 0; k < MT_XR; k++) {
                // the lane's Y rows come in ascending order: strict < keeps the first minimum
                if (acc[k] < best[k]) { second[k] = best[k]; best[k] = acc[k]; bidx[k] = y; }
                else if (acc[k] < second[k]) second[k] = acc[k];
            }
        }
    }
    // one lexicographic (distance, row) reduction per X row; the second smallest distance over the union of two sets is
    // min(max(best_a, best_b), second_a, second_b)
#pragma unroll
    for (int k = 0; k < MT_XR; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best[k], o), os = __shfl_xor_sync(0xffffffffu, second[k], o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx[k], o);
            second[k] = fminf(fmaxf(best[k], ob), fminf(second[k], os));
            if (ob < best[k] || (ob == best[k] && oi < bidx[k])) { best[k] = ob; bidx[k] = oi; }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < MT_XR; k++) {
            const int x = x0 + warp * MT_XR + k;
            if (x < nx) {
                idx_out[x] = bidx[k] == 0x7fffffff ? -1 : bidx[k];
                d2_out[x] = best[k];
                if (!rev && second_ab) second_ab[x] = second[k];
            }
        }
    }
}

__global__ void bf_cross_kernel(const int32_t* __restrict__ idx_ab, const float* __restrict__ d2_ab,
                                const int32_t* __restrict__ idx_ba, const float* __restrict__ second_ab, double max_ratio,
                                int na, int cross_check, int32_t* __restrict__ train_out, float* __restrict__ dist_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= na) return;
    const int j = idx_ab[i];
    bool keep = j >= 0 && (!cross_check || idx_ba[j] == i);
    if (keep && second_ab) {
        // skimage: ratio = best / second_best in float64, second_best == 0 -> eps; with a single train descriptor
        // there is no other candidate and the "second best" read back is the inf written over the best: ratio 0
        double sb = sqrt((double)second_ab[i]);
        if (sb == 0.0) sb = 2.220446049250313e-16;
        keep = sqrt((double)d2_ab[i]) / sb < max_ratio;
    }
    train_out[i] = keep ? j : -1;
    dist_out[i] = keep ? sqrtf(d2_ab[i]) : __int_as_float(0x7f800000);
}

struct MatchWs {
    int32_t* idx_ab;
    float* d2_ab;
    int32_t* idx_ba;
    float* d2_ba;
    float* second_ab;
};

static void match_layout(MatchWs& w, Arena& a, int n1, int n2) {
    w.idx_ab = a.take<int32_t>((size_t)n1);
    w.d2_ab = a.take<float>((size_t)n1);
    w.idx_ba = a.take<int32_t>((size_t)n2);
    w.d2_ba = a.take<float>((size_t)n2);
    w.second_ab = a.take<float>((size_t)n1);
}

}  // namespace r3d

using namespace r3d;

extern "C" size_t r3d_match_workspace_bytes(int n1, int n2) {
    if (n1 <= 0 || n2 <= 0) return 0;
    MatchWs w;
    Arena a(nullptr, 0);
    match_layout(w, a, n1, n2);
    return a.used + 512;
}

static int match_run(const float* des1, int n1, const float* des2, int n2, int dim, int cross_check, bool ratio_test,
                     double max_ratio, int32_t* train_idx_out, float* dist_out, void* workspace, size_t workspace_bytes,
                     void* stream) {
    if (!des1 || !des2 || !train_idx_out || !dist_out || n1 <= 0 || n2 <= 0 || dim <= 0) return R3D_ERR_INVALID;
    if (ratio_test && !(max_ratio == max_ratio)) return R3D_ERR_INVALID;
    const int dim_pad = (dim + 31) / 32 * 32;
    const size_t smem = ((size_t)MT_X_TILE * dim_pad + (size_t)MT_C_TILE * (MT_Y_TILE + 1)) * sizeof(float);
    if (smem > 200 * 1024) return R3D_ERR_INVALID;      // descriptors longer than ~1000 floats are not supported
    cudaStream_t st = (cudaStream_t)stream;
    MatchWs w;
    Arena a(workspace, workspace_bytes);
    match_layout(w, a, n1, n2);
    if (!a.ok) return R3D_ERR_WORKSPACE;
    if (smem > 48 * 1024) R3D_CUDA(cudaFuncSetAttribute(bf_argmin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nmax = n1 > n2 ? n1 : n2;
    float* second = ratio_test ? w.second_ab : nullptr;
    R3D_LAUNCH(bf_argmin_kernel, dim3((nmax + MT_X_TILE - 1) / MT_X_TILE, 2), MT_THREADS, smem, st, des1, n1, des2, n2, dim,
               dim_pad, w.idx_ab, w.d2_ab, w.idx_ba, w.d2_ba, second);
    R3D_LAUNCH(bf_cross_kernel, (n1 + 255) / 256, 256, 0, st, w.idx_ab, w.d2_ab, w.idx_ba, (const float*)second, max_ratio, n1,
               cross_check, train_idx_out, dist_out);
    return R3D_OK;
}

extern "C" int r3d_match_descriptors(const float* des1, int n1, const float* des2, int n2, int dim, int cross_check,
                                     int32_t* train_idx_out, float* dist_out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    return match_run(des1, n1, des2, n2, dim, cross_check, false, 1.0, train_idx_out, dist_out, workspace, workspace_bytes, stream);
}

extern "C" int r3d_match_descriptors_ratio(const float* des1, int n1, const float* des2, int n2, int dim, int cross_check,
                                           double max_ratio, int32_t* train_idx_out, float* dist_out, void* workspace,
                                           size_t workspace_bytes, void* stream) {
    return match_run(des1, n1, des2, n2, dim, cross_check, max_ratio < 1.0, max_ratio, train_idx_out, dist_out, workspace,
                     workspace_bytes, stream);
}
