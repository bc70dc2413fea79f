// K4: RANSAC hypothesis scoring.
//  K4A r3d_ransac2d_score: the body of q9.ransac_loop's hypothesis loop
//      (utils/homography_utils/q9.py:169-191) for every hypothesis at once:
//      q9.py:52-69 get_transformed_points, q9.py:129-144 make_binned_array (as a CSR cell table
//      with the reference's [int(r)-1][int(c)-1] negative-index wrap) and q9.py:92-127 get_ssd_v2.
//      The reference's sequential "first keypoint to claim j wins" rule is evaluated in parallel:
//      every keypoint i finds its windowed first-minimum j*(i); a claim succeeds iff i is the
//      smallest claimant of j*, which is exactly what the sequential loop accepts.
//  K4B r3d_ransac3d_score: inlier counts of 4x4 rigid/affine hypotheses over 3-D correspondences.
// All scoring arithmetic is fp64 so that threshold and truncation decisions match NumPy.
#include "common.cuh"
#include <atomic>

namespace r3d {

constexpr int RS2_THREADS = 256;

struct Ransac2dWs {
    int32_t* cell_of;      // [n2] cell id of every image-2 keypoint (or -1 when out of range)
    int32_t* cell_start;   // [h*w + 1]
    int32_t* items;        // [n2]
    int32_t* first_i;      // [n_hyp][n2]
    int32_t* jstar;        // [n_hyp][n1]
    double* ssd;           // [n_hyp][n1]
    int32_t* err;          // [1] build error flag
};

static void ransac2d_layout(Ransac2dWs& w, Arena& a, int n_hyp, int n1, int n2, int h, int wd) {
    w.cell_of = a.take<int32_t>((size_t)n2);
    w.cell_start = a.take<int32_t>((size_t)h * wd + 1);
    w.items = a.take<int32_t>((size_t)n2);
    w.first_i = a.take<int32_t>((size_t)n_hyp * n2);
    w.jstar = a.take<int32_t>((size_t)n_hyp * n1);
    w.ssd = a.take<double>((size_t)n_hyp * n1);
    w.err = a.take<int32_t>(1);
}

// Python int(float): truncation toward zero.  Clamped so the cast is defined; the clamp range is far
// outside any image, so window arithmetic on the clamped value gives the same (empty) ranges.
__device__ __forceinline__ int py_int(double x) {
    x = fmin(fmax(x, -1.0e9), 1.0e9);
    return (int)x;   // C cast truncates toward zero
}

__global__ void r2_cells_kernel(const double* __restrict__ kp2, int n2, int h, int w, Ransac2dWs ws) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) *ws.err = 0;
    if (j >= n2) return;
    const double pr = kp2[2 * j], pc = kp2[2 * j + 1];
    int cell = -1;
    if (isfinite(pr) && isfinite(pc)) {
        const int r = py_int(pr) - 1, c = py_int(pc) - 1;
        // Python list indexing: -len <= index < len, negative wraps (q9.py:142)
        if (r >= -h && r < h && c >= -w && c < w) cell = ((r + h) % h) * w + ((c + w) % w);
    }
    ws.cell_of[j] = cell;
}

// One CTA: histogram of cells -> exclusive scan -> stable fill (ascending j inside a cell).
__global__ void __launch_bounds__(1024)
r2_csr_kernel(int n2, int n_cells, Ransac2dWs ws) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = threadIdx.x; c <= n_cells; c += 1024) ws.cell_start[c] = 0;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < n2; j += 1024) {
        const int c = ws.cell_of[j];
        if (c < 0) atomicExch(ws.err, 1); else atomicAdd(&ws.cell_start[c], 1);
    }
    __syncthreads();
    for (int start = 0; start <= n_cells; start += 1024) {
        const int i = start + threadIdx.x;
        const int v = (i <= n_cells) ? ws.cell_start[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int t = warp_tot[lane], ts = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, ts, o);
                if (lane >= o) ts += y;
            }
            warp_tot[lane] = ts - t;
        }
        __syncthreads();
        const int excl = carry + warp_tot[warp] + x - v;
        if (i <= n_cells) ws.cell_start[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    // rank of j inside its cell = number of earlier keypoints in the same cell (n2 is ~1e3)
    for (int j = threadIdx.x; j < n2; j += 1024) {
        const int c = ws.cell_of[j];
        if (c < 0) continue;
        int rank = 0;
        for (int k = 0; k < j; k++) rank += (ws.cell_of[k] == c);
        ws.items[ws.cell_start[c] + rank] = j;
    }
}

__global__ void r2_fill_kernel(int32_t* p, size_t n, int32_t v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// numpy complex division (Smith's algorithm, as in numpy's loops for complex128), real part only
__device__ __forceinline__ double cdiv_real(double ar, double ai, double br, double bi) {
    if (fabs(br) >= fabs(bi)) {
        if (br == 0.0 && bi == 0.0) return ar / fabs(br);
        const double rat = bi / br, scl = 1.0 / __dadd_rn(br, __dmul_rn(bi, rat));
        return __dmul_rn(__dadd_rn(ar, __dmul_rn(ai, rat)), scl);
    }
    const double rat = br / bi, scl = 1.0 / __dadd_rn(bi, __dmul_rn(br, rat));
    return __dmul_rn(__dadd_rn(__dmul_rn(ar, rat), ai), scl);
}

// (row, col) -> H * (col, row, 1)^T, divide by w, back to (row, col)   (q9.py:62-67).  Hs holds the
// 9 real parts followed by the 9 imaginary parts.  Products and sums are rounded one by one.
__device__ __forceinline__ void homography_point(const double* Hs, bool cplx, double row, double col, double& tr, double& tc) {
    const double x = col, y = row;
    const double px = __dadd_rn(__dadd_rn(__dmul_rn(Hs[0], x), __dmul_rn(Hs[1], y)), Hs[2]);
    const double py = __dadd_rn(__dadd_rn(__dmul_rn(Hs[3], x), __dmul_rn(Hs[4], y)), Hs[5]);
    const double pw = __dadd_rn(__dadd_rn(__dmul_rn(Hs[6], x), __dmul_rn(Hs[7], y)), Hs[8]);
    if (!cplx) {
        tc = px / pw;
        tr = py / pw;
    } else {
        const double ix = __dadd_rn(__dadd_rn(__dmul_rn(Hs[9], x), __dmul_rn(Hs[10], y)), Hs[11]);
        const double iy = __dadd_rn(__dadd_rn(__dmul_rn(Hs[12], x), __dmul_rn(Hs[13], y)), Hs[14]);
        const double iw = __dadd_rn(__dadd_rn(__dmul_rn(Hs[15], x), __dmul_rn(Hs[16], y)), Hs[17]);
        tc = cdiv_real(px, ix, pw, iw);
        tr = cdiv_real(py, iy, pw, iw);
    }
}

__global__ void __launch_bounds__(256)
r2_points_kernel(const double* __restrict__ Hre, const double* __restrict__ Him, const double* __restrict__ kp, int n,
                 double* __restrict__ out) {
    __shared__ double Hs[18];
    if (threadIdx.x < 9) {
        Hs[threadIdx.x] = Hre[threadIdx.x];
        Hs[9 + threadIdx.x] = Him ? Him[threadIdx.x] : 0.0;
    }
    __syncthreads();
    bool cplx = false;
#pragma unroll
    for (int k = 0; k < 9; k++) cplx |= (Hs[9 + k] != 0.0);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double tr, tc;
    homography_point(Hs, cplx, kp[2 * i], kp[2 * i + 1], tr, tc);
    out[2 * i] = tr;
    out[2 * i + 1] = tc;
}

__global__ void __launch_bounds__(RS2_THREADS)
r2_score_kernel(const double* __restrict__ Hre, const double* __restrict__ Him, const double* __restrict__ kp1, int n1,
                const double* __restrict__ kp2, int n2, int h, int w, Ransac2dWs ws, int32_t* __restrict__ counts_out,
                int32_t* __restrict__ match_j_out, double* __restrict__ match_ssd_out, int32_t* __restrict__ flags_out) {
    const int hyp = blockIdx.x;
    __shared__ double Hs[18];
    __shared__ int s_cnt[RS2_THREADS / 32];
    __shared__ int s_flags;
    if (threadIdx.x < 9) {
        Hs[threadIdx.x] = Hre[(size_t)hyp * 9 + threadIdx.x];
        Hs[9 + threadIdx.x] = Him ? Him[(size_t)hyp * 9 + threadIdx.x] : 0.0;
    }
    if (threadIdx.x == 0) s_flags = (*ws.err) ? 4 : 0;
    __syncthreads();
    bool cplx = false;
#pragma unroll
    for (int k = 0; k < 9; k++) cplx |= (Hs[9 + k] != 0.0);

    int32_t* first_i = ws.first_i + (size_t)hyp * n2;
    int32_t* jstar = ws.jstar + (size_t)hyp * n1;
    double* ssd = ws.ssd + (size_t)hyp * n1;
    int flags = 0;

    for (int i = threadIdx.x; i < n1; i += RS2_THREADS) {
        double tr, tc;
        homography_point(Hs, cplx, kp1[2 * i], kp1[2 * i + 1], tr, tc);
        int best_j = -1;
        double best = 0.0;
        if (isnan(tr) || isnan(tc)) flags |= 1;
        else if (isinf(tr) || isinf(tc)) flags |= 2;
        else {
            const int r0 = py_int(tr), c0 = py_int(tc);
            const int ylo = max(0, r0 - 6), yhi = min(r0 + 5, h);
            const int xlo = max(0, c0 - 6), xhi = min(c0 + 5, w);
            if (xhi > xlo) {
                for (int yy = ylo; yy < yhi; yy++) {
                    const int s = __ldg(ws.cell_start + yy * w + xlo), e = __ldg(ws.cell_start + yy * w + xhi);
                    for (int k = s; k < e; k++) {
                        const int j = __ldg(ws.items + k);
                        const double dr = __dsub_rn(kp2[2 * j], tr), dc = __dsub_rn(kp2[2 * j + 1], tc);
                        const double v = __dadd_rn(__dmul_rn(dr, dr), __dmul_rn(dc, dc));
                        if (best_j < 0 || v < best) { best = v; best_j = j; }   // first minimum
                    }
                }
            }
        }
        if (best_j >= 0 && best < 10.0) {
            atomicMin(first_i + best_j, i);
        } else {
            best_j = -1;
        }
        jstar[i] = best_j;
        ssd[i] = best;
    }
    __syncthreads();
    int cnt = 0;
    for (int i = threadIdx.x; i < n1; i += RS2_THREADS) {
        const int j = jstar[i];
        const bool ok = (j >= 0) && (first_i[j] == i);
        cnt += ok;
        if (match_j_out) match_j_out[(size_t)hyp * n1 + i] = ok ? j : -1;
        if (match_ssd_out) match_ssd_out[(size_t)hyp * n1 + i] = ok ? ssd[i] : 0.0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    flags = __reduce_or_sync(0xffffffffu, flags);
    if ((threadIdx.x & 31) == 0) {
        s_cnt[threadIdx.x >> 5] = cnt;
        if (flags) atomicOr(&s_flags, flags);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < RS2_THREADS / 32; k++) t += s_cnt[k];
        counts_out[hyp] = t;
        if (flags_out) flags_out[hyp] = s_flags;
    }
}

// ---- K4B -----------------------------------------------------------------------------------------
// One hypothesis per lane.  The CTA streams tiles of correspondences through shared memory and every lane
// reads them by broadcast, so a correspondence costs four LDS per warp for 32 evaluations.  The tiles are
// staged by TMA: one thread issues two bulk async copies per tile (the P and Q records are contiguous,
// 8 KB each) into one of two buffers and the copy of the next tile runs while the current one is scored.
// grid.y splits the correspondences; per-hypothesis partial counts are combined with atomicAdd.
#ifndef R3D_RS3_TILE
#define R3D_RS3_TILE 64
#endif
#ifndef R3D_RS3_MINB
#define R3D_RS3_MINB 8
#endif
constexpr int RS3_THREADS = 128;
constexpr int RS3_TILE = R3D_RS3_TILE;    // correspondences per stage: 2 x 64 x 32 B = 4 KB, two stages (measured best of 64..512)

__global__ void __launch_bounds__(RS3_THREADS, R3D_RS3_MINB)
r3_score_kernel(const double* __restrict__ T, int n_hyp, const Point4* __restrict__ P, const Point4* __restrict__ Q,
                int n_pts, int pts_per_cta, double thresh, int32_t* __restrict__ counts) {
    __shared__ __align__(128) double2 sP[2][RS3_TILE * 2];     // raw 32-byte records: (x, y), (z, tag)
    __shared__ __align__(128) double2 sQ[2][RS3_TILE * 2];
    __shared__ __align__(8) unsigned long long bars[2];
    const int hyp = blockIdx.x * RS3_THREADS + threadIdx.x;
    double t[12];
#pragma unroll
    for (int k = 0; k < 12; k++) t[k] = (hyp < n_hyp) ? T[(size_t)hyp * 16 + k] : 0.0;
    const int p0 = blockIdx.y * pts_per_cta, p1 = min(n_pts, p0 + pts_per_cta);
    const int tiles = (p1 - p0 + RS3_TILE - 1) / RS3_TILE;
    const uint32_t bar0 = smem_u32(&bars[0]);
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
    }
    mbar_init_fence();
    __syncthreads();
    auto issue = [&](int tile) {      // thread 0 only
        const int base = p0 + tile * RS3_TILE, m = min(RS3_TILE, p1 - base), b = tile & 1;
        mbar_arrive_expect_tx(bar0 + 8u * b, 2u * (uint32_t)m * 32u);
        bulk_copy_g2s(smem_u32(sP[b]), P + base, (uint32_t)m * 32u, bar0 + 8u * b);
        bulk_copy_g2s(smem_u32(sQ[b]), Q + base, (uint32_t)m * 32u, bar0 + 8u * b);
    };
    if (threadIdx.x == 0 && tiles > 0) issue(0);
    int cnt = 0;
    unsigned phase = 0;
    for (int tile = 0; tile < tiles; tile++) {
        const int b = tile & 1;
        // buffer b ^ 1 was last read in tile - 1, and every thread has passed the barrier that ended it
        if (threadIdx.x == 0 && tile + 1 < tiles) issue(tile + 1);
        mbar_wait(bar0 + 8u * b, (phase >> b) & 1u);
        phase ^= 1u << b;
        const int m = min(RS3_TILE, p1 - (p0 + tile * RS3_TILE));
        const double2* sp = sP[b];
        const double2* sq = sQ[b];
#pragma unroll 4
        for (int k = 0; k < m; k++) {
            const double2 pxy = sp[2 * k];
            const double pz = sp[2 * k + 1].x;
            const double2 qxy = sq[2 * k];
            const double qz = sq[2 * k + 1].x;
            // three fused multiply-adds per coordinate, the translation as the innermost addend (16 fp64
            // instructions per evaluation instead of 19); differs from an unfused left-to-right sum by rounding only
            const double dx = fma(t[0], pxy.x, fma(t[1], pxy.y, fma(t[2], pz, t[3]))) - qxy.x;
            const double dy = fma(t[4], pxy.x, fma(t[5], pxy.y, fma(t[6], pz, t[7]))) - qxy.y;
            const double dz = fma(t[8], pxy.x, fma(t[9], pxy.y, fma(t[10], pz, t[11]))) - qz;
            const double r2 = dx * dx + dy * dy + dz * dz;
            cnt += (r2 < thresh);
        }
        __syncthreads();      // the buffer may be refilled
    }
    if (hyp < n_hyp && cnt) atomicAdd(counts + hyp, cnt);
}


// The same sweep with an fp32 FILTER and an fp64 DECISION (opt-in: R3D_OPT_K4B_FP32; counts identical to the all-fp64 kernel
// above, which stays the default because it is still the faster one: 5.55 ms against 6.0 ms on BASELINE.json configs[4],
// profiles/r02_k4b_fp32.md).  The fp64 pipe is the limiter there (78.7 % active, 16 lanes/clk/SMSP); fp32 has twice the
// lanes and FFMA issues every cycle.  Per tile the CTA converts the staged fp64 records once into fp32
// (coordinates relative to the tile's first source / target point, which keeps the operands a cloud-extent small) and
// every lane scores its hypothesis on them: x' = R p' + t' with t' = R o_p + t - o_q folded in fp64 per tile.  The fp32
// residual r2f differs from the fp64 one by at most `band` (operands rounded to fp32: 2^-24 relative each; three FFMA and
// a subtraction per coordinate; see r3_band).  r2f < thresh - band counts, r2f > thresh + band does not, and the rare
// evaluation in between is REDONE with the all-fp64 expression of the kernel above on the original records -- so every
// count is bit for bit the fp64 kernel's.
__device__ __forceinline__ float r3_band(float rows_l1_max, float pmax, float tmax, float qmax, float thresh) {
    // per coordinate: |x'_f - x'| <= u (10 (|R|_1 |p|_inf + |t'|) + 2 |q| + |d|), u = 2^-24 (rounded up to 6.1e-8), |d| <= 2 sqrt(thresh)
    // in the zone that matters; r2: 2 sqrt(3) |d| e + 3 e^2 + 5 u r2
    const float d = 2.0f * sqrtf(thresh);
    const float e = 6.1e-8f * (10.0f * (rows_l1_max * pmax + tmax) + 2.0f * qmax + d);
    return 1.5f * (3.47f * d * e + 3.0f * e * e + 3.1e-7f * 4.0f * thresh) + 1e-30f;
}

// the rare evaluation the fp32 filter cannot decide: the fp64 expression of r3_score_kernel on the original records.  Out of
// line, and reading the hypothesis from memory, so that the hot loop keeps its fp32 operands in registers (inlined, the
// twelve fp64 coefficients crowd them out and the compiler re-converts them inside the loop).
__device__ __noinline__ int r3_recheck(const double* __restrict__ t, const double2* sp, const double2* sq, int k, double thresh) {
    const double2 pxy = sp[2 * k];
    const double pz = sp[2 * k + 1].x;
    const double2 qxy = sq[2 * k];
    const double qz = sq[2 * k + 1].x;
    const double ex = fma(__ldg(t + 0), pxy.x, fma(__ldg(t + 1), pxy.y, fma(__ldg(t + 2), pz, __ldg(t + 3)))) - qxy.x;
    const double ey = fma(__ldg(t + 4), pxy.x, fma(__ldg(t + 5), pxy.y, fma(__ldg(t + 6), pz, __ldg(t + 7)))) - qxy.y;
    const double ez = fma(__ldg(t + 8), pxy.x, fma(__ldg(t + 9), pxy.y, fma(__ldg(t + 10), pz, __ldg(t + 11)))) - qz;
    return (ex * ex + ey * ey + ez * ez < thresh) ? 1 : 0;
}

// HPL hypotheses per lane: every staged point is read once (two LDS.128) and scored against HPL hypotheses held in registers
// -- at one hypothesis per lane the shared-memory pipe, not the FMA pipe, was the limit (ncu: l1tex 79 %, issue 50 %).
constexpr int RS3_HPL = 2;

__global__ void __launch_bounds__(RS3_THREADS, R3D_RS3_MINB)
r3_score_f32_kernel(const double* __restrict__ T, int n_hyp, const Point4* __restrict__ P, const Point4* __restrict__ Q,
                    int n_pts, int pts_per_cta, double thresh, int32_t* __restrict__ counts) {
    __shared__ __align__(128) double2 sP[2][RS3_TILE * 2];     // raw 32-byte records: (x, y), (z, tag)
    __shared__ __align__(128) double2 sQ[2][RS3_TILE * 2];
    __shared__ __align__(16) float4 fP[RS3_TILE], fQ[RS3_TILE];      // the current tile in fp32, relative to its first points
    __shared__ unsigned s_pmax, s_qmax;                              // max |coordinate| of fP / fQ (float bits)
    __shared__ __align__(8) unsigned long long bars[2];
    const double* t[RS3_HPL];
    float r[RS3_HPL][9], l1[RS3_HPL];
#pragma unroll
    for (int h = 0; h < RS3_HPL; h++) {
        const int hyp = (blockIdx.x * RS3_HPL + h) * RS3_THREADS + threadIdx.x;
        t[h] = T + (size_t)min(hyp, n_hyp - 1) * 16;      // (lanes past the end score the last hypothesis and drop the result)
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) r[h][i * 3 + j] = (float)__ldg(t[h] + i * 4 + j);
        l1[h] = fmaxf(fmaxf(fabsf(r[h][0]) + fabsf(r[h][1]) + fabsf(r[h][2]), fabsf(r[h][3]) + fabsf(r[h][4]) + fabsf(r[h][5])),
                      fabsf(r[h][6]) + fabsf(r[h][7]) + fabsf(r[h][8])) * 1.000001f;
    }
    const float thr = (float)thresh;
    const int p0 = blockIdx.y * pts_per_cta, p1 = min(n_pts, p0 + pts_per_cta);
    const int tiles = (p1 - p0 + RS3_TILE - 1) / RS3_TILE;
    const uint32_t bar0 = smem_u32(&bars[0]);
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8u, 1);
    }
    mbar_init_fence();
    __syncthreads();
    auto issue = [&](int tile) {      // thread 0 only
        const int base = p0 + tile * RS3_TILE, m = min(RS3_TILE, p1 - base), b = tile & 1;
        mbar_arrive_expect_tx(bar0 + 8u * b, 2u * (uint32_t)m * 32u);
        bulk_copy_g2s(smem_u32(sP[b]), P + base, (uint32_t)m * 32u, bar0 + 8u * b);
        bulk_copy_g2s(smem_u32(sQ[b]), Q + base, (uint32_t)m * 32u, bar0 + 8u * b);
    };
    if (threadIdx.x == 0 && tiles > 0) issue(0);
    int cnt[RS3_HPL];
#pragma unroll
    for (int h = 0; h < RS3_HPL; h++) cnt[h] = 0;
    unsigned phase = 0;
    for (int tile = 0; tile < tiles; tile++) {
        const int b = tile & 1;
        // buffer b ^ 1 was last read in tile - 1, and every thread has passed the barrier that ended it
        if (threadIdx.x == 0 && tile + 1 < tiles) issue(tile + 1);
        if (threadIdx.x == 0) { s_pmax = 0u; s_qmax = 0u; }
        mbar_wait(bar0 + 8u * b, (phase >> b) & 1u);
        phase ^= 1u << b;
        const int m = min(RS3_TILE, p1 - (p0 + tile * RS3_TILE));
        const double2* sp = sP[b];
        const double2* sq = sQ[b];
        // origins of the tile, and the tile in fp32 (threads 0..m-1: a source point, threads 64..64+m-1: a target point)
        const double opx = sp[0].x, opy = sp[0].y, opz = sp[1].x, oqx = sq[0].x, oqy = sq[0].y, oqz = sq[1].x;
        __syncthreads();      // (s_pmax / s_qmax are reset; fP / fQ of the previous tile are no longer read)
        {
            const int k = threadIdx.x & (RS3_TILE - 1);
            const bool src_side = threadIdx.x < RS3_TILE;
            if (k < m && threadIdx.x < 2 * RS3_TILE) {
                const double2* rec = src_side ? sp : sq;
                const float x = (float)(rec[2 * k].x - (src_side ? opx : oqx)), y = (float)(rec[2 * k].y - (src_side ? opy : oqy)),
                            z = (float)(rec[2 * k + 1].x - (src_side ? opz : oqz));
                (src_side ? fP : fQ)[k] = make_float4(x, y, z, 0.0f);
                const unsigned mx = __reduce_max_sync(__activemask(), __float_as_uint(fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z))));
                if ((threadIdx.x & 31) == 0 || k == 0) atomicMax(src_side ? &s_pmax : &s_qmax, mx);
            } else if (threadIdx.x < 2 * RS3_TILE) {
                // past the end of the last tile: a pair infinitely far apart (never counted, never undecided)
                (src_side ? fP : fQ)[k] = src_side ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : make_float4(3.0e19f, 3.0e19f, 3.0e19f, 0.0f);
            }
        }
        __syncthreads();
        // t' = R o_p + t - o_q in fp64, then fp32; the band of this tile
        float tx[RS3_HPL], ty[RS3_HPL], tz[RS3_HPL], lo[RS3_HPL], hi[RS3_HPL];
#pragma unroll
        for (int h = 0; h < RS3_HPL; h++) {
            const double* th = t[h];
            tx[h] = (float)(fma(__ldg(th + 0), opx, fma(__ldg(th + 1), opy, fma(__ldg(th + 2), opz, __ldg(th + 3)))) - oqx);
            ty[h] = (float)(fma(__ldg(th + 4), opx, fma(__ldg(th + 5), opy, fma(__ldg(th + 6), opz, __ldg(th + 7)))) - oqy);
            tz[h] = (float)(fma(__ldg(th + 8), opx, fma(__ldg(th + 9), opy, fma(__ldg(th + 10), opz, __ldg(th + 11)))) - oqz);
            const float band = r3_band(l1[h], __uint_as_float(s_pmax), fmaxf(fmaxf(fabsf(tx[h]), fabsf(ty[h])), fabsf(tz[h])) * 1.000001f,
                                       __uint_as_float(s_qmax), thr);
            lo[h] = thr - band;
            hi[h] = thr + band;
        }
        // 32 points without a branch (undecided evaluations are remembered as bits), then the rare rechecks
#pragma unroll 1
        for (int kb = 0; kb < RS3_TILE; kb += 32) {
            unsigned undecided[RS3_HPL];
#pragma unroll
            for (int h = 0; h < RS3_HPL; h++) undecided[h] = 0u;
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const float4 p = fP[kb + j], q = fQ[kb + j];
#pragma unroll
                for (int h = 0; h < RS3_HPL; h++) {
                    const float dx = fmaf(r[h][0], p.x, fmaf(r[h][1], p.y, fmaf(r[h][2], p.z, tx[h]))) - q.x;
                    const float dy = fmaf(r[h][3], p.x, fmaf(r[h][4], p.y, fmaf(r[h][5], p.z, ty[h]))) - q.y;
                    const float dz = fmaf(r[h][6], p.x, fmaf(r[h][7], p.y, fmaf(r[h][8], p.z, tz[h]))) - q.z;
                    const float r2f = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    cnt[h] += (r2f < lo[h]);
                    undecided[h] |= (r2f >= lo[h] && r2f <= hi[h]) ? (1u << j) : 0u;
                }
            }
#pragma unroll
            for (int h = 0; h < RS3_HPL; h++) {
                while (undecided[h]) {
                    const int j = __ffs(undecided[h]) - 1;
                    undecided[h] &= undecided[h] - 1u;
                    cnt[h] += r3_recheck(t[h], sp, sq, kb + j, thresh);
                }
            }
        }
        __syncthreads();      // the buffers may be refilled
    }
#pragma unroll
    for (int h = 0; h < RS3_HPL; h++) {
        const int hyp = (blockIdx.x * RS3_HPL + h) * RS3_THREADS + threadIdx.x;
        if (hyp < n_hyp && cnt[h]) atomicAdd(counts + hyp, cnt[h]);
    }
}

std::atomic<int> g_k4b_fp32{0};

}  // namespace r3d

using namespace r3d;

extern "C" size_t r3d_ransac2d_workspace_bytes(int n_hyp, int n1, int n2, int img_h, int img_w) {
    if (n_hyp <= 0 || n1 <= 0 || n2 <= 0 || img_h <= 0 || img_w <= 0) return 0;
    Ransac2dWs w;
    Arena a(nullptr, 0);
    ransac2d_layout(w, a, n_hyp, n1, n2, img_h, img_w);
    return a.used + 512;
}

extern "C" int r3d_ransac2d_score(const double* H, const double* H_imag, int n_hyp, const double* kp1, int n1,
                                  const double* kp2, int n2, int img_h, int img_w, int32_t* counts_out, int32_t* match_j_out,
                                  double* match_ssd_out, int32_t* flags_out, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    if (!H || !kp1 || !kp2 || !counts_out || n_hyp <= 0 || n1 <= 0 || n2 <= 0 || img_h <= 0 || img_w <= 0)
        return R3D_ERR_INVALID;
    if ((int64_t)img_h * img_w >= 0x7fffffffll) return R3D_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    Ransac2dWs w;
    Arena a(workspace, workspace_bytes);
    ransac2d_layout(w, a, n_hyp, n1, n2, img_h, img_w);
    if (!a.ok) return R3D_ERR_WORKSPACE;
    R3D_LAUNCH(r2_cells_kernel, (n2 + 255) / 256, 256, 0, st, kp2, n2, img_h, img_w, w);
    R3D_LAUNCH(r2_csr_kernel, 1, 1024, 0, st, n2, img_h * img_w, w);
    R3D_LAUNCH(r2_fill_kernel, 148, 256, 0, st, w.first_i, (size_t)n_hyp * n2, (int32_t)0x7fffffff);
    R3D_LAUNCH(r2_score_kernel, n_hyp, RS2_THREADS, 0, st, H, H_imag, kp1, n1, kp2, n2, img_h, img_w, w, counts_out,
               match_j_out, match_ssd_out, flags_out);
    return R3D_OK;
}

extern "C" int r3d_ransac3d_score(const double* T, int n_hyp, const double* P, const double* Q, int n_pts, double thresh,
                                  int32_t* counts_out, void* stream) {
    if (!T || !P || !Q || !counts_out || n_hyp <= 0 || n_pts < 0) return R3D_ERR_INVALID;
    if ((((uintptr_t)P) & 31) || (((uintptr_t)Q) & 31)) return R3D_ERR_INVALID;      // xyzw records: bulk copies / 256-bit loads
    cudaStream_t st = (cudaStream_t)stream;
    R3D_CUDA(cudaMemsetAsync(counts_out, 0, (size_t)n_hyp * sizeof(int32_t), st));
    if (n_pts == 0) return R3D_OK;
    const int gx = (n_hyp + RS3_THREADS - 1) / RS3_THREADS;
    // split the correspondences so that the grid is close to a whole number of waves of 148 SMs x resident CTAs
    int resident = 0;
    R3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, r3_score_kernel, RS3_THREADS, 0));
    if (resident < 1) resident = 1;
    const int slots = 148 * resident;
    const int max_gy = (n_pts + RS3_TILE - 1) / RS3_TILE;
    int gy = 1;
    double best_eff = 0.0;
    for (int cand = 1; cand <= max_gy && cand <= 64; cand++) {
        const long long ctas = (long long)gx * cand;
        const long long waves = (ctas + slots - 1) / slots;
        const double eff = (double)ctas / (double)(waves * slots);
        // prefer at least ~3 waves (load balance between CTAs that share an SM), then the fullest last wave
        if ((waves >= 3 || cand == max_gy) && eff > best_eff + 1e-9) { best_eff = eff; gy = cand; }
        if (waves >= 8) break;
    }
    int per = (n_pts + gy - 1) / gy;
    per = (per + RS3_TILE - 1) / RS3_TILE * RS3_TILE;
    gy = (n_pts + per - 1) / per;
    // non-finite or huge thresholds / hypotheses: the fp32 filter has nothing to offer, and NaN needs the fp64 comparison
    if (!g_k4b_fp32.load() || !(thresh > 1e-30 && thresh < 1e30))
        R3D_LAUNCH(r3_score_kernel, dim3(gx, gy), RS3_THREADS, 0, st, T, n_hyp, (const Point4*)P, (const Point4*)Q, n_pts, per,
                   thresh, counts_out);
    else
        R3D_LAUNCH(r3_score_f32_kernel, dim3((gx + RS3_HPL - 1) / RS3_HPL, gy), RS3_THREADS, 0, st, T, n_hyp, (const Point4*)P,
                   (const Point4*)Q, n_pts, per, thresh, counts_out);
    return R3D_OK;
}

extern "C" int r3d_homography_points(const double* H, const double* H_imag, const double* kp, int n, double* out, void* stream) {
    if (!H || !kp || !out || n < 0) return R3D_ERR_INVALID;
    if (n == 0) return R3D_OK;
    R3D_LAUNCH(r2_points_kernel, (n + 255) / 256, 256, 0, (cudaStream_t)stream, H, H_imag, kp, n, out);
    return R3D_OK;
}
