// Shared device/host helpers for the r3d_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include "../../include/r3d_b200.h"

namespace r3d {

extern std::atomic<uint64_t> g_launches;
void set_cuda_error(cudaError_t e, const char* where);

#define R3D_CUDA(call)                                                   \
    do {                                                                 \
        cudaError_t _e = (call);                                         \
        if (_e != cudaSuccess) {                                         \
            r3d::set_cuda_error(_e, #call);                              \
            return R3D_ERR_CUDA;                                         \
        }                                                                \
    } while (0)

// Kernel launch + launch-error check + launch counter.
#define R3D_LAUNCH(kernel, grid, block, smem, stream, ...)               \
    do {                                                                 \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);      \
        r3d::g_launches.fetch_add(1, std::memory_order_relaxed);         \
        cudaError_t _e = cudaPeekAtLastError();                          \
        if (_e != cudaSuccess) {                                         \
            r3d::set_cuda_error(_e, #kernel);                            \
            return R3D_ERR_CUDA;                                         \
        }                                                                \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace (256-byte aligned slices).
struct Arena {
    char* base;
    size_t cap;
    size_t used;
    bool ok;
    Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), used(0), ok(true) {
        size_t mis = ((uintptr_t)p) & 255;
        if (mis) used = 256 - mis;
    }
    template <typename T>
    T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T), 256);
        if (base == nullptr) { used += bytes; return nullptr; }   // sizing pass
        if (used + bytes > cap) { ok = false; return nullptr; }
        T* r = (T*)(base + used);
        used += bytes;
        return r;
    }
};

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
struct __align__(32) Point4 {      // one cloud point: xyz + the point's index in its cloud
    double x, y, z;
    int32_t idx;
    int32_t pad;
};

// One 32-byte record = one 256-bit read-only load (sm_100: LDG.E.256.CONSTANT).  R3D_LOAD256=0: two 128-bit loads.
#ifndef R3D_LOAD256
#define R3D_LOAD256 1
#endif
__device__ __forceinline__ void load_record(const void* p, double2& a, double2& b) {
#if R3D_LOAD256
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
#else
    const double2* q = reinterpret_cast<const double2*>(p);
    a = __ldg(q);
    b = __ldg(q + 1);
#endif
}

__device__ __forceinline__ Point4 load_point(const Point4* p) {
    double2 a, b;
    load_record(p, a, b);
    Point4 r;
    r.x = a.x; r.y = a.y; r.z = b.x;
    long long w = __double_as_longlong(b.y);
    r.idx = (int32_t)(w & 0xffffffffll);
    r.pad = 0;
    return r;
}

// One 32-byte record = one 256-bit store (sm_100: STG.E.256; the record is 32-byte aligned).  R3D_STORE256=0 falls
// back to two 128-bit stores.
#ifndef R3D_STORE256
#define R3D_STORE256 1
#endif
__device__ __forceinline__ void store_record(Point4* p, double x, double y, double z, double w) {
#if R3D_STORE256
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(x), "d"(y), "d"(z), "d"(w) : "memory");
#else
    double2* q = reinterpret_cast<double2*>(p);
    q[0] = make_double2(x, y);
    q[1] = make_double2(z, w);
#endif
}

__device__ __forceinline__ void store_point(Point4* p, double x, double y, double z, int32_t idx) {
    store_record(p, x, y, z, __longlong_as_double((long long)(uint32_t)idx));
}

// tag = cloud index (low word) | a second index (high word); load_point / consider() read the low word only
__device__ __forceinline__ void store_point_tagged(Point4* p, double x, double y, double z, int32_t idx, int32_t hi) {
    store_record(p, x, y, z, __hiloint2double(hi, idx));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- TMA (bulk async copy) plumbing: 1-D cp.async.bulk global -> shared, completion counted on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 24); spin++) {
        asm volatile(
            "{\n"
            " .reg .pred p;\n"
            " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();      // a lost byte count would otherwise hang the GPU
}

// Order-preserving map double -> uint64 (for atomicMin/atomicMax on doubles).
__device__ __forceinline__ unsigned long long enc_double(double x) {
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_double(unsigned long long e) {
    unsigned long long b = (e & 0x8000000000000000ull) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)b);
}

// ---------------------------------------------------------------------------------------
// 3x3 Kabsch solve in fp64 (one thread).  H = sum (a - ca)(b - cb)^T.
// R = V U^T from the SVD H = U S V^T, with the sign of the last right-singular vector
// flipped when det(V U^T) < 0 -- utils/icp.py:39-45, utils/r3d.py:49-56.
// Implemented as a one-sided (Hestenes) Jacobi SVD: columns of G = H V are orthogonalised;
// then R = v1 u1^T + v2 u2^T + (v1 x v2)(u1 x u2)^T, which equals the reference's
// reflection-corrected product without needing the smallest singular value's sign.
// ---------------------------------------------------------------------------------------
__device__ inline void kabsch_rotation(const double Hin[9], double R[9]) {
    double G[3][3], V[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) { G[i][j] = Hin[i * 3 + j]; V[i][j] = (i == j) ? 1.0 : 0.0; }

    // (The rotation is written with one square root, one division and one reciprocal square root -- t = tan of the Jacobi
    // angle as 2 gamma sgn(d) / (|d| + sqrt(d^2 + 4 gamma^2)), d = beta - alpha -- and the size tests on squares: this loop is a
    // single thread's dependent chain on the critical path of every ICP iteration, and fp64 sqrt / div are 25-40 instructions
    // each.  Same angles as the textbook zeta form up to rounding.)
    for (int sweep = 0; sweep < 30; sweep++) {
        bool rotated = false;
#pragma unroll
        for (int pq = 0; pq < 3; pq++) {
            const int p = (pq == 2) ? 1 : 0;
            const int q = (pq == 0) ? 1 : 2;
            double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
            for (int i = 0; i < 3; i++) {
                alpha += G[i][p] * G[i][p];
                beta += G[i][q] * G[i][q];
                gamma += G[i][p] * G[i][q];
            }
            const double ab = alpha * beta, g2 = gamma * gamma;
            // rotate while |gamma| > 1e-30 + 1e-16 sqrt(alpha beta); converged once every |gamma| <= 1e-15 sqrt(alpha beta)
            if (g2 > 1e-60 + 1e-32 * ab) {
                rotated = rotated || (g2 >= 1e-30 * ab);
                const double d = beta - alpha;
                const double t = ((d >= 0) ? 2.0 * gamma : -2.0 * gamma) / (fabs(d) + sqrt(d * d + 4.0 * g2));
                const double c = rsqrt(1.0 + t * t);
                const double s = c * t;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    double gp = G[i][p], gq = G[i][q];
                    G[i][p] = c * gp - s * gq;
                    G[i][q] = s * gp + c * gq;
                    double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - s * vq;
                    V[i][q] = s * vp + c * vq;
                }
            }
        }
        if (!rotated) break;
    }
    // singular values = column norms of G; order descending
    double sv[3];
#pragma unroll
    for (int j = 0; j < 3; j++) sv[j] = G[0][j] * G[0][j] + G[1][j] * G[1][j] + G[2][j] * G[2][j];
    int i0 = 0, i1 = 1, i2 = 2;
    if (sv[i0] < sv[i1]) { int t = i0; i0 = i1; i1 = t; }
    if (sv[i0] < sv[i2]) { int t = i0; i0 = i2; i2 = t; }
    if (sv[i1] < sv[i2]) { int t = i1; i1 = i2; i2 = t; }

    double v1[3], v2[3], u1[3], u2[3];
    // select columns without dynamic register indexing
#pragma unroll
    for (int i = 0; i < 3; i++) {
        v1[i] = (i0 == 0) ? V[i][0] : (i0 == 1) ? V[i][1] : V[i][2];
        v2[i] = (i1 == 0) ? V[i][0] : (i1 == 1) ? V[i][1] : V[i][2];
        u1[i] = (i0 == 0) ? G[i][0] : (i0 == 1) ? G[i][1] : G[i][2];
        u2[i] = (i1 == 0) ? G[i][0] : (i1 == 1) ? G[i][1] : G[i][2];
    }
    double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    if (!(n1 > 0.0)) {   // H == 0: no rotation information
#pragma unroll
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
#pragma unroll
    for (int i = 0; i < 3; i++) u1[i] /= n1;
    double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
#pragma unroll
    for (int i = 0; i < 3; i++) u2[i] -= d12 * u1[i];
    double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    if (!(n2 > 1e-300 * n1) || n2 < 1e-14 * n1) {
        // rank-1 H: any unit vector orthogonal to u1 completes the basis (LAPACK's choice is
        // equally arbitrary there); pick the one built from the smallest component of u1.
        double ax = fabs(u1[0]), ay = fabs(u1[1]), az = fabs(u1[2]);
        double e[3] = {0, 0, 0};
        if (ax <= ay && ax <= az) e[0] = 1; else if (ay <= az) e[1] = 1; else e[2] = 1;
        double d = u1[0] * e[0] + u1[1] * e[1] + u1[2] * e[2];
#pragma unroll
        for (int i = 0; i < 3; i++) u2[i] = e[i] - d * u1[i];
        n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) u2[i] /= n2;
    double u3[3] = {u1[1] * u2[2] - u1[2] * u2[1], u1[2] * u2[0] - u1[0] * u2[2], u1[0] * u2[1] - u1[1] * u2[0]};
    double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) R[i * 3 + j] = v1[i] * u1[j] + v2[i] * u2[j] + v3[i] * u3[j];
}

// moments layout shared by every reduction that feeds kabsch: 16 doubles
//   [0..2] sum a   [3..5] sum b   [6..14] sum a_i b_j (row-major)   [15] sum of distances
// with a and b already shifted by a common offset `shift` (numerical conditioning only).
// Builds the 4x4 (row-major) mapping a -> b.
__device__ inline void kabsch_from_moments(const double m[16], double n, const double shift[3], double T[16]) {
    double ca[3] = {m[0] / n, m[1] / n, m[2] / n};
    double cb[3] = {m[3] / n, m[4] / n, m[5] / n};
    double H[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) H[i * 3 + j] = m[6 + i * 3 + j] - n * ca[i] * cb[j];
    double R[9];
    kabsch_rotation(H, R);
    // centroids back in world coordinates
#pragma unroll
    for (int i = 0; i < 3; i++) { ca[i] += shift[i]; cb[i] += shift[i]; }
#pragma unroll
    for (int i = 0; i < 3; i++) {
        T[i * 4 + 0] = R[i * 3 + 0];
        T[i * 4 + 1] = R[i * 3 + 1];
        T[i * 4 + 2] = R[i * 3 + 2];
        T[i * 4 + 3] = cb[i] - (R[i * 3 + 0] * ca[0] + R[i * 3 + 1] * ca[1] + R[i * 3 + 2] * ca[2]);
    }
    T[12] = 0; T[13] = 0; T[14] = 0; T[15] = 1;
}

}  // namespace r3d
