// K1: depth image -> compacted point cloud.
// Replaces utils/utils.py:21-45 (depth_to_voxel), utils/utils.py:48-76 (depth_to_voxel_ld) and
// the per-frame list comprehension of reconstruct.py:336-354, for a whole batch of frames.
//
// HBM-bound: 2 B read per pixel (uint16), 24/32 B written per valid point.  ONE pass (bp_onepass, the default): a CTA takes
// the next 4096-pixel tile off a ticket counter, reads it once, publishes its count of valid pixels and finds its output
// position by decoupled look-back over the tiles before it (a chained scan: status word = flag | global count | count
// inside the frame), then writes.  The round-1 form is kept behind R3D_BP_ONEPASS=0 and for the v8 variant:
//   bp_count   per 2048-pixel tile: number of pixels whose z = (int16)d * zscale is non-zero
//   bp_scan    one CTA: exclusive scan of all tile counts (frame-major, so the result is the
//              global output position of every tile, and of every frame)
//   bp_write   per tile: recompute, ballot-rank inside the tile, coalesced stores
// The tile is walked in rounds of 256 consecutive pixels so that consecutive threads read
// consecutive pixels and, after compaction, write consecutive points.
#include "common.cuh"
#include <type_traits>

namespace r3d {

constexpr int BP_THREADS = 256;
constexpr int BP_ROUNDS = 8;
constexpr int BP_TILE = BP_THREADS * BP_ROUNDS;

template <typename T> struct DepthTraits;
template <> struct DepthTraits<uint16_t> {
    __device__ static double as_double(uint16_t v) { return (double)v; }
    __device__ static double z_of(uint16_t v) { return (double)(int16_t)v; }   // astype(int16): wraps
};
template <> struct DepthTraits<float> {
    __device__ static double as_double(float v) { return (double)v; }
    // astype(int16) on a float truncates toward zero (x86: cvtt to int32, low 16 bits kept)
    __device__ static double z_of(float v) { return (double)(int16_t)(int32_t)v; }
};
template <> struct DepthTraits<double> {
    __device__ static double as_double(double v) { return v; }
    __device__ static double z_of(double v) { return (double)(int16_t)(int32_t)v; }
};

template <typename T>
__global__ void __launch_bounds__(BP_THREADS)
bp_count(const T* __restrict__ depth, int hw, int tiles_per_frame, double zscale, int32_t* __restrict__ tile_count) {
    const int frame = blockIdx.y;
    const int tile = blockIdx.x;
    const T* d = depth + (size_t)frame * hw;
    const int base = tile * BP_TILE;
    int c = 0;
#pragma unroll
    for (int r = 0; r < BP_ROUNDS; r++) {
        int px = base + r * BP_THREADS + threadIdx.x;
        if (px < hw) {
            double z = DepthTraits<T>::z_of(d[px]) * zscale;
            c += (z != 0.0);
        }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ int s[BP_THREADS / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < BP_THREADS / 32; w++) t += s[w];
        tile_count[(size_t)frame * tiles_per_frame + tile] = t;
    }
}

// Exclusive scan of n int32 values by one CTA of 1024 threads; writes frame offsets as well.
__global__ void __launch_bounds__(1024)
bp_scan(const int32_t* __restrict__ tile_count, int32_t* __restrict__ tile_base, int n, int tiles_per_frame,
        int frames, int32_t* __restrict__ frame_offsets) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < n; start += 1024) {
        int i = start + threadIdx.x;
        int v = (i < n) ? tile_count[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
            int ws = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += y;
            }
            warp_tot[lane] = ws - w;   // exclusive over warps
        }
        __syncthreads();
        int excl = carry + warp_tot[warp] + x - v;
        if (i < n) {
            tile_base[i] = excl;
            if (i % tiles_per_frame == 0) frame_offsets[i / tiles_per_frame] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) frame_offsets[frames] = carry;
}

// n / b for the reference's intrinsics (utils/utils.py:63-66: fx = fy = 525, cx = 319.5, cy = 239.5) on frames of
// at most 640 x 480 uint16 pixels: reciprocal multiply + one fused residual correction (Markstein) instead of the
// ~35-instruction IEEE division routine.  The result equals the correctly rounded quotient for EVERY numerator the
// domain can produce, (u - 319.5) * d and (v - 239.5) * d with d in 0..65535: checked exhaustively against NumPy by
// tests/test_gpu_backproject.py::test_fast_division_is_exact_over_the_whole_uint16_domain.  Any other intrinsics,
// size or depth type take __ddiv_rn.
__device__ __forceinline__ double div_markstein(double n, double b, double rb) {
    const double q = n * rb;
    const double e = fma(-q, b, n);
    return fma(e, rb, q);
}

#ifndef R3D_BP_MIN_CTAS
#define R3D_BP_MIN_CTAS 1
#endif
template <typename T, int MODE, int LAYOUT, bool FASTDIV>
__global__ void __launch_bounds__(BP_THREADS, R3D_BP_MIN_CTAS)
bp_write(const T* __restrict__ depth, int hw, int width, int tiles_per_frame, double fx, double fy, double cx,
         double cy, double zscale, const int32_t* __restrict__ tile_base, const int32_t* __restrict__ frame_offsets,
         void* __restrict__ out) {
    const int frame = blockIdx.y;
    const int tile = blockIdx.x;
    const T* d = depth + (size_t)frame * hw;
    const int base = tile * BP_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = BP_THREADS / 32;

    __shared__ int cnt[BP_ROUNDS * NW];   // [round][warp] valid counts -> exclusive offsets

    T raw[BP_ROUNDS];
    unsigned ballots[BP_ROUNDS];
#pragma unroll
    for (int r = 0; r < BP_ROUNDS; r++) {
        int px = base + r * BP_THREADS + threadIdx.x;
        bool ok = false;
        raw[r] = T(0);
        if (px < hw) {
            raw[r] = d[px];
            ok = (DepthTraits<T>::z_of(raw[r]) * zscale) != 0.0;
        }
        ballots[r] = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) cnt[r * NW + warp] = __popc(ballots[r]);
    }
    __syncthreads();
    if (warp == 0) {
        // 64 counts, two per lane, exclusive scan in (round, warp) order
        int a = cnt[2 * lane], b = cnt[2 * lane + 1];
        int s = a + b, x = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        int excl = x - s;
        cnt[2 * lane] = excl;
        cnt[2 * lane + 1] = excl + a;
    }
    __syncthreads();

    const int out_tile = tile_base[(size_t)frame * tiles_per_frame + tile];
    const int frame_first = frame_offsets[frame];
#pragma unroll
    for (int r = 0; r < BP_ROUNDS; r++) {
        const unsigned bal = ballots[r];
        if (!((bal >> lane) & 1u)) continue;
        const int px = base + r * BP_THREADS + threadIdx.x;
        const int pos = out_tile + cnt[r * NW + warp] + __popc(bal & ((1u << lane) - 1u));
        const int v = px / width, u = px - v * width;
        const double dv = DepthTraits<T>::as_double(raw[r]);
        const double z = DepthTraits<T>::z_of(raw[r]) * zscale;
        double x, y;
        if (MODE == R3D_BP_PINHOLE) {
            // ((u - cx) * d) / fx, in this order, IEEE double: bit-exact with utils/utils.py:72
            if (FASTDIV) {
                x = div_markstein(__dmul_rn((double)u - cx, dv), fx, 1.0 / 525.0);
                y = div_markstein(__dmul_rn((double)v - cy, dv), fy, 1.0 / 525.0);
            } else {
                x = __ddiv_rn(__dmul_rn((double)u - cx, dv), fx);
                y = __ddiv_rn(__dmul_rn((double)v - cy, dv), fy);
            }
        } else {
            x = (double)u;
            y = (double)v;
        }
        if (LAYOUT == R3D_POINTS_XYZ) {
            double* o = (double*)out + (size_t)pos * 3;
            o[0] = x; o[1] = y; o[2] = z;
        } else if (LAYOUT == R3D_POINTS_XYZW) {
            store_point((Point4*)out + pos, x, y, z, pos - frame_first);
        } else {
            ((float4*)out)[pos] = make_float4((float)x, (float)y, (float)z, 1.0f);
        }
    }
}

// ---- one pass: read, count, chained scan with decoupled look-back, write -------------------------------------------
#ifndef R3D_BP_ONEPASS
#define R3D_BP_ONEPASS 1
#endif
constexpr int OP_ROUNDS = 16;
constexpr int OP_TILE = BP_THREADS * OP_ROUNDS;      // 4096 pixels: ~118 KB of output per tile hide the look-back
// status word of a tile: [63:62] 0 = nothing yet, 1 = the tile's own count, 2 = inclusive prefix; [61:31] count over all
// frames so far; [30:0] count inside the tile's frame (for the per-frame point index of the xyzw layout)
__device__ __forceinline__ unsigned long long op_pack(unsigned flag, unsigned global_cnt, unsigned frame_cnt) {
    return ((unsigned long long)flag << 62) | ((unsigned long long)global_cnt << 31) | (unsigned long long)frame_cnt;
}

template <typename T, int MODE, int LAYOUT, bool FASTDIV>
__global__ void __launch_bounds__(BP_THREADS)
bp_onepass(const T* __restrict__ depth, int hw, int width, int tiles_per_frame, int n_tiles, double fx, double fy, double cx, double cy,
           double zscale, unsigned long long* __restrict__ status, unsigned* __restrict__ ticket, int32_t* __restrict__ frame_offsets,
           void* __restrict__ out) {
    constexpr int NW = BP_THREADS / 32;
    __shared__ int cnt[OP_ROUNDS * NW];      // [round][warp] valid counts -> exclusive offsets
    __shared__ int s_tile, s_total, s_excl_global, s_excl_frame;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);      // tiles are handed out in launch order: every predecessor is running or done
    __syncthreads();
    const int g = s_tile;
    if (g >= n_tiles) return;
    const int frame = g / tiles_per_frame, tile = g - frame * tiles_per_frame;
    const T* d = depth + (size_t)frame * hw;
    const int base = tile * OP_TILE;

    T raw[OP_ROUNDS];
    unsigned ballots[OP_ROUNDS];
#pragma unroll
    for (int r = 0; r < OP_ROUNDS; r++) {
        const int px = base + r * BP_THREADS + threadIdx.x;
        bool ok = false;
        raw[r] = T(0);
        if (px < hw) {
            raw[r] = d[px];
            ok = (DepthTraits<T>::z_of(raw[r]) * zscale) != 0.0;
        }
        ballots[r] = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) cnt[r * NW + warp] = __popc(ballots[r]);
    }
    __syncthreads();
    if (warp == 0) {
        // OP_ROUNDS * NW = 128 counts, four per lane, exclusive scan in (round, warp) order
        int c[4], s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { c[k] = cnt[4 * lane + k]; s += c[k]; }
        int x = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        int run = x - s;
#pragma unroll
        for (int k = 0; k < 4; k++) { cnt[4 * lane + k] = run; run += c[k]; }
        const int total = __shfl_sync(0xffffffffu, x, 31);
        // publish the tile's own count, then look back
        if (lane == 0) {
            s_total = total;
            if (g > 0) atomicExch(status + g, op_pack(1u, (unsigned)total, (unsigned)total));
        }
        unsigned excl_global = 0, excl_frame = 0;
        const int frame_start = frame * tiles_per_frame;
        bool in_frame = true;      // the look-back has not left this tile's frame yet
        int j0 = g - 1;
        while (j0 >= 0) {
            const int j = j0 - lane;
            unsigned long long word = op_pack(2u, 0u, 0u);      // before the first tile: an empty prefix
            if (j >= 0) {
                const volatile unsigned long long* p = status + j;
                do { word = *p; } while ((word >> 62) == 0ull);
            }
            const unsigned flag = (unsigned)(word >> 62);
            const unsigned pm = __ballot_sync(0xffffffffu, flag == 2u);
            const int stop = pm ? __ffs(pm) - 1 : 32;      // nearest tile that already knows its prefix
            const unsigned gl = (unsigned)((word >> 31) & 0x7fffffffull), fr = (unsigned)(word & 0x7fffffffull);
            const bool take = lane <= stop;
            excl_global += __reduce_add_sync(0xffffffffu, take ? gl : 0u);
            // inside the frame: own counts of the frame's tiles, or the frame part of a prefix taken inside the frame
            excl_frame += __reduce_add_sync(0xffffffffu, (take && in_frame && j >= frame_start) ? fr : 0u);
            if (j0 - min(stop, 31) < frame_start) in_frame = false;
            if (pm) break;
            j0 -= 32;
        }
        if (lane == 0) {
            s_excl_global = (int)excl_global;
            s_excl_frame = (int)excl_frame;
            __threadfence();
            atomicExch(status + g, op_pack(2u, excl_global + (unsigned)total, (tile == 0 ? 0u : excl_frame) + (unsigned)total));
            if (tile == 0) frame_offsets[frame] = (int)excl_global;
            if (g == n_tiles - 1) frame_offsets[frame + 1] = (int)excl_global + total;
        }
    }
    __syncthreads();
    const int out_tile = s_excl_global;
    const int frame_first = s_excl_global - (tile == 0 ? 0 : s_excl_frame);
#pragma unroll
    for (int r = 0; r < OP_ROUNDS; r++) {
        const unsigned bal = ballots[r];
        if (!((bal >> lane) & 1u)) continue;
        const int px = base + r * BP_THREADS + threadIdx.x;
        const int pos = out_tile + cnt[r * NW + warp] + __popc(bal & ((1u << lane) - 1u));
        const int v = px / width, u = px - v * width;
        const double dv = DepthTraits<T>::as_double(raw[r]);
        const double z = DepthTraits<T>::z_of(raw[r]) * zscale;
        double x, y;
        if (MODE == R3D_BP_PINHOLE) {
            // ((u - cx) * d) / fx, in this order, IEEE double: bit-exact with utils/utils.py:72
            if (FASTDIV) {
                x = div_markstein(__dmul_rn((double)u - cx, dv), fx, 1.0 / 525.0);
                y = div_markstein(__dmul_rn((double)v - cy, dv), fy, 1.0 / 525.0);
            } else {
                x = __ddiv_rn(__dmul_rn((double)u - cx, dv), fx);
                y = __ddiv_rn(__dmul_rn((double)v - cy, dv), fy);
            }
        } else {
            x = (double)u;
            y = (double)v;
        }
        if (LAYOUT == R3D_POINTS_XYZ) {
            double* o = (double*)out + (size_t)pos * 3;
            o[0] = x; o[1] = y; o[2] = z;
        } else if (LAYOUT == R3D_POINTS_XYZW) {
            store_point((Point4*)out + pos, x, y, z, pos - frame_first);
        } else {
            ((float4*)out)[pos] = make_float4((float)x, (float)y, (float)z, 1.0f);
        }
    }
}

// uint16 frames whose width is a multiple of 8: a thread owns 8 CONSECUTIVE pixels (one 128-bit load; they share an
// image row), ranks them with one block-wide scan of the per-thread counts and writes its points back to back, so the
// output order is still the pixel order.  Same tiles as bp_count / bp_scan.  Measured: 0.135 ms against 0.133 ms for the
// round-robin kernel above (64 frames): the kernel is bound by its 32-byte-per-point store stream, not by the depth
// loads or the ballots, so this variant is off by default (R3D_BP_V8=1 selects it; it passes the same tests).
#ifndef R3D_BP_V8
#define R3D_BP_V8 0
#endif
template <int MODE, int LAYOUT, bool FASTDIV>
__global__ void __launch_bounds__(BP_THREADS, R3D_BP_MIN_CTAS)
bp_write_v8(const uint16_t* __restrict__ depth, int hw, int width, int tiles_per_frame, double fx, double fy, double cx,
            double cy, double zscale, const int32_t* __restrict__ tile_base, const int32_t* __restrict__ frame_offsets,
            void* __restrict__ out) {
    const int frame = blockIdx.y;
    const int tile = blockIdx.x;
    const uint16_t* d = depth + (size_t)frame * hw;
    const int px0 = tile * BP_TILE + 8 * threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = BP_THREADS / 32;
    __shared__ int wtot[NW];
    uint4 raw = make_uint4(0u, 0u, 0u, 0u);
    if (px0 < hw) raw = __ldg(reinterpret_cast<const uint4*>(d + px0));      // hw is a multiple of 8: all eight in range
    const uint32_t words[4] = {raw.x, raw.y, raw.z, raw.w};
    unsigned mask = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const uint16_t v = (uint16_t)(words[k >> 1] >> ((k & 1) * 16));
        if (px0 < hw && (DepthTraits<uint16_t>::z_of(v) * zscale) != 0.0) mask |= 1u << k;
    }
    const int cnt = __popc(mask);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    int before = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) before += (w < warp) ? wtot[w] : 0;
    int pos = tile_base[(size_t)frame * tiles_per_frame + tile] + before + incl - cnt;
    const int frame_first = frame_offsets[frame];
    const int v = px0 / width, u0 = px0 - v * width;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (!((mask >> k) & 1u)) continue;
        const uint16_t rv = (uint16_t)(words[k >> 1] >> ((k & 1) * 16));
        const double dv = DepthTraits<uint16_t>::as_double(rv);
        const double z = DepthTraits<uint16_t>::z_of(rv) * zscale;
        const int u = u0 + k;
        double x, y;
        if (MODE == R3D_BP_PINHOLE) {
            // ((u - cx) * d) / fx, in this order, IEEE double: bit-exact with utils/utils.py:72
            if (FASTDIV) {
                x = div_markstein(__dmul_rn((double)u - cx, dv), fx, 1.0 / 525.0);
                y = div_markstein(__dmul_rn((double)v - cy, dv), fy, 1.0 / 525.0);
            } else {
                x = __ddiv_rn(__dmul_rn((double)u - cx, dv), fx);
                y = __ddiv_rn(__dmul_rn((double)v - cy, dv), fy);
            }
        } else {
            x = (double)u;
            y = (double)v;
        }
        if (LAYOUT == R3D_POINTS_XYZ) {
            double* o = (double*)out + (size_t)pos * 3;
            o[0] = x; o[1] = y; o[2] = z;
        } else if (LAYOUT == R3D_POINTS_XYZW) {
            store_point((Point4*)out + pos, x, y, z, pos - frame_first);
        } else {
            ((float4*)out)[pos] = make_float4((float)x, (float)y, (float)z, 1.0f);
        }
        pos++;
    }
}

template <typename T, int M, int L, bool F>
static int bp_launch_write(const T* depth, int hw, int width, int tiles, double fx, double fy, double cx, double cy, double zscale,
                           const int32_t* tile_base, const int32_t* out_offsets, void* out, dim3 grid, cudaStream_t st) {
    if (R3D_BP_V8 && std::is_same<T, uint16_t>::value && width % 8 == 0 && (((uintptr_t)depth) & 15) == 0) {
        R3D_LAUNCH((bp_write_v8<M, L, F>), grid, BP_THREADS, 0, st, (const uint16_t*)depth, hw, width, tiles, fx, fy, cx, cy, zscale,
                   tile_base, out_offsets, out);
    } else {
        R3D_LAUNCH((bp_write<T, M, L, F>), grid, BP_THREADS, 0, st, depth, hw, width, tiles, fx, fy, cx, cy, zscale, tile_base,
                   out_offsets, out);
    }
    return R3D_OK;
}

template <typename T>
static int bp_dispatch(const T* depth, int frames, int height, int width, int mode, double fx, double fy, double cx,
                       double cy, double zscale, int layout, void* out, int32_t* out_offsets, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
    const int hw = height * width;
    // the reference's hard-coded intrinsics on a uint16 frame no larger than 640 x 480: exact fast division
    const bool fast = std::is_same<T, uint16_t>::value && mode == R3D_BP_PINHOLE && fx == 525.0 && fy == 525.0 && cx == 319.5 &&
                      cy == 239.5 && width <= 640 && height <= 480;
    if (R3D_BP_ONEPASS && !R3D_BP_V8) {
        const int tpf = (hw + OP_TILE - 1) / OP_TILE, n_tiles = frames * tpf;
        Arena a1(ws, ws_bytes);
        unsigned long long* status = a1.take<unsigned long long>((size_t)n_tiles);
        unsigned* ticket = a1.take<unsigned>(1);
        if (!a1.ok) return R3D_ERR_WORKSPACE;
        R3D_CUDA(cudaMemsetAsync(status, 0, (size_t)((char*)(ticket + 1) - (char*)status), st));
#define OP_CASE(M, L)                                                                                                                \
    if (mode == M && layout == L) {                                                                                                  \
        if (fast)                                                                                                                    \
            R3D_LAUNCH((bp_onepass<T, M, L, true>), n_tiles, BP_THREADS, 0, st, depth, hw, width, tpf, n_tiles, fx, fy, cx, cy, zscale, \
                       status, ticket, out_offsets, out);                                                                           \
        else                                                                                                                         \
            R3D_LAUNCH((bp_onepass<T, M, L, false>), n_tiles, BP_THREADS, 0, st, depth, hw, width, tpf, n_tiles, fx, fy, cx, cy, zscale, \
                       status, ticket, out_offsets, out);                                                                           \
        return R3D_OK;                                                                                                               \
    }
        OP_CASE(R3D_BP_PIXEL, R3D_POINTS_XYZ)
        OP_CASE(R3D_BP_PIXEL, R3D_POINTS_XYZW)
        OP_CASE(R3D_BP_PIXEL, R3D_POINTS_XYZW_F32)
        OP_CASE(R3D_BP_PINHOLE, R3D_POINTS_XYZ)
        OP_CASE(R3D_BP_PINHOLE, R3D_POINTS_XYZW)
        OP_CASE(R3D_BP_PINHOLE, R3D_POINTS_XYZW_F32)
#undef OP_CASE
        return R3D_ERR_INVALID;
    }
    const int tiles = (hw + BP_TILE - 1) / BP_TILE;
    Arena a(ws, ws_bytes);
    int32_t* tile_count = a.take<int32_t>((size_t)frames * tiles);
    int32_t* tile_base = a.take<int32_t>((size_t)frames * tiles);
    if (!a.ok) return R3D_ERR_WORKSPACE;
    dim3 grid(tiles, frames);
    R3D_LAUNCH((bp_count<T>), grid, BP_THREADS, 0, st, depth, hw, tiles, zscale, tile_count);
    R3D_LAUNCH(bp_scan, 1, 1024, 0, st, tile_count, tile_base, frames * tiles, tiles, frames, out_offsets);
#define BP_CASE(M, L)                                                                                                          \
    if (mode == M && layout == L) {                                                                                            \
        if (fast)                                                                                                              \
            return bp_launch_write<T, M, L, true>(depth, hw, width, tiles, fx, fy, cx, cy, zscale, tile_base, out_offsets, out, \
                                                  grid, st);                                                                   \
        return bp_launch_write<T, M, L, false>(depth, hw, width, tiles, fx, fy, cx, cy, zscale, tile_base, out_offsets, out,    \
                                               grid, st);                                                                      \
    }
    BP_CASE(R3D_BP_PIXEL, R3D_POINTS_XYZ)
    BP_CASE(R3D_BP_PIXEL, R3D_POINTS_XYZW)
    BP_CASE(R3D_BP_PIXEL, R3D_POINTS_XYZW_F32)
    BP_CASE(R3D_BP_PINHOLE, R3D_POINTS_XYZ)
    BP_CASE(R3D_BP_PINHOLE, R3D_POINTS_XYZW)
    BP_CASE(R3D_BP_PINHOLE, R3D_POINTS_XYZW_F32)
#undef BP_CASE
    return R3D_ERR_INVALID;
}

}  // namespace r3d

extern "C" size_t r3d_backproject_workspace_bytes(int frames, int height, int width) {
    if (frames <= 0 || height <= 0 || width <= 0) return 0;
    const size_t tiles = ((size_t)height * width + r3d::BP_TILE - 1) / r3d::BP_TILE;
    return 2 * r3d::align_up((size_t)frames * tiles * sizeof(int32_t), 256) + 256;
}

extern "C" int r3d_backproject(const void* depth, int depth_dtype, int frames, int height, int width, int mode,
                               double fx, double fy, double cx, double cy, double zscale, int out_layout,
                               void* out_points, int32_t* out_offsets, void* workspace, size_t workspace_bytes,
                               void* stream) {
    if (!depth || !out_points || !out_offsets || frames <= 0 || height <= 0 || width <= 0) return R3D_ERR_INVALID;
    if ((int64_t)frames * height * width > 0x7fffffffll) return R3D_ERR_INVALID;
    if (frames > 65535) return R3D_ERR_INVALID;
    if (out_layout == R3D_POINTS_XYZW && (((uintptr_t)out_points) & 31)) return R3D_ERR_INVALID;      // 256-bit stores
    cudaStream_t st = (cudaStream_t)stream;
    switch (depth_dtype) {
        case R3D_DEPTH_U16:
            return r3d::bp_dispatch((const uint16_t*)depth, frames, height, width, mode, fx, fy, cx, cy, zscale,
                                    out_layout, out_points, out_offsets, workspace, workspace_bytes, st);
        case R3D_DEPTH_F32:
            return r3d::bp_dispatch((const float*)depth, frames, height, width, mode, fx, fy, cx, cy, zscale,
                                    out_layout, out_points, out_offsets, workspace, workspace_bytes, st);
        case R3D_DEPTH_F64:
            return r3d::bp_dispatch((const double*)depth, frames, height, width, mode, fx, fy, cx, cy, zscale,
                                    out_layout, out_points, out_offsets, workspace, workspace_bytes, st);
        default:
            return R3D_ERR_INVALID;
    }
}
