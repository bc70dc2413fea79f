// K7: voxel down-sampling of a merged cloud -- the step right after registration.
// Replaces open3d.geometry.PointCloud.voxel_down_sample(voxel_size) as the reference calls it in the merge loop of
// chain_transformation (reconstruct.py:212-220: transform cloud i, append, down-sample with voxel_size=2).  Open3D
// (0.9.0, requirements.txt) is a third-party dependency that is not vendored; its published algorithm
// (PointCloud::VoxelDownSample) is:
//     min_bound = min over points - voxel_size / 2
//     voxel(p)  = floor((p - min_bound) / voxel_size)             per axis, as int
//     one output point per occupied voxel = (sum of its points, accumulated in index order) / count
// Open3D emits the voxels in the iteration order of a std::unordered_map (unspecified); here they come out in
// ascending (z, y, x) voxel order, and the points of a voxel are summed sequentially in ascending index order, so the
// result is a deterministic function of the input and bit-identical to the NumPy restatement in oracle/ref_port.py.
//
// HBM-bound byte work: bounding box (two-level min/max), three voxel coordinates per point, a stable LSD sort of the
// point indices by (x, then y, then z) coordinate with the batched radix sort of grid.cu (one pass per 8 bits a
// coordinate really uses: 2 passes per axis for clouds a few thousand voxels wide), head flags + scan, and one thread per
// voxel walking its run.  Positions only: the reference's merged clouds carry neither colours nor normals
// (reconstruct.py:62-77).
#include <cmath>
#include "common.cuh"
#include "grid.cuh"

namespace r3d {

constexpr int VX_THREADS = 256;
constexpr int VX_MAX_BLOCKS = 148 * 8;
constexpr int VX_TILE = 1024;       // sorted positions per CTA of the head-count / reduce kernels

struct VoxState {
    double lo[3], hi[3];
    double minb[3];
    double voxel;
    int32_t status;      // 0 ok, 1 voxel_size too small for the extent (Open3D raises), 2 non-finite coordinate
    int32_t n_vox;
    uint32_t maxc[3];
    int32_t pad;
};

__global__ void __launch_bounds__(VX_THREADS)
vx_bbox_partial(const double* __restrict__ xyz, int64_t n, double* __restrict__ partial) {
    double lo[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, hi[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
    double bad = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * VX_THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * VX_THREADS) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const double v = xyz[i * 3 + d];
            lo[d] = fmin(lo[d], v);
            hi[d] = fmax(hi[d], v);
            if (!(fabs(v) < CUDART_INF)) bad = 1.0;
        }
    }
    __shared__ double s[VX_THREADS / 32][7];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            lo[d] = fmin(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmax(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        bad = fmax(bad, __shfl_xor_sync(0xffffffffu, bad, o));
    }
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 3; d++) { s[warp][d] = lo[d]; s[warp][3 + d] = hi[d]; }
        s[warp][6] = bad;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double r = s[0][threadIdx.x];
        for (int w = 1; w < VX_THREADS / 32; w++) r = threadIdx.x < 3 ? fmin(r, s[w][threadIdx.x]) : fmax(r, s[w][threadIdx.x]);
        partial[(size_t)blockIdx.x * 7 + threadIdx.x] = r;
    }
}

__device__ __forceinline__ uint32_t voxel_coord(double v, double minb, double voxel, uint32_t maxc) {
    // Open3D: int(floor((p - min_bound) / voxel_size)); never negative, never above the coordinate of the bbox maximum
    const long long c = __double2ll_rd((v - minb) / voxel);
    return (uint32_t)min(max(c, 0ll), (long long)maxc);
}

__global__ void vx_bbox_final(const double* __restrict__ partial, int nb, int64_t n, double voxel, VoxState* st, SortDesc* desc) {
    const int t = threadIdx.x;      // 32 threads
    double r = t < 3 ? CUDART_INF : -CUDART_INF;
    if (t < 7)
        for (int b = 0; b < nb; b++) r = t < 3 ? fmin(r, partial[(size_t)b * 7 + t]) : fmax(r, partial[(size_t)b * 7 + t]);
    __shared__ double s[7];
    if (t < 7) s[t] = r;
    __syncwarp();
    if (t == 0) {
        int status = 0;
        if (s[6] != 0.0) status = 2;
        double ext = 0.0;
        for (int d = 0; d < 3; d++) {
            st->lo[d] = s[d];
            st->hi[d] = s[3 + d];
            st->minb[d] = s[d] - voxel * 0.5;
            // Open3D's own test uses the padded bounds: (max + voxel/2) - (min - voxel/2)
            ext = fmax(ext, (s[3 + d] + voxel * 0.5) - st->minb[d]);
        }
        if (status == 0 && voxel * 2147483647.0 < ext) status = 1;
        st->voxel = voxel;
        for (int d = 0; d < 3; d++) {
            uint32_t mc = 0;
            if (status == 0) mc = (uint32_t)__double2ll_rd((s[3 + d] - st->minb[d]) / voxel);
            st->maxc[d] = mc;
            const int bits = mc ? 32 - __clz(mc) : 1;
            desc[d].n = status == 0 ? (int)n : 0;
            desc[d].n_passes = (bits + 7) / 8;
        }
        st->status = status;
        st->n_vox = 0;
    }
}

__global__ void __launch_bounds__(VX_THREADS)
vx_coords(const double* __restrict__ xyz, int n, const VoxState* __restrict__ st, uint32_t* __restrict__ cx,
          uint32_t* __restrict__ cy, uint32_t* __restrict__ cz, uint32_t* __restrict__ keys0, uint32_t* __restrict__ vals0) {
    if (st->status != 0) return;
    const double vs = st->voxel;
    for (int i = blockIdx.x * VX_THREADS + threadIdx.x; i < n; i += gridDim.x * VX_THREADS) {
        const uint32_t x = voxel_coord(xyz[(size_t)i * 3], st->minb[0], vs, st->maxc[0]);
        cx[i] = x;
        cy[i] = voxel_coord(xyz[(size_t)i * 3 + 1], st->minb[1], vs, st->maxc[1]);
        cz[i] = voxel_coord(xyz[(size_t)i * 3 + 2], st->minb[2], vs, st->maxc[2]);
        keys0[i] = x;
        vals0[i] = (uint32_t)i;
    }
}

// after the sort by one axis (result in buffer prev->n_passes & 1): keys of the next axis in the current order -> buffer 0
__global__ void __launch_bounds__(VX_THREADS)
vx_next_keys(const uint32_t* __restrict__ c_next, const uint32_t* vals_a, const uint32_t* vals_b, const SortDesc* __restrict__ prev,
             uint32_t* keys0, uint32_t* vals0, int n, const VoxState* __restrict__ st) {      // vals0 may alias vals_a
    if (st->status != 0) return;
    const uint32_t* perm = (prev->n_passes & 1) ? vals_b : vals_a;
    for (int i = blockIdx.x * VX_THREADS + threadIdx.x; i < n; i += gridDim.x * VX_THREADS) {
        const uint32_t p = perm[i];
        keys0[i] = c_next[p];
        vals0[i] = p;       // same slot when perm is buffer 0: every thread rewrites only what it read
    }
}

__device__ __forceinline__ bool vx_is_head(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ cx,
                                           const uint32_t* __restrict__ cy, const uint32_t* __restrict__ cz, int i) {
    if (i == 0) return true;
    const uint32_t a = perm[i], b = perm[i - 1];
    return cx[a] != cx[b] || cy[a] != cy[b] || cz[a] != cz[b];
}

__global__ void __launch_bounds__(VX_THREADS)
vx_count_heads(const uint32_t* __restrict__ vals_a, const uint32_t* __restrict__ vals_b, const SortDesc* __restrict__ last,
               const uint32_t* __restrict__ cx, const uint32_t* __restrict__ cy, const uint32_t* __restrict__ cz, int n,
               uint32_t* __restrict__ tile_count, const VoxState* __restrict__ st) {
    if (st->status != 0) return;
    const uint32_t* perm = (last->n_passes & 1) ? vals_b : vals_a;
    int c = 0;
    for (int k = threadIdx.x; k < VX_TILE; k += VX_THREADS) {
        const int i = blockIdx.x * VX_TILE + k;
        if (i < n && vx_is_head(perm, cx, cy, cz, i)) c++;
    }
    __shared__ int s[VX_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < VX_THREADS / 32; w++) t += s[w];
        tile_count[blockIdx.x] = (uint32_t)t;
    }
}

// exclusive scan of the tile counts (one CTA of 1024 threads), total -> state and n_out
__global__ void __launch_bounds__(1024)
vx_scan_tiles(uint32_t* __restrict__ tile_count, int tiles, VoxState* st, int32_t* __restrict__ n_out, int32_t* __restrict__ status_out) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (st->status == 0) {
        for (int start = 0; start < tiles; start += 1024) {
            const int i = start + threadIdx.x;
            const uint32_t v = i < tiles ? tile_count[i] : 0u;
            uint32_t x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) warp_tot[warp] = x;
            __syncthreads();
            if (warp == 0) {
                const uint32_t w = warp_tot[lane];
                uint32_t ws = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
                    if (lane >= o) ws += y;
                }
                warp_tot[lane] = ws - w;
            }
            __syncthreads();
            const uint32_t excl = carry + warp_tot[warp] + x - v;
            if (i < tiles) tile_count[i] = excl;
            __syncthreads();
            if (threadIdx.x == 1023) carry = excl + v;
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        st->n_vox = (int32_t)carry;
        if (n_out) *n_out = (int32_t)carry;
        if (status_out) *status_out = st->status;
    }
}

// one thread per sorted position; the heads walk their voxel's run in order and write its mean
__global__ void __launch_bounds__(VX_THREADS)
vx_reduce(const double* __restrict__ xyz, const uint32_t* __restrict__ vals_a, const uint32_t* __restrict__ vals_b,
          const SortDesc* __restrict__ last, const uint32_t* __restrict__ cx, const uint32_t* __restrict__ cy,
          const uint32_t* __restrict__ cz, int n, const uint32_t* __restrict__ tile_base, double* __restrict__ out,
          const VoxState* __restrict__ st) {
    if (st->status != 0) return;
    const uint32_t* perm = (last->n_passes & 1) ? vals_b : vals_a;
    __shared__ uint32_t warp_cnt[VX_THREADS / 32];
    __shared__ uint32_t running;
    if (threadIdx.x == 0) running = tile_base[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k0 = 0; k0 < VX_TILE; k0 += VX_THREADS) {
        const int i = blockIdx.x * VX_TILE + k0 + threadIdx.x;
        const bool head = i < n && vx_is_head(perm, cx, cy, cz, i);
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = running + __popc(bal & ((1u << lane) - 1u));
        for (int w = 0; w < warp; w++) before += warp_cnt[w];
        if (head) {
            uint32_t p = perm[i];
            const uint32_t vx = cx[p], vy = cy[p], vz = cz[p];
            double sx = 0.0, sy = 0.0, sz = 0.0;
            int cnt = 0;
            int j = i;
            while (true) {
                sx += xyz[(size_t)p * 3];
                sy += xyz[(size_t)p * 3 + 1];
                sz += xyz[(size_t)p * 3 + 2];
                cnt++;
                if (++j >= n) break;
                p = perm[j];
                if (cx[p] != vx || cy[p] != vy || cz[p] != vz) break;
            }
            const double c = (double)cnt;
            out[(size_t)before * 3] = sx / c;
            out[(size_t)before * 3 + 1] = sy / c;
            out[(size_t)before * 3 + 2] = sz / c;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int w = 0; w < VX_THREADS / 32; w++) t += warp_cnt[w];
            running += t;
        }
        __syncthreads();
    }
}

struct VoxWs {
    VoxState* state;
    SortDesc* desc;      // [3]: x, y, z
    double* partial;
    uint32_t *cx, *cy, *cz;
    uint32_t* keys[2];
    uint32_t* vals[2];
    uint32_t* hist;
    uint32_t* tile_count;
    int bbox_blocks, rtiles, tiles;
};

static void vox_layout(VoxWs& w, Arena& a, int64_t n) {
    int64_t nb = (n + VX_THREADS - 1) / VX_THREADS;
    w.bbox_blocks = (int)(nb < 1 ? 1 : (nb > VX_MAX_BLOCKS ? VX_MAX_BLOCKS : nb));
    w.rtiles = radix_tiles((int)n);
    w.tiles = (int)((n + VX_TILE - 1) / VX_TILE);
    w.state = a.take<VoxState>(1);
    w.desc = a.take<SortDesc>(3);
    w.partial = a.take<double>((size_t)w.bbox_blocks * 7);
    w.cx = a.take<uint32_t>((size_t)n);
    w.cy = a.take<uint32_t>((size_t)n);
    w.cz = a.take<uint32_t>((size_t)n);
    for (int k = 0; k < 2; k++) {
        w.keys[k] = a.take<uint32_t>((size_t)n);
        w.vals[k] = a.take<uint32_t>((size_t)n);
    }
    w.hist = a.take<uint32_t>((size_t)256 * w.rtiles);
    w.tile_count = a.take<uint32_t>((size_t)w.tiles + 1);
}

}  // namespace r3d

using namespace r3d;

extern "C" size_t r3d_voxel_downsample_workspace_bytes(int64_t n) {
    if (n <= 0 || n > (1ll << 30)) return 0;
    VoxWs w;
    Arena a(nullptr, 0);
    vox_layout(w, a, n);
    return a.used + 512;
}

extern "C" int r3d_voxel_downsample(const double* xyz, int64_t n, double voxel_size, double* xyz_out, int32_t* n_out,
                                    int32_t* status_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!xyz || !xyz_out || !n_out || n <= 0 || n > (1ll << 30) || !(voxel_size > 0.0) || !std::isfinite(voxel_size))
        return R3D_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    VoxWs w;
    Arena a(workspace, workspace_bytes);
    vox_layout(w, a, n);
    if (!a.ok) return R3D_ERR_WORKSPACE;
    const int ni = (int)n;
    const int blocks = w.bbox_blocks;
    R3D_LAUNCH(vx_bbox_partial, blocks, VX_THREADS, 0, st, xyz, n, w.partial);
    R3D_LAUNCH(vx_bbox_final, 1, 32, 0, st, w.partial, blocks, n, voxel_size, w.state, w.desc);
    R3D_LAUNCH(vx_coords, blocks, VX_THREADS, 0, st, xyz, ni, w.state, w.cx, w.cy, w.cz, w.keys[0], w.vals[0]);
    int rc = radix_sort_pairs(w.keys, w.vals, w.hist, w.desc + 0, 1, ni, w.rtiles, st);
    if (rc != R3D_OK) return rc;
    R3D_LAUNCH(vx_next_keys, blocks, VX_THREADS, 0, st, w.cy, w.vals[0], w.vals[1], w.desc + 0, w.keys[0], w.vals[0], ni, w.state);
    rc = radix_sort_pairs(w.keys, w.vals, w.hist, w.desc + 1, 1, ni, w.rtiles, st);
    if (rc != R3D_OK) return rc;
    R3D_LAUNCH(vx_next_keys, blocks, VX_THREADS, 0, st, w.cz, w.vals[0], w.vals[1], w.desc + 1, w.keys[0], w.vals[0], ni, w.state);
    rc = radix_sort_pairs(w.keys, w.vals, w.hist, w.desc + 2, 1, ni, w.rtiles, st);
    if (rc != R3D_OK) return rc;
    R3D_LAUNCH(vx_count_heads, w.tiles, VX_THREADS, 0, st, w.vals[0], w.vals[1], w.desc + 2, w.cx, w.cy, w.cz, ni, w.tile_count,
               w.state);
    R3D_LAUNCH(vx_scan_tiles, 1, 1024, 0, st, w.tile_count, w.tiles, w.state, n_out, status_out);
    R3D_LAUNCH(vx_reduce, w.tiles, VX_THREADS, 0, st, xyz, w.vals[0], w.vals[1], w.desc + 2, w.cx, w.cy, w.cz, ni, w.tile_count,
               xyz_out, w.state);
    return R3D_OK;
}
