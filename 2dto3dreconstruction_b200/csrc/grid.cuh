// Sparse-block uniform voxel grid over a destination cloud, and the exact bounded
// nearest-neighbour search on it.
//
// Layout (per pair):
//   - fine cells of edge `cell`; cell (ix,iy,iz) = floor((p - origin) / cell)
//   - 4x4x4 fine cells form a block; only occupied blocks exist.  A block is found through an
//     open-addressing hash of its linear block index (x-adjacent blocks hash to adjacent slots)
//   - points are radix-sorted by key = (block_linear << 6) | (fz<<4 | fy<<2 | fx), stable, so
//     every fine cell, every x-run of fine cells inside a block, and every block is ONE
//     contiguous range of the sorted array, and points inside a cell keep ascending cloud index
//   - block_info[b] = {occupancy mask of the block's 64 cells (2 x 32 bits), first sorted position, end};
//     cell_first[first + j] = first sorted position of the block's j-th OCCUPIED cell (in cell order), and the
//     entry after the last occupied cell is the end of the block: the points of cells [f0, f1) of block b are the
//     range cell_first[first + popc(mask below f0)] .. cell_first[first + popc(mask below f1)]   (cell_range below;
//     a dense 64-entry table per block was 79 MB per 640x480 pair, this is 6 MB, and an empty run costs no load)
//
// Search contract (lean_search.cuh): given an upper bound that is the distance to an
// actual point of the cloud, every cell that intersects the ball of that radius around the query is
// visited and the lexicographic minimum of (squared distance, cloud index) is returned.  Every point
// closer than the bound lies in that ball, so the result is the exact nearest neighbour with
// lowest-index tie-break: the same answer as float64 brute force (oracle/ref_port.py
// nearest_neighbor_brute), which is what utils/icp.py:58-71 nearest_neighbor computes up to exact ties.
#pragma once
#include "common.cuh"

namespace r3d {

constexpr uint32_t HASH_EMPTY = 0xffffffffu;
constexpr int MAX_BLOCK_BITS = 26;     // key = block (<= 26 bits) << 6 | fine (6 bits)

struct GridParams {
    double ox, oy, oz;        // origin = bbox min of the cloud
    double cell, inv_cell;
    double shift[3];          // bbox centre: conditioning offset for moment sums
    int nx, ny, nz;           // fine-cell dims
    int bx, by, bz;           // block dims
    int key_bits;             // bits used by the sort key
    int n_passes;             // radix passes executed = ceil(key_bits / 8)
    int n_pts;
    int n_blocks;             // filled by the block-table kernels
    uint32_t hash_mask;       // slots - 1
    int pad;
};

struct SortDesc {             // per pair: how many keys, how many 8-bit radix passes they need
    int n;
    int n_passes;
};

struct GridView {             // raw pointers of one pair's grid
    const Point4* pts;        // sorted points (idx = position in the original cloud)
    const uint2* hash;        // {block_linear, block_id}
    const uint4* block_info;
    const uint32_t* cell_first;
    const uint32_t* block_key;    // block_id -> block_linear
    const uint32_t* block_first;  // block_id -> first sorted position (n_blocks + 1 entries)
};

// World coordinate of the first cell (index c4 = 4 * block coordinate) of a block along one axis.  Explicitly
// rounded operations: the grid build and the search must get the same bits whatever the compiler contracts.
__device__ __forceinline__ double block_origin(double o, int c4, double cell) {
    return __dadd_rn(o, __dmul_rn((double)c4, cell));
}

__device__ __forceinline__ uint32_t block_hash(int bx, int by, int bz) {
    return (uint32_t)bx + (uint32_t)by * 0x9E3779B1u + (uint32_t)bz * 0x85EBCA77u;
}

__device__ __forceinline__ int find_block(const uint2* __restrict__ hash, uint32_t mask, uint32_t key, uint32_t h) {
    uint32_t s = h & mask;
    while (true) {
        uint2 e = __ldg(hash + s);
        if (e.x == key) return (int)e.y;
        if (e.x == HASH_EMPTY) return -1;
        s = (s + 1) & mask;
    }
}

// Sorted range [s, s + cnt) of the points in cells [f0, f1) (0 <= f0 < f1 <= 64, fine-cell numbers z<<4 | y<<2 | x) of block b.
__device__ __forceinline__ void cell_range(const GridView& v, int b, int f0, int f1, uint32_t& s, uint32_t& cnt) {
    const uint4 bi = __ldg(v.block_info + b);
    const unsigned long long mask = ((unsigned long long)bi.y << 32) | (unsigned long long)bi.x;
    const unsigned long long below0 = mask & ((1ull << f0) - 1ull);
    const unsigned long long below1 = f1 >= 64 ? mask : (mask & ((1ull << f1) - 1ull));
    const int j0 = __popcll(below0), j1 = __popcll(below1);
    s = 0;
    cnt = 0;
    if (j1 > j0) {
        const uint32_t* cf = v.cell_first + bi.z;
        s = __ldg(cf + j0);
        cnt = __ldg(cf + j1) - s;
    }
}

// cell coordinate along one axis, clamped into [0, n-1]; robust to NaN/huge input
__device__ __forceinline__ int cell_coord(double t, int n) {
    t = fmin(fmax(t, 0.0), (double)(n - 1));
    return __double2int_rd(t);
}

struct Best {
    double d2;
    int idx;      // index in the original cloud
    int pos;      // position in the sorted array
};

__device__ __forceinline__ void consider(Best& b, const Point4* __restrict__ pts, int k, double qx, double qy, double qz) {
    double2 xy, zw;
    load_record(pts + k, xy, zw);
    double dx = qx - xy.x, dy = qy - xy.y, dz = qz - zw.x;
    double d2 = dx * dx + dy * dy + dz * dz;
    int idx = (int)(__double_as_longlong(zw.y) & 0xffffffffll);
    if (d2 < b.d2 || (d2 == b.d2 && idx < b.idx)) { b.d2 = d2; b.idx = idx; b.pos = k; }
}

// ---------------------------------------------------------------------------------------
// Host-side description of the per-chunk grid workspace (all arrays have a fixed per-pair
// stride derived from max_dst, so kernels index them with blockIdx.y = pair).
// ---------------------------------------------------------------------------------------
// consider() that also keeps the winner's coordinates
__device__ __forceinline__ void consider_xyz(Best& b, double& bx, double& by, double& bz, const Point4* __restrict__ pts, int k,
                                             double qx, double qy, double qz) {
    double2 xy, zw;
    load_record(pts + k, xy, zw);
    double dx = qx - xy.x, dy = qy - xy.y, dz = qz - zw.x;
    double d2 = dx * dx + dy * dy + dz * dz;
    int idx = (int)(__double_as_longlong(zw.y) & 0xffffffffll);
    if (d2 < b.d2 || (d2 == b.d2 && idx < b.idx)) { b.d2 = d2; b.idx = idx; b.pos = k; bx = xy.x; by = xy.y; bz = zw.x; }
}

struct GridWorkspace {
    int n_pairs, max_dst;
    int tiles;                      // radix tiles per pair
    size_t hash_slots;              // per pair (power of two)
    unsigned long long* bbox;       // [pair][6] encoded min xyz, max xyz
    GridParams* params;             // [pair]
    SortDesc* sort_desc;            // [pair]
    uint32_t* keys[2];              // [pair][max_dst]
    uint32_t* vals[2];
    uint32_t* hist;                 // [pair][256][tiles]
    Point4* sorted;                 // [pair][max_dst]
    uint32_t* tile_cnt;             // [pair][tiles] block-start counts per tile
    uint2* hash;                    // [pair][hash_slots]
    uint32_t* block_key;            // [pair][max_dst]
    uint32_t* block_first;          // [pair][max_dst + 1]
    uint4* block_info;              // [pair][max_dst]
    uint32_t* cell_first;           // [pair][max_dst + 1]
};

size_t grid_workspace_layout(GridWorkspace& w, Arena& a, int n_pairs, int max_dst);
int radix_sort_pairs(uint32_t* const keys[2], uint32_t* const vals[2], uint32_t* hist, const SortDesc* desc, int n_pairs,
                     int cap, int tiles, cudaStream_t st);
int radix_tiles(int cap);
extern std::atomic<double> g_cell_factor;
// Builds the grids of all pairs (async on `st`).  dst: xyzw points, offsets on device.
int grid_build(const GridWorkspace& w, const Point4* dst, const int32_t* dst_off, double cell_size, cudaStream_t st);

__device__ __forceinline__ GridView grid_view(const GridWorkspace& w, int pair) {
    GridView v;
    v.pts = w.sorted + (size_t)pair * w.max_dst;
    v.hash = w.hash + (size_t)pair * w.hash_slots;
    v.block_info = w.block_info + (size_t)pair * w.max_dst;
    v.cell_first = w.cell_first + (size_t)pair * (w.max_dst + 1);
    v.block_key = w.block_key + (size_t)pair * w.max_dst;
    v.block_first = w.block_first + (size_t)pair * (w.max_dst + 1);
    return v;
}

}  // namespace r3d
