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
//   - cell_start[b*64 + f] = first sorted position of cell f of block b (lower bound, so empty
//     cells have start == next start and the end of cell f is cell_start[b*64 + f + 1])
//
// Search contract: given an upper bound (best_d2, best_idx) that is the distance to an actual
// point of the cloud, `bounded_search` visits every cell that intersects the axis-aligned box
// of half-width sqrt(best_d2) around the query and returns the lexicographic minimum of
// (squared distance, cloud index) over everything it sees.  Every point closer than the bound
// lies in that box, so the result is the exact nearest neighbour with lowest-index tie-break:
// the same answer as float64 brute force (oracle/ref_port.py nearest_neighbor_brute), which
// is what utils/icp.py:58-71 nearest_neighbor computes up to exact ties.
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
    const uint32_t* cell_start;
    const uint32_t* block_key;    // block_id -> block_linear
    const uint32_t* block_first;  // block_id -> first sorted position (n_blocks + 1 entries)
};

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

// Scan the part of block `b` (block coords cbx,cby,cbz) that intersects the fine-cell box
// [lo, hi], pruning (y,z) rows whose distance lower bound already exceeds the best.
__device__ __forceinline__ void scan_block_box(Best& best, const GridView& v, const GridParams& g, int b, int cbx, int cby,
                                               int cbz, int lox, int loy, int loz, int hix, int hiy, int hiz, double qx,
                                               double qy, double qz, double tqx, double tqy, double tqz) {
    const int x0 = max(lox, cbx * 4) - cbx * 4, x1 = min(hix, cbx * 4 + 3) - cbx * 4;
    const int y0 = max(loy, cby * 4), y1 = min(hiy, cby * 4 + 3);
    const int z0 = max(loz, cbz * 4), z1 = min(hiz, cbz * 4 + 3);
    // gap in cell units between the query and the x-range of cells that will be read; with the (y,z) gaps below it
    // prunes whole rows -- and for a query far outside the cloud, whose box covers everything, nearly all of them
    double gx = fmax(fmax((double)(cbx * 4 + x0) - tqx, tqx - (double)(cbx * 4 + x1 + 1)), 0.0);
    gx = fmax(gx - 1e-6, 0.0) * g.cell;
    {   // the block as a whole
        double by_ = fmax(fmax((double)y0 - tqy, tqy - (double)(y1 + 1)), 0.0), bz_ = fmax(fmax((double)z0 - tqz, tqz - (double)(z1 + 1)), 0.0);
        by_ = fmax(by_ - 1e-6, 0.0) * g.cell;
        bz_ = fmax(bz_ - 1e-6, 0.0) * g.cell;
        if (gx * gx + by_ * by_ + bz_ * bz_ > best.d2) return;
    }
    const uint32_t* cs = v.cell_start + (size_t)b * 64;
    for (int iz = z0; iz <= z1; iz++) {
        // gap in cell units between the query and the slab [iz, iz+1); 1e-6 cells of slack
        // covers the rounding of (p - o) * inv_cell
        double gz = fmax(fmax((double)iz - tqz, tqz - (double)(iz + 1)), 0.0);
        gz = fmax(gz - 1e-6, 0.0) * g.cell;
        for (int iy = y0; iy <= y1; iy++) {
            double gy = fmax(fmax((double)iy - tqy, tqy - (double)(iy + 1)), 0.0);
            gy = fmax(gy - 1e-6, 0.0) * g.cell;
            if (gx * gx + gz * gz + gy * gy > best.d2) continue;
            const int f = ((iz & 3) << 4) | ((iy & 3) << 2);
            const uint32_t s = __ldg(cs + f + x0);
            const uint32_t e = __ldg(cs + f + x1 + 1);
            for (uint32_t k = s; k < e; k++) consider(best, v.pts, (int)k, qx, qy, qz);
        }
    }
}

// Exact search given a valid upper bound in `best` (see file header).
__device__ __forceinline__ void bounded_search(Best& best, const GridView& v, const GridParams& g, double qx, double qy,
                                               double qz) {
    // half-width of the box, inflated by a relative 1e-12 so that rounding in (q -+ r - o)*inv
    // can never exclude a point at distance exactly sqrt(best.d2)
    const double r = sqrt(best.d2) * (1.0 + 1e-12) + 1e-300;
    const double tqx = (qx - g.ox) * g.inv_cell, tqy = (qy - g.oy) * g.inv_cell, tqz = (qz - g.oz) * g.inv_cell;
    const double rc = r * g.inv_cell * (1.0 + 1e-12) + 1e-9;
    // box entirely outside the grid: the bound itself is the answer
    if (!(tqx + rc >= 0.0 && tqx - rc < (double)g.nx && tqy + rc >= 0.0 && tqy - rc < (double)g.ny &&
          tqz + rc >= 0.0 && tqz - rc < (double)g.nz))
        return;
    const int lox = cell_coord(tqx - rc, g.nx), hix = cell_coord(tqx + rc, g.nx);
    const int loy = cell_coord(tqy - rc, g.ny), hiy = cell_coord(tqy + rc, g.ny);
    const int loz = cell_coord(tqz - rc, g.nz), hiz = cell_coord(tqz + rc, g.nz);
    const int bx0 = lox >> 2, bx1 = hix >> 2, by0 = loy >> 2, by1 = hiy >> 2, bz0 = loz >> 2, bz1 = hiz >> 2;
    const long long nbox = (long long)(bx1 - bx0 + 1) * (by1 - by0 + 1) * (bz1 - bz0 + 1);
    if (nbox <= (long long)g.n_blocks) {
        for (int cbz = bz0; cbz <= bz1; cbz++)
            for (int cby = by0; cby <= by1; cby++)
                for (int cbx = bx0; cbx <= bx1; cbx++) {
                    const uint32_t key = (uint32_t)((cbz * g.by + cby) * g.bx + cbx);
                    const int b = find_block(v.hash, g.hash_mask, key, block_hash(cbx, cby, cbz));
                    if (b >= 0) scan_block_box(best, v, g, b, cbx, cby, cbz, lox, loy, loz, hix, hiy, hiz, qx, qy, qz, tqx, tqy, tqz);
                }
    } else {
        // far query: the box covers more block slots than there are blocks; walk the block list
        for (int b = 0; b < g.n_blocks; b++) {
            const uint32_t key = __ldg(v.block_key + b);
            const int cbx = (int)(key % (uint32_t)g.bx);
            const int cby = (int)((key / (uint32_t)g.bx) % (uint32_t)g.by);
            const int cbz = (int)(key / ((uint32_t)g.bx * (uint32_t)g.by));
            if (cbx < bx0 || cbx > bx1 || cby < by0 || cby > by1 || cbz < bz0 || cbz > bz1) continue;
            scan_block_box(best, v, g, b, cbx, cby, cbz, lox, loy, loz, hix, hiy, hiz, qx, qy, qz, tqx, tqy, tqz);
        }
    }
}

__device__ __forceinline__ void scan_whole_block(Best& best, const GridView& v, int b, double qx, double qy, double qz) {
    const uint32_t s = __ldg(v.block_first + b), e = __ldg(v.block_first + b + 1);
    for (uint32_t k = s; k < e; k++) consider(best, v.pts, (int)k, qx, qy, qz);
}

// Find SOME point of the cloud to serve as the first upper bound: expanding shells of blocks
// around the query's (clamped) block until an occupied one is met; if the shells would cost
// more probes than there are blocks, take the block whose centre is nearest instead.
__device__ __forceinline__ void seed_search(Best& best, const GridView& v, const GridParams& g, double qx, double qy, double qz) {
    const int cx = cell_coord((qx - g.ox) * g.inv_cell, g.nx) >> 2;
    const int cy = cell_coord((qy - g.oy) * g.inv_cell, g.ny) >> 2;
    const int cz = cell_coord((qz - g.oz) * g.inv_cell, g.nz) >> 2;
    const int kmax = max(g.bx, max(g.by, g.bz));
    for (int k = 0; k <= kmax; k++) {
        const long long shell = (long long)(2 * k + 1) * (2 * k + 1) * (2 * k + 1);
        if (k > 1 && shell > (long long)g.n_blocks) break;
        for (int dz = -k; dz <= k; dz++) {
            const int bz = cz + dz;
            if (bz < 0 || bz >= g.bz) continue;
            for (int dy = -k; dy <= k; dy++) {
                const int by = cy + dy;
                if (by < 0 || by >= g.by) continue;
                const bool face = (dz == -k || dz == k || dy == -k || dy == k);
                const int step = face ? 1 : max(2 * k, 1);
                for (int dx = -k; dx <= k; dx += step) {
                    const int bx = cx + dx;
                    if (bx < 0 || bx >= g.bx) continue;
                    const uint32_t key = (uint32_t)((bz * g.by + by) * g.bx + bx);
                    const int b = find_block(v.hash, g.hash_mask, key, block_hash(bx, by, bz));
                    if (b >= 0) { scan_whole_block(best, v, b, qx, qy, qz); return; }
                }
            }
        }
    }
    // nearest block centre over the block list
    double bd = 1e300;
    int bb = 0;
    const double bw = 4.0 * g.cell;
    for (int b = 0; b < g.n_blocks; b++) {
        const uint32_t key = __ldg(v.block_key + b);
        const double ccx = g.ox + ((double)(key % (uint32_t)g.bx) + 0.5) * bw;
        const double ccy = g.oy + ((double)((key / (uint32_t)g.bx) % (uint32_t)g.by) + 0.5) * bw;
        const double ccz = g.oz + ((double)(key / ((uint32_t)g.bx * (uint32_t)g.by)) + 0.5) * bw;
        const double d = (ccx - qx) * (ccx - qx) + (ccy - qy) * (ccy - qy) + (ccz - qz) * (ccz - qz);
        if (d < bd) { bd = d; bb = b; }
    }
    scan_whole_block(best, v, bb, qx, qy, qz);
}

// ---------------------------------------------------------------------------------------
// Host-side description of the per-chunk grid workspace (all arrays have a fixed per-pair
// stride derived from max_dst, so kernels index them with blockIdx.y = pair).
// ---------------------------------------------------------------------------------------
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
    uint32_t* cell_start;           // [pair][64 * max_dst + 64]
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
    v.cell_start = w.cell_start + (size_t)pair * ((size_t)64 * w.max_dst + 64);
    v.block_key = w.block_key + (size_t)pair * w.max_dst;
    v.block_first = w.block_first + (size_t)pair * (w.max_dst + 1);
    return v;
}

}  // namespace r3d
