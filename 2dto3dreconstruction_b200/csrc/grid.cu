// Grid build: bounding box, cell-size plan, cell keys, batched LSD radix sort (8-bit digits,
// stable), gather of the sorted points, block table (hash + per-block 64-entry cell starts).
// Everything is batched over the pairs of a chunk (blockIdx.y = pair) and runs without any
// host round trip: per-pair decisions (cell size, key width, table size) live in GridParams.
#include "grid.cuh"

namespace r3d {

std::atomic<double> g_cell_factor{2.0};   // automatic cell size = factor x estimated point spacing

constexpr int GB_THREADS = 256;
constexpr int RS_ROUNDS = 8;
constexpr int RS_TILE = GB_THREADS * RS_ROUNDS;    // 2048 keys per tile

// ---- bbox -----------------------------------------------------------------------------------
__global__ void bbox_init(unsigned long long* bbox, int n_pairs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pairs * 6) bbox[i] = ((i % 6) < 3) ? 0xffffffffffffffffull : 0ull;
}

__global__ void __launch_bounds__(GB_THREADS)
bbox_kernel(const Point4* __restrict__ dst, const int32_t* __restrict__ dst_off, unsigned long long* __restrict__ bbox) {
    const int pair = blockIdx.y;
    const int first = dst_off[pair], n = dst_off[pair + 1] - first;
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    bool any = false;
    for (int i = blockIdx.x * GB_THREADS + threadIdx.x; i < n; i += gridDim.x * GB_THREADS) {
        Point4 p = load_point(dst + first + i);
        mn[0] = fmin(mn[0], p.x); mn[1] = fmin(mn[1], p.y); mn[2] = fmin(mn[2], p.z);
        mx[0] = fmax(mx[0], p.x); mx[1] = fmax(mx[1], p.y); mx[2] = fmax(mx[2], p.z);
        any = true;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    any = __any_sync(0xffffffffu, any);
    if ((threadIdx.x & 31) == 0 && any) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            atomicMin(bbox + pair * 6 + k, enc_double(mn[k]));
            atomicMax(bbox + pair * 6 + 3 + k, enc_double(mx[k]));
        }
    }
}

// ---- plan -----------------------------------------------------------------------------------
__device__ inline int bits_for(unsigned long long count) {   // bits needed to index `count` values
    int b = 0;
    while ((1ull << b) < count) b++;
    return b;
}

__global__ void plan_kernel(const unsigned long long* __restrict__ bbox, const int32_t* __restrict__ dst_off,
                            double cell_size, double cell_factor, int n_pairs, GridParams* __restrict__ params,
                            SortDesc* __restrict__ desc) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= n_pairs) return;
    GridParams g;
    g.n_pts = dst_off[pair + 1] - dst_off[pair];
    double mn[3], mx[3];
    for (int k = 0; k < 3; k++) {
        mn[k] = dec_double(bbox[pair * 6 + k]);
        mx[k] = dec_double(bbox[pair * 6 + 3 + k]);
        if (g.n_pts <= 0 || !(mx[k] >= mn[k])) { mn[k] = 0; mx[k] = 0; }
    }
    double e[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
    double c = cell_size;
    if (!(c > 0.0) || !isfinite(c)) {
        // automatic: depth-image clouds are surfaces; with N points over the largest bbox face the
        // mean spacing is sqrt(area / N).  cell_factor spacings per cell (default 1.5: ~2 points per
        // occupied cell).
        double face = fmax(e[0] * e[1], fmax(e[0] * e[2], e[1] * e[2]));
        c = cell_factor * sqrt(face / (double)max(g.n_pts, 1));
        if (!(c > 0.0) || !isfinite(c)) {
            double emax = fmax(e[0], fmax(e[1], e[2]));
            c = (emax > 0.0) ? emax / 64.0 : 1.0;
        }
    }
    // keep every axis below 2^20 cells and the block count below 2^26
    for (int it = 0; it < 200; it++) {
        const double nmax = fmax(e[0], fmax(e[1], e[2])) / c;
        const double bxx = floor(e[0] / c / 4.0) + 1.0, byy = floor(e[1] / c / 4.0) + 1.0, bzz = floor(e[2] / c / 4.0) + 1.0;
        if (nmax < 1000000.0 && bxx * byy * bzz <= (double)(1ull << MAX_BLOCK_BITS)) break;
        c *= 1.25;
    }
    g.cell = c;
    g.inv_cell = 1.0 / c;
    g.ox = mn[0]; g.oy = mn[1]; g.oz = mn[2];
    g.nx = __double2int_rd(e[0] * g.inv_cell) + 1;
    g.ny = __double2int_rd(e[1] * g.inv_cell) + 1;
    g.nz = __double2int_rd(e[2] * g.inv_cell) + 1;
    g.bx = (g.nx + 3) >> 2; g.by = (g.ny + 3) >> 2; g.bz = (g.nz + 3) >> 2;
    g.key_bits = bits_for((unsigned long long)g.bx * g.by * g.bz) + 6;
    g.n_passes = (g.key_bits + 7) / 8;
    g.n_blocks = 0;
    g.hash_mask = 0;
    g.pad = 0;
    for (int k = 0; k < 3; k++) g.shift[k] = 0.5 * (mn[k] + mx[k]);
    params[pair] = g;
    desc[pair].n = g.n_pts;
    desc[pair].n_passes = g.n_passes;
}

// ---- keys -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(GB_THREADS)
keys_kernel(const Point4* __restrict__ dst, const int32_t* __restrict__ dst_off, const GridParams* __restrict__ params,
            uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, int max_dst) {
    const int pair = blockIdx.y;
    __shared__ GridParams g;
    if (threadIdx.x == 0) g = params[pair];
    __syncthreads();
    const int i = blockIdx.x * GB_THREADS + threadIdx.x;
    if (i >= g.n_pts) return;
    Point4 p = load_point(dst + dst_off[pair] + i);
    const int ix = cell_coord((p.x - g.ox) * g.inv_cell, g.nx);
    const int iy = cell_coord((p.y - g.oy) * g.inv_cell, g.ny);
    const int iz = cell_coord((p.z - g.oz) * g.inv_cell, g.nz);
    const uint32_t blk = (uint32_t)(((iz >> 2) * g.by + (iy >> 2)) * g.bx + (ix >> 2));
    const uint32_t fine = (uint32_t)(((iz & 3) << 4) | ((iy & 3) << 2) | (ix & 3));
    keys[(size_t)pair * max_dst + i] = (blk << 6) | fine;
    vals[(size_t)pair * max_dst + i] = (uint32_t)i;
}

// ---- radix sort -----------------------------------------------------------------------------
// Pass k reads buffer (k & 1) and writes buffer ((k + 1) & 1).  Passes beyond a pair's
// n_passes do nothing for that pair, so its result stays in buffer (n_passes & 1).
__global__ void __launch_bounds__(GB_THREADS)
rs_hist(const uint32_t* __restrict__ keys, const SortDesc* __restrict__ desc, uint32_t* __restrict__ hist, int pass,
        int max_dst, int tiles) {
    const int pair = blockIdx.y, tile = blockIdx.x;
    const int n = desc[pair].n;
    if (pass >= desc[pair].n_passes || tile * RS_TILE >= n) return;
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t* k = keys + (size_t)pair * max_dst;
    const int shift = pass * 8;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int i = tile * RS_TILE + r * GB_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(k[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    const int my_tiles = (n + RS_TILE - 1) / RS_TILE;
    hist[(size_t)pair * 256 * tiles + (size_t)threadIdx.x * my_tiles + tile] = h[threadIdx.x];
}

// exclusive scan over a pair's [256][tiles] histogram (digit-major), one CTA per pair: every warp sums its 8 digit
// rows (coalesced, no block barriers), the 256 row totals are scanned once, and every warp then scans its rows from
// their bases.  (The first version walked all 256 x tiles entries 1024 at a time with two block barriers per step:
// 60 us for the 2 M-key sorts of configs[3] and of the voxel down-sampling, against ~10 us.)
__global__ void __launch_bounds__(1024)
rs_scan(uint32_t* __restrict__ hist, const SortDesc* __restrict__ desc, int pass, int tiles_cap) {
    const int pair = blockIdx.x;
    if (pass >= desc[pair].n_passes) return;
    // only the tiles this pair really has take part: the histogram is [256][tiles] with the pair's
    // own tile count as row length, so the scan length follows the cloud size, not the capacity
    const int tiles = (desc[pair].n + RS_TILE - 1) / RS_TILE;
    uint32_t* h = hist + (size_t)pair * 256 * tiles_cap;
    __shared__ uint32_t row_base[256];
    __shared__ uint32_t warp_tot[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = 0; r < 8; r++) {
        const uint32_t* row = h + (size_t)(warp * 8 + r) * tiles;
        uint32_t s = 0;
        for (int j = lane; j < tiles; j += 32) s += row[j];
        s = __reduce_add_sync(0xffffffffu, s);
        if (lane == 0) row_base[warp * 8 + r] = s;
    }
    __syncthreads();
    uint32_t v = 0, x = 0;
    if (threadIdx.x < 256) {
        v = row_base[threadIdx.x];
        x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        uint32_t before = 0;
        for (int w = 0; w < warp; w++) before += warp_tot[w];
        row_base[threadIdx.x] = before + x - v;
    }
    __syncthreads();
    for (int r = 0; r < 8; r++) {
        uint32_t* row = h + (size_t)(warp * 8 + r) * tiles;
        uint32_t carry = row_base[warp * 8 + r];
        for (int j0 = 0; j0 < tiles; j0 += 32) {
            const int j = j0 + lane;
            const uint32_t c = j < tiles ? row[j] : 0u;
            uint32_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (j < tiles) row[j] = carry + incl - c;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

__global__ void __launch_bounds__(GB_THREADS)
rs_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
           uint32_t* __restrict__ vals_out, const uint32_t* __restrict__ offsets, const SortDesc* __restrict__ desc,
           int pass, int max_dst, int tiles) {
    const int pair = blockIdx.y, tile = blockIdx.x;
    const int n = desc[pair].n;
    if (pass >= desc[pair].n_passes || tile * RS_TILE >= n) return;
    constexpr int NW = GB_THREADS / 32;
    __shared__ uint32_t wcnt[NW][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int w = 0; w < NW; w++) wcnt[w][threadIdx.x] = 0;
    __syncthreads();

    const size_t pbase = (size_t)pair * max_dst;
    const int shift = pass * 8;
    // warp `warp` owns keys [tile*RS_TILE + warp*256, +256): order inside the tile is
    // (warp, round, lane) = ascending position, which makes the pass stable
    const int wbase = tile * RS_TILE + warp * (32 * RS_ROUNDS);
    uint32_t key[RS_ROUNDS], val[RS_ROUNDS], rank[RS_ROUNDS];
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int i = wbase + r * 32 + lane;
        const bool ok = i < n;
        key[r] = ok ? keys_in[pbase + i] : 0u;
        val[r] = ok ? vals_in[pbase + i] : 0u;
        const uint32_t d = (key[r] >> shift) & 255u;
        // lanes past the end form singleton groups that never touch the counters
        const unsigned grp = __match_any_sync(0xffffffffu, ok ? d : (256u + (uint32_t)lane));
        const uint32_t before = __popc(grp & ((1u << lane) - 1u));
        uint32_t old = ok ? wcnt[warp][d] : 0u;
        __syncwarp();
        if (ok && before == 0) wcnt[warp][d] = old + __popc(grp);
        __syncwarp();
        rank[r] = old + before;
    }
    __syncthreads();
    {   // digit = threadIdx.x: running base over warps, seeded with the global offset
        const uint32_t d = threadIdx.x;
        const int my_tiles = (n + RS_TILE - 1) / RS_TILE;
        uint32_t run = offsets[(size_t)pair * 256 * tiles + (size_t)d * my_tiles + tile];
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const uint32_t c = wcnt[w][d];
            wcnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int i = wbase + r * 32 + lane;
        if (i < n) {
            const uint32_t d = (key[r] >> shift) & 255u;
            const uint32_t pos = wcnt[warp][d] + rank[r];
            keys_out[pbase + pos] = key[r];
            vals_out[pbase + pos] = val[r];
        }
    }
}

// ---- gather sorted points + block starts ------------------------------------------------------
__device__ __forceinline__ const uint32_t* sorted_keys(const GridWorkspace& w, const GridParams& g, int pair) {
    return w.keys[g.n_passes & 1] + (size_t)pair * w.max_dst;
}

__global__ void __launch_bounds__(GB_THREADS)
gather_kernel(GridWorkspace w, const Point4* __restrict__ dst, const int32_t* __restrict__ dst_off) {
    const int pair = blockIdx.y, tile = blockIdx.x;
    __shared__ GridParams g;
    __shared__ uint32_t warp_cnt[GB_THREADS / 32];
    if (threadIdx.x == 0) g = w.params[pair];
    __syncthreads();
    const uint32_t* keys = sorted_keys(w, g, pair);
    const uint32_t* vals = w.vals[g.n_passes & 1] + (size_t)pair * w.max_dst;
    Point4* out = w.sorted + (size_t)pair * w.max_dst;
    const Point4* src = dst + dst_off[pair];
    uint32_t starts = 0;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int i = tile * RS_TILE + r * GB_THREADS + threadIdx.x;
        if (i < g.n_pts) {
            const uint32_t v = vals[i];
            Point4 p = load_point(src + v);
            // high word of the tag = the point's own sorted position: a raw copy of the record (TMA staging in
            // lean_search.cuh) then carries everything the search needs
            store_point_tagged(out + i, p.x, p.y, p.z, (int32_t)v, (int32_t)i);
            const uint32_t k = keys[i];
            starts += (i == 0) || ((keys[i - 1] >> 6) != (k >> 6));
        }
    }
    starts = __reduce_add_sync(0xffffffffu, starts);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = starts;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int k = 0; k < GB_THREADS / 32; k++) t += warp_cnt[k];
        w.tile_cnt[(size_t)pair * w.tiles + tile] = t;
    }
}

// one CTA per pair: exclusive scan of the tile counts (tiles <= a few hundred), n_blocks, table size,
// then clear the hash slots that will be used
__global__ void __launch_bounds__(1024)
blocks_scan_kernel(GridWorkspace w) {
    const int pair = blockIdx.x;
    uint32_t* cnt = w.tile_cnt + (size_t)pair * w.tiles;
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    __shared__ uint32_t slots_s;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < w.tiles; start += 1024) {
        const int i = start + threadIdx.x;
        const uint32_t v = (i < w.tiles) ? cnt[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t t = warp_tot[lane], ts = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, ts, o);
                if (lane >= o) ts += y;
            }
            warp_tot[lane] = ts - t;
        }
        __syncthreads();
        const uint32_t excl = carry + warp_tot[warp] + x - v;
        if (i < w.tiles) cnt[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t nb = carry;
        uint32_t slots = 64;
        while (slots < 4u * nb && slots < (uint32_t)w.hash_slots) slots <<= 1;
        GridParams* g = w.params + pair;
        g->n_blocks = (int)nb;
        g->hash_mask = slots - 1;
        slots_s = slots;
        w.block_first[(size_t)pair * (w.max_dst + 1) + nb] = (uint32_t)g->n_pts;
    }
    __syncthreads();
    uint2* hash = w.hash + (size_t)pair * w.hash_slots;
    for (uint32_t s = threadIdx.x; s < slots_s; s += 1024) hash[s] = make_uint2(HASH_EMPTY, 0u);
}

// every block start: id = (scanned tile base) + rank inside the tile; record and hash-insert
__global__ void __launch_bounds__(GB_THREADS)
blocks_assign_kernel(GridWorkspace w) {
    const int pair = blockIdx.y, tile = blockIdx.x;
    __shared__ GridParams g;
    __shared__ uint32_t cnt[RS_ROUNDS * (GB_THREADS / 32)];
    if (threadIdx.x == 0) g = w.params[pair];
    __syncthreads();
    if (tile * RS_TILE >= g.n_pts) return;
    const uint32_t* keys = sorted_keys(w, g, pair);
    constexpr int NW = GB_THREADS / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned ballots[RS_ROUNDS];
    uint32_t myk[RS_ROUNDS];
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        const int i = tile * RS_TILE + r * GB_THREADS + threadIdx.x;
        bool st = false;
        myk[r] = 0;
        if (i < g.n_pts) {
            myk[r] = keys[i];
            st = (i == 0) || ((keys[i - 1] >> 6) != (myk[r] >> 6));
        }
        ballots[r] = __ballot_sync(0xffffffffu, st);
        if (lane == 0) cnt[r * NW + warp] = __popc(ballots[r]);
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t a = cnt[2 * lane], b = cnt[2 * lane + 1], s = a + b, x = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        cnt[2 * lane] = x - s;
        cnt[2 * lane + 1] = x - s + a;
    }
    __syncthreads();
    const uint32_t tile_base = w.tile_cnt[(size_t)pair * w.tiles + tile];
    uint2* hash = w.hash + (size_t)pair * w.hash_slots;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; r++) {
        if (!((ballots[r] >> lane) & 1u)) continue;
        const int i = tile * RS_TILE + r * GB_THREADS + threadIdx.x;
        const uint32_t b = tile_base + cnt[r * NW + warp] + __popc(ballots[r] & ((1u << lane) - 1u));
        const uint32_t blk = myk[r] >> 6;
        w.block_key[(size_t)pair * w.max_dst + b] = blk;
        w.block_first[(size_t)pair * (w.max_dst + 1) + b] = (uint32_t)i;
        const int cbx = (int)(blk % (uint32_t)g.bx), cby = (int)((blk / (uint32_t)g.bx) % (uint32_t)g.by),
                  cbz = (int)(blk / ((uint32_t)g.bx * (uint32_t)g.by));
        uint32_t s = block_hash(cbx, cby, cbz) & g.hash_mask;
        while (true) {
            const uint32_t prev = atomicCAS(&hash[s].x, HASH_EMPTY, blk);
            if (prev == HASH_EMPTY) { hash[s].y = b; break; }
            s = (s + 1) & g.hash_mask;
        }
    }
}

// one warp per block: the 64 cell starts are lower bounds of (block<<6 | f) in the block's key range.  Stored compactly:
// a 64-bit occupancy mask per block, and the starts of the OCCUPIED cells only, in cell order, at cell_first[first ..
// first + n_occ) where first = the block's first sorted position (a block has at least as many points as occupied cells,
// so the block's own range of positions has room for them), followed by the end of the block -- explicitly when the range
// has a free slot, else it is the next block's first entry (= this block's end) or the global sentinel at n_pts.
__global__ void __launch_bounds__(GB_THREADS)
cell_table_kernel(GridWorkspace w) {
    const int pair = blockIdx.y;
    const GridParams* gp = w.params + pair;
    const int n_blocks = gp->n_blocks;
    const int lane = threadIdx.x & 31;
    uint4* info = w.block_info + (size_t)pair * w.max_dst;
    uint32_t* cf = w.cell_first + (size_t)pair * (w.max_dst + 1);
    if (blockIdx.x == 0 && threadIdx.x == 0) cf[gp->n_pts] = (uint32_t)gp->n_pts;   // global sentinel
    const uint32_t* keys = w.keys[gp->n_passes & 1] + (size_t)pair * w.max_dst;
    const uint32_t* bf = w.block_first + (size_t)pair * (w.max_dst + 1);
    const int warps = gridDim.x * (GB_THREADS / 32);
    for (int b = blockIdx.x * (GB_THREADS / 32) + (threadIdx.x >> 5); b < n_blocks; b += warps) {
        const uint32_t first = bf[b], last = bf[b + 1];
        const uint32_t blk = w.block_key[(size_t)pair * w.max_dst + b];
        uint32_t start[2], next[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t f = (uint32_t)(lane + 32 * h);
            const uint32_t target = (blk << 6) | f;
            uint32_t lo = first, hi = last;          // first position with key >= target
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (keys[mid] < target) lo = mid + 1; else hi = mid;
            }
            start[h] = lo;
        }
        // a cell is occupied iff the next cell starts later
        next[0] = __shfl_down_sync(0xffffffffu, start[0], 1);
        const uint32_t s32 = __shfl_sync(0xffffffffu, start[1], 0);
        if (lane == 31) next[0] = s32;
        next[1] = __shfl_down_sync(0xffffffffu, start[1], 1);
        if (lane == 31) next[1] = last;
        const unsigned m0 = __ballot_sync(0xffffffffu, next[0] > start[0]), m1 = __ballot_sync(0xffffffffu, next[1] > start[1]);
        const unsigned below = (1u << lane) - 1u;
        if ((m0 >> lane) & 1u) cf[first + __popc(m0 & below)] = start[0];
        if ((m1 >> lane) & 1u) cf[first + __popc(m0) + __popc(m1 & below)] = start[1];
        if (lane == 0) {
            const uint32_t n_occ = (uint32_t)(__popc(m0) + __popc(m1));
            if (first + n_occ < last) cf[first + n_occ] = last;
            info[b] = make_uint4(m0, m1, first, last);
        }
    }
}

// ---- host ---------------------------------------------------------------------------------------
// Batched stable LSD radix sort of 32-bit keys with 32-bit payloads: 4 passes of 8 bits; pass k reads
// buffer (k & 1) and writes the other one; a pair whose keys need fewer bits skips the top passes,
// so its result is in buffer (desc.n_passes & 1).
int radix_sort_pairs(uint32_t* const keys[2], uint32_t* const vals[2], uint32_t* hist, const SortDesc* desc, int n_pairs,
                     int cap, int tiles, cudaStream_t st) {
    for (int pass = 0; pass < 4; pass++) {
        const int in = pass & 1, out = in ^ 1;
        R3D_LAUNCH(rs_hist, dim3(tiles, n_pairs), GB_THREADS, 0, st, keys[in], desc, hist, pass, cap, tiles);
        R3D_LAUNCH(rs_scan, n_pairs, 1024, 0, st, hist, desc, pass, tiles);
        R3D_LAUNCH(rs_scatter, dim3(tiles, n_pairs), GB_THREADS, 0, st, keys[in], vals[in], keys[out], vals[out], hist, desc,
                   pass, cap, tiles);
    }
    return R3D_OK;
}

int radix_tiles(int cap) { return (cap + RS_TILE - 1) / RS_TILE; }

size_t grid_workspace_layout(GridWorkspace& w, Arena& a, int n_pairs, int max_dst) {
    w.n_pairs = n_pairs;
    w.max_dst = max_dst;
    w.tiles = (max_dst + RS_TILE - 1) / RS_TILE;
    size_t slots = 64;
    while (slots < (size_t)max_dst + (size_t)max_dst / 2) slots <<= 1;      // load factor <= 2/3 even if every point had its own block
    w.hash_slots = slots;
    const size_t P = (size_t)n_pairs, M = (size_t)max_dst;
    w.bbox = a.take<unsigned long long>(P * 6);
    w.params = a.take<GridParams>(P);
    w.sort_desc = a.take<SortDesc>(P);
    w.keys[0] = a.take<uint32_t>(P * M);
    w.keys[1] = a.take<uint32_t>(P * M);
    w.vals[0] = a.take<uint32_t>(P * M);
    w.vals[1] = a.take<uint32_t>(P * M);
    w.hist = a.take<uint32_t>(P * 256 * w.tiles);
    w.sorted = a.take<Point4>(P * M);
    w.tile_cnt = a.take<uint32_t>(P * w.tiles);
    w.hash = a.take<uint2>(P * slots);
    w.block_key = a.take<uint32_t>(P * M);
    w.block_first = a.take<uint32_t>(P * (M + 1));
    w.block_info = a.take<uint4>(P * M);
    w.cell_first = a.take<uint32_t>(P * (M + 1));
    return a.used;
}

int grid_build(const GridWorkspace& w, const Point4* dst, const int32_t* dst_off, double cell_size, cudaStream_t st) {
    const int P = w.n_pairs;
    const int pt_blocks = (w.max_dst + GB_THREADS - 1) / GB_THREADS;
    R3D_LAUNCH(bbox_init, (P * 6 + 127) / 128, 128, 0, st, w.bbox, P);
    {
        int bx = pt_blocks < 64 ? pt_blocks : 64;
        R3D_LAUNCH(bbox_kernel, dim3(bx, P), GB_THREADS, 0, st, dst, dst_off, w.bbox);
    }
    R3D_LAUNCH(plan_kernel, (P + 127) / 128, 128, 0, st, w.bbox, dst_off, cell_size, g_cell_factor.load(), P, w.params,
               w.sort_desc);
    R3D_LAUNCH(keys_kernel, dim3(pt_blocks, P), GB_THREADS, 0, st, dst, dst_off, w.params, w.keys[0], w.vals[0], w.max_dst);
    int rc = radix_sort_pairs(w.keys, w.vals, w.hist, w.sort_desc, P, w.max_dst, w.tiles, st);
    if (rc != R3D_OK) return rc;
    R3D_LAUNCH(gather_kernel, dim3(w.tiles, P), GB_THREADS, 0, st, w, dst, dst_off);
    R3D_LAUNCH(blocks_scan_kernel, P, 1024, 0, st, w);
    R3D_LAUNCH(blocks_assign_kernel, dim3(w.tiles, P), GB_THREADS, 0, st, w);
    R3D_LAUNCH(cell_table_kernel, dim3(148, P), GB_THREADS, 0, st, w);
    return R3D_OK;
}

}  // namespace r3d
