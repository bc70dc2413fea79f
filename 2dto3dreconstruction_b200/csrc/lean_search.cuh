// Warp-union exact nearest-neighbour search on the sparse-block grid.
//
// Contract: exact NN, lowest-index tie-break, equal to float64 brute force (utils/icp.py:58-71 up to
// exact ties).  A warp owns 32 consecutive queries of the
// spatially sorted source cloud, shares ONE candidate stream, and every lane evaluates every
// candidate that survives a box filter:
//   1. every lane turns its upper bound into a ball and the ball into a box of cells; the warp
//      takes the union box (REDUX min/max) and the union of the balls' bounding boxes in world
//      coordinates (the filter box);
//   2. the (y,z) rows of the union box, cut at block borders, are the warp's segments: 32 of them
//      are looked up at a time, one per lane (one hash probe + two cell-start loads), and each lane
//      keeps the sorted-cloud range [s, s+cnt) of its segment in registers;
//   3. the concatenation of the segments is the candidate stream.  Per batch of 32 stream indices
//      every lane resolves one index to a sorted position (binary search over the warp's prefix
//      sums, by SHFL) and loads THAT point: consecutive lanes read consecutive 32-byte records
//      (coalesced, 32 independent loads in flight) and the next batch's loads are issued before
//      the current one is consumed.  Points outside the filter box are dropped (they are outside
//      every lane's ball); the survivors are ballot-compacted into a small per-warp shared-memory
//      list;
//   4. when the list holds enough points every lane scans it by broadcast LDS and keeps the
//      lexicographic minimum of (squared distance, cloud index) in fp64.
// A lane may evaluate points it did not need (they are real points of the cloud, so any of them
// is a valid bound), which is why nothing in the scan is predicated per lane.
// Boxes with many row segments are walked block by block (one segment = one block's points); boxes of
// more than LEAN_BLOCKS_MAX blocks (far outliers, queries far outside the cloud, badly misaligned pairs) are
// handled one query at a time by the whole warp: blocks are pruned by their distance to the query and the
// survivors scanned 32 points at a time (far_query_search).
#pragma once
#include "grid.cuh"

// Candidate stream staged into shared memory by cp.async.bulk (1) or by per-lane LDG (0).  The TMA path is
// correct and tested (build with R3D_LEAN_TMA=1) but measured 15% slower, 333 vs 289 us per launch
// (profiles/r01_nn_sweeps.md, sweep_12): a segment is ~5 points (160 B), every copy has its own source,
// destination and size, and UBLKCP takes them from uniform registers, so the ~30 copies of a task are issued
// one lane at a time (ELECT + 5 R2UR + UBLKCP + branch each) -- more instructions than the binary search and
// the two coalesced LDG.128 per lane they replace, in a kernel that is issue bound.
#ifndef R3D_LEAN_TMA
#define R3D_LEAN_TMA 0
#endif

namespace r3d {

#ifndef R3D_LEAN_ROW_SEGS_MAX
#define R3D_LEAN_ROW_SEGS_MAX 96
#endif
#ifndef R3D_LEAN_BLOCKS_MAX
#define R3D_LEAN_BLOCKS_MAX 256
#endif
constexpr int LEAN_ROW_SEGS_MAX = R3D_LEAN_ROW_SEGS_MAX;  // boxes with more row segments than this are walked block by block
constexpr int LEAN_BLOCKS_MAX = R3D_LEAN_BLOCKS_MAX;      // ... and boxes with more blocks than this go to the per-lane search
constexpr int LEAN_LIST_CAP = 64;      // survivors list: at most LEAN_FLUSH - 1 + 32 entries between flushes
#ifndef R3D_LEAN_FLUSH
#define R3D_LEAN_FLUSH 24
#endif
constexpr int LEAN_FLUSH = R3D_LEAN_FLUSH;      // scan the list as soon as it holds this many points (<= 32)
// shared memory of one warp, in double2: survivor list (xy + zw), and with TMA staging two buffers of 32 raw
// records plus their two mbarriers
constexpr int LEAN_WARP_SMEM = 2 * LEAN_LIST_CAP + (R3D_LEAN_TMA ? 128 + 2 : 0);
#ifndef R3D_FULL_MASK_DEFINED
#define R3D_FULL_MASK_DEFINED
constexpr unsigned FULL = 0xffffffffu;
#endif

// Diagnostic counters (builds with -DR3D_NN_STATS only; read by r3d_nn_stats_read):
//   0 tasks  1 search rounds  2 candidate evaluations per lane  3 lanes whose search was skipped
//   4 segment lookup rounds   5 segments looked up              6 non-empty segments
//   7 skipped lanes whose seed was beaten by an evaluated point (a violated bound: must stay 0)
//   8 single-query (far) searches  9 their block-list sweep steps  10 their 32-point scan steps  11 block-granular walks
//   12 sum of task durations (clock64)  13 longest task (clock64)          (8..31: r3d_nn_stats_read_ext)
//   group search: 16 tasks searched by groups  17 group passes (4 queries each)  18 queries sent on to the union walk
//   19 candidate steps (8 points per group each)
#ifdef R3D_NN_STATS
static __device__ unsigned long long g_nn_stats[32];     // this header is included by icp.cu only
#define R3D_STAT_MAX(slot, n)                                                               \
    do {                                                                                    \
        const unsigned long long _v = (unsigned long long)(n);                              \
        if ((threadIdx.x & 31) == 0) atomicMax(&g_nn_stats[slot], _v);                      \
    } while (0)
#define R3D_STAT(slot, n)                                                                   \
    do {                                                                                    \
        const unsigned long long _v = (unsigned long long)(n); /* evaluated by every lane */ \
        if ((threadIdx.x & 31) == 0) atomicAdd(&g_nn_stats[slot], _v);                      \
    } while (0)
#else
#define R3D_STAT(slot, n) do { } while (0)
#define R3D_STAT_MAX(slot, n) do { } while (0)
#endif

// one candidate of the list: xy = (x, y), zw = (z, cloud index | sorted position << 32).
// `other` is the lexicographic minimum of (squared distance, cloud index) over every evaluated point
// EXCEPT the lane's seed (sorted position seed_pos; -1 = none): the seed's exact distance is known
// before the scan, and keeping the two apart is what yields the distance to the second-nearest point
// whenever the seed stays the nearest (lean_search, "budget").
// (ptxas turns the update into predicated moves whatever the source says -- a vote-guarded branch was
// measured slower, profiles/r01_nn_sweeps.md -- so the test is kept minimal)
__device__ __forceinline__ void eval_staged(Best& other, int seed_pos, const double2 xy, const double2 zw, double qx, double qy,
                                            double qz) {
    const double dx = qx - xy.x, dy = qy - xy.y, dz = qz - zw.x;
    const double d2 = dx * dx + dy * dy + dz * dz;
    const int idx = __double2loint(zw.y), pos = __double2hiint(zw.y);
    if (pos != seed_pos && (d2 < other.d2 || (d2 == other.d2 && idx < other.idx))) {
        other.d2 = d2;
        other.idx = idx;
        other.pos = pos;
    }
}

struct WarpList {
    double2* sxy;      // [LEAN_LIST_CAP] (x, y)
    double2* szw;      // [LEAN_LIST_CAP] (z, tag)
    int n;             // warp-uniform
    int nrep, rep;     // small pairs: every query sits on nrep = 4 (or 2) lanes, replica rep = lane / (32 / nrep); else nrep = 1, rep = 0
    double lo[3], hi[3];      // filter box (world coordinates)
    float4* sq;               // [32] shared: every lane's query in cell coordinates (xyz) and the radius it asked for in cells
                              // (w, margins included; < 0: the lane is not served by this round) -- for the block-granular walk
#if R3D_LEAN_TMA
    const double2* stage;     // two staging buffers of 32 raw 32-byte records each, filled by cp.async.bulk
    uint32_t stage_addr;      // shared-window address of `stage`
    uint32_t mbar0;           // shared-window address of buffer 0's mbarrier; buffer 1's is 8 bytes further
    unsigned phase;           // bit b: parity the next wait on mbarrier b uses
#endif
};

// Every lane scans the list, four entries at a time.  Slots past n hold points of earlier batches (or
// the point the list was initialised with): real points of the same cloud, harmless to re-evaluate.
__device__ __forceinline__ void list_flush(WarpList& L, Best& best, int seed_pos, double qx, double qy, double qz) {
    __syncwarp();
    const int n = L.n;
    R3D_STAT(2, (n + 3) & ~3);
    if (L.nrep > 1) {
        // replica r of a query takes entries r, r + nrep, ...: a quarter (half) of the scan each (lean_scan_box merges them)
        const int nrep = L.nrep;
        if (nrep <= 4) {
#pragma unroll 1
            for (int u = L.rep; u < n; u += 2 * nrep) {      // reads up to slot n - 1 + nrep < LEAN_LIST_CAP: a real point
                const double2 a0 = L.sxy[u], b0 = L.szw[u], a1 = L.sxy[u + nrep], b1 = L.szw[u + nrep];
                eval_staged(best, seed_pos, a0, b0, qx, qy, qz);
                eval_staged(best, seed_pos, a1, b1, qx, qy, qz);
            }
        } else {
#pragma unroll 1
            for (int u = L.rep; u < n; u += nrep) eval_staged(best, seed_pos, L.sxy[u], L.szw[u], qx, qy, qz);
        }
        __syncwarp();
        L.n = 0;
        return;
    }
#pragma unroll 1
    for (int u = 0; u < n; u += 4) {
        const double2 a0 = L.sxy[u], b0 = L.szw[u], a1 = L.sxy[u + 1], b1 = L.szw[u + 1];
        const double2 a2 = L.sxy[u + 2], b2 = L.szw[u + 2], a3 = L.sxy[u + 3], b3 = L.szw[u + 3];
        eval_staged(best, seed_pos, a0, b0, qx, qy, qz);
        eval_staged(best, seed_pos, a1, b1, qx, qy, qz);
        eval_staged(best, seed_pos, a2, b2, qx, qy, qz);
        eval_staged(best, seed_pos, a3, b3, qx, qy, qz);
    }
    __syncwarp();
    L.n = 0;
}

// Fill every slot of a warp's list with a real point of the cloud (once per kernel and pair).
__device__ __forceinline__ void list_init(double2* list, const Point4* __restrict__ pts) {
    const int lane = threadIdx.x & 31;
    const double2* p = reinterpret_cast<const double2*>(pts);
    const double2 xy = __ldg(p);
    const double2 zw = __ldg(p + 1);      // the tag carries the cloud index and the sorted position (0)
#pragma unroll
    for (int k = lane; k < LEAN_LIST_CAP; k += 32) {
        list[k] = xy;
        list[LEAN_LIST_CAP + k] = zw;
    }
    __syncwarp();
}

// A block-granular box is mostly corners no ball reaches (the box of a ball of 7 cells has 2.6x its volume, and the cloud is a
// surface): a block is kept only if its cube comes within the radius of one of the served queries.  sq[j] = query j in cell
// coordinates + the radius it asked for in cells (margins of the box computation included; < 0: not served by this round);
// with replicated queries (small pairs) the n_q distinct ones sit on lanes 0 .. n_q - 1.  Float arithmetic: the test only has
// to be conservative.  Out of line: block-granular walks are the exception, and inlined the loop costs the common row walk
// registers.
__device__ __noinline__ bool block_reached(const float4* sq, int n_q, int cbx, int cby, int cbz) {
    const float cx0 = (float)(4 * cbx), cy0 = (float)(4 * cby), cz0 = (float)(4 * cbz);
    bool reached = false;
#pragma unroll 4
    for (int j = 0; j < n_q; j++) {
        const float4 a = sq[j];      // (broadcast)
        const float dx = fmaxf(fmaxf(cx0 - a.x, a.x - (cx0 + 4.0f)), 0.0f), dy = fmaxf(fmaxf(cy0 - a.y, a.y - (cy0 + 4.0f)), 0.0f),
                    dz = fmaxf(fmaxf(cz0 - a.z, a.z - (cz0 + 4.0f)), 0.0f);
        reached = reached || (a.w >= 0.0f && fmaf(dz, dz, fmaf(dy, dy, dx * dx)) <= a.w * a.w);
    }
    return reached;
}

// Walk every cell of the warp's box [lo, hi] (warp-uniform ints; size limits: see lean_search).
// PRUNE: block-granular walks keep only the blocks some served query's ball reaches (block_reached).  Compiled in for the
// launches whose pairs are all small (the reference's sub-sampled, often badly aligned clouds, where such walks are the
// rule: 15.2 -> 10.0 ms for the sofa pair's 100 iterations); in the kernel that serves the large pairs the extra live
// state costs the common row walk 2 % (250 -> 256 us per launch on the bench workload) and the walks are rare.
template <bool PRUNE>
__device__ __forceinline__ void lean_scan_box(WarpList& L, Best& best, int seed_pos, const GridView& v, const GridParams& g, int lox,
                                              int loy, int loz, int hix, int hiy, int hiz, double qx, double qy, double qz) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int ny = hiy - loy + 1, nz = hiz - loz + 1;
    const int bx0 = lox >> 2, bx1 = hix >> 2;
    const int nxb = bx1 - bx0 + 1;
    const int nseg = ny * nz * nxb;
    // Large boxes (misaligned clouds, outliers: the ball is many cells wide) are mostly empty rows.  They are walked
    // block by block instead: one segment = all points of a block (contiguous in the sorted cloud), an order of
    // magnitude fewer lookups, and the filter box drops the extra points when they are loaded.
    const int by0 = loy >> 2, nby = (hiy >> 2) - by0 + 1, bz0 = loz >> 2, nbz = (hiz >> 2) - bz0 + 1;
    const int nblk = nxb * nby * nbz;
    const bool by_block = nseg > LEAN_ROW_SEGS_MAX;      // warp-uniform
    R3D_STAT(11, by_block ? 1 : 0);
    const int n_units = by_block ? nblk : nseg;
    // exact for the small integers involved (a few hundred units): quotient by reciprocal, then fix up
    const int d0 = nxb, d1 = by_block ? nby : ny;
    const float inv_d0 = 1.0f / (float)d0, inv_d1 = 1.0f / (float)d1;
    R3D_STAT(5, n_units);
    for (int base = 0; base < n_units; base += 32) {
        R3D_STAT(4, 1);
        const int sidx = base + lane;
        uint32_t s = 0, cnt = 0;
        if (sidx < n_units) {
            // sidx = (k2 * d1 + k1) * d0 + k0
            int row = (int)((float)sidx * inv_d0);
            row += (sidx - row * d0 >= d0) - (sidx - row * d0 < 0);
            const int k0 = sidx - row * d0;
            int k2 = (int)((float)row * inv_d1);
            k2 += (row - k2 * d1 >= d1) - (row - k2 * d1 < 0);
            const int k1 = row - k2 * d1;
            const int cbx = bx0 + k0;
            const int cby = by_block ? by0 + k1 : (loy + k1) >> 2, cbz = by_block ? bz0 + k2 : (loz + k2) >> 2;
            const uint32_t key = (uint32_t)((cbz * g.by + cby) * g.bx + cbx);
            // (block-granular walk: only blocks that some served query's ball reaches; shared-memory reads only, so the
            // call may sit in divergent code)
            int b = -1;
            bool wanted = true;
            if constexpr (PRUNE) wanted = !by_block || block_reached(L.sq, 32 / L.nrep, cbx, cby, cbz);      // (shared-memory reads only: fine in divergent code)
            if (wanted) b = find_block(v.hash, g.hash_mask, key, block_hash(cbx, cby, cbz));
            if (b >= 0) {
                if (by_block) {
                    s = __ldg(v.block_first + b);
                    cnt = __ldg(v.block_first + b + 1) - s;
                } else {
                    const int iy = loy + k1, iz = loz + k2;
                    const int x0 = max(lox, cbx * 4) - cbx * 4, x1 = min(hix, cbx * 4 + 3) - cbx * 4;
                    const int f = ((iz & 3) << 4) | ((iy & 3) << 2);
                    cell_range(v, b, f + x0, f + x1 + 1, s, cnt);
                }
            }
        }
        // candidate stream = concatenation of the 32 segments; stream index t lives in the segment j
        // with excl_j <= t < excl_j + cnt_j (exclusive prefix sums of cnt)
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += y;
        }
        const uint32_t excl = incl - cnt;
        const uint32_t total = __shfl_sync(FULL, incl, 31);
        R3D_STAT(6, __popc(__ballot_sync(FULL, cnt > 0)));
        if (total == 0u) continue;
        double2 xy = make_double2(0.0, 0.0), zw = make_double2(0.0, 0.0);
        bool have = false;
#if R3D_LEAN_TMA
        // The stream is staged 32 records at a time into one of two shared-memory buffers by bulk async copies:
        // the lane that owns a segment copies the part of it that falls into the chunk's window straight to its
        // place in the buffer (the prefix sums are the compaction), lane 0 announces the chunk's byte count to
        // the buffer's mbarrier, and the next chunk is in flight while the current one is filtered and scanned.
        const uint32_t nchunks = (total + 31u) >> 5;
        auto issue = [&](uint32_t c) {
            const uint32_t c0 = c << 5, b = c & 1u;
            if (lane == 0) mbar_arrive_expect_tx(L.mbar0 + 8u * b, min(32u, total - c0) * 32u);
            const uint32_t lo = max(excl, c0), hi = min(excl + cnt, c0 + 32u);
            if (hi > lo) bulk_copy_g2s(L.stage_addr + (b * 32u + (lo - c0)) * 32u, v.pts + s + (lo - excl), (hi - lo) * 32u, L.mbar0 + 8u * b);
        };
        issue(0u);
        for (uint32_t c = 0; c < nchunks; c++) {
            if (c + 1u < nchunks) issue(c + 1u);      // its buffer was last read two chunks ago, before a __syncwarp
            const uint32_t b = c & 1u;
            mbar_wait(L.mbar0 + 8u * b, (L.phase >> b) & 1u);
            L.phase ^= 1u << b;
            have = (c << 5) + (uint32_t)lane < total;
            if (have) {
                xy = L.stage[(b * 32u + lane) * 2u];
                zw = L.stage[(b * 32u + lane) * 2u + 1u];
            }
            __syncwarp();      // every lane has its record: the buffer may be refilled
            const bool keep = have && xy.x >= L.lo[0] && xy.x <= L.hi[0] && xy.y >= L.lo[1] && xy.y <= L.hi[1] &&
                              zw.x >= L.lo[2] && zw.x <= L.hi[2];
            const unsigned bal = __ballot_sync(FULL, keep);
            if (keep) {
                const int slot = L.n + __popc(bal & lt);
                L.sxy[slot] = xy;
                L.szw[slot] = zw;
            }
            L.n += __popc(bal);
            if (L.n >= LEAN_FLUSH) list_flush(L, best, seed_pos, qx, qy, qz);
        }
#else
        // resolve stream index t = t0 + lane and issue its loads
        auto fetch = [&](uint32_t t0) {
            const uint32_t t = t0 + (uint32_t)lane;
            have = t < total;
            const uint32_t tc = min(t, total - 1u);
            int seg = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const uint32_t ex = __shfl_sync(FULL, excl, seg + step);
                if (ex <= tc) seg += step;
            }
            const uint32_t pos = __shfl_sync(FULL, s, seg) + (tc - __shfl_sync(FULL, excl, seg));
            if (have) load_record(v.pts + pos, xy, zw);
        };
        fetch(0u);
        for (uint32_t t0 = 0; t0 < total; t0 += 32) {
            const bool keep = have && xy.x >= L.lo[0] && xy.x <= L.hi[0] && xy.y >= L.lo[1] && xy.y <= L.hi[1] &&
                              zw.x >= L.lo[2] && zw.x <= L.hi[2];
            const unsigned bal = __ballot_sync(FULL, keep);
            if (keep) {
                const int slot = L.n + __popc(bal & lt);
                L.sxy[slot] = xy;
                L.szw[slot] = zw;
            }
            L.n += __popc(bal);
            if (t0 + 32 < total) fetch(t0 + 32);          // next batch's loads fly while the list is scanned
            if (L.n >= LEAN_FLUSH) list_flush(L, best, seed_pos, qx, qy, qz);
        }
#endif
    }
    if (L.n > 0) list_flush(L, best, seed_pos, qx, qy, qz);
    if (L.nrep > 1) {
        // the four replicas of a query have each seen a quarter of the candidates: give all of them the minimum, so
        // that they keep taking the same decisions
        for (int o = 32 / L.nrep; o < 32; o <<= 1) {
            const double od2 = __shfl_xor_sync(FULL, best.d2, o);
            const int oidx = __shfl_xor_sync(FULL, best.idx, o), opos = __shfl_xor_sync(FULL, best.pos, o);
            if (od2 < best.d2 || (od2 == best.d2 && oidx < best.idx)) { best.d2 = od2; best.idx = oidx; best.pos = opos; }
        }
    }
}

// Exact NN for the warp's 32 queries.
//   seed    a real point of the cloud with its exact squared distance (the previous iteration's
//           neighbour), or d2 = +inf / pos = -1 when the lane has none;
//   other   out: lexicographic minimum over every evaluated point except the seed (d2 = +inf if none).
//           The lane's nearest neighbour is the lexicographic minimum of seed and other;
//   slack   extra radius (>= 0) a seeded lane asks for beyond its bound;
//   cover   out: every point within this distance of the query was evaluated (0 = no claim).  With
//           cover and other, sqrt(other.d2) and cover are both lower bounds of the distance to any
//           point that is not the seed, which is what the caller's skip test needs.
// Inactive lanes take part in the collectives only.  `list`: 2 * LEAN_LIST_CAP double2 of shared memory
// owned by the warp.
//
// Expanding search in rounds.  Every unfinished lane asks for the ball of radius min(its bound + slack,
// its rcap), i.e. a box of cells.  Consecutive queries are usually neighbours in space, but a run of 32
// can straddle two blocks that are far apart (a surface leaves and re-enters a row of blocks), so a
// round serves the lanes around a LEADER (the first unfinished lane): those whose box lies inside a
// window around the leader's.  The union box of the served lanes is walked; a served lane is
// finished once its best distance is within the radius it asked for (nothing unseen can be closer),
// otherwise its rcap doubles.  Lanes left out wait for a later round with a leader near them.  Lanes
// that arrive with a seed ask for their full ball at once, so a coherent warp takes one round.
#ifndef R3D_LEAN_WIN
#define R3D_LEAN_WIN 4
#endif
constexpr int LEAN_WIN = R3D_LEAN_WIN;      // served: box inside the leader's box widened by max(4 cells, its own extent)

// All 32 lanes work for ONE query (lane `owner`'s): exact search of the ball of radius R around it when its box is far
// too large for the union walk (a far outlier, a query far outside the cloud, a badly misaligned pair).  The block
// list is swept 32 blocks at a time: a lane computes the distance from the query to its block's bounding box and the
// blocks that can reach into the ball are scanned by the whole warp, 32 points at a time.  Every point within R of
// the query is evaluated (the seed excepted), so afterwards nothing unseen is closer than R.  Returns, on every lane,
// the lexicographic minimum of (squared distance, cloud index) over the evaluated points other than the seed, and
// whether every block was scanned.
__device__ __forceinline__ Best far_query_search(int owner, const Best& seed, const GridView& v, const GridParams& g, double qx,
                                                 double qy, double qz, float R, bool& all_scanned) {
    const int lane = threadIdx.x & 31;
    const double ax = __shfl_sync(0xffffffffu, qx, owner), ay = __shfl_sync(0xffffffffu, qy, owner), az = __shfl_sync(0xffffffffu, qz, owner);
    const int seed_pos = __shfl_sync(0xffffffffu, seed.pos, owner);
    const float Rw = __shfl_sync(0xffffffffu, R, owner);
    // squared reach, rounded up generously (the block test only has to be conservative)
    const double reach2 = (Rw < 1e18f) ? (double)Rw * (double)Rw * (1.0 + 1e-5) : CUDART_INF;
    const double bw = 4.0 * g.cell, pad = 1e-6 * g.cell;
    Best loc;
    loc.d2 = CUDART_INF; loc.idx = 0x7fffffff; loc.pos = -1;
    bool all = true;
    R3D_STAT(8, 1);
    for (int b0 = 0; b0 < g.n_blocks; b0 += 32) {
        R3D_STAT(9, 1);
        const int b = b0 + lane;
        bool hit = false;
        if (b < g.n_blocks) {
            const uint32_t key = __ldg(v.block_key + b);
            const int cbx = (int)(key % (uint32_t)g.bx);
            const int cby = (int)((key / (uint32_t)g.bx) % (uint32_t)g.by);
            const int cbz = (int)(key / ((uint32_t)g.bx * (uint32_t)g.by));
            const double lx = g.ox + (double)cbx * bw - pad, ly = g.oy + (double)cby * bw - pad, lz = g.oz + (double)cbz * bw - pad;
            const double dx = fmax(fmax(lx - ax, ax - (lx + bw + 2.0 * pad)), 0.0);
            const double dy = fmax(fmax(ly - ay, ay - (ly + bw + 2.0 * pad)), 0.0);
            const double dz = fmax(fmax(lz - az, az - (lz + bw + 2.0 * pad)), 0.0);
            hit = !(dx * dx + dy * dy + dz * dz > reach2);      // (the grid spans the cloud's bounding box: no point lies outside its block)
        }
        unsigned m = __ballot_sync(0xffffffffu, hit);
        const int in_range = min(32, g.n_blocks - b0);
        all = all && (__popc(m) == in_range);
        while (m) {
            // a run of consecutive hit blocks is one contiguous range of the sorted cloud (a block holds ~10 points:
            // scanned one block at a time most lanes would idle)
            const int j = __ffs(m) - 1;
            const unsigned rest = ~(m >> j);
            const int len = rest ? __ffs(rest) - 1 : 32;
            m = (j + len >= 32) ? 0u : (m >> (j + len)) << (j + len);
            const uint32_t s0 = __ldg(v.block_first + b0 + j), e0 = __ldg(v.block_first + b0 + j + len);
            R3D_STAT(10, (e0 - s0 + 31) / 32);
            for (uint32_t k = s0 + (uint32_t)lane; k < e0; k += 32) {
                double2 xy, zw;
                load_record(v.pts + k, xy, zw);
                const double ex = ax - xy.x, ey = ay - xy.y, ez = az - zw.x;
                const double d2 = ex * ex + ey * ey + ez * ez;
                const int idx = __double2loint(zw.y);
                if ((int)k != seed_pos && (d2 < loc.d2 || (d2 == loc.d2 && idx < loc.idx))) { loc.d2 = d2; loc.idx = idx; loc.pos = (int)k; }
            }
        }
    }
    // lexicographic minimum over the lanes
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od2 = __shfl_xor_sync(0xffffffffu, loc.d2, o);
        const int oidx = __shfl_xor_sync(0xffffffffu, loc.idx, o), opos = __shfl_xor_sync(0xffffffffu, loc.pos, o);
        if (od2 < loc.d2 || (od2 == loc.d2 && oidx < loc.idx)) { loc.d2 = od2; loc.idx = oidx; loc.pos = opos; }
    }
    all_scanned = all;
    return loc;
}

// float -> int map that preserves order under SIGNED comparison (and is its own inverse)
__device__ __forceinline__ int f2sord(float f) {
    const int b = __float_as_int(f);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float sord2f(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7fffffff)); }
// warp minimum / maximum of floats: sm_100a reduces fp32 in one instruction (CREDUX.MIN/MAX.F32; NaN inputs are ignored, as by
// fminf / fmaxf) -- the order-preserving integer map above is what older parts needed
__device__ __forceinline__ float wmin_f(float v) {
    float r;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float wmax_f(float v) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
    return r;
}
// Square root for BOUNDS: sqrt.approx.f32 (MUFU.SQRT, denormals handled) is within 2^-23 relative of the exact root, half of
// what the IEEE sequence costs; every use multiplies by 1 +- 1e-6 (= 2^-19.9) towards the safe side, which covers it together
// with the roundings of the conversion before and the product after.
__device__ __forceinline__ float sqrt_bound(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// cell index of a float cell coordinate, clamped into [0, n-1] (NaN -> 0)
__device__ __forceinline__ int cell_lo(float t, int n) { return min(max(__float2int_rd(t), 0), n - 1); }
__device__ __forceinline__ int cell_hi(float t, int n) { return min(max(__float2int_rd(t), 0), n - 1); }

// `sq`: 32 float4 of shared memory owned by the warp (may be reused by the caller between calls; used with PRUNE only).
template <bool PRUNE>
__device__ __forceinline__ void lean_search(double2* list, float4* sq, unsigned& tma_phase, const Best& seed, Best& other, const GridView& v,
                                            const GridParams& g, bool active, double qx, double qy, double qz, float slack,
                                            float& cover, int qpt = 32) {
    // Cell coordinates in float.  The box only has to CONTAIN the ball, so every rounding is covered by a
    // margin: 2e-3 cells absolute plus 2.4e-7 relative to the coordinate (one float ulp), on top of the
    // 1e-6 relative slack of the radius.
    const float tqx = (float)((qx - g.ox) * g.inv_cell), tqy = (float)((qy - g.oy) * g.inv_cell),
                tqz = (float)((qz - g.oz) * g.inv_cell);
    const float mgx = fmaf(fabsf(tqx), 2.4e-7f, 2e-3f), mgy = fmaf(fabsf(tqy), 2.4e-7f, 2e-3f),
                mgz = fmaf(fabsf(tqz), 2.4e-7f, 2e-3f);
    const float inv_cell = (float)g.inv_cell * 1.000001f;
    const float qxf_lo = __double2float_rd(qx), qxf_hi = __double2float_ru(qx);
    const float qyf_lo = __double2float_rd(qy), qyf_hi = __double2float_ru(qy);
    const float qzf_lo = __double2float_rd(qz), qzf_hi = __double2float_ru(qz);
    const int big = 0x7fffffff;
    const float inf = __int_as_float(0x7f800000);
    WarpList L;
    L.sxy = list;
    L.szw = list + LEAN_LIST_CAP;
    L.n = 0;
    // qpt < 32 (8 or 16): lanes l, l + qpt, ... carry the same query, seed and state; they differ in the list entries
    // they evaluate, and are merged after every walk
    L.nrep = 32 / qpt;
    L.rep = (int)(threadIdx.x & 31) / qpt;
    L.sq = sq;
#if R3D_LEAN_TMA
    L.stage = list + 2 * LEAN_LIST_CAP;
    L.stage_addr = smem_u32(L.stage);
    L.mbar0 = smem_u32(list + 2 * LEAN_LIST_CAP + 128);
    L.phase = tma_phase;
#endif
#ifndef R3D_LEAN_RCAP0
#define R3D_LEAN_RCAP0 1.0f
#endif
    float rcap = (seed.d2 < CUDART_INF) ? inf : R3D_LEAN_RCAP0 * (float)g.cell;
    bool fin = !active;
    cover = 0.0f;
    while (true) {
        const unsigned unfin = __ballot_sync(FULL, !fin);
        if (unfin == 0u) {
#if R3D_LEAN_TMA
            tma_phase = L.phase;
#endif
            return;
        }
        R3D_STAT(1, 1);
        const int leader = __ffs(unfin) - 1;
        // float sqrt rounded up
        const float rl = sqrt_bound((float)fmin(seed.d2, other.d2)) * 1.000001f + 1e-30f;
        const float ruse = fminf(rl + slack, rcap);
        const float rc = ruse * inv_cell;
        const int mlox = cell_lo(tqx - rc - mgx, g.nx), mhix = cell_hi(tqx + rc + mgx, g.nx);
        const int mloy = cell_lo(tqy - rc - mgy, g.ny), mhiy = cell_hi(tqy + rc + mgy, g.ny);
        const int mloz = cell_lo(tqz - rc - mgz, g.nz), mhiz = cell_hi(tqz + rc + mgz, g.nz);
        // window around the leader's box (always contains the leader's own box)
        const int Llox = __shfl_sync(FULL, mlox, leader), Lhix = __shfl_sync(FULL, mhix, leader);
        const int Lloy = __shfl_sync(FULL, mloy, leader), Lhiy = __shfl_sync(FULL, mhiy, leader);
        const int Lloz = __shfl_sync(FULL, mloz, leader), Lhiz = __shfl_sync(FULL, mhiz, leader);
        const int wx = max(LEAN_WIN, Lhix - Llox + 1), wy = max(LEAN_WIN, Lhiy - Lloy + 1), wz = max(LEAN_WIN, Lhiz - Lloz + 1);
        const bool part = !fin && mlox >= Llox - wx && mhix <= Lhix + wx && mloy >= Lloy - wy && mhiy <= Lhiy + wy &&
                          mloz >= Lloz - wz && mhiz <= Lhiz + wz;
        const int lox = __reduce_min_sync(FULL, part ? mlox : big), hix = __reduce_max_sync(FULL, part ? mhix : -1);
        const int loy = __reduce_min_sync(FULL, part ? mloy : big), hiy = __reduce_max_sync(FULL, part ? mhiy : -1);
        const int loz = __reduce_min_sync(FULL, part ? mloz : big), hiz = __reduce_max_sync(FULL, part ? mhiz : -1);
        const bool whole = lox == 0 && loy == 0 && loz == 0 && hix == g.nx - 1 && hiy == g.ny - 1 && hiz == g.nz - 1;
        // (counts as float products: exact below 2^24, and anything larger is above both limits below -- an axis has up to
        // 2^20 cells, so the integer products would need 64 bits)
        const float nxbf = (float)((hix >> 2) - (lox >> 2) + 1);
        const float nseg = (float)(hiy - loy + 1) * (float)(hiz - loz + 1) * nxbf;
        const float nblk = (float)((hiy >> 2) - (loy >> 2) + 1) * (float)((hiz >> 2) - (loz >> 2) + 1) * nxbf;
        // small boxes are walked row by row, larger ones block by block.  A box of more than LEAN_BLOCKS_MAX block slots
        // that is also a large part of the cloud (more than an eighth of its occupied blocks: a small cloud with far
        // outliers, a badly misaligned pair) is searched one query at a time by sweeping the cloud's block list with
        // distance pruning; in a large cloud the same number of slots is a small neighbourhood and is still walked
        // (the sweep costs n_blocks / 32 steps per query, the walk nblk / 32 lookups per warp).
        if (nseg <= (float)LEAN_ROW_SEGS_MAX || nblk <= (float)max(LEAN_BLOCKS_MAX, g.n_blocks >> 3)) {
            // the union of the served balls' bounding boxes in world coordinates: a point outside it is
            // outside every served ball.  When the box is the whole grid everything is looked at once and
            // the answer is final for every lane, so nothing may be filtered.
            const float rw = ruse * 1.000001f;
            const float l0 = wmin_f(part ? __fsub_rd(qxf_lo, rw) : inf), h0 = wmax_f(part ? __fadd_ru(qxf_hi, rw) : -inf);
            const float l1 = wmin_f(part ? __fsub_rd(qyf_lo, rw) : inf), h1 = wmax_f(part ? __fadd_ru(qyf_hi, rw) : -inf);
            const float l2 = wmin_f(part ? __fsub_rd(qzf_lo, rw) : inf), h2 = wmax_f(part ? __fadd_ru(qzf_hi, rw) : -inf);
            L.lo[0] = whole ? -CUDART_INF : (double)l0; L.hi[0] = whole ? CUDART_INF : (double)h0;
            L.lo[1] = whole ? -CUDART_INF : (double)l1; L.hi[1] = whole ? CUDART_INF : (double)h1;
            L.lo[2] = whole ? -CUDART_INF : (double)l2; L.hi[2] = whole ? CUDART_INF : (double)h2;
            if constexpr (PRUNE) {
                __syncwarp();
                L.sq[threadIdx.x & 31] = make_float4(tqx, tqy, tqz, (part && !whole) ? rc + (mgx + mgy + mgz) : (part ? 3.0e18f : -1.0f));
                __syncwarp();
            }
            lean_scan_box<PRUNE>(L, other, seed.pos, v, g, lox, loy, loz, hix, hiy, hiz, qx, qy, qz);
            if (part) {
                const float rn = sqrt_bound((float)fmin(seed.d2, other.d2)) * 1.000001f + 1e-30f;
                fin = whole || (rn <= ruse) || !(ruse < 1e37f);
                if (fin) cover = whole ? inf : ruse;
                rcap *= 2.0f;   // a poor bound (a point seen in a far corner of the union) must not set the next radius
            }
        } else {
            // far outliers / queries far outside the cloud / badly misaligned pairs: the served lanes are searched one
            // at a time, the whole warp working for each (far_query_search)
            unsigned todo = __ballot_sync(FULL, part && (int)(threadIdx.x & 31) < qpt);      // one replica of each query
            while (todo) {
                const int owner = __ffs(todo) - 1;
                todo &= todo - 1u;
                bool all_scanned = false;
                const Best found = far_query_search(owner, seed, v, g, qx, qy, qz, ruse, all_scanned);
                if ((int)(threadIdx.x & 31) % qpt == owner) {
                    if (found.d2 < other.d2 || (found.d2 == other.d2 && found.idx < other.idx)) other = found;
                    const float rn = sqrt_bound((float)fmin(seed.d2, other.d2)) * 1.000001f + 1e-30f;
                    fin = all_scanned || (rn <= ruse) || !(ruse < 1e37f);
                    if (fin) cover = all_scanned ? inf : ruse;
                    rcap *= 2.0f;
                }
            }
        }
    }
}


// ---- group search ---------------------------------------------------------------------------------------------
// The union walk above costs about the same whether 32 lanes of a task search or two: rounds, ~70 segment lookups and a
// candidate stream for the union box come first.  Once ICP has nearly converged most lanes keep their neighbour without a
// search (nn_query_kernel's skip test) and the tasks that are left have a handful of searching lanes each.  Those are served
// here: the warp splits into four groups of eight lanes and every group works for ONE searching query at a time -- the
// (y, z) rows of the cells under that query's own ball, cut at block borders, are looked up one per lane (at most two
// rounds of eight), their points are loaded eight at a time, one per lane, evaluated in fp64 and reduced over the group.
// Same contract as lean_search for the lanes it serves (bit `lane` of `todo`; every one of them seeded, i.e. with a finite
// bound): `other` = lexicographic minimum of (squared distance, cloud index) over every point but the seed within the
// ball of radius `cover` = bound + slack.  Returns the lanes it did NOT serve (their ball spans more than GRP_SEGS_MAX row
// segments): the caller sends those through lean_search.
constexpr int GRP_SEGS_MAX = 16;

__device__ __forceinline__ unsigned group_search(unsigned todo, const Best& seed, Best& other, const GridView& v, const GridParams& g,
                                                 double qx, double qy, double qz, float slack, float& cover) {
    const int lane = threadIdx.x & 31;
    const int grp = lane >> 3, t = lane & 7;
    // the lane's own ball as a box of cells (the arithmetic and margins of lean_search)
    const float tqx = (float)((qx - g.ox) * g.inv_cell), tqy = (float)((qy - g.oy) * g.inv_cell),
                tqz = (float)((qz - g.oz) * g.inv_cell);
    const float mgx = fmaf(fabsf(tqx), 2.4e-7f, 2e-3f), mgy = fmaf(fabsf(tqy), 2.4e-7f, 2e-3f),
                mgz = fmaf(fabsf(tqz), 2.4e-7f, 2e-3f);
    const float rl = sqrt_bound((float)seed.d2) * 1.000001f + 1e-30f;
    const float ruse = rl + slack;
    const float rc = ruse * ((float)g.inv_cell * 1.000001f);
    const int mlox = cell_lo(tqx - rc - mgx, g.nx), mhix = cell_hi(tqx + rc + mgx, g.nx);
    const int mloy = cell_lo(tqy - rc - mgy, g.ny), mhiy = cell_hi(tqy + rc + mgy, g.ny);
    const int mloz = cell_lo(tqz - rc - mgz, g.nz), mhiz = cell_hi(tqz + rc + mgz, g.nz);
    const int mseg = (mhiy - mloy + 1) * (mhiz - mloz + 1) * ((mhix >> 2) - (mlox >> 2) + 1);
    const bool mine = ((todo >> lane) & 1u) != 0u;
    unsigned rem = __ballot_sync(FULL, mine && ruse < 1e30f && mseg <= GRP_SEGS_MAX);
    const unsigned unserved = todo & ~rem;
    R3D_STAT(16, rem ? 1 : 0);
    R3D_STAT(18, __popc(unserved));
    while (rem) {
        R3D_STAT(17, 1);
        // the next four searching lanes, one per group
        const int o0 = __ffs(rem) - 1;
        rem &= rem - 1u;
        const int o1 = __ffs(rem) - 1;      // (-1 once rem is exhausted: rem & (rem - 1) stays 0)
        rem &= rem - 1u;
        const int o2 = __ffs(rem) - 1;
        rem &= rem - 1u;
        const int o3 = __ffs(rem) - 1;
        rem &= rem - 1u;
        const int owner = grp == 0 ? o0 : grp == 1 ? o1 : grp == 2 ? o2 : o3;
        const int src = max(owner, 0);
        const double ax = __shfl_sync(FULL, qx, src), ay = __shfl_sync(FULL, qy, src), az = __shfl_sync(FULL, qz, src);
        const int lox = __shfl_sync(FULL, mlox, src), hix = __shfl_sync(FULL, mhix, src);
        const int loy = __shfl_sync(FULL, mloy, src), hiy = __shfl_sync(FULL, mhiy, src);
        const int loz = __shfl_sync(FULL, mloz, src), hiz = __shfl_sync(FULL, mhiz, src);
        const int spos = __shfl_sync(FULL, seed.pos, src);
        const int ny = hiy - loy + 1, bx0 = lox >> 2, nxb = (hix >> 2) - bx0 + 1;
        const int nseg = owner >= 0 ? ny * (hiz - loz + 1) * nxb : 0;
        const float inv_nxb = 1.0f / (float)nxb, inv_ny = 1.0f / (float)ny;
        Best loc;
        loc.d2 = CUDART_INF; loc.idx = 0x7fffffff; loc.pos = -1;
#pragma unroll 1
        for (int base_seg = 0; __any_sync(FULL, base_seg < nseg); base_seg += 8) {
            const int sidx = base_seg + t;
            uint32_t s = 0, cnt = 0;
            if (sidx < nseg) {
                // sidx = (k2 * ny + k1) * nxb + k0; quotients of small integers by reciprocal, then fixed up
                int row = (int)((float)sidx * inv_nxb);
                row += (sidx - row * nxb >= nxb) - (sidx - row * nxb < 0);
                const int k0 = sidx - row * nxb;
                int k2 = (int)((float)row * inv_ny);
                k2 += (row - k2 * ny >= ny) - (row - k2 * ny < 0);
                const int k1 = row - k2 * ny;
                const int cbx = bx0 + k0, iy = loy + k1, iz = loz + k2;
                const int cby = iy >> 2, cbz = iz >> 2;
                const uint32_t key = (uint32_t)((cbz * g.by + cby) * g.bx + cbx);
                const int b = find_block(v.hash, g.hash_mask, key, block_hash(cbx, cby, cbz));
                if (b >= 0) {
                    const int x0 = max(lox, cbx * 4) - cbx * 4, x1 = min(hix, cbx * 4 + 3) - cbx * 4;
                    const int f = ((iz & 3) << 4) | ((iy & 3) << 2);
                    cell_range(v, b, f + x0, f + x1 + 1, s, cnt);
                }
            }
            // the group's candidate stream = concatenation of its eight segments
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL, incl, o, 8);
                if (t >= o) incl += y;
            }
            const uint32_t excl = incl - cnt;
            const uint32_t total = __shfl_sync(FULL, incl, 7, 8);
#pragma unroll 1
            for (uint32_t base = 0; __any_sync(FULL, base < total); base += 8) {
                R3D_STAT(19, 1);
                const uint32_t tc = base + (uint32_t)t;
                const bool have = tc < total;
                const uint32_t tcc = min(tc, total - 1u);      // (total = 0: have is false and the position is not used)
                int seg = 0;
#pragma unroll
                for (int step = 4; step >= 1; step >>= 1) {
                    const uint32_t ex = __shfl_sync(FULL, excl, seg + step, 8);
                    if (ex <= tcc) seg += step;
                }
                const uint32_t pos = __shfl_sync(FULL, s, seg, 8) + (tcc - __shfl_sync(FULL, excl, seg, 8));
                if (have) {
                    double2 xy, zw;
                    load_record(v.pts + pos, xy, zw);
                    const double dx = ax - xy.x, dy = ay - xy.y, dz = az - zw.x;
                    const double d2 = dx * dx + dy * dy + dz * dz;
                    const int idx = __double2loint(zw.y);
                    if ((int)pos != spos && (d2 < loc.d2 || (d2 == loc.d2 && idx < loc.idx))) { loc.d2 = d2; loc.idx = idx; loc.pos = (int)pos; }
                }
            }
        }
        // lexicographic minimum over the group, then to the owner
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const double od2 = __shfl_xor_sync(FULL, loc.d2, o);
            const int oidx = __shfl_xor_sync(FULL, loc.idx, o), opos = __shfl_xor_sync(FULL, loc.pos, o);
            if (od2 < loc.d2 || (od2 == loc.d2 && oidx < loc.idx)) { loc.d2 = od2; loc.idx = oidx; loc.pos = opos; }
        }
        const int gi = lane == o0 ? 0 : lane == o1 ? 1 : lane == o2 ? 2 : lane == o3 ? 3 : -1;
        const int from = gi >= 0 ? gi * 8 : lane;
        const double rd2 = __shfl_sync(FULL, loc.d2, from);
        const int ridx = __shfl_sync(FULL, loc.idx, from), rpos = __shfl_sync(FULL, loc.pos, from);
        if (gi >= 0) {
            if (rd2 < other.d2 || (rd2 == other.d2 && ridx < other.idx)) { other.d2 = rd2; other.idx = ridx; other.pos = rpos; }
            cover = ruse;
        }
    }
    return unserved;
}

}  // namespace r3d
