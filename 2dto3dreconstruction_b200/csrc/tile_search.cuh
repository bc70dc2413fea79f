// Tile search: exact nearest neighbours on the sparse-block grid, candidates staged into shared memory by
// TMA bulk copies, every lane walking only the cells of its OWN ball.
//
// Contract: the same as lean_search.cuh -- exact NN in fp64 with the lowest cloud index on exact ties, equal
// to float64 brute force (utils/icp.py:58-71 nearest_neighbor up to exact ties) -- and the same interface,
// so the two can be swapped under the query kernel (R3D_OPT_NN_IMPL) and must return identical bits.
//
// What changed against the warp-union walk of round 1 (which evaluated ~80 candidates per query where the
// query's own ball holds ~10-25, and was instruction-issue bound doing so):
//   1. staging unit = BLOCK.  The warp takes the union box of its lanes' balls as before, but looks up the
//      BLOCKS under it (<= 32 slots per batch, one per lane: one hash probe + two cell-start loads) and copies,
//      per occupied block, ONE contiguous range of the block-relative fp32 records (grid.cuh `packed`, 16 B
//      per point: the range from the first to the last cell of the block that the union box touches) plus
//      the block's 64-entry cell table, with two cp.async.bulk copies counted on one mbarrier per warp.  A
//      copy is ~0.3-1.5 KB instead of the 160 B row segments that made TMA lose in round 1;
//   2. walk = per lane, in two phases.  (A) every lane visits the staged blocks (a warp-uniform loop), intersects
//      its own cell box with the block, drops the z-slabs farther than its radius and pushes, per remaining
//      slab, ONE range of staged records (from the first to the last cell of its box in that slab: cells are
//      stored z, y, x-major) onto a small per-lane queue in shared memory; (B) every lane runs over its queued
//      ranges: an LDS.128, three FADD, FMUL + two FFMA, two min/max and a compare-select per candidate.
//      Splitting the walk keeps the candidate loop out of the block/slab loops: nested, the lanes of a warp
//      sat in different blocks and rows at any time and the candidate loop ran at 7 % lane utilisation
//      (profiles/r02_nn_tile_steps.md);
//   3. filter = fp32, decision = fp64.  With block-relative coordinates the fp32 distance is off by at most
//      `marg` (a few 1e-7 of the block size).  The walk keeps the fp32 minimum and runner-up; afterwards the
//      nearest candidate -- if its fp32 distance is within the lane's exact bound + marg -- is re-evaluated
//      in fp64 from the original record, by all lanes at once, and only that value is compared and stored.
//      When the runner-up also falls inside the error band (exact ties, duplicates, lattices) the queue is
//      walked again with an fp64 re-evaluation of every candidate in the band.  Everything else is provably
//      farther than the current best, so the result is the fp64 brute-force answer bit for bit.  The fp32
//      minimum over the non-seed candidates, rounded inward, is the lower bound the caller's skip budget
//      needs (it never needed more than a bound).
//
// Status (round 2): exact on every test, 14 candidate evaluations per query instead of 80, TMA-staged -- and
// 1.3-2x SLOWER than the warp-union walk on the bench workload, so it is the opt-in search (R3D_OPT_NN_IMPL = 2).
// The per-round machinery (union box, block lookups, staging, range building: ~1500 warp instructions) costs
// what the cheaper evaluations save, and at 96 registers + 8.6 KB of staging per warp only 20 warps are
// resident per SM.  Measurements and the variants that were tried: profiles/r02_nn_tile_search.md.
#pragma once
#include "lean_search.cuh"

namespace r3d {

#ifndef R3D_TILE_CAP
#define R3D_TILE_CAP 384
#endif
#ifndef R3D_TILE_TBLS
#define R3D_TILE_TBLS 8
#endif
constexpr int TILE_CAP = R3D_TILE_CAP;        // staged records per warp and pass (16 B each); < 512 (queue entry format)
constexpr int TILE_TBLS = R3D_TILE_TBLS;      // staged blocks per warp and pass; <= 16 (queue entry format)
constexpr int TILE_TBL_WORDS = 68;            // 64 cell starts + the end of the last cell, padded to a multiple of 16 bytes
constexpr int TILE_QCAP = 16;                 // queued ranges per lane (the queue is flushed when full)
constexpr int TILE_QUEUE_WORDS = TILE_QCAP * 32;
static_assert(TILE_CAP < 512 && TILE_TBLS <= 16, "queue entries hold 9 + 9 + 4 bits");

struct __align__(16) TileSlot {               // one staged block
    uint32_t win_lo, win_hi;                  // sorted positions [win_lo, win_hi) are in shared memory ...
    int32_t rel;                              // ... record of sorted position k at pts[k + rel]
    int32_t pad;
    int32_t cx4, cy4, cz4;                    // first cell of the block along each axis (4 x block coordinate)
    int32_t pad2;
};

struct __align__(16) TileWarp {               // shared memory of one warp
    float4 pts[TILE_CAP];                     // staged records (coordinates relative to their own block's origin)
    uint32_t tbl[TILE_TBLS][TILE_TBL_WORDS];  // staged cell tables
    TileSlot slot[TILE_TBLS];
    float4 boff[TILE_TBLS];                   // block origin relative to the round's frame (xyz), w = TileSlot::rel
    unsigned long long mbar;
    unsigned long long pad;
};

// Every point whose exact squared distance is <= best_d2 has an fp32 squared distance <= this value, when the fp32
// distance is off by at most marg (see tile_search).
__device__ __forceinline__ float tile_threshold(double best_d2, float marg) {
    const float r = sqrtf(__double2float_ru(best_d2));      // +inf stays +inf
    const float rr = fmaf(r, 1.000001f, marg);
    return rr * rr * 1.000001f;
}

// fp64 evaluation of the sorted point k from its original record; the expression is eval_staged()'s
__device__ __forceinline__ bool tile_exact(Best& other, const Point4* __restrict__ pts, int k, double qx, double qy, double qz) {
    double2 xy, zw;
    load_record(pts + k, xy, zw);
    const double dx = qx - xy.x, dy = qy - xy.y, dz = qz - zw.x;
    const double d2 = dx * dx + dy * dy + dz * dz;
    const int idx = __double2loint(zw.y);
    if (d2 < other.d2 || (d2 == other.d2 && idx < other.idx)) {
        other.d2 = d2;
        other.idx = idx;
        other.pos = k;
        return true;
    }
    return false;
}

// Diagnostic counters of this search (R3D_NN_STATS builds): 2 = candidate evaluations summed over the LANES
// (divide by 32 for the per-lane figure the lean search reports), 5 = block slots looked up, 6 = blocks staged,
// 4 = staging passes, 14 = fp64 re-evaluations (lanes), 15 = records staged.

// Arguments as lean_search; `lb_other` (out) = lower bound of the distance from the query to every evaluated
// point other than the seed (+inf if there was none); `queue`: TILE_QUEUE_WORDS words of shared memory owned by
// the warp (may be reused by the caller between calls).
__device__ __forceinline__ void tile_search(TileWarp& S, uint32_t* queue, unsigned& phase, const Best& seed, Best& other, float& lb_other,
                                            const GridView& v, const GridParams& g, bool active, double qx, double qy, double qz,
                                            float slack, float& cover, int qpt = 32) {
    const float inf = __int_as_float(0x7f800000);
    cover = 0.0f;
    lb_other = inf;
    if (__ballot_sync(FULL, active) == 0u) return;      // (every lane of the task skips its search: the common case late in ICP)
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    // cell coordinates in float, margins as in lean_search
    const float tqx = (float)((qx - g.ox) * g.inv_cell), tqy = (float)((qy - g.oy) * g.inv_cell),
                tqz = (float)((qz - g.oz) * g.inv_cell);
    const float mgx = fmaf(fabsf(tqx), 2.4e-7f, 2e-3f), mgy = fmaf(fabsf(tqy), 2.4e-7f, 2e-3f),
                mgz = fmaf(fabsf(tqz), 2.4e-7f, 2e-3f);
    const float inv_cell = (float)g.inv_cell * 1.000001f;
    const float cellf = (float)g.cell;
    // fp64 rounding of the block origins themselves (matters only when coordinates are ~1e8 times the cell size)
    const float eps_o = __double2float_ru(1e-15 * (fabs(g.ox) + fabs(g.oy) + fabs(g.oz) + (double)(g.nx + g.ny + g.nz) * g.cell));
    const int big = 0x7fffffff;
    const int nrep = 32 / qpt, rep = lane / qpt;
    const uint32_t mbar = smem_u32(&S.mbar), pts_addr = smem_u32(S.pts), tbl_addr = smem_u32(S.tbl);
    float rcap = (seed.d2 < CUDART_INF) ? inf : R3D_LEAN_RCAP0 * (float)g.cell;
    bool fin = !active;
    float m2 = inf;            // fp32 minimum of the squared distances of the evaluated non-seed points
    float margmax = 0.0f;      // largest error bound of the rounds this lane took part in
    while (true) {
        const unsigned unfin = __ballot_sync(FULL, !fin);
        if (unfin == 0u) break;
        R3D_STAT(1, 1);
        const int leader = __ffs(unfin) - 1;
        const float rl = sqrtf((float)fmin(seed.d2, other.d2)) * 1.000001f + 1e-30f;      // float sqrt rounded up
        const float ruse = fminf(rl + slack, rcap);
        const float rc = ruse * inv_cell;
        const int mlox = cell_lo(tqx - rc - mgx, g.nx), mhix = cell_hi(tqx + rc + mgx, g.nx);
        const int mloy = cell_lo(tqy - rc - mgy, g.ny), mhiy = cell_hi(tqy + rc + mgy, g.ny);
        const int mloz = cell_lo(tqz - rc - mgz, g.nz), mhiz = cell_hi(tqz + rc + mgz, g.nz);
        // lanes around the leader are served together (lean_search explains the window)
        const int Llox = __shfl_sync(FULL, mlox, leader), Lhix = __shfl_sync(FULL, mhix, leader);
        const int Lloy = __shfl_sync(FULL, mloy, leader), Lhiy = __shfl_sync(FULL, mhiy, leader);
        const int Lloz = __shfl_sync(FULL, mloz, leader), Lhiz = __shfl_sync(FULL, mhiz, leader);
        const int wx = max(LEAN_WIN, Lhix - Llox + 1), wy = max(LEAN_WIN, Lhiy - Lloy + 1), wz = max(LEAN_WIN, Lhiz - Lloz + 1);
        const bool part = !fin && mlox >= Llox - wx && mhix <= Lhix + wx && mloy >= Lloy - wy && mhiy <= Lhiy + wy &&
                          mloz >= Lloz - wz && mhiz <= Lhiz + wz;
        const int lox = __reduce_min_sync(FULL, part ? mlox : big), hix = __reduce_max_sync(FULL, part ? mhix : -1);
        const int loy = __reduce_min_sync(FULL, part ? mloy : big), hiy = __reduce_max_sync(FULL, part ? mhiy : -1);
        const int loz = __reduce_min_sync(FULL, part ? mloz : big), hiz = __reduce_max_sync(FULL, part ? mhiz : -1);
        const int ubx0 = lox >> 2, nbx = (hix >> 2) - ubx0 + 1, uby0 = loy >> 2, nby = (hiy >> 2) - uby0 + 1, ubz0 = loz >> 2,
                  nbz = (hiz >> 2) - ubz0 + 1;
        const long long nblk = (long long)nbx * nby * nbz;
        if (nblk <= max((long long)LEAN_BLOCKS_MAX, (long long)(g.n_blocks >> 3))) {
            // a lane whose own box is the whole grid looks at every point: nothing may be pruned, and it is finished
            const bool own_whole = mlox == 0 && mloy == 0 && mloz == 0 && mhix == g.nx - 1 && mhiy == g.ny - 1 && mhiz == g.nz - 1;
            const float rc2 = own_whole ? inf : rc * rc * 1.00001f;      // slabs farther than this (cells^2) cannot reach into the ball
            // This round's frame: the origin of the first block of the union box.  The query is taken relative to it
            // once, in fp64; a staged record is relative to its own block's origin, and the block's offset in the frame
            // (boff, a small multiple of the cell size) is subtracted from the query in fp32 when a range of that
            // block is walked.  Error of an fp32 distance formed this way, per component of the difference:
            //   2^-24 |qu| (rounding of qu) + 2^-23 |boff| (cell size and product rounded) + 2^-24 |qu - boff| (the
            //   subtraction) + 2^-24 4.01 cell (the record) + 2^-24 |difference|  <=  2^-24 (3 |qu| + 4 |boff| + 8.02 cell);
            // the norm moves by at most sqrt(3) times that; the fp32 products and sums of the norm add 2^-22 relative
            // (inside tile_threshold's 1e-6).  4e-7 (sum |qu| + 2 max|boff| + 4 cell) is > 1.29 times the bound.
            const float qux = (float)(qx - block_origin(g.ox, 4 * ubx0, g.cell));
            const float quy = (float)(qy - block_origin(g.oy, 4 * uby0, g.cell));
            const float quz = (float)(qz - block_origin(g.oz, 4 * ubz0, g.cell));
            const float boff_max = 4.0f * cellf * (float)(max(nbx, max(nby, nbz)) - 1);
            const float marg = fmaf(4e-7f, fabsf(qux) + fabsf(quy) + fabsf(quz) + 2.0f * boff_max + 4.0f * cellf, eps_o);
            const int nslots = (int)nblk;
            const float inv_nbx = __fdividef(1.0f, (float)nbx), inv_nby = __fdividef(1.0f, (float)nby);      // (the quotients below are fixed up)
            int dealt = 0;      // ranges met so far (replicas of a query share them out)
#ifdef R3D_NN_STATS
            unsigned n_eval = 0, n_exact = 0;
#endif
            R3D_STAT(5, nslots);
            for (int base = 0; base < nslots; base += 32) {
                // ---- one block slot per lane
                const int sidx = base + lane;
                uint32_t cur = 0, end = 0, blk_id = 0;
                int cx4 = 0, cy4 = 0, cz4 = 0;
                if (sidx < nslots) {
                    // sidx = (k2 * nby + k1) * nbx + k0: quotient by reciprocal, then fix up (exact for these small integers)
                    int row = (int)((float)sidx * inv_nbx);
                    row += (sidx - row * nbx >= nbx) - (sidx - row * nbx < 0);
                    const int k0 = sidx - row * nbx;
                    int k2 = (int)((float)row * inv_nby);
                    k2 += (row - k2 * nby >= nby) - (row - k2 * nby < 0);
                    const int k1 = row - k2 * nby;
                    const int cbx = ubx0 + k0, cby = uby0 + k1, cbz = ubz0 + k2;
                    const uint32_t key = (uint32_t)((cbz * g.by + cby) * g.bx + cbx);
                    const int b = find_block(v.hash, g.hash_mask, key, block_hash(cbx, cby, cbz));
                    if (b >= 0) {
                        cx4 = cbx * 4; cy4 = cby * 4; cz4 = cbz * 4;
                        // cells of the block inside the union box lie between its first and last such cell (z, y, x order)
                        const int x0 = max(lox, cx4) - cx4, x1 = min(hix, cx4 + 3) - cx4;
                        const int y0 = max(loy, cy4) - cy4, y1 = min(hiy, cy4 + 3) - cy4;
                        const int z0 = max(loz, cz4) - cz4, z1 = min(hiz, cz4 + 3) - cz4;
                        const uint32_t* cs = v.cell_start + (size_t)b * 64;
                        cur = __ldg(cs + ((z0 << 4) | (y0 << 2) | x0));
                        end = __ldg(cs + ((z1 << 4) | (y1 << 2) | x1) + 1);
                        blk_id = (uint32_t)b;
                    }
                }
                // ---- stage and walk; several passes when the slots do not fit at once
                while (true) {
                    const uint32_t c = min(end - cur, (uint32_t)TILE_CAP);
                    const unsigned hm = __ballot_sync(FULL, c > 0u);
                    if (hm == 0u) break;
                    uint32_t incl = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t y = __shfl_up_sync(FULL, incl, o);
                        if (lane >= o) incl += y;
                    }
                    const int rank = __popc(hm & lt);
                    const bool fits = c > 0u && incl <= (uint32_t)TILE_CAP && rank < TILE_TBLS;      // a prefix of the occupied slots
                    const unsigned fm = __ballot_sync(FULL, fits);
                    const int nfit = __popc(fm);
                    const uint32_t totpts = __shfl_sync(FULL, incl, 31 - __clz(fm));
                    R3D_STAT(4, 1);
                    R3D_STAT(6, nfit);
                    R3D_STAT(15, totpts);
                    if (lane == 0) mbar_arrive_expect_tx(mbar, totpts * 16u + (uint32_t)nfit * (TILE_TBL_WORDS * 4u));
                    __syncwarp();
                    if (fits) {
                        const uint32_t excl = incl - c;
                        bulk_copy_g2s(pts_addr + excl * 16u, v.packed + cur, c * 16u, mbar);
                        bulk_copy_g2s(tbl_addr + (uint32_t)rank * (TILE_TBL_WORDS * 4u), v.cell_start + (size_t)blk_id * 64, TILE_TBL_WORDS * 4u, mbar);
                        TileSlot sl;
                        sl.win_lo = cur; sl.win_hi = cur + c; sl.rel = (int32_t)excl - (int32_t)cur; sl.pad = 0;
                        sl.cx4 = cx4; sl.cy4 = cy4; sl.cz4 = cz4; sl.pad2 = 0;
                        S.slot[rank] = sl;
                        S.boff[rank] = make_float4((float)(cx4 - 4 * ubx0) * cellf, (float)(cy4 - 4 * uby0) * cellf,
                                                   (float)(cz4 - 4 * ubz0) * cellf, __int_as_float(sl.rel));
                    }
                    __syncwarp();
                    mbar_wait(mbar, phase & 1u);
                    phase ^= 1u;
                    if (part) {
                        float m1 = inf, c2 = inf;      // nearest and second-nearest fp32 squared distance among the queued candidates
                        int k1 = -1;                   // sorted position of the nearest
                        int nq = 0;
                        // Phase B: the lane's queued ranges.  exact = false: fp32 only, tracking (m1, k1, c2);
                        // exact = true: every candidate within thr of the query, except k1, is re-evaluated in fp64.
                        auto walk = [&](const bool exact, float thr) {
                            for (int qi = 0; qi < nq; qi++) {
                                const uint32_t u = queue[qi * 32 + lane];
                                uint32_t k = u & 0x1ffu;
                                const uint32_t e = k + ((u >> 9) & 0x1ffu);
                                const float4 bo = S.boff[u >> 18];
                                const float qrx = qux - bo.x, qry = quy - bo.y, qrz = quz - bo.z;
                                const int rel = __float_as_int(bo.w);
                                const uint32_t seed_k = (uint32_t)(seed.pos + rel);      // where the seed sits if this range holds it
                                do {
                                    const float4 p = S.pts[k];
                                    const float dx = qrx - p.x, dy = qry - p.y, dz = qrz - p.z;
                                    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                                    const float d2e = (k == seed_k) ? inf : d2;
                                    if (!exact) {
                                        c2 = fminf(c2, fmaxf(m1, d2e));
                                        if (d2e < m1) { m1 = d2e; k1 = (int)k - rel; }
#ifdef R3D_NN_STATS
                                        n_eval++;
#endif
                                    } else if (d2e <= thr && (int)k - rel != k1) {
#ifdef R3D_NN_STATS
                                        n_exact++;
#endif
                                        if (tile_exact(other, v.pts, (int)k - rel, qx, qy, qz)) thr = tile_threshold(fmin(seed.d2, other.d2), marg);
                                    }
                                } while (++k < e);
                            }
                        };
                        // The fp32 walk remembers the nearest candidate (m1 at k1) and the runner-up distance (c2).  k1 alone
                        // is then re-evaluated in fp64 -- by all lanes at once: a load inside the walk would stall the whole
                        // warp once per lane that needs one -- and unless the runner-up also lies within the error band of
                        // the new bound, nothing else that was queued can be the neighbour.  Otherwise (exact or near ties:
                        // duplicate points, lattices) the queue is walked again, re-evaluating every candidate in the band.
                        auto flush = [&]() {
                            if (nq > 0) {
                                walk(false, 0.0f);
                                m2 = fminf(m2, m1);
                                if (k1 >= 0 && m1 <= tile_threshold(fmin(seed.d2, other.d2), marg)) {
#ifdef R3D_NN_STATS
                                    n_exact++;
#endif
                                    tile_exact(other, v.pts, k1, qx, qy, qz);
                                    const float thr = tile_threshold(fmin(seed.d2, other.d2), marg);
                                    if (c2 <= thr) walk(true, thr);
                                }
                            }
                            m1 = inf; c2 = inf; k1 = -1; nq = 0;
                        };
                        // Phase A: one range per staged block and z-slab of the lane's box
                        for (int r = 0; r < nfit; r++) {
                            const TileSlot sl = S.slot[r];
                            const int x0 = max(mlox, sl.cx4) - sl.cx4, x1 = min(mhix, sl.cx4 + 3) - sl.cx4;
                            const int y0 = max(mloy, sl.cy4) - sl.cy4, y1 = min(mhiy, sl.cy4 + 3) - sl.cy4;
                            const int z0 = max(mloz, sl.cz4) - sl.cz4, z1 = min(mhiz, sl.cz4 + 3) - sl.cz4;
                            if (x0 > x1 || y0 > y1 || z0 > z1) continue;
                            // gaps (cells) between the query and the part of the block that is looked at
                            const float ux = tqx - (float)sl.cx4, uy = tqy - (float)sl.cy4, uz = tqz - (float)sl.cz4;
                            const float gx = fmaxf(fmaxf(fmaxf((float)x0 - ux, ux - (float)(x1 + 1)), 0.0f) - mgx, 0.0f);
                            const float gy = fmaxf(fmaxf(fmaxf((float)y0 - uy, uy - (float)(y1 + 1)), 0.0f) - mgy, 0.0f);
                            const float gxy = fmaf(gy, gy, gx * gx);
                            if (gxy > rc2) continue;
                            const uint32_t* tb = S.tbl[r];
                            for (int iz = z0; iz <= z1; iz++) {
                                const float gz = fmaxf(fmaxf(fmaxf((float)iz - uz, uz - (float)(iz + 1)), 0.0f) - mgz, 0.0f);
                                if (fmaf(gz, gz, gxy) > rc2) continue;
                                const uint32_t s = max(tb[(iz << 4) | (y0 << 2) | x0], sl.win_lo);
                                const uint32_t e = min(tb[((iz << 4) | (y1 << 2) | x1) + 1], sl.win_hi);
                                if (s >= e) continue;
                                if (nrep > 1) {      // the replicas of a query share its ranges out
                                    const bool mine = (dealt & (nrep - 1)) == rep;
                                    dealt++;
                                    if (!mine) continue;
                                }
                                queue[nq * 32 + lane] = (uint32_t)((int)s + sl.rel) | ((e - s) << 9) | ((uint32_t)r << 18);
                                if (++nq == TILE_QCAP) flush();
                            }
                        }
                        flush();
                    }
                    __syncwarp();      // every lane is done with the staged data: the next pass may overwrite it
                    if (fits) cur += c;
                }
            }
#ifdef R3D_NN_STATS
            R3D_STAT(2, __reduce_add_sync(FULL, n_eval));
            R3D_STAT(14, __reduce_add_sync(FULL, n_exact));
#endif
            if (nrep > 1) {
                // the replicas of a query have each walked a share of its ranges: give all of them the minimum, so that they
                // keep taking the same decisions
                for (int o = qpt; o < 32; o <<= 1) {
                    const double od2 = __shfl_xor_sync(FULL, other.d2, o);
                    const int oidx = __shfl_xor_sync(FULL, other.idx, o), opos = __shfl_xor_sync(FULL, other.pos, o);
                    if (od2 < other.d2 || (od2 == other.d2 && oidx < other.idx)) { other.d2 = od2; other.idx = oidx; other.pos = opos; }
                    m2 = fminf(m2, __shfl_xor_sync(FULL, m2, o));
                }
            }
            if (part) {
                margmax = fmaxf(margmax, marg);
                const float rn = sqrtf((float)fmin(seed.d2, other.d2)) * 1.000001f + 1e-30f;
                fin = own_whole || (rn <= ruse) || !(ruse < 1e37f);
                if (fin) cover = own_whole ? inf : ruse;
                rcap *= 2.0f;   // a poor bound (a point seen in a far corner of the box) must not set the next radius
            }
        } else {
            // far outliers / queries far outside the cloud / badly misaligned pairs: the served lanes are searched one
            // at a time, the whole warp working for each (far_query_search, exact in fp64)
            unsigned todo = __ballot_sync(FULL, part && lane < qpt);      // one replica of each query
            while (todo) {
                const int owner = __ffs(todo) - 1;
                todo &= todo - 1u;
                bool all_scanned = false;
                const Best found = far_query_search(owner, seed, v, g, qx, qy, qz, ruse, all_scanned);
                if (lane % qpt == owner) {
                    if (found.d2 < other.d2 || (found.d2 == other.d2 && found.idx < other.idx)) other = found;
                    m2 = fminf(m2, __double2float_rd(found.d2));
                    const float rn = sqrtf((float)fmin(seed.d2, other.d2)) * 1.000001f + 1e-30f;
                    fin = all_scanned || (rn <= ruse) || !(ruse < 1e37f);
                    if (fin) cover = all_scanned ? inf : ruse;
                    rcap *= 2.0f;
                }
            }
        }
    }
    lb_other = fmaxf(sqrtf(m2) * 0.999999f - margmax, 0.0f);
}

}  // namespace r3d
