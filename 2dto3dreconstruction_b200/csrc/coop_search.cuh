// Group-cooperative exact nearest-neighbour search on the sparse-block grid.
//
// A per-lane search walks nested, data-dependent loops, so the 32 lanes of a warp run at ~27% SIMT
// efficiency (profiles/r01_nn_lane_search.md).  Here queries are spatial neighbours in memory (the
// source cloud is sorted once along a Hilbert curve; rigid updates keep neighbours together) and
// G consecutive lanes (G = 8, 16 or 32) form a group that works as a unit:
//   1. every lane turns its upper bound into a ball; the group takes the union box of its balls;
//   2. the group's lanes walk the box row by row ((y,z) rows of cells; the x-run of a row inside
//      a block is one contiguous range of the sorted cloud) and copy the points that fall inside
//      the union box into a shared-memory list owned by the group;
//   3. the group's lanes scan that list (one broadcast LDS per point per group), each keeping the
//      lexicographic minimum of (squared distance, cloud index) in fp64.
// All loop bounds are taken warp-wide (max over the warp's groups) and group-specific work is
// predicated, so the warp never diverges.  A group's list is a superset of what each of its lanes
// needs, hence the result is still the exact nearest neighbour with lowest-index tie-break, equal
// to float64 brute force (oracle nearest_neighbor_brute; utils/icp.py:58-71 up to exact ties).
#pragma once
#include "grid.cuh"

namespace r3d {

constexpr int COOP_WARP_CAP = 192;       // staged points per warp (32 B each), split among its groups

template <int G>
struct GroupStage {
    static constexpr int NG = 32 / G;                    // groups per warp
    static constexpr int CAP = COOP_WARP_CAP / NG;       // list capacity per group
    static constexpr int ROUND = CAP / G;                // points one lane may add per staging round
    Point4* buf;     // shared memory, CAP entries, owned by this lane's group
    int n;           // points in the group's list (same value in all lanes of the group)
    // union over the group's lanes of [q - r, q + r] per axis (conservative floats): a point outside
    // it is outside every lane's ball and need not be staged
    float lo[3], hi[3];
};

// order-preserving float <-> uint so that min/max can use the integer REDUX unit
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

template <int G>
__device__ __forceinline__ unsigned group_mask() {
    const int lane = threadIdx.x & 31;
    return (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
}
template <int G> __device__ __forceinline__ int gmin_i(int v) { return __reduce_min_sync(group_mask<G>(), v); }
template <int G> __device__ __forceinline__ int gmax_i(int v) { return __reduce_max_sync(group_mask<G>(), v); }
template <int G> __device__ __forceinline__ float gmin_f(float v) { return ord2f(__reduce_min_sync(group_mask<G>(), f2ord(v))); }
template <int G> __device__ __forceinline__ float gmax_f(float v) { return ord2f(__reduce_max_sync(group_mask<G>(), f2ord(v))); }

// every lane scans its group's list; trip count = longest list in the warp, shorter ones predicated
template <int G>
__device__ __forceinline__ void coop_flush(GroupStage<G>& st, Best& best, double qx, double qy, double qz) {
    __syncwarp();
    const int nmax = __reduce_max_sync(0xffffffffu, st.n);
#pragma unroll 2
    for (int c = 0; c < nmax; c++) {
        if (c < st.n) {
            const double2* p = reinterpret_cast<const double2*>(st.buf + c);
            const double2 xy = p[0];
            const double2 zw = p[1];
            const double dx = qx - xy.x, dy = qy - xy.y, dz = qz - zw.x;
            const double d2 = dx * dx + dy * dy + dz * dz;
            const long long wv = __double_as_longlong(zw.y);
            const int idx = (int)(wv & 0xffffffffll);
            if (d2 < best.d2 || (d2 == best.d2 && idx < best.idx)) {
                best.d2 = d2;
                best.idx = idx;
                best.pos = (int)(wv >> 32);
            }
        }
    }
    __syncwarp();
    st.n = 0;
}

// Load sorted point `pos`, tag it with its position; returns whether it lies inside the group's union box.
template <int G>
__device__ __forceinline__ bool load_candidate(const GroupStage<G>& st, const Point4* __restrict__ pts, uint32_t pos, double2& xy,
                                               double2& zw) {
    const double2* p = reinterpret_cast<const double2*>(pts + pos);
    xy = __ldg(p);
    zw = __ldg(p + 1);
    // w lane: low 32 bits = cloud index (already there), high 32 bits = sorted position
    const long long wv = (__double_as_longlong(zw.y) & 0xffffffffll) | ((long long)pos << 32);
    zw.y = __longlong_as_double(wv);
    return xy.x >= (double)st.lo[0] && xy.x <= (double)st.hi[0] && xy.y >= (double)st.lo[1] && xy.y <= (double)st.hi[1] &&
           zw.x >= (double)st.lo[2] && zw.x <= (double)st.hi[2];
}

// append this lane's point (if keep) to its group's list; ballot-compacted inside the group
template <int G>
__device__ __forceinline__ void group_append(GroupStage<G>& st, bool keep, const double2& xy, const double2& zw) {
    const int lane = threadIdx.x & 31;
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const unsigned gbits = (G == 32) ? bal : ((bal >> (lane & ~(G - 1))) & ((1u << G) - 1u));
    if (keep) {
        double2* d = reinterpret_cast<double2*>(st.buf + st.n + __popc(gbits & ((1u << (lane & (G - 1))) - 1u)));
        d[0] = xy;
        d[1] = zw;
    }
    st.n += __popc(gbits);
}

// Each lane contributes the sorted-cloud range [s, s+cnt); ranges may be empty.  Staged in rounds of
// at most ROUND points per lane, so a round always fits a group's list after a flush.
template <int G>
__device__ __forceinline__ void coop_stage_ranges(GroupStage<G>& st, Best& best, const Point4* __restrict__ pts, uint32_t s,
                                                  uint32_t cnt, double qx, double qy, double qz) {
    while (true) {
        const uint32_t take = min(cnt, (uint32_t)GroupStage<G>::ROUND);
        const uint32_t tmax = __reduce_max_sync(0xffffffffu, take);
        if (tmax == 0) break;
        const int room = __reduce_max_sync(0xffffffffu, st.n) + G * (int)tmax;
        if (room > GroupStage<G>::CAP) coop_flush<G>(st, best, qx, qy, qz);
        for (uint32_t k = 0; k < tmax; k++) {
            double2 xy, zw;
            bool keep = false;
            if (k < take) keep = load_candidate<G>(st, pts, s + k, xy, zw);
            group_append<G>(st, keep, xy, zw);
        }
        s += take;
        cnt -= take;
    }
}

// One contiguous range [first, last) (warp-uniform) staged by every group for which `want` holds,
// each with its own G lanes and its own filter.
template <int G>
__device__ __forceinline__ void coop_stage_span(GroupStage<G>& st, Best& best, const Point4* __restrict__ pts, uint32_t first,
                                                uint32_t last, bool want, double qx, double qy, double qz) {
    const int lg = (threadIdx.x & 31) & (G - 1);
    for (uint32_t t0 = first; t0 < last; t0 += G) {
        if (__reduce_max_sync(0xffffffffu, st.n) + G > GroupStage<G>::CAP) coop_flush<G>(st, best, qx, qy, qz);
        const uint32_t t = t0 + lg;
        double2 xy, zw;
        bool keep = false;
        if (want && t < last) keep = load_candidate<G>(st, pts, t, xy, zw);
        group_append<G>(st, keep, xy, zw);
    }
}

// Stage everything inside each group's cell box [lo, hi] (group-uniform ints; an empty box has
// hiy < loy) and scan it.
template <int G>
__device__ __forceinline__ void coop_scan_box(GroupStage<G>& st, Best& best, const GridView& v, const GridParams& g, int lox, int loy,
                                              int loz, int hix, int hiy, int hiz, double qx, double qy, double qz) {
    const int lg = (threadIdx.x & 31) & (G - 1);
    const bool some = hiy >= loy;
    const int ny = some ? hiy - loy + 1 : 0, nz = some ? hiz - loz + 1 : 0;
    const long long nrows_ll = (long long)ny * nz;
    const int bx0 = lox >> 2, bx1 = hix >> 2;
    const int nxb = some ? bx1 - bx0 + 1 : 0;
    const bool list_mode = nrows_ll * nxb > 8ll * g.n_blocks + 256;
    if (!__any_sync(0xffffffffu, list_mode)) {
        const int nrows = (int)nrows_ll;
        const int nrows_max = __reduce_max_sync(0xffffffffu, nrows);
        const int nxb_max = __reduce_max_sync(0xffffffffu, nxb);
        for (int row0 = 0; row0 < nrows_max; row0 += G) {
            const int row = row0 + lg;
            const bool rv = row < nrows;
            const int rz = rv ? row / ny : 0;
            const int iy = loy + (row - rz * ny), iz = loz + rz;
            for (int j = 0; j < nxb_max; j++) {
                const int cbx = bx0 + j;
                uint32_t s = 0, cnt = 0;
                if (rv && j < nxb) {
                    const int cby = iy >> 2, cbz = iz >> 2;
                    const uint32_t key = (uint32_t)((cbz * g.by + cby) * g.bx + cbx);
                    const int b = find_block(v.hash, g.hash_mask, key, block_hash(cbx, cby, cbz));
                    if (b >= 0) {
                        const int x0 = max(lox, cbx * 4) - cbx * 4, x1 = min(hix, cbx * 4 + 3) - cbx * 4;
                        const uint32_t* cs = v.cell_start + (size_t)b * 64 + (((iz & 3) << 4) | ((iy & 3) << 2));
                        s = __ldg(cs + x0);
                        cnt = __ldg(cs + x1 + 1) - s;
                    }
                }
                coop_stage_ranges<G>(st, best, v.pts, s, cnt, qx, qy, qz);
            }
        }
    } else {
        // some box covers far more rows than there are blocks (far or outlying queries): walk the
        // block list once for the whole warp; every group stages the blocks that touch its box
        const int by0 = loy >> 2, by1 = hiy >> 2, bz0 = loz >> 2, bz1 = hiz >> 2;
        const int lane = threadIdx.x & 31;
        // warp-wide union of the block boxes, to preselect blocks
        const int big = 0x7fffffff;
        const int wx0 = __reduce_min_sync(0xffffffffu, some ? bx0 : big), wx1 = __reduce_max_sync(0xffffffffu, some ? bx1 : -1);
        const int wy0 = __reduce_min_sync(0xffffffffu, some ? by0 : big), wy1 = __reduce_max_sync(0xffffffffu, some ? by1 : -1);
        const int wz0 = __reduce_min_sync(0xffffffffu, some ? bz0 : big), wz1 = __reduce_max_sync(0xffffffffu, some ? bz1 : -1);
        for (int b0 = 0; b0 < g.n_blocks; b0 += 32) {
            const int b = b0 + lane;
            bool hit = false;
            uint32_t key = 0;
            if (b < g.n_blocks) {
                key = __ldg(v.block_key + b);
                const int cbx = (int)(key % (uint32_t)g.bx);
                const int cby = (int)((key / (uint32_t)g.bx) % (uint32_t)g.by);
                const int cbz = (int)(key / ((uint32_t)g.bx * (uint32_t)g.by));
                hit = cbx >= wx0 && cbx <= wx1 && cby >= wy0 && cby <= wy1 && cbz >= wz0 && cbz <= wz1;
            }
            unsigned m = __ballot_sync(0xffffffffu, hit);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int bb = b0 + src;
                const uint32_t kk = __shfl_sync(0xffffffffu, key, src);
                const int cbx = (int)(kk % (uint32_t)g.bx);
                const int cby = (int)((kk / (uint32_t)g.bx) % (uint32_t)g.by);
                const int cbz = (int)(kk / ((uint32_t)g.bx * (uint32_t)g.by));
                const bool want = some && cbx >= bx0 && cbx <= bx1 && cby >= by0 && cby <= by1 && cbz >= bz0 && cbz <= bz1;
                coop_stage_span<G>(st, best, v.pts, __ldg(v.block_first + bb), __ldg(v.block_first + bb + 1), want, qx, qy, qz);
            }
        }
    }
    coop_flush<G>(st, best, qx, qy, qz);
}

// Exact NN for the warp's 32 queries.  `best` holds a valid upper bound (a real point of the cloud)
// for lanes that have one, d2 = +inf otherwise.  Inactive lanes take part in the collectives only.
//
// Expanding search: every lane asks for the ball of radius min(its bound, rcap); the union box of
// the group's balls is staged and scanned; a lane is finished once its best distance is within the
// radius it asked for (nothing unseen can be closer).  A group's rcap doubles until all its lanes
// are finished or its box is the whole grid.  With a previous-iteration neighbour as the bound this
// is a single pass over a box ~1-2 cells wider than the group's own footprint.
template <int G>
__device__ __forceinline__ void coop_search(GroupStage<G>& st, Best& best, const GridView& v, const GridParams& g, bool active,
                                            double qx, double qy, double qz) {
    const double tqx = (qx - g.ox) * g.inv_cell, tqy = (qy - g.oy) * g.inv_cell, tqz = (qz - g.oz) * g.inv_cell;
    const int big = 0x7fffffff;
    const float inf = __int_as_float(0x7f800000);
    float rcap = (float)(2.0 * g.cell);
    bool group_done = false;     // group-uniform
    while (true) {
        // float sqrt rounded up: the box only has to CONTAIN the ball
        const float rl = sqrtf((float)best.d2) * 1.000001f + 1e-30f;
        const float ruse = fminf(rl, rcap);
        const double rc = (double)ruse * g.inv_cell * (1.0 + 1e-12) + 1e-9;
        const bool part = active && !group_done;
        int lox = gmin_i<G>(part ? cell_coord(tqx - rc, g.nx) : big);
        int loy = gmin_i<G>(part ? cell_coord(tqy - rc, g.ny) : big);
        int loz = gmin_i<G>(part ? cell_coord(tqz - rc, g.nz) : big);
        int hix = gmax_i<G>(part ? cell_coord(tqx + rc, g.nx) : -1);
        int hiy = gmax_i<G>(part ? cell_coord(tqy + rc, g.ny) : -1);
        int hiz = gmax_i<G>(part ? cell_coord(tqz + rc, g.nz) : -1);
        const bool some = hix >= lox;    // the group still has work
        const bool whole = some && lox == 0 && loy == 0 && loz == 0 && hix == g.nx - 1 && hiy == g.ny - 1 && hiz == g.nz - 1;
        // the same union in world coordinates, for the per-point filter; when the box is the whole
        // grid everything is looked at once and the answer is final, so nothing may be filtered
        const double rw = (double)ruse * (1.0 + 1e-9);
        const float l0 = gmin_f<G>(part ? __double2float_rd(qx - rw) : inf), h0 = gmax_f<G>(part ? __double2float_ru(qx + rw) : -inf);
        const float l1 = gmin_f<G>(part ? __double2float_rd(qy - rw) : inf), h1 = gmax_f<G>(part ? __double2float_ru(qy + rw) : -inf);
        const float l2 = gmin_f<G>(part ? __double2float_rd(qz - rw) : inf), h2 = gmax_f<G>(part ? __double2float_ru(qz + rw) : -inf);
        st.lo[0] = whole ? -inf : l0; st.hi[0] = whole ? inf : h0;
        st.lo[1] = whole ? -inf : l1; st.hi[1] = whole ? inf : h1;
        st.lo[2] = whole ? -inf : l2; st.hi[2] = whole ? inf : h2;
        if (!__any_sync(0xffffffffu, some)) return;
        if (!some) { loy = 0; hiy = -1; }      // empty box for groups that are done
        coop_scan_box<G>(st, best, v, g, lox, loy, loz, hix, hiy, hiz, qx, qy, qz);
        const bool fin = !part || (sqrtf((float)best.d2) * 1.000001f + 1e-30f <= ruse);
        const bool all_fin = __all_sync(group_mask<G>(), fin);
        group_done = group_done || whole || all_fin;
        if (__all_sync(0xffffffffu, group_done)) return;
        rcap *= 2.0f;
    }
}

}  // namespace r3d
