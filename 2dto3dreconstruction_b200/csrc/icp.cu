// K2 query + K3: the ICP loop of utils/icp.py:74-133, batched over pairs, entirely on the device.
//
// Per iteration (no host round trip; converged pairs turn into early-exit CTAs):
//   nn_query_kernel  q = T_cum * a (fp64), exact NN on the sparse-block grid, one warp per 32
//                    spatially adjacent queries (lean_search.cuh: warp-union walk, and a group search
//                    for tasks with few searching lanes; the source cloud is sorted by cell key
//                    once, before the loop).  Iterations >= 2 seed the upper bound with
//                    the previous neighbour and skip the search altogether when the neighbour
//                    provably cannot change.  With partial_set_size == 1 the 16 Kabsch moments are
//                    reduced in the same kernel.
//   (partial < 1)    selection of the cutoff-th smallest squared distance = np.argpartition(distances, cutoff-1)[:cutoff],
//                    utils/icp.py:109-110: the first histogram pass is counted inside nn_query_kernel, fsel_hist_kernel
//                    counts the second and brackets the cutoff (small pairs: sel_small_kernel, one launch)
//   moments_kernel   (partial < 1) moments over the kept correspondences, list of the bracket's members
//   solve_kernel     fixed-order sum of the partial rows, 3x3 Kabsch (utils/icp.py:22-55),
//                    T_cum <- T_iter * T_cum, mean error, |prev - mean| < tol test (icp.py:123-126);
//                    resolve_solve_kernel (partial < 1, large pairs): the bracket's exact finish, then the same step
// After the loop, final_kernel reproduces utils/icp.py:131 best_fit_transform(A, src) from the
// second moments of A (src is the affine image T_cum * A, so H = C_A * M^T in closed form).
#include <type_traits>
#include "lean_search.cuh"
#include <mutex>
#include <vector>

namespace r3d {

constexpr int NN_THREADS = 128;
// resident CTAs per SM the kernel is compiled for: the search is latency bound and wants 8 (64 registers)
#ifndef R3D_LEAN_MIN_CTAS
#define R3D_LEAN_MIN_CTAS 8
#endif
// partial-set selection (radix select, below): passes x bins of the per-pair histograms in the workspace
constexpr int SEL_PASSES = 6;
constexpr int SEL_BINS = 2048;
constexpr int SEL_TILE = 4096;      // keys per CTA of the histogram kernel
constexpr int FSEL_PASSES = 2;      // histogram passes of the large-pair selection (11 bits of the fp32 key each)
constexpr int NN_VWARPS = 2048;      // virtual warps per pair of the persistent nn kernel (see nn_kernel) ...
constexpr int NN_VWARPS_BIG = 8192;  // ... and for a pair of more than NN_VWARPS_BIG_FROM source points: a single large pair must
constexpr int NN_VWARPS_BIG_FROM = 1 << 19;      // still fill the GPU.  A function of the pair alone, so batching stays irrelevant.
// Queries per task of the lean kernel: a warp normally takes 32 consecutive queries; a SMALL pair (the reference's
// own sizes: 13k-26k points after its sub-sampling) would then give a few hundred warps, one per SM sub-partition,
// each running its whole search serially.  Such pairs use 4 queries per task, every query on eight lanes that each
// evaluate an eighth of the candidates: eight times the warps, an eighth of the per-task serial scan.  Also a
// function of the pair alone.  (8 / 4 / 2 queries per task on the sofa pair: 149 / 140 / 164 us per launch at
// partial_set_size 1.0, 203 / 163 / 149 us at 0.7.)
constexpr int NN_SMALL_PAIR = 1 << 15;
// Pairs of at most this many source points take the one-CTA selection (sel_small_kernel: six 64-bit histogram passes back to
// back); larger ones the bracket selection (first pass in the nn kernel, fsel_hist_kernel, moments_kernel, resolve_solve_kernel).
// (0 = every pair through the bracket selection: correct, and 12 % slower on the reference's 13 k-point pairs -- 16.3 against
// 14.6 ms for 100 iterations -- whose whole key set a single CTA walks in a few microseconds.)
#ifndef R3D_SEL_SMALL_MAX
#define R3D_SEL_SMALL_MAX (1 << 15)
#endif
constexpr int SEL_SMALL_MAX = R3D_SEL_SMALL_MAX;
#ifndef R3D_NN_MID_PAIR
#define R3D_NN_MID_PAIR 0
#endif
constexpr int NN_MID_PAIR = R3D_NN_MID_PAIR;      // up to this many points: 16 queries per task on two lanes each (0 = off)
#ifndef R3D_NN_SMALL_QPT
#define R3D_NN_SMALL_QPT 4
#endif
__host__ __device__ __forceinline__ int queries_per_task(int n_src) {
    return n_src <= NN_SMALL_PAIR ? R3D_NN_SMALL_QPT : (n_src <= NN_MID_PAIR ? 16 : 32);
}
__host__ __device__ __forceinline__ int vwarps_for(int n_src) { return n_src > NN_VWARPS_BIG_FROM ? NN_VWARPS_BIG : NN_VWARPS; }

struct PairState {
    double T[12];        // cumulative transform, rows 0..2 of the 4x4
    double Tp[12];       // the cumulative transform one iteration earlier (how far did every query just move?)
    double prev_err;
    double mean_err;
    double tau;          // partial mode: squared-distance threshold
    int idx_cut;         // partial mode: ties at tau are kept up to this source index
    int iters;           // reference convention (starts at 1)
    int done;
    int n_src;
    int cutoff;
    int pad;
};

struct SelState {          // radix-select progress of one pair
    unsigned long long prefix;
    uint32_t k;            // rank still to find inside the prefix
    uint32_t neq;          // keys in the chosen bin
    int finished;          // the descent ended early: every key of the chosen bin belongs to the kept set
    int pad;
};

struct FSel {              // selection of a large pair: where the k-th smallest squared distance lies
    uint32_t prefix;       // after pass p: the top 11 (p + 1) bits of its fp32 key
    uint32_t k;            // its rank among the keys that share the prefix
    uint32_t all;          // cutoff >= n: everything is kept
    uint32_t pad;
};

struct IcpWorkspace {
    GridWorkspace grid;
    int shard_rank, shard_world;      // query sharding of r3d_icp_batch_sharded: this process runs virtual warps rank, rank + world, ...
    int n_pairs, max_src, src_tiles, src_rows;
    PairState* state;      // [pair]
    SortDesc* src_desc;    // [pair]
    uint32_t* skeys[2];    // [pair][max_src]  source sort keys (ping-pong)
    uint32_t* svals[2];    // [pair][max_src]
    uint32_t* shist;       // [pair][256][src radix tiles]
    int src_rtiles;
    Point4* src_sorted;    // [pair][max_src]  source cloud in cell-key order, idx = original index
    int* vw_counter;       // [pair] next virtual warp of the running nn launch (reset by its consumer)
    int32_t* nn_pos;       // [pair][max_src]  sorted position of the current neighbour
    float* budget;         // [pair][max_src]  lower bound of the distance from the query to every point but its neighbour
    double* d2;            // [pair][max_src]
    double* partial;       // [pair][src_rows][16]  one row of moment sums per query warp
    double* amom;          // [pair][16]  moments of A
    SelState* sel;         // [pair]
    uint32_t* sel_hist;    // [pair][FSEL_PASSES][SEL_BINS]  zero between selections
    FSel* fsel;            // [pair]  selection state of the large pairs
    uint32_t* fsel_done;   // [pair][FSEL_PASSES]  histogram CTAs that have finished (the last one picks); zero between selections
    uint32_t* cand_n;      // [pair]  correspondences inside the bracket of the k-th smallest distance ...
    int32_t* cand_idx;     // [pair][max_src]  ... and their (sorted) source positions, in arrival order
};

static size_t icp_workspace_layout(IcpWorkspace& w, Arena& a, int n_pairs, int max_src, int max_dst) {
    grid_workspace_layout(w.grid, a, n_pairs, max_dst);
    w.n_pairs = n_pairs;
    w.max_src = max_src;
    w.shard_rank = 0;
    w.shard_world = 1;
    w.src_tiles = (max_src + NN_THREADS - 1) / NN_THREADS;      // CTAs per pair
    // rows of `partial`: one per warp of the widest launch (whole CTAs, so round up per CTA)
    w.src_rows = w.src_tiles * (NN_THREADS / 32) + 1;      // (+ 1: the bracket row of fsel_resolve_body)
    if (w.src_rows < vwarps_for(max_src)) w.src_rows = vwarps_for(max_src);
    const size_t P = (size_t)n_pairs;
    w.state = a.take<PairState>(P);
    w.src_desc = a.take<SortDesc>(P);
    w.src_rtiles = radix_tiles(max_src);
    for (int k = 0; k < 2; k++) {
        w.skeys[k] = a.take<uint32_t>(P * max_src);
        w.svals[k] = a.take<uint32_t>(P * max_src);
    }
    w.shist = a.take<uint32_t>(P * 256 * w.src_rtiles);
    w.src_sorted = a.take<Point4>(P * max_src);
    w.vw_counter = a.take<int>(P);
    w.nn_pos = a.take<int32_t>(P * max_src);
    w.budget = a.take<float>(P * max_src);
    w.d2 = a.take<double>(P * max_src);
    w.partial = a.take<double>(P * w.src_rows * 16);
    w.amom = a.take<double>(P * 16);
    w.sel = a.take<SelState>(P);
    w.sel_hist = a.take<uint32_t>(P * FSEL_PASSES * SEL_BINS);
    w.fsel = a.take<FSel>(P);
    w.fsel_done = a.take<uint32_t>(P * FSEL_PASSES);
    w.cand_n = a.take<uint32_t>(P);
    w.cand_idx = a.take<int32_t>(P * max_src);
    return a.used;
}

// ---- state -------------------------------------------------------------------------------------
__global__ void state_init_kernel(IcpWorkspace w, const int32_t* __restrict__ src_off, const double* __restrict__ init_pose,
                                  double partial_set_size) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= w.n_pairs) return;
    PairState s;
    for (int k = 0; k < 12; k++) s.T[k] = init_pose ? init_pose[(size_t)pair * 16 + k] : ((k % 5 == 0) ? 1.0 : 0.0);
    for (int k = 0; k < 12; k++) s.Tp[k] = s.T[k];
    s.prev_err = 0.0;
    s.mean_err = 0.0;
    s.tau = CUDART_INF;      // keep everything until the selection says otherwise
    s.idx_cut = 0x7fffffff;
    s.iters = 1;
    s.done = 0;
    s.n_src = src_off[pair + 1] - src_off[pair];
    s.cutoff = (int)((double)s.n_src * partial_set_size);   // int(N * p), utils/icp.py:109
    s.pad = 0;
    if (s.n_src <= 0 || w.grid.params[pair].n_pts <= 0) s.done = 1;
    // int(N * p) == 0: the reference fits a transform to ZERO correspondences (means of empty arrays, SVD of NaN ->
    // LinAlgError).  Here such a pair does nothing and reports iters = 0, a value the loop cannot produce otherwise;
    // the Python drop-in raises on it (utils/icp.py).
    if (s.cutoff <= 0 && !s.done) { s.done = 1; s.iters = 0; }
    w.state[pair] = s;
    w.vw_counter[pair] = 0;
    w.src_desc[pair].n = s.n_src > 0 ? s.n_src : 0;
    w.src_desc[pair].n_passes = w.grid.params[pair].n_passes;
}

// ---- one-time spatial sort of the source cloud, and query groups ------------------------------------
// key = (block << 6 | octant-major fine cell) of the destination-grid cell of the (initially posed)
// source point, clamped into the grid.  Queries of one block (4x4x4 cells) are contiguous after the
// sort and, inside a block, ordered octant by octant.  Groups of up to G consecutive queries are then
// cut so that no group straddles two blocks: a group's queries are within one block edge of each
// other, and rigid updates move them together, so the union box of a group stays a few cells wide
// for every iteration.  (Cutting fixed-size chunks of a space-filling-curve order does not work for
// a surface: consecutive occupied cells along the curve can be arbitrarily far apart -- measured
// union boxes of up to the whole scene with a Hilbert order.)
__global__ void __launch_bounds__(256)
src_keys_kernel(IcpWorkspace w, const Point4* __restrict__ src, const int32_t* __restrict__ src_off) {
    const int pair = blockIdx.y;
    __shared__ GridParams g;
    __shared__ double T[12];
    __shared__ int s_n;
    if (threadIdx.x == 0) { g = w.grid.params[pair]; s_n = w.state[pair].n_src; }
    if (threadIdx.x < 12) T[threadIdx.x] = w.state[pair].T[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= s_n) return;
    const Point4 a = load_point(src + src_off[pair] + i);
    const double qx = T[0] * a.x + T[1] * a.y + T[2] * a.z + T[3];
    const double qy = T[4] * a.x + T[5] * a.y + T[6] * a.z + T[7];
    const double qz = T[8] * a.x + T[9] * a.y + T[10] * a.z + T[11];
    const int ix = cell_coord((qx - g.ox) * g.inv_cell, g.nx);
    const int iy = cell_coord((qy - g.oy) * g.inv_cell, g.ny);
    const int iz = cell_coord((qz - g.oz) * g.inv_cell, g.nz);
    // Block order of the QUERY sort: the axis with the fewest blocks runs fastest.  A depth-image cloud
    // is a height field over its two long axes, so stepping along a long axis moves to a block that
    // touches the previous one; with the destination grid's own x-fastest order a run of consecutive
    // queries would keep jumping between far-apart blocks of one block row.
    const int cb[3] = {ix >> 2, iy >> 2, iz >> 2};
    const int nb[3] = {g.bx, g.by, g.bz};
    int o0 = 0, o1 = 1, o2 = 2;      // o0 fastest ... o2 slowest
    if (nb[o1] < nb[o0]) { const int t = o0; o0 = o1; o1 = t; }
    if (nb[o2] < nb[o0]) { const int t = o0; o0 = o2; o2 = t; }
    if (nb[o2] < nb[o1]) { const int t = o1; o1 = o2; o2 = t; }
    const int c0 = (o0 == 0) ? cb[0] : (o0 == 1) ? cb[1] : cb[2], n0 = (o0 == 0) ? nb[0] : (o0 == 1) ? nb[1] : nb[2];
    const int c1 = (o1 == 0) ? cb[0] : (o1 == 1) ? cb[1] : cb[2], n1 = (o1 == 0) ? nb[0] : (o1 == 1) ? nb[1] : nb[2];
    const int c2 = (o2 == 0) ? cb[0] : (o2 == 1) ? cb[1] : cb[2];
    const uint32_t blk = (uint32_t)((c2 * n1 + c1) * n0 + c0);
    const uint32_t fine = (uint32_t)((((iz >> 1) & 1) << 5) | (((iy >> 1) & 1) << 4) | (((ix >> 1) & 1) << 3) |
                                     ((iz & 1) << 2) | ((iy & 1) << 1) | (ix & 1));
    w.skeys[0][(size_t)pair * w.max_src + i] = (blk << 6) | fine;
    w.svals[0][(size_t)pair * w.max_src + i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
src_gather_kernel(IcpWorkspace w, const Point4* __restrict__ src, const int32_t* __restrict__ src_off) {
    const int pair = blockIdx.y;
    const int n = w.state[pair].n_src;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint32_t* vals = w.svals[w.src_desc[pair].n_passes & 1] + (size_t)pair * w.max_src;
    const uint32_t v = vals[i];
    const Point4 a = load_point(src + src_off[pair] + v);
    store_point(w.src_sorted + (size_t)pair * w.max_src + i, a.x, a.y, a.z, (int32_t)v);
}

// ---- per-warp reduction of 16 doubles per lane (no CTA barrier) --------------------------------------
// `scratch` is >= 16*33 doubles of shared memory owned by the warp.  Rows are skewed by one double
// so that the 16 row-summing lanes hit distinct banks.  Fixed summation order: deterministic.
__device__ __forceinline__ double warp_reduce16(const double (&m)[16], double* scratch) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; k++) scratch[k * 33 + lane] = m[k];
    __syncwarp();
    const int row = lane & 15, half = lane >> 4;
    const double* r = scratch + row * 33 + half * 16;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 16; j++) s += r[j];
    s += __shfl_down_sync(0xffffffffu, s, 16);
    __syncwarp();
    return s;      // lanes 0..15 hold the sum of moment `lane`
}

// Same sums in two passes of eight moments, with half the scratch (8 * 33 doubles per warp).
__device__ __forceinline__ double warp_reduce16_small(const double (&m)[16], double* scratch) {
    const int lane = threadIdx.x & 31;
    const int row = lane & 7, quarter = lane >> 3;
    double result = 0.0;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) scratch[k * 33 + lane] = m[pass * 8 + k];
        __syncwarp();
        const double* r = scratch + row * 33 + quarter * 8;
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) s += r[j];
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if (quarter == pass) result = s;      // lanes 0..7: moments 0..7; lanes 8..15: moments 8..15
    }
    __syncwarp();
    return result;      // lanes 0..15 hold the sum of moment `lane`
}

__device__ __forceinline__ void warp_reduce16_store(const double (&m)[16], double* scratch, double* __restrict__ out16) {
    const double s = warp_reduce16(m, scratch);
    if ((threadIdx.x & 31) < 16) out16[threadIdx.x & 31] = s;
}

__device__ __forceinline__ void accumulate_pair(double (&m)[16], double ax, double ay, double az, double bx, double by,
                                                double bz, double dist) {
    m[0] = ax; m[1] = ay; m[2] = az;
    m[3] = bx; m[4] = by; m[5] = bz;
    m[6] = ax * bx; m[7] = ax * by; m[8] = ax * bz;
    m[9] = ay * bx; m[10] = ay * by; m[11] = ay * bz;
    m[12] = az * bx; m[13] = az * by; m[14] = az * bz;
    m[15] = dist;
}

// fp32 key of a squared distance for the top-fraction selection (monotone: rounded towards zero); see fsel_hist_kernel
__device__ __forceinline__ uint32_t key32(double d2) { return __float_as_uint(__double2float_rz(d2)); }

// ---- nearest neighbour ----------------------------------------------------------------------------
// Persistent query kernel: grid (CTAs per pair, pairs); a task is 32 consecutive queries of the sorted source cloud.
// Work is dealt to vwarps_for(n_src) "virtual warps" per pair (virtual warp u owns tasks u, u + nvw, ...), handed
// out dynamically, and each physical warp runs some of them back to back.  The moment sums are kept per VIRTUAL
// warp, so their summation order -- and therefore every bit of the result -- does not depend on how many CTAs
// this launch happened to get (chunk size, number of pairs in flight, GPU count).
// IMPL = 1: the search of lean_search.cuh (a second, TMA-staged tile search lived here in round 2; it was exact and slower,
// profiles/r02_nn_tile_search.md, and is gone again).
//
// Skipping searches that cannot change the answer.  When a seeded search ends with the seed p still the
// nearest point, it also knows a lower bound L2 of the distance to every OTHER point (the runner-up among
// the evaluated points, or the radius it covered, whichever is smaller); that bound is the query's stored
// `budget`.  If the query later moves by delta in total, every other distance shrinks by at most delta
// (triangle inequality), so while d(q_now, p) < L2 - sum(delta) the neighbour is still p, strictly, and no
// search is needed -- only the exact new distance to p, which the seed evaluation computes anyway (a query
// that moves TOWARDS its neighbour, the usual case in a converging ICP, keeps its budget longer than the
// symmetric bound L2 - d - 2 sum(delta) > 0 would allow).  The per-iteration movement |T_k a - T_{k-1} a| is
// computed exactly per query and subtracted from the stored bound with outward rounding.  To have a bound worth keeping, a seeded
// lane asks for a ball a little larger than its bound (`slack`, proportional to its current movement):
// costs a few candidates now, saves whole searches once ICP has nearly converged.
struct LeanTuning {
    float slack_mult;        // slack = slack_mult * (movement of the query in this iteration) ...
    float slack_max_frac;    // ... if that is at most slack_max_frac * cell (else 0: still moving too fast to bother)
    int group_max;           // tasks with at most this many searching lanes go through group_search (0 = never)
    int slack_clamp;         // a query that moves faster than that still asks for slack_max_frac * cell (else for nothing extra)
};

#ifndef R3D_NN_PREFETCH
#define R3D_NN_PREFETCH 0
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int IMPL> struct SearchSmem;
template <> struct SearchSmem<1> { double2 list[LEAN_WARP_SMEM]; };      // survivors list; never shared with the reduction: its stale slots must stay real points

// SMALL: every pair of the launch has at most NN_SMALL_PAIR source points (decided by the host from max_src): the search is
// compiled with the block pruning of lean_search.cuh (PRUNE).
template <int IMPL, bool SEEDED, bool FUSED, bool SMALL>
__global__ void __launch_bounds__(NN_THREADS, R3D_LEAN_MIN_CTAS)
nn_query_kernel(IcpWorkspace w, int write_d2, LeanTuning tune) {
    const int pair = blockIdx.y;
    __shared__ GridParams g;
    __shared__ double T[12], D[12];
    __shared__ int s_n, s_done, s_cutoff;
    // the grid's base pointers live in shared memory: the compiler would otherwise recompute these 64-bit
    // addresses inside every loop of the search
    __shared__ GridView s_view;
    __shared__ __align__(16) double scratch[NN_THREADS / 32][8 * 33];      // moment reduction scratch (two passes of 8)
    __shared__ __align__(16) SearchSmem<IMPL> ssm[NN_THREADS / 32];
    if (threadIdx.x == 0) {
        g = w.grid.params[pair];
        s_n = w.state[pair].n_src;
        s_done = w.state[pair].done;
        s_cutoff = w.state[pair].cutoff;
        s_view = grid_view(w.grid, pair);
    }
    if (threadIdx.x < 12) {
        T[threadIdx.x] = w.state[pair].T[threadIdx.x];
        D[threadIdx.x] = w.state[pair].T[threadIdx.x] - w.state[pair].Tp[threadIdx.x];
    }
    __syncthreads();
    if (s_done) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const GridView& v = s_view;
    const Point4* src = w.src_sorted + (size_t)pair * w.max_src;
    int32_t* nn_pos = w.nn_pos + (size_t)pair * w.max_src;
    float* budget = w.budget + (size_t)pair * w.max_src;
    double* d2_out = w.d2 + (size_t)pair * w.max_src;
    // The task size is a compile-time constant in the launches whose pairs are all small (QPT = R3D_NN_SMALL_QPT: 1-4 % off
    // their nn launches) and read from the pair's size otherwise (QPT = 0).  A second, QPT = 32 copy of the loop for the large
    // pairs was measured: 1.7 % fewer warp instructions (the replica paths fold away), no change in time, twice the code.
    auto run = [&](auto qpt_c) {
    constexpr int QPT = decltype(qpt_c)::value;
    const int qpt = QPT > 0 ? QPT : queries_per_task(s_n);
    const int n_tasks = (s_n + qpt - 1) / qpt;
    const int nvw = vwarps_for(s_n);
    const float slack_max = tune.slack_max_frac * (float)g.cell;
    int* vw_counter = w.vw_counter + pair;
    // (selection of a large pair at partial_set_size < 1: see the epilogue of the task loop)
    uint32_t* hist0 = (!FUSED && write_d2 == 2 && s_n > SEL_SMALL_MAX && s_cutoff < s_n)
                          ? w.sel_hist + ((size_t)pair * FSEL_PASSES + 0) * SEL_BINS : nullptr;
    unsigned phase = 0;      // parity of the warp's staging mbarrier(s)
    if constexpr (IMPL == 1) {
        list_init(ssm[warp].list, v.pts);
#if R3D_LEAN_TMA
        if (lane == 0) {
            mbar_init(smem_u32(ssm[warp].list + 2 * LEAN_LIST_CAP + 128), 1);
            mbar_init(smem_u32(ssm[warp].list + 2 * LEAN_LIST_CAP + 128) + 8u, 1);
        }
        mbar_init_fence();
        __syncwarp();
#endif
    }
    while (true) {
        int vw = 0;
        if (lane == 0) vw = atomicAdd(vw_counter, 1);
        vw = __shfl_sync(0xffffffffu, vw, 0) * w.shard_world + w.shard_rank;      // (1, 0 unless the queries are sharded over GPUs)
        if (vw >= nvw) break;
        double acc = 0.0;      // lanes 0..15: running sum of moment `lane` over this virtual warp's tasks
        for (int task = vw; task < n_tasks; task += nvw) {
            const int i = task * qpt + (lane & (qpt - 1));      // qpt < 32: 32 / qpt lanes per query (replicas)
            const bool active = i < s_n;
            const bool primary = lane < qpt;                    // the replica that writes the results
#if R3D_NN_PREFETCH
            // The virtual warp's next task is 2048 tasks away: nothing of it is in cache.  Once most lanes skip their
            // search a task is a chain of dependent loads (query, previous neighbour's position, that point), so ask
            // for the next task's lines now and for its neighbour points at the end of this task.
            int pos_next = -1;
            {
                const int in = (task + nvw) * qpt + (lane & (qpt - 1));
                if (task + nvw < n_tasks && in < s_n) {
                    prefetch_l2(src + in);
                    if (SEEDED) {
                        pos_next = nn_pos[in];
                        prefetch_l2(budget + in);
                    }
                }
            }
#endif
            double qx = 0.0, qy = 0.0, qz = 0.0;
            Best seed, other;
            seed.d2 = CUDART_INF; seed.idx = 0x7fffffff; seed.pos = -1;
            other.d2 = CUDART_INF; other.idx = 0x7fffffff; other.pos = -1;
            float left = 0.0f, slack = 0.0f;      // bound on the other points left after this iteration's movement
            bool skip = false;
            if (active) {
                const Point4 a = load_point(src + i);
                qx = T[0] * a.x + T[1] * a.y + T[2] * a.z + T[3];
                qy = T[4] * a.x + T[5] * a.y + T[6] * a.z + T[7];
                qz = T[8] * a.x + T[9] * a.y + T[10] * a.z + T[11];
                if (SEEDED) {
                    consider(seed, v.pts, nn_pos[i], qx, qy, qz);      // (at 64 registers the seed's coordinates are reloaded for the moment sums)
                    const double mx = D[0] * a.x + D[1] * a.y + D[2] * a.z + D[3];
                    const double my = D[4] * a.x + D[5] * a.y + D[6] * a.z + D[7];
                    const double mz = D[8] * a.x + D[9] * a.y + D[10] * a.z + D[11];
                    // rounded up; the last term covers the fp64 rounding of the two transformed positions themselves
                    // (a few ulps of the coordinates), which matters only when coordinates are ~1e9 times the distances
                    const float moved = sqrt_bound(__double2float_ru(mx * mx + my * my + mz * mz)) * 1.000001f +
                                        __double2float_ru(4.0e-15 * (fabs(qx) + fabs(qy) + fabs(qz))) + 1e-30f;
                    left = __fsub_rd(budget[i], moved);
                    skip = left > sqrt_bound(__double2float_ru(seed.d2)) * 1.000001f;
                    const float want = moved * tune.slack_mult;
                    slack = (want <= slack_max) ? want + 1e-4f * slack_max : (tune.slack_clamp ? slack_max : 0.0f);
                }
            }
            R3D_STAT(0, 1);
#ifdef R3D_NN_STATS
            const long long t_task = clock64();
#endif
            float cover = 0.0f, lb_other = 0.0f;
            if constexpr (IMPL == 1) {
                // tasks with few searching lanes: eight lanes per query on that query's own cells (group_search); everything
                // else, and whatever the groups could not take, through the warp-union walk
                unsigned walk_m = __ballot_sync(0xffffffffu, active && !skip);
                if (SEEDED && walk_m != 0u && tune.group_max > 0) {
                    if (qpt == 32) {
                        if (__popc(walk_m) <= tune.group_max) walk_m = group_search(walk_m, seed, other, v, g, qx, qy, qz, slack, cover);
                    } else {
                        // small pairs: the task's few queries sit on several lanes each (replicas); one group per query is
                        // exactly the group search's shape.  The primary replicas own the results; a query the groups could
                        // not take goes to the union walk with all its replicas.
                        const unsigned un = group_search(walk_m & ((1u << qpt) - 1u), seed, other, v, g, qx, qy, qz, slack, cover);
                        walk_m = __ballot_sync(0xffffffffu, active && !skip && ((un >> (lane & (qpt - 1))) & 1u));
                    }
                }
                if (walk_m != 0u) {
                    float cover_w = 0.0f;
                    // (the walk's per-lane query table shares the warp's reduction scratch: the two are never live at the same time)
                    lean_search<SMALL>(ssm[warp].list, reinterpret_cast<float4*>(scratch[warp]), phase, seed, other, v, g,
                                active && !skip && ((walk_m >> lane) & 1u), qx, qy, qz, slack, cover_w, qpt);
                    if ((walk_m >> lane) & 1u) cover = cover_w;
                }
                lb_other = sqrt_bound(__double2float_rd(other.d2));
            }
            const bool seed_wins = skip || !(other.d2 < seed.d2 || (other.d2 == seed.d2 && other.idx < seed.idx));
#ifdef R3D_NN_STATS
            R3D_STAT(3, __popc(__ballot_sync(0xffffffffu, active && skip && primary)));
            // a skipped lane of the lean search still evaluates the warp's candidates: none of them may beat its seed
            R3D_STAT(7, __popc(__ballot_sync(0xffffffffu, active && skip && (other.d2 < seed.d2 || (other.d2 == seed.d2 && other.idx < seed.idx)))));
            R3D_STAT(12, clock64() - t_task);
            R3D_STAT_MAX(13, clock64() - t_task);
#endif
            const Best best = seed_wins ? seed : other;
            if constexpr (!FUSED) {
                // partial_set_size < 1, large pairs: the first histogram pass of the selection (top 11 bits of the fp32 key of
                // every distance, fsel_hist_kernel) is counted here, where the distances are made: one pass over d2 and one
                // launch per iteration less.  Warp-aggregated: the 32 distances of a task fall into a handful of bins.
                if (hist0 != nullptr) {
                    const bool mine = active && primary;
                    const unsigned am = __ballot_sync(0xffffffffu, mine);
                    if (mine) {
                        const uint32_t bin = key32(best.d2) >> 21;
                        const unsigned grp = __match_any_sync(am, bin);
                        if ((grp & ((1u << lane) - 1u)) == 0u) atomicAdd(hist0 + bin, (uint32_t)__popc(grp));
                    }
                }
            }
            if (active && primary) {
                nn_pos[i] = best.pos;
                if (!FUSED || write_d2) d2_out[i] = best.d2;
                float nb = 0.0f;
                if (skip) {
                    nb = left;
                } else if (SEEDED && seed_wins) {
                    // lower bound of the distance to any point other than the seed: the nearest of the evaluated ones or
                    // the radius the search covered (float square roots: the 1e-6 factors cover their rounding)
                    nb = fminf(lb_other, cover) * 0.999999f;
                }
                budget[i] = nb;
            }
#if R3D_NN_PREFETCH
            if (SEEDED && pos_next >= 0) prefetch_l2(v.pts + pos_next);
#endif
            if (FUSED) {
                double m[16];
#pragma unroll
                for (int k = 0; k < 16; k++) m[k] = 0.0;
                if (active && primary) {
                    const Point4 b = load_point(v.pts + best.pos);
                    accumulate_pair(m, qx - g.shift[0], qy - g.shift[1], qz - g.shift[2], b.x - g.shift[0], b.y - g.shift[1],
                                    b.z - g.shift[2], sqrt(best.d2));
                }
                acc += warp_reduce16_small(m, scratch[warp]);
            }
        }
        if (FUSED && lane < 16) w.partial[((size_t)pair * w.src_rows + (size_t)vw) * 16 + lane] = acc;
    }
    };      // run
    if constexpr (SMALL) {
        run(std::integral_constant<int, (NN_MID_PAIR > 0 ? 0 : R3D_NN_SMALL_QPT)>{});
    } else {
        run(std::integral_constant<int, 0>{});
    }
}

// ---- partial-set selection ---------------------------------------------------------------------------
// k-th smallest of the non-negative doubles d2[0..n) by MSB-first radix select on their bit patterns
// (order preserving for non-negative IEEE doubles): 9 + 5 x 11 bits.  Every pass is a wide histogram
// kernel over all keys that still match the prefix (shared-memory bins, warp-aggregated, merged into
// the pair's global bins) followed by a one-CTA-per-pair pick of the bin that holds rank k.
// = np.argpartition(distances, cutoff - 1)[:cutoff], utils/icp.py:109-110.
__host__ __device__ __forceinline__ int sel_shift(int pass) { return pass == 0 ? 55 : 55 - 11 * pass; }
__host__ __device__ __forceinline__ int sel_bits(int pass) { return pass == 0 ? 9 : 11; }

// ---- selection of a large pair: two histogram passes on an fp32 key, a bracket, an exact finish ------------------
// The k-th smallest of n squared distances (n up to millions) used to take six histogram + six pick launches on the 64-bit
// keys.  Now: key32 = bits of the distance rounded TOWARDS ZERO to fp32 -- a monotone map, so the k-th smallest double has the
// k-th smallest key32 -- and two histogram passes (11 bits each; the first counted by the nn kernel, the second by
// fsel_hist_kernel, whose last CTA picks) leave a BRACKET: the distances whose key32 shares the 22 leading bits of the k-th
// one's, a relative width of 2^-13, a few dozen of 276 k.  Everything below the bracket is kept for certain and its moments are
// summed by moments_kernel in the same pass that lists the bracket's members; fsel_resolve_body (resolve_solve_kernel) then
// orders the members by (distance, position) exactly, keeps the first k', publishes tau / idx_cut (kept() below) and adds their
// moments as one more row.  Exact ties at the cutoff keep the lowest positions, as before; a bracket that holds many points
// (lattices: every distance the same double) takes a slower, still exact route through fsel_resolve_body.

// Which of the SEL_BINS counters of `gh` (global) holds rank k, and the rank left inside it: every thread of a CTA of 256 gets
// the answer (8 bins per thread, block scan).  zero: the counters are also reset for the next iteration's selection.
__device__ __forceinline__ void hist_pick256(uint32_t* gh, uint32_t k, bool zero, uint32_t* warp_tot /* [8] shared */,
                                             uint32_t* s_out /* [2] shared */, uint32_t& bin, uint32_t& kk) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t c[8], tot = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        c[j] = __ldcg(gh + threadIdx.x * 8 + j);
        if (zero) gh[threadIdx.x * 8 + j] = 0;
        tot += c[j];
    }
    uint32_t x = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();      // (warp_tot / s_out may still be read from an earlier call)
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    uint32_t before = x - tot;
    for (int wv = 0; wv < warp; wv++) before += warp_tot[wv];
    if (k >= before && k < before + tot) {      // exactly one thread
        uint32_t r = k - before;
        int j = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (j == q && r >= c[q]) { r -= c[q]; j = q + 1; }
        }
        s_out[0] = (uint32_t)(threadIdx.x * 8 + j);
        s_out[1] = r;
    }
    __syncthreads();
    bin = s_out[0];
    kk = s_out[1];
}

// The second (and last) histogram pass of a large pair's selection.  The FIRST pass -- the top 11 bits of every key -- is
// counted by the nn kernel itself as it produces the distances (nn_query_kernel, `hist0`): every CTA here starts by reading
// those 2048 counters and picking the bin of rank cutoff - 1 (the same answer in every CTA), counts the next 11 bits of the
// keys inside that bin, and the last CTA to finish picks again, publishes the 22-bit bracket and resets both histograms.
__global__ void __launch_bounds__(256)
fsel_hist_kernel(IcpWorkspace w) {
    const int pair = blockIdx.y;
    const PairState* st = w.state + pair;
    const int n = st->n_src;
    if (st->done || n <= SEL_SMALL_MAX) return;
    const int base = blockIdx.x * SEL_TILE;
    if (base >= n) return;
    FSel* fs = w.fsel + pair;
    const int cutoff = st->cutoff;
    if (cutoff >= n) {      // nothing to select: everything is kept
        if (blockIdx.x == 0 && threadIdx.x == 0) { fs->all = 1u; fs->prefix = 0u; fs->k = 0u; }
        return;
    }
    __shared__ uint32_t hist[SEL_BINS];
    __shared__ uint32_t warp_tot[8];
    __shared__ uint32_t s_pick[2];
    __shared__ int s_last;
    for (int k = threadIdx.x; k < SEL_BINS; k += 256) hist[k] = 0;
    uint32_t* gh0 = w.sel_hist + ((size_t)pair * FSEL_PASSES + 0) * SEL_BINS;
    uint32_t* gh = w.sel_hist + ((size_t)pair * FSEL_PASSES + 1) * SEL_BINS;
    uint32_t prefix, k1;
    hist_pick256(gh0, (uint32_t)(cutoff - 1), false, warp_tot, s_pick, prefix, k1);      // (ends with a barrier: hist[] is zero for all)
    const double* d2 = w.d2 + (size_t)pair * w.max_src;
    const int lane = threadIdx.x & 31;
#pragma unroll 4
    for (int r = 0; r < SEL_TILE / 256; r++) {
        const int i = base + r * 256 + threadIdx.x;
        bool in = false;
        uint32_t d = 0;
        if (i < n) {
            const uint32_t key = key32(d2[i]);
            in = (key >> 21) == prefix;
            d = (key >> 10) & 2047u;
        }
        const unsigned active = __ballot_sync(0xffffffffu, in);
        if (active == 0u) continue;
        if (in) {
            // warp-aggregated increment among the lanes that match
            const unsigned grp = __match_any_sync(active, d);
            if ((grp & ((1u << lane) - 1u)) == 0) atomicAdd(&hist[d], (uint32_t)__popc(grp));
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < SEL_BINS; k += 256) {
        const uint32_t c = hist[k];
        if (c) atomicAdd(gh + k, c);
    }
    // the last CTA of this pair to get here picks the bin that holds the rank
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t ctas = (uint32_t)((n + SEL_TILE - 1) / SEL_TILE);
        s_last = atomicAdd(w.fsel_done + pair, 1u) == ctas - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    uint32_t bin1, k2;
    hist_pick256(gh, k1, true, warp_tot, s_pick, bin1, k2);
    for (int k = threadIdx.x; k < SEL_BINS; k += 256) gh0[k] = 0;      // every CTA of the pair has read it: ready for the next nn launch
    if (threadIdx.x == 0) {
        fs->prefix = (prefix << 11) | bin1;
        fs->k = k2;
        fs->all = 0u;
        w.fsel_done[pair] = 0u;
        w.cand_n[pair] = 0u;      // (moments_kernel fills the list next)
    }
}

// One step of the descent, by a CTA of 1024 threads: find the bin of `gh` (SEL_BINS counters, global or shared; zeroed
// here for the next selection) that holds rank k and descend into it; after the last pass publish the threshold and,
// when exact ties straddle the cutoff, the largest source index kept among them.
__device__ __forceinline__ void sel_pick_step(IcpWorkspace& w, int pair, int pass, uint32_t* gh) {
    PairState* st = w.state + pair;
    SelState* sel = w.sel + pair;
    const int n = st->n_src, cutoff = st->cutoff;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t warp_cnt[32];
    __shared__ int s_idx_cut;
    // nothing to select: keep everything / nothing
    const bool trivial = cutoff >= n || cutoff <= 0;
    if (pass == 0) {
        if (threadIdx.x == 0) {
            sel->prefix = 0ull;
            sel->k = (uint32_t)max(cutoff - 1, 0);
            sel->finished = 0;
            if (trivial) { st->tau = (cutoff >= n) ? CUDART_INF : -1.0; st->idx_cut = 0x7fffffff; }
        }
        __syncthreads();
    }
    if (trivial) return;
    if (pass > 0 && sel->finished) return;      // (uniform: written before the last barrier of an earlier step)
    // 2 bins per thread, exclusive scan over the CTA
    const uint32_t c0 = gh[2 * threadIdx.x], c1 = gh[2 * threadIdx.x + 1];
    gh[2 * threadIdx.x] = 0;          // ready for the next pass / the next iteration's selection
    gh[2 * threadIdx.x + 1] = 0;
    const uint32_t tot = c0 + c1;
    uint32_t x = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
        const uint32_t t = warp_tot[lane];
        uint32_t ts = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, ts, o);
            if (lane >= o) ts += y;
        }
        warp_tot[lane] = ts - t;
    }
    __syncthreads();
    const uint32_t before = warp_tot[warp] + x - tot;
    const uint32_t k = sel->k;
    const unsigned long long prefix = sel->prefix;
    __syncthreads();
    if (k >= before && k < before + tot) {          // exactly one thread
        const bool first = k < before + c0;
        const uint32_t bin = 2 * threadIdx.x + (first ? 0 : 1);
        sel->prefix = (prefix << sel_bits(pass)) | (unsigned long long)bin;
        const uint32_t kk = k - before - (first ? 0u : c0), neq = first ? c0 : c1;
        sel->k = kk;
        sel->neq = neq;
        if (pass != SEL_PASSES - 1 && kk + 1 == neq) {
            // The rank falls on the LAST key of its bin: every key of the bin is kept, so the low bits of the threshold
            // do not matter -- keep everything up to the largest key the bin can hold and skip the remaining passes
            // (with 2 M distinct distances the bin of the third pass usually holds one key).
            const unsigned long long pre = (prefix << sel_bits(pass)) | (unsigned long long)bin;
            const int sh = sel_shift(pass);
            st->tau = __longlong_as_double((long long)((pre << sh) | ((1ull << sh) - 1ull)));
            st->idx_cut = 0x7fffffff;
            sel->finished = 1;
        }
    }
    __syncthreads();
    if (pass != SEL_PASSES - 1 || sel->finished) return;
    const unsigned long long tau_bits = sel->prefix;
    const uint32_t quota = sel->k + 1;     // how many elements equal to tau belong to the kept set
    const uint32_t neq = sel->neq;
    if (threadIdx.x == 0) s_idx_cut = 0x7fffffff;
    __syncthreads();
    if (quota < neq) {
        // exact ties straddle the cutoff: keep the lowest-index `quota` of them
        const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(w.d2 + (size_t)pair * w.max_src);
        uint32_t running = 0;
        for (int start = 0; start < n; start += 1024) {
            const int i = start + threadIdx.x;
            const bool eq = (i < n) && (keys[i] == tau_bits);
            const unsigned bal = __ballot_sync(0xffffffffu, eq);
            if (lane == 0) warp_cnt[warp] = __popc(bal);
            __syncthreads();
            uint32_t bef = running;
            for (int wv = 0; wv < warp; wv++) bef += warp_cnt[wv];
            const uint32_t incl = bef + __popc(bal & ((2u << lane) - 1u));
            if (eq && incl == quota) s_idx_cut = i;
            uint32_t t2 = 0;
            for (int wv = 0; wv < 32; wv++) t2 += warp_cnt[wv];
            running += t2;
            __syncthreads();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st->tau = __longlong_as_double((long long)tau_bits);
        st->idx_cut = s_idx_cut;
    }
}

// Pairs of at most NN_SMALL_PAIR source points are selected by sel_small_kernel below, the others by the
// multi-CTA histogram passes.
__device__ __forceinline__ bool sel_is_small(int n_src) { return n_src <= SEL_SMALL_MAX; }

// The whole selection of a small pair in one launch: one CTA, the six histogram + pick steps back to back on a
// shared-memory histogram (the keys, at most 256 KB, stay in L1/L2).  For the reference's own cloud sizes the twelve
// launches of the general path cost more than the search itself.
__global__ void __launch_bounds__(1024)
sel_small_kernel(IcpWorkspace w) {
    const int pair = blockIdx.x;
    const PairState* st = w.state + pair;
    const int n = st->n_src;
    if (st->done || !sel_is_small(n)) return;
    __shared__ uint32_t hist[SEL_BINS];
    for (int k = threadIdx.x; k < SEL_BINS; k += 1024) hist[k] = 0;
    const bool trivial = st->cutoff >= n || st->cutoff <= 0;
    const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(w.d2 + (size_t)pair * w.max_src);
    for (int pass = 0; pass < SEL_PASSES; pass++) {
        __syncthreads();
        if (!trivial) {
            const unsigned long long prefix = pass == 0 ? 0ull : w.sel[pair].prefix;
            const int shift = sel_shift(pass);
            const uint32_t mask = (1u << sel_bits(pass)) - 1u;
            for (int i = threadIdx.x; i < n; i += 1024) {
                const unsigned long long key = keys[i];
                if (pass == 0 || (key >> (shift + sel_bits(pass))) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & mask], 1u);
            }
        }
        __syncthreads();
        sel_pick_step(w, pair, pass, hist);
        if (trivial || w.sel[pair].finished) return;      // pass 0 has published "keep everything / nothing" / early end
    }
}

__device__ __forceinline__ bool kept(double d2, int i, double tau, int idx_cut) {
    return d2 < tau || (d2 == tau && i <= idx_cut);
}

// Moments over the kept correspondences.  A CTA takes MOM_TILE = NN_THREADS x MOM_PT consecutive queries, MOM_PT per thread
// (query base + r NN_THREADS + thread: coalesced), each thread summing its own in order before ONE warp reduction: a row of the
// pair's table per warp and tile, i.e. per 32 x MOM_PT queries (mom_rows() of them), a fixed order whatever the launch.  (One
// query per thread and a row per 32 queries made the reductions -- not the 76 bytes per correspondence -- the kernel's cost, and
// gave solve_kernel four times the rows to add.)  Small pairs: tau / idx_cut are final (sel_small_kernel).  Large pairs:
// everything below the bracket of the k-th smallest distance is kept and summed here; the bracket's members are listed for
// fsel_resolve_body; everything above is dropped.
constexpr int MOM_PT = 4;
constexpr int MOM_TILE = NN_THREADS * MOM_PT;
__host__ __device__ __forceinline__ int mom_rows(int n_src) { return ((n_src + MOM_TILE - 1) / MOM_TILE) * (NN_THREADS / 32); }

__global__ void __launch_bounds__(NN_THREADS)
moments_kernel(IcpWorkspace w) {
    const int pair = blockIdx.y;
    __shared__ GridParams g;
    __shared__ double T[12];
    __shared__ PairState s;
    __shared__ FSel fs;
    __shared__ double red[NN_THREADS / 32][16 * 33];
    if (threadIdx.x == 0) { g = w.grid.params[pair]; s = w.state[pair]; fs = w.fsel[pair]; }
    __syncthreads();
    if (threadIdx.x < 12) T[threadIdx.x] = s.T[threadIdx.x];
    __syncthreads();
    if (s.done || blockIdx.x * MOM_TILE >= s.n_src) return;
    const bool small = sel_is_small(s.n_src);
    const int lane = threadIdx.x & 31;
    double m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m[k] = 0.0;
#pragma unroll 1
    for (int r = 0; r < MOM_PT; r++) {
        const int i = blockIdx.x * MOM_TILE + r * NN_THREADS + threadIdx.x;
        bool member = false;
        if (i < s.n_src) {
            // the three streams indexed by i are requested together (one memory latency instead of two before the gather)
            const double d2 = w.d2[(size_t)pair * w.max_src + i];
            const int pos = w.nn_pos[(size_t)pair * w.max_src + i];
            const Point4 a = load_point(w.src_sorted + (size_t)pair * w.max_src + i);
            bool keep;
            if (small) {
                keep = kept(d2, i, s.tau, s.idx_cut);
            } else {
                const uint32_t k22 = key32(d2) >> 10;
                keep = fs.all || k22 < fs.prefix;
                member = !fs.all && k22 == fs.prefix;
            }
            if (keep) {
                const double qx = T[0] * a.x + T[1] * a.y + T[2] * a.z + T[3];
                const double qy = T[4] * a.x + T[5] * a.y + T[6] * a.z + T[7];
                const double qz = T[8] * a.x + T[9] * a.y + T[10] * a.z + T[11];
                const Point4 b = load_point(w.grid.sorted + (size_t)pair * w.grid.max_dst + pos);
                double t[16];
                accumulate_pair(t, qx - g.shift[0], qy - g.shift[1], qz - g.shift[2], b.x - g.shift[0], b.y - g.shift[1],
                                b.z - g.shift[2], sqrt(d2));
#pragma unroll
                for (int k = 0; k < 16; k++) m[k] += t[k];
            }
        }
        // members of the bracket: one slot each in the pair's list (warp-aggregated)
        const unsigned mm = __ballot_sync(0xffffffffu, member);
        if (mm) {
            uint32_t base = 0;
            if (lane == __ffs(mm) - 1) base = atomicAdd(w.cand_n + pair, (uint32_t)__popc(mm));
            base = __shfl_sync(0xffffffffu, base, __ffs(mm) - 1);
            if (member) w.cand_idx[(size_t)pair * w.max_src + base + __popc(mm & ((1u << lane) - 1u))] = i;
        }
    }
    warp_reduce16_store(m, red[threadIdx.x >> 5],
                        w.partial + ((size_t)pair * w.src_rows + (size_t)(blockIdx.x * (NN_THREADS / 32) + (threadIdx.x >> 5))) * 16);
}

// The bracket of a large pair (one CTA per pair): its c members ordered by (squared distance, position), the first k' kept.
// Publishes tau / idx_cut and writes the kept members' moments as row `mom_rows(n)` of the pair's table.  Up to
// FSEL_RANKED members (the normal case: a few dozen) every member's rank is counted directly and the kept ones are summed
// in rank order; a larger bracket (many equal distances) is resolved by bisection on the key bits and summed in position
// order by a sweep over all queries -- slower, deterministic either way.
constexpr int SOLVE_THREADS = 512;
constexpr int FSEL_RANKED = 2048;

__device__ __forceinline__ void fsel_resolve_body(IcpWorkspace& w, int pair) {
    PairState* st = w.state + pair;
    const int n = st->n_src;
    if (st->done) return;
    __shared__ double sm[SOLVE_THREADS];
    __shared__ double out16[16];
    __shared__ unsigned long long s_key[FSEL_RANKED];
    __shared__ int s_pos[FSEL_RANKED];
    __shared__ int s_byrank[FSEL_RANKED];
    __shared__ unsigned long long s_tau;
    __shared__ int s_cut;
    double* row = w.partial + ((size_t)pair * w.src_rows + (size_t)mom_rows(n)) * 16;
    const FSel fs = w.fsel[pair];
    if (sel_is_small(n)) return;      // (selected by sel_small_kernel; solve_kernel does not read a bracket row for them)
    if (fs.all) {
        if (threadIdx.x < 16) row[threadIdx.x] = 0.0;
        if (threadIdx.x == 0) { st->tau = CUDART_INF; st->idx_cut = 0x7fffffff; }
        return;
    }
    const int c = (int)w.cand_n[pair];
    const int kq = (int)fs.k;      // 0-based rank of the k-th smallest distance inside the bracket: members of rank <= kq are kept
    const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(w.d2 + (size_t)pair * w.max_src);
    const int32_t* cand = w.cand_idx + (size_t)pair * w.max_src;
    const GridParams& g = w.grid.params[pair];
    double m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m[k] = 0.0;
    auto add = [&](int i) {
        const double d2 = w.d2[(size_t)pair * w.max_src + i];
        const Point4 a = load_point(w.src_sorted + (size_t)pair * w.max_src + i);
        const Point4 b = load_point(w.grid.sorted + (size_t)pair * w.grid.max_dst + w.nn_pos[(size_t)pair * w.max_src + i]);
        const double qx = st->T[0] * a.x + st->T[1] * a.y + st->T[2] * a.z + st->T[3];
        const double qy = st->T[4] * a.x + st->T[5] * a.y + st->T[6] * a.z + st->T[7];
        const double qz = st->T[8] * a.x + st->T[9] * a.y + st->T[10] * a.z + st->T[11];
        double t[16];
        accumulate_pair(t, qx - g.shift[0], qy - g.shift[1], qz - g.shift[2], b.x - g.shift[0], b.y - g.shift[1], b.z - g.shift[2], sqrt(d2));
#pragma unroll
        for (int k = 0; k < 16; k++) m[k] += t[k];
    };
    if (c <= FSEL_RANKED) {
        for (int j = threadIdx.x; j < c; j += SOLVE_THREADS) {
            const int i = cand[j];
            s_pos[j] = i;
            s_key[j] = keys[i];      // non-negative doubles order as their bit patterns
        }
        __syncthreads();
        for (int j = threadIdx.x; j < c; j += SOLVE_THREADS) {
            const unsigned long long kj = s_key[j];
            const int pj = s_pos[j];
            int rank = 0;
            for (int u = 0; u < c; u++) rank += (s_key[u] < kj || (s_key[u] == kj && s_pos[u] < pj)) ? 1 : 0;
            s_byrank[rank] = j;
            if (rank == kq) { s_tau = kj; s_cut = pj; }
        }
        __syncthreads();
        // kept members in rank order: thread t takes ranks t, t + SOLVE_THREADS, ...
        for (int r = threadIdx.x; r <= kq && r < c; r += SOLVE_THREADS) add(s_pos[s_byrank[r]]);
    } else {
        // bisection on the 64 key bits, then on the 31 position bits among the keys equal to tau; counts by __syncthreads_count
        unsigned long long tau = 0ull;
        int need = kq;      // members with a key below the prefix under construction are already accounted for
        for (int bit = 62; bit >= 0; bit--) {      // (bit 63 is the sign: never set)
            const unsigned long long probe = tau | (1ull << bit), mask = ~((1ull << bit) - 1ull);
            int local = 0;
            for (int j = threadIdx.x; j < c; j += SOLVE_THREADS) {
                const unsigned long long kj = keys[cand[j]] & mask;
                local += (kj >= tau && kj < probe) ? 1 : 0;      // keys that have this bit clear under the current prefix
            }
            sm[threadIdx.x] = (double)local;
            __syncthreads();
            for (int o = SOLVE_THREADS / 2; o > 0; o >>= 1) {
                if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
                __syncthreads();
            }
            const int zeros = (int)sm[0];
            __syncthreads();
            if (need >= zeros) { need -= zeros; tau = probe; }
        }
        // `need` = rank among the members whose key equals tau, by position
        int cut = 0;
        for (int bit = 30; bit >= 0; bit--) {
            const int probe = cut | (1 << bit), mask = ~((1 << bit) - 1);
            int local = 0;
            for (int j = threadIdx.x; j < c; j += SOLVE_THREADS) {
                const int i = cand[j];
                local += (keys[i] == tau && (i & mask) >= cut && (i & mask) < probe) ? 1 : 0;
            }
            sm[threadIdx.x] = (double)local;
            __syncthreads();
            for (int o = SOLVE_THREADS / 2; o > 0; o >>= 1) {
                if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
                __syncthreads();
            }
            const int zeros = (int)sm[0];
            __syncthreads();
            if (need >= zeros) { need -= zeros; cut = probe; }
        }
        if (threadIdx.x == 0) { s_tau = tau; s_cut = cut; }
        __syncthreads();
        // kept members in position order: a sweep over all queries, thread t takes positions t, t + SOLVE_THREADS, ...
        const uint32_t prefix = fs.prefix;
        for (int i = threadIdx.x; i < n; i += SOLVE_THREADS) {
            const unsigned long long ki = keys[i];
            if ((key32(__longlong_as_double((long long)ki)) >> 10) == prefix && (ki < tau || (ki == tau && i <= cut))) add(i);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st->tau = __longlong_as_double((long long)s_tau);
        st->idx_cut = s_cut;
    }
    // fixed-order sum of the threads' partial moments
    for (int k = 0; k < 16; k++) {
        sm[threadIdx.x] = m[k];
        __syncthreads();
        for (int o = SOLVE_THREADS / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out16[k] = sm[0];
        __syncthreads();
    }
    if (threadIdx.x < 16) row[threadIdx.x] = out16[threadIdx.x];
}

// ---- solve ---------------------------------------------------------------------------------------------
// Sum `count` rows of 16 doubles in a fixed order with SOLVE_THREADS threads; result in out16 (shared).
__device__ __forceinline__ void sum_rows16(const double* __restrict__ rows, int count, double* sm /* [SOLVE_THREADS] */,
                                           double* out16) {
    constexpr int NGRP = SOLVE_THREADS / 16;
    const int q = threadIdx.x & 15, grp = threadIdx.x >> 4;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;   // four independent chains hide the load latency
    int r = grp;
    for (; r + 3 * NGRP < count; r += 4 * NGRP) {
        s0 += rows[(size_t)r * 16 + q];
        s1 += rows[(size_t)(r + NGRP) * 16 + q];
        s2 += rows[(size_t)(r + 2 * NGRP) * 16 + q];
        s3 += rows[(size_t)(r + 3 * NGRP) * 16 + q];
    }
    for (; r < count; r += NGRP) s0 += rows[(size_t)r * 16 + q];
    sm[grp * 16 + q] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
        for (int k = 0; k < NGRP; k++) t += sm[k * 16 + threadIdx.x];
        out16[threadIdx.x] = t;
    }
    __syncthreads();
}

// One ICP step of a pair from its table of moment rows (every thread of a CTA of SOLVE_THREADS): fixed-order sum, Kabsch,
// T <- T_iter T, convergence test (utils/icp.py:113-128).
__device__ __forceinline__ void solve_body(IcpWorkspace& w, int pair, double tolerance, int rows, double* sm, double* m) {
    PairState* st = w.state + pair;
    sum_rows16(w.partial + (size_t)pair * w.src_rows * 16, rows, sm, m);
    if (threadIdx.x == 0) {
        const GridParams& g = w.grid.params[pair];
        const double cnt = (double)st->cutoff;
        double Ti[16];
        kabsch_from_moments(m, cnt, g.shift, Ti);
        double Tn[12];
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 4; c++) {
                double acc = Ti[r * 4 + 0] * st->T[0 * 4 + c] + Ti[r * 4 + 1] * st->T[1 * 4 + c] + Ti[r * 4 + 2] * st->T[2 * 4 + c];
                if (c == 3) acc += Ti[r * 4 + 3];
                Tn[r * 4 + c] = acc;
            }
        }
        for (int k = 0; k < 12; k++) { st->Tp[k] = st->T[k]; st->T[k] = Tn[k]; }
        const double mean = m[15] / cnt;
        st->mean_err = mean;
        if (fabs(st->prev_err - mean) < tolerance) {
            st->done = 1;                       // break before iters += 1 (utils/icp.py:124-128)
        } else {
            st->prev_err = mean;
            st->iters += 1;
        }
    }
}

__global__ void __launch_bounds__(SOLVE_THREADS)
solve_kernel(IcpWorkspace w, double tolerance, int rows_fixed) {
    const int pair = blockIdx.x;
    PairState* st = w.state + pair;
    if (threadIdx.x == 0) w.vw_counter[pair] = 0;      // the nn launch of this iteration has finished
    if (st->done) return;
    __shared__ double sm[SOLVE_THREADS];
    __shared__ double m[16];
    const int n = st->n_src;
    // fused mode: one row per virtual warp of the persistent nn kernel; otherwise moments_kernel's rows
    const int rows = rows_fixed > 0 ? vwarps_for(n) : mom_rows(n) + (sel_is_small(n) ? 0 : 1);      // (+ 1: the bracket row of fsel_resolve_body, large pairs)
    solve_body(w, pair, tolerance, rows, sm, m);
}

// partial_set_size < 1, large pairs: the bracket's exact finish and the ICP step in one launch (both are one CTA per pair,
// the second reads the row the first has just written).
__global__ void __launch_bounds__(SOLVE_THREADS)
resolve_solve_kernel(IcpWorkspace w, double tolerance) {
    const int pair = blockIdx.x;
    fsel_resolve_body(w, pair);
    __threadfence();
    __syncthreads();
    PairState* st = w.state + pair;
    if (threadIdx.x == 0) w.vw_counter[pair] = 0;      // the nn launch of this iteration has finished
    if (st->done) return;
    __shared__ double sm[SOLVE_THREADS];
    __shared__ double m[16];
    const int n = st->n_src;
    const int rows = mom_rows(n) + (sel_is_small(n) ? 0 : 1);
    solve_body(w, pair, tolerance, rows, sm, m);
}

// ---- moments of A and the final fit -----------------------------------------------------------------------
__global__ void __launch_bounds__(NN_THREADS)
src_moments_kernel(IcpWorkspace w) {
    const int pair = blockIdx.y;
    __shared__ double red[NN_THREADS / 32][16 * 33];
    __shared__ double shift[3];
    __shared__ int s_n;
    if (threadIdx.x == 0) {
        s_n = w.state[pair].n_src;
        for (int k = 0; k < 3; k++) shift[k] = w.grid.params[pair].shift[k];
    }
    __syncthreads();
    if (blockIdx.x * NN_THREADS >= s_n) return;
    const int i = blockIdx.x * NN_THREADS + threadIdx.x;
    double m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m[k] = 0.0;
    if (i < s_n) {
        const Point4 a = load_point(w.src_sorted + (size_t)pair * w.max_src + i);
        const double x = a.x - shift[0], y = a.y - shift[1], z = a.z - shift[2];
        m[0] = x; m[1] = y; m[2] = z;
        m[3] = x * x; m[4] = x * y; m[5] = x * z; m[6] = y * y; m[7] = y * z; m[8] = z * z;
    }
    warp_reduce16_store(m, red[threadIdx.x >> 5],
                        w.partial + ((size_t)pair * w.src_rows + (size_t)(blockIdx.x * (NN_THREADS / 32) + (threadIdx.x >> 5))) * 16);
}

__global__ void __launch_bounds__(SOLVE_THREADS)
src_moments_reduce_kernel(IcpWorkspace w) {
    const int pair = blockIdx.x;
    __shared__ double sm[SOLVE_THREADS];
    __shared__ double m[16];
    const int n = w.state[pair].n_src;
    const int rows = (n + 31) / 32;
    sum_rows16(w.partial + (size_t)pair * w.src_rows * 16, rows, sm, m);
    if (threadIdx.x < 16) w.amom[(size_t)pair * 16 + threadIdx.x] = m[threadIdx.x];
}

__global__ void final_kernel(IcpWorkspace w, double* __restrict__ T_out, int32_t* __restrict__ iters_out,
                             double* __restrict__ mean_err_out) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= w.n_pairs) return;
    const PairState& st = w.state[pair];
    const double* am = w.amom + (size_t)pair * 16;
    const double* sh = w.grid.params[pair].shift;
    const double n = (double)max(st.n_src, 1);
    double ma[3] = {am[0] / n, am[1] / n, am[2] / n};
    const double raw[3][3] = {{am[3], am[4], am[5]}, {am[4], am[6], am[7]}, {am[5], am[7], am[8]}};
    double H[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double acc = 0.0;   // H = C_A * M^T
            for (int k = 0; k < 3; k++) acc += (raw[i][k] - n * ma[i] * ma[k]) * st.T[j * 4 + k];
            H[i * 3 + j] = acc;
        }
    double R[9];
    kabsch_rotation(H, R);
    double ca[3] = {ma[0] + sh[0], ma[1] + sh[1], ma[2] + sh[2]};   // centroid of A
    double* To = T_out + (size_t)pair * 16;
    for (int r = 0; r < 3; r++) {
        const double cs = st.T[r * 4 + 0] * ca[0] + st.T[r * 4 + 1] * ca[1] + st.T[r * 4 + 2] * ca[2] + st.T[r * 4 + 3];
        To[r * 4 + 0] = R[r * 3 + 0];
        To[r * 4 + 1] = R[r * 3 + 1];
        To[r * 4 + 2] = R[r * 3 + 2];
        To[r * 4 + 3] = cs - (R[r * 3 + 0] * ca[0] + R[r * 3 + 1] * ca[1] + R[r * 3 + 2] * ca[2]);
    }
    To[12] = 0; To[13] = 0; To[14] = 0; To[15] = 1;
    if (iters_out) iters_out[pair] = st.iters;
    if (mean_err_out) mean_err_out[pair] = st.mean_err;
}

__global__ void __launch_bounds__(256)
export_kernel(IcpWorkspace w, const int32_t* __restrict__ src_off, int32_t* __restrict__ out_idx, double* __restrict__ out_dist,
              uint8_t* __restrict__ out_keep, int sqrt_dist) {
    const int pair = blockIdx.y;
    const PairState& st = w.state[pair];
    const int n = st.n_src;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (w.grid.params[pair].n_pts <= 0 || st.iters == 0) {     // empty destination cloud / cutoff 0 (no search ran): nothing to report
        const size_t o = (size_t)src_off[pair] + i;
        if (out_idx) out_idx[o] = -1;
        if (out_dist) out_dist[o] = CUDART_INF;
        if (out_keep) out_keep[o] = 0;
        return;
    }
    // i runs over the SORTED source order; results go back to the caller's order
    const Point4 a = load_point(w.src_sorted + (size_t)pair * w.max_src + i);
    const size_t o = (size_t)src_off[pair] + a.idx;
    const double d2 = w.d2[(size_t)pair * w.max_src + i];
    if (out_idx) {
        const Point4 b = load_point(w.grid.sorted + (size_t)pair * w.grid.max_dst + w.nn_pos[(size_t)pair * w.max_src + i]);
        out_idx[o] = b.idx;
    }
    if (out_dist) out_dist[o] = sqrt_dist ? sqrt(d2) : d2;
    if (out_keep) out_keep[o] = kept(d2, i, st.tau, st.idx_cut) ? 1 : 0;
}

// ---- optional event timing of nn_kernel launches ----------------------------------------------------------
static std::mutex g_time_mutex;
static bool g_time_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_time_events;
static size_t g_time_used = 0;

struct NnTimer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t st;
    bool on;
    explicit NnTimer(cudaStream_t s) : st(s) {
        std::lock_guard<std::mutex> lock(g_time_mutex);
        on = g_time_on;
        if (!on) return;
        if (g_time_used == g_time_events.size()) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            g_time_events.emplace_back(a, b);
        }
        e0 = g_time_events[g_time_used].first;
        e1 = g_time_events[g_time_used].second;
        g_time_used++;
        cudaEventRecord(e0, st);
    }
    ~NnTimer() {
        if (on) cudaEventRecord(e1, st);
    }
};

// ---- host orchestration -----------------------------------------------------------------------------------------
static std::atomic<int> g_nn_impl{1};      // 1 = the search of lean_search.cuh (the only one left; the option is kept for ABI stability)
extern std::atomic<int> g_k4b_fp32;      // csrc/ransac.cu
static std::atomic<int> g_force_unfused{0}; // test knob: run partial_set_size == 1 through the select + moments kernels

static std::atomic<float> g_slack_mult{8.0f}, g_slack_max_frac{0.25f};
static std::atomic<int> g_slack_clamp{0};
static std::atomic<int> g_group_max{12};      // group_search for tasks with at most this many searching lanes (results do not depend on it)

template <int IMPL>
static int launch_query(IcpWorkspace& w, dim3 grid, bool seeded, bool fused, int write_d2, cudaStream_t st) {
    LeanTuning tune;
    tune.slack_mult = g_slack_mult.load();
    tune.slack_max_frac = g_slack_max_frac.load();
    tune.group_max = g_group_max.load();
    tune.slack_clamp = g_slack_clamp.load();
    const bool small = w.max_src <= NN_SMALL_PAIR;
#define NN_LAUNCH(S, F, M) R3D_LAUNCH((nn_query_kernel<IMPL, S, F, M>), grid, NN_THREADS, 0, st, w, write_d2, tune)
    if (seeded) {
        if (fused) { if (small) NN_LAUNCH(true, true, true); else NN_LAUNCH(true, true, false); }
        else { if (small) NN_LAUNCH(true, false, true); else NN_LAUNCH(true, false, false); }
    } else {
        if (fused) { if (small) NN_LAUNCH(false, true, true); else NN_LAUNCH(false, true, false); }
        else { if (small) NN_LAUNCH(false, false, true); else NN_LAUNCH(false, false, false); }
    }
#undef NN_LAUNCH
    return R3D_OK;
}

// Persistent query kernel: as many CTAs as are RESIDENT at once (148 SMs x the occupancy the kernel is compiled for)
// spread over the chunk's pairs -- a CTA that has to wait for a slot would leave its pair without workers while the
// others run -- and never more than one warp per task (32 queries; 4 for small pairs).
static int query_ctas_per_pair(const IcpWorkspace& w, int impl) {
    (void)impl;
    int per_pair = (148 * R3D_LEAN_MIN_CTAS) / w.n_pairs;
    const int per_cta = queries_per_task(w.max_src) * (NN_THREADS / 32);
    const int cap = (w.max_src + per_cta - 1) / per_cta;
    if (per_pair > cap) per_pair = cap;
    if (per_pair > vwarps_for(w.max_src) / (NN_THREADS / 32)) per_pair = vwarps_for(w.max_src) / (NN_THREADS / 32);
    return per_pair < 1 ? 1 : per_pair;
}

static int launch_nn(IcpWorkspace& w, bool seeded, bool fused, int write_d2, cudaStream_t st) {
    const int impl = g_nn_impl.load();
    const dim3 grid(query_ctas_per_pair(w, impl), w.n_pairs);
    return launch_query<1>(w, grid, seeded, fused, write_d2, st);
}

static int sort_source(IcpWorkspace& w, const Point4* src, const int32_t* src_off, cudaStream_t st) {
    const dim3 g256((w.max_src + 255) / 256, w.n_pairs);
    R3D_LAUNCH(src_keys_kernel, g256, 256, 0, st, w, src, src_off);
    int rc = radix_sort_pairs(w.skeys, w.svals, w.shist, w.src_desc, w.n_pairs, w.max_src, w.src_rtiles, st);
    if (rc != R3D_OK) return rc;
    R3D_LAUNCH(src_gather_kernel, g256, 256, 0, st, w, src, src_off);
    return R3D_OK;
}

// Query sharding over GPUs (r3d_icp_batch_sharded): every process holds both clouds and runs the same launches, but its
// nn kernel only takes the virtual warps vw = rank (mod world).  A virtual warp's row of moment sums is written by
// exactly one process and is zero everywhere else, so the SUM exchange of the row table reproduces the single-GPU table
// bit for bit (x + 0 + ... + 0 = x whatever the reduction order) and solve_kernel, which adds the rows in a fixed
// order on every process, produces the single-GPU transform on all of them.
struct IcpShard {
    int rank, world;
    r3d_exchange_fn exchange;
    void* ctx;
};

int icp_run(IcpWorkspace& w, const Point4* src, const int32_t* src_off, const Point4* dst, const int32_t* dst_off,
            const double* init_pose, const r3d_icp_params& prm, double* T_out, int32_t* iters_out, double* mean_err_out,
            int32_t* last_idx, double* last_dist, uint8_t* last_keep, cudaStream_t st, const IcpShard* shard = nullptr) {
    const int P = w.n_pairs;
    const bool sharded = shard != nullptr && shard->world > 1;
    if (sharded) {
        w.shard_rank = shard->rank;
        w.shard_world = shard->world;
    }
    const size_t shard_doubles = (size_t)vwarps_for(w.max_src) * 16;      // one pair (checked by the caller): rows [0, nvw)
    int rc = grid_build(w.grid, dst, dst_off, prm.cell_size, st);
    if (rc != R3D_OK) return rc;
    R3D_LAUNCH(state_init_kernel, (P + 127) / 128, 128, 0, st, w, src_off, init_pose, prm.partial_set_size);
    const dim3 qgrid(w.src_tiles, P);
    rc = sort_source(w, src, src_off, st);
    if (rc != R3D_OK) return rc;
    R3D_LAUNCH(src_moments_kernel, qgrid, NN_THREADS, 0, st, w);
    R3D_LAUNCH(src_moments_reduce_kernel, P, SOLVE_THREADS, 0, st, w);
    const bool fused = prm.partial_set_size >= 1.0 && !g_force_unfused.load();
    const int want_d2 = (last_dist != nullptr || last_keep != nullptr) ? 1 : 0;
    if (!fused) {
        R3D_CUDA(cudaMemsetAsync(w.sel_hist, 0, (size_t)P * FSEL_PASSES * SEL_BINS * sizeof(uint32_t), st));
        R3D_CUDA(cudaMemsetAsync(w.fsel_done, 0, (size_t)P * FSEL_PASSES * sizeof(uint32_t), st));
        R3D_CUDA(cudaMemsetAsync(w.fsel, 0, (size_t)P * sizeof(FSel), st));
    }
    for (int it = 0; it < prm.max_iterations; it++) {
        if (sharded) R3D_CUDA(cudaMemsetAsync(w.partial, 0, shard_doubles * sizeof(double), st));
        {
            NnTimer timer(st);
            rc = launch_nn(w, it > 0, fused, fused ? want_d2 : 2, st);      // (2: distances are written AND their first selection histogram counted)
            if (rc != R3D_OK) return rc;
        }
        if (sharded) {
            // the caller's collective (NCCL all-reduce through torch.distributed in the shipped binding), ordered on `st`
            if (shard->exchange(shard->ctx, w.partial, shard_doubles, (void*)st) != 0) return R3D_ERR_INVALID;
        }
        if (!fused) {
            if (SEL_SMALL_MAX > 0) R3D_LAUNCH(sel_small_kernel, P, 1024, 0, st, w);
            if (w.max_src > SEL_SMALL_MAX) R3D_LAUNCH(fsel_hist_kernel, dim3((w.max_src + SEL_TILE - 1) / SEL_TILE, P), 256, 0, st, w);
            R3D_LAUNCH(moments_kernel, dim3((w.max_src + MOM_TILE - 1) / MOM_TILE, P), NN_THREADS, 0, st, w);
        }
        if (!fused && w.max_src > SEL_SMALL_MAX) R3D_LAUNCH(resolve_solve_kernel, P, SOLVE_THREADS, 0, st, w, prm.tolerance);
        else R3D_LAUNCH(solve_kernel, P, SOLVE_THREADS, 0, st, w, prm.tolerance, fused ? 1 : 0);
    }
    R3D_LAUNCH(final_kernel, (P + 63) / 64, 64, 0, st, w, T_out, iters_out, mean_err_out);
    if ((last_idx || last_dist || last_keep) && prm.max_iterations > 0) {
        R3D_LAUNCH(export_kernel, dim3((w.max_src + 255) / 256, P), 256, 0, st, w, src_off, last_idx, last_dist, last_keep, 1);
    }
    return R3D_OK;
}

size_t icp_workspace_bytes_internal(int n_pairs, int max_src, int max_dst) {
    IcpWorkspace w;
    Arena a(nullptr, 0);
    icp_workspace_layout(w, a, n_pairs, max_src, max_dst);
    return a.used + 512;
}

int icp_workspace_bind(IcpWorkspace& w, void* ws, size_t bytes, int n_pairs, int max_src, int max_dst) {
    Arena a(ws, bytes);
    icp_workspace_layout(w, a, n_pairs, max_src, max_dst);
    return a.ok ? R3D_OK : R3D_ERR_WORKSPACE;
}

// nearest_neighbor(src, dst) for a batch of pairs: grid build + one unseeded exact query
int nn_run(IcpWorkspace& w, const Point4* src, const int32_t* src_off, const Point4* dst, const int32_t* dst_off,
           double cell_size, int32_t* out_idx, double* out_dist, cudaStream_t st) {
    const int P = w.n_pairs;
    int rc = grid_build(w.grid, dst, dst_off, cell_size, st);
    if (rc != R3D_OK) return rc;
    R3D_LAUNCH(state_init_kernel, (P + 127) / 128, 128, 0, st, w, src_off, (const double*)nullptr, 1.0);
    rc = sort_source(w, src, src_off, st);
    if (rc != R3D_OK) return rc;
    {
        NnTimer timer(st);
        rc = launch_nn(w, false, false, 1, st);
        if (rc != R3D_OK) return rc;
    }
    R3D_LAUNCH(export_kernel, dim3((w.max_src + 255) / 256, P), 256, 0, st, w, src_off, out_idx, out_dist, (uint8_t*)nullptr, 1);
    return R3D_OK;
}

}  // namespace r3d

using namespace r3d;

extern "C" size_t r3d_icp_workspace_bytes(int n_pairs, int max_src, int max_dst) {
    if (n_pairs <= 0 || max_src <= 0 || max_dst <= 0) return 0;
    return icp_workspace_bytes_internal(n_pairs, max_src, max_dst);
}

extern "C" size_t r3d_nn_workspace_bytes(int n_pairs, int max_src, int max_dst) {
    return r3d_icp_workspace_bytes(n_pairs, max_src, max_dst);
}

extern "C" int r3d_icp_batch(const double* src, const int32_t* src_off, const double* dst, const int32_t* dst_off,
                             int n_pairs, int max_src, int max_dst, const double* init_pose, const r3d_icp_params* params,
                             double* T_out, int32_t* iters_out, double* mean_err_out, int32_t* last_idx, double* last_dist,
                             uint8_t* last_keep, void* workspace, size_t workspace_bytes, void* stream) {
    if (!src || !src_off || !dst || !dst_off || !params || !T_out || n_pairs <= 0 || max_src <= 0 || max_dst <= 0)
        return R3D_ERR_INVALID;
    if (n_pairs > 65535 || params->max_iterations < 0) return R3D_ERR_INVALID;
    if (!(params->partial_set_size > 0.0) || params->partial_set_size > 1.0) return R3D_ERR_INVALID;
    if ((((uintptr_t)src) & 31) || (((uintptr_t)dst) & 31)) return R3D_ERR_INVALID;
    IcpWorkspace w;
    int rc = icp_workspace_bind(w, workspace, workspace_bytes, n_pairs, max_src, max_dst);
    if (rc != R3D_OK) return rc;
    return icp_run(w, (const Point4*)src, src_off, (const Point4*)dst, dst_off, init_pose, *params, T_out, iters_out,
                   mean_err_out, last_idx, last_dist, last_keep, (cudaStream_t)stream);
}

extern "C" int r3d_icp_batch_sharded(const double* src, const int32_t* src_off, const double* dst, const int32_t* dst_off,
                                     int n_pairs, int max_src, int max_dst, const double* init_pose, const r3d_icp_params* params,
                                     int rank, int world, r3d_exchange_fn exchange, void* exchange_ctx, double* T_out,
                                     int32_t* iters_out, double* mean_err_out, void* workspace, size_t workspace_bytes,
                                     void* stream) {
    if (!src || !src_off || !dst || !dst_off || !params || !T_out || max_src <= 0 || max_dst <= 0) return R3D_ERR_INVALID;
    if (n_pairs != 1) return R3D_ERR_INVALID;                       // one (large) pair; batches shard by pair instead
    if (world < 1 || rank < 0 || rank >= world || world > NN_VWARPS || (world > 1 && !exchange)) return R3D_ERR_INVALID;
    if (params->max_iterations < 0 || params->partial_set_size != 1.0) return R3D_ERR_INVALID;      // see include/r3d_b200.h
    if ((((uintptr_t)src) & 31) || (((uintptr_t)dst) & 31)) return R3D_ERR_INVALID;
    IcpWorkspace w;
    int rc = icp_workspace_bind(w, workspace, workspace_bytes, n_pairs, max_src, max_dst);
    if (rc != R3D_OK) return rc;
    IcpShard shard{rank, world, exchange, exchange_ctx};
    return icp_run(w, (const Point4*)src, src_off, (const Point4*)dst, dst_off, init_pose, *params, T_out, iters_out,
                   mean_err_out, nullptr, nullptr, nullptr, (cudaStream_t)stream, &shard);
}

extern "C" int r3d_nn_batch(const double* src, const int32_t* src_off, const double* dst, const int32_t* dst_off, int n_pairs,
                            int max_src, int max_dst, double cell_size, int32_t* out_idx, double* out_dist, void* workspace,
                            size_t workspace_bytes, void* stream) {
    if (!src || !src_off || !dst || !dst_off || !out_idx || !out_dist || n_pairs <= 0 || max_src <= 0 || max_dst <= 0)
        return R3D_ERR_INVALID;
    if (n_pairs > 65535) return R3D_ERR_INVALID;
    if ((((uintptr_t)src) & 31) || (((uintptr_t)dst) & 31)) return R3D_ERR_INVALID;
    IcpWorkspace w;
    int rc = icp_workspace_bind(w, workspace, workspace_bytes, n_pairs, max_src, max_dst);
    if (rc != R3D_OK) return rc;
    return nn_run(w, (const Point4*)src, src_off, (const Point4*)dst, dst_off, cell_size, out_idx, out_dist, (cudaStream_t)stream);
}

extern "C" void r3d_nn_timing_enable(int on) {
    std::lock_guard<std::mutex> lock(g_time_mutex);
    g_time_on = on != 0;
}

extern "C" int r3d_nn_timing_read(double* total_ms, int64_t* launches, int64_t* queries, int reset) {
    std::lock_guard<std::mutex> lock(g_time_mutex);
    double ms = 0.0;
    for (size_t k = 0; k < g_time_used; k++) {
        cudaError_t e = cudaEventSynchronize(g_time_events[k].second);
        if (e != cudaSuccess) { set_cuda_error(e, "nn timing"); return R3D_ERR_CUDA; }
        float t = 0.f;
        cudaEventElapsedTime(&t, g_time_events[k].first, g_time_events[k].second);
        ms += t;
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = (int64_t)g_time_used;
    if (queries) *queries = 0;
    if (reset) g_time_used = 0;
    return R3D_OK;
}

static int nn_stats_read(uint64_t* out_host, int n, int reset) {
#ifdef R3D_NN_STATS
    if (out_host) {
        cudaError_t e = cudaMemcpyFromSymbol(out_host, g_nn_stats, (size_t)n * sizeof(unsigned long long));
        if (e != cudaSuccess) { set_cuda_error(e, "nn stats"); return R3D_ERR_CUDA; }
    }
    if (reset) {
        const unsigned long long z[32] = {0};
        cudaError_t e = cudaMemcpyToSymbol(g_nn_stats, z, sizeof(z));
        if (e != cudaSuccess) { set_cuda_error(e, "nn stats"); return R3D_ERR_CUDA; }
    }
    return R3D_OK;
#else
    (void)out_host; (void)n; (void)reset;
    return R3D_ERR_INVALID;      // the library was built without -DR3D_NN_STATS
#endif
}

extern "C" int r3d_nn_stats_read(uint64_t* out8_host, int reset) { return nn_stats_read(out8_host, 8, reset); }
extern "C" int r3d_nn_stats_read_ext(uint64_t* out32_host, int reset) { return nn_stats_read(out32_host, 32, reset); }

extern "C" int r3d_set_option(int option, double value) {
    switch (option) {
        case R3D_OPT_NN_IMPL: {
            const int impl = (int)value;
            if (impl != 1) return R3D_ERR_INVALID;
            g_nn_impl.store(impl);
            return R3D_OK;
        }
        case R3D_OPT_GROUP_SEARCH_MAX:
            if (!(value >= 0.0 && value <= 32.0)) return R3D_ERR_INVALID;
            g_group_max.store((int)value);
            return R3D_OK;
        case R3D_OPT_SLACK_CLAMP:
            g_slack_clamp.store(value != 0.0 ? 1 : 0);
            return R3D_OK;
        case R3D_OPT_K4B_FP32:
            g_k4b_fp32.store(value != 0.0 ? 1 : 0);
            return R3D_OK;
        case R3D_OPT_FORCE_UNFUSED:
            g_force_unfused.store(value != 0.0 ? 1 : 0);
            return R3D_OK;
        case R3D_OPT_SLACK_MULT:
            if (!(value >= 0.0 && value <= 1.0e6)) return R3D_ERR_INVALID;
            g_slack_mult.store((float)value);
            return R3D_OK;
        case R3D_OPT_SLACK_MAX_FRAC:
            if (!(value >= 0.0 && value <= 64.0)) return R3D_ERR_INVALID;
            g_slack_max_frac.store((float)value);
            return R3D_OK;
        case R3D_OPT_CELL_FACTOR:
            if (!(value > 0.05 && value < 100.0)) return R3D_ERR_INVALID;
            g_cell_factor.store(value);
            return R3D_OK;
        default:
            return R3D_ERR_INVALID;
    }
}
