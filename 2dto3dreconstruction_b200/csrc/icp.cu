// K2 query + K3: the ICP loop of utils/icp.py:74-133, batched over pairs, entirely on the device.
//
// Per iteration (no host round trip; converged pairs turn into early-exit CTAs):
//   nn_kernel        q = T_cum * a (fp64), exact NN on the sparse-block grid.  Iterations >= 2
//                    seed the upper bound with the previous neighbour (the cloud moves little
//                    between iterations, so the search box is ~1-2 cells wide).  With
//                    partial_set_size == 1 the 16 Kabsch moments are reduced in the same kernel.
//   select_kernel    (partial < 1) radix-select of the cutoff-th smallest squared distance
//                    = np.argpartition(distances, cutoff-1)[:cutoff], utils/icp.py:109-110
//   moments_kernel   (partial < 1) moments over the kept correspondences
//   solve_kernel     fixed-order sum of the per-CTA partials, 3x3 Kabsch (utils/icp.py:22-55),
//                    T_cum <- T_iter * T_cum, mean error, |prev - mean| < tol test (icp.py:123-126)
// After the loop, final_kernel reproduces utils/icp.py:131 best_fit_transform(A, src) from the
// second moments of A (src is the affine image T_cum * A, so H = C_A * M^T in closed form).
#include "grid.cuh"
#include <mutex>
#include <vector>

namespace r3d {

constexpr int NN_THREADS = 128;

struct PairState {
    double T[12];        // cumulative transform, rows 0..2 of the 4x4
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

struct IcpWorkspace {
    GridWorkspace grid;
    int n_pairs, max_src, src_tiles;
    PairState* state;      // [pair]
    int32_t* nn_pos;       // [pair][max_src]  sorted position of the current neighbour
    double* d2;            // [pair][max_src]
    double* partial;       // [pair][src_tiles][16]
    double* amom;          // [pair][16]  moments of A
};

static size_t icp_workspace_layout(IcpWorkspace& w, Arena& a, int n_pairs, int max_src, int max_dst) {
    grid_workspace_layout(w.grid, a, n_pairs, max_dst);
    w.n_pairs = n_pairs;
    w.max_src = max_src;
    w.src_tiles = (max_src + NN_THREADS - 1) / NN_THREADS;
    const size_t P = (size_t)n_pairs;
    w.state = a.take<PairState>(P);
    w.nn_pos = a.take<int32_t>(P * max_src);
    w.d2 = a.take<double>(P * max_src);
    w.partial = a.take<double>(P * w.src_tiles * 16);
    w.amom = a.take<double>(P * 16);
    return a.used;
}

// ---- state -------------------------------------------------------------------------------------
__global__ void state_init_kernel(IcpWorkspace w, const int32_t* __restrict__ src_off, const double* __restrict__ init_pose,
                                  double partial_set_size) {
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= w.n_pairs) return;
    PairState s;
    for (int k = 0; k < 12; k++) s.T[k] = init_pose ? init_pose[(size_t)pair * 16 + k] : ((k % 5 == 0) ? 1.0 : 0.0);
    s.prev_err = 0.0;
    s.mean_err = 0.0;
    s.tau = 0.0;
    s.idx_cut = 0x7fffffff;
    s.iters = 1;
    s.done = 0;
    s.n_src = src_off[pair + 1] - src_off[pair];
    s.cutoff = (int)((double)s.n_src * partial_set_size);   // int(N * p), utils/icp.py:109
    s.pad = 0;
    if (s.n_src <= 0 || w.grid.params[pair].n_pts <= 0) s.done = 1;
    w.state[pair] = s;
}

// ---- block reduction of 16 doubles per thread ----------------------------------------------------
__device__ __forceinline__ void reduce16_store(const double (&m)[16], double* red /* [16][NN_THREADS] */,
                                               double* __restrict__ out16) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int k = 0; k < 16; k++) red[k * NN_THREADS + t] = m[k];
    __syncthreads();
    constexpr int NW = NN_THREADS / 32;
#pragma unroll
    for (int j = 0; j < 16 / NW; j++) {
        const int k = warp * (16 / NW) + j;
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < NW; c++) s += red[k * NN_THREADS + c * 32 + lane];
        s = warp_sum(s);
        if (lane == 0) out16[k] = s;
    }
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

// ---- nearest neighbour ----------------------------------------------------------------------------
template <bool SEEDED, bool FUSED>
__global__ void __launch_bounds__(NN_THREADS)
nn_kernel(IcpWorkspace w, const Point4* __restrict__ src, const int32_t* __restrict__ src_off, int write_d2) {
    const int pair = blockIdx.y;
    __shared__ GridParams g;
    __shared__ double T[12];
    __shared__ int s_n, s_done;
    __shared__ double red[FUSED ? 16 * NN_THREADS : 1];
    if (threadIdx.x == 0) {
        g = w.grid.params[pair];
        s_n = w.state[pair].n_src;
        s_done = w.state[pair].done;
    }
    if (threadIdx.x < 12) T[threadIdx.x] = w.state[pair].T[threadIdx.x];
    __syncthreads();
    if (s_done || blockIdx.x * NN_THREADS >= s_n) return;

    const GridView v = grid_view(w.grid, pair);
    const int i = blockIdx.x * NN_THREADS + threadIdx.x;
    const bool active = i < s_n;
    double m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m[k] = 0.0;

    if (active) {
        const Point4 a = load_point(src + src_off[pair] + i);
        const double qx = T[0] * a.x + T[1] * a.y + T[2] * a.z + T[3];
        const double qy = T[4] * a.x + T[5] * a.y + T[6] * a.z + T[7];
        const double qz = T[8] * a.x + T[9] * a.y + T[10] * a.z + T[11];
        Best best;
        best.d2 = CUDART_INF;
        best.idx = 0x7fffffff;
        best.pos = 0;
        int32_t* my_pos = w.nn_pos + (size_t)pair * w.max_src + i;
        if (SEEDED) {
            consider(best, v.pts, *my_pos, qx, qy, qz);
        } else {
            seed_search(best, v, g, qx, qy, qz);
        }
        bounded_search(best, v, g, qx, qy, qz);
        *my_pos = best.pos;
        if (!FUSED || write_d2) w.d2[(size_t)pair * w.max_src + i] = best.d2;
        if (FUSED) {
            const Point4 b = load_point(v.pts + best.pos);
            accumulate_pair(m, qx - g.shift[0], qy - g.shift[1], qz - g.shift[2], b.x - g.shift[0], b.y - g.shift[1],
                            b.z - g.shift[2], sqrt(best.d2));
        }
    }
    if (FUSED) reduce16_store(m, red, w.partial + ((size_t)pair * w.src_tiles + blockIdx.x) * 16);
}

// ---- partial-set selection ---------------------------------------------------------------------------
// k-th smallest of the non-negative doubles d2[0..n) by MSB-first 8-bit radix select on their bit
// patterns (order preserving for non-negative IEEE doubles).  One CTA per pair.
__global__ void __launch_bounds__(1024)
select_kernel(IcpWorkspace w) {
    const int pair = blockIdx.x;
    PairState* st = w.state + pair;
    if (st->done) return;
    const int n = st->n_src, cutoff = st->cutoff;
    if (cutoff >= n || cutoff <= 0) {
        if (threadIdx.x == 0) { st->tau = (cutoff >= n) ? CUDART_INF : -1.0; st->idx_cut = 0x7fffffff; }
        return;
    }
    const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(w.d2 + (size_t)pair * w.max_src);
    __shared__ uint32_t hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ uint32_t s_k, s_neq;
    __shared__ uint32_t warp_cnt[32];
    __shared__ int s_idx_cut;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_prefix = 0ull; s_k = (uint32_t)(cutoff - 1); }
    const int n_round = (n + 1023) / 1024 * 1024;
    for (int pass = 7; pass >= 0; pass--) {
        const int shift = pass * 8;
        if (threadIdx.x < 256) hist[threadIdx.x] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        for (int i = threadIdx.x; i < n_round; i += 1024) {
            bool in = false;
            uint32_t d = 0;
            if (i < n) {
                const unsigned long long key = keys[i];
                in = (pass == 7) || ((key >> (shift + 8)) == prefix);
                d = (uint32_t)(key >> shift) & 255u;
            }
            const unsigned grp = __match_any_sync(0xffffffffu, in ? d : (256u + (uint32_t)lane));
            if (in && (grp & ((1u << lane) - 1u)) == 0) atomicAdd(&hist[d], (uint32_t)__popc(grp));
        }
        __syncthreads();
        if (warp == 0) {
            // 8 bins per lane; find the bin containing rank k
            uint32_t c[8], tot = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) { c[j] = hist[lane * 8 + j]; tot += c[j]; }
            uint32_t x = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            uint32_t before = x - tot;
            const uint32_t k = s_k;
            if (k >= before && k < before + tot) {
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (k >= before && k < before + c[j]) {
                        s_prefix = (prefix << 8) | (unsigned long long)(lane * 8 + j);
                        s_k = k - before;
                        s_neq = c[j];
                    }
                    before += c[j];
                }
            }
        }
        __syncthreads();
    }
    const unsigned long long tau_bits = s_prefix;
    const uint32_t quota = s_k + 1;     // how many elements equal to tau belong to the kept set
    if (threadIdx.x == 0) s_idx_cut = 0x7fffffff;
    __syncthreads();
    if (quota < s_neq) {
        // exact ties straddle the cutoff: keep the lowest-index `quota` of them
        uint32_t running = 0;
        for (int start = 0; start < n; start += 1024) {
            const int i = start + threadIdx.x;
            const bool eq = (i < n) && (keys[i] == tau_bits);
            const unsigned bal = __ballot_sync(0xffffffffu, eq);
            if (lane == 0) warp_cnt[warp] = __popc(bal);
            __syncthreads();
            uint32_t before = running;
            for (int wv = 0; wv < warp; wv++) before += warp_cnt[wv];
            const uint32_t incl = before + __popc(bal & ((2u << lane) - 1u));
            if (eq && incl == quota) s_idx_cut = i;
            uint32_t tot = 0;
            for (int wv = 0; wv < 32; wv++) tot += warp_cnt[wv];
            running += tot;
            __syncthreads();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st->tau = __longlong_as_double((long long)tau_bits);
        st->idx_cut = s_idx_cut;
    }
}

__device__ __forceinline__ bool kept(double d2, int i, double tau, int idx_cut) {
    return d2 < tau || (d2 == tau && i <= idx_cut);
}

__global__ void __launch_bounds__(NN_THREADS)
moments_kernel(IcpWorkspace w, const Point4* __restrict__ src, const int32_t* __restrict__ src_off) {
    const int pair = blockIdx.y;
    __shared__ GridParams g;
    __shared__ double T[12];
    __shared__ PairState s;
    __shared__ double red[16 * NN_THREADS];
    if (threadIdx.x == 0) { g = w.grid.params[pair]; s = w.state[pair]; }
    __syncthreads();
    if (threadIdx.x < 12) T[threadIdx.x] = s.T[threadIdx.x];
    __syncthreads();
    if (s.done || blockIdx.x * NN_THREADS >= s.n_src) return;
    const int i = blockIdx.x * NN_THREADS + threadIdx.x;
    double m[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m[k] = 0.0;
    if (i < s.n_src) {
        const double d2 = w.d2[(size_t)pair * w.max_src + i];
        if (kept(d2, i, s.tau, s.idx_cut)) {
            const Point4 a = load_point(src + src_off[pair] + i);
            const double qx = T[0] * a.x + T[1] * a.y + T[2] * a.z + T[3];
            const double qy = T[4] * a.x + T[5] * a.y + T[6] * a.z + T[7];
            const double qz = T[8] * a.x + T[9] * a.y + T[10] * a.z + T[11];
            const Point4 b = load_point(w.grid.sorted + (size_t)pair * w.grid.max_dst + w.nn_pos[(size_t)pair * w.max_src + i]);
            accumulate_pair(m, qx - g.shift[0], qy - g.shift[1], qz - g.shift[2], b.x - g.shift[0], b.y - g.shift[1],
                            b.z - g.shift[2], sqrt(d2));
        }
    }
    reduce16_store(m, red, w.partial + ((size_t)pair * w.src_tiles + blockIdx.x) * 16);
}

// ---- solve ---------------------------------------------------------------------------------------------
// Sum `count` rows of 16 doubles in a fixed order with 128 threads; result in out16 (shared).
__device__ __forceinline__ void sum_rows16(const double* __restrict__ rows, int count, double* sm /* [8][16] */, double* out16) {
    const int q = threadIdx.x & 15, grp = threadIdx.x >> 4;   // 8 groups
    double s = 0.0;
    for (int r = grp; r < count; r += 8) s += rows[(size_t)r * 16 + q];
    sm[grp * 16 + q] = s;
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) t += sm[k * 16 + threadIdx.x];
        out16[threadIdx.x] = t;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(128)
solve_kernel(IcpWorkspace w, double tolerance) {
    const int pair = blockIdx.x;
    PairState* st = w.state + pair;
    if (st->done) return;
    __shared__ double sm[8 * 16];
    __shared__ double m[16];
    const int n = st->n_src;
    const int tiles = (n + NN_THREADS - 1) / NN_THREADS;
    sum_rows16(w.partial + (size_t)pair * w.src_tiles * 16, tiles, sm, m);
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
        for (int k = 0; k < 12; k++) st->T[k] = Tn[k];
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

// ---- moments of A and the final fit -----------------------------------------------------------------------
__global__ void __launch_bounds__(NN_THREADS)
src_moments_kernel(IcpWorkspace w, const Point4* __restrict__ src, const int32_t* __restrict__ src_off) {
    const int pair = blockIdx.y;
    __shared__ double red[16 * NN_THREADS];
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
        const Point4 a = load_point(src + src_off[pair] + i);
        const double x = a.x - shift[0], y = a.y - shift[1], z = a.z - shift[2];
        m[0] = x; m[1] = y; m[2] = z;
        m[3] = x * x; m[4] = x * y; m[5] = x * z; m[6] = y * y; m[7] = y * z; m[8] = z * z;
    }
    reduce16_store(m, red, w.partial + ((size_t)pair * w.src_tiles + blockIdx.x) * 16);
}

__global__ void __launch_bounds__(128)
src_moments_reduce_kernel(IcpWorkspace w) {
    const int pair = blockIdx.x;
    __shared__ double sm[8 * 16];
    __shared__ double m[16];
    const int n = w.state[pair].n_src;
    const int tiles = (n + NN_THREADS - 1) / NN_THREADS;
    sum_rows16(w.partial + (size_t)pair * w.src_tiles * 16, tiles, sm, m);
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
    const size_t o = (size_t)src_off[pair] + i;
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
int icp_run(IcpWorkspace& w, const Point4* src, const int32_t* src_off, const Point4* dst, const int32_t* dst_off,
            const double* init_pose, const r3d_icp_params& prm, double* T_out, int32_t* iters_out, double* mean_err_out,
            int32_t* last_idx, double* last_dist, uint8_t* last_keep, cudaStream_t st) {
    const int P = w.n_pairs;
    int rc = grid_build(w.grid, dst, dst_off, prm.cell_size, st);
    if (rc != R3D_OK) return rc;
    R3D_LAUNCH(state_init_kernel, (P + 127) / 128, 128, 0, st, w, src_off, init_pose, prm.partial_set_size);
    const dim3 qgrid(w.src_tiles, P);
    R3D_LAUNCH(src_moments_kernel, qgrid, NN_THREADS, 0, st, w, src, src_off);
    R3D_LAUNCH(src_moments_reduce_kernel, P, 128, 0, st, w);
    const bool fused = prm.partial_set_size >= 1.0;
    const int want_d2 = (last_dist != nullptr || last_keep != nullptr) ? 1 : 0;
    for (int it = 0; it < prm.max_iterations; it++) {
        {
            NnTimer timer(st);
            if (fused) {
                if (it == 0) R3D_LAUNCH((nn_kernel<false, true>), qgrid, NN_THREADS, 0, st, w, src, src_off, want_d2);
                else R3D_LAUNCH((nn_kernel<true, true>), qgrid, NN_THREADS, 0, st, w, src, src_off, want_d2);
            } else {
                if (it == 0) R3D_LAUNCH((nn_kernel<false, false>), qgrid, NN_THREADS, 0, st, w, src, src_off, 1);
                else R3D_LAUNCH((nn_kernel<true, false>), qgrid, NN_THREADS, 0, st, w, src, src_off, 1);
            }
        }
        if (!fused) {
            R3D_LAUNCH(select_kernel, P, 1024, 0, st, w);
            R3D_LAUNCH(moments_kernel, qgrid, NN_THREADS, 0, st, w, src, src_off);
        }
        R3D_LAUNCH(solve_kernel, P, 128, 0, st, w, prm.tolerance);
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
    {
        NnTimer timer(st);
        R3D_LAUNCH((nn_kernel<false, false>), dim3(w.src_tiles, P), NN_THREADS, 0, st, w, src, src_off, 1);
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
