// Fused registration of batches of depth-frame pairs: K1 (both frames) -> grid build -> ICP.
// This is the batched form of the reference's pair loop body (reconstruct.py:319-324 with the
// clouds of reconstruct_rgbd.py:48): pairs are independent, so a chunk of them shares every launch.
// Chunks alternate between `n_streams` CUDA streams, each with its own workspace slice, so the
// tiny per-pair kernels (solve, scans) of one chunk overlap the wide kernels of another and, in the
// host variant, the H2D copy of the next chunk overlaps compute.
#include "grid.cuh"
#include <mutex>
#include <vector>

namespace r3d {

struct IcpWorkspace;   // defined in icp.cu
size_t icp_workspace_bytes_internal(int n_pairs, int max_src, int max_dst);

}  // namespace r3d

// icp.cu entry points reused here through the public C ABI
using namespace r3d;

namespace {

struct ChunkBuffers {
    uint16_t* depth;        // [2 * ppc][h][w] staging (host variant only): src frames then dst frames
    double* src_points;     // [ppc * h * w] xyzw
    double* dst_points;     // [ppc * h * w] xyzw
    int32_t* src_off;       // [ppc + 1]
    int32_t* dst_off;       // [ppc + 1]
    void* bp_ws;
    size_t bp_ws_bytes;
    void* icp_ws;
    size_t icp_ws_bytes;
};

size_t layout_chunk(ChunkBuffers& c, Arena& a, int ppc, int h, int w) {
    const size_t hw = (size_t)h * w;
    c.depth = a.take<uint16_t>(2 * (size_t)ppc * hw);
    c.src_points = (double*)a.take<Point4>((size_t)ppc * hw);
    c.dst_points = (double*)a.take<Point4>((size_t)ppc * hw);
    c.src_off = a.take<int32_t>((size_t)ppc + 1);
    c.dst_off = a.take<int32_t>((size_t)ppc + 1);
    c.bp_ws_bytes = r3d_backproject_workspace_bytes(ppc, h, w);
    c.bp_ws = a.take<char>(c.bp_ws_bytes);
    c.icp_ws_bytes = r3d_icp_workspace_bytes(ppc, (int)hw, (int)hw);
    c.icp_ws = a.take<char>(c.icp_ws_bytes);
    return a.used;
}

__global__ void counts_kernel(const int32_t* __restrict__ src_off, const int32_t* __restrict__ dst_off, int n,
                              int32_t* __restrict__ n_src, int32_t* __restrict__ n_dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (n_src) n_src[i] = src_off[i + 1] - src_off[i];
    if (n_dst) n_dst[i] = dst_off[i + 1] - dst_off[i];
}

// ---- side streams --------------------------------------------------------------------------------------------
constexpr int MAX_STREAMS = 4;
struct StreamSet {
    cudaStream_t streams[MAX_STREAMS];
    cudaEvent_t done[MAX_STREAMS];
    cudaEvent_t fork;
    int device;
    bool pooled;
};
std::mutex g_pool_mutex;
std::vector<StreamSet*> g_pool_free;      // idle sets of any device (a handful at most)

int make_stream_set(StreamSet*& out, int device) {
    StreamSet* s = new StreamSet();
    s->device = device;
    s->pooled = true;
    cudaError_t e = cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming);
    for (int k = 0; k < MAX_STREAMS && e == cudaSuccess; k++) {
        e = cudaStreamCreateWithFlags(&s->streams[k], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->done[k], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { set_cuda_error(e, "stream pool"); delete s; return R3D_ERR_CUDA; }      // (process is in trouble anyway)
    out = s;
    return R3D_OK;
}

int acquire_streams(StreamSet*& out) {
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) return R3D_ERR_CUDA;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        for (size_t k = 0; k < g_pool_free.size(); k++) {
            if (g_pool_free[k]->device == device) {
                out = g_pool_free[k];
                g_pool_free.erase(g_pool_free.begin() + (long)k);
                return R3D_OK;
            }
        }
    }
    return make_stream_set(out, device);
}

void release_streams(StreamSet* s) {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    g_pool_free.push_back(s);
}

int default_ppc(int ppc) { return ppc > 0 ? ppc : 16; }
int default_streams(int s) { return s > 0 ? (s > MAX_STREAMS ? MAX_STREAMS : s) : 3; }

int register_impl(const uint16_t* depth_src, const uint16_t* depth_dst, bool host_input, int n_pairs, int h, int w, double fx,
                  double fy, double cx, double cy, const r3d_icp_params* params, int ppc, int n_streams, double* T_out,
                  int32_t* iters_out, double* mean_err_out, int32_t* n_src_out, int32_t* n_dst_out, void* workspace,
                  size_t workspace_bytes, cudaStream_t user_stream) {
    if (!depth_src || !depth_dst || !params || !T_out || n_pairs <= 0 || h <= 0 || w <= 0) return R3D_ERR_INVALID;
    ppc = default_ppc(ppc);
    if (ppc > n_pairs) ppc = n_pairs;
    n_streams = default_streams(n_streams);
    {
        // A handful of chunks that do not divide over the streams leave one stream a whole chunk behind the others (128 pairs in
        // chunks of 32 on three streams: 2 + 1 + 1): cut them a little smaller instead (six chunks of 22).  Never larger than the
        // caller's chunk (the workspace was sized for it); per-pair results do not depend on the chunking.
        // Fewer chunks than streams (a batch of 32 in one chunk): one chunk per stream, so that the one-CTA-per-pair steps of a
        // chunk overlap the searches of the others (32 pairs: 1038 -> 1233 pairs/s; 64 pairs: no change).
        int n_chunks = (n_pairs + ppc - 1) / ppc;
        if (n_chunks < n_streams && n_pairs >= 4 * n_streams) {
            n_chunks = n_streams;
            ppc = (n_pairs + n_chunks - 1) / n_chunks;
        } else if (n_chunks > n_streams && n_chunks < 4 * n_streams && n_chunks % n_streams != 0) {
            n_chunks = ((n_chunks + n_streams - 1) / n_streams) * n_streams;
            ppc = (n_pairs + n_chunks - 1) / n_chunks;
        }
    }
    const size_t hw = (size_t)h * w;
    if ((size_t)ppc * hw > 0x7fffffffull) return R3D_ERR_INVALID;

    Arena arena(workspace, workspace_bytes);
    std::vector<ChunkBuffers> cb(n_streams);
    for (int s = 0; s < n_streams; s++) layout_chunk(cb[s], arena, ppc, h, w);
    // device-side result staging for the host variant
    double* T_dev = T_out;
    int32_t* iters_dev = iters_out;
    double* err_dev = mean_err_out;
    int32_t* nsrc_dev = n_src_out;
    int32_t* ndst_dev = n_dst_out;
    if (host_input) {
        T_dev = arena.take<double>((size_t)n_pairs * 16);
        iters_dev = arena.take<int32_t>((size_t)n_pairs);
        err_dev = arena.take<double>((size_t)n_pairs);
        nsrc_dev = arena.take<int32_t>((size_t)n_pairs);
        ndst_dev = arena.take<int32_t>((size_t)n_pairs);
    }
    if (!arena.ok) return R3D_ERR_WORKSPACE;

    // side streams + fork/join events: created once per process and device, lent to one call at a time (a concurrent call
    // from another host thread gets a set of its own and destroys it afterwards)
    StreamSet* set = nullptr;
    int rc = acquire_streams(set);
    if (rc != R3D_OK) return rc;
    cudaStream_t* streams = set->streams;
    cudaEvent_t* done = set->done;
    cudaEvent_t fork = set->fork;
    cudaError_t e0 = cudaEventRecord(fork, user_stream);
    for (int s = 0; s < n_streams && e0 == cudaSuccess; s++) e0 = cudaStreamWaitEvent(streams[s], fork, 0);
    if (e0 != cudaSuccess) {
        set_cuda_error(e0, "register_depth_pairs fork");
        release_streams(set);
        return R3D_ERR_CUDA;
    }

    int chunk_index = 0;
    for (int first = 0; first < n_pairs && rc == R3D_OK; first += ppc, chunk_index++) {
        const int np = (n_pairs - first < ppc) ? (n_pairs - first) : ppc;
        const int s = chunk_index % n_streams;
        cudaStream_t st = streams[s];
        ChunkBuffers& c = cb[s];
        const uint16_t* d_src = depth_src + (size_t)first * hw;
        const uint16_t* d_dst = depth_dst + (size_t)first * hw;
        if (host_input) {
            // stage the chunk's frames; the copy of this chunk overlaps the kernels of the other streams
            cudaError_t ec = cudaMemcpyAsync(c.depth, d_src, (size_t)np * hw * sizeof(uint16_t), cudaMemcpyHostToDevice, st);
            if (ec == cudaSuccess)
                ec = cudaMemcpyAsync(c.depth + (size_t)np * hw, d_dst, (size_t)np * hw * sizeof(uint16_t), cudaMemcpyHostToDevice, st);
            if (ec != cudaSuccess) { set_cuda_error(ec, "register_depth_pairs H2D"); rc = R3D_ERR_CUDA; break; }
            d_src = c.depth;
            d_dst = c.depth + (size_t)np * hw;
        }
        rc = r3d_backproject(d_src, R3D_DEPTH_U16, np, h, w, R3D_BP_PINHOLE, fx, fy, cx, cy, 1.0, R3D_POINTS_XYZW,
                             c.src_points, c.src_off, c.bp_ws, c.bp_ws_bytes, st);
        if (rc != R3D_OK) break;
        rc = r3d_backproject(d_dst, R3D_DEPTH_U16, np, h, w, R3D_BP_PINHOLE, fx, fy, cx, cy, 1.0, R3D_POINTS_XYZW,
                             c.dst_points, c.dst_off, c.bp_ws, c.bp_ws_bytes, st);
        if (rc != R3D_OK) break;
        rc = r3d_icp_batch(c.src_points, c.src_off, c.dst_points, c.dst_off, np, (int)hw, (int)hw, nullptr, params,
                           T_dev + (size_t)first * 16, iters_dev ? iters_dev + first : nullptr,
                           err_dev ? err_dev + first : nullptr, nullptr, nullptr, nullptr, c.icp_ws, c.icp_ws_bytes, st);
        if (rc != R3D_OK) break;
        if (nsrc_dev || ndst_dev) {
            counts_kernel<<<(np + 127) / 128, 128, 0, st>>>(c.src_off, c.dst_off, np, nsrc_dev ? nsrc_dev + first : nullptr,
                                                            ndst_dev ? ndst_dev + first : nullptr);
            g_launches.fetch_add(1);
        }
    }
    if (rc == R3D_OK && host_input) {
        cudaStream_t st = streams[0];
        // results leave from stream 0 after every stream has finished
        for (int s = 1; s < n_streams; s++) {
            cudaEventRecord(done[s], streams[s]);
            cudaStreamWaitEvent(st, done[s], 0);
        }
        cudaMemcpyAsync(T_out, T_dev, (size_t)n_pairs * 16 * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (iters_out) cudaMemcpyAsync(iters_out, iters_dev, (size_t)n_pairs * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (mean_err_out) cudaMemcpyAsync(mean_err_out, err_dev, (size_t)n_pairs * sizeof(double), cudaMemcpyDeviceToHost, st);
        if (n_src_out) cudaMemcpyAsync(n_src_out, nsrc_dev, (size_t)n_pairs * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
        if (n_dst_out) cudaMemcpyAsync(n_dst_out, ndst_dev, (size_t)n_pairs * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
    }
    // join
    for (int s = 0; s < n_streams; s++) {
        cudaEventRecord(done[s], streams[s]);
        cudaStreamWaitEvent(user_stream, done[s], 0);
    }
    cudaError_t e = cudaSuccess;
    if (host_input) e = cudaStreamSynchronize(user_stream);
    release_streams(set);      // (work still queued on the side streams is ordered before anything a later call puts there)
    if (rc != R3D_OK) return rc;
    if (e != cudaSuccess) { set_cuda_error(e, "register_depth_pairs sync"); return R3D_ERR_CUDA; }
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_cuda_error(e, "register_depth_pairs"); return R3D_ERR_CUDA; }
    return R3D_OK;
}

}  // namespace

extern "C" size_t r3d_register_workspace_bytes(int n_pairs, int pairs_per_chunk, int height, int width, int n_streams) {
    if (height <= 0 || width <= 0 || n_pairs <= 0) return 0;
    int ppc = default_ppc(pairs_per_chunk);
    if (ppc > n_pairs) ppc = n_pairs;
    const int ns = default_streams(n_streams);
    Arena a(nullptr, 0);
    ChunkBuffers c;
    for (int s = 0; s < ns; s++) layout_chunk(c, a, ppc, height, width);
    // device-side result staging of the *_host variant
    a.take<double>((size_t)n_pairs * 16);
    a.take<int32_t>((size_t)n_pairs);
    a.take<double>((size_t)n_pairs);
    a.take<int32_t>((size_t)n_pairs);
    a.take<int32_t>((size_t)n_pairs);
    return a.used + 1024;
}

extern "C" int r3d_register_depth_pairs(const uint16_t* depth_src, const uint16_t* depth_dst, int n_pairs, int height, int width,
                                        double fx, double fy, double cx, double cy, const r3d_icp_params* params,
                                        int pairs_per_chunk, int n_streams, double* T_out, int32_t* iters_out,
                                        double* mean_err_out, int32_t* n_src_out, int32_t* n_dst_out, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    return register_impl(depth_src, depth_dst, false, n_pairs, height, width, fx, fy, cx, cy, params, pairs_per_chunk, n_streams,
                         T_out, iters_out, mean_err_out, n_src_out, n_dst_out, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int r3d_register_depth_pairs_host(const uint16_t* depth_src_host, const uint16_t* depth_dst_host, int n_pairs,
                                             int height, int width, double fx, double fy, double cx, double cy,
                                             const r3d_icp_params* params, int pairs_per_chunk, int n_streams,
                                             double* T_out_host, int32_t* iters_out_host, double* mean_err_out_host,
                                             int32_t* n_src_out_host, int32_t* n_dst_out_host, void* workspace,
                                             size_t workspace_bytes, void* stream) {
    return register_impl(depth_src_host, depth_dst_host, true, n_pairs, height, width, fx, fy, cx, cy, params, pairs_per_chunk,
                         n_streams, T_out_host, iters_out_host, mean_err_out_host, n_src_out_host, n_dst_out_host, workspace,
                         workspace_bytes, (cudaStream_t)stream);
}
