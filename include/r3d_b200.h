/*
 * r3d_b200.h -- C ABI of the B200-native registration hot path.
 *
 * The reference (3370sohail/2dto3dreconstruction) has no FFI layer: its "operator API" is
 * module-level Python functions.  Each entry point below is what a ctypes binding for one
 * of those functions calls; the reference function it replaces is cited as file:line into
 * the reference tree.  The Python drop-ins in 2dto3dreconstruction_b200/utils/ are such
 * bindings.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - nothing allocates: scratch memory is caller-provided, sized by *_workspace_bytes();
 *   - return value: 0 on success, negative r3d_status on failure (r3d_status_string());
 *   - calls are asynchronous on `stream` unless stated otherwise;
 *   - points are double precision.  "xyz" = packed (n,3) row-major doubles (NumPy layout),
 *     "xyzw" = (n,4) doubles, 32-byte aligned (records move as 256-bit loads/stores; a misaligned pointer
 *     is rejected with R3D_ERR_INVALID), w unused on input.
 *   - 4x4 transforms are 16 doubles, row-major, acting on column vectors (NumPy layout).
 */
#ifndef R3D_B200_H
#define R3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    R3D_OK = 0,
    R3D_ERR_INVALID = -1,      /* bad argument (shape, null pointer, enum) */
    R3D_ERR_WORKSPACE = -2,    /* workspace too small / misaligned */
    R3D_ERR_CUDA = -3,         /* a CUDA runtime call failed; see r3d_last_cuda_error() */
    R3D_ERR_NO_DEVICE = -4,    /* no sm_100 device visible */
    R3D_ERR_NUMERIC = -5       /* non-finite value where the reference would raise */
} r3d_status;

int r3d_version(void);
const char* r3d_status_string(int status);
const char* r3d_last_cuda_error(void);
/* 0 if a CUDA device of compute capability 10.x is current, else R3D_ERR_NO_DEVICE */
int r3d_check_device(void);

/* ------------------------------------------------------------------------------------
 * K1  back-projection
 *   replaces utils/utils.py:21-45  depth_to_voxel      (mode = R3D_BP_PIXEL)
 *            utils/utils.py:48-76  depth_to_voxel_ld   (mode = R3D_BP_PINHOLE)
 *            reconstruct.py:336-354 depth_images_to_3d_pts[_ld] (frames > 1, one launch)
 *   x = ((u - cx) * d) / fx, y = ((v - cy) * d) / fy, z = (double)(int16)d * zscale,
 *   rows with z == 0 dropped, row-major pixel order kept; bit-exact with NumPy float64.
 * ------------------------------------------------------------------------------------ */
typedef enum { R3D_DEPTH_U16 = 0, R3D_DEPTH_F32 = 1, R3D_DEPTH_F64 = 2 } r3d_depth_dtype;
typedef enum { R3D_BP_PIXEL = 0, R3D_BP_PINHOLE = 1 } r3d_bp_mode;
typedef enum { R3D_POINTS_XYZ = 0, R3D_POINTS_XYZW = 1, R3D_POINTS_XYZW_F32 = 2 } r3d_point_layout;

size_t r3d_backproject_workspace_bytes(int frames, int height, int width);

/* depth: [frames, height, width] contiguous.  out_points: capacity frames*height*width
 * points of the chosen layout; frame f occupies points [out_offsets[f], out_offsets[f+1]).
 * out_offsets: int32[frames+1] (device). */
int r3d_backproject(const void* depth, int depth_dtype, int frames, int height, int width,
                    int mode, double fx, double fy, double cx, double cy, double zscale,
                    int out_layout, void* out_points, int32_t* out_offsets,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * K3  closed-form rigid fit on given correspondences (Kabsch / Arun, reflection fixed)
 *   replaces utils/icp.py:22-55   best_fit_transform(A, B)   (point stride 3, dim stride 1)
 *            utils/r3d.py:22-60   rigid_transform_3D(A, B)   (point stride 1, dim stride K)
 *   a, b: doubles; element (point i, dim d) is at base[i*pt_stride + d*dim_stride].
 *   T_out: 16 doubles.
 * ------------------------------------------------------------------------------------ */
size_t r3d_kabsch_workspace_bytes(int64_t n);
int r3d_kabsch(const double* a, int64_t a_pt_stride, int64_t a_dim_stride,
               const double* b, int64_t b_pt_stride, int64_t b_dim_stride,
               int64_t n, double* T_out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * K3c 12-dof affine least squares  q ~ M p + t
 *   replaces utils/transformation3d.py:163-180 homo_rigid_transform_3d(img1_kp, img2_kp)
 *   p, q: packed xyz doubles (n,3); p is rounded through float32 first, as the reference's
 *   float32 design matrix does (utils/transformation3d.py:146).  T_out: 16 doubles with
 *   last row 0 0 0 1.  status_out (device int32): 0 ok, 1 = rank-deficient point set.
 * ------------------------------------------------------------------------------------ */
size_t r3d_affine_fit_workspace_bytes(int64_t n);
int r3d_affine_fit(const double* p, const double* q, int64_t n, double* T_out, int32_t* status_out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* replaces utils/transformation3d.py:183-187 get_transformed_points(points, h):
 * out = [P 1] h^T, xyz kept, no perspective divide.  in/out packed xyz (n,3); may alias. */
int r3d_transform_points(const double* points, int64_t n, const double* T /* device, 16 */,
                         double* out, void* stream);

/* packed (n,3) doubles -> (n,4) xyzw doubles (32-byte aligned), the layout K2/K3 consume.
 * Used by the drop-ins for clouds that arrive as NumPy (n,3) arrays. */
int r3d_xyz_to_xyzw(const double* xyz, int64_t n, double* xyzw, void* stream);

/* ------------------------------------------------------------------------------------
 * K2  exact nearest neighbour on a uniform voxel grid (sparse 4x4x4 blocks, radix-sorted)
 *   replaces utils/icp.py:58-71 nearest_neighbor(src, dst), batched over pairs.
 *   Exact = identical to float64 brute force with lowest-index tie-break.
 *   src / dst: xyzw points; pair p owns src[src_off[p] .. src_off[p+1]) and likewise dst.
 *   src_off / dst_off: device int32[n_pairs+1].  max_src / max_dst: upper bounds on any
 *   pair's point count (launch geometry; the true counts are read on the device).
 *   out_idx: int32 index into the pair's dst cloud, out_dist: Euclidean distance (sqrt).
 * ------------------------------------------------------------------------------------ */
size_t r3d_nn_workspace_bytes(int n_pairs, int max_src, int max_dst);
int r3d_nn_batch(const double* src, const int32_t* src_off, const double* dst, const int32_t* dst_off,
                 int n_pairs, int max_src, int max_dst, double cell_size /* <=0: automatic */,
                 int32_t* out_idx, double* out_dist,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * K2+K3  batched ICP, whole loop on the device
 *   replaces utils/icp.py:74-133 icp(A, B, init_pose, max_iterations, tolerance,
 *   partial_set_size) for n_pairs independent (A, B) pairs at once.
 *   init_pose: NULL or device double[n_pairs,16].
 *   Outputs (device): T_out double[n_pairs,16]; iters_out int32[n_pairs] with the
 *   reference's convention (1-based, max_iterations+1 when never converged);
 *   mean_err_out double[n_pairs] (mean kept distance of the last executed iteration).
 *   Optional (may be NULL): last_idx int32 / last_dist double, laid out like src: for every
 *   source point the NN index and distance of the last executed iteration; last_keep uint8:
 *   1 where the point was in the kept top-fraction set.
 * ------------------------------------------------------------------------------------ */
typedef struct {
    int32_t max_iterations;      /* reference default 20 */
    int32_t reserved;
    double tolerance;            /* reference default 0.001 */
    double partial_set_size;     /* reference default 1.0, in (0, 1] */
    double cell_size;            /* <= 0: automatic */
} r3d_icp_params;

size_t r3d_icp_workspace_bytes(int n_pairs, int max_src, int max_dst);
int r3d_icp_batch(const double* src, const int32_t* src_off, const double* dst, const int32_t* dst_off,
                  int n_pairs, int max_src, int max_dst, const double* init_pose,
                  const r3d_icp_params* params,
                  double* T_out, int32_t* iters_out, double* mean_err_out,
                  int32_t* last_idx, double* last_dist, uint8_t* last_keep,
                  void* workspace, size_t workspace_bytes, void* stream);

/*   Query-sharded ICP of ONE large pair over the GPUs of a box (BASELINE.json configs[3]).
 *   Every process passes the same clouds and parameters plus its rank; each runs the search of every world-th slice
 *   of the queries and calls `exchange` once per iteration on a device buffer of n_doubles float64 that must be
 *   SUMMED IN PLACE over all ranks, ordered on `stream` (an NCCL all-reduce).  Each element is non-zero on exactly
 *   one rank, so the sums are exact and T_out / iters_out / mean_err_out are bit-identical on every rank and to
 *   r3d_icp_batch on one GPU.  partial_set_size must be 1.0 (the top-fraction selection would need a distributed
 *   k-th select) and n_pairs 1 (batches are sharded by pair: no exchange at all).  exchange returns 0 on success. */
typedef int (*r3d_exchange_fn)(void* ctx, void* device_buf, size_t n_doubles, void* stream);
int r3d_icp_batch_sharded(const double* src, const int32_t* src_off, const double* dst, const int32_t* dst_off,
                          int n_pairs, int max_src, int max_dst, const double* init_pose, const r3d_icp_params* params,
                          int rank, int world, r3d_exchange_fn exchange, void* exchange_ctx,
                          double* T_out, int32_t* iters_out, double* mean_err_out,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * K4A  2-D homography RANSAC scoring, reference-exact
 *   replaces the body of the hypothesis loop of utils/homography_utils/q9.py:146-210:
 *   q9.py:52-69 get_transformed_points, q9.py:129-144 make_binned_array and
 *   q9.py:92-127 get_ssd_v2, for all hypotheses in one launch.
 *   kp1, kp2: (n,2) doubles as (row, col).  H: n_hyp x 9 doubles (real part) and, when
 *   H_imag != NULL, n_hyp x 9 imaginary parts (numpy.linalg.eig can hand the reference a
 *   complex homography).  counts_out int32[n_hyp] = len(matches) per hypothesis.
 *   match_j_out (optional, int32[n_hyp, n1]): for keypoint i the accepted kp2 index or -1;
 *   match_ssd_out (optional, double[n_hyp, n1]).  flags_out int32[n_hyp]: bit0 = a
 *   transformed coordinate was NaN, bit1 = infinite (the reference raises in int()).
 * ------------------------------------------------------------------------------------ */
size_t r3d_ransac2d_workspace_bytes(int n_hyp, int n1, int n2, int img_h, int img_w);
int r3d_ransac2d_score(const double* H, const double* H_imag, int n_hyp,
                       const double* kp1, int n1, const double* kp2, int n2,
                       int img_h, int img_w,
                       int32_t* counts_out, int32_t* match_j_out, double* match_ssd_out,
                       int32_t* flags_out,
                       void* workspace, size_t workspace_bytes, void* stream);

/* replaces utils/homography_utils/q9.py:52-69 get_transformed_points(keypoints, H_matrix):
 * (row, col) keypoints through one 3x3 homography (9 doubles, + optional imaginary parts),
 * divided by w, returned as (row, col).  kp / out: (n,2) doubles. */
int r3d_homography_points(const double* H, const double* H_imag, const double* kp, int n,
                          double* out, void* stream);

/* ------------------------------------------------------------------------------------
 * K4B  3-D hypothesis scoring (rigid or affine 4x4): counts[h] = #{ i : |T_h p_i - q_i|^2 < thresh }
 *   No reference implementation exists; composition of
 *   utils/transformation3d.py:183-187 and the SSD threshold of q9.py:122.
 *   T: n_hyp x 16 doubles (rows 0..2 used).  P, Q: xyzw doubles.
 * ------------------------------------------------------------------------------------ */
int r3d_ransac3d_score(const double* T, int n_hyp, const double* P, const double* Q, int n_pts,
                       double thresh, int32_t* counts_out, void* stream);

/* ------------------------------------------------------------------------------------
 * K6  minimal-sample hypothesis generation, one hypothesis per thread
 *   r3d_rigid_hypotheses:  3 correspondences -> rigid 4x4, utils/r3d.py:22-60 rigid_transform_3D on the sample
 *   r3d_affine_hypotheses: 4 correspondences -> affine 4x4, utils/transformation3d.py:163-180
 *                          homo_rigid_transform_3d on the sample (exactly determined: a 4x4 solve)
 *   P, Q: xyzw doubles (n_pts); samples: int32 [n_hyp, 3 | 4] indices into P/Q, drawn by the caller (the
 *   reference samples with NumPy's global RNG); T_out: [n_hyp, 16]; status_out (optional) int32 [n_hyp]:
 *   0 ok, 1 degenerate sample (repeated / collinear / coplanar points: the fit is not unique).
 * ------------------------------------------------------------------------------------ */
int r3d_rigid_hypotheses(const double* P, const double* Q, int n_pts, const int32_t* samples, int n_hyp,
                         double* T_out, int32_t* status_out, void* stream);
int r3d_affine_hypotheses(const double* P, const double* Q, int n_pts, const int32_t* samples, int n_hyp,
                          double* T_out, int32_t* status_out, void* stream);

/* ------------------------------------------------------------------------------------
 * K5  brute-force descriptor matching with cross-check (the step before RANSAC)
 *   replaces cv2.BFMatcher(normType=cv2.NORM_L2, crossCheck=True).match(des1, des2) as called at
 *   utils/transformation3d.py:31-32 (OpenCV is a third-party dependency of the reference, not vendored).
 *   des1: [n1, dim] float32 query descriptors, des2: [n2, dim] float32 train descriptors.
 *   train_idx_out int32[n1]: the matched train index of every query, or -1 (cross_check != 0: kept only when
 *   the two are each other's nearest, first minimum on ties); dist_out float32[n1]: L2 distance (inf if none).
 * ------------------------------------------------------------------------------------ */
size_t r3d_match_workspace_bytes(int n1, int n2);
int r3d_match_descriptors(const float* des1, int n1, const float* des2, int n2, int dim, int cross_check,
                          int32_t* train_idx_out, float* dist_out,
                          void* workspace, size_t workspace_bytes, void* stream);
/*   ... with the second-best ratio test of skimage.feature.match_descriptors(des1, des2, max_ratio=0.8,
 *   cross_check=True) as called at utils/homography_utils/q8.py:90 (scikit-image 0.17.2, not vendored): after the
 *   mutual check a query is kept iff best / second_best < max_ratio (float64 distances; second_best = nearest OTHER
 *   train descriptor, 0 replaced by DBL_EPSILON).  max_ratio >= 1 disables the test, as in skimage. */
int r3d_match_descriptors_ratio(const float* des1, int n1, const float* des2, int n2, int dim, int cross_check,
                                double max_ratio, int32_t* train_idx_out, float* dist_out,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * K7  voxel down-sampling of a (merged) cloud (the step right after registration)
 *   replaces open3d.geometry.PointCloud.voxel_down_sample(voxel_size) as called in the merge loop of
 *   chain_transformation, reconstruct.py:212-220 (Open3D is a third-party dependency of the reference, not vendored):
 *   min_bound = min(points) - voxel_size/2; voxel = floor((p - min_bound) / voxel_size); one output point per occupied
 *   voxel = mean of its points, summed in ascending index order.  Output order: ascending (z, y, x) voxel index
 *   (Open3D's is the iteration order of an unordered_map, i.e. unspecified).
 *   xyz: [n, 3] float64 packed (device); xyz_out: [<= n, 3]; n_out: int32 (device) number of voxels;
 *   status_out (device, optional): 0 ok, 1 = voxel_size too small for the extent (Open3D raises), 2 = a non-finite
 *   coordinate; n_out = 0 in both cases.  The transform of the merge step is r3d_transform_points.
 * ------------------------------------------------------------------------------------ */
size_t r3d_voxel_downsample_workspace_bytes(int64_t n);
int r3d_voxel_downsample(const double* xyz, int64_t n, double voxel_size, double* xyz_out, int32_t* n_out,
                         int32_t* status_out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused registration of depth-frame pairs: K1 on both frames, grid build, ICP.
 *   The batched form of the reference's pair loop body (reconstruct.py:319-324 /
 *   reconstruct_rgbd.py:48) for the synthetic-depth configuration.
 *   depth_src / depth_dst: uint16 [n_pairs, height, width] (device, or pinned host for the
 *   *_host variant, which stages chunks through the workspace with copies overlapped).
 *   n_src_out / n_dst_out: int32[n_pairs] cloud sizes.  pairs_per_chunk: the LARGEST number of
 *   pairs processed per launch group (<=0: default); a batch that gives fewer chunks than
 *   streams, or a handful that do not divide over them, is cut into smaller, equal chunks
 *   (the workspace is sized for the value given; per-pair results do not depend on it).
 *   The *_host variant synchronises before returning; outputs are HOST pointers there.
 * ------------------------------------------------------------------------------------ */
size_t r3d_register_workspace_bytes(int n_pairs, int pairs_per_chunk, int height, int width, int n_streams);
int r3d_register_depth_pairs(const uint16_t* depth_src, const uint16_t* depth_dst,
                             int n_pairs, int height, int width,
                             double fx, double fy, double cx, double cy,
                             const r3d_icp_params* params, int pairs_per_chunk, int n_streams,
                             double* T_out, int32_t* iters_out, double* mean_err_out,
                             int32_t* n_src_out, int32_t* n_dst_out,
                             void* workspace, size_t workspace_bytes, void* stream);
int r3d_register_depth_pairs_host(const uint16_t* depth_src_host, const uint16_t* depth_dst_host,
                                  int n_pairs, int height, int width,
                                  double fx, double fy, double cx, double cy,
                                  const r3d_icp_params* params, int pairs_per_chunk, int n_streams,
                                  double* T_out_host, int32_t* iters_out_host, double* mean_err_out_host,
                                  int32_t* n_src_out_host, int32_t* n_dst_out_host,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* Tuning knobs; none of them changes results (process-wide, read when a call is issued).
 *   R3D_OPT_NN_IMPL        1 (the only value: the search of csrc/lean_search.cuh; round 2's TMA-staged tile search was exact and
 *                          slower -- profiles/r02_nn_tile_search.md -- and has been removed; the option stays for ABI stability)
 *   R3D_OPT_FORCE_UNFUSED  non-zero: partial_set_size == 1 also runs through the selection + moments kernels (test knob)
 *   R3D_OPT_CELL_FACTOR    automatic grid cell edge in units of the estimated point spacing (used when cell_size <= 0; default 2)
 *   R3D_OPT_SLACK_MULT, R3D_OPT_SLACK_MAX_FRAC
 *                          how much larger than its bound a seeded search ball is made: a multiple of the query's movement in
 *                          this iteration (default 8), used while that is at most a fraction of the cell edge (default 0.25);
 *                          a multiple of 0 disables the skipping of searches that cannot change the neighbour
 *   R3D_OPT_GROUP_SEARCH_MAX  0..32 (default 12): seeded tasks of the ICP loop in which at most this many of the 32 lanes still
 *                          search are served by csrc/lean_search.cuh group_search (eight lanes per query, only that query's
 *                          own cells) instead of the warp-union walk; 0 = always the union walk
 *   R3D_OPT_SLACK_CLAMP    non-zero: a query that moves too fast for the SLACK_MULT rule still widens its ball by SLACK_MAX_FRAC cells
 *                          (instead of not at all), which buys it a skip budget for the next iteration or two
 *   R3D_OPT_K4B_FP32       non-zero: r3d_ransac3d_score filters in fp32 and redoes the residual in fp64 inside the rounding band
 *                          (same counts as the default all-fp64 kernel; measured slower, kept as the cross-check) */
typedef enum {
    R3D_OPT_CELL_FACTOR = 2, R3D_OPT_NN_IMPL = 3, R3D_OPT_SLACK_MULT = 4, R3D_OPT_SLACK_MAX_FRAC = 5, R3D_OPT_FORCE_UNFUSED = 6,
    R3D_OPT_GROUP_SEARCH_MAX = 7, R3D_OPT_K4B_FP32 = 8, R3D_OPT_SLACK_CLAMP = 9
} r3d_option;
int r3d_set_option(int option, double value);

/* Number of kernels this library has launched since load (all entry points). */
uint64_t r3d_launch_count(void);

/* Optional device-event timing of the dominant kernel (nn search): when enabled, every
 * nn-search launch is bracketed by cudaEvents on its own stream.  r3d_nn_timing_read
 * synchronises those events and returns accumulated milliseconds / launch count /
 * queries processed since the last reset. */
void r3d_nn_timing_enable(int on);
int r3d_nn_timing_read(double* total_ms, int64_t* launches, int64_t* queries, int reset);

/* Diagnostic counters of the nn search (only in builds with -DR3D_NN_STATS, otherwise returns
 * R3D_ERR_INVALID): out8_host[0..7] = tasks (warps of 32 queries), search rounds, candidate
 * evaluations per lane, searches skipped (lanes), segment lookup rounds, segments looked up, non-empty
 * segments, skipped lanes whose seed was beaten by an evaluated point (must be 0).  Synchronous. */
int r3d_nn_stats_read(uint64_t* out8_host, int reset);
int r3d_nn_stats_read_ext(uint64_t* out32_host, int reset);   /* 32 counters: + far-path / task-duration / group-search counters (csrc/lean_search.cuh) */

#ifdef __cplusplus
}
#endif
#endif /* R3D_B200_H */
