/*
 * fm3d.h -- C-ABI of libfm3d: the B200 (sm_100a) implementation of the
 * match -> triangulate -> normal-optimise -> patch-extract hot path of
 * caomw/3DFeatureMatcher.
 *
 * The reference has no FFI of its own: its boundary is the public surface of four
 * C++ classes.  Every entry point below names the reference method(s) it replaces
 * (file:line under the reference tree).  The class adapters in
 * 3dfeaturematcher_b200/host/ (same class names and signatures as the reference) and
 * the ctypes binding in 3dfeaturematcher_b200/api.py are thin wrappers over exactly
 * these symbols.
 *
 * Conventions
 *   - plain pointers and sizes only; all arrays row-major and caller-allocated;
 *   - functions without a _dev suffix take HOST pointers, copy in/out on the context's
 *     stream and return after the result is in the caller's buffers;
 *   - functions with a _dev suffix take DEVICE pointers (memory of the context's GPU),
 *     enqueue work on the context's stream and return without synchronising
 *     (fm3d_sync waits);
 *   - every function returns FM3D_OK (0) or a negative fm3d_status; fm3d_last_error
 *     gives the text.  Nothing calls exit(), nothing throws across the boundary;
 *   - there is no CPU implementation behind any of these calls: without a usable
 *     sm_100 device fm3d_ctx_create fails with FM3D_ERR_NO_DEVICE.
 *   - one fm3d_ctx per host thread / per GPU; a context is not re-entrant.
 */
#ifndef FM3D_H_
#define FM3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FM3D_VERSION 100 /* 0.1.0 */

typedef struct fm3d_ctx fm3d_ctx;

typedef enum fm3d_status {
    FM3D_OK = 0,
    FM3D_ERR_INVALID_ARG = -1,
    FM3D_ERR_CUDA = -2,
    FM3D_ERR_NO_DEVICE = -3,
    FM3D_ERR_STATE = -4,       /* camera / images / pose not set yet */
    FM3D_ERR_UNSUPPORTED = -5,
    FM3D_ERR_NOMEM = -6,
    FM3D_ERR_COMM = -7         /* NCCL could not be loaded, or a collective failed */
} fm3d_status;

/* Per-feature outcome of the normal optimiser.  The reference silently erases a failed
 * feature from points3D ("Bad point!" / "Not enough pixels!",
 * Triangulator/normaloptimizer.cpp:364-382); here every feature keeps its index and
 * gets a status instead. */
typedef enum fm3d_feature_status {
    FM3D_FEAT_OK = 0,
    FM3D_FEAT_NO_PIXELS = 1,  /* m_dat == 0,                     normaloptimizer.cpp:364-369 */
    FM3D_FEAT_ABORT_BBOX = 2, /* isInBoundingBox failed,         singlecameratriangulator.cpp:557-560 */
    FM3D_FEAT_ABORT_PIXEL = 3,/* isPixelGood failed (img 1 or 2), singlecameratriangulator.cpp:580-584,623-626 */
    FM3D_FEAT_ABORT_NAN = 4   /* NaN normal / NaN plane point,   normaloptimizer.cpp:81-85, singlecameratriangulator.cpp:465-469 */
} fm3d_feature_status;

/* Semantics of the unqualified abs() in the angle-penalty wall of evaluateNormal
 * (Triangulator/normaloptimizer.cpp:133-138). */
typedef enum fm3d_penalty_mode {
    FM3D_PENALTY_FABS = 0,    /* abs() == fabs(): the source as compiled by g++ >= 6 */
    FM3D_PENALTY_INT_ABS = 1, /* abs() == int abs(int): pre-C++11 toolchains */
    FM3D_PENALTY_OFF = 2
} fm3d_penalty_mode;

/* What the normal search minimises (option "normals_cost").  The reference knows only the first: plain SSD of the
 * bilinear intensities, fvec[i] = w (I1_i - I2_i) (Triangulator/normaloptimizer.cpp:145-148; SURVEY fact 4: "there
 * is no NCC anywhere").  FM3D_COST_NCC is an extension the north star names: the zero-mean normalised residual
 * fvec[i] = w ((I1_i - mean I1)/|I1 - mean I1| - (I2_i - mean I2)/|I2 - mean I2|), whose squared sum is
 * 2 - 2 NCC(I1, I2) -- invariant to gain and offset between the two frames -- minimised by the same
 * Levenberg-Marquardt, level by level.  Same gates, same statuses; a patch without texture is dropped (NaN). */
typedef enum fm3d_cost_mode {
    FM3D_COST_SSD = 0,
    FM3D_COST_NCC = 1
} fm3d_cost_mode;

/* ------------------------------------------------------------------ context ---- */

int fm3d_version(void);

/* Creates a context bound to CUDA device `device` (one stream, scratch buffers). */
int fm3d_ctx_create(int device, fm3d_ctx** out);
void fm3d_ctx_destroy(fm3d_ctx* ctx);
const char* fm3d_last_error(const fm3d_ctx* ctx);
/* Waits for the context's stream.  Errors that an asynchronous (_dev) call could only detect on the device
 * are reported here: FM3D_ERR_CUDA if a TMA window load of fm3d_optimize_normals_dev timed out (the results
 * are then valid but were taken through the slow global-memory sampler). */
int fm3d_sync(fm3d_ctx* ctx);
/* cudaStream_t of the context, as an opaque pointer (for event timing by the caller). */
void* fm3d_stream(fm3d_ctx* ctx);
/* Device memory of the context's GPU for callers of the _dev entry points that do not link a CUDA runtime
 * themselves (the C++ class adapters when they shard over several GPUs).  The copies run on the context's
 * stream and return when they are complete. */
int fm3d_dev_malloc(fm3d_ctx* ctx, size_t bytes, void** out);
int fm3d_dev_free(fm3d_ctx* ctx, void* p);
int fm3d_copy_h2d(fm3d_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int fm3d_copy_d2h(fm3d_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int fm3d_device_info(fm3d_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor,
                     char* name, int name_len);

/* Tuning / mode knobs that are not part of the reference's settings.yml:
 *   "geometry_f32"      0 (default): fp64 ray/plane/projection arithmetic as the reference;
 *                       1: fp32 geometry (faster, normals agree to <<0.5 deg)
 *   "matcher_tensor"    1 (default): use the tcgen05 contraction -- directly when descriptors are
 *                       integer-valued in [0,255] and dim == 128, as a filter followed by an exact
 *                       fp32 decision for other float descriptors (dim <= 128, dim % 4 == 0, >= 2^22
 *                       pairs); results are identical either way.  0: always the exact CUDA-core path
 *   "matcher_sp_tile"   train rows per tile of that filter when dim > 80: 128 (default, two stages) or 256 (one stage)
 *   "matcher_persistent" 1 (default): both tensor-core matchers run one persistent CTA per SM over equal ranges of the (query
 *                       tile, train tile) sequence while the re-tiled train operand fits L2 (64 MB); 2: always; 0: one CTA per
 *                       (query tile, train split).  Same result bits either way
 *   "matcher_min_tiles" persistent matchers: fewest train tiles worth a CTA of its own (default 1)
 *   "matcher_splits"    one-CTA-per-item kernels: train splits per query tile (0 = split only below two CTAs per SM); a value
 *                       > 0 also selects those kernels
 *   "matcher_exact_fallback"  read-only: queries of the last filtered match that the exact path decided
 *   "lm_patience"       lm_control.patience / maxcall (default 100 -> 300 evaluations/level)
 *   "normals_threads"   threads per CTA of the normal optimiser (default 512)
 *   "normals_fast"      1 (default): fm3d_normals_fast.cu (fp32 offset-form geometry, analytic
 *                       Jacobian); 0: the evaluation-by-evaluation fp64 kernel; 2: also route
 *                       fm3d_evaluate_normals through the fast kernel
 *   "normals_cost"      fm3d_cost_mode: 0 (default) the reference's SSD, 1 NCC (fast kernel only: with
 *                       normals_fast = 0, and in fm3d_sweep_normals, NCC returns FM3D_ERR_UNSUPPORTED)
 *   "normals_fuse"      fast kernel: evaluate the Jacobian together with a trial point, so that an accepted trial opens the
 *                       next LM iteration without another pass: 3 (default) every trial, 1 the first trial of an iteration
 *                       only, 2 that one if the previous first trial was accepted, 0 never.  Same iterates either way
 *   "normals_memo"      fast kernel: answer from memory what a pass would return bit for bit -- 1: trial points whose fp32
 *                       coefficients equal the iterate's; 2: also Jacobian requests at such points (SSD cost); 3 (default):
 *                       also trial / Jacobian requests whose coefficients equal those of one of the last four Jacobian
 *                       passes of the level; 0: evaluate everything
 *   "normals_level_sync" two-slot kernel: 1 = the warps of a slot wait at the end of a level set-up until the slot's LM warp has
 *                       started the level's LM; 0 (default) = they go on to the partner slot (same results)
 *   "normals_groups"    fast kernel: independent feature pipelines per CTA: 1, 2, or 0 (default:
 *                       2 when there are more features than SMs and the layout fits)
 *   "normals_sweep_batch"  fast kernel, fm3d_sweep_normals: 4 (default) evaluates four candidate normals per
 *                       pass over the disc (bit-identical costs, one barrier pair per batch); 1: one pass each
 *   "normals_tma"       stage the image window with a TMA tensor-tile load (1)
 *   "pyramid_fused"     1 (default): fm3d_set_images[_dev] builds three pyramid levels (and, for device frames, level 0)
 *                       per launch; 0: one launch per level plus the copies into level 0 (same bits)
 * Returns FM3D_ERR_INVALID_ARG for an unknown key. */
int fm3d_set_option(fm3d_ctx* ctx, const char* key, double value);
int fm3d_get_option(fm3d_ctx* ctx, const char* key, double* value);

/* Counters of the last call that launched kernels on this context (for benchmarking):
 * number of kernel launches and of bulk-copy (memcpy/memset) operations. */
int fm3d_get_launch_counters(fm3d_ctx* ctx, int64_t* kernel_launches, int64_t* copies);

/* ------------------------------------------------------------- several GPUs ---- */

/* The path shards by query keypoint / by feature over the GPUs of one box (the reference is one
 * sequential loop, Triangulator/normaloptimizer.cpp:335-449, and has no counterpart for any of this).  Two exchanges
 * exist, both on the context's stream, both asynchronous: replicate the inputs every GPU needs in full (train
 * descriptors, train keypoints, both frames: pack them into ONE device buffer and call fm3d_broadcast_dev once) and
 * collect the per-shard results (fm3d_pack_shard_dev writes one fixed-size block from device-resident counts,
 * fm3d_allgather_dev gathers the blocks of all ranks).  NCCL is loaded at run time; without it these calls return
 * FM3D_ERR_COMM and everything else keeps working.
 *
 * One process per GPU: rank 0 calls fm3d_comm_unique_id and hands the 128 bytes to the other ranks over whatever
 * channel the application has; every rank then calls fm3d_comm_init_rank (collective).
 * One process, several GPUs: fm3d_comm_init_all over the contexts of the process; use the *_all_dev forms, which issue
 * the per-context calls inside one NCCL group (ctxs[k] is rank k). */
#define FM3D_COMM_ID_BYTES 128
int fm3d_comm_unique_id(uint8_t id[FM3D_COMM_ID_BYTES]);
int fm3d_comm_init_rank(fm3d_ctx* ctx, const uint8_t id[FM3D_COMM_ID_BYTES], int nranks, int rank);
int fm3d_comm_init_all(fm3d_ctx** ctxs, int n);
int fm3d_comm_info(fm3d_ctx* ctx, int* nranks, int* rank, int* nccl_version);
int fm3d_comm_destroy(fm3d_ctx* ctx);
int fm3d_broadcast_dev(fm3d_ctx* ctx, void* buf, size_t bytes, int root);
int fm3d_allgather_dev(fm3d_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank);
int fm3d_broadcast_all_dev(fm3d_ctx** ctxs, int n, void* const* bufs, size_t bytes, int root);
int fm3d_allgather_all_dev(fm3d_ctx** ctxs, int n, const void* const* send, void* const* recv,
                           size_t bytes_per_rank);

/* Block a rank contributes to the gather of a step.  Header (int32): [0] matches, [1] inliers, [2] rank, [3] cap,
 * [4] query offset of the shard; then `cap` rows of every part at the byte offsets below (rows beyond the counts are
 * zero).  qidx are GLOBAL query indices (local index + query_offset), so the blocks of all ranks in rank order list the
 * matches in ascending query order, as DescriptorsMatcher::compareWithNNDR returns them (descriptorsmatcher.cpp:119-129);
 * src_idx[k] is the row, in the same block's match list, of the match inlier k was triangulated from. */
typedef struct fm3d_shard_layout {
    int cap;
    size_t off_qidx, off_tidx, off_dist;      /* int32, int32, float   x cap : the NNDR matches of the shard          */
    size_t off_src, off_normals, off_status;  /* int32, 3 x double, int32 x cap : the depth-gated, refined features   */
    size_t bytes;                              /* size of one block (multiple of 256)                                  */
} fm3d_shard_layout;
int fm3d_shard_block_layout(int cap, fm3d_shard_layout* layout);
int fm3d_pack_shard_dev(fm3d_ctx* ctx, int cap, int rank, int query_offset, const int32_t* n_match_dev,
                        const int32_t* n_inl_dev, const int32_t* qidx, const int32_t* tidx,
                        const float* dist, const int32_t* src_idx, const double* normals,
                        const int32_t* status, void* block);

/* ------------------------------------------------------------------- camera ---- */

/* Replaces the camera part of SingleCameraTriangulator::SingleCameraTriangulator
 * (Triangulator/singlecameratriangulator.cpp:67-112): K row-major 3x3,
 * dist = (k1,k2,p1,p2,k3) in OpenCV order (settings.yml names them k0,k1,p1,p2,k2),
 * and the depth gate zThresholdMin/zThresholdMax. */
int fm3d_set_camera(fm3d_ctx* ctx, const double K[9], const double dist[5],
                    double z_min, double z_max);

/* Replaces the result of SingleCameraTriangulator::setg12
 * (Triangulator/singlecameratriangulator.cpp:123-143): g12 row-major 4x4,
 * X_cam2 = R * X_cam1 + t.  Composing g12 from (T1,T2,rodrigues1,rodrigues2,g_IC) is
 * host arithmetic done by the class adapter (fm3d_compose_g12 below). */
int fm3d_set_g12(fm3d_ctx* ctx, const double g12[16]);

/* Host helper (no GPU work): g12 = g_IC^-1 * g2^-1 * g1 * g_IC with g_i = [Rodrigues(r_i) | T_i]
 * (singlecameratriangulator.cpp:123-143, tools.cpp:87-99). */
int fm3d_compose_g12(const double T1[3], const double T2[3], const double rod1[3],
                     const double rod2[3], const double rodIC[3], const double tIC[3],
                     double g12_out[16]);

/* ----------------------------------------------------------------- matching ---- */

/* Exact brute-force 2-nearest-neighbour search, query = frame A, train = frame B.
 * Replaces matcher_->knnMatch(desc_a, desc_b, matches, 2)
 * (DescriptorsMatcher/descriptorsmatcher.cpp:84-85,101,117) with the exact search the
 * north star asks for (the reference's FLANN index is approximate and randomised).
 *   q: nq x dim float32, t: nt x dim float32.
 *   idx[2*i+k]  train index of the k-th neighbour of query i (-1 if nt <= k)
 *   dist[2*i+k] L2 distance (square-rooted, float32) as cv::DMatch::distance
 * Ties are broken by the lower train index. */
int fm3d_match_knn2_f32(fm3d_ctx* ctx, const float* q, int nq, const float* t, int nt,
                        int dim, int32_t* idx, float* dist);

/* Same for binary descriptors (ORB 32 B, BRISK/FREAK 64 B): Hamming distance as float. */
int fm3d_match_knn2_hamming(fm3d_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t,
                            int nt, int nbytes, int32_t* idx, float* dist);

/* Replaces the matching part of DescriptorsMatcher::compareWithNNDR
 * (descriptorsmatcher.cpp:117-129): keep the best match of query i iff it has two
 * neighbours and (double)d0 <= eps * (double)d1.  Output is ordered by ascending query
 * index; *nmatch entries are written to qidx/tidx/dist (capacity nq each).
 * If mutual != NULL the role-swapped search is also run and mutual[j] = 1 iff the best
 * train of qidx[j] has qidx[j] as its own best query (the fused cross-check; the
 * reference's crosscompare, :74-87, only returns both raw lists). */
int fm3d_match_nndr_f32(fm3d_ctx* ctx, const float* q, int nq, const float* t, int nt,
                        int dim, double eps, int32_t* qidx, int32_t* tidx, float* dist,
                        uint8_t* mutual, int* nmatch);
int fm3d_match_nndr_hamming(fm3d_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t,
                            int nt, int nbytes, double eps, int32_t* qidx, int32_t* tidx,
                            float* dist, uint8_t* mutual, int* nmatch);

/* Device-pointer variants (asynchronous).  nmatch_dev is a device int. */
int fm3d_match_knn2_f32_dev(fm3d_ctx* ctx, const float* q, int nq, const float* t, int nt,
                            int dim, int32_t* idx, float* dist);
int fm3d_match_knn2_hamming_dev(fm3d_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t,
                                int nt, int nbytes, int32_t* idx, float* dist);
int fm3d_nndr_filter_dev(fm3d_ctx* ctx, const int32_t* idx, const float* dist, int nq,
                         double eps, int32_t* qidx, int32_t* tidx, float* dist_out,
                         int* nmatch_dev);

/* ------------------------------------------------------------ triangulation ---- */

/* Replaces SingleCameraTriangulator::setKeypoints + ::triangulate
 * (singlecameratriangulator.cpp:145-230): for match i take kp1[qidx[i]], kp2[tidx[i]]
 * (float x,y pairs, as cv::KeyPoint::pt), undistort both (5 fixed-point iterations, as
 * cv::undistortPoints), 4x4 DLT with P1=[I|0], P2=g12[0:3,:], dehomogenise, and keep the
 * point iff z_min <= Z < z_max.
 *   xyz_all  n x 3 : dehomogenised point of every match (also the gated-out ones)
 *   mask     n     : 1 = inlier (the reference's outliersMask)
 *   xyz      n x 3 : order-preserving compaction of the inliers (first *ninl rows valid)
 *   src_idx  n     : match index of every compacted row (nullable)
 * Pass qidx == tidx == NULL to use kp1[i], kp2[i] directly. */
int fm3d_triangulate(fm3d_ctx* ctx, const float* kp1, int n1, const float* kp2, int n2,
                     const int32_t* qidx, const int32_t* tidx, int n, double* xyz_all,
                     uint8_t* mask, double* xyz, int32_t* src_idx, int* ninl);
int fm3d_triangulate_dev(fm3d_ctx* ctx, const float* kp1, int n1, const float* kp2, int n2,
                         const int32_t* qidx, const int32_t* tidx, int n, double* xyz_all,
                         uint8_t* mask, double* xyz, int32_t* src_idx, int* ninl_dev);

/* cv::undistortPoints as called at singlecameratriangulator.cpp:169-170,542 (no R, no P:
 * normalised ideal coordinates), exposed for parity tests of the primitive. */
int fm3d_undistort_points(fm3d_ctx* ctx, const double* pts, int n, double* out);

/* ------------------------------------------------------------------- images ---- */

/* Replaces NormalOptimizer::setImages / compute_pyramids
 * (Triangulator/normaloptimizer.cpp:191-221) and SingleCameraTriangulator::setImages
 * (:116-120): uploads both 8-bit images (row stride `stride` bytes) and builds
 * `pyramids` cv::pyrDown levels of each on the GPU (pyramids+1 images per view). */
int fm3d_set_images(fm3d_ctx* ctx, const uint8_t* img1, const uint8_t* img2, int w, int h,
                    int stride, int pyramids);
int fm3d_set_images_dev(fm3d_ctx* ctx, const uint8_t* img1, const uint8_t* img2, int w,
                        int h, int stride, int pyramids);
/* The same with one row stride per image: the reference keeps each cv::Mat at its own step
 * (singlecameratriangulator.cpp:116-120), so a ROI or a padded frame may differ from the other. */
int fm3d_set_images2(fm3d_ctx* ctx, const uint8_t* img1, int stride1, const uint8_t* img2,
                     int stride2, int w, int h, int pyramids);
/* Copies pyramid level `level` of image `image` (1 or 2) to the host (tightly packed). */
int fm3d_get_pyramid_level(fm3d_ctx* ctx, int image, int level, uint8_t* out, int* w,
                           int* h);

/* ------------------------------------------------------- normal optimisation ---- */

/* Replaces NormalOptimizer::computeOptimizedNormals (normaloptimizer.cpp:321-452) with
 * everything below it: extractPixelsContour (singlecameratriangulator.cpp:341-397),
 * optimize_pyramid/optimize (normaloptimizer.cpp:223-292), lmmin (lmfit), evaluateNormal
 * (normaloptimizer.cpp:65-149) and the three per-evaluation helpers
 * (singlecameratriangulator.cpp:530-665).
 *   xyz        n x 3  feature points in camera-1 coordinates
 *   pixels_ray        Neighborhoods.pixelsRay   (disc radius in image-1 pixels)
 *   epsilon_lmmin     Neighborhoods.epsilonLMMIN (lm_control.epsilon)
 *   penalty_mode      fm3d_penalty_mode
 *   normals    n x 3  refined unit normals (initial guess P/|P| if the feature failed)
 *   status     n      fm3d_feature_status
 *   nfev       n x (pyramids+1), column l = evaluations spent at pyramid level l (nullable)
 *   npenalty   n      evaluations that entered the penalty branch (nullable)
 *   cost       n      final sum of squared residuals at level 0 (nullable)
 * The number of pyramid levels is the one given to fm3d_set_images. */
int fm3d_optimize_normals(fm3d_ctx* ctx, const double* xyz, int n, int pixels_ray,
                          double epsilon_lmmin, int penalty_mode, double* normals,
                          int32_t* status, int32_t* nfev, int32_t* npenalty, double* cost);
int fm3d_optimize_normals_dev(fm3d_ctx* ctx, const double* xyz, int n, int pixels_ray,
                              double epsilon_lmmin, int penalty_mode, double* normals,
                              int32_t* status, int32_t* nfev, int32_t* npenalty,
                              double* cost);

/* Work executed by the last fm3d_optimize_normals[_dev] / fm3d_evaluate_normals call of this
 * context (synchronises the stream):
 *   out[0] value-only passes over the disc      out[1] value+Jacobian passes at the LM iterate
 *   out[2] trial passes evaluated with Jacobian out[3] ... of which lmfit accepted the trial
 *   out[4] passes that sampled from global memory (window miss)
 *   out[5] pixel evaluations of value-only passes   out[6] of value+Jacobian passes
 *   out[7] features processed
 *   out[8..10] SM cycles of thread 0 in: the pixel loop of the passes / waiting for the slowest
 *   warp / the serial reduction + LM step + homography set-up;  out[11], out[12] split of out[10]
 *   into LM algebra and homography set-up;  out[13] trial points answered without a pass (their
 *   fp32 homography coefficients equal the iterate's);  out[14] cycles of the per-feature prologue
 *   (disc lattice + undistorted rays), out[15] of the per-level set-up (window staging + image-1 samples)
 * nfev (above) counts what lmfit would have evaluated; these count what the GPU did. */
int fm3d_get_normals_stats(fm3d_ctx* ctx, int64_t out[16]);

/* One evaluation of the cost the optimiser minimises (evaluateNormal,
 * normaloptimizer.cpp:65-149) for given normals at pyramid level `level`:
 * cost[i] = sum of squared weighted residuals, m[i] = number of disc pixels,
 * status[i] as above.  Used by parity tests and by the dense candidate-normal sweep. */
int fm3d_evaluate_normals(fm3d_ctx* ctx, const double* xyz, const double* normals_phi_theta,
                          int n, int pixels_ray, int level, int penalty_mode, double* cost,
                          int32_t* m, int32_t* status);

/* ---- the public per-evaluation helpers of SingleCameraTriangulator, one by one ----
 * evaluateNormal (normaloptimizer.cpp:65-149) is built from four public methods; the normal-search
 * kernels fuse them, these entry points keep them callable (element-wise kernels; host pointers).
 * `info` mirrors the reference's return value: 0, -1 (a point left the bounding box / a pixel is
 * not good), -6 (NaN plane point: the reference exit(-6)s, singlecameratriangulator.cpp:465-469). */

/* SingleCameraTriangulator::extractPixelsContour(Vec3d) (singlecameratriangulator.cpp:341-397): the
 * disc of image-1 pixels around the projection of P, in the reference's order (x offset outer loop),
 * clipped to the image given to fm3d_set_images (the reference hard-codes 1024x768).  *m = number of
 * pixels; at most `cap` are written to xy (m x 2). */
int fm3d_disc_pixels(fm3d_ctx* ctx, const double P[3], int pixels_ray, double* xy, int cap, int* m);

/* ::get3dPointsFromImage1Pixels (:530-565) with projectPointToPlane (:421-470) and isInBoundingBox
 * (:646-655): undistort every pixel, intersect its ray with the plane (P, normal).  All m points are
 * written (the reference stops at the first one outside the box). */
int fm3d_plane_points(fm3d_ctx* ctx, const double P[3], const double normal[3], const double* xy, int m,
                      double* xyz, int* info);

/* ::updateImage1PixelsIntensity (:576-589) for image = 1; the same sampler on image 2.  Samples
 * pyramid level `level` of the image at (float)(scale x), (float)(scale y) with
 * getBilinearInterpPix32f (tools.cpp:129-142); gate != 0 applies isPixelGood (:657-665). */
int fm3d_sample_pixels(fm3d_ctx* ctx, int image, int level, double scale, int gate, const double* xy, int m,
                       float* intensity, int* info);

/* ::projectPointsToImage2 (:591-632): cv::projectPoints with g12, isPixelGood, bilinear sample of
 * image 2 at pyramid level `level`.  intensity may be NULL (projection only, no gate: the first half of
 * projectPointsToImages, :232-276). */
int fm3d_project_to_image2(fm3d_ctx* ctx, const double* xyz, int m, int level, double scale, double* xy2,
                           float* intensity, int* info);

/* Dense search over candidate plane normals (BASELINE configs[4]): evaluates the same cost on a
 * regular n_phi x n_theta grid of (phi, theta) = centre + ((i - (n_phi-1)/2) dphi, (j - (n_theta-1)/2) dtheta),
 * candidate index c = i * n_theta + j, at pyramid level `level`.  The centre is center_phi_theta[f]
 * (n x 2), or, when NULL, the optimiser's initial normal P/|P| (normaloptimizer.cpp:343).  One CTA
 * group per feature stages rays, image-1 samples and the image-2 window once and then runs one
 * value-only pass per candidate.
 *   cost       n x n_phi*n_theta  (nullable; NaN where a candidate fails a bounding-box / pixel gate)
 *   best_idx   n   argmin over the grid, -1 if no candidate is valid (nullable)
 *   best_cost  n   (nullable)
 *   status     n   FM3D_FEAT_OK / FM3D_FEAT_NO_PIXELS */
int fm3d_sweep_normals(fm3d_ctx* ctx, const double* xyz, const double* center_phi_theta, int n,
                       int pixels_ray, int level, int penalty_mode, int n_phi, int n_theta,
                       double dphi, double dtheta, double* cost, int32_t* best_idx,
                       double* best_cost, int32_t* status);
int fm3d_sweep_normals_dev(fm3d_ctx* ctx, const double* xyz, const double* center_phi_theta, int n,
                           int pixels_ray, int level, int penalty_mode, int n_phi, int n_theta,
                           double dphi, double dtheta, double* cost, int32_t* best_idx,
                           double* best_cost, int32_t* status);

/* Replaces NormalOptimizer::computeFeaturesFrames (normaloptimizer.cpp:454-505):
 * frame = [x y z P; 0 0 0 1], z = n, x = normalize(g x z), y = normalize(z x x),
 * row-major 4x4 per feature.  gravity as NormalOptimizer::getGravity (:185-188). */
int fm3d_feature_frames(fm3d_ctx* ctx, const double* xyz, const double* normals, int n,
                        const double gravity[3], double* frames);
int fm3d_feature_frames_dev(fm3d_ctx* ctx, const double* xyz, const double* normals, int n,
                            const double gravity[3], double* frames);

/* ---------------------------------------------------------- patch extraction ---- */

/* Patch edge S = 2*floor(epsilon / (0.01*cmPerPixel))
 * (Triangulator/neighborhoodsgenerator.cpp:136-137). */
int fm3d_patch_size(double epsilon_m, double cm_per_pixel);

/* Replaces NeighborhoodsGenerator::getReferenceSquaredNeighborhood
 * (neighborhoodsgenerator.cpp:134-158) + SingleCameraTriangulator::
 * projectReferencePointsToImageWithFrames (singlecameratriangulator.cpp:769-849): for
 * every frame project the S x S metric grid on the feature plane into image 1 (with lens
 * distortion), sample bilinearly and truncate to u8; 0 where the pixel is outside.
 *   patches       n x S x S u8, patch(row=j, col=i) <- grid point (i,j)   (as the reference
 *                 writes patch.at<uchar>(col,row))
 *   image_points  n x S*S x 2 f64 in grid order idx = i*S + j (nullable: 16 B per pixel) */
int fm3d_extract_patches(fm3d_ctx* ctx, const double* frames, int n, double epsilon_m,
                         double cm_per_pixel, uint8_t* patches, double* image_points);
int fm3d_extract_patches_dev(fm3d_ctx* ctx, const double* frames, int n, double epsilon_m,
                             double cm_per_pixel, uint8_t* patches, double* image_points);

/* Replaces SingleCameraTriangulator::projectPointsToImage
 * (singlecameratriangulator.cpp:667-767) for explicit 3-D groups (n groups of S*S points,
 * camera-1 coordinates) and either image (image = 1 or 2). */
int fm3d_project_groups(fm3d_ctx* ctx, int image, const double* groups, int n, int S,
                        uint8_t* patches, double* image_points);

/* Replaces NeighborhoodsGenerator::computeSquareNeighborhoodsByNormals
 * (neighborhoodsgenerator.cpp:76-132): out n x S*S x 3 = frame * (grid point, 1). */
int fm3d_square_neighborhoods(fm3d_ctx* ctx, const double* frames, int n, double epsilon_m,
                              double cm_per_pixel, double* out);

/* Replaces NeighborhoodsGenerator::computeCircularNeighborhood(s)ByNormal(s)
 * (neighborhoodsgenerator.cpp:160-277, look-up table of the constructor :46-66): n_rays concentric
 * circles of n_angles samples each on the plane through points[f] with normal normals[f]; sample
 * k = (ray - 1) * n_angles + angle.  An all-zero normal means "initial guess": it is replaced by
 * P/|P| and written back.  out: n x n_rays*n_angles x 3.  (Settings: Neighborhoods.epsilon, thetas, rays.) */
int fm3d_circular_neighborhoods(fm3d_ctx* ctx, const double* points, double* normals, int n, double epsilon_m,
                                int n_angles, int n_rays, double* out);

/* ---------------------------------------------------------- patch descriptors ---- */

/* Replaces DescriptorsMatcher::extractDescriptorsFromPatches
 * (DescriptorsMatcher/descriptorsmatcher.cpp:133-174) for ExtractorType SIFT (:302-314): one keypoint
 * per patch at (floor(S/2), floor(S/2)) with size = S, angle = -1, octave = 0, response = 1,
 * described by cv::SIFT::compute (nOctaveLayers 3, sigma 1.6).  patches: n x S x S u8 as written by
 * fm3d_extract_patches; descriptors: n x 128 f32, integer-valued in [0, 255] (cv::SIFT's CV_32F
 * output), row k = patch k (the reference copies descriptorsVector[k] into row k, :168-172).
 * 8 <= S <= 160 (two padded S x S float planes in shared memory). */
int fm3d_describe_patches_sift(fm3d_ctx* ctx, const uint8_t* patches, int n, int S, float* descriptors);
int fm3d_describe_patches_sift_dev(fm3d_ctx* ctx, const uint8_t* patches, int n, int S, float* descriptors);

/* ---------------------------------------------------------- keypoint detection ---- */

/* Replaces feature_detector_->detect(frame, keypoints) of DescriptorsMatcher::compareWithNNDR /
 * compare / crosscompare (DescriptorsMatcher/descriptorsmatcher.cpp:110-111, :92-93, :77-78) for
 * DetectorType FAST, DetectorMode STATIC (:215-222: cv::FastFeatureDetector(Threshold,
 * NonMaxSuppression)): FAST-9-16 corners, keypoints in row-major order (ascending y, then x).
 *   img            h rows of w u8 pixels, `stride` bytes apart
 *   threshold      FeatureOptions.FastDetector.Threshold (clamped to [0, 255] as cv::FAST does)
 *   nonmax         FeatureOptions.FastDetector.NonMaxSuppression != 0: keep a corner only if its
 *                  cornerScore is strictly above its 8 neighbours'; response = score (else 0)
 *   xy             max_keypoints x 2 f32 (KeyPoint::pt), response max_keypoints f32
 *   n              total number of corners found; when n > max_keypoints only the first
 *                  max_keypoints (row-major order) were written.  max_keypoints = 0 only counts.
 * KeyPoint::size is 7, angle -1, octave 0, class_id -1 for every keypoint (constant, not returned). */
int fm3d_detect_fast(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int threshold,
                     int nonmax, int max_keypoints, float* xy, float* response, int* n);
int fm3d_detect_fast_dev(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int threshold,
                         int nonmax, int max_keypoints, float* xy, float* response, int* n_dev);

/* Replaces feature_detector_->detect(frame, keypoints) (DescriptorsMatcher/descriptorsmatcher.cpp:110-111, :91-92, :76-77) for
 * DetectorType SIFT (:243-256: cv::SIFT(NumFeatures, NumOctaveLayers, ContrastThreshold, EdgeThreshold, Sigma); settings.yml
 * FeatureOptions.SiftDetector): cv::SIFT's scale-space detector -- doubled base image, Gaussian / DoG pyramids, 26-neighbour
 * extrema, sub-pixel refinement with the contrast and edge tests, one keypoint per orientation peak -- followed by
 * KeyPointsFilter::removeDuplicatedSorted (the output order) and retainBest(nfeatures) (nfeatures <= 0: all; the kept SET is
 * OpenCV's, their order stays the sorted one).  img: host u8, stride bytes per row.  Outputs (host, up to max_keypoints rows):
 * xy (x, y pairs), size, angle (degrees), response, octave (cv::KeyPoint::octave as cv::SIFT packs it).  *n = keypoints
 * found (may exceed max_keypoints; only max_keypoints are written). */
int fm3d_detect_sift(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int nfeatures, int n_octave_layers,
                     double contrast_threshold, double edge_threshold, double sigma, int max_keypoints, float* xy, float* size,
                     float* angle, float* response, int32_t* octave, int* n);

/* Replaces feature_detector_->detect followed by descriptor_extractor_->compute on the same frame
 * (DescriptorsMatcher/descriptorsmatcher.cpp:110-115, :91-96, :76-81) for DetectorType SIFT + ExtractorType SIFT with the same
 * settings (:243-256, :302-314): fm3d_detect_sift, then the descriptors of its keypoints read from THE PYRAMID THEY WERE FOUND
 * ON -- the images cv::SIFT::compute would build a second time (doubled frame, every octave) -- i.e. what
 * cv::SIFT::detectAndCompute does.  Keypoints and descriptors are bit-identical to fm3d_detect_sift followed by
 * fm3d_describe_keypoints_sift_oct.  descriptors: max_keypoints x 128 f32; when *n > max_keypoints nothing is described
 * (call again with room for *n). */
int fm3d_detect_and_describe_sift(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int nfeatures, int n_octave_layers,
                                  double contrast_threshold, double edge_threshold, double sigma, int max_keypoints, float* xy, float* size,
                                  float* angle, float* response, int32_t* octave, int* n, float* descriptors);

/* Replaces feature_detector_->detect followed by descriptor_extractor_->compute (DescriptorsMatcher/descriptorsmatcher.cpp:110-115,
 * :91-96, :76-81) for DetectorType ORB + ExtractorType ORB (:273-279, :336-342: cv::ORB(OrbDetector.NumFeatures, ScaleFactor,
 * NumLevels), the other arguments at cv::ORB's defaults; fast_threshold = 20 is its default): the INTER_LINEAR_EXACT pyramid,
 * FAST + Harris ranking + intensity-centroid angles per level, rBRIEF rows on the blurred level images -- the keypoint SET and
 * the rows of cv::ORB::detectAndCompute; the order is (level, y, x) (OpenCV's is whatever std::nth_element leaves).  Outputs
 * (host, up to max_keypoints rows): xy in frame coordinates, size = 31 s^level, angle (degrees), response (Harris measure),
 * octave = level, descriptors (32 bytes per keypoint; may be NULL).  *n = keypoints found. */
int fm3d_detect_orb(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int nfeatures, double scale_factor, int nlevels,
                    int fast_threshold, int max_keypoints, float* xy, float* size, float* angle, float* response, int32_t* octave,
                    uint8_t* descriptors, int* n);

/* ---------------------------------------------------------- keypoint description ---- */

/* Replaces descriptor_extractor_->compute(frame, keypoints, descriptors) of
 * DescriptorsMatcher::compareWithNNDR / compare / crosscompare
 * (DescriptorsMatcher/descriptorsmatcher.cpp:114-115, :95-96, :80-81) for ExtractorType SIFT (:302-314,
 * default octave layers 3 / sigma 1.6) and keypoints of octave 0 -- what DetectorType FAST produces
 * (:215-222: size 7, angle -1) -- for which cv::SIFT::compute reads every descriptor from
 * GaussianBlur(float(frame), sqrt(1.6^2 - 0.5^2)) without building a scale space.
 *   img            h rows of w u8 pixels, `stride` bytes apart
 *   kps            n x 4 f32: KeyPoint::pt.x, pt.y, size, angle (angle -1 = "not set", as FAST leaves it)
 *   descriptors    n x 128 f32, integer-valued in [0, 255] (cv::SIFT's CV_32F rows), row k = keypoint k
 * A keypoint outside the image or of size <= FLT_EPSILON (DescriptorExtractor::compute removes those before
 * the extractor runs; the adapters do the same) gets an all-zero row. */
int fm3d_describe_keypoints_sift(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps,
                                 int n, float* descriptors);
int fm3d_describe_keypoints_sift_dev(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride,
                                     const float* kps, int n, float* descriptors);

/* The same call for keypoints that carry an octave -- what cv::SIFT's own detector (fm3d_detect_sift) returns: cv::SIFT::compute
 * with provided keypoints builds the Gaussian pyramid over the octave range of the keypoints (the doubled frame first if one
 * of them has octave -1) and reads every descriptor from the image of the keypoint's octave and layer, at pt * scale with
 * size * scale (calcDescriptors / unpackOctave).  kps: n x 4 (x, y, size, angle) in frame coordinates, octaves: n packed
 * cv::KeyPoint::octave values.  n_octave_layers / sigma: FeatureOptions.SiftDetector.NumOctaveLayers / Sigma. */
int fm3d_describe_keypoints_sift_oct(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps,
                                     const int32_t* octaves, int n, int n_octave_layers, double sigma, float* descriptors);
/* The base image alone (createInitialImage): base is w x h f32, rows w floats apart, device memory. */
int fm3d_sift_base_image_dev(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, float* base);

/* The same call for ExtractorType BRISK (:343-349: cv::BRISK(BriskDetector.Threshold, BriskDetector.Octaves) -- both
 * knobs only steer BRISK's own detector, which the reference does not run): the binary descriptors that the Hamming
 * matcher consumes (:64-67).
 *   kps                  n x 4 f32 as above
 *   compute_orientation  != 0: orientation from the long pairs, as cv::BRISK::compute of OpenCV >= 3 does for provided
 *                        keypoints (pinned against cv2 4.13); 0: the keypoint's own angle is used and angle -1 means
 *                        "unrotated" (OpenCV 2.4's behaviour for provided keypoints)
 *   descriptors          n x 64 u8; kept n u8; angles n f32 (KeyPoint::angle after the call, degrees in [0, 360))
 * cv::BRISK REMOVES keypoints whose pattern would leave the image (and their rows): here kept[k] = 0 and row k is
 * zero; the adapters erase those keypoints as the reference's call would.  w * h * 255 must fit an int32 (the integral
 * image is CV_32S as in cv::integral). */
int fm3d_describe_keypoints_brisk(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps,
                                  int n, int compute_orientation, uint8_t* descriptors, uint8_t* kept, float* angles);
int fm3d_describe_keypoints_brisk_dev(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride,
                                      const float* kps, int n, int compute_orientation, uint8_t* descriptors,
                                      uint8_t* kept, float* angles);

/* The same call for ExtractorType ORB (:336-342: cv::ORB(OrbDetector.NumFeatures, ScaleFactor, NumLevels); the knobs
 * steer ORB's own detector and pyramid, which keypoints of octave 0 do not touch): 256-bit rBRIEF rows.
 *   kps            n x 4 f32 as above (octave 0; angle in degrees -- FAST's -1 is a rotation by -1 degree, as in cv::ORB)
 *   descriptors    n x 32 u8; kept n u8
 * cv::ORB REMOVES keypoints whose rounded position is within 31 pixels of the border (and their rows): here
 * kept[k] = 0 and row k is zero; the adapters erase those keypoints as the reference's call would. */
int fm3d_describe_keypoints_orb(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps, int n,
                                uint8_t* descriptors, uint8_t* kept);
int fm3d_describe_keypoints_orb_dev(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps,
                                    int n, uint8_t* descriptors, uint8_t* kept);

/* DescriptorsMatcher::extractDescriptorsFromPatches (:133-174) with ExtractorType ORB: one keypoint per patch at
 * (floor(S/2), floor(S/2)), size S, angle -1; descriptors n x 32 u8.  cv::ORB keeps that keypoint only when
 * 31 <= S/2 < S - 31 (FM3D_ERR_UNSUPPORTED otherwise: the reference's own call would return an empty row). */
int fm3d_describe_patches_orb(fm3d_ctx* ctx, const uint8_t* patches, int n, int S, uint8_t* descriptors);
int fm3d_describe_patches_orb_dev(fm3d_ctx* ctx, const uint8_t* patches, int n, int S, uint8_t* descriptors);

#ifdef __cplusplus
}
#endif
#endif /* FM3D_H_ */
