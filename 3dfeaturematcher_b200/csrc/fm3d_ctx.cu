// fm3d_ctx.cu -- context, options, camera state and small host-side helpers of libfm3d.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "fm3d_internal.cuh"

int fm3d_fail(fm3d_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

int fm3d_bind(fm3d_ctx* ctx) {
    FM3D_CUDA(ctx, cudaSetDevice(ctx->device));
    return FM3D_OK;
}

int fm3d_scratch(fm3d_ctx* ctx, int slot, size_t bytes, void** out) {
    if (bytes == 0) bytes = 16;
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) {
            FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            FM3D_CUDA(ctx, cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_bytes[slot] = 0;
        }
        size_t cap = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&ctx->scratch[slot], cap);
        if (e != cudaSuccess)
            return fm3d_fail(ctx, FM3D_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
        ctx->scratch_bytes[slot] = cap;
    }
    *out = ctx->scratch[slot];
    return FM3D_OK;
}

int fm3d_pinned(fm3d_ctx* ctx, size_t bytes, void** out) {
    if (bytes == 0) bytes = 16;
    if (ctx->pinned_bytes < bytes) {
        if (ctx->pinned) {
            FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            FM3D_CUDA(ctx, cudaFreeHost(ctx->pinned));
            ctx->pinned = nullptr;
            ctx->pinned_bytes = 0;
        }
        cudaError_t e = cudaMallocHost(&ctx->pinned, bytes);
        if (e != cudaSuccess)
            return fm3d_fail(ctx, FM3D_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
        ctx->pinned_bytes = bytes;
    }
    *out = ctx->pinned;
    return FM3D_OK;
}

int fm3d_h2d(fm3d_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return FM3D_OK;
    ctx->n_copy++;
    FM3D_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return FM3D_OK;
}

int fm3d_d2h(fm3d_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return FM3D_OK;
    ctx->n_copy++;
    FM3D_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return FM3D_OK;
}

typedef CUresult (*fm3d_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int fm3d_encode_tmap_2d_u8(fm3d_ctx* ctx, CUtensorMap* map, const void* base, int w, int h,
                           int pitch, int box_w, int box_h) {
    static fm3d_encode_fn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        FM3D_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (!p || qres != cudaDriverEntryPointSuccess)
            return fm3d_fail(ctx, FM3D_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
        fn = (fm3d_encode_fn)p;
    }
    cuuint64_t dims[2] = {(cuuint64_t)w, (cuuint64_t)h};
    cuuint64_t strides[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fm3d_fail(ctx, FM3D_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) w=%d h=%d pitch=%d box=%dx%d",
                         (int)r, w, h, pitch, box_w, box_h);
    return FM3D_OK;
}

extern "C" {

int fm3d_version(void) { return FM3D_VERSION; }

int fm3d_ctx_create(int device, fm3d_ctx** out) {
    if (!out) return FM3D_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return FM3D_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return FM3D_ERR_NO_DEVICE;
    fm3d_ctx* ctx = new fm3d_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess) {
        delete ctx;
        return FM3D_ERR_NO_DEVICE;
    }
    if (ctx->prop.major != 10) {  // the fatbin carries sm_100a SASS only: no other GPU can run it
        delete ctx;
        return FM3D_ERR_NO_DEVICE;
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return FM3D_ERR_CUDA;
    }
    *out = ctx;
    return FM3D_OK;
}

void fm3d_ctx_destroy(fm3d_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    fm3d_comm_destroy(ctx);
    for (int i = 0; i < FM3D_SCRATCH_SLOTS; i++)
        if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
    if (ctx->pyr_mem) cudaFree(ctx->pyr_mem);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* fm3d_last_error(const fm3d_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int fm3d_sync(fm3d_ctx* ctx) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->normals_flag_pending && ctx->scratch[1]) {
        // the device-resident normal search reports a timed-out TMA window load here (the host-buffer entry point
        // reports it itself): the results are valid, taken through the global-memory sampler, but slow
        ctx->normals_flag_pending = false;
        int flags[2] = {0, 0};
        FM3D_CUDA(ctx, cudaMemcpy(flags, ctx->scratch[1], sizeof(flags), cudaMemcpyDeviceToHost));
        if (flags[1]) return fm3d_fail(ctx, FM3D_ERR_CUDA, "normal optimiser: a TMA window load timed out (results used the global-memory path)");
    }
    return FM3D_OK;
}

void* fm3d_stream(fm3d_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// Device memory for callers of the _dev entry points that have no CUDA runtime of their own (the C++ class adapters).
int fm3d_dev_malloc(fm3d_ctx* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return FM3D_ERR_INVALID_ARG;
    *out = nullptr;
    if (int rc = fm3d_bind(ctx)) return rc;
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess) return fm3d_fail(ctx, FM3D_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return FM3D_OK;
}

int fm3d_dev_free(fm3d_ctx* ctx, void* p) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    if (!p) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FM3D_CUDA(ctx, cudaFree(p));
    return FM3D_OK;
}

int fm3d_copy_h2d(fm3d_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, bytes == 0 || (dst_dev && src_host));
    if (int rc = fm3d_bind(ctx)) return rc;
    if (int rc = fm3d_h2d(ctx, dst_dev, src_host, bytes)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));      // the host buffer may be reused on return
    return FM3D_OK;
}

int fm3d_copy_d2h(fm3d_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, bytes == 0 || (dst_host && src_dev));
    if (int rc = fm3d_bind(ctx)) return rc;
    if (int rc = fm3d_d2h(ctx, dst_host, src_dev, bytes)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

int fm3d_device_info(fm3d_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, char* name,
                     int name_len) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    if (name && name_len > 0) {
        strncpy(name, ctx->prop.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    return FM3D_OK;
}

static int* fm3d_option_slot(fm3d_ctx* ctx, const char* key) {
    if (!strcmp(key, "geometry_f32")) return &ctx->opt_geometry_f32;
    if (!strcmp(key, "matcher_tensor")) return &ctx->opt_matcher_tensor;
    if (!strcmp(key, "matcher_sp_tile")) return &ctx->opt_matcher_sp_tile;
    if (!strcmp(key, "matcher_splits")) return &ctx->opt_matcher_splits;
    if (!strcmp(key, "matcher_persistent")) return &ctx->opt_matcher_persistent;
    if (!strcmp(key, "matcher_min_tiles")) return &ctx->opt_matcher_min_tiles;
    if (!strcmp(key, "lm_patience")) return &ctx->opt_lm_patience;
    if (!strcmp(key, "normals_threads")) return &ctx->opt_normals_threads;
    if (!strcmp(key, "normals_tma")) return &ctx->opt_normals_tma;
    if (!strcmp(key, "normals_fast")) return &ctx->opt_normals_fast;
    if (!strcmp(key, "normals_cost")) return &ctx->opt_normals_cost;
    if (!strcmp(key, "pyramid_fused")) return &ctx->opt_pyramid_fused;
    if (!strcmp(key, "normals_fuse")) return &ctx->opt_normals_fuse;
    if (!strcmp(key, "normals_memo")) return &ctx->opt_normals_memo;
    if (!strcmp(key, "normals_groups")) return &ctx->opt_normals_groups;
    if (!strcmp(key, "normals_pingpong")) return &ctx->opt_normals_pingpong;
    if (!strcmp(key, "normals_level_sync")) return &ctx->opt_normals_level_sync;
    if (!strcmp(key, "normals_sweep_batch")) return &ctx->opt_normals_sweep_batch;
    if (!strcmp(key, "matcher_exact_fallback")) return &ctx->n_matcher_exact_fallback;   // read-out of the last float match
    return nullptr;
}

int fm3d_set_option(fm3d_ctx* ctx, const char* key, double value) {
    if (!ctx || !key) return FM3D_ERR_INVALID_ARG;
    int* s = fm3d_option_slot(ctx, key);
    if (!s) return fm3d_fail(ctx, FM3D_ERR_INVALID_ARG, "unknown option '%s'", key);
    *s = (int)value;
    return FM3D_OK;
}

int fm3d_get_option(fm3d_ctx* ctx, const char* key, double* value) {
    if (!ctx || !key || !value) return FM3D_ERR_INVALID_ARG;
    int* s = fm3d_option_slot(ctx, key);
    if (!s) return fm3d_fail(ctx, FM3D_ERR_INVALID_ARG, "unknown option '%s'", key);
    *value = (double)*s;
    return FM3D_OK;
}

int fm3d_get_launch_counters(fm3d_ctx* ctx, int64_t* kernel_launches, int64_t* copies) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    if (kernel_launches) *kernel_launches = ctx->n_launch;
    if (copies) *copies = ctx->n_copy;
    return FM3D_OK;
}

int fm3d_set_camera(fm3d_ctx* ctx, const double K[9], const double dist[5], double z_min,
                    double z_max) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, K && dist);
    FM3D_CHECK_ARG(ctx, K[0] != 0.0 && K[4] != 0.0);
    ctx->cam.fx = K[0]; ctx->cam.fy = K[4]; ctx->cam.cx = K[2]; ctx->cam.cy = K[5];
    ctx->cam.ifx = 1.0 / K[0]; ctx->cam.ify = 1.0 / K[4];
    ctx->cam.k1 = dist[0]; ctx->cam.k2 = dist[1]; ctx->cam.p1 = dist[2]; ctx->cam.p2 = dist[3];
    ctx->cam.k3 = dist[4];
    ctx->cam.zmin = z_min; ctx->cam.zmax = z_max;
    ctx->has_cam = true;
    return FM3D_OK;
}

int fm3d_set_g12(fm3d_ctx* ctx, const double g12[16]) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, g12 != nullptr);
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) ctx->cam.R[i * 3 + j] = g12[i * 4 + j];
        ctx->cam.t[i] = g12[i * 4 + 3];
    }
    ctx->has_g12 = true;
    return FM3D_OK;
}

// ---- host arithmetic for setg12 (cv::Rodrigues vector -> matrix, 4x4 inverse and products)
static void rodrigues_to_matrix(const double r[3], double R[9]) {
    double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (th < 2.220446049250313e-16) {
        for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    double c = cos(th), s = sin(th), c1 = 1.0 - c, it = 1.0 / th;
    double k[3] = {r[0] * it, r[1] * it, r[2] * it};
    double rrt[9] = {k[0] * k[0], k[0] * k[1], k[0] * k[2], k[0] * k[1], k[1] * k[1],
                     k[1] * k[2], k[0] * k[2], k[1] * k[2], k[2] * k[2]};
    double rx[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
    for (int i = 0; i < 9; i++) R[i] = c * ((i % 4 == 0) ? 1.0 : 0.0) + c1 * rrt[i] + s * rx[i];
}

static void compose4(const double R[9], const double t[3], double G[16]) {
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) G[i * 4 + j] = R[i * 3 + j];
        G[i * 4 + 3] = t[i];
    }
    G[12] = G[13] = G[14] = 0.0;
    G[15] = 1.0;
}

static int inverse4(const double A[16], double Ai[16]) {  // Gauss-Jordan with partial pivoting
    double m[4][8];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) { m[i][j] = A[i * 4 + j]; m[i][j + 4] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; c++) {
        int p = c;
        for (int r = c + 1; r < 4; r++) if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
        if (m[p][c] == 0.0) return -1;
        if (p != c) for (int j = 0; j < 8; j++) { double t = m[c][j]; m[c][j] = m[p][j]; m[p][j] = t; }
        double d = 1.0 / m[c][c];
        for (int j = 0; j < 8; j++) m[c][j] *= d;
        for (int r = 0; r < 4; r++) {
            if (r == c) continue;
            double f = m[r][c];
            if (f != 0.0) for (int j = 0; j < 8; j++) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Ai[i * 4 + j] = m[i][j + 4];
    return 0;
}

static void mul4(const double A[16], const double B[16], double C[16]) {
    double T[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0;
            for (int k = 0; k < 4; k++) s += A[i * 4 + k] * B[k * 4 + j];
            T[i * 4 + j] = s;
        }
    memcpy(C, T, sizeof(T));
}

int fm3d_compose_g12(const double T1[3], const double T2[3], const double rod1[3],
                     const double rod2[3], const double rodIC[3], const double tIC[3],
                     double g12_out[16]) {
    if (!T1 || !T2 || !rod1 || !rod2 || !rodIC || !tIC || !g12_out) return FM3D_ERR_INVALID_ARG;
    double R1[9], R2[9], RIC[9], g1[16], g2[16], gIC[16], gICi[16], g2i[16], tmp[16];
    rodrigues_to_matrix(rod1, R1);
    rodrigues_to_matrix(rod2, R2);
    rodrigues_to_matrix(rodIC, RIC);
    compose4(R1, T1, g1);
    compose4(R2, T2, g2);
    compose4(RIC, tIC, gIC);
    if (inverse4(gIC, gICi) || inverse4(g2, g2i)) return FM3D_ERR_INVALID_ARG;
    mul4(gICi, g2i, tmp);   // g_IC^-1 * g2^-1
    mul4(tmp, g1, tmp);     //   * g1
    mul4(tmp, gIC, g12_out);//   * g_IC        (singlecameratriangulator.cpp:140)
    return FM3D_OK;
}

}  // extern "C"
