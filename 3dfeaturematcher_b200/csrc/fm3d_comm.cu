// fm3d_comm.cu -- several GPUs behind the C-ABI.
//
// Every stage of the path is independent per query keypoint / per feature, so the only exchanges are the two the north
// star names: replicate the train-descriptor set and the two frames (one ncclBroadcast of ONE packed buffer), and collect
// the per-shard matches and normals (one ncclAllGather of fixed-size blocks whose header and rows are written by a kernel
// from device-resident counts: nothing of a step touches the host).  No collective sits inside a compute kernel.
//
// Two deployments:
//   one process per GPU   fm3d_comm_unique_id on rank 0, the 128 bytes reach the other ranks by any channel the caller has
//                         (bench.py: torch.distributed), fm3d_comm_init_rank everywhere;
//   one process, n GPUs   fm3d_comm_init_all over the n contexts of the process (ncclCommInitAll), collectives through the
//                         *_all_dev entry points, which put the n per-context calls into one NCCL group.
//
// NCCL is loaded at run time (dlopen libnccl.so.2): libfm3d.so has no link-time dependency on it and a single-GPU user
// never needs it.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "fm3d_internal.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string why;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return; }
#define FM3D_NCCL_SYM(field, name)                                                                 \
        api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));                \
        if (!api.field) { api.why = std::string("NCCL symbol missing: ") + name; api.handle = nullptr; return; }
        FM3D_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        FM3D_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        FM3D_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        FM3D_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        FM3D_NCCL_SYM(Broadcast, "ncclBroadcast")
        FM3D_NCCL_SYM(AllGather, "ncclAllGather")
        FM3D_NCCL_SYM(GroupStart, "ncclGroupStart")
        FM3D_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        FM3D_NCCL_SYM(GetErrorString, "ncclGetErrorString")
        FM3D_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef FM3D_NCCL_SYM
    });
    return api.handle ? &api : nullptr;
}

int nccl_fail(fm3d_ctx* ctx, const char* what, ncclResult_t r) {
    NcclApi* a = nccl_api();
    return fm3d_fail(ctx, FM3D_ERR_COMM, "%s failed: %s", what, a ? a->GetErrorString(r) : "NCCL not loaded");
}

#define FM3D_NCCL(ctx, expr)                                                    \
    do {                                                                        \
        ncclResult_t r__ = (expr);                                              \
        if (r__ != ncclSuccess) return nccl_fail((ctx), #expr, r__);            \
    } while (0)

int need_comm(fm3d_ctx* ctx) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    if (!ctx->comm) return fm3d_fail(ctx, FM3D_ERR_STATE, "no communicator: call fm3d_comm_init_rank / fm3d_comm_init_all first");
    return FM3D_OK;
}

// One thread block per part: header + rows of a shard block (layout: fm3d_shard_layout in fm3d.h).
__global__ void __launch_bounds__(256)
pack_shard_kernel(fm3d_shard_layout L, int rank, const int32_t* __restrict__ n_match, const int32_t* __restrict__ n_inl,
                  const int32_t* __restrict__ qidx, const int32_t* __restrict__ tidx, const float* __restrict__ dist,
                  const int32_t* __restrict__ src, const double* __restrict__ normals, const int32_t* __restrict__ status,
                  int query_offset, uint8_t* __restrict__ block) {
    const int nm = min(max(*n_match, 0), L.cap), ni = min(max(*n_inl, 0), L.cap);
    const int part = blockIdx.y;
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    if (part == 0 && i0 == 0) {
        int32_t* h = reinterpret_cast<int32_t*>(block);
        h[0] = nm; h[1] = ni; h[2] = rank; h[3] = L.cap; h[4] = query_offset;
    }
    // rows beyond the counts are zero: the block is fully defined whatever the previous step left in it
    if (part == 0) { int32_t* o = reinterpret_cast<int32_t*>(block + L.off_qidx); for (int i = i0; i < L.cap; i += stride) o[i] = i < nm ? qidx[i] + query_offset : 0; }
    else if (part == 1) { int32_t* o = reinterpret_cast<int32_t*>(block + L.off_tidx); for (int i = i0; i < L.cap; i += stride) o[i] = i < nm ? tidx[i] : 0; }
    else if (part == 2) { float* o = reinterpret_cast<float*>(block + L.off_dist); for (int i = i0; i < L.cap; i += stride) o[i] = i < nm ? dist[i] : 0.f; }
    else if (part == 3) { int32_t* o = reinterpret_cast<int32_t*>(block + L.off_src); for (int i = i0; i < L.cap; i += stride) o[i] = i < ni ? src[i] : 0; }
    else if (part == 4) { double* o = reinterpret_cast<double*>(block + L.off_normals); for (int i = i0; i < 3 * L.cap; i += stride) o[i] = i < 3 * ni ? normals[i] : 0.0; }
    else { int32_t* o = reinterpret_cast<int32_t*>(block + L.off_status); for (int i = i0; i < L.cap; i += stride) o[i] = i < ni ? status[i] : 0; }
}

}  // namespace

extern "C" {

int fm3d_comm_unique_id(uint8_t id[FM3D_COMM_ID_BYTES]) {
    if (!id) return FM3D_ERR_INVALID_ARG;
    NcclApi* a = nccl_api();
    if (!a) return FM3D_ERR_COMM;
    static_assert(sizeof(ncclUniqueId) == FM3D_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    if (a->GetUniqueId(&u) != ncclSuccess) return FM3D_ERR_COMM;
    memcpy(id, &u, sizeof(u));
    return FM3D_OK;
}

int fm3d_comm_init_rank(fm3d_ctx* ctx, const uint8_t id[FM3D_COMM_ID_BYTES], int nranks, int rank) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, id && nranks >= 1 && rank >= 0 && rank < nranks);
    if (ctx->comm) return fm3d_fail(ctx, FM3D_ERR_STATE, "the context already has a communicator");
    NcclApi* a = nccl_api();
    if (!a) return fm3d_fail(ctx, FM3D_ERR_COMM, "NCCL unavailable");
    if (int rc = fm3d_bind(ctx)) return rc;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t c = nullptr;
    FM3D_NCCL(ctx, a->CommInitRank(&c, nranks, u, rank));
    ctx->comm = c; ctx->comm_nranks = nranks; ctx->comm_rank = rank;
    return FM3D_OK;
}

int fm3d_comm_init_all(fm3d_ctx** ctxs, int n) {
    if (!ctxs || n < 1) return FM3D_ERR_INVALID_ARG;
    for (int k = 0; k < n; k++) if (!ctxs[k]) return FM3D_ERR_INVALID_ARG;
    fm3d_ctx* c0 = ctxs[0];
    for (int k = 0; k < n; k++) if (ctxs[k]->comm) return fm3d_fail(c0, FM3D_ERR_STATE, "context %d already has a communicator", k);
    NcclApi* a = nccl_api();
    if (!a) return fm3d_fail(c0, FM3D_ERR_COMM, "NCCL unavailable");
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n, nullptr);
    for (int k = 0; k < n; k++) devs[k] = ctxs[k]->device;
    FM3D_NCCL(c0, a->CommInitAll(comms.data(), n, devs.data()));
    for (int k = 0; k < n; k++) { ctxs[k]->comm = comms[k]; ctxs[k]->comm_nranks = n; ctxs[k]->comm_rank = k; }
    return FM3D_OK;
}

int fm3d_comm_info(fm3d_ctx* ctx, int* nranks, int* rank, int* nccl_version) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    if (nranks) *nranks = ctx->comm ? ctx->comm_nranks : 1;
    if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
    if (nccl_version) {
        *nccl_version = 0;
        if (NcclApi* a = nccl_api()) a->GetVersion(nccl_version);
    }
    return FM3D_OK;
}

int fm3d_comm_destroy(fm3d_ctx* ctx) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    if (!ctx->comm) return FM3D_OK;
    NcclApi* a = nccl_api();
    if (int rc = fm3d_bind(ctx)) return rc;
    cudaStreamSynchronize(ctx->stream);
    if (a) a->CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr; ctx->comm_nranks = 1; ctx->comm_rank = 0;
    return FM3D_OK;
}

int fm3d_broadcast_dev(fm3d_ctx* ctx, void* buf, size_t bytes, int root) {
    if (int rc = need_comm(ctx)) return rc;
    FM3D_CHECK_ARG(ctx, buf && root >= 0 && root < ctx->comm_nranks);
    if (bytes == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    ctx->n_coll++;
    FM3D_NCCL(ctx, nccl_api()->Broadcast(buf, buf, bytes, ncclUint8, root, (ncclComm_t)ctx->comm, ctx->stream));
    return FM3D_OK;
}

int fm3d_allgather_dev(fm3d_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
    if (int rc = need_comm(ctx)) return rc;
    FM3D_CHECK_ARG(ctx, send && recv);
    if (bytes_per_rank == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    ctx->n_coll++;
    FM3D_NCCL(ctx, nccl_api()->AllGather(send, recv, bytes_per_rank, ncclUint8, (ncclComm_t)ctx->comm, ctx->stream));
    return FM3D_OK;
}

int fm3d_broadcast_all_dev(fm3d_ctx** ctxs, int n, void* const* bufs, size_t bytes, int root) {
    if (!ctxs || !bufs || n < 1) return FM3D_ERR_INVALID_ARG;
    for (int k = 0; k < n; k++) if (int rc = need_comm(ctxs[k])) return rc;
    NcclApi* a = nccl_api();
    FM3D_NCCL(ctxs[0], a->GroupStart());
    for (int k = 0; k < n; k++) {
        cudaSetDevice(ctxs[k]->device);
        ctxs[k]->n_coll++;
        ncclResult_t r = a->Broadcast(bufs[k], bufs[k], bytes, ncclUint8, root, (ncclComm_t)ctxs[k]->comm, ctxs[k]->stream);
        if (r != ncclSuccess) { a->GroupEnd(); return nccl_fail(ctxs[k], "ncclBroadcast", r); }
    }
    FM3D_NCCL(ctxs[0], a->GroupEnd());
    return FM3D_OK;
}

int fm3d_allgather_all_dev(fm3d_ctx** ctxs, int n, const void* const* send, void* const* recv, size_t bytes_per_rank) {
    if (!ctxs || !send || !recv || n < 1) return FM3D_ERR_INVALID_ARG;
    for (int k = 0; k < n; k++) if (int rc = need_comm(ctxs[k])) return rc;
    NcclApi* a = nccl_api();
    FM3D_NCCL(ctxs[0], a->GroupStart());
    for (int k = 0; k < n; k++) {
        cudaSetDevice(ctxs[k]->device);
        ctxs[k]->n_coll++;
        ncclResult_t r = a->AllGather(send[k], recv[k], bytes_per_rank, ncclUint8, (ncclComm_t)ctxs[k]->comm, ctxs[k]->stream);
        if (r != ncclSuccess) { a->GroupEnd(); return nccl_fail(ctxs[k], "ncclAllGather", r); }
    }
    FM3D_NCCL(ctxs[0], a->GroupEnd());
    return FM3D_OK;
}

int fm3d_shard_block_layout(int cap, fm3d_shard_layout* L) {
    if (!L || cap < 0) return FM3D_ERR_INVALID_ARG;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 256;                                   // header
    L->cap = cap;
    L->off_qidx = off; off = al(off + sizeof(int32_t) * (size_t)cap);
    L->off_tidx = off; off = al(off + sizeof(int32_t) * (size_t)cap);
    L->off_dist = off; off = al(off + sizeof(float) * (size_t)cap);
    L->off_src = off; off = al(off + sizeof(int32_t) * (size_t)cap);
    L->off_normals = off; off = al(off + sizeof(double) * 3 * (size_t)cap);
    L->off_status = off; off = al(off + sizeof(int32_t) * (size_t)cap);
    L->bytes = off;
    return FM3D_OK;
}

int fm3d_pack_shard_dev(fm3d_ctx* ctx, int cap, int rank, int query_offset, const int32_t* n_match_dev, const int32_t* n_inl_dev,
                        const int32_t* qidx, const int32_t* tidx, const float* dist, const int32_t* src_idx,
                        const double* normals, const int32_t* status, void* block) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, cap >= 0 && n_match_dev && n_inl_dev && block && (cap == 0 || (qidx && tidx && dist && src_idx && normals && status)));
    if (int rc = fm3d_bind(ctx)) return rc;
    fm3d_shard_layout L;
    fm3d_shard_block_layout(cap, &L);
    const int bx = cap > 0 ? (cap + 1023) / 1024 : 1;
    pack_shard_kernel<<<dim3(bx < 64 ? bx : 64, 6), 256, 0, ctx->stream>>>(L, rank, n_match_dev, n_inl_dev, qidx, tidx, dist, src_idx, normals, status,
                                                                       query_offset, (uint8_t*)block);
    FM3D_LAUNCH_CHECK(ctx);
    return FM3D_OK;
}

}  // extern "C"
