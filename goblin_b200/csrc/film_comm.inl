// Film::mergeTile across GPUs, inside the library (src/GoblinFilm.cpp:140-153): every worker's full-frame
// (colour, weight) buffer is summed into the film.  Across GPUs that is one all-reduce(sum) of the 4 W H float
// device film over NVLink / NVSwitch (north_star); NCCL is bound at run time (dlopen of libnccl.so.2: the
// library itself does not link against it, so a one-GPU host without NCCL still loads it).
//   * single process, one context per GPU (g_ray --gpus N): gb_comm_init_all + gb_film_allreduce_all;
//   * one process per GPU (torchrun): rank 0 calls gb_comm_unique_id, the launcher's own channel carries the
//     128 bytes to the other ranks, every rank calls gb_comm_init_rank, then gb_film_allreduce per render;
//   * a communicator the caller already owns: gb_comm_attach.
// Included at the end of device.cu (it needs gb_context).
#include <dlfcn.h>
#include <mutex>

#include <nccl.h> // types and prototypes only; no symbol of it is referenced directly

namespace {

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) getUniqueId = nullptr;
    decltype(&ncclCommInitRank) commInitRank = nullptr;
    decltype(&ncclCommInitAll) commInitAll = nullptr;
    decltype(&ncclCommDestroy) commDestroy = nullptr;
    decltype(&ncclAllReduce) allReduce = nullptr;
    decltype(&ncclGroupStart) groupStart = nullptr;
    decltype(&ncclGroupEnd) groupEnd = nullptr;
    decltype(&ncclGetErrorString) getErrorString = nullptr;
    decltype(&ncclGetVersion) getVersion = nullptr;
    std::string error;
};

NcclApi* ncclApi() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        // a process that already carries an NCCL (torch's bundled one) gets that one: same soname
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("NCCL not found: ") + dlerror(); return; }
#define GB_NCCL_SYM(field, name)                                                     \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, #name));     \
    if (!api.field) { api.error = "libnccl lacks " #name; return; }
        GB_NCCL_SYM(getUniqueId, ncclGetUniqueId)
        GB_NCCL_SYM(commInitRank, ncclCommInitRank)
        GB_NCCL_SYM(commInitAll, ncclCommInitAll)
        GB_NCCL_SYM(commDestroy, ncclCommDestroy)
        GB_NCCL_SYM(allReduce, ncclAllReduce)
        GB_NCCL_SYM(groupStart, ncclGroupStart)
        GB_NCCL_SYM(groupEnd, ncclGroupEnd)
        GB_NCCL_SYM(getErrorString, ncclGetErrorString)
        GB_NCCL_SYM(getVersion, ncclGetVersion)
#undef GB_NCCL_SYM
    });
    return &api;
}

#define GB_NCCL(api, call)                                                                              \
    do {                                                                                                \
        ncclResult_t r_ = (call);                                                                       \
        if (r_ != ncclSuccess) return gb::failWith(GB_ERR_CUDA, std::string(#call) + ": " + (api)->getErrorString(r_)); \
    } while (0)

int needNccl(NcclApi** out) {
    NcclApi* api = ncclApi();
    if (!api->error.empty()) return gb::failWith(GB_ERR_STATE, api->error);
    *out = api;
    return GB_OK;
}

int dropComm(gb_context* ctx) {
    if (ctx->comm && ctx->commOwned) {
        NcclApi* api = nullptr;
        if (needNccl(&api) == GB_OK) api->commDestroy(static_cast<ncclComm_t>(ctx->comm));
    }
    ctx->comm = nullptr;
    ctx->commOwned = false;
    ctx->commRanks = 0;
    return GB_OK;
}

} // namespace

extern "C" {

int gb_comm_init_all(gb_context** ctxs, int n) {
    if (!ctxs || n < 1) return gb::failWith(GB_ERR_INVALID, "null argument");
    NcclApi* api = nullptr;
    int rc = needNccl(&api);
    if (rc != GB_OK) return rc;
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i]) return gb::failWith(GB_ERR_INVALID, "null context");
        for (int j = 0; j < i; ++j) if (ctxs[j]->device == ctxs[i]->device) return gb::failWith(GB_ERR_INVALID, "two contexts on one device");
        devs[i] = ctxs[i]->device;
        dropComm(ctxs[i]);
    }
    std::vector<ncclComm_t> comms(n, nullptr);
    GB_NCCL(api, api->commInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
        ctxs[i]->comm = comms[i];
        ctxs[i]->commOwned = true;
        ctxs[i]->commRanks = n;
    }
    return GB_OK;
}

int gb_comm_unique_id(void* id, size_t bytes) {
    if (!id || bytes < sizeof(ncclUniqueId)) return gb::failWith(GB_ERR_INVALID, "the id buffer must hold GB_COMM_ID_BYTES bytes");
    NcclApi* api = nullptr;
    int rc = needNccl(&api);
    if (rc != GB_OK) return rc;
    ncclUniqueId u;
    GB_NCCL(api, api->getUniqueId(&u));
    std::memcpy(id, &u, sizeof u);
    return GB_OK;
}

int gb_comm_init_rank(gb_context* ctx, const void* id, size_t bytes, int nranks, int rank) {
    if (!ctx || !id || bytes < sizeof(ncclUniqueId)) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return gb::failWith(GB_ERR_INVALID, "bad rank / world size");
    NcclApi* api = nullptr;
    int rc = needNccl(&api);
    if (rc != GB_OK) return rc;
    GB_CUDA(cudaSetDevice(ctx->device));
    dropComm(ctx);
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    ncclComm_t comm = nullptr;
    GB_NCCL(api, api->commInitRank(&comm, nranks, u, rank));
    ctx->comm = comm;
    ctx->commOwned = true;
    ctx->commRanks = nranks;
    return GB_OK;
}

int gb_comm_attach(gb_context* ctx, void* nccl_comm, int nranks) {
    if (!ctx || !nccl_comm || nranks < 1) return gb::failWith(GB_ERR_INVALID, "null argument");
    NcclApi* api = nullptr;
    int rc = needNccl(&api);
    if (rc != GB_OK) return rc;
    dropComm(ctx);
    ctx->comm = nccl_comm;
    ctx->commOwned = false;
    ctx->commRanks = nranks;
    return GB_OK;
}

int gb_comm_destroy(gb_context* ctx) {
    if (!ctx) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (ctx->comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
    }
    return dropComm(ctx);
}

int gb_comm_size(gb_context* ctx, int* nranks) {
    if (!ctx || !nranks) return gb::failWith(GB_ERR_INVALID, "null argument");
    *nranks = ctx->comm ? ctx->commRanks : 0;
    return GB_OK;
}

// One process per GPU: sum this rank's device film with every other rank's, in place, on the context's stream
// (ordered after the render kernels, before a later gb_film_download).  A context without a communicator, or a
// communicator of one rank, has nothing to add: the call succeeds.
int gb_film_allreduce(gb_context* ctx) {
    if (!ctx) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (!ctx->haveScene) return gb::failWith(GB_ERR_STATE, "no scene uploaded");
    if (!ctx->comm || ctx->commRanks < 2) return GB_OK;
    NcclApi* api = nullptr;
    int rc = needNccl(&api);
    if (rc != GB_OK) return rc;
    GB_CUDA(cudaSetDevice(ctx->device));
    GB_NCCL(api, api->allReduce(ctx->film, ctx->film, ctx->filmPixels * 4, ncclFloat, ncclSum, static_cast<ncclComm_t>(ctx->comm), ctx->stream));
    return GB_OK;
}

// Single process, several contexts of one communicator: the calls of all ranks go into one NCCL group.
int gb_film_allreduce_all(gb_context** ctxs, int n) {
    if (!ctxs || n < 1) return gb::failWith(GB_ERR_INVALID, "null argument");
    if (n == 1) return GB_OK;
    NcclApi* api = nullptr;
    int rc = needNccl(&api);
    if (rc != GB_OK) return rc;
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i] || !ctxs[i]->haveScene) return gb::failWith(GB_ERR_STATE, "a context has no scene");
        if (!ctxs[i]->comm || ctxs[i]->commRanks != n) return gb::failWith(GB_ERR_STATE, "gb_comm_init_all was not called for these contexts");
        if (ctxs[i]->filmPixels != ctxs[0]->filmPixels) return gb::failWith(GB_ERR_INVALID, "films differ in size");
    }
    GB_NCCL(api, api->groupStart());
    ncclResult_t first = ncclSuccess;
    for (int i = 0; i < n; ++i) {
        cudaSetDevice(ctxs[i]->device);
        ncclResult_t r = api->allReduce(ctxs[i]->film, ctxs[i]->film, ctxs[i]->filmPixels * 4, ncclFloat, ncclSum,
            static_cast<ncclComm_t>(ctxs[i]->comm), ctxs[i]->stream);
        if (r != ncclSuccess && first == ncclSuccess) first = r;
    }
    GB_NCCL(api, api->groupEnd());
    if (first != ncclSuccess) return gb::failWith(GB_ERR_CUDA, std::string("ncclAllReduce: ") + api->getErrorString(first));
    return GB_OK;
}

int gb_nccl_version(int* version) {
    if (!version) return gb::failWith(GB_ERR_INVALID, "null argument");
    NcclApi* api = nullptr;
    int rc = needNccl(&api);
    if (rc != GB_OK) return rc;
    GB_NCCL(api, api->getVersion(version));
    return GB_OK;
}

} // extern "C"
