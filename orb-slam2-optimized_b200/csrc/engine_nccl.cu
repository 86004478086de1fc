// engine_nccl.cu -- native NCCL exchange of the per-candidate records.
//
// The only exchange step of the path is one all-gather of fixed-size records (rsac_result, 96 B
// per candidate) per sweep; the communicator is persistent and the collective is enqueued on
// the engine's stream so it orders after the replay kernel without a host round trip.
// libnccl.so.2 is resolved at run time (dlopen): the library has no link-time NCCL dependency
// and picks up the copy the host process already loaded (e.g. the one bundled with torch).
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <string>
#include "../../include/ransac_b200.h"
#include "engine_state.cuh"

namespace {
typedef struct { char internal[128]; } nccl_unique_id_t;
typedef void* nccl_comm_t;
typedef int (*fn_get_unique_id)(nccl_unique_id_t*);
typedef int (*fn_comm_init_rank)(nccl_comm_t*, int, nccl_unique_id_t, int);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_comm_destroy)(nccl_comm_t);
typedef const char* (*fn_get_error_string)(int);

struct NcclApi {
    void* lib = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_get_error_string get_error_string = nullptr;
    bool ok() const { return lib && get_unique_id && comm_init_rank && all_gather && comm_destroy; }
};

NcclApi& nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) {
            api.get_unique_id = (fn_get_unique_id)dlsym(api.lib, "ncclGetUniqueId");
            api.comm_init_rank = (fn_comm_init_rank)dlsym(api.lib, "ncclCommInitRank");
            api.all_gather = (fn_all_gather)dlsym(api.lib, "ncclAllGather");
            api.comm_destroy = (fn_comm_destroy)dlsym(api.lib, "ncclCommDestroy");
            api.get_error_string = (fn_get_error_string)dlsym(api.lib, "ncclGetErrorString");
        }
    });
    return api;
}
}  // namespace

int rsac_nccl_get_unique_id(void* id128)
{
    if (!id128) return RSAC_ERR_INVALID;
    NcclApi& n = nccl_api();
    if (!n.ok()) return RSAC_ERR_STATE;
    nccl_unique_id_t id;
    if (n.get_unique_id(&id) != 0) return RSAC_ERR_CUDA;
    memcpy(id128, &id, sizeof(id));
    return RSAC_OK;
}

int rsac_nccl_init(rsac_engine* e, const void* id128, int rank, int world)
{
    if (!e || !id128 || world < 1 || rank < 0 || rank >= world) return RSAC_ERR_INVALID;
    NcclApi& n = nccl_api();
    if (!n.ok()) { e->err = "libnccl.so.2 not found"; return RSAC_ERR_STATE; }
    RSAC_CUDA(e, cudaSetDevice(e->device));
    nccl_unique_id_t id;
    memcpy(&id, id128, sizeof(id));
    nccl_comm_t comm = nullptr;
    const int rc = n.comm_init_rank(&comm, world, id, rank);
    if (rc != 0) { e->err = std::string("ncclCommInitRank: ") + (n.get_error_string ? n.get_error_string(rc) : "error"); return RSAC_ERR_CUDA; }
    e->nccl_comm = comm;
    return RSAC_OK;
}

int rsac_nccl_allgather_results(rsac_engine* e, const void* d_send, int count_per_rank, void* d_gathered)
{
    if (!e || !d_send || !d_gathered || count_per_rank < 0) return RSAC_ERR_INVALID;
    if (!e->nccl_comm) { e->err = "rsac_nccl_allgather_results before rsac_nccl_init"; return RSAC_ERR_STATE; }
    NcclApi& n = nccl_api();
    const int rc = n.all_gather(d_send, d_gathered, (size_t)count_per_rank * sizeof(rsac_result), /*ncclChar*/ 0,
                                (nccl_comm_t)e->nccl_comm, e->stream);
    if (rc != 0) { e->err = std::string("ncclAllGather: ") + (n.get_error_string ? n.get_error_string(rc) : "error"); return RSAC_ERR_CUDA; }
    return RSAC_OK;
}

int rsac_nccl_destroy(rsac_engine* e)
{
    if (!e) return RSAC_ERR_INVALID;
    if (e->nccl_comm) {
        nccl_api().comm_destroy((nccl_comm_t)e->nccl_comm);
        e->nccl_comm = nullptr;
    }
    return RSAC_OK;
}
