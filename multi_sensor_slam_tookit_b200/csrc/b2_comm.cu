// NCCL communicator of the C ABI (b2_comm_*). See b2_comm.cuh. The NCCL entry points are resolved with dlopen/dlsym
// from the prototypes of nccl.h 2.27/2.28 (ncclUniqueId is 128 opaque bytes passed by value).
#include "b2_comm.cuh"
#include <dlfcn.h>
#include <cstdlib>
#include <mutex>

namespace b2 {

struct NcclId { char internal[128]; };
typedef int (*fn_get_unique_id)(NcclId*);
typedef int (*fn_comm_init_rank)(void**, int, NcclId, int);
typedef int (*fn_comm_destroy)(void*);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*fn_error_string)(int);

struct NcclApi {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_all_gather all_gather = nullptr;
    fn_error_string error_string = nullptr;
    bool ok = false;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;

static void load_nccl() {
    const char* names[3] = {std::getenv("B2_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        g_nccl.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) return;
    g_nccl.get_unique_id = (fn_get_unique_id)dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.comm_init_rank = (fn_comm_init_rank)dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.comm_destroy = (fn_comm_destroy)dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.all_reduce = (fn_all_reduce)dlsym(g_nccl.handle, "ncclAllReduce");
    g_nccl.all_gather = (fn_all_gather)dlsym(g_nccl.handle, "ncclAllGather");
    g_nccl.error_string = (fn_error_string)dlsym(g_nccl.handle, "ncclGetErrorString");
    g_nccl.ok = g_nccl.get_unique_id && g_nccl.comm_init_rank && g_nccl.comm_destroy && g_nccl.all_reduce && g_nccl.all_gather && g_nccl.error_string;
}

static int nccl_api(NcclApi** out) {
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.ok) { set_error("NCCL is not available: %s", g_nccl.handle ? "missing symbols" : "libnccl.so.2 could not be opened (set B2_NCCL_LIB)"); return B2_ERR_NCCL; }
    *out = &g_nccl;
    return B2_OK;
}

#define B2_NCCL(api, expr)                                                                         \
    do {                                                                                           \
        int _r = (expr);                                                                           \
        if (_r != 0) { set_error("%s -> NCCL error %d: %s", #expr, _r, (api)->error_string(_r)); return B2_ERR_NCCL; } \
    } while (0)

int comm_allreduce_sum_f64(b2_comm_s* c, const double* d_send, double* d_recv, size_t n, cudaStream_t s) {
    NcclApi* api;
    B2_CHECK(nccl_api(&api));
    B2_NCCL(api, api->all_reduce(d_send, d_recv, n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, c->comm, s));
    return B2_OK;
}

// all-gather of `count` doubles per rank into d_recv (rank r's block at d_recv + r * count); in place when d_send is that block
int comm_allgather_f64(b2_comm_s* c, const double* d_send, double* d_recv, size_t count, cudaStream_t s) {
    NcclApi* api;
    B2_CHECK(nccl_api(&api));
    B2_NCCL(api, api->all_gather(d_send, d_recv, count, /*ncclFloat64*/ 8, c->comm, s));
    return B2_OK;
}

}  // namespace b2

using namespace b2;

extern "C" {

int b2_comm_unique_id(unsigned char id[128]) {
    if (!id) return B2_ERR_ARG;
    NcclApi* api;
    B2_CHECK(nccl_api(&api));
    NcclId nid;
    B2_NCCL(api, api->get_unique_id(&nid));
    memcpy(id, nid.internal, 128);
    return B2_OK;
}

int b2_comm_create(b2_comm_t* out, const unsigned char id[128], int rank, int world) {
    if (!out || !id || world < 1 || rank < 0 || rank >= world) return B2_ERR_ARG;
    *out = nullptr;
    NcclApi* api;
    B2_CHECK(nccl_api(&api));
    b2_comm_s* c = new b2_comm_s();
    c->rank = rank; c->world = world;
    if (cudaGetDevice(&c->device) != cudaSuccess) { set_error("b2_comm_create: no CUDA device"); delete c; return B2_ERR_CUDA; }
    NcclId nid;
    memcpy(nid.internal, id, 128);
    int r = api->comm_init_rank(&c->comm, world, nid, rank);
    if (r != 0) { set_error("ncclCommInitRank -> NCCL error %d: %s", r, api->error_string(r)); delete c; return B2_ERR_NCCL; }
    *out = c;
    return B2_OK;
}

int b2_comm_destroy(b2_comm_t c) {
    if (!c) return B2_OK;
    if (c->comm && g_nccl.ok) g_nccl.comm_destroy(c->comm);
    delete c;
    return B2_OK;
}

int b2_comm_rank(b2_comm_t c, int* rank, int* world) {
    if (!c) return B2_ERR_ARG;
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    return B2_OK;
}

/* sum-all-reduce of n host doubles through the device (used by tests and by callers that own no device buffers) */
int b2_comm_allreduce_f64(b2_comm_t c, double* values, size_t n) {
    if (!c || (n && !values)) return B2_ERR_ARG;
    if (!n) return B2_OK;
    double* d = nullptr;
    B2_CUDA(cudaMalloc(&d, n * 8));
    cudaStream_t s;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("b2_comm_allreduce_f64: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(d);
        return B2_ERR_CUDA;
    }
    int st = B2_OK;
    if (cudaMemcpyAsync(d, values, n * 8, cudaMemcpyHostToDevice, s) != cudaSuccess) st = B2_ERR_CUDA;
    if (st == B2_OK) st = comm_allreduce_sum_f64(c, d, d, n, s);
    if (st == B2_OK && (cudaMemcpyAsync(values, d, n * 8, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)) {
        set_error("b2_comm_allreduce_f64: %s", cudaGetErrorString(cudaGetLastError())); st = B2_ERR_CUDA;
    }
    cudaStreamDestroy(s); cudaFree(d);
    return st;
}

}  // extern "C"
