// NCCL plumbing for the sample-sharded Newton phase.  Residual samples are
// independent, so the sampler needs no communication; the Newton phase sums the KL value, the
// gradient input sum_p g'(s_p) dl/dlambda(p) and the mean metric weight over ranks -- one
// all-reduce (double, sum) each.  libnccl is resolved at run time (dlopen) so the library loads
// on machines without NCCL and picks up the copy already mapped by the host process.
#include "engine.h"

#if defined(__CUDACC__) && !defined(MGVI_SIMT_EMU)
#include <dlfcn.h>

namespace mgvi {

typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_comm_init_rank)(nccl_comm*, int, nccl_uid, int);
typedef int (*fn_comm_destroy)(nccl_comm);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef const char* (*fn_err_string)(int);

struct NcclApi {
    void* lib = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_comm_init_rank init_rank = nullptr;
    fn_comm_destroy destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_err_string err_string = nullptr;
};

static NcclApi& nccl_api() {
    static NcclApi api;
    if (api.lib) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) throw RtError{"libnccl.so.2 not found (multi-GPU needs NCCL)", 4};
    api.get_uid = (fn_get_uid)dlsym(api.lib, "ncclGetUniqueId");
    api.init_rank = (fn_comm_init_rank)dlsym(api.lib, "ncclCommInitRank");
    api.destroy = (fn_comm_destroy)dlsym(api.lib, "ncclCommDestroy");
    api.all_reduce = (fn_all_reduce)dlsym(api.lib, "ncclAllReduce");
    api.err_string = (fn_err_string)dlsym(api.lib, "ncclGetErrorString");
    if (!api.get_uid || !api.init_rank || !api.all_reduce) throw RtError{"libnccl: missing symbols", 4};
    return api;
}

struct Comm {
    nccl_comm comm = nullptr;
};

static void nccl_check(int rc, const char* what) {
    if (rc != 0) {
        NcclApi& a = nccl_api();
        std::string m = std::string(what) + " failed: " + (a.err_string ? a.err_string(rc) : "nccl error");
        throw RtError{m, 4};
    }
}

void comm_unique_id(uint8_t out[128]) {
    nccl_uid id;
    nccl_check(nccl_api().get_uid(&id), "ncclGetUniqueId");
    memcpy(out, id.internal, 128);
}

void comm_init(mgvib200_ctx* ctx, int rank, int world, const uint8_t idb[128]) {
    if (world <= 1) { ctx->rank = 0; ctx->world = 1; return; }
    nccl_uid id;
    memcpy(id.internal, idb, 128);
    Comm* c = new Comm();
    nccl_check(nccl_api().init_rank(&c->comm, world, id, rank), "ncclCommInitRank");
    ctx->comm = c;
    ctx->rank = rank;
    ctx->world = world;
    // NCCL connects lazily on the first collective (hundreds of ms): pay that here, not inside the
    // first Newton step
    double* d = (double*)rt_malloc(sizeof(double) * 8);
    rt_memset(d, 0, sizeof(double) * 8, ctx->stream);
    comm_allreduce_sum(ctx, d, 8);
    rt_sync(ctx->stream);
    rt_free(d);
}

void comm_destroy(mgvib200_ctx* ctx) {
    if (ctx->comm) {
        if (ctx->comm->comm) nccl_api().destroy(ctx->comm->comm);
        delete ctx->comm;
        ctx->comm = nullptr;
    }
}

void comm_allreduce_sum(mgvib200_ctx* ctx, double* dev, long long count) {
    if (!ctx->comm || ctx->world <= 1) return;
    // ncclDouble = 8, ncclSum = 0
    nccl_check(nccl_api().all_reduce(dev, dev, (size_t)count, 8, 0, ctx->comm->comm, ctx->stream), "ncclAllReduce");
    ctx->launches++;
}

int64_t comm_global_count(mgvib200_ctx* ctx, int64_t local) {
    if (!ctx->comm || ctx->world <= 1) return local;
    ctx->count_scalar.ensure(sizeof(double));
    double* d = ctx->count_scalar.as<double>();
    double h = (double)local;
    rt_h2d(d, &h, sizeof(double), ctx->stream);
    comm_allreduce_sum(ctx, d, 1);
    rt_d2h(&h, d, sizeof(double), ctx->stream);
    rt_sync(ctx->stream);
    return (int64_t)llround(h);
}

}  // namespace mgvi

#else   // emulator build: single rank only

namespace mgvi {
struct Comm {};
void comm_unique_id(uint8_t out[128]) { memset(out, 0, 128); }
void comm_init(mgvib200_ctx* ctx, int, int world, const uint8_t*) {
    if (world > 1) throw RtError{"emulator build is single-rank", 3};
    ctx->rank = 0; ctx->world = 1;
}
void comm_destroy(mgvib200_ctx*) {}
void comm_allreduce_sum(mgvib200_ctx*, double*, long long) {}
int64_t comm_global_count(mgvib200_ctx*, int64_t local) { return local; }
}  // namespace mgvi

#endif
