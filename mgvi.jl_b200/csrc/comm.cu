// Collective plumbing for the sample-sharded Newton phase.  Residual samples
// are independent, so the sampler needs no communication; the Newton phase sums the KL value, the
// gradient input sum_p g'(s_p) dl/dlambda(p) and the mean metric weight over ALL residual samples.
// To make the result independent of the number of GPUs the sums run in one fixed, canonical order
// (engine.h: CanonShape): every rank forms the partial sums of the canonical chunks it owns, the
// chunk partials are ALL-GATHERED (no arithmetic inside the collective, so its internal order does
// not matter) and every rank adds the chunks up in chunk order.  libnccl is resolved at run time
// (dlopen) so the library loads on machines without NCCL and picks up the copy already mapped by
// the host process; alternatively the host supplies its own all-gather
// (mgvib200_comm_init_external: MPI from Julia, gloo in the CPU tests).
#include "engine.h"

#if defined(__CUDACC__) && !defined(MGVI_SIMT_EMU)
#include <dlfcn.h>
#define MGVI_HAVE_NCCL 1
#endif

namespace mgvi {

struct Comm {
    void* nccl = nullptr;                    // ncclComm_t
    mgvib200_allgather_fn ext = nullptr;     // host-supplied all-gather
    void* ext_user = nullptr;
};

#ifdef MGVI_HAVE_NCCL
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_comm_init_rank)(nccl_comm*, int, nccl_uid, int);
typedef int (*fn_comm_destroy)(nccl_comm);
typedef int (*fn_all_gather)(const void*, void*, size_t, int, nccl_comm, cudaStream_t);
typedef const char* (*fn_err_string)(int);

struct NcclApi {
    void* lib = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_comm_init_rank init_rank = nullptr;
    fn_comm_destroy destroy = nullptr;
    fn_all_gather all_gather = nullptr;
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
    api.all_gather = (fn_all_gather)dlsym(api.lib, "ncclAllGather");
    api.err_string = (fn_err_string)dlsym(api.lib, "ncclGetErrorString");
    if (!api.get_uid || !api.init_rank || !api.all_gather) throw RtError{"libnccl: missing symbols", 4};
    return api;
}

static void nccl_check(int rc, const char* what) {
    if (rc != 0) {
        NcclApi& a = nccl_api();
        std::string m = std::string(what) + " failed: " + (a.err_string ? a.err_string(rc) : "nccl error");
        throw RtError{m, 4};
    }
}
#endif

void comm_unique_id(uint8_t out[128]) {
#ifdef MGVI_HAVE_NCCL
    nccl_uid id;
    nccl_check(nccl_api().get_uid(&id), "ncclGetUniqueId");
    memcpy(out, id.internal, 128);
#else
    memset(out, 0, 128);
#endif
}

static void comm_reset(mgvib200_ctx* ctx) {
    ctx->n_cache_local = -1;
    ctx->n_cache_glob = -1;
}

static void comm_warmup(mgvib200_ctx* ctx) {
    // NCCL connects lazily on the first collective (hundreds of ms): pay that here, not inside the
    // first Newton step
    double* d = (double*)rt_malloc(sizeof(double) * 8 * ctx->world);
    rt_memset(d, 0, sizeof(double) * 8 * ctx->world, ctx->stream);
    comm_allgather(ctx, d + 8 * ctx->rank, d, 8);
    rt_sync(ctx->stream);
    rt_free(d);
}

void comm_init(mgvib200_ctx* ctx, int rank, int world, const uint8_t idb[128]) {
    if (ctx->comm) throw RtError{"comm_init: the context already has a communicator", 1};
    comm_reset(ctx);
    if (world <= 1) { ctx->rank = 0; ctx->world = 1; return; }
    if (rank < 0 || rank >= world) throw RtError{"comm_init: rank out of range", 1};
    if (ctx->n_models_single_rank > 0)
        throw RtError{"comm_init: the context holds POLY / DENSE_LINEAR models, which are single-GPU ('replicas only')", 3};
#ifdef MGVI_HAVE_NCCL
    nccl_uid id;
    memcpy(id.internal, idb, 128);
    Comm* c = new Comm();
    nccl_check(nccl_api().init_rank(&c->nccl, world, id, rank), "ncclCommInitRank");
    ctx->comm = c;
    ctx->rank = rank;
    ctx->world = world;
    comm_warmup(ctx);
#else
    (void)idb;
    throw RtError{"this build has no NCCL: use mgvib200_comm_init_external", 3};
#endif
}

void comm_init_external(mgvib200_ctx* ctx, int rank, int world, mgvib200_allgather_fn fn, void* user) {
    if (ctx->comm) throw RtError{"comm_init_external: the context already has a communicator", 1};
    comm_reset(ctx);
    if (world <= 1) { ctx->rank = 0; ctx->world = 1; return; }
    if (rank < 0 || rank >= world) throw RtError{"comm_init_external: rank out of range", 1};
    if (!fn) throw RtError{"comm_init_external: all-gather callback is null", 1};
    if (ctx->n_models_single_rank > 0)
        throw RtError{"comm_init_external: the context holds POLY / DENSE_LINEAR models, which are single-GPU", 3};
    Comm* c = new Comm();
    c->ext = fn;
    c->ext_user = user;
    ctx->comm = c;
    ctx->rank = rank;
    ctx->world = world;
}

void comm_destroy(mgvib200_ctx* ctx) {
    if (ctx->comm) {
#ifdef MGVI_HAVE_NCCL
        if (ctx->comm->nccl) nccl_api().destroy(ctx->comm->nccl);
#endif
        delete ctx->comm;
        ctx->comm = nullptr;
    }
}

// recv[r * count .. (r+1) * count) = rank r's `send` (count doubles); in place when
// send == recv + rank * count.  Device pointers; ordered on the context's stream.
void comm_allgather(mgvib200_ctx* ctx, const double* send, double* recv, long long count) {
    if (!ctx->comm || ctx->world <= 1) {
        if (send != recv) rt_d2d(recv, send, sizeof(double) * (size_t)count, ctx->stream);
        return;
    }
    if (ctx->comm->ext) {
        rt_sync(ctx->stream);        // the callback is host code: inputs must be complete, it returns when done
        const int rc = ctx->comm->ext(ctx->comm->ext_user, send, recv, (int64_t)count);
        if (rc != 0) throw RtError{"external all-gather callback failed", 4};
        return;
    }
#ifdef MGVI_HAVE_NCCL
    // ncclDouble = 8
    nccl_check(nccl_api().all_gather(send, recv, (size_t)count, 8, ctx->comm->nccl, ctx->stream), "ncclAllGather");
    ctx->launches++;
#endif
}

// total residual count over the ranks; every rank must hold the same number (contiguous equal
// shards).  One tiny all-gather the first time a given local count is seen, cached
// afterwards (the count is constant for a whole optimisation).
int64_t comm_global_count(mgvib200_ctx* ctx, int64_t local) {
    if (!ctx->comm || ctx->world <= 1) return local;
    if (ctx->n_cache_local == local) return ctx->n_cache_glob;
    ctx->count_scalar.ensure(sizeof(double) * ctx->world);
    double* d = ctx->count_scalar.as<double>();
    double h = (double)local;
    rt_h2d(d + ctx->rank, &h, sizeof(double), ctx->stream);
    comm_allgather(ctx, d + ctx->rank, d, 1);
    std::vector<double> all(ctx->world);
    rt_d2h(all.data(), d, sizeof(double) * ctx->world, ctx->stream);
    rt_sync(ctx->stream);
    for (int r = 0; r < ctx->world; ++r)
        if (all[r] != h) throw RtError{"multi-GPU: every rank must hold the same number of residual samples", 1};
    ctx->n_cache_local = local;
    ctx->n_cache_glob = local * ctx->world;
    return ctx->n_cache_glob;
}

// canonical chunking of the n_glob residual samples (engine.h)
CanonShape canon_shape(mgvib200_ctx* ctx, int64_t n_local) {
    CanonShape c;
    c.n_glob = comm_global_count(ctx, n_local);
    int64_t g = c.n_glob, e = 8;
    while (e) { const int64_t t = g % e; g = e; e = t; }      // gcd(n_glob, 8)
    c.chunks = g;
    if (ctx->world > 1 && (c.chunks % ctx->world) != 0)
        throw RtError{"multi-GPU: the number of ranks must be 1, 2, 4 or 8 (and divide the residual count)", 1};
    c.per_chunk = c.n_glob / c.chunks;
    c.chunks_local = c.chunks / ctx->world;
    return c;
}

}  // namespace mgvi
