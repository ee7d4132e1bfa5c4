// Thin runtime layer under the host orchestration code: device memory, copies, stream sync,
// events and kernel launch.  The product build (nvcc) maps it onto the CUDA runtime.  The
// -DMGVI_SIMT_EMU build (tests/simt_emu/, TEST INFRASTRUCTURE ONLY, see simt.h) maps it onto
// malloc/memcpy and the fiber emulator so the host logic can be exercised without a GPU.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <string>
#include <utility>
#include <vector>
#include "simt.h"

namespace mgvi {

struct RtError {
    std::string msg;
    int code;
};

#if defined(__CUDACC__) && !defined(MGVI_SIMT_EMU)

typedef cudaStream_t rt_stream;
typedef cudaEvent_t rt_event;

#define RT_CUDA(call)                                                                           \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            char b_[512];                                                                       \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            throw mgvi::RtError{b_, 2};                                                         \
        }                                                                                       \
    } while (0)

inline void rt_set_device(int dev) { RT_CUDA(cudaSetDevice(dev)); }
inline rt_stream rt_stream_create() { rt_stream s; RT_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)); return s; }
inline void rt_stream_destroy(rt_stream s) { cudaStreamDestroy(s); }
inline void* rt_malloc(size_t bytes) { void* p = nullptr; RT_CUDA(cudaMalloc(&p, bytes ? bytes : 16)); return p; }
inline void rt_free(void* p) { if (p) cudaFree(p); }
inline void* rt_host_alloc(size_t bytes) { void* p = nullptr; RT_CUDA(cudaMallocHost(&p, bytes ? bytes : 16)); return p; }
inline void rt_host_free(void* p) { if (p) cudaFreeHost(p); }
inline void rt_h2d(void* d, const void* h, size_t b, rt_stream s) { RT_CUDA(cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, s)); }
inline void rt_d2h(void* h, const void* d, size_t b, rt_stream s) { RT_CUDA(cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, s)); }
inline void rt_d2d(void* d, const void* s_, size_t b, rt_stream s) { RT_CUDA(cudaMemcpyAsync(d, s_, b, cudaMemcpyDeviceToDevice, s)); }
inline void rt_memset(void* d, int v, size_t b, rt_stream s) { RT_CUDA(cudaMemsetAsync(d, v, b, s)); }
inline void rt_sync(rt_stream s) { RT_CUDA(cudaStreamSynchronize(s)); }
inline rt_event rt_event_create() { rt_event e; RT_CUDA(cudaEventCreate(&e)); return e; }
inline void rt_event_destroy(rt_event e) { cudaEventDestroy(e); }
inline void rt_event_record(rt_event e, rt_stream s) { RT_CUDA(cudaEventRecord(e, s)); }
inline double rt_event_ms(rt_event a, rt_event b) { float ms = 0; RT_CUDA(cudaEventSynchronize(b)); RT_CUDA(cudaEventElapsedTime(&ms, a, b)); return ms; }

// opt-in to > 48 KB of dynamic shared memory is a per-(kernel, device) attribute: set it once per pair
// (and again only if a larger size is asked for), not on every launch
inline void rt_ensure_smem_attr(const void* fn, size_t smem) {
    struct Key { const void* fn; int dev; };
    static std::vector<std::pair<Key, size_t>> cache;       // a few dozen entries: linear scan
    int dev = 0;
    RT_CUDA(cudaGetDevice(&dev));
    for (auto& e : cache)
        if (e.first.fn == fn && e.first.dev == dev) {
            if (e.second >= smem) return;
            RT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            e.second = smem;
            return;
        }
    RT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cache.push_back({Key{fn, dev}, smem});
}

template <typename K, typename P>
inline void rt_launch(K kernel, long long grid, int block, size_t smem, rt_stream s, const P& params) {
    if (grid <= 0) return;
    if (grid > 2147483647LL) throw RtError{"grid too large", 1};
    if (smem > 48 * 1024) rt_ensure_smem_attr((const void*)kernel, smem);
    kernel<<<(unsigned)grid, block, smem, s>>>(params);
    RT_CUDA(cudaGetLastError());
}

// Programmatic dependent launch (MGVIB200_PDL=1): the kernel may be scheduled while its predecessor in the
// stream drains; it must execute SIMT_PDL_ENTRY() (griddepcontrol.wait) before touching global memory.  Only
// kernels whose entry point does that are launched through here.
inline bool rt_pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("MGVIB200_PDL"); on = (e && atoi(e)) ? 1 : 0; }
    return on == 1;
}
template <typename K, typename P>
inline void rt_launch_pdl(K kernel, long long grid, int block, size_t smem, rt_stream s, const P& params) {
    if (!rt_pdl_enabled()) { rt_launch(kernel, grid, block, smem, s, params); return; }
    if (grid <= 0) return;
    if (grid > 2147483647LL) throw RtError{"grid too large", 1};
    if (smem > 48 * 1024) rt_ensure_smem_attr((const void*)kernel, smem);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    RT_CUDA(cudaLaunchKernelEx(&cfg, kernel, params));
}

// CUDA graphs: a fixed launch sequence (several lock-step CG iterations -- every per-iteration scalar
// lives in device memory, so the kernel parameters do not change) captured once and replayed
typedef cudaGraphExec_t rt_graph;
static const bool rt_has_graphs = true;
inline void rt_capture_begin(rt_stream s) { RT_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal)); }
inline rt_graph rt_capture_end(rt_stream s) {
    cudaGraph_t g = nullptr;
    RT_CUDA(cudaStreamEndCapture(s, &g));
    cudaGraphExec_t e = nullptr;
    cudaError_t rc = cudaGraphInstantiate(&e, g, 0);
    cudaGraphDestroy(g);
    RT_CUDA(rc);
    return e;
}
inline void rt_capture_abort(rt_stream s) { cudaGraph_t g = nullptr; cudaStreamEndCapture(s, &g); if (g) cudaGraphDestroy(g); cudaGetLastError(); }
inline void rt_graph_launch(rt_graph g, rt_stream s) { RT_CUDA(cudaGraphLaunch(g, s)); }
inline void rt_graph_destroy(rt_graph g) { if (g) cudaGraphExecDestroy(g); }
inline void rt_event_sync(rt_event e) { RT_CUDA(cudaEventSynchronize(e)); }
// make `s` wait (on the device) for everything recorded into `e`
inline void rt_stream_wait(rt_stream s, rt_event e) { RT_CUDA(cudaStreamWaitEvent(s, e, 0)); }

#else  // ---------------------------------- emulation --------------------------------------------

typedef int rt_stream;
typedef int rt_event;
inline void rt_set_device(int) {}
inline rt_stream rt_stream_create() { return 0; }
inline void rt_stream_destroy(rt_stream) {}
// Emulated device allocations carry a guard zone on both sides: NaN-filled (so an out-of-bounds
// READ poisons the results the tests compare) and checked on free (an out-of-bounds WRITE aborts
// the test run with a message).  This stands in for compute-sanitizer memcheck, which the GPU
// pool does not offer.
static const size_t kEmuGuard = 4096;
struct EmuAllocHeader { size_t bytes; unsigned long long magic; };
inline void* rt_malloc(size_t bytes) {
    size_t b = bytes ? bytes : 16;
    const size_t body = (b + 255) / 256 * 256;
    unsigned char* raw = (unsigned char*)aligned_alloc(256, kEmuGuard + body + kEmuGuard);
    memset(raw, 0xff, kEmuGuard + body + kEmuGuard);      // NaN-poison: uninitialised / stray reads show up in the results
    EmuAllocHeader h{b, 0x6d67766962323030ULL};
    memcpy(raw, &h, sizeof h);
    return raw + kEmuGuard;
}
inline void rt_free(void* p) {
    if (!p) return;
    unsigned char* raw = (unsigned char*)p - kEmuGuard;
    EmuAllocHeader h;
    memcpy(&h, raw, sizeof h);
    bool ok = (h.magic == 0x6d67766962323030ULL);
    if (ok) {
        for (size_t i = sizeof h; i < kEmuGuard && ok; ++i) ok = (raw[i] == 0xff);
        const unsigned char* tail = (const unsigned char*)p + h.bytes;
        const size_t body = (h.bytes + 255) / 256 * 256;
        for (size_t i = 0; i < (body - h.bytes) + kEmuGuard && ok; ++i) ok = (tail[i] == 0xff);
    }
    if (!ok) {
        fprintf(stderr, "simt emulator: out-of-bounds WRITE next to a device allocation of %zu bytes\n", h.bytes);
        abort();
    }
    free(raw);
}
inline void* rt_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 16); }
inline void rt_host_free(void* p) { free(p); }
inline void rt_h2d(void* d, const void* h, size_t b, rt_stream) { memcpy(d, h, b); }
inline void rt_d2h(void* h, const void* d, size_t b, rt_stream) { memcpy(h, d, b); }
inline void rt_d2d(void* d, const void* s_, size_t b, rt_stream) { memmove(d, s_, b); }
inline void rt_memset(void* d, int v, size_t b, rt_stream) { memset(d, v, b); }
inline void rt_sync(rt_stream) {}
inline rt_event rt_event_create() { return 0; }
inline void rt_event_destroy(rt_event) {}
inline void rt_event_record(rt_event, rt_stream) {}
inline double rt_event_ms(rt_event, rt_event) { return 0.0; }

typedef int rt_graph;
static const bool rt_has_graphs = false;
inline void rt_capture_begin(rt_stream) {}
inline rt_graph rt_capture_end(rt_stream) { return 0; }
inline void rt_capture_abort(rt_stream) {}
inline void rt_graph_launch(rt_graph, rt_stream) {}
inline void rt_graph_destroy(rt_graph) {}
inline void rt_event_sync(rt_event) {}
inline void rt_stream_wait(rt_stream, rt_event) {}      // the emulator executes launches in program order

template <typename K, typename P>
inline void rt_launch(K kernel, long long grid, int block, size_t smem, rt_stream, const P& params);
template <typename K, typename P>
inline void rt_launch_pdl(K kernel, long long grid, int block, size_t smem, rt_stream s, const P& params) {
    rt_launch(kernel, grid, block, smem, s, params);
}
template <typename K, typename P>
inline void rt_launch(K kernel, long long grid, int block, size_t smem, rt_stream, const P& params) {
    if (grid <= 0) return;
    if (block % 32 != 0) throw RtError{"emu: block size must be a multiple of 32", 1};
    if (smem > 227 * 1024) throw RtError{"emu: shared memory request above 227 KB", 1};
    unsigned char* sm = (unsigned char*)aligned_alloc(1024, (smem + 1023) / 1024 * 1024 + 1024);
    memset(sm, 0xff, smem + 1024);
    P local = params;
    simt_emu::launch(grid, block, [&]() { kernel(local, sm); });
    free(sm);
}

#endif

}  // namespace mgvi
