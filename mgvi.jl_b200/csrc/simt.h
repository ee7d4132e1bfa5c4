// SIMT portability macros.
//
// Kernel BODIES in this directory are written once against the tiny vocabulary below.
//  * Under nvcc (the product build, sm_100a) every macro maps 1:1 onto the CUDA builtin.
//  * Under g++ with -DMGVI_SIMT_EMU (tests/simt_emu/, TEST INFRASTRUCTURE ONLY) the same bodies
//    run on a single-threaded fiber scheduler so that index arithmetic, barrier placement and
//    reduction logic can be checked on a box without a GPU.  The emulator runtime lives under
//    tests/ and is never compiled into, linked with or loaded by libmgvib200.so.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(MGVI_SIMT_EMU)

#include <cuda_runtime.h>
#define SIMT_DEVICE __device__ __forceinline__
#define SIMT_HD __host__ __device__ __forceinline__
#define SIMT_TID ((int)threadIdx.x)
#define SIMT_NTID ((int)blockDim.x)
#define SIMT_BID ((long long)blockIdx.x)
#define SIMT_NBID ((long long)gridDim.x)
#define SIMT_SYNC() __syncthreads()
#define SIMT_FENCE() __threadfence()
#define SIMT_SHFL_XOR_F64(v, m) __shfl_xor_sync(0xffffffffu, (v), (m))
#define SIMT_SHFL_IDX_I32(v, l) __shfl_sync(0xffffffffu, (v), (l))
#define SIMT_ATOMIC_INC_U32(p) atomicAdd((p), 1u)
#define SIMT_ATOMIC_ADD_I32(p, v) atomicAdd((p), (v))
#define SIMT_LDG(p) __ldg(p)
#define SIMT_LDCG(p) __ldcg(p)
// L2 prefetch hints (no register, no scoreboard): one contiguous chunk / one line
#define SIMT_PREFETCH_BULK_L2(ptr, bytes)                                                          \
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"((unsigned)(bytes)) : "memory")
#define SIMT_PREFETCH_L2(ptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr))
#define SIMT_PREFETCH_L1(ptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr))
// TMA bulk copy global -> shared with mbarrier completion (SASS UBLKCP): one thread arms the barrier
// with the byte count and issues the copy; consumers wait on the barrier phase.  No registers, no LSU.
#define SIMT_MBAR_INIT(mbar, count)                                                                 \
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n\tfence.mbarrier_init.release.cluster;\n\t" \
                 "fence.proxy.async.shared::cta;" ::"r"((unsigned)__cvta_generic_to_shared(mbar)), "r"((unsigned)(count)) : "memory")
#define SIMT_BULK_G2S(dst, src, bytes, mbar)                                                        \
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"                        \
                 "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3], %1, [%0];" \
                 ::"r"((unsigned)__cvta_generic_to_shared(mbar)), "r"((unsigned)(bytes)),            \
                   "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory")
#define SIMT_MBAR_WAIT(mbar, parity)                                                                \
    do {                                                                                            \
        unsigned done_ = 0;                                                                         \
        while (!done_)                                                                              \
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" \
                         : "=r"(done_) : "r"((unsigned)__cvta_generic_to_shared(mbar)), "r"((unsigned)(parity)) : "memory"); \
    } while (0)
// optimisation fence on a 32-bit value: stops the compiler from precomputing (and spilling) the
// addresses of a later phase at the top of the kernel
#define SIMT_OPAQUE_U32(x) asm volatile("" : "+r"(x))
#define SIMT_RSQRT(x) rsqrt(x)
// programmatic dependent launch: let the successor be scheduled now, then wait until the predecessor grid has
// completed and its writes are visible (a no-op when the kernel was not launched as a dependent)
#define SIMT_PDL_ENTRY()                                                   \
    do {                                                                   \
        asm volatile("griddepcontrol.launch_dependents;");                 \
        asm volatile("griddepcontrol.wait;" ::: "memory");                 \
    } while (0)
#define SIMT_KERNEL(name, params_t, body_call)                                   \
    __global__ void name(const params_t P) {                                     \
        extern __shared__ __align__(16) unsigned char simt_smem[];               \
        body_call;                                                               \
    }

#else  // ------------------------------- emulation ---------------------------------------------

#include "simt_emu_runtime.h"   // tests/simt_emu/ (passed with -I); defines simt_emu::*
#include <cmath>
#define SIMT_DEVICE inline
#define SIMT_HD inline
#define SIMT_TID (simt_emu::tid())
#define SIMT_NTID (simt_emu::ntid())
#define SIMT_BID ((long long)simt_emu::bid())
#define SIMT_NBID ((long long)simt_emu::nbid())
#define SIMT_SYNC() simt_emu::block_barrier()
#define SIMT_FENCE() ((void)0)
#define SIMT_SHFL_XOR_F64(v, m) simt_emu::shfl_xor_f64((v), (m))
#define SIMT_SHFL_IDX_I32(v, l) simt_emu::shfl_idx_i32((v), (l))
#define SIMT_ATOMIC_INC_U32(p) simt_emu::atomic_inc_u32(p)
#define SIMT_ATOMIC_ADD_I32(p, v) simt_emu::atomic_add_i32((p), (v))
#define SIMT_LDG(p) (*(p))
#define SIMT_LDCG(p) (*(p))
#define SIMT_PREFETCH_BULK_L2(ptr, bytes) ((void)0)
#define SIMT_PREFETCH_L2(ptr) ((void)0)
#define SIMT_PREFETCH_L1(ptr) ((void)0)
#define SIMT_OPAQUE_U32(x) ((void)0)
#define SIMT_RSQRT(x) (1.0 / sqrt(x))
#define SIMT_PDL_ENTRY() ((void)0)
#define SIMT_MBAR_INIT(mbar, count) ((void)0)
#define SIMT_BULK_G2S(dst, src, bytes, mbar) memcpy((dst), (src), (bytes))
#define SIMT_MBAR_WAIT(mbar, parity) ((void)0)
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
#define __restrict__
#define SIMT_KERNEL(name, params_t, body_call)                                   \
    static void name(const params_t& P, unsigned char* simt_smem) { body_call; }

#endif
