// __global__ entry points (or their emulator stand-ins) for the kernel bodies.
#pragma once
#include "vec_ops.h"
#include "big1d_pass.h"

#if defined(__CUDACC__) && !defined(MGVI_SIMT_EMU)
#define SIMT_GLOBAL(maxt) __global__ void __launch_bounds__(maxt)
#define SIMT_KARGS(T) const T P
#define SIMT_SMEM_DECL extern __shared__ __align__(16) unsigned char simt_smem[];
#else
#define SIMT_GLOBAL(maxt) static void
#define SIMT_KARGS(T) const T &P, unsigned char *simt_smem
#define SIMT_SMEM_DECL
#endif

namespace mgvi {

#if defined(__CUDACC__) && !defined(MGVI_SIMT_EMU)
#define SIMT_GLOBAL2(maxt, minb) __global__ void __launch_bounds__(maxt, minb)
#else
#define SIMT_GLOBAL2(maxt, minb) static void
#endif

// MAXT = 256: up to 256 threads, MINB resident blocks per SM wanted; MAXT = 1024: wide tiles.
// LG > 0 / LGL >= 0: log2 of the transform length / of the tile width fixed at compile time (hot
// sizes), LOOP = false: one vector per block; LG = 0, LGL = -1, LOOP = true: everything at run time.
template <int GEO, int PRO, int EPI, bool DOUBLE, int MAXT, int MINB, int LG, int LGL, bool LOOP, int PF = 0>
SIMT_GLOBAL2(MAXT, MINB) k_field_pass(SIMT_KARGS(FieldPassParams)) {
    SIMT_SMEM_DECL
    SIMT_PDL_ENTRY();
    field_pass_body<GEO, PRO, EPI, DOUBLE, LG, LGL, LOOP, PF>(P, simt_smem);
}

// MAXT = 256: blocks of up to 256 threads, two resident per SM (128 registers: the run-time-shape body with
// its 8-element operand batches spills at the 64 registers a 1024-thread bound allows); MAXT = 1024: long lines.
template <int KIND, int PRO, int EPI, int MAXT = 1024>
SIMT_GLOBAL2(MAXT, (MAXT <= 256 ? 2 : 1)) k_big1d(SIMT_KARGS(BigParams)) {
    SIMT_SMEM_DECL
    SIMT_PDL_ENTRY();
    big1d_body<KIND, PRO, EPI>(P, simt_smem);
}

// ------------------------------------------------------------------------------------------
// Persistent lock-step CG for long 1-D fields whose whole CG state is L2-resident (C2: 2^16 pixels,
// 16 samples = 8 MB per array): ONE cooperative launch runs every remaining iteration.  A resident
// grid walks over the block indices of the four phases of an iteration (column FFTs with p = r + beta p
// | row FFTs, weight, row FFTs | column FFTs, + p, p.Ap | x += alpha p, r -= alpha Ap, r.r) with a grid
// barrier between phases instead of a kernel boundary; the per-sample scalar tails stay where they are
// ("last block of the sample", field_pass.h), so the arithmetic -- and every result bit -- is that of the
// host-launched sequence.  Product build only (the fiber emulator runs blocks one after another).
// ------------------------------------------------------------------------------------------
struct BigCgParams {
    BigParams a, b, ap;
    VecParams u;
    long long grid_a, grid_b, grid_ap, grid_u;
    unsigned* bar;              // grid barrier counter, zero at launch
    long long max_iters;        // lock-step iterations this launch may run (the device stops every sample at itmax itself)
};

#if defined(__CUDACC__) && !defined(MGVI_SIMT_EMU)
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                       // release: this block's writes (cumulative over bar.sync)
        atomicAdd(bar, 1u);
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
        } while ((int)(seen - target) < 0);
        __threadfence();
    }
    __syncthreads();
}

// The parameters live in DEVICE memory, not in the kernel's parameter space: loads from the parameter bank are
// free to be hoisted out of the iteration loop, which keeps dozens of pointers live across all four phases
// (560 bytes of spill at 128 registers); loads from global memory cannot cross the barrier's memory clobber,
// so every phase starts with an empty register file like the stand-alone kernel does.
template <int UNUSED>      // a template so that the header can be included by several translation units
__global__ void __launch_bounds__(256, 2) k_big1d_cg_persist(const BigCgParams* __restrict__ Qp) {
    extern __shared__ __align__(16) unsigned char simt_smem[];
    const unsigned nb = gridDim.x;
    unsigned gen = 0;
    const long long max_iters = Qp->max_iters;
    for (long long it = 0; it < max_iters; ++it) {
        // uniform over the grid: n_active only changes in the update phase, three barriers away
        int na;
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(na) : "l"(Qp->u.cg.n_active) : "memory");
        if (na <= 0) break;
#pragma unroll 1
        for (long long vb = blockIdx.x; vb < Qp->grid_a; vb += nb) {
            asm volatile("" : "+l"(vb) :: "memory");
            big1d_body<BIG_A, PRO_CG, EPI_STORE>(Qp->a, simt_smem, vb);
            __syncthreads();
        }
        grid_barrier(Qp->bar, (++gen) * nb);
#pragma unroll 1
        for (long long vb = blockIdx.x; vb < Qp->grid_b; vb += nb) {
            asm volatile("" : "+l"(vb) :: "memory");
            big1d_body<BIG_B_MID, PRO_LOAD, EPI_STORE>(Qp->b, simt_smem, vb);
            __syncthreads();
        }
        grid_barrier(Qp->bar, (++gen) * nb);
#pragma unroll 1
        for (long long vb = blockIdx.x; vb < Qp->grid_ap; vb += nb) {
            asm volatile("" : "+l"(vb) :: "memory");
            big1d_body<BIG_AP, PRO_LOAD, EPI_METRIC>(Qp->ap, simt_smem, vb);
            __syncthreads();
        }
        grid_barrier(Qp->bar, (++gen) * nb);
#pragma unroll 1
        for (long long vb = blockIdx.x; vb < Qp->grid_u; vb += nb) {
            asm volatile("" : "+l"(vb) :: "memory");
            vec_body(Qp->u, simt_smem, vb);
            __syncthreads();
        }
        grid_barrier(Qp->bar, (++gen) * nb);
    }
}
#endif

template <int PRO>
SIMT_GLOBAL(256) k_elem_pro(SIMT_KARGS(ElemParams)) {
    SIMT_SMEM_DECL
    elem_pro_body<PRO>(P, simt_smem);
}

template <int EPI>
SIMT_GLOBAL(256) k_elem_epi(SIMT_KARGS(ElemParams)) {
    SIMT_SMEM_DECL
    elem_epi_body<EPI>(P, simt_smem);
}

}  // namespace mgvi
