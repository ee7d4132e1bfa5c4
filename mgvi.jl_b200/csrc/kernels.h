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
