// Single-threaded SIMT emulator (TEST INFRASTRUCTURE ONLY -- never part of libmgvib200.so).
//
// Runs a CUDA-style kernel body, written against mgvi.jl_b200/csrc/simt.h, on the CPU: every
// thread of a block is a ucontext fiber; __syncthreads() and warp shuffles are cooperative
// barriers of a round-robin scheduler; blocks run one after another.  Deterministic, no OS
// threads.  Its only purpose is to let `pytest -m "not gpu"` exercise the index arithmetic of
// the real kernel sources on a box without a GPU.
#pragma once
#include <ucontext.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <vector>

namespace simt_emu {

enum State { RUNNABLE, WAIT_BLOCK, WAIT_WARP, DONE };

struct Fiber {
    ucontext_t ctx;
    State state = RUNNABLE;
    int tid = 0;
    char* stack = nullptr;
};

struct Runtime {
    ucontext_t sched_ctx;
    std::vector<Fiber> fibers;
    std::vector<char> stacks;
    int cur = -1;
    int ntid = 0;
    long long bid = 0, nbid = 0;
    std::function<void()> body;
    double warp_f64[64][32];
    double warp_f64b[64][32];
    int warp_i32[64][32];
    static constexpr size_t kStack = 256 * 1024;
};

inline Runtime& rt() { static Runtime r; return r; }

inline int tid() { return rt().fibers[rt().cur].tid; }
inline int ntid() { return rt().ntid; }
inline long long bid() { return rt().bid; }
inline long long nbid() { return rt().nbid; }

inline void yield_to_sched(State s) {
    Runtime& r = rt();
    Fiber& f = r.fibers[r.cur];
    f.state = s;
    swapcontext(&f.ctx, &r.sched_ctx);
}

inline void block_barrier() { yield_to_sched(WAIT_BLOCK); }
inline void warp_barrier() { yield_to_sched(WAIT_WARP); }

inline double shfl_xor_f64(double v, int mask) {
    Runtime& r = rt();
    int t = tid(), w = t / 32, l = t % 32;
    r.warp_f64[w][l] = v;
    warp_barrier();
    int src = l ^ mask;
    int warp_size = (r.ntid - w * 32) < 32 ? (r.ntid - w * 32) : 32;
    double out = (src < warp_size) ? r.warp_f64[w][src] : v;
    warp_barrier();
    return out;
}

inline double shfl_idx_f64(double v, int lane) {
    Runtime& r = rt();
    int t = tid(), w = t / 32, l = t % 32;
    r.warp_f64[w][l] = v;
    warp_barrier();
    double out = r.warp_f64[w][lane];
    warp_barrier();
    return out;
}

// all lanes publish (a, b); every lane gets the whole warp's values (2 barriers instead of 2 per shuffle)
inline void warp_allgather2(double a, double b, double (&oa)[32], double (&ob)[32]) {
    Runtime& r = rt();
    int t = tid(), w = t / 32, l = t % 32;
    r.warp_f64[w][l] = a;
    r.warp_f64b[w][l] = b;
    warp_barrier();
    for (int i = 0; i < 32; ++i) { oa[i] = r.warp_f64[w][i]; ob[i] = r.warp_f64b[w][i]; }
    warp_barrier();
}

inline int shfl_idx_i32(int v, int lane) {
    Runtime& r = rt();
    int t = tid(), w = t / 32, l = t % 32;
    r.warp_i32[w][l] = v;
    warp_barrier();
    int out = r.warp_i32[w][lane];
    warp_barrier();
    return out;
}

inline unsigned atomic_inc_u32(unsigned* p) { unsigned o = *p; *p = o + 1; return o; }
inline int atomic_add_i32(int* p, int v) { int o = *p; *p = o + v; return o; }

inline void fiber_entry() {
    Runtime& r = rt();
    r.body();
    r.fibers[r.cur].state = DONE;
    swapcontext(&r.fibers[r.cur].ctx, &r.sched_ctx);
}

// Run `body` for every (block, thread).  `body` reads tid()/bid() and may call the barriers.
inline void launch(long long nblocks, int nthreads, const std::function<void()>& body) {
    Runtime& r = rt();
    if (nthreads <= 0 || nthreads > 1024) throw std::runtime_error("simt_emu: bad block size");
    r.ntid = nthreads;
    r.nbid = nblocks;
    r.body = body;
    if (r.stacks.size() < (size_t)nthreads * Runtime::kStack) r.stacks.resize((size_t)nthreads * Runtime::kStack);
    r.fibers.resize(nthreads);
    for (long long b = 0; b < nblocks; ++b) {
        r.bid = b;
        for (int t = 0; t < nthreads; ++t) {
            Fiber& f = r.fibers[t];
            f.tid = t;
            f.state = RUNNABLE;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = r.stacks.data() + (size_t)t * Runtime::kStack;
            f.ctx.uc_stack.ss_size = Runtime::kStack;
            f.ctx.uc_link = &r.sched_ctx;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        int done = 0;
        while (done < nthreads) {
            bool progressed = false;
            for (int t = 0; t < nthreads; ++t) {
                Fiber& f = r.fibers[t];
                if (f.state != RUNNABLE) continue;
                r.cur = t;
                swapcontext(&r.sched_ctx, &f.ctx);
                progressed = true;
                if (f.state == DONE) ++done;
            }
            // release barriers
            int nwait_block = 0, nlive = 0;
            for (auto& f : r.fibers) { if (f.state != DONE) ++nlive; if (f.state == WAIT_BLOCK) ++nwait_block; }
            if (nlive > 0 && nwait_block == nlive) {
                for (auto& f : r.fibers) if (f.state == WAIT_BLOCK) f.state = RUNNABLE;
                progressed = true;
            }
            int nwarps = (nthreads + 31) / 32;
            for (int w = 0; w < nwarps; ++w) {
                int lo = w * 32, hi = lo + 32 < nthreads ? lo + 32 : nthreads;
                int live = 0, waiting = 0;
                for (int t = lo; t < hi; ++t) {
                    if (r.fibers[t].state != DONE) ++live;
                    if (r.fibers[t].state == WAIT_WARP) ++waiting;
                }
                if (live > 0 && waiting == live) {
                    for (int t = lo; t < hi; ++t) if (r.fibers[t].state == WAIT_WARP) r.fibers[t].state = RUNNABLE;
                    progressed = true;
                }
            }
            if (!progressed) throw std::runtime_error("simt_emu: deadlock (divergent barrier)");
        }
    }
    r.cur = -1;
}

}  // namespace simt_emu
