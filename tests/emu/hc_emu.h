// hc_emu.h -- host-side SIMT emulator.  TEST HARNESS ONLY.
//
// Compiled only with -DHC_EMU into tests/emu/_build/libhc_emu.so, which lets the CPU test
// suite execute the *same kernel source* (one ucontext fiber per CUDA thread, CTAs run one
// after another) on tiny inputs, so that logic errors are found before GPU minutes are spent.
// It is never part of libhc_b200.so and the package never loads it: the product has no CPU
// path.  Limitations (by design): no inter-CTA waiting, 1-D blocks, full-warp collectives.
#pragma once
#ifndef HC_EMU
#error "hc_emu.h is only for the -DHC_EMU test build"
#endif

#include <ucontext.h>
#ifdef HC_EMU_DEBUG
#include <dlfcn.h>
#endif

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct hc_emu_idx { unsigned x, y, z; };

inline hc_emu_idx threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace hc_emu {

constexpr size_t kStack = 256 * 1024;

struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = false;
};

struct Warp {
    uint32_t live = 0, arrived = 0, gen = 0;
    uint64_t buf[32], snap[32];
    void *site[32];          // HC_EMU_DEBUG: caller of the collective each lane is waiting in
};

struct State {
    ucontext_t main;
    std::vector<Fiber> fibers;
    std::vector<char *> stacks;
    Warp warps[32];
    unsigned nthreads = 0, live = 0, cur = 0;
    unsigned bar_arrived = 0, bar_gen = 0;
    std::function<void()> body;
    std::vector<char> dyn;
};

inline State &st()
{
    static State s;
    return s;
}

inline char *dyn_smem() { return st().dyn.data(); }

inline void yield()
{
    State &s = st();
    swapcontext(&s.fibers[s.cur].ctx, &s.main);
}

inline void bar_release_if_complete()
{
    State &s = st();
    if (s.bar_arrived > 0 && s.bar_arrived >= s.live) {
        s.bar_arrived = 0;
        s.bar_gen++;
    }
}

inline void warp_release_if_complete(Warp &w)
{
    if (w.arrived != 0 && (w.arrived & w.live) == w.live) {
#ifdef HC_EMU_DEBUG
        // every lane of a converged warp must be inside the SAME collective call
        void *first = nullptr;
        for (int i = 0; i < 32; i++)
            if ((w.live >> i) & 1) {
                if (!first) first = w.site[i];
                else if (w.site[i] != first) {
                    fprintf(stderr, "hc_emu: lanes of a warp wait in different collectives:");
                    for (int j = 0; j < 32; j++)
                        if (((w.live >> j) & 1) && (j < 2 || w.site[j] != w.site[j - 1])) {
                            Dl_info di;
                            if (dladdr(w.site[j], &di) && di.dli_sname)
                                fprintf(stderr, "\n  lane %d: %s+0x%lx (lib offset 0x%lx)", j, di.dli_sname, (unsigned long)((char *)w.site[j] - (char *)di.dli_saddr),
                                        (unsigned long)((char *)w.site[j] - (char *)di.dli_fbase));
                            else
                                fprintf(stderr, "\n  lane %d: %p (lib offset 0x%lx)", j, w.site[j], di.dli_fbase ? (unsigned long)((char *)w.site[j] - (char *)di.dli_fbase) : 0ul);
                        }
                    fprintf(stderr, "\n");
                    abort();
                }
            }
#endif
        memcpy(w.snap, w.buf, sizeof w.snap);
        w.arrived = 0;
        w.gen++;
    }
}

inline void trampoline()
{
    State &s = st();
    s.body();
    unsigned tid = s.cur;
    s.fibers[tid].done = true;
    s.live--;
    Warp &w = s.warps[tid / 32];
    w.live &= ~(1u << (tid % 32));
    warp_release_if_complete(w);
    bar_release_if_complete();
    swapcontext(&s.fibers[tid].ctx, &s.main);
}

inline void syncthreads()
{
    State &s = st();
    unsigned g = s.bar_gen;
    s.bar_arrived++;
    bar_release_if_complete();
    while (s.bar_gen == g) yield();
}

// all live lanes of the calling warp exchange one 64-bit value; returns pointer to the snapshot
inline const uint64_t *warp_exchange(uint64_t v, void *site = nullptr)
{
    State &s = st();
    unsigned tid = s.cur;
    Warp &w = s.warps[tid / 32];
    unsigned g = w.gen;
    w.buf[tid % 32] = v;
    w.site[tid % 32] = site;
    w.arrived |= 1u << (tid % 32);
    warp_release_if_complete(w);
    while (w.gen == g) yield();
    return w.snap;
}

inline unsigned lane() { return st().cur % 32; }
inline uint32_t warp_live() { return st().warps[st().cur / 32].live; }

inline void run_block(unsigned nthreads, const std::function<void()> &body)
{
    State &s = st();
    if (nthreads > 1024 || nthreads % 32 != 0) {
        fprintf(stderr, "hc_emu: block size %u unsupported\n", nthreads);
        abort();
    }
    s.body = body;
    s.nthreads = s.live = nthreads;
    s.bar_arrived = 0;
    if (s.fibers.size() < nthreads) s.fibers.resize(nthreads);
    while (s.stacks.size() < nthreads) s.stacks.push_back((char *)malloc(kStack));
    for (unsigned w = 0; w < 32; w++) {
        s.warps[w].arrived = 0;
        unsigned lo = w * 32;
        s.warps[w].live = lo >= nthreads ? 0u : (nthreads - lo >= 32 ? 0xffffffffu : ((1u << (nthreads - lo)) - 1));
    }
    for (unsigned t = 0; t < nthreads; t++) {
        Fiber &f = s.fibers[t];
        f.done = false;
        f.stack = s.stacks[t];
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    while (s.live > 0) {
        for (unsigned t = 0; t < nthreads; t++) {
            if (s.fibers[t].done) continue;
            s.cur = t;
            threadIdx.x = t;
            threadIdx.y = threadIdx.z = 0;
            swapcontext(&s.main, &s.fibers[t].ctx);
        }
    }
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
    // one emulated launch at a time (the emulator state is global); host threads of hc_pipeline take turns
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    State &s = st();
    gridDim = grid;
    blockDim = block;
    if (s.dyn.size() < smem + 16) s.dyn.resize(smem + 16);
    for (unsigned by = 0; by < grid.y; by++)
        for (unsigned bx = 0; bx < grid.x; bx++) {
            blockIdx.x = bx;
            blockIdx.y = by;
            blockIdx.z = 0;
            run_block(block.x, body);
        }
}

}  // namespace hc_emu
