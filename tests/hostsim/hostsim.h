// hostsim.h -- TEST INFRASTRUCTURE ONLY.
//
// A single-OS-thread fibre emulator for the SPMD execution model the CUDA kernels in
// sbayes_b200/csrc are written against (CTAs of threads, 32-wide warps, block barriers,
// warp shuffles, global atomics, dynamic shared memory).  The kernels touch the machine
// only through the `sb_*` wrappers of sb_platform.cuh; compiling the same source with
// -DSB_HOSTSIM binds those wrappers to this file so that the *control flow* of the step
// kernel (phase structure, barrier placement, uniformity of team decisions) can be
// checked against the oracle in the GPU-less build container before GPU minutes are
// spent.  It is never loaded by the product (`sbayes_b200/backend.py` only ever opens
// libsbayes_b200.so) and is built into tests/_hostsim/ by the tests that use it.
#pragma once
#include <ucontext.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <vector>
#include <functional>

namespace hostsim {

struct WarpSlot { uint64_t val[32]; int arrived; unsigned gen; };
struct Cta {
    int arrived = 0; unsigned gen = 0; char *smem = nullptr; std::vector<WarpSlot> warps;
};
struct Fiber {
    ucontext_t ctx; char *stack = nullptr; int tid = 0, cta = 0; bool done = false;
};
struct Grid {
    int n_cta = 0, n_thr = 0; std::vector<Cta> ctas; std::vector<Fiber> fibers;
    ucontext_t sched; Fiber *cur = nullptr; std::function<void()> body;
};
inline Grid *&grid() { static Grid *g = nullptr; return g; }

inline void yield() { Grid *g = grid(); swapcontext(&g->cur->ctx, &g->sched); }

inline void trampoline() {
    Grid *g = grid();
    g->body();
    g->cur->done = true;
    swapcontext(&g->cur->ctx, &g->sched);
}

inline void launch(std::function<void()> body, int n_cta, int n_thr, size_t smem_bytes) {
    Grid G; G.n_cta = n_cta; G.n_thr = n_thr; G.body = body;
    G.ctas.resize(n_cta);
    for (auto &c : G.ctas) {
        c.smem = (char *)aligned_alloc(128, ((smem_bytes + 127) / 128 + 1) * 128);
        memset(c.smem, 0xA5, smem_bytes);
        c.warps.resize((n_thr + 31) / 32);
        for (auto &w : c.warps) { w.arrived = 0; w.gen = 0; }
    }
    const size_t STK = 256 * 1024;
    G.fibers.resize((size_t)n_cta * n_thr);
    for (int c = 0; c < n_cta; c++)
        for (int t = 0; t < n_thr; t++) {
            Fiber &f = G.fibers[(size_t)c * n_thr + t];
            f.tid = t; f.cta = c; f.stack = (char *)malloc(STK);
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STK; f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, (void (*)())trampoline, 0);
        }
    Grid *prev = grid();
    grid() = &G;
    size_t alive = G.fibers.size();
    unsigned long rounds = 0;
    // SB_HOSTSIM_SKEW=k: before every round one randomly favoured CTA runs k extra rounds on
    // its own, so CTAs of a team drift apart as far as the barriers allow (exposes missing
    // barriers between a fast CTA's writes and a slow CTA's reads).
    const char *sk = getenv("SB_HOSTSIM_SKEW");
    const int skew = sk ? atoi(sk) : 0;
    unsigned long long lcg = 0x9E3779B97F4A7C15ull;
    while (alive) {
        if (skew > 0) {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            int fav = (int)((lcg >> 33) % (unsigned)n_cta);
            for (int r = 0; r < skew; r++)
                for (int t = 0; t < n_thr; t++) {
                    Fiber &f = G.fibers[(size_t)fav * n_thr + t];
                    if (f.done) continue;
                    G.cur = &f;
                    swapcontext(&G.sched, &f.ctx);
                }
        }
        alive = 0;
        for (auto &f : G.fibers) {
            if (f.done) continue;
            G.cur = &f;
            swapcontext(&G.sched, &f.ctx);
            if (!f.done) alive++;
        }
        if (++rounds > 2000000000UL) { fprintf(stderr, "hostsim: livelock\n"); abort(); }
    }
    grid() = prev;
    for (auto &f : G.fibers) free(f.stack);
    for (auto &c : G.ctas) free(c.smem);
}

inline int tid() { return grid()->cur->tid; }
inline int cta() { return grid()->cur->cta; }
inline int nthr() { return grid()->n_thr; }
inline int ncta() { return grid()->n_cta; }
inline char *smem() { return grid()->ctas[cta()].smem; }

inline void cta_sync() {
    Grid *g = grid(); Cta &c = g->ctas[cta()];
    unsigned gen = c.gen;
    if (++c.arrived == g->n_thr) { c.arrived = 0; c.gen++; }
    else while (c.gen == gen) yield();
}
inline void warp_sync() {
    Grid *g = grid(); WarpSlot &w = g->ctas[cta()].warps[tid() / 32];
    unsigned gen = w.gen;
    if (++w.arrived == 32) { w.arrived = 0; w.gen++; }
    else while (w.gen == gen) yield();
}
inline uint64_t shfl(uint64_t v, int src_lane) {
    Grid *g = grid(); WarpSlot &w = g->ctas[cta()].warps[tid() / 32];
    w.val[tid() % 32] = v;
    warp_sync();
    uint64_t r = w.val[src_lane & 31];
    warp_sync();
    return r;
}
inline unsigned ballot(bool p) {
    Grid *g = grid(); WarpSlot &w = g->ctas[cta()].warps[tid() / 32];
    w.val[tid() % 32] = p ? 1 : 0;
    warp_sync();
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (unsigned)(w.val[i] & 1) << i;
    warp_sync();
    return r;
}
}  // namespace hostsim
