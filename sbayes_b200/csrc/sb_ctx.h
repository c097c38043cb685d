// sb_ctx.h -- the host-side context behind the C ABI (include/sbayes_b200.h), shared by the translation units of
// the library: sb_api.cu (problem upload, chain state, the persistent step kernel, MC3 swap) and sb_api_eval.cu
// (the stand-alone streaming kernels: statistics pass, terms, observation likelihood, Gibbs sweep, initializer EM,
// effect draws).  The split keeps the slow compile of the step kernel out of the edit cycle of the others.
#pragma once
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <map>
#include <functional>
#include <cstddef>
#include <unistd.h>
#include "sb_platform.cuh"
#include "sb_types.h"
#include "sb_plans.h"

#ifdef SB_HOSTSIM
#define SB_API extern "C" __attribute__((visibility("default")))
typedef void *sb_stream_t;
#else
#define SB_API extern "C" __attribute__((visibility("default")))
typedef cudaStream_t sb_stream_t;
#endif

extern std::string g_create_error;          // defined in sb_api.cu

struct sb_ctx {
    int device = 0, n_chains = 0, team = 1, n_teams = 1, grid = 1, block = 256, n_sm = 1, smem_members = 0, clock_khz = 1000000;
    size_t smem = 0;
    long long launches = 0;
    SbProblem P;
    std::vector<SbShared> h_states;
    SbShared *d_states = nullptr;
    SbPlace place;
    char *d_rep = nullptr;
    int use_cluster = 0, req_team = 0;
    size_t rep_alloc = 0;
    const void *kernel = nullptr;
    std::function<int(int, int)> place_fn;
    std::vector<sb_chain_cfg> h_cfg;
    std::vector<int32_t> cg_ptr, cg_idx;
    SbSlot *d_slots = nullptr;
    unsigned long long *d_bars = nullptr;
    double *d_tred = nullptr;
    // everything another GPU may have to reach when the objects of a chain are sharded over GPUs (per-chain
    // state, reduction buffer, flags, statistics blocks) is carved from ONE allocation, the arena: a peer maps
    // it once (CUDA IPC) and finds every array at the same offset as in its own arena
    char *arena = nullptr; size_t arena_size = 0, arena_used = 0; bool arena_dry = false;
    int xrank = 0, xworld = 1;
    char *peer_arena[SB_MAX_RANKS] = {nullptr};
    bool peer_opened[SB_MAX_RANKS] = {false};
    unsigned long long xlaunches = 0;
    unsigned long long *d_xflag = nullptr, *d_lrelease = nullptr;
    int32_t *d_lstat = nullptr, *d_gstat = nullptr, *d_lcl = nullptr;
    uint8_t *d_losel = nullptr;
    int *d_xabort = nullptr;
    double *d_glh_all = nullptr, *d_geo_all = nullptr;   // [n_chains, G_tot], [n_chains, K] export copies, one allocation each
    SbExport exp_;                       // staging buffers of sb_export_samples (allocated on first use)
    int exp_ok = 0;
    int32_t *d_exp_idx = nullptr;
    double *d_post_all = nullptr;        // [n_chains, 8] the post blocks of all states, one allocation
    std::vector<int> h_slot_state;       // host mirror of SbSlot::state (changed only by sb_swap / sb_set_slots)
    uint8_t *d_src_tmp = nullptr;
    int32_t *d_counts_tmp = nullptr;
    // stand-alone evaluation (sb_eval.cuh): the int32 statistics of all states are one block
    int32_t *d_stats = nullptr;
    size_t stats_stride = 0;          // int32 per state: counts | sizes | H, padded to 16 B
    SbEvalPlan eplan;
    int eplan_ok = 0;
    SbMmaPlan mplan;                  // tcgen05 statistics pass over all states
    int mplan_ok = 0, stats_mode = 0; // 0 = shared-memory histogram, 1 = int8 tensor-core contraction
    double *d_epartial = nullptr, *d_egeo = nullptr, *d_epost = nullptr;
    unsigned *d_eticket = nullptr;
    int *d_eerr = nullptr;
    // observation likelihood (allocated by the first sb_obs_loglik)
    SbObsPlan oplan;
    float *d_obs_out = nullptr;
    double *d_obs_part = nullptr, *d_obs_ll = nullptr, *d_obs_sums = nullptr;
    unsigned long long *d_gibbs_changed = nullptr;
    void *eval_ext = nullptr;                        // private state of sb_api_eval.cu (tables of the fine-row statistics pass)
    void (*eval_ext_free)(void *) = nullptr;
    std::vector<void *> allocs;
    sb_stream_t stream = 0;
    std::string err;
};

// ---------------------------------------------------------------------------------------
// device memory shim
// ---------------------------------------------------------------------------------------
#ifdef SB_HOSTSIM
static inline int dev_alloc(sb_ctx *c, void **p, size_t n) { *p = calloc(1, n ? n : 1); c->allocs.push_back(*p); return *p ? 0 : -1; }
static inline void dev_free_all(sb_ctx *c) { for (void *p : c->allocs) free(p); c->allocs.clear(); }
static inline int h2d(sb_ctx *, void *d, const void *h, size_t n) { memcpy(d, h, n); return 0; }
static inline int d2h(sb_ctx *, void *h, const void *d, size_t n) { memcpy(h, d, n); return 0; }
static inline int dev_zero(sb_ctx *, void *d, size_t n) { memset(d, 0, n); return 0; }
static inline int dev_sync(sb_ctx *) { return 0; }
#else
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            c->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return -2;                                                                             \
        }                                                                                          \
    } while (0)
static inline int dev_alloc(sb_ctx *c, void **p, size_t n) {
    CK(cudaMalloc(p, n ? n : 1));
    c->allocs.push_back(*p);
    CK(cudaMemsetAsync(*p, 0, n ? n : 1, c->stream));
    return 0;
}
static inline void dev_free_all(sb_ctx *c) { for (void *p : c->allocs) cudaFree(p); c->allocs.clear(); }
static inline int h2d(sb_ctx *c, void *d, const void *h, size_t n) { CK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, c->stream)); CK(cudaStreamSynchronize(c->stream)); return 0; }
static inline int d2h(sb_ctx *c, void *h, const void *d, size_t n) { CK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream)); return 0; }
static inline int dev_zero(sb_ctx *c, void *d, size_t n) { CK(cudaMemsetAsync(d, 0, n, c->stream)); return 0; }
static inline int dev_sync(sb_ctx *c) { CK(cudaStreamSynchronize(c->stream)); return 0; }
#endif

#define TRY(x) do { int r_ = (x); if (r_) return r_; } while (0)
template <class T> static inline int up(sb_ctx *c, const T **dst, const std::vector<T> &v) {
    void *p; TRY(dev_alloc(c, &p, v.size() * sizeof(T)));
    if (!v.empty()) TRY(h2d(c, p, v.data(), v.size() * sizeof(T)));
    *dst = (const T *)p; return 0;
}
template <class T> static inline int mk(sb_ctx *c, T **dst, size_t n) { void *p; TRY(dev_alloc(c, &p, n * sizeof(T))); *dst = (T *)p; return 0; }

// carve n elements from the arena (256-byte aligned); in the dry pass only the size is accumulated
template <class T> static inline int amk(sb_ctx *c, T **dst, size_t n) {
    const size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
    if (!c->arena_dry) {
        if (c->arena_used + bytes > c->arena_size) { c->err = "internal: arena overflow"; return -1; }
        *dst = (T *)(c->arena + c->arena_used);
    } else *dst = nullptr;
    c->arena_used += bytes;
    return 0;
}

static inline int fail(sb_ctx *c, const char *msg) { if (c) c->err = msg; else g_create_error = msg; return -1; }
static inline int check_chain(sb_ctx *c, int chain) {
    if (!c) return -1;
    if (chain < 0 || chain >= c->n_chains) return fail(c, "chain index out of range");
#ifndef SB_HOSTSIM
    CK(cudaSetDevice(c->device));
#endif
    return 0;
}
static inline int get_slot(sb_ctx *c, int chain, SbSlot *s) { return d2h(c, s, c->d_slots + chain, sizeof(SbSlot)); }

// sb_api_eval.cu: plans, kernel attributes and scratch buffers of the stand-alone kernels (called by sb_create)
int sb_eval_setup(sb_ctx *c);

