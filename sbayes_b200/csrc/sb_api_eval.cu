// sb_api_eval.cu -- C ABI of the stand-alone streaming kernels (include/sbayes_b200.h): statistics pass and terms
// (sb_eval.cuh), int8 tensor-core count rebuild (sb_counts.cuh), observation likelihood, full-sweep source Gibbs,
// initializer EM (sb_init.cuh).  Its own translation unit: these kernels do not need the step kernel (sb_core.cuh).
#include "sb_ctx.h"
#include "sb_philox.cuh"
#include "sb_counts.cuh"
#include "sb_eval.cuh"
#include "sb_init.cuh"

// ---------------------------------------------------------------------------------------
// fine-row statistics pass (sb_eval.cuh: ev_hist_fine): row tables, the NA-mapped copy of the state bytes, launch plan
// ---------------------------------------------------------------------------------------
struct FineState {
    SbFineArgs fa; int ok = 0; const void *kernel = nullptr;
    double *d_eff = nullptr; size_t eff_cap = 0; int32_t *d_eff_chains = nullptr;    // staging of sb_effect_sample
};
static void fine_free(void *p) { delete (FineState *)p; }
static FineState *fine_of(sb_ctx *c) { return (FineState *)c->eval_ext; }

#ifndef SB_HOSTSIM
static const void *fine_kernel_ptr(int S1) {
    switch (S1) {
    case 2: return (const void *)sb_eval_fine_kernel<2>; case 3: return (const void *)sb_eval_fine_kernel<3>;
    case 4: return (const void *)sb_eval_fine_kernel<4>; case 5: return (const void *)sb_eval_fine_kernel<5>;
    case 6: return (const void *)sb_eval_fine_kernel<6>; case 7: return (const void *)sb_eval_fine_kernel<7>;
    case 8: return (const void *)sb_eval_fine_kernel<8>; case 9: return (const void *)sb_eval_fine_kernel<9>;
    }
    return nullptr;
}
#else
static void fine_kernel_host(int S1, const SbEvalArgs &A, const SbFineArgs &FA) {
    switch (S1) {
    case 2: sb_eval_fine_kernel<2>(A, FA); break; case 3: sb_eval_fine_kernel<3>(A, FA); break;
    case 4: sb_eval_fine_kernel<4>(A, FA); break; case 5: sb_eval_fine_kernel<5>(A, FA); break;
    case 6: sb_eval_fine_kernel<6>(A, FA); break; case 7: sb_eval_fine_kernel<7>(A, FA); break;
    case 8: sb_eval_fine_kernel<8>(A, FA); break; case 9: sb_eval_fine_kernel<9>(A, FA); break;
    }
}
#endif

static int fine_setup(sb_ctx *c) {
    const SbProblem &P = c->P;
    FineState *fs = new FineState();
    memset(&fs->fa, 0, sizeof fs->fa);
    c->eval_ext = fs; c->eval_ext_free = fine_free;
    const int N = P.N, C = P.C, K = P.K, np = P.n_pat, S1 = P.S + 1;
    if (!P.obj_pack || S1 < 2 || S1 > 9 || getenv("SB_EVAL_NO_FINE")) return 0;
    std::vector<uint32_t> pack(N);
    std::vector<int32_t> pat_bits(np);
    TRY(d2h(c, pack.data(), P.obj_pack, sizeof(uint32_t) * N));
    TRY(d2h(c, pat_bits.data(), P.pat_bits, sizeof(int32_t) * np));
    // fine rows: [0, K*np) (cluster, pattern); [K*np, (K+1)*np) (free, pattern); then one pair (not in a cluster, in a
    // cluster) per distinct (component, group or none, pattern); last the trash pair (components the problem does not have)
    std::map<uint32_t, int> pair_of;
    int next = (K + 1) * np;
    auto key_of = [](int u, uint32_t g, uint32_t pat) { return ((uint32_t)u << 24) | (g << 8) | pat; };
    for (int pat = 0; pat < np; pat++)                 // a component the pattern lacks: counts nowhere, but H sees it
        for (int u = 1; u < C; u++) if (!((pat_bits[pat] >> u) & 1)) { pair_of[key_of(u, 255u, (uint32_t)pat)] = next; next += 2; }
    for (int o = 0; o < N; o++) {
        const uint32_t pat = pack[o] & 255u;
        for (int u = 1; u < C && u < 4; u++) {
            const uint32_t k = key_of(u, (pack[o] >> (8 * u)) & 255u, pat);
            if (!pair_of.count(k)) { pair_of[k] = next; next += 2; }
        }
    }
    const int trash = next, R = next + 2;
    if (trash + 1 > 255) return 0;                     // row ids travel as bytes: the two-increment pass serves
    std::vector<uint32_t> fine(N);
    for (int o = 0; o < N; o++) {
        const uint32_t pat = pack[o] & 255u;
        uint32_t w = pat;
        for (int u = 1; u < 4; u++) {
            int row = trash;
            if (u < C) row = pair_of[key_of(u, (pack[o] >> (8 * u)) & 255u, pat)];
            w |= (uint32_t)row << (8 * u);
        }
        fine[o] = w;
    }
    std::vector<std::vector<int32_t>> cl(P.G_tot), hl((size_t)np * 2 * C);
    for (int k = 0; k < K; k++)
        for (int p = 0; p < np; p++) { cl[k].push_back(k * np + p); hl[(size_t)(p * 2 + 1) * C].push_back(k * np + p); }
    for (int p = 0; p < np; p++) hl[(size_t)(p * 2) * C].push_back(K * np + p);
    for (auto &kv : pair_of) {
        const int u = (int)(kv.first >> 24), g = (int)((kv.first >> 8) & 255u), pat = (int)(kv.first & 255u);
        for (int in = 0; in < 2; in++) {
            if (g != 255) cl[g].push_back(kv.second + in);
            hl[(size_t)(pat * 2 + in) * C + u].push_back(kv.second + in);
        }
    }
    auto csr = [](const std::vector<std::vector<int32_t>> &l, std::vector<int32_t> &ptr, std::vector<int32_t> &idx) {
        ptr.assign(1, 0);
        for (auto &v : l) { idx.insert(idx.end(), v.begin(), v.end()); ptr.push_back((int32_t)idx.size()); }
    };
    std::vector<int32_t> cp, ci, hp, hi;
    csr(cl, cp, ci); csr(hl, hp, hi);
    // launch plan: the widest feature tile whose bins fit the budget, balanced tiles, object blocks by the cost model of the
    // two-increment pass (one block per (state, tile) lets the CTA own its slice of the statistics)
    const size_t per_f = (size_t)R * S1 * 4, fixed = (size_t)K * 4 + (size_t)R * 4 + 64 + ((size_t)P.G_tot + (size_t)np * 2 * C + 2 + 2 * (size_t)R) * 4;
    // measured (profiles/r4e_finesweep.txt): 32-feature tiles in about 55 KB at config 4 (four CTAs per SM) 122 us against 164 us
    // of the two-increment pass; where 32-feature tiles of fine rows do not fit that budget (config 5: 170 rows x 9 bins) the
    // two-increment pass with its fewer rows stays faster, so the fine pass is only planned when they fit
    size_t budget = 60 * 1024;
#ifdef SB_HOSTSIM
    if (getenv("SB_HOSTSIM_EVAL_SMEM")) budget = (size_t)atol(getenv("SB_HOSTSIM_EVAL_SMEM"));
    if (budget < fixed + 16 * per_f) budget = fixed + 16 * per_f;         // (the emulator has no capacity limit)
#else
    if (getenv("SB_FINE_SMEM_KB")) budget = (size_t)atol(getenv("SB_FINE_SMEM_KB")) * 1024;
    if (budget < fixed + (P.ldx < 32 ? P.ldx : 32) * per_f) return 0;
#endif
    int ft_max = (int)((budget - fixed) / per_f) / 16 * 16;
    if (ft_max < 16) return 0;
    if (ft_max > 32 && !getenv("SB_FINE_WIDE")) ft_max = 32;                     // (wider tiles measured slower at config 2 and 4)
#ifdef SB_HOSTSIM
    if (getenv("SB_HOSTSIM_FINE_FT")) ft_max = atoi(getenv("SB_HOSTSIM_FINE_FT"));      // tests: several feature tiles
#endif
    int n_ftiles = (P.ldx + ft_max - 1) / ft_max;
    const int ft = ((P.ldx + n_ftiles - 1) / n_ftiles + 15) / 16 * 16;
    n_ftiles = (P.ldx + ft - 1) / ft;
    const size_t smem = fixed + (size_t)ft * per_f;
    long long per_sm = (long long)((227 * 1024) / (smem + 1024)); if (per_sm > 4) per_sm = 4; if (per_sm < 1) per_sm = 1;
    const long long slots = per_sm * c->n_sm, base = (long long)c->n_chains * n_ftiles;
    long long max_blk = N / 128; if (max_blk < 1) max_blk = 1; if (max_blk > 4096) max_blk = 4096;
    long long n_oblk = 1; double best = -1.0;
    for (long long n = 1; n <= max_blk; n++) {
        const long long waves = (base * n + slots - 1) / slots;
        const double cost = (double)waves * ((double)((N + n - 1) / n) * ft + 3.0 * (double)ft * R * S1);
        if (best < 0.0 || cost < best) { best = cost; n_oblk = n; }
    }
    if (getenv("SB_FINE_OBLK")) n_oblk = atoi(getenv("SB_FINE_OBLK"));      // tuning knob
#ifdef SB_HOSTSIM
    if (getenv("SB_HOSTSIM_EVAL_OBLK")) n_oblk = atoi(getenv("SB_HOSTSIM_EVAL_OBLK"));
#endif
    int rows = (int)((N + n_oblk - 1) / n_oblk);
    n_oblk = (N + rows - 1) / rows;
    // device copies
    std::vector<uint8_t> xs((size_t)N * P.ldx);
    TRY(d2h(c, xs.data(), P.x, xs.size()));
    for (auto &v : xs) if (v >= (uint8_t)P.S) v = (uint8_t)P.S;
    SbFineArgs &fa = fs->fa;
    TRY(up(c, &fa.xs, xs)); TRY(up(c, &fa.obj_fine, fine));
    TRY(up(c, &fa.cnt_ptr, cp)); TRY(up(c, &fa.cnt_idx, ci)); TRY(up(c, &fa.h_ptr, hp)); TRY(up(c, &fa.h_idx, hi));
    fa.n_rows = R; fa.ft = ft; fa.n_ftiles = n_ftiles; fa.n_oblk = (int)n_oblk; fa.rows_per_blk = rows; fa.smem = (unsigned)smem;
#ifndef SB_HOSTSIM
    fs->kernel = fine_kernel_ptr(S1);
    CK(cudaFuncSetAttribute(fs->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
    fs->ok = 1;
    return 0;
}

int sb_eval_setup(sb_ctx *c) {
    const SbProblem &P = c->P;
    const int n_chains = c->n_chains, K = P.K, N = P.N; (void)N;

    {
        size_t budget = 56 * 1024;       // measured best (profiles/r1k_eval_smem_sweep.txt): three CTAs per SM by registers
#ifdef SB_HOSTSIM
        if (getenv("SB_HOSTSIM_EVAL_SMEM")) budget = (size_t)atol(getenv("SB_HOSTSIM_EVAL_SMEM"));
#else
        if (getenv("SB_EVAL_SMEM_KB")) budget = (size_t)atol(getenv("SB_EVAL_SMEM_KB")) * 1024;   // tuning knob
#endif
        c->eplan_ok = sb_eval_make_plan(P, n_chains, c->n_sm, budget, &c->eplan) == 0;
        if (!c->eplan_ok) c->eplan_ok = sb_eval_make_plan(P, n_chains, c->n_sm, 220 * 1024, &c->eplan) == 0;
#ifdef SB_HOSTSIM
        if (c->eplan_ok && getenv("SB_HOSTSIM_EVAL_OBLK")) {     // tests: force several object blocks
            int nb = atoi(getenv("SB_HOSTSIM_EVAL_OBLK"));
            c->eplan.rows_per_blk = (N + nb - 1) / nb; c->eplan.n_oblk = (N + c->eplan.rows_per_blk - 1) / c->eplan.rows_per_blk;
        }
#else
        if (c->eplan_ok) {
            CK(cudaFuncSetAttribute((const void *)sb_eval_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->eplan.smem));
            if (P.geo) CK(cudaFuncSetAttribute((const void *)sb_eval_geo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c->smem_members * 12 + 64)));
        }
#endif
        c->mplan_ok = sb_mma_make_plan(P, n_chains, c->n_sm, &c->mplan) == 0;
#ifndef SB_HOSTSIM
        if (c->mplan_ok) CK(cudaFuncSetAttribute((const void *)sb_counts_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->mplan.smem));
#else
        c->mplan_ok = 0;
#endif
        TRY(mk(c, &c->d_epartial, (size_t)n_chains * SB_EVAL_MAXPARTS * 4)); TRY(mk(c, &c->d_eticket, (size_t)n_chains));
        TRY(mk(c, &c->d_egeo, (size_t)n_chains * K)); TRY(mk(c, &c->d_epost, (size_t)n_chains * 8)); TRY(mk(c, &c->d_eerr, 1));
    }
    TRY(fine_setup(c));
    return 0;
}

// ---------------------------------------------------------------------------------------
// stand-alone full evaluation of every chain (sb_eval.cuh)
// ---------------------------------------------------------------------------------------
static SbEvalArgs eval_args(sb_ctx *c, long long n_global) {
    SbEvalArgs A; memset(&A, 0, sizeof A);
    A.P = c->P; A.states = c->d_states; A.slots = c->d_slots; A.n_states = c->n_chains; A.plan = c->eplan;
    A.partial = c->d_epartial; A.ticket = c->d_eticket; A.geo_val = c->d_egeo; A.post = c->d_epost; A.err = c->d_eerr;
    A.n_global = n_global > 0 ? n_global : c->P.N; A.geo_cap = c->smem_members;
    return A;
}
#ifndef SB_HOSTSIM
struct EvTimer {
    cudaEvent_t e0 = nullptr, e1 = nullptr; cudaStream_t s; float *out;
    EvTimer(cudaStream_t st, float *o) : s(st), out(o) { if (out) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, s); } }
    void stop() { if (out) { cudaEventRecord(e1, s); cudaEventSynchronize(e1); cudaEventElapsedTime(out, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1); out = nullptr; } }
};
#endif

// statistics pass: counts, source histogram and cluster sizes of every state (zeroed first)
SB_API int sb_eval_stats(sb_ctx *c, float *ms) {
    TRY(check_chain(c, 0));
    FineState *fs = fine_of(c);
    const bool fine = fs && fs->ok && c->stats_mode == 0;
    if (!fine && !c->eplan_ok && c->stats_mode == 0) return fail(c, "evaluation histogram does not fit in shared memory (too many groups x states)");
    SbEvalArgs A = eval_args(c, 0);
    const int grid = fine ? c->n_chains * fs->fa.n_ftiles * fs->fa.n_oblk : c->n_chains * c->eplan.n_ftiles * c->eplan.n_oblk;
    // (the fine-row pass with one object block per (state, tile) stores every word of the statistics itself)
    const bool zero = !(fine && fs->fa.n_oblk == 1);
#ifdef SB_HOSTSIM
    if (ms) *ms = 0.f;
    if (zero) TRY(dev_zero(c, c->d_stats, sizeof(int32_t) * c->stats_stride * c->n_chains));
    if (fine) hostsim::launch([&]() { fine_kernel_host(c->P.S + 1, A, fs->fa); }, grid, 64, fs->fa.smem);
    else hostsim::launch([&]() { sb_eval_hist_kernel(A); }, grid, 64, c->eplan.smem);
#else
    EvTimer t(c->stream, ms);
    if (zero) TRY(dev_zero(c, c->d_stats, sizeof(int32_t) * c->stats_stride * c->n_chains));
    if (c->stats_mode == 1) {
        if (!c->mplan_ok) return fail(c, "the tensor-core statistics pass does not fit this problem (more than 256 group columns)");
        sb_counts_mma_kernel<<<c->n_chains * c->mplan.n_ftiles * c->mplan.n_oblk, SB_MMA_BLOCK, c->mplan.smem, c->stream>>>(A, c->mplan);
    } else if (fine) {
        void *args[2] = {(void *)&A, (void *)&fs->fa};
        CK(cudaLaunchKernel(fs->kernel, dim3(grid), dim3(SB_EVAL_BLOCK), args, fs->fa.smem, c->stream));
    } else {
        sb_eval_hist_kernel<<<grid, SB_EVAL_BLOCK, c->eplan.smem, c->stream>>>(A);
    }
    CK(cudaGetLastError());
    t.stop();
#endif
    c->launches++;
    return 0;
}

// which statistics pass sb_eval_stats runs and its launch geometry:
// out8 = {pass (0 two-increment histogram, 1 int8 tensor-core contraction, 2 fine-row histogram), feature-tile width,
//         feature tiles, object blocks, dynamic shared memory, fine rows, CTAs, block size}
SB_API int sb_eval_info(sb_ctx *c, int32_t *out8) {
    if (!c || !out8) return -1;
    FineState *fs = fine_of(c);
    memset(out8, 0, sizeof(int32_t) * 8);
    if (c->stats_mode == 1) {
        out8[0] = 1; out8[1] = c->mplan.ftw; out8[2] = c->mplan.n_ftiles; out8[3] = c->mplan.n_oblk; out8[4] = (int)c->mplan.smem;
        out8[6] = c->n_chains * c->mplan.n_ftiles * c->mplan.n_oblk; out8[7] = SB_MMA_BLOCK;
    } else if (fs && fs->ok) {
        out8[0] = 2; out8[1] = fs->fa.ft; out8[2] = fs->fa.n_ftiles; out8[3] = fs->fa.n_oblk; out8[4] = (int)fs->fa.smem; out8[5] = fs->fa.n_rows;
        out8[6] = c->n_chains * fs->fa.n_ftiles * fs->fa.n_oblk; out8[7] = SB_EVAL_BLOCK;
    } else {
        out8[1] = c->eplan.ft; out8[2] = c->eplan.n_ftiles; out8[3] = c->eplan.n_oblk; out8[4] = (int)c->eplan.smem;
        out8[6] = c->n_chains * c->eplan.n_ftiles * c->eplan.n_oblk; out8[7] = SB_EVAL_BLOCK;
    }
    return 0;
}

// effect draws of the listed chains at log time (sb_eval.cuh: sb_effect_sample_kernel): out [n, G_tot, F, S] float64
SB_API int sb_effect_sample(sb_ctx *c, int32_t n, const int32_t *chains, uint64_t step, double *out, float *ms) {
    TRY(check_chain(c, 0));
    if (n <= 0 || !chains || !out) return fail(c, "sb_effect_sample: nothing to do");
    for (int i = 0; i < n; i++) if (chains[i] < 0 || chains[i] >= c->n_chains) return fail(c, "chain index out of range");
    FineState *fs = fine_of(c);
    const SbProblem &P = c->P;
    const size_t per = (size_t)P.G_tot * P.F * P.S, need = per * (size_t)n;
    if (fs->eff_cap < need) {
        void *p; TRY(dev_alloc(c, &p, need * sizeof(double))); fs->d_eff = (double *)p; fs->eff_cap = need;     // (earlier buffers go with the ctx)
    }
    if (!fs->d_eff_chains) TRY(mk(c, &fs->d_eff_chains, (size_t)c->n_chains));
    if (n > c->n_chains) return fail(c, "sb_effect_sample: more entries than chains");
    // the count tensors of the committed states (a state set by sb_set_state has none yet); with the objects sharded over GPUs the
    // statistics block holds the all-reduced counts of the last launch and a local pass would replace them by this rank's share
    if (c->xworld == 1) TRY(sb_eval_stats(c, nullptr));
    TRY(h2d(c, fs->d_eff_chains, chains, sizeof(int32_t) * n));
    SbEffectArgs A; memset(&A, 0, sizeof A);
    A.P = P; A.states = c->d_states; A.slots = c->d_slots; A.chains = fs->d_eff_chains; A.n = n; A.step = step; A.out = fs->d_eff;
    const long long items = (long long)P.G_tot * P.F * n;
    int grid = (int)((items + 127) / 128); if (grid > 8 * c->n_sm) grid = 8 * c->n_sm; if (grid < 1) grid = 1;
#ifdef SB_HOSTSIM
    if (ms) *ms = 0.f;
    hostsim::launch([&]() { sb_effect_sample_kernel(A); }, grid > 4 ? 4 : grid, 64, 0);
#else
    EvTimer t(c->stream, ms);
    sb_effect_sample_kernel<<<grid, 128, 0, c->stream>>>(A);
    CK(cudaGetLastError());
    t.stop();
#endif
    c->launches++;
    return d2h(c, out, fs->d_eff, need * sizeof(double));
}

// posterior terms of every chain from the statistics currently in the ctx; `n_global` = number
// of objects of the whole problem when this ctx holds a shard of the objects (0: not sharded)
SB_API int sb_eval_terms(sb_ctx *c, int64_t n_global, sb_posterior_t *out, float *ms) {
    TRY(check_chain(c, 0));
    if (!c->eplan_ok) return fail(c, "evaluation plan unavailable (the histogram bins of one 16-feature tile do not fit in shared memory)");
    SbEvalArgs A = eval_args(c, (long long)n_global);
    if (c->P.geo && n_global > 0 && n_global != c->P.N) return fail(c, "the geo prior is not supported with sharded objects");
    TRY(dev_zero(c, c->d_eticket, sizeof(unsigned) * c->n_chains));
    TRY(dev_zero(c, c->d_eerr, sizeof(int)));
#ifdef SB_HOSTSIM
    if (ms) *ms = 0.f;
    if (c->P.geo) hostsim::launch([&]() { sb_eval_geo_kernel(A); }, c->n_chains * c->P.K, 64, (size_t)c->smem_members * 12 + 64);
    hostsim::launch([&]() { sb_eval_terms_kernel(A); }, c->n_chains * c->eplan.n_parts, 64, 1024);
#else
    EvTimer t(c->stream, ms);
    if (c->P.geo) {
        sb_eval_geo_kernel<<<c->n_chains * c->P.K, 128, (size_t)c->smem_members * 12 + 64, c->stream>>>(A);
        CK(cudaGetLastError());
        c->launches++;
    }
    sb_eval_terms_kernel<<<c->n_chains * c->eplan.n_parts, SB_EVAL_BLOCK, 1024, c->stream>>>(A);
    CK(cudaGetLastError());
    t.stop();
#endif
    c->launches++;
    if (out) {
        std::vector<double> post((size_t)c->n_chains * 8);
        TRY(d2h(c, post.data(), c->d_epost, sizeof(double) * post.size()));
        int err = 0; TRY(d2h(c, &err, c->d_eerr, sizeof(int)));
        if (err == 2) return fail(c, "source component out of range in state");
        if (err == 3) return fail(c, "cluster larger than the geo-prior member list");
        std::vector<SbSlot> slots(c->n_chains);
        TRY(d2h(c, slots.data(), c->d_slots, sizeof(SbSlot) * c->n_chains));
        for (int i = 0; i < c->n_chains; i++) {
            const double *p = &post[(size_t)slots[i].state * 8];
            out[i].ll = p[0]; out[i].size_prior = p[1]; out[i].geo_prior = p[2]; out[i].weights_prior = p[3]; out[i].source_prior = p[4];
            out[i].prior = p[1] + p[2] + p[3] + p[4];
        }
    }
    return 0;
}

// which statistics pass sb_eval_stats / sb_eval_batch run: 0 = shared-memory histogram,
// 1 = int8 tcgen05 contraction (sb_counts.cuh)
SB_API int sb_eval_set_mode(sb_ctx *c, int mode) {
    if (!c) return -1;
    if (mode != 0 && mode != 1) return fail(c, "unknown statistics mode");
#ifdef SB_HOSTSIM
    if (mode == 1) return fail(c, "the tensor-core statistics pass needs a GPU");
#endif
    if (mode == 1 && !c->mplan_ok) return fail(c, "the tensor-core statistics pass does not fit this problem (more than 256 group columns)");
    c->stats_mode = mode;
    return 0;
}

// uncached likelihood + prior of every chain: out[n_chains]; ms (may be NULL) receives the
// device time of {statistics pass, terms}
SB_API int sb_eval_batch(sb_ctx *c, sb_posterior_t *out, float *ms2) {
    TRY(sb_eval_stats(c, ms2 ? ms2 : nullptr));
    return sb_eval_terms(c, 0, out, ms2 ? ms2 + 1 : nullptr);
}

// per-object partial sums / per-chain sums shared by the observation-likelihood and the Gibbs-sweep kernels
static int ensure_obs_buffers(sb_ctx *c) {
    if (c->d_obs_part) return 0;
    const SbProblem &P = c->P;
    if (sb_obs_make_plan(P, c->n_chains, c->n_sm, &c->oplan)) return fail(c, "per-tile tables (groups x states x 16 features) do not fit in shared memory");
    TRY(mk(c, &c->d_obs_part, (size_t)c->n_chains * c->oplan.n_ftiles * P.N));
    TRY(mk(c, &c->d_obs_ll, (size_t)c->n_chains * P.N)); TRY(mk(c, &c->d_obs_sums, (size_t)c->n_chains));
    TRY(mk(c, &c->d_gibbs_changed, (size_t)c->n_chains));
#ifndef SB_HOSTSIM
    CK(cudaFuncSetAttribute((const void *)sb_obs_lik_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->oplan.smem));
    CK(cudaFuncSetAttribute((const void *)sb_gibbs_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->oplan.smem));
#endif
    return 0;
}

// Full-sweep source Gibbs of every chain (GibbsSampleSource with ObjectSelector.ALL,
// sbayes/sampling/operators.py:496-574; the initializer's use, sbayes/sampling/initializers.py:285-310).
//   mode 0  resample the source of EVERY observation of every chain from its conditional given the current
//           counts and weights (draws: Philox (chain seed, step, SITE_GIBBS, o * F + f)); commits the new
//           sources; log_q_out[chain] = sum log q[new], n_changed_out[chain] = reassigned observations
//   mode 1  log_q_out = sum log q[previous sources] under the CURRENT counts (the log_q_back of the proposal,
//           to be called after a mode-0 sweep); mode 2: the same for the committed sources
// The statistics pass runs first (the tables need the committed counts).  The cached posterior of the chains
// (sb_posterior) is stale after a mode-0 sweep until the next sb_run / sb_eval.  Outputs may be NULL; ms
// receives the device time of the sweep kernel + reduction.
SB_API int sb_gibbs_sweep(sb_ctx *c, uint64_t step, int mode, double *log_q_out, int64_t *n_changed_out, float *ms) {
    TRY(check_chain(c, 0));
    if (mode < 0 || mode > 2) return fail(c, "unknown sweep mode");
    const SbProblem &P = c->P;
    TRY(ensure_obs_buffers(c));
    TRY(sb_eval_stats(c, nullptr));
    SbGibbsArgs G; memset(&G, 0, sizeof G);
    G.P = P; G.states = c->d_states; G.slots = c->d_slots; G.n_states = c->n_chains; G.plan = c->oplan;
    G.obj_part = c->d_obs_part; G.n_changed = c->d_gibbs_changed; G.step = step; G.mode = mode;
    SbObsArgs A; memset(&A, 0, sizeof A);
    A.P = P; A.states = c->d_states; A.n_states = c->n_chains; A.plan = c->oplan;
    A.obj_part = c->d_obs_part; A.obj_ll = c->d_obs_ll; A.sums = c->d_obs_sums;
    const int grid = c->n_chains * c->oplan.n_ftiles * c->oplan.n_oblk;
    TRY(dev_zero(c, c->d_gibbs_changed, sizeof(unsigned long long) * c->n_chains));
    TRY(dev_zero(c, c->d_eticket, sizeof(unsigned) * c->n_chains));
#ifdef SB_HOSTSIM
    if (ms) *ms = 0.f;
    hostsim::launch([&]() { sb_gibbs_sweep_kernel(G); }, grid, 64, c->oplan.smem);
    hostsim::launch([&]() { sb_obs_reduce_kernel(A, c->d_epartial, c->d_eticket); }, c->n_chains * SB_OBS_RPARTS, 64, 1024);
    if (mode == 0) hostsim::launch([&]() { sb_flip_osel_kernel(c->d_states, c->n_chains, P.N); }, 2, 64, 0);
#else
    EvTimer t(c->stream, ms);
    sb_gibbs_sweep_kernel<<<grid, SB_EVAL_BLOCK, c->oplan.smem, c->stream>>>(G);
    CK(cudaGetLastError());
    sb_obs_reduce_kernel<<<c->n_chains * SB_OBS_RPARTS, SB_EVAL_BLOCK, 1024, c->stream>>>(A, c->d_epartial, c->d_eticket);
    CK(cudaGetLastError());
    t.stop();
    if (mode == 0) { sb_flip_osel_kernel<<<c->n_sm * 2, 256, 0, c->stream>>>(c->d_states, c->n_chains, P.N); CK(cudaGetLastError()); }
#endif
    c->launches += mode == 0 ? 3 : 2;
    std::vector<double> sums(c->n_chains);
    std::vector<unsigned long long> ch(c->n_chains);
    TRY(d2h(c, sums.data(), c->d_obs_sums, sizeof(double) * c->n_chains));
    TRY(d2h(c, ch.data(), c->d_gibbs_changed, sizeof(unsigned long long) * c->n_chains));
    for (int i = 0; i < c->n_chains; i++) {
        const int s = c->h_slot_state[i];
        if (log_q_out) log_q_out[i] = sums[s];
        if (n_changed_out) n_changed_out[i] = (int64_t)ch[s];
    }
    return 0;
}

// Observation likelihood of every chain (LikelihoodLogger._write_sample, sbayes/sampling/
// loggers.py:354-359; likelihood_per_component_exact, sbayes/sampling/conditionals.py:300-367).
// Runs the statistics pass first (the tables need the committed counts), then one streaming
// kernel.  lh_out [n_chains, N, F] float32, obj_ll_out [n_chains, N] and sums_out [n_chains]
// (sum of log lh per object / per chain) may each be NULL; ms receives the device time of the
// likelihood kernel + reduction.
SB_API int sb_obs_loglik(sb_ctx *c, float *lh_out, double *obj_ll_out, double *sums_out, float *ms) {
    TRY(check_chain(c, 0));
    const SbProblem &P = c->P;
    const size_t nNF = (size_t)P.N * P.F;
    TRY(ensure_obs_buffers(c));
    if (!c->d_obs_out) TRY(mk(c, &c->d_obs_out, nNF * c->n_chains));
    TRY(sb_eval_stats(c, nullptr));
    SbObsArgs A; memset(&A, 0, sizeof A);
    A.P = P; A.states = c->d_states; A.n_states = c->n_chains; A.plan = c->oplan;
    A.out = c->d_obs_out; A.obj_part = c->d_obs_part; A.obj_ll = c->d_obs_ll; A.sums = c->d_obs_sums;
    const int grid = c->n_chains * c->oplan.n_ftiles * c->oplan.n_oblk;
#ifdef SB_HOSTSIM
    if (ms) *ms = 0.f;
    hostsim::launch([&]() { sb_obs_lik_kernel(A); }, grid, 64, c->oplan.smem);
    TRY(dev_zero(c, c->d_eticket, sizeof(unsigned) * c->n_chains));
    hostsim::launch([&]() { sb_obs_reduce_kernel(A, c->d_epartial, c->d_eticket); }, c->n_chains * SB_OBS_RPARTS, 64, 1024);
#else
    EvTimer t(c->stream, ms);
    sb_obs_lik_kernel<<<grid, SB_EVAL_BLOCK, c->oplan.smem, c->stream>>>(A);
    CK(cudaGetLastError());
    TRY(dev_zero(c, c->d_eticket, sizeof(unsigned) * c->n_chains));
    sb_obs_reduce_kernel<<<c->n_chains * SB_OBS_RPARTS, SB_EVAL_BLOCK, 1024, c->stream>>>(A, c->d_epartial, c->d_eticket);
    CK(cudaGetLastError());
    t.stop();
#endif
    c->launches += 2;
    std::vector<SbSlot> slots(c->n_chains);
    TRY(d2h(c, slots.data(), c->d_slots, sizeof(SbSlot) * c->n_chains));
    for (int i = 0; i < c->n_chains; i++) {
        const int s = slots[i].state;
        if (lh_out) TRY(d2h(c, lh_out + (size_t)i * nNF, c->d_obs_out + (size_t)s * nNF, sizeof(float) * nNF));
        if (obj_ll_out) TRY(d2h(c, obj_ll_out + (size_t)i * P.N, c->d_obs_ll + (size_t)s * P.N, sizeof(double) * P.N));
        if (sums_out) TRY(d2h(c, sums_out + i, c->d_obs_sums + s, sizeof(double)));
    }
    if (!lh_out && !obj_ll_out && !sums_out) TRY(dev_sync(c));
    return 0;
}

// EM-style cluster initializer (SbayesInitializer.generate_clusters_em, sbayes/sampling/initializers.py:93-169):
// `n_steps` EM steps from the soft assignment z0 [G_tot, N] (float32; rows = clusters, then every confounder's
// groups -- the reference's order); z_out receives the final assignment.  Uniform geo prior only (the reference
// adds a geo term for cost-based priors).  ms = device time of all steps.
SB_API int sb_init_em(sb_ctx *c, int32_t n_steps, const float *z0, float *z_out, float *ms) {
    TRY(check_chain(c, 0));
    if (!z0 || !z_out || n_steps < 1) return fail(c, "bad arguments");
    if (c->P.geo) return fail(c, "sb_init_em: the geo term of the cost-based prior is not implemented (use the reference's initializer)");
    const SbProblem &P = c->P;
    const size_t G = P.G_tot, N = P.N, FS = (size_t)P.F * P.S;
    const int n_ft = (P.F + SB_EM_FT - 1) / SB_EM_FT, n_gb = (int)((G + SB_EM_GB - 1) / SB_EM_GB);
    int n_oblk = std::max(1, (8 * c->n_sm + n_ft * n_gb - 1) / (n_ft * n_gb));
    n_oblk = (int)std::min<size_t>((size_t)n_oblk, std::max<size_t>(1, N / 64));
    const int rows = (int)((N + n_oblk - 1) / n_oblk);
    n_oblk = (int)((N + rows - 1) / rows);
    const size_t keep = c->allocs.size();
    struct Guard { sb_ctx *c; size_t keep; ~Guard() { while (c->allocs.size() > keep) {
#ifdef SB_HOSTSIM
        free(c->allocs.back());
#else
        cudaFree(c->allocs.back());
#endif
        c->allocs.pop_back(); } } } guard{c, keep};
    SbEmArgs A; memset(&A, 0, sizeof A);
    A.P = P; A.n_oblk = n_oblk; A.rows_per_blk = rows;
    TRY(mk(c, &A.z, G * N)); TRY(mk(c, &A.partial, (size_t)n_oblk * G * FS)); TRY(mk(c, &A.logp, G * FS));
    TRY(h2d(c, A.z, z0, sizeof(float) * G * N));
    const size_t smem_counts = (size_t)(SB_EM_BLOCK / 32) * SB_EM_GB * P.S * 32 * 8, smem_assign = (size_t)(SB_EM_BLOCK / 32) * G * 8;
#ifndef SB_HOSTSIM
    if (smem_counts > 200 * 1024 || smem_assign > 200 * 1024) return fail(c, "sb_init_em: too many states / groups for the shared-memory bins");
    CK(cudaFuncSetAttribute((const void *)sb_em_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_counts));
    CK(cudaFuncSetAttribute((const void *)sb_em_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_assign));
    EvTimer t(c->stream, ms);
#else
    if (ms) *ms = 0.f;
#endif
    for (int i = 0; i < n_steps; i++) {
        const double temperature = pow((double)n_steps / (1.0 + (double)i), 3.0);
        A.inv_temperature = 1.0 / temperature;
#ifdef SB_HOSTSIM
        hostsim::launch([&]() { sb_em_counts_kernel(A); }, n_ft * n_gb * n_oblk, 64, smem_counts);
        hostsim::launch([&]() { sb_em_tables_kernel(A); }, (int)((G * P.F + 63) / 64), 64, 0);
        hostsim::launch([&]() { sb_em_assign_kernel(A); }, 2, 64, smem_assign);
#else
        sb_em_counts_kernel<<<n_ft * n_gb * n_oblk, SB_EM_BLOCK, smem_counts, c->stream>>>(A);
        sb_em_tables_kernel<<<(unsigned)((G * P.F + SB_EM_BLOCK - 1) / SB_EM_BLOCK), SB_EM_BLOCK, 0, c->stream>>>(A);
        sb_em_assign_kernel<<<c->n_sm * 8, SB_EM_BLOCK, smem_assign, c->stream>>>(A);
        CK(cudaGetLastError());
#endif
        c->launches += 3;
    }
#ifndef SB_HOSTSIM
    t.stop();
#endif
    return d2h(c, z_out, A.z, sizeof(float) * G * N);
}

// bulk count rebuild as stand-alone kernels (sb_counts.cuh): shared-memory histogram and the
// int8 tensor-core contraction.  Both read the committed state of `chain` and return the
// stacked counts; `ms` receives the device time of the kernel.
static int counts_common(sb_ctx *c, int chain, int32_t *counts_out, float *ms, bool mma) {
    TRY(check_chain(c, chain));
    const SbProblem &P = c->P;
    SbSlot slot; TRY(get_slot(c, chain, &slot));
    SbShared &s = c->h_states[slot.state];
    const size_t nG = (size_t)P.G_tot * P.F * P.S;
    TRY(dev_zero(c, c->d_counts_tmp, sizeof(int32_t) * nG));
    float t = 0.f;
#ifdef SB_HOSTSIM
    const int32_t *result = c->d_counts_tmp;
    if (mma) return fail(c, "the tensor-core rebuild needs a GPU");
    hostsim::launch([&]() { sb_counts_hist_kernel(P, s, c->d_counts_tmp); }, 2, 64, sb_counts_hist_smem(P));
#else
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, c->stream));
    const int32_t *result = c->d_counts_tmp;
    if (mma) {
        // single state: the tensor-core pass writes the state's own statistics block
        SbMmaPlan mp;
        if (sb_mma_make_plan(P, 1, c->n_sm, &mp)) return fail(c, "the tensor-core statistics pass does not fit this problem (more than 256 group columns)");
        SbEvalArgs A = eval_args(c, 0);
        A.states = c->d_states + slot.state; A.n_states = 1;
        CK(cudaFuncSetAttribute((const void *)sb_counts_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mp.smem));
        CK(cudaMemsetAsync(s.counts, 0, sizeof(int32_t) * c->stats_stride, c->stream));
        sb_counts_mma_kernel<<<mp.n_ftiles * mp.n_oblk, SB_MMA_BLOCK, mp.smem, c->stream>>>(A, mp);
        result = s.counts;
    } else {
        if (sb_counts_hist_launch(P, s, c->d_counts_tmp, c->n_sm, c->stream)) return fail(c, "count rebuild launch failed");
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
#endif
    c->launches++;
    if (ms) *ms = t;
    if (counts_out) TRY(d2h(c, counts_out, result, sizeof(int32_t) * nG));
    return 0;
}
SB_API int sb_counts_rebuild_hist(sb_ctx *c, int chain, int32_t *counts_out, float *ms) { return counts_common(c, chain, counts_out, ms, false); }
SB_API int sb_counts_rebuild_mma(sb_ctx *c, int chain, int32_t *counts_out, float *ms) { return counts_common(c, chain, counts_out, ms, true); }
