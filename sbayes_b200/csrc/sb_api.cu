// sb_api.cu -- C ABI of the B200 backend (include/sbayes_b200.h): problem upload, chain
// state in HBM, launches of the persistent step kernel, MC3 swap.
//
// Product build: nvcc -gencode arch=compute_100a,code=sm_100a.  With -DSB_HOSTSIM (tests
// only) the device memory is host memory and launches go to the fibre emulator.
#include "sb_ctx.h"
// the step kernel, instantiated for up to 4 and up to 8 mixture components (the loops over
// components are unrolled to that width; the narrow one is what the reference's configs need)
#define SB_UC 4
#define SB_NS uc4
#include "sb_core.cuh"
#undef SB_UC
#undef SB_NS
// (the instantiation for 5..8 components is its own translation unit, sb_uc8.cu: the two compile in parallel)
const void *sb_uc8_run_kernel_ptr();
void sb_uc8_run_kernel_host(const SbLaunch &L);        // emulator builds: the kernel body as a host function
long sb_uc8_tiled_calls(int kind);
// (and the one for exactly 3 components, sb_uc3.cu: what every BASELINE configuration runs)
const void *sb_uc3_run_kernel_ptr();
void sb_uc3_run_kernel_host(const SbLaunch &L);
long sb_uc3_tiled_calls(int kind);
#if defined(SB_PHASE_PROF) || defined(SB_NO_UC3)
static inline bool use_uc3(int) { return false; }      // (the phase counters live in the 4-component instantiation)
#else
static inline bool use_uc3(int C) { return C == 3; }
#endif
using uc4::FixedSmem; using uc4::sb_fixed_smem; using uc4::Draws; using uc4::Philox4; using uc4::philox4x32_10; using uc4::u01_double;
using uc4::sb_export_kernel; using uc4::sb_export_samples_kernel;

std::string g_create_error;

// ---------------------------------------------------------------------------------------
SB_API const char *sb_last_error(const sb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

SB_API void sb_destroy(sb_ctx *c) {
    if (!c) return;
#ifndef SB_HOSTSIM
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (int p = 0; p < SB_MAX_RANKS; p++) if (c->peer_opened[p]) cudaIpcCloseMemHandle(c->peer_arena[p]);
#endif
    dev_free_all(c);
    if (c->eval_ext && c->eval_ext_free) c->eval_ext_free(c->eval_ext);
#ifndef SB_HOSTSIM
    if (c->stream) cudaStreamDestroy(c->stream);
#endif
    delete c;
}

static int create_impl(sb_ctx *c, const sb_problem_desc *d, int device, int n_chains, int team) {
    const int N = d->n_objects, F = d->n_features, S = d->n_states, K = d->n_clusters, nc = d->n_confounders;
    const int C = 1 + nc;
    if (N <= 0 || F <= 0 || S <= 0 || S > 254 || K <= 0 || K > SB_MAXK) return fail(c, "bad problem dimensions");
    if (C > SB_MAX_COMPONENTS) return fail(c, "too many mixture components (max 8)");
    if (n_chains <= 0) return fail(c, "n_chains must be positive");
    c->device = device; c->n_chains = n_chains;
#ifndef SB_HOSTSIM
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(c, "sbayes_b200 needs an sm_100a (B200) device");
    c->n_sm = prop.multiProcessorCount;
    { int khz = 0; if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device) == cudaSuccess && khz > 0) c->clock_khz = khz; }
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
#else
    c->n_sm = 4;
#endif
    SbProblem &P = c->P;
    memset(&P, 0, sizeof(P));
    P.N = N; P.F = F; P.S = S; P.K = K; P.C = C; P.n_conf = nc;
    P.ldx = (F + 15) / 16 * 16;
    P.goff[0] = 0; P.goff[1] = K;
    for (int i = 0; i < nc; i++) { P.conf_ng[i] = d->conf_ngroups[i]; P.goff[i + 2] = P.goff[i + 1] + d->conf_ngroups[i]; }
    P.G_tot = P.goff[C];
    P.min_size = d->min_size; P.max_size = d->max_size; P.size_prior = d->size_prior;
    P.wprior_uniform = d->weights_prior_uniform;
    P.geo = d->geo_enabled && d->geo_cost; P.geo_agg = d->geo_aggregation; P.geo_pf = d->geo_prob_fn;
    P.geo_rate = d->geo_rate; P.geo_infl = d->geo_inflection;

    // packed observations, row stride padded to 16 bytes
    std::vector<uint8_t> x((size_t)N * P.ldx, SB_NA_U8);
    for (int o = 0; o < N; o++) memcpy(&x[(size_t)o * P.ldx], d->x + (size_t)o * F, F);
    for (size_t i = 0; i < x.size(); i++) if (x[i] != SB_NA_U8 && x[i] >= S) return fail(c, "state index out of range");
    TRY(up(c, &P.x, x));
    {
        // (one combo's pair table, 8 B x pairs x (S+1)^2, has to leave room for more than itself in shared memory)
        bool pairs = (S + 1) * (S + 1) <= 256 && (size_t)((F + 1) / 2 + 15) / 16 * 16 * (S + 1) * (S + 1) * 8 <= (size_t)64 * 1024;
#ifdef SB_HOSTSIM
        if (getenv("SB_NO_PAIRS")) pairs = false;
#else
        if (getenv("SB_NO_PAIRS")) pairs = false;      // tuning knob
#endif
        if (pairs) {
            const int np = (F + 1) / 2;
            P.ldp = (np + 15) / 16 * 16;
            std::vector<uint8_t> xp((size_t)N * P.ldp, (uint8_t)((S + 1) * (S + 1) - 1));
            for (int o = 0; o < N; o++)
                for (int i = 0; i < np; i++) {
                    const int a = x[(size_t)o * P.ldx + 2 * i], b = 2 * i + 1 < F ? x[(size_t)o * P.ldx + 2 * i + 1] : SB_NA_U8;
                    xp[(size_t)o * P.ldp + i] = (uint8_t)((a == SB_NA_U8 ? S : a) * (S + 1) + (b == SB_NA_U8 ? S : b));
                }
            TRY(up(c, &P.xp, xp));
        }
    }
    std::vector<int32_t> gid(d->conf_gid, d->conf_gid + (size_t)nc * N);
    for (int i = 0; i < nc; i++) for (int o = 0; o < N; o++) if (gid[(size_t)i * N + o] >= P.conf_ng[i]) return fail(c, "confounder group id out of range");
    TRY(up(c, &P.conf_gid, gid));
    // distinct confounder has-patterns
    std::map<int, int> pat_index; std::vector<int32_t> pat_bits; std::vector<uint8_t> pat_id(N);
    for (int o = 0; o < N; o++) {
        int bits = 0;
        for (int i = 0; i < nc; i++) if (gid[(size_t)i * N + o] >= 0) bits |= 1 << (i + 1);
        auto it = pat_index.find(bits);
        if (it == pat_index.end()) { it = pat_index.emplace(bits, (int)pat_bits.size()).first; pat_bits.push_back(bits); }
        pat_id[o] = (uint8_t)it->second;
    }
    P.n_pat = (int)pat_bits.size();
    TRY(up(c, &P.pat_id, pat_id)); TRY(up(c, &P.pat_bits, pat_bits));
    // distinct tuples of confounder groups ("combos", sb_types.h)
    {
        std::map<std::vector<int>, int> combo_index;
        std::vector<int32_t> combo_id(N), combo_rows, combo_pat;
        for (int o = 0; o < N; o++) {
            std::vector<int> key(nc);
            for (int i = 0; i < nc; i++) key[i] = gid[(size_t)i * N + o];
            auto it = combo_index.find(key);
            if (it == combo_index.end()) {
                it = combo_index.emplace(key, (int)combo_pat.size()).first;
                for (int i = 0; i < nc; i++) combo_rows.push_back(key[i] < 0 ? -1 : P.goff[i + 1] + key[i]);
                combo_pat.push_back(pat_id[o]);
            }
            combo_id[o] = it->second;
        }
        P.n_combos = (int)combo_pat.size();
        if (combo_rows.empty()) combo_rows.push_back(-1);
        TRY(up(c, &P.combo_id, combo_id)); TRY(up(c, &P.combo_rows, combo_rows)); TRY(up(c, &P.combo_pat, combo_pat));
        if (C <= 4 && P.G_tot < 255 && P.n_pat < 256) {
            std::vector<uint32_t> pack(N);
            for (int o = 0; o < N; o++) {
                uint32_t w = (uint32_t)pat_id[o];
                for (int u = 1; u < 4; u++) {
                    int g = 255;
                    if (u <= nc && gid[(size_t)(u - 1) * N + o] >= 0) g = P.goff[u] + gid[(size_t)(u - 1) * N + o];
                    w |= (uint32_t)g << (8 * u);
                }
                pack[o] = w;
            }
            TRY(up(c, &P.obj_pack, pack));
        }
        {   // objects sorted by combo (stable counting sort) for the tiled passes
            std::vector<int32_t> start(P.n_combos + 1, 0), perm(N);
            for (int o = 0; o < N; o++) start[combo_id[o] + 1]++;
            for (int j = 0; j < P.n_combos; j++) start[j + 1] += start[j];
            std::vector<int32_t> fill(start.begin(), start.end() - 1);
            for (int o = 0; o < N; o++) perm[fill[combo_id[o]]++] = o;
            TRY(up(c, &P.perm, perm)); TRY(up(c, &P.combo_start, start));
        }
        // per-combo pass tables: 16 B per (combo, feature, state) and CTA; kept whole (shared memory or the
        // CTA's HBM slice) while they are small, built tile by tile in shared memory otherwise
        P.use_tables = 1;
#ifdef SB_HOSTSIM
        if (getenv("SB_NO_TABLES")) P.use_tables = 0;
#endif
    }
    // concentrations and the count-independent part of the Dirichlet-categorical marginal;
    // the K cluster rows are stored once when they are identical
    const size_t nG = (size_t)P.G_tot * F * S;
    {
        const size_t FS = (size_t)F * S;
        bool same = true;
        for (int k = 1; k < K && same; k++) same = memcmp(d->alpha, d->alpha + (size_t)k * FS, FS * sizeof(double)) == 0;
        P.dedup = same ? 1 : 0; P.arow_off = same ? K - 1 : 0; P.A_rows = P.G_tot - P.arow_off;
        const double *src = d->alpha + (size_t)P.arow_off * FS;
        std::vector<double> alpha(src, src + (size_t)P.A_rows * FS), asum((size_t)P.A_rows * F), dcc((size_t)P.A_rows * F), unif(FS);
        for (size_t gf = 0; gf < (size_t)P.A_rows * F; gf++) {
            double sa = 0.0, sl = 0.0;
            for (int s = 0; s < S; s++) { double a = alpha[gf * S + s]; sa += a; if (a > 0.0) sl += lgamma(a); }
            asum[gf] = sa; dcc[gf] = lgamma(sa) - sl;
        }
        for (size_t i = 0; i < unif.size(); i++) unif[i] = d->applicable[i] ? 1.0 : 0.0;
        P.use_mask = S <= 64;
        std::vector<uint64_t> applm(F, 0);
        if (P.use_mask) for (int f = 0; f < F; f++) for (int s = 0; s < S; s++) if (d->applicable[(size_t)f * S + s]) applm[f] |= 1ull << s;
        TRY(up(c, &P.alpha, alpha)); TRY(up(c, &P.asum, asum)); TRY(up(c, &P.dcc, dcc)); TRY(up(c, &P.unif, unif));
        TRY(up(c, &P.applm, applm));
    }
    std::vector<double> aw(d->alpha_w, d->alpha_w + (size_t)F * C), wlnB(F);
    for (int f = 0; f < F; f++) {
        double sa = 0.0, sl = 0.0;
        for (int cc = 0; cc < C; cc++) { sa += aw[(size_t)f * C + cc]; sl += lgamma(aw[(size_t)f * C + cc]); }
        wlnB[f] = sl - lgamma(sa);
    }
    TRY(up(c, &P.alpha_w, aw)); TRY(up(c, &P.wlnB, wlnB));
    {
        const int n_tab = (int)std::min<long long>((long long)N + S + 64, 1LL << 22);
        std::vector<double> lf(n_tab);
        for (int i = 0; i < n_tab; i++) lf[i] = lgamma((double)i + 1.0);
        TRY(up(c, &P.lfact, lf));
        P.lfact_n = n_tab;
    }
    if (P.geo) { std::vector<double> cost(d->geo_cost, d->geo_cost + (size_t)N * N); TRY(up(c, &P.cost, cost)); }
    if (d->adj_indptr) {
        std::vector<int32_t> ip(d->adj_indptr, d->adj_indptr + N + 1);
        std::vector<int32_t> ix(d->adj_indices, d->adj_indices + ip[N]);
        TRY(up(c, &P.adj_ptr, ip)); TRY(up(c, &P.adj_idx, ix));
    }
    // members of every confounder group (CSR, sorted): the GROUPS selector of the source operator
    {
        std::vector<int32_t> ptr(1, 0), idx;
        for (int i = 0; i < nc; i++)
            for (int g = 0; g < P.conf_ng[i]; g++) {
                for (int o = 0; o < N; o++) if (gid[(size_t)i * N + o] == g) idx.push_back(o);
                ptr.push_back((int32_t)idx.size());
            }
        if (idx.empty()) idx.push_back(0);
        TRY(up(c, &P.cg_ptr, ptr)); TRY(up(c, &P.cg_idx, idx));
    }
    c->smem_members = std::max(P.G_tot, 32);
    if (P.geo) c->smem_members = std::max(c->smem_members, std::min(std::max(P.max_size + 1, 32), N));
    if (c->smem_members > 8192) return fail(c, "geo prior: max cluster size above 8192 is not supported");

    // replica placement: shared memory first (priority = enum order of sb_types.h), HBM slice otherwise
    const size_t nS = (size_t)N * P.ldx, nH = (size_t)P.n_pat * 2 * F * C;
    c->req_team = team;
    c->place_fn = [c, N, F, S, K, C, nG, nH, device](int team_for_xblk, int alias_w) -> int {
        const SbProblem &P = c->P;
        SbPlace &pl = c->place;
        memset(&pl, 0, sizeof(pl));
        const size_t nGF = (size_t)P.G_tot * F, nT = P.use_tables ? (size_t)P.n_combos * F * S : 1;
        size_t bytes[A_N];
        bytes[A_TAB0] = nT * 16; bytes[A_TAB1] = 0;      // both pass tables interleaved in A_TAB0
        // smallest useful tile: one combo's dual table or its pair table (whichever is larger) + the unit entry
        const size_t pair_combo = P.xp ? (size_t)P.ldp * (S + 1) * (S + 1) * 8 : 0;
        const size_t one_combo = std::max((size_t)F * S * 16, pair_combo) + 32;
        bytes[A_WQ] = nH * 8; bytes[A_COUNTS] = nG * 4; bytes[A_NN] = nGF * 4;
        bytes[A_RT] = nGF * 8; bytes[A_RP] = nGF * 8; bytes[A_WP] = alias_w ? 0 : nH * 8; bytes[A_W1] = alias_w ? 0 : nH * 8; bytes[A_LW] = nH * 8;
        bytes[A_H0] = nH * 4; bytes[A_H1] = nH * 4; bytes[A_DCNT] = nG * 4; bytes[A_LLGF] = nGF * 8; bytes[A_BST] = nGF * 8;
        bytes[A_BK_LLGF] = nGF * 8; bytes[A_BK_RP] = nGF * 8; bytes[A_BK_RT] = nGF * 8;
        bytes[A_ALPHA] = (size_t)P.A_rows * F * S * 8; bytes[A_ASUM] = (size_t)P.A_rows * F * 8;
        bytes[A_APPLM] = P.use_mask ? (size_t)F * 8 : 0;
        bytes[A_MCL] = K < 32767 ? (size_t)N * 2 : 0; bytes[A_MCOMBO] = P.n_combos < 65535 ? (size_t)N * 2 : 0; bytes[A_MOSEL] = (size_t)N;
        bytes[A_COMBOT] = (size_t)P.n_combos * (P.n_conf + 1) * 4;
        // per-(combo, feature) ratio rows of the pair pass (HBM slice only; used when the passes are tiled)
        bytes[A_RS] = P.xp ? (size_t)P.n_combos * F * (S + 1) * 8 : 0;
        pl.alias_w = alias_w;
        pl.xblk_rows = team_for_xblk > 0 ? (N + team_for_xblk - 1) / team_for_xblk : 0;
        bytes[A_XBLK] = (size_t)pl.xblk_rows * P.ldx;
        size_t budget = 200 * 1024;
#ifdef SB_HOSTSIM
        const char *eb2 = getenv("SB_HOSTSIM_SMEM");
        if (eb2) budget = (size_t)atol(eb2);
#else
        int max_optin = 0;
        CK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        budget = (size_t)max_optin;
        // when the pass tables cannot live in shared memory anyway (large N x F), the passes gather
        // them through L1: leave most of the SM's 256 KB to the cache (measured at config 4:
        // +10 % steps/s, profiles/r2e_c4_smem_cap.txt)
        if (getenv("SB_SMEM_CAP_KB")) budget = std::min((size_t)max_optin, (size_t)atol(getenv("SB_SMEM_CAP_KB")) * 1024);   // tuning knob
#endif
        FixedSmem fx0 = sb_fixed_smem(P.G_tot, K, F, C, c->smem_members, 0);
        if (fx0.end + 1024 > budget) return fail(c, "problem too large for the fixed shared-memory part of the step kernel");
        // arrays that only exist in shared memory (the global originals serve otherwise)
        auto smem_only = [](int a) { return a == A_XBLK || a == A_ALPHA || a == A_ASUM || a == A_APPLM || a == A_MCL || a == A_MCOMBO || a == A_MOSEL || a == A_COMBOT; };
        size_t soff = 0, goff = 0;
        const bool force_tiled =
#ifdef SB_HOSTSIM
            getenv("SB_HOSTSIM_TILED") != nullptr;
#else
            getenv("SB_FORCE_TILED") != nullptr;
#endif
        size_t avail = budget > fx0.end + 256 ? (budget - fx0.end - 256) & ~(size_t)15 : 0;
        // does the whole pass table find room behind the arrays that precede it?
        bool tiled = false;
        if (P.use_tables) {
            size_t pre = 0;
            for (int a = 0; a < A_TAB0; a++) { size_t b = (bytes[a] + 15) & ~(size_t)15; if (pre + b <= avail) pre += b; }
            tiled = force_tiled || pre + ((bytes[A_TAB0] + 15) & ~(size_t)15) > avail;
        }
#ifndef SB_HOSTSIM
        if (tiled && !getenv("SB_SMEM_CAP_KB")) {
            // tiled placement: the replicas live in the HBM slice and are reached through L1, so a part of the
            // SM's 256 KB is left to the cache (measured at config 4, profiles/r3g_c4_tile_sweep.txt: 168 KB of
            // shared memory 205 us per step, all 227 KB 225 us) -- unless one combo's table needs more
            const size_t want = std::max((size_t)168 * 1024, one_combo + fx0.end + 256 + 4096);
            budget = std::min(budget, want);
            avail = budget > fx0.end + 256 ? (budget - fx0.end - 256) & ~(size_t)15 : 0;
        }
#endif
        size_t tile_bytes = 0, tile_full = 0;
        if (tiled) {
            // the table is built tile by tile in shared memory (sb_core.cuh: tiled_pass).  The tile buffer is
            // reserved first (at least one combo's dual table, 96 KB when the table is that large), the other
            // arrays follow in priority order.
            const size_t full = std::max((bytes[A_TAB0] + 15) & ~(size_t)15, (pair_combo * P.n_combos + 47) & ~(size_t)15);
            if (avail < one_combo) return fail(c, "problem too large: one combo's pass table (16 B x features x states) does not fit in shared memory");
            size_t tile_target = (size_t)96 * 1024;
#ifndef SB_HOSTSIM
            if (getenv("SB_TILE_KB")) tile_target = (size_t)atol(getenv("SB_TILE_KB")) * 1024;     // tuning knob
#endif
            tile_bytes = std::min(avail, std::max(one_combo, std::min(full, tile_target)));
            if (force_tiled) tile_bytes = std::min(tile_bytes, std::max(one_combo, ((full / 3) + 15) & ~(size_t)15));   // tests: several tiles
            tile_bytes = (tile_bytes + 15) & ~(size_t)15;
            tile_full = full;
        }
        const size_t room_others = avail - tile_bytes;
        for (int a = 0; a < A_N; a++) {
            size_t b = (bytes[a] + 15) & ~(size_t)15;
            pl.glob_off[a] = goff;
            if (a == A_TAB0 && tiled) continue;            // placed below; no copy in the HBM slice
            if (a == A_RS) { if (tiled) goff += b; continue; }      // HBM slice only
            if (!smem_only(a)) goff += b;
            if (b > 0 && soff + b <= room_others) { pl.in_smem[a] = 1; pl.smem_off[a] = (unsigned)soff; soff += b; }
        }
        if (tiled) {
            (void)tile_full;     // (the tile keeps its target size: what is left over serves the cache, see above)
            pl.tiled = 1; pl.tile_doubles = (unsigned)(tile_bytes / 8);
            pl.in_smem[A_TAB0] = 1; pl.smem_off[A_TAB0] = (unsigned)soff; soff += tile_bytes;
        }
        if (!pl.in_smem[A_XBLK]) pl.xblk_rows = 0;
        pl.sel_off = goff; goff += ((size_t)N + 15) & ~(size_t)15;
        pl.rep_bytes = (goff + 255) & ~(size_t)255;
        pl.fixed_off = (unsigned)soff;
        pl.smem_bytes = sb_fixed_smem(P.G_tot, K, F, C, c->smem_members, pl.fixed_off).end;
        c->smem = pl.smem_bytes;
        return 0;
    };
    auto place = [c](int t) { return c->place_fn(t, 0); };
    TRY(place(0));

    // launch geometry: teams of `team` CTAs; a team of <= 16 CTAs is a thread-block cluster
#ifdef SB_HOSTSIM
    const char *eb = getenv("SB_HOSTSIM_BLOCK");
    c->block = eb ? atoi(eb) : 64;
    int capacity = 64;
#else
    c->block = SB_BLOCK;
    c->kernel = use_uc3(C) ? sb_uc3_run_kernel_ptr() : (C <= 4 ? (const void *)uc4::sb_run_kernel : sb_uc8_run_kernel_ptr());
    CK(cudaFuncSetAttribute(c->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem));
    CK(cudaFuncSetAttribute(c->kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, c->kernel, c->block, c->smem));
    if (per_sm < 1) return fail(c, "step kernel does not fit on an SM");
    int capacity = per_sm * c->n_sm;
#endif
    if (team <= 0) {   // auto: about 8K observations per CTA and pass, at most a 16-CTA cluster per chain
        long long want = ((long long)N * F + 8191) / 8192;
        long long cap_per_chain = std::max(1, capacity / std::max(1, std::min(n_chains, capacity)));
        team = (int)std::max(1LL, std::min(want, cap_per_chain));
        if (want > 8 && want <= 32 && cap_per_chain >= 16) team = 16;   // a full 16-CTA cluster measured fastest (profiles/r1e_team_sweep)
    }
    if (team > capacity) team = capacity;
#ifndef SB_HOSTSIM
    // the capacity above came from the placement without the per-CTA block of observation rows
    // (A_XBLK depends on the team size): re-place with the chosen team and re-query the occupancy
    // until it is consistent, so the cooperative launch never asks for more CTAs than can be resident
    for (int iter = 0; iter < 4; iter++) {
        TRY(c->place_fn(team, 0));
        CK(cudaFuncSetAttribute(c->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, c->kernel, c->block, c->smem));
        if (per_sm < 1) return fail(c, "step kernel does not fit on an SM");
        capacity = per_sm * c->n_sm;
        if (team <= capacity) break;
        team = capacity;
    }
#endif
    c->team = team;
    c->use_cluster = 0;
#ifndef SB_HOSTSIM
    if (team > 1 && team <= 16 && !getenv("SB_NO_CLUSTER")) {
        // largest number of co-resident clusters of this shape
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(team * std::max(1, capacity / team)); cfg.blockDim = dim3(c->block); cfg.dynamicSmemBytes = c->smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = team; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, c->kernel, &cfg) == cudaSuccess && n_clusters >= 1) {
            c->use_cluster = 1;
            capacity = n_clusters * team;
        } else cudaGetLastError();
    }
#endif
    c->n_teams = std::max(1, std::min(n_chains, capacity / team));
    c->grid = c->n_teams * team;
    // worst-case private slice (no aliasing); the placement itself is recomputed at every launch
    TRY(c->place_fn(team, 0));
    c->rep_alloc = c->place.rep_bytes;
    c->h_cfg.assign(n_chains, sb_chain_cfg());
    for (int i = 0; i < n_chains; i++) { c->h_cfg[i].temperature = 1.0; c->h_cfg[i].prior_temperature = 1.0; }

    // per-chain shared state and per-CTA replica slices
    c->h_states.resize(n_chains);
    c->stats_stride = (nG + nH + (size_t)K + 3) & ~(size_t)3;
    c->h_slot_state.resize(n_chains);
    for (int i = 0; i < n_chains; i++) c->h_slot_state[i] = i;
    auto carve = [&]() -> int {
        c->arena_used = 0;
        TRY(amk(c, &c->d_tred, (size_t)c->n_teams * 2 * team * SB_MAX_RANKS * SB_REDW));
        TRY(amk(c, &c->d_xflag, (size_t)c->n_teams * SB_MAX_RANKS));
        TRY(amk(c, &c->d_stats, c->stats_stride * (size_t)n_chains));
        TRY(amk(c, &c->d_lstat, c->stats_stride * (size_t)n_chains));
        TRY(amk(c, &c->d_gstat, c->stats_stride * (size_t)n_chains));
        TRY(amk(c, &c->d_post_all, (size_t)n_chains * 8));
        TRY(amk(c, &c->d_glh_all, (size_t)n_chains * P.G_tot)); TRY(amk(c, &c->d_geo_all, (size_t)n_chains * K));
        for (int i = 0; i < n_chains; i++) {
            SbShared &s = c->h_states[i];
            TRY(amk(c, &s.cluster_of, N)); TRY(amk(c, &s.W, (size_t)F * C));
            for (int b = 0; b < 2; b++) { TRY(amk(c, &s.src[b], nS)); TRY(amk(c, &s.val[b], N)); }
            TRY(amk(c, &s.osel, N));
            TRY(amk(c, &s.list_a, N)); TRY(amk(c, &s.list_b, N)); TRY(amk(c, &s.flag_a, N)); TRY(amk(c, &s.flag_b, N));
            if (!c->arena_dry) {
                s.counts = c->d_stats + (size_t)i * c->stats_stride; s.sizes = s.counts + nG; s.H = s.sizes + K;   // counts | sizes | H
                s.post = c->d_post_all + (size_t)i * 8;
                s.glh = c->d_glh_all + (size_t)i * P.G_tot; s.geo_out = c->d_geo_all + (size_t)i * K;
            }
        }
        return 0;
    };
    c->arena_dry = true;
    TRY(carve());
    c->arena_size = c->arena_used;
    { void *p; TRY(dev_alloc(c, &p, c->arena_size)); c->arena = (char *)p; }
    c->arena_dry = false;
    TRY(carve());
    c->peer_arena[0] = c->arena;
    TRY(mk(c, &c->d_lrelease, (size_t)c->n_teams + 1)); TRY(mk(c, &c->d_xabort, 1));
    TRY(mk(c, &c->d_rep, c->rep_alloc * (size_t)c->grid));
    TRY(mk(c, &c->d_states, n_chains)); TRY(h2d(c, c->d_states, c->h_states.data(), sizeof(SbShared) * n_chains));
    std::vector<SbSlot> slots(n_chains);
    memset(slots.data(), 0, sizeof(SbSlot) * n_chains);
    for (int i = 0; i < n_chains; i++) {
        slots[i].state = i; slots[i].cfg.temperature = 1.0; slots[i].cfg.prior_temperature = 1.0; slots[i].cfg.p_grow = 0.5;
        slots[i].cfg.seed = (uint64_t)i;
    }
    TRY(mk(c, &c->d_slots, n_chains)); TRY(h2d(c, c->d_slots, slots.data(), sizeof(SbSlot) * n_chains));
    TRY(mk(c, &c->d_bars, c->n_teams + 1));
    TRY(mk(c, &c->d_src_tmp, nS)); TRY(mk(c, &c->d_counts_tmp, nG));
    TRY(sb_eval_setup(c));        // plans and scratch buffers of the stand-alone kernels (sb_api_eval.cu)
    return dev_sync(c);
}

SB_API int sb_create(const sb_problem_desc *desc, int device, int n_chains, int team_ctas, sb_ctx **out) {
    if (!desc || !out) { g_create_error = "null argument"; return -1; }
    sb_ctx *c = new sb_ctx();
    int r = create_impl(c, desc, device, n_chains, team_ctas);
    if (r) { g_create_error = c->err; sb_destroy(c); *out = nullptr; return r; }
    *out = c;
    return 0;
}

SB_API int sb_chain_config(sb_ctx *c, int chain, const sb_chain_cfg *cfg) {
    TRY(check_chain(c, chain));
    if (!(cfg->temperature > 0) || !(cfg->prior_temperature > 0)) return fail(c, "temperatures must be positive");
    c->h_cfg[chain] = *cfg;
    return h2d(c, &c->d_slots[chain].cfg, cfg, sizeof(sb_chain_cfg));
}

SB_API int sb_set_state(sb_ctx *c, int chain, const int32_t *cluster_of, const float *W, const uint8_t *src, int64_t i_step) {
    TRY(check_chain(c, chain));
    const SbProblem &P = c->P;
    SbSlot slot; TRY(get_slot(c, chain, &slot));
    SbShared &s = c->h_states[slot.state];
    if (!cluster_of && !W && !src) {
        // a rank > 0 of a sharded chain: the state is rank 0's (sb_shard_attach); only mark the chain active here
        if (c->xworld <= 1 || c->xrank == 0) return fail(c, "sb_set_state without arrays is for ranks > 0 of a sharded chain");
        int one = 1;
        return h2d(c, &c->d_slots[chain].active, &one, sizeof(int));
    }
    if (!cluster_of || !W || !src) return fail(c, "null state array");
    std::vector<int> size(P.K, 0);
    for (int o = 0; o < P.N; o++) {
        if (cluster_of[o] >= P.K) return fail(c, "cluster index out of range");
        if (cluster_of[o] >= 0) size[cluster_of[o]]++;
    }
    std::vector<uint8_t> sp((size_t)P.N * P.ldx, SB_NA_U8);
    for (int o = 0; o < P.N; o++) memcpy(&sp[(size_t)o * P.ldx], src + (size_t)o * P.F, P.F);
    for (int o = 0; o < P.N; o++)
        for (int f = 0; f < P.F; f++) {
            uint8_t v = sp[(size_t)o * P.ldx + f];
            if (v != SB_NA_U8 && v >= P.C) return fail(c, "source component out of range");
        }
    TRY(h2d(c, s.cluster_of, cluster_of, sizeof(int32_t) * P.N));
    TRY(h2d(c, s.W, W, sizeof(float) * P.F * P.C));
    TRY(h2d(c, s.src[0], sp.data(), sp.size()));
    TRY(dev_zero(c, s.osel, P.N));
    double post[8] = {0, 0, 0, 0, 0, (double)i_step, 0, 0};
    TRY(h2d(c, s.post, post, sizeof(post)));
    int one = 1;
    TRY(h2d(c, &c->d_slots[chain].active, &one, sizeof(int)));
    return 0;
}

static int check_errors(sb_ctx *c);
static int check_errors_fwd(sb_ctx *c) { return check_errors(c); }
template <class T> static T *peer_ptr(sb_ctx *c, int p, T *own) { return (T *)(c->peer_arena[p] + ((char *)own - c->arena)); }

static int prepare_launch(sb_ctx *c, long long n_steps, const sb_tape *tape_dev, sb_step_record *rec_dev, SbLaunch *Lp) {
    int alias_w = 1;
    for (const sb_chain_cfg &g : c->h_cfg) if (g.prior_temperature != 1.0) alias_w = 0;
    TRY(c->place_fn(c->team, alias_w));
    if (c->place.rep_bytes > c->rep_alloc) return fail(c, "internal: replica slice larger than allocated");
#ifndef SB_HOSTSIM
    CK(cudaFuncSetAttribute(c->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem));
#endif
    SbLaunch &L = *Lp;
    memset(&L, 0, sizeof(L));
    L.P = c->P; L.states = c->d_states; L.slots = c->d_slots; L.n_slots = c->n_chains; L.team = c->team;
    L.n_steps = n_steps; L.bars = c->d_bars; L.tred = c->d_tred; L.records = rec_dev; L.smem_members = c->smem_members;
    L.rep_base = c->d_rep; L.place = c->place; L.use_cluster = c->use_cluster; L.rep_stride = c->rep_alloc;
    L.cta_base = 0; L.lgrid = c->grid;
    L.xworld = c->xworld; L.xrank = c->xrank;
    if (c->xworld > 1) {
        // collective launch: every rank launches the same number of steps; the flag values of this launch lie
        // above those of all earlier ones
        L.use_cluster = 0;
        L.xepoch_base = (++c->xlaunches) << 32;
        for (int p = 0; p < c->xworld; p++) {
            L.arena_peer[p] = c->peer_arena[p];
            L.tred_peer[p] = peer_ptr(c, p, c->d_tred); L.xflag_peer[p] = peer_ptr(c, p, c->d_xflag);
            L.lstat_peer[p] = peer_ptr(c, p, c->d_lstat); L.gstat_peer[p] = peer_ptr(c, p, c->d_gstat);
        }
        L.lrelease = c->d_lrelease; L.stats_stride = c->stats_stride; L.lcl = c->d_lcl; L.losel = c->d_losel; L.xabort = c->d_xabort;
        TRY(dev_zero(c, c->d_lrelease, sizeof(unsigned long long) * (c->n_teams + 1)));
        TRY(dev_zero(c, c->d_xabort, sizeof(int)));
    }
    if (tape_dev) L.tape = *tape_dev;
    TRY(dev_zero(c, c->d_bars, sizeof(unsigned long long) * (c->n_teams + 1)));
    c->launches++;
    return 0;
}

static int launch(sb_ctx *c, long long n_steps, const sb_tape *tape_dev, sb_step_record *rec_dev) {
    SbLaunch L;
    TRY(prepare_launch(c, n_steps, tape_dev, rec_dev, &L));
#ifdef SB_HOSTSIM
    if (c->xworld > 1) return fail(c, "emulator: chains sharded over ranks are launched together (sb_run_group)");
    if (use_uc3(c->P.C)) hostsim::launch([&]() { sb_uc3_run_kernel_host(L); }, c->grid, c->block, c->smem);
    else if (c->P.C <= 4) hostsim::launch([&]() { uc4::sb_run_kernel(L); }, c->grid, c->block, c->smem);
    else hostsim::launch([&]() { sb_uc8_run_kernel_host(L); }, c->grid, c->block, c->smem);
#else
    if (c->use_cluster && c->xworld == 1) {
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(c->grid); cfg.blockDim = dim3(c->block); cfg.dynamicSmemBytes = c->smem; cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c->team; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        void *args[] = {&L};
        CK(cudaLaunchKernelExC(&cfg, c->kernel, args));
    } else {
        void *args[] = {&L};
        CK(cudaLaunchCooperativeKernel(c->kernel, dim3(c->grid), dim3(c->block), args, c->smem, c->stream));
    }
#endif
    return 0;
}

// ---------------------------------------------------------------------------------------
// objects of a chain sharded over the GPUs of one box (BASELINE config 5)
// ---------------------------------------------------------------------------------------
struct ShardHandle {            // 128 bytes exchanged between the ranks (torch.distributed all_gather)
    unsigned char ipc[64];      // cudaIpcMemHandle_t of the arena
    uint64_t raw;               // the arena's address (same-process peers, emulator)
    uint64_t size;
    int32_t device, pid;
    unsigned char pad[128 - 64 - 8 - 8 - 4 - 4];
};
static_assert(sizeof(ShardHandle) == 128, "handle is 128 bytes");

SB_API int sb_shard_handle(sb_ctx *c, void *out128) {
    TRY(check_chain(c, 0));
    ShardHandle h; memset(&h, 0, sizeof h);
    h.raw = (uint64_t)(uintptr_t)c->arena; h.size = c->arena_size; h.device = c->device;
#ifndef SB_HOSTSIM
    h.pid = (int32_t)getpid();
    cudaIpcMemHandle_t ih;
    CK(cudaIpcGetMemHandle(&ih, c->arena));
    static_assert(sizeof(ih) == 64, "ipc handle size");
    memcpy(h.ipc, &ih, 64);
#endif
    memcpy(out128, &h, sizeof h);
    return 0;
}

// Makes this ctx rank `rank` of `world` ranks that advance the SAME chains together, every rank processing
// its block of the objects.  Every rank was created with the same (whole) problem, chain count and team size.
// handles = the 128-byte handles of all ranks in rank order (sb_shard_handle).  The per-chain state of rank 0
// becomes THE state of the chains: the other ranks reach it through the mapped arena of rank 0.
SB_API int sb_shard_attach(sb_ctx *c, int rank, int world, const void *handles) {
    TRY(check_chain(c, 0));
    if (world < 1 || world > SB_MAX_RANKS || rank < 0 || rank >= world) return fail(c, "bad rank / world");
    if (c->xworld > 1) return fail(c, "ctx is already attached");
    if (world == 1) return 0;
    if (c->P.geo) return fail(c, "the cost-based geo prior is not supported with sharded objects");
    const ShardHandle *H = (const ShardHandle *)handles;
    for (int p = 0; p < world; p++) {
        if (H[p].size != c->arena_size) return fail(c, "the ranks were not created with the same problem / chains / team");
        if (p == rank) { c->peer_arena[p] = c->arena; continue; }
#ifdef SB_HOSTSIM
        c->peer_arena[p] = (char *)(uintptr_t)H[p].raw;
#else
        if (H[p].pid == (int32_t)getpid()) {      // ranks of one process: plain peer access
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, c->device, H[p].device));
            if (!can) return fail(c, "no peer access between the GPUs");
            cudaError_t e = cudaDeviceEnablePeerAccess(H[p].device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { c->err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e); return -2; }
            cudaGetLastError();
            c->peer_arena[p] = (char *)(uintptr_t)H[p].raw;
        } else {
            cudaIpcMemHandle_t ih; memcpy(&ih, H[p].ipc, 64);
            void *ptr = nullptr;
            CK(cudaIpcOpenMemHandle(&ptr, ih, cudaIpcMemLazyEnablePeerAccess));
            c->peer_arena[p] = (char *)ptr; c->peer_opened[p] = true;
        }
#endif
    }
    c->xrank = rank; c->xworld = world;
    const size_t N = c->P.N;
    TRY(mk(c, &c->d_lcl, (size_t)c->n_chains * N)); TRY(mk(c, &c->d_losel, (size_t)c->n_chains * N));
    if (rank != 0) {
        // the chains' state is rank 0's: same offsets inside rank 0's arena
        for (SbShared &st : c->h_states) {
            st.cluster_of = peer_ptr(c, 0, st.cluster_of); st.W = peer_ptr(c, 0, st.W); st.osel = peer_ptr(c, 0, st.osel);
            for (int b = 0; b < 2; b++) { st.src[b] = peer_ptr(c, 0, st.src[b]); st.val[b] = peer_ptr(c, 0, st.val[b]); }
            st.counts = peer_ptr(c, 0, st.counts); st.H = peer_ptr(c, 0, st.H); st.sizes = peer_ptr(c, 0, st.sizes);
            st.list_a = peer_ptr(c, 0, st.list_a); st.list_b = peer_ptr(c, 0, st.list_b);
            st.flag_a = peer_ptr(c, 0, st.flag_a); st.flag_b = peer_ptr(c, 0, st.flag_b);
            st.post = peer_ptr(c, 0, st.post); st.glh = peer_ptr(c, 0, st.glh); st.geo_out = peer_ptr(c, 0, st.geo_out);
        }
        TRY(h2d(c, c->d_states, c->h_states.data(), sizeof(SbShared) * c->n_chains));
    }
    c->use_cluster = 0;
    return dev_sync(c);
}

// Launches `n` attached ctxs of ONE process together (all ranks of a sharded chain) and waits for all of them:
// the single-process way to drive a box, and the only way in the emulator (one fibre pool runs all ranks).
SB_API int sb_run_group(sb_ctx **ctxs, int n, int64_t n_steps, sb_step_record *records0) {
    if (!ctxs || n < 1 || n > SB_MAX_RANKS) return -1;
    std::vector<SbLaunch> L(n);
    sb_step_record *rd = nullptr;
    sb_ctx *c0 = ctxs[0];
    const size_t n_rec = (size_t)c0->n_chains * (size_t)n_steps;
    size_t keep0 = c0->allocs.size();
    if (records0 && n_rec) TRY(mk(c0, &rd, n_rec));
    int r = 0;
    for (int i = 0; i < n && !r; i++) {
#ifndef SB_HOSTSIM
        cudaSetDevice(ctxs[i]->device);
#endif
        r = prepare_launch(ctxs[i], n_steps, nullptr, ctxs[i]->xrank == 0 ? rd : nullptr, &L[i]);
    }
#ifdef SB_HOSTSIM
    if (!r) {
        int total = 0;
        for (int i = 0; i < n; i++) { L[i].cta_base = total; L[i].lgrid = ctxs[i]->grid; total += ctxs[i]->grid; }
        size_t smem = 0;
        for (int i = 0; i < n; i++) smem = std::max(smem, ctxs[i]->smem);
        hostsim::launch([&]() {
            const int cta = hostsim::cta();
            int i = 0;
            while (i + 1 < n && cta >= L[i + 1].cta_base) i++;
            if (use_uc3(ctxs[i]->P.C)) sb_uc3_run_kernel_host(L[i]); else if (ctxs[i]->P.C <= 4) uc4::sb_run_kernel(L[i]); else sb_uc8_run_kernel_host(L[i]);
        }, total, c0->block, smem);
    }
#else
    for (int i = 0; i < n && !r; i++) {
        sb_ctx *c = ctxs[i];
        cudaSetDevice(c->device);
        void *args[] = {&L[i]};
        cudaError_t e = cudaLaunchCooperativeKernel(c->kernel, dim3(c->grid), dim3(c->block), args, c->smem, c->stream);
        if (e != cudaSuccess) { c->err = std::string("cudaLaunchCooperativeKernel: ") + cudaGetErrorString(e); r = -2; }
    }
    for (int i = 0; i < n; i++) { cudaSetDevice(ctxs[i]->device); cudaError_t e = cudaStreamSynchronize(ctxs[i]->stream); if (e != cudaSuccess && !r) { ctxs[i]->err = cudaGetErrorString(e); r = -2; } }
    cudaSetDevice(c0->device);
#endif
    if (!r && rd) r = d2h(c0, records0, rd, sizeof(sb_step_record) * n_rec);
    while (c0->allocs.size() > keep0) {
#ifdef SB_HOSTSIM
        free(c0->allocs.back());
#else
        cudaFree(c0->allocs.back());
#endif
        c0->allocs.pop_back();
    }
    if (r) return r;
    for (int i = 0; i < n; i++) TRY(check_errors_fwd(ctxs[i]));
    return 0;
}

static int check_errors(sb_ctx *c) {
    std::vector<SbSlot> slots(c->n_chains);
    TRY(d2h(c, slots.data(), c->d_slots, sizeof(SbSlot) * c->n_chains));
    for (int i = 0; i < c->n_chains; i++)
        if (slots[i].err) {
            static const char *msg[] = {"", "non-finite posterior accepted (sbayes/sampling/mcmc_chain.py:231-236)",
                                        "source component out of range in state", "cluster larger than the geo-prior member list",
                                        "wide operator could not change the cluster",
                                        "a rank of the sharded chain did not arrive at a team barrier (timeout)"};
            char buf[256];
            snprintf(buf, sizeof buf, "chain %d: %s", i, msg[std::min(slots[i].err, 5)]);
            c->err = buf;
            int zero = 0; h2d(c, &c->d_slots[i].err, &zero, sizeof(int));
            return -3;
        }
    return 0;
}

SB_API int sb_sync(sb_ctx *c) {
    if (!c) return -1;
    TRY(dev_sync(c));
    return check_errors(c);
}

SB_API int sb_run_async(sb_ctx *c, int64_t n_steps) {
    TRY(check_chain(c, 0));
    return launch(c, n_steps, nullptr, nullptr);
}

SB_API int sb_run(sb_ctx *c, int64_t n_steps, const sb_tape *tape, sb_step_record *records) {
    TRY(check_chain(c, 0));
    sb_tape td; memset(&td, 0, sizeof td);
    // tape / record buffers live for this call only: released on every path out of it
    struct TmpGuard {
        sb_ctx *c; size_t n_keep;
        ~TmpGuard() {
            while (c->allocs.size() > n_keep) {
#ifdef SB_HOSTSIM
                free(c->allocs.back());
#else
                cudaFree(c->allocs.back());
#endif
                c->allocs.pop_back();
            }
        }
    } guard{c, c->allocs.size()};
    if (tape && tape->vals) {
        double *v; int32_t *o, *l, *w;
        TRY(mk(c, &v, (size_t)std::max<int64_t>(tape->n_vals, 1))); TRY(h2d(c, v, tape->vals, sizeof(double) * tape->n_vals));
        TRY(mk(c, &o, (size_t)n_steps * SB_N_SITES)); TRY(h2d(c, o, tape->off, sizeof(int32_t) * n_steps * SB_N_SITES));
        TRY(mk(c, &l, (size_t)n_steps * SB_N_SITES)); TRY(h2d(c, l, tape->len, sizeof(int32_t) * n_steps * SB_N_SITES));
        td.vals = v; td.off = o; td.len = l; td.n_vals = tape->n_vals;
        if (tape->wide_rows) { TRY(mk(c, &w, (size_t)n_steps)); TRY(h2d(c, w, tape->wide_rows, sizeof(int32_t) * n_steps)); td.wide_rows = w; }
    }
    sb_step_record *rd = nullptr;
    size_t n_rec = (size_t)c->n_chains * (size_t)n_steps;
    if (records && n_rec) TRY(mk(c, &rd, n_rec));
    int r = launch(c, n_steps, (tape && tape->vals) ? &td : nullptr, rd);
    if (!r) r = dev_sync(c);
    if (!r && rd) r = d2h(c, records, rd, sizeof(sb_step_record) * n_rec);
    if (r) return r;
    return check_errors(c);
}

SB_API int sb_posterior(sb_ctx *c, int chain, sb_posterior_t *out) {
    TRY(check_chain(c, chain));
    SbSlot slot; TRY(get_slot(c, chain, &slot));
    double post[8];
    TRY(d2h(c, post, c->h_states[slot.state].post, sizeof(post)));
    out->ll = post[0]; out->size_prior = post[1]; out->geo_prior = post[2]; out->weights_prior = post[3]; out->source_prior = post[4];
    out->prior = post[1] + post[2] + post[3] + post[4];
    return 0;
}

SB_API int sb_eval(sb_ctx *c, int chain, sb_posterior_t *out) {
    TRY(check_chain(c, chain));
    TRY(launch(c, 0, nullptr, nullptr));
    TRY(dev_sync(c));
    TRY(check_errors(c));
    return sb_posterior(c, chain, out);
}

SB_API int sb_get_state(sb_ctx *c, int chain, int32_t *cluster_of, float *W, uint8_t *src, int32_t *counts) {
    TRY(check_chain(c, chain));
    const SbProblem &P = c->P;
    SbSlot slot; TRY(get_slot(c, chain, &slot));
    SbShared &s = c->h_states[slot.state];
    if (cluster_of) TRY(d2h(c, cluster_of, s.cluster_of, sizeof(int32_t) * P.N));
    if (W) TRY(d2h(c, W, s.W, sizeof(float) * P.F * P.C));
    if (src) {
#ifdef SB_HOSTSIM
        hostsim::launch([&]() { sb_export_kernel(P, s, c->d_src_tmp); }, 2, 64, 0);
#else
        sb_export_kernel<<<c->n_sm * 2, 256, 0, c->stream>>>(P, s, c->d_src_tmp);
        CK(cudaGetLastError());
#endif
        c->launches++;
        {
            std::vector<uint8_t> sp((size_t)P.N * P.ldx);
            TRY(d2h(c, sp.data(), c->d_src_tmp, sp.size()));
            for (int o = 0; o < P.N; o++) memcpy(src + (size_t)o * P.F, &sp[(size_t)o * P.ldx], P.F);
        }
    }
    if (counts) TRY(d2h(c, counts, s.counts, sizeof(int32_t) * P.G_tot * P.F * P.S));
    return 0;
}

// Batched sample export: the committed states of `n` chains (chains[] = chain indices) straight into the caller's
// host buffers -- one gather kernel into staging buffers in the host layout, one asynchronous copy per array,
// ONE synchronisation.  Besides what sb_get_state returns it delivers what the reference's loggers read from the
// cache nodes of a Sample (sbayes/sampling/loggers.py:157-262, sbayes/sampling/state.py:396-449): posterior
// parts, per-group log-likelihoods, per-cluster geo prior, per-object source prior.  Pinned host buffers make
// the copies truly asynchronous; any pointer of `out` may be NULL.
SB_API int sb_export_samples(sb_ctx *c, int32_t n, const int32_t *chains, const sb_sample_out *out) {
    TRY(check_chain(c, 0));
    if (n <= 0 || n > c->n_chains || !chains || !out) return fail(c, "bad export request");
    const SbProblem &P = c->P;
    const size_t N = P.N, F = P.F, C = P.C, nG = (size_t)P.G_tot * F * P.S, nc = c->n_chains;
    if (!c->exp_ok) {
        memset(&c->exp_, 0, sizeof(SbExport));
        TRY(mk(c, &c->exp_.cluster_of, nc * N)); TRY(mk(c, &c->exp_.W, nc * F * C)); TRY(mk(c, &c->exp_.src, nc * N * F));
        TRY(mk(c, &c->exp_.counts, nc * nG)); TRY(mk(c, &c->exp_.post, nc * 8)); TRY(mk(c, &c->exp_.glh, nc * P.G_tot));
        TRY(mk(c, &c->exp_.geo, nc * P.K)); TRY(mk(c, &c->exp_.sp_obj, nc * N)); TRY(mk(c, &c->d_exp_idx, nc));
        c->exp_ok = 1;
    }
    std::vector<int32_t> idx(n);
    for (int i = 0; i < n; i++) {
        if (chains[i] < 0 || chains[i] >= c->n_chains) return fail(c, "chain index out of range");
        idx[i] = c->h_slot_state[chains[i]];
    }
    TRY(h2d(c, c->d_exp_idx, idx.data(), sizeof(int32_t) * n));
    SbExport E = c->exp_;
    E.state_idx = c->d_exp_idx; E.n = n;
    E.want = (out->cluster_of ? 1u : 0u) | (out->W ? 2u : 0u) | (out->src ? 4u : 0u) | (out->counts ? 8u : 0u) | (out->post ? 16u : 0u) |
             (out->group_ll ? 32u : 0u) | (out->geo_prior ? 64u : 0u) | (out->source_prior ? 128u : 0u);
#ifdef SB_HOSTSIM
    hostsim::launch([&]() { sb_export_samples_kernel(P, c->d_states, E); }, 2 * n, 64, 0);
#define SB_D2H(dst, srcp, bytes) if (dst) memcpy(dst, srcp, bytes)
#else
    const int blocks = std::max(1, std::min(c->n_sm * 4 / n + 1, (int)((N + 7) / 8)));
    sb_export_samples_kernel<<<dim3(blocks, n), 256, 0, c->stream>>>(P, c->d_states, E);
    CK(cudaGetLastError());
#define SB_D2H(dst, srcp, bytes) if (dst) CK(cudaMemcpyAsync(dst, srcp, bytes, cudaMemcpyDeviceToHost, c->stream))
#endif
    c->launches++;
    SB_D2H(out->cluster_of, E.cluster_of, sizeof(int32_t) * n * N);
    SB_D2H(out->W, E.W, sizeof(float) * n * F * C);
    SB_D2H(out->src, E.src, (size_t)n * N * F);
    SB_D2H(out->counts, E.counts, sizeof(int32_t) * n * nG);
    SB_D2H(out->post, E.post, sizeof(double) * n * 8);
    SB_D2H(out->group_ll, E.glh, sizeof(double) * n * P.G_tot);
    SB_D2H(out->geo_prior, E.geo, sizeof(double) * n * P.K);
    SB_D2H(out->source_prior, E.sp_obj, sizeof(double) * n * N);
#undef SB_D2H
    return dev_sync(c);
}

SB_API int sb_op_stats(sb_ctx *c, int chain, int64_t *accepts, int64_t *rejects, int64_t *cycles) {
    TRY(check_chain(c, chain));
    SbSlot slot; TRY(get_slot(c, chain, &slot));
    for (int i = 0; i < SB_N_OPS; i++) { accepts[i] = slot.accepts[i]; rejects[i] = slot.rejects[i]; if (cycles) cycles[i] = slot.cycles[i]; }
    return 0;
}

SB_API int sb_get_slots(sb_ctx *c, int32_t *state_of_slot) {
    TRY(check_chain(c, 0));
    std::vector<SbSlot> slots(c->n_chains);
    TRY(d2h(c, slots.data(), c->d_slots, sizeof(SbSlot) * c->n_chains));
    for (int i = 0; i < c->n_chains; i++) state_of_slot[i] = slots[i].state;
    return 0;
}
SB_API int sb_set_slots(sb_ctx *c, const int32_t *state_of_slot) {
    TRY(check_chain(c, 0));
    std::vector<char> seen(c->n_chains, 0);
    for (int i = 0; i < c->n_chains; i++) {
        int s = state_of_slot[i];
        if (s < 0 || s >= c->n_chains || seen[s]) return fail(c, "state_of_slot must be a permutation");
        seen[s] = 1;
    }
    for (int i = 0; i < c->n_chains; i++) { TRY(h2d(c, &c->d_slots[i].state, &state_of_slot[i], sizeof(int))); c->h_slot_state[i] = state_of_slot[i]; }
    return 0;
}

// cached posterior terms of EVERY chain (by slot) in one device-to-host copy: what the MC3 swap reads
// between two launches (sbayes/mcmc_setup.py:389-390)
SB_API int sb_posteriors(sb_ctx *c, sb_posterior_t *out) {
    TRY(check_chain(c, 0));
    std::vector<double> post((size_t)c->n_chains * 8);
    TRY(d2h(c, post.data(), c->d_post_all, sizeof(double) * post.size()));
    for (int i = 0; i < c->n_chains; i++) {
        const double *p = &post[(size_t)c->h_slot_state[i] * 8];
        out[i].ll = p[0]; out[i].size_prior = p[1]; out[i].geo_prior = p[2]; out[i].weights_prior = p[3]; out[i].source_prior = p[4];
        out[i].prior = p[1] + p[2] + p[3] + p[4];
    }
    return 0;
}

// new temperatures for every chain (an accepted MC3 swap across GPUs exchanges temperatures, not states):
// stream-ordered copies of 16 bytes per chain, no host synchronisation
SB_API int sb_set_temperatures(sb_ctx *c, const double *temps, const double *prior_temps) {
    TRY(check_chain(c, 0));
    for (int i = 0; i < c->n_chains; i++) {
        if (!(temps[i] > 0) || !(prior_temps[i] > 0)) return fail(c, "temperatures must be positive");
        if (c->h_cfg[i].temperature == temps[i] && c->h_cfg[i].prior_temperature == prior_temps[i]) continue;
        c->h_cfg[i].temperature = temps[i]; c->h_cfg[i].prior_temperature = prior_temps[i];
#ifdef SB_HOSTSIM
        c->d_slots[i].cfg.temperature = temps[i]; c->d_slots[i].cfg.prior_temperature = prior_temps[i];
#else
        static_assert(offsetof(sb_chain_cfg, prior_temperature) == 8, "temperature pair must be adjacent");
        CK(cudaMemcpyAsync(&c->d_slots[i].cfg.temperature, &c->h_cfg[i].temperature, 16, cudaMemcpyHostToDevice, c->stream));
#endif
    }
    return 0;
}

// swap_chains (sbayes/mcmc_setup.py:365-418) on scalars: THE swap implementation of the
// library (sb_swap and the multi-rank ladder of sbayes_b200/mc3.py both end here).
//   ll / prior [n]        un-tempered posterior terms of the state with index s
//   temps / ptemps [n]    ladder: temperature and prior temperature of slot i
//   state_at_slot [n]     in/out: which state sits at slot i; an accepted swap exchanges two entries
// Draws: Philox stream (seed, step = round, SITE_SWAP; sub 0 = partial Fisher-Yates over the
// candidate pairs, sub 1 = uniforms), same addressing as oracle/draws.py: swap_pairs /
// swap_uniform; or, replaying a recording of the reference, tape_pairs [attempts, 2] (slot
// indices in the order the reference drew them) and tape_u [attempts].
static int swap_decide(int n, const double *ll, const double *pr, const double *temps, const double *ptemps, int attempts,
                       int only_neighbours, uint64_t seed, int64_t round, const int32_t *tape_pairs, const double *tape_u,
                       int32_t *state_at_slot, int32_t *accepted_pairs, int32_t *n_accepted) {
    std::vector<std::pair<int, int>> pairs;
    if (only_neighbours) for (int i = 0; i + 1 < n; i++) pairs.push_back({i, i + 1});
    else for (int i = 0; i + 1 < n; i++) for (int j = i + 1; j < n; j++) pairs.push_back({i, j});
    attempts = std::min<int>(attempts, (int)pairs.size());
    const uint32_t k0 = (uint32_t)(seed & 0xFFFFFFFFull), k1 = (uint32_t)(seed >> 32);
    auto uni = [&](uint32_t idx, int sub) {
        uint32_t c1 = (uint32_t)SITE_SWAP | ((uint32_t)sub << 8);
        Philox4 p = philox4x32_10(idx, c1, (uint32_t)((uint64_t)round & 0xFFFFFFFFull), (uint32_t)((uint64_t)round >> 32), k0, k1);
        return u01_double(p.w[0], p.w[1]);
    };
    std::vector<int> idx(pairs.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = (int)i;
    int acc = 0;
    for (int a = 0; a < attempts; a++) {
        int i0, j0;
        if (tape_pairs) { i0 = tape_pairs[2 * a]; j0 = tape_pairs[2 * a + 1]; if (i0 < 0 || j0 < 0 || i0 >= n || j0 >= n) return -1; }
        else {
            int rem = (int)idx.size() - a;
            int j = a + std::min((int)(uni((uint32_t)a, 0) * rem), rem - 1);
            std::swap(idx[a], idx[j]);
            i0 = pairs[idx[a]].first; j0 = pairs[idx[a]].second;
        }
        const int si = state_at_slot[i0], sj = state_at_slot[j0];
        const double r = -((pr[si] - pr[sj]) * (1.0 / ptemps[i0] - 1.0 / ptemps[j0]) + (ll[si] - ll[sj]) * (1.0 / temps[i0] - 1.0 / temps[j0]));
        const double u = tape_u ? tape_u[a] : uni((uint32_t)a, 1);
        const double lu = u > 0 ? log(u) : -INFINITY;
        if (lu < r) {
            state_at_slot[i0] = sj; state_at_slot[j0] = si;
            if (accepted_pairs) { accepted_pairs[2 * acc] = i0; accepted_pairs[2 * acc + 1] = j0; }
            acc++;
        }
    }
    if (n_accepted) *n_accepted = acc;
    return 0;
}

SB_API int sb_swap_scalars(int32_t n, const double *ll, const double *prior, const double *temps, const double *prior_temps,
                           int32_t attempts, int32_t only_neighbours, uint64_t seed, int64_t round, const int32_t *tape_pairs,
                           const double *tape_u, int32_t *state_at_slot, int32_t *accepted_pairs, int32_t *n_accepted) {
    if (n <= 0 || !ll || !prior || !temps || !prior_temps || !state_at_slot) return -1;
    return swap_decide(n, ll, prior, temps, prior_temps, attempts, only_neighbours, seed, round, tape_pairs, tape_u,
                       state_at_slot, accepted_pairs, n_accepted);
}

// swap of the chains of ONE ctx: the states stay where they are in HBM; a swap exchanges which
// state a temperature slot points at.  Decisions on the host from the per-chain scalars.
SB_API int sb_swap(sb_ctx *c, int attempts, int only_neighbours, uint64_t seed, int64_t round,
                   int32_t *state_of_slot, int32_t *n_accepted) {
    TRY(check_chain(c, 0));
    TRY(dev_sync(c));
    const int n = c->n_chains;
    std::vector<SbSlot> slots(n);
    TRY(d2h(c, slots.data(), c->d_slots, sizeof(SbSlot) * n));
    std::vector<double> ll(n), pr(n), T(n), Tp(n);
    std::vector<int32_t> sas(n);
    for (int s = 0; s < n; s++) {
        double post[8];
        TRY(d2h(c, post, c->h_states[s].post, sizeof(post)));
        ll[s] = post[0]; pr[s] = post[1] + post[2] + post[3] + post[4];
        T[s] = slots[s].cfg.temperature; Tp[s] = slots[s].cfg.prior_temperature; sas[s] = slots[s].state;
    }
    int acc = 0;
    if (swap_decide(n, ll.data(), pr.data(), T.data(), Tp.data(), attempts, only_neighbours, seed, round, nullptr, nullptr,
                    sas.data(), nullptr, &acc)) return fail(c, "swap failed");
    for (int i = 0; i < n; i++) {
        if (sas[i] != slots[i].state) TRY(h2d(c, &c->d_slots[i].state, &sas[i], sizeof(int)));
        c->h_slot_state[i] = sas[i];
        if (state_of_slot) state_of_slot[i] = sas[i];
    }
    if (n_accepted) *n_accepted = acc;
    return 0;
}

// device address and length (int32 elements) of the statistics block of all states, laid out
// [state][counts G_tot*F*S | sizes K | H n_pat*2*F*C | pad]; layout3 = {n_counts, n_hist, n_sizes}.
// A sharded run all-reduces the leading counts | sizes part (the histogram is shard-local)
SB_API int sb_stats_buffer(sb_ctx *c, uint64_t *dev_ptr, int64_t *n_int32, int64_t *stride, int64_t *layout3) {
    if (!c) return -1;
    if (dev_ptr) *dev_ptr = (uint64_t)(uintptr_t)c->d_stats;
    if (n_int32) *n_int32 = (int64_t)(c->stats_stride * c->n_chains);
    if (stride) *stride = (int64_t)c->stats_stride;
    if (layout3) { layout3[0] = (int64_t)c->P.G_tot * c->P.F * c->P.S; layout3[1] = (int64_t)c->P.n_pat * 2 * c->P.F * c->P.C; layout3[2] = c->P.K; }
    return 0;
}

SB_API int sb_info(const sb_ctx *c, int32_t *out8) {
    if (!c) return -1;
    out8[0] = c->device; out8[1] = c->n_chains; out8[2] = c->team; out8[3] = c->grid; out8[4] = c->block;
    out8[5] = (int)c->smem; out8[6] = c->n_sm; out8[7] = (int)c->launches;
    return 0;
}

SB_API uint64_t sb_stream(const sb_ctx *c) { return c ? (uint64_t)(uintptr_t)c->stream : 0; }
SB_API int64_t sb_clock_khz(const sb_ctx *c) { return c ? (int64_t)c->clock_khz : 0; }


#ifdef SB_HOSTSIM
// emulator builds only (not part of the ABI): calls of the tiled passes per kind {alter, wide, jump, alter on feature pairs}
SB_API long sb_hostsim_tiled_calls(int kind) { return kind >= 0 && kind < 4 ? uc4::g_tiled_calls[kind] + sb_uc8_tiled_calls(kind) + sb_uc3_tiled_calls(kind) : -1; }
#endif

#if defined(SB_PHASE_PROF) && !defined(SB_HOSTSIM)
// debug builds only (not part of the ABI): phase cycle counters of the step kernel
SB_API int sb_debug_phases(sb_ctx *c, long long *phase /* [SB_N_OPS * SB_NPH] */, long long *n /* [SB_N_OPS] */, int reset) {
    if (!c) return -1;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpyFromSymbol(phase, uc4::g_phase, sizeof(long long) * SB_N_OPS * SB_NPH));
    CK(cudaMemcpyFromSymbol(n, uc4::g_phase_n, sizeof(long long) * SB_N_OPS));
    if (reset) {
        static long long z[SB_N_OPS * SB_NPH];
        CK(cudaMemcpyToSymbol(uc4::g_phase, z, sizeof(long long) * SB_N_OPS * SB_NPH));
        CK(cudaMemcpyToSymbol(uc4::g_phase_n, z, sizeof(long long) * SB_N_OPS));
    }
    return 0;
}
#endif
