// sb_eval.cuh -- stand-alone full evaluation of every chain of a ctx: the uncached
//   Model.likelihood(sample, caching=False) + Model.prior(sample, caching=False)
// of the reference (sbayes/model/model.py:47-51) as three kernels
//
//   sb_eval_hist_kernel    one pass over the packed observations (2 B per observation:
//                          state byte + source byte): the count tensors
//                          (recalculate_feature_counts, sbayes/sampling/counts.py:10-52), the
//                          (has-pattern, feature, component) histogram of the source array (the
//                          sufficient statistic of SourcePrior, sbayes/model/prior.py:568-611)
//                          and the cluster sizes.  HBM bound.
//   sb_eval_geo_kernel     GeoPrior.__call__ (prior.py:769-805,950-962): Prim's minimum spanning
//                          tree of every cluster over the cost matrix (cost_based configs only)
//   sb_eval_terms_kernel   Dirichlet-categorical log-likelihood of the counts
//                          (sbayes/model/likelihood.py:47-101, sbayes/util.py:1373-1394), source /
//                          weights / size priors (prior.py:522-558,568-611,647-681)
//
// The statistics of the first kernel are plain int32 sums over objects: with the objects of a
// very large matrix sharded over GPUs they are what gets all-reduced (sbayes_b200/shard.py).
#pragma once
#include "sb_platform.cuh"
#include "sb_types.h"
#include "sb_plans.h"
#include "sb_philox.cuh"
#include <stdlib.h>

#define SB_EVAL_BLOCK 256
#define SB_EVAL_MAXPARTS 512


struct SbEvalArgs {
    SbProblem P;
    const SbShared *states;      // [n_states]
    const SbSlot *slots;         // [n_states] chain configuration (sample_from_prior)
    int n_states;
    SbEvalPlan plan;
    double *partial;             // [n_states, n_parts, 4]
    unsigned *ticket;            // [n_states] zeroed before the launch
    double *geo_val;             // [n_states, K] per-cluster geo prior (geo only)
    double *post;                // [n_states, 8] {ll, size, geo, weights, source, 0, 0, 0}
    int *err;                    // [1]
    long long n_global;          // number of objects of the whole problem (== P.N unless sharded)
    int geo_cap;                 // capacity of the shared-memory member list (geo only)
};

struct sb_u4 { uint32_t x, y, z, w; };
SB_DEV sb_u4 ev_load16(const uint8_t *p) {      // p is 16-byte aligned
#ifdef SB_HOSTSIM
    sb_u4 v; memcpy(&v, p, 16); return v;
#else
    const uint4 t = *(const uint4 *)p; sb_u4 v; v.x = t.x; v.y = t.y; v.z = t.z; v.w = t.w; return v;
#endif
}
#ifdef SB_HOSTSIM
SB_DEV void ev_store_f32x4(float *p, float a, float b, float c, float d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
SB_DEV void ev_store16(uint8_t *p, const sb_u4 &v) { memcpy(p, &v, 16); }
#else
SB_DEV void ev_store_f32x4(float *p, float a, float b, float c, float d) { *(float4 *)p = make_float4(a, b, c, d); }   // p is 16-byte aligned
SB_DEV void ev_store16(uint8_t *p, const sb_u4 &v) { *(uint4 *)p = make_uint4(v.x, v.y, v.z, v.w); }                 // p is 16-byte aligned
#endif
SB_DEV int ev_arow(const SbProblem &P, int g) { return g < P.K ? (P.dedup ? 0 : g) : g - P.arow_off; }

// -----------------------------------------------------------------------------------------
// histogram pass.  CTA = (state, feature tile, object block).  The tile's bins live in shared
// memory: counts [G_tot][ft*S (+pad)], H [n_pat*2][ft*C (+pad)] with odd row strides (rows of
// different groups fall into different banks).  A warp takes 32 consecutive objects and walks
// their rows in 16-byte chunks: in one instruction all lanes look at the same feature, so the
// bins they hit differ only by group / state (or coincide, which the hardware merges); the
// group row of an observation is selected branch-free from registers.  The non-zero bins are
// added to the state's int32 statistics at the end.
// -----------------------------------------------------------------------------------------
#define SB_EV_NEG (-(1 << 28))
SB_HD int ev_odd(int v) { return v | 1; }

// predicated +1 on a shared-memory bin (straight-line: no branch around the atomic)
#ifdef SB_HOSTSIM
typedef int32_t *ev_bins_t;
SB_DEV ev_bins_t ev_bins(int32_t *p) { return p; }
SB_DEV void ev_inc(ev_bins_t b, int idx, bool pred) { if (pred) b[idx]++; }
#else
typedef uint32_t ev_bins_t;
SB_DEV ev_bins_t ev_bins(int32_t *p) { return (uint32_t)__cvta_generic_to_shared(p); }
SB_DEV void ev_inc(ev_bins_t b, int idx, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %1, 0;\n\t@p red.shared.add.u32 [%0], 1;\n\t}" ::"r"(b + ((uint32_t)idx << 2)), "r"((int)pred) : "memory");
}
#endif

template <int UC>
SB_DEV void ev_hist_word(ev_bins_t bins, uint32_t xw, uint32_t sw, int C, int S, int fl0, const int (&gb)[UC], int hb) {
#ifndef SB_HOSTSIM
#pragma unroll
#endif
    for (int j = 0; j < 4; j++) {
        const int x = (xw >> (8 * j)) & 255, c = (sw >> (8 * j)) & 255;
        int sel = gb[UC - 1];
#ifndef SB_HOSTSIM
#pragma unroll
#endif
        for (int u = UC - 2; u >= 0; u--) sel = c == u ? gb[u] : sel;
        const bool ok = x != SB_NA_U8 && c < C;
        const int fl = fl0 + j;
        const int ci = sel + fl * S + x;
        ev_inc(bins, ci, ok && ci >= 0);
        ev_inc(bins, hb + fl * C + c, ok);
    }
}

template <int UC>
SB_DEV void ev_hist_body(const SbEvalArgs &A) {
    const SbProblem &P = A.P;
    const SbEvalPlan &pl = A.plan;
    const int F = P.F, S = P.S, C = P.C, nconf = P.n_conf;
    int w = sb_cta();
    const int oblk = w % pl.n_oblk; w /= pl.n_oblk;
    const int tile = w % pl.n_ftiles; w /= pl.n_ftiles;
    const SbShared st = A.states[w];
    const int f0 = tile * pl.ft;                                  // first byte column of the tile
    const int ftw = (P.ldx - f0 < pl.ft ? P.ldx - f0 : pl.ft);    // tile width in bytes (multiple of 16)
    const int ft = pl.ft;
    const int gs = ev_odd(ft * S), hs = ev_odd(ft * C);           // padded row strides
    const int lo = oblk * pl.rows_per_blk, hi = lo + pl.rows_per_blk < P.N ? lo + pl.rows_per_blk : P.N;
    int32_t *cnt = (int32_t *)sb_smem();
    const int n_cnt = P.G_tot * gs, n_H = P.n_pat * 2 * hs;
    int32_t *H = cnt + n_cnt;
    int32_t *sz = H + n_H;                                        // [K] (tile 0 only)
    const int n_bins = n_cnt + n_H + (tile == 0 ? P.K : 0);
    for (int i = sb_tid(); i < n_bins; i += sb_nthr()) cnt[i] = 0;
    sb_cta_sync();
    const int cpr = ftw >> 4;                                     // chunks per row
    const ev_bins_t bins = ev_bins(cnt);
    const uint8_t *src0 = st.src[0], *src1 = st.src[1];
    for (int o0 = lo + sb_warp() * 32; o0 < hi; o0 += sb_nwarps() * 32) {
        const int o = o0 + sb_lane();
        const bool live = o < hi;
        int gb[UC];
        int hb = 0, k = -1;
        const uint8_t *xr = P.x, *sr = src0;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
        for (int u = 0; u < UC; u++) gb[u] = SB_EV_NEG;
        if (live) {
            k = st.cluster_of[o];
            const int combo = sb_ldg(P.combo_id + o);
            const int32_t *rows = P.combo_rows + (size_t)combo * nconf;
            gb[0] = k >= 0 ? k * gs : SB_EV_NEG;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
            for (int u = 1; u < UC; u++)
                if (u <= nconf) { const int g = sb_ldg(rows + u - 1); gb[u] = g >= 0 ? g * gs : SB_EV_NEG; }
            hb = n_cnt + (sb_ldg(P.combo_pat + combo) * 2 + (k >= 0 ? 1 : 0)) * hs;     // H follows the counts
            const size_t off = (size_t)o * P.ldx + f0;
            xr = P.x + off; sr = (st.osel[o] ? src1 : src0) + off;
            if (tile == 0 && k >= 0) sb_atomic_add_i32(sz + k, 1);
        }
        sb_u4 xv, sv, xn, sn;
        xn.x = xn.y = xn.z = xn.w = 0xFFFFFFFFu; sn = xn;
        if (live) { xn = ev_load16(xr); sn = ev_load16(sr); }
        for (int ch = 0; ch < cpr; ch++) {
            xv = xn; sv = sn;
            if (live && ch + 1 < cpr) { xn = ev_load16(xr + (ch + 1) * 16); sn = ev_load16(sr + (ch + 1) * 16); }
            const int fl = ch << 4;
            ev_hist_word<UC>(bins, xv.x, sv.x, C, S, fl, gb, hb);
            ev_hist_word<UC>(bins, xv.y, sv.y, C, S, fl + 4, gb, hb);
            ev_hist_word<UC>(bins, xv.z, sv.z, C, S, fl + 8, gb, hb);
            ev_hist_word<UC>(bins, xv.w, sv.w, C, S, fl + 12, gb, hb);
        }
    }
    sb_cta_sync();
    // flush: shared bin (g, fl, s) -> counts[(g*F + f0 + fl)*S + s]; padding columns hold nothing
    const int fv = F - f0 < ftw ? F - f0 : ftw;                    // real features in the tile
    for (int i = sb_tid(); i < n_cnt; i += sb_nthr()) {
        const int v = cnt[i];
        if (v == 0) continue;
        const int g = i / gs, r = i % gs, fl = r / S, s = r % S;
        if (fl < fv) sb_atomic_add_i32(st.counts + ((size_t)g * F + f0 + fl) * S + s, v);
    }
    for (int i = sb_tid(); i < n_H; i += sb_nthr()) {
        const int v = H[i];
        if (v == 0) continue;
        const int pz = i / hs, r = i % hs, fl = r / C, c = r % C;
        if (fl < fv) sb_atomic_add_i32(st.H + ((size_t)pz * F + f0 + fl) * C + c, v);
    }
    if (tile == 0)
        for (int k = sb_tid(); k < P.K; k += sb_nthr()) if (sz[k]) sb_atomic_add_i32(st.sizes + k, sz[k]);
}

// Fast path (at most 4 components, fewer than 255 group rows -- every BASELINE config): the group
// rows of an object travel as four bytes of one register and are picked with a byte permute; both
// atomics are unconditional, observations that count nowhere (missing value, component without a
// group) are redirected to a trash word.
#ifdef SB_HOSTSIM
SB_DEV uint32_t ev_byte(uint32_t w, uint32_t i) { return i < 4 ? (w >> (8 * i)) & 255u : 0u; }
SB_DEV void ev_inc1(ev_bins_t b, int idx) { b[idx]++; }
#else
SB_DEV uint32_t ev_byte(uint32_t w, uint32_t i) { return __byte_perm(w, 0u, 0x4440u | i); }   // i > 3 selects the zero operand
SB_DEV void ev_inc1(ev_bins_t b, int idx) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(b + ((uint32_t)idx << 2)) : "memory"); }
#endif

SB_DEV void ev_fast_word(ev_bins_t bins, uint32_t xw, uint32_t sw, int C, int S, int gs, int fl0, uint32_t gw, int hb, int trash) {
#ifndef SB_HOSTSIM
#pragma unroll
#endif
    for (uint32_t j = 0; j < 4; j++) {
        const uint32_t x = ev_byte(xw, j), c = ev_byte(sw, j);
        const uint32_t g = ev_byte(gw, c);
        const bool ok = x != SB_NA_U8;
        const int fl = fl0 + (int)j;
        const int ci = (int)g * gs + (fl * S + (int)x);
        const int hi = hb + fl * C + (int)c;
        ev_inc1(bins, (ok && g != 255u) ? ci : trash);
        ev_inc1(bins, ok ? hi : trash);
    }
}

SB_DEV void ev_hist_fast(const SbEvalArgs &A) {
    const SbProblem &P = A.P;
    const SbEvalPlan &pl = A.plan;
    const int F = P.F, S = P.S, C = P.C;
    int w = sb_cta();
    const int oblk = w % pl.n_oblk; w /= pl.n_oblk;
    const int tile = w % pl.n_ftiles; w /= pl.n_ftiles;
    const SbShared st = A.states[w];
    const int f0 = tile * pl.ft;
    const int ftw = (P.ldx - f0 < pl.ft ? P.ldx - f0 : pl.ft);
    const int ft = pl.ft;
    const int gs = ev_odd(ft * S), hs = ev_odd(ft * C);
    const int lo = oblk * pl.rows_per_blk, hi = lo + pl.rows_per_blk < P.N ? lo + pl.rows_per_blk : P.N;
    int32_t *cnt = (int32_t *)sb_smem();
    const int n_cnt = P.G_tot * gs, n_H = P.n_pat * 2 * hs;
    int32_t *H = cnt + n_cnt;
    int32_t *sz = H + n_H;                                        // [K], then the trash word
    const int trash = n_cnt + n_H + P.K;
    for (int i = sb_tid(); i < trash + 1; i += sb_nthr()) cnt[i] = 0;
    sb_cta_sync();
    const int cpr = ftw >> 4;
    const ev_bins_t bins = ev_bins(cnt);
    const uint8_t *src0 = st.src[0], *src1 = st.src[1];
    // software pipeline over the warp's trips (see ev_hist_fine): per-object words of trip t + 2 and the first chunks of
    // trip t + 1 are in flight while trip t is counted
    const int stride = sb_nwarps() * 32;
    const int o_first = lo + sb_warp() * 32 + sb_lane(), o_last = hi - 1;
    auto clampo = [&](int o) { return o < hi ? o : o_last; };
    if (lo + sb_warp() * 32 < hi) {
        int kA, kB; uint32_t pkA, pkB; uint8_t selA, selB;
        { const int o = clampo(o_first); kA = st.cluster_of[o]; pkA = sb_ldg(P.obj_pack + o); selA = st.osel[o]; }
        { const int o = clampo(o_first + stride); kB = st.cluster_of[o]; pkB = sb_ldg(P.obj_pack + o); selB = st.osel[o]; }
        const uint8_t *xrA, *srA;
        { const size_t off = (size_t)clampo(o_first) * P.ldx + f0; xrA = P.x + off; srA = (selA ? src1 : src0) + off; }
        sb_u4 xn = ev_load16(xrA), sn = ev_load16(srA);
        for (int o = o_first; o - sb_lane() < hi; o += stride) {
            const uint8_t *xrB, *srB;
            { const size_t off = (size_t)clampo(o + stride) * P.ldx + f0; xrB = P.x + off; srB = (selB ? src1 : src0) + off; }
            const sb_u4 xB = ev_load16(xrB), sB = ev_load16(srB);
            int kC; uint32_t pkC; uint8_t selC;
            { const int oc = clampo(o + 2 * stride); kC = st.cluster_of[oc]; pkC = sb_ldg(P.obj_pack + oc); selC = st.osel[oc]; }
            if (o < hi) {                                            // (no warp-wide operation below)
                const int k = kA;
                const uint32_t gw = (uint32_t)(k >= 0 ? k : 255) | (pkA & 0xFFFFFF00u);
                const int hb = n_cnt + ((int)(pkA & 255u) * 2 + (k >= 0 ? 1 : 0)) * hs;
                if (tile == 0 && k >= 0) sb_atomic_add_i32(sz + k, 1);
                sb_u4 xv, sv;
                for (int ch = 0; ch < cpr; ch++) {
                    xv = xn; sv = sn;
                    if (ch + 1 < cpr) { xn = ev_load16(xrA + (ch + 1) * 16); sn = ev_load16(srA + (ch + 1) * 16); }
                    const int fl = ch << 4;
                    ev_fast_word(bins, xv.x, sv.x, C, S, gs, fl, gw, hb, trash);
                    ev_fast_word(bins, xv.y, sv.y, C, S, gs, fl + 4, gw, hb, trash);
                    ev_fast_word(bins, xv.z, sv.z, C, S, gs, fl + 8, gw, hb, trash);
                    ev_fast_word(bins, xv.w, sv.w, C, S, gs, fl + 12, gw, hb, trash);
                }
            }
            kA = kB; pkA = pkB; xrA = xrB; srA = srB; xn = xB; sn = sB;
            kB = kC; pkB = pkC; selB = selC;
        }
    }
    sb_cta_sync();
    // flush: the first fv*S words of bin row g are exactly the contiguous global segment
    // counts[g, f0 .. f0+fv, :] (same for H with C), so no index arithmetic per bin
    const int fv = F - f0 < ftw ? F - f0 : ftw;
    for (int g = sb_warp(); g < P.G_tot; g += sb_nwarps()) {
        const int32_t *row = cnt + g * gs;
        int32_t *dst = st.counts + ((size_t)g * F + f0) * S;
        for (int r = sb_lane(); r < fv * S; r += 32) { const int v = row[r]; if (v) sb_atomic_add_i32(dst + r, v); }
    }
    for (int pz = sb_warp(); pz < P.n_pat * 2; pz += sb_nwarps()) {
        const int32_t *row = H + pz * hs;
        int32_t *dst = st.H + ((size_t)pz * F + f0) * C;
        for (int r = sb_lane(); r < fv * C; r += 32) { const int v = row[r]; if (v) sb_atomic_add_i32(dst + r, v); }
    }
    if (tile == 0)
        for (int k = sb_tid(); k < P.K; k += sb_nthr()) if (sz[k]) sb_atomic_add_i32(st.sizes + k, sz[k]);
}

// -----------------------------------------------------------------------------------------
// Fine-row statistics pass (every BASELINE config: at most 4 components, fewer than 255 fine rows).
// ONE shared-memory increment per observation instead of two: the bins are indexed by a *fine row*
//   (cluster k, has-pattern)                     for component 0 of a cluster member
//   (no cluster, has-pattern)                    for component 0 of a free object (counts nowhere, H only)
//   (confounder group, has-pattern, in-cluster?) for component c >= 1
// so that both the count tensors (sum of the fine rows of a group) and the source histogram H[pattern,
// in-cluster, f, c] (sum over the fine rows of (pattern, in-cluster, c) and over the states) are sums of
// bins, taken once per CTA in the flush.  Missing values go to an extra bin per (row, feature) (`xs`: the
// state bytes with NA and padding mapped to S), components without a group to a trash row: the inner
// loop has no predicate and no select.  Per observation: byte extracts of the state and of the
// (component + 1) selector, a byte permute that picks the fine row from two registers of the object, one
// multiply-add, one shifted add, one `red.shared` whose feature offset is an immediate.
// With one object block per (state, tile) -- every BASELINE shape -- a CTA owns its slice of the
// statistics block and flushes with plain stores: no memset, no global atomics.
// -----------------------------------------------------------------------------------------
struct SbFineArgs {
    const uint8_t *xs;          // [N, ldx] state bytes, NA / padding = S
    const uint32_t *obj_fine;   // [N] byte 0 = has-pattern, bytes 1..3 = fine row (not-in-cluster flavour) of the object's group
                                //     in confounder 1..3; the in-cluster flavour is the next row; trash row if none
    const int32_t *cnt_ptr, *cnt_idx;   // CSR: fine rows of count row g            [G_tot + 1]
    const int32_t *h_ptr, *h_idx;       // CSR: fine rows of H row (pz * C + c)     [n_pat * 2 * C + 1]
    int n_rows;                 // fine rows including the two trash rows
    int ft, n_ftiles, n_oblk, rows_per_blk;
    unsigned smem;
};

#ifdef SB_HOSTSIM
SB_DEV uint32_t ev_prmt(uint32_t a, uint32_t b, uint32_t s) {
    const uint64_t pool = (uint64_t)a | ((uint64_t)b << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) r |= (uint32_t)((pool >> (8 * ((s >> (4 * i)) & 7))) & 255u) << (8 * i);
    return r;
}
template <int OFF> SB_DEV void ev_inc_at(ev_bins_t b, uint32_t byte_addr) { b[(byte_addr >> 2) + OFF]++; }
SB_DEV uint32_t ev_bins_base(ev_bins_t) { return 0u; }
#else
SB_DEV uint32_t ev_prmt(uint32_t a, uint32_t b, uint32_t s) {     // (selector nibbles are 0..7 here: no masking, unlike __byte_perm)
    uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d;
}
template <int OFF> SB_DEV void ev_inc_at(ev_bins_t, uint32_t byte_addr) {
    asm volatile("red.shared.add.u32 [%0+%1], 1;" ::"r"(byte_addr), "n"(OFF * 4) : "memory");
}
SB_DEV uint32_t ev_bins_base(ev_bins_t b) { return b; }
#endif

// the four observations of one word; t0 = byte address of the chunk's first feature in row 0
template <int S1, int J0>
SB_DEV void ev_fine_word(ev_bins_t bins, uint32_t xw, uint32_t vw, uint32_t lo, uint32_t hi, uint32_t gs4, uint32_t t0) {
#define SB_FINE_OBS(j)                                                                                   \
    {                                                                                                    \
        const uint32_t v = ev_prmt(vw, 0u, 0x4440u | (j)), x = ev_prmt(xw, 0u, 0x4440u | (j));            \
        const uint32_t g = ev_prmt(lo, hi, v);                                                           \
        ev_inc_at<(J0 + (j)) * S1>(bins, g * gs4 + t0 + (x << 2));                                        \
    }
    SB_FINE_OBS(0) SB_FINE_OBS(1) SB_FINE_OBS(2) SB_FINE_OBS(3)
#undef SB_FINE_OBS
}

template <int S1>
SB_DEV void ev_hist_fine(const SbEvalArgs &A, const SbFineArgs &FA) {
    const SbProblem &P = A.P;
    const int F = P.F, S = P.S, C = P.C, K = P.K;
    int w = sb_cta();
    const int oblk = w % FA.n_oblk; w /= FA.n_oblk;
    const int tile = w % FA.n_ftiles; w /= FA.n_ftiles;
    const SbShared st = A.states[w];
    const int f0 = tile * FA.ft;
    const int ftw = (P.ldx - f0 < FA.ft ? P.ldx - f0 : FA.ft);
    const int gs = ev_odd(FA.ft * S1);
    const int lo_o = oblk * FA.rows_per_blk, hi_o = lo_o + FA.rows_per_blk < P.N ? lo_o + FA.rows_per_blk : P.N;
    int32_t *cnt = (int32_t *)sb_smem();
    const int n_bins = FA.n_rows * gs;
    int32_t *sz = cnt + n_bins;                                   // [K]
    // the row lists of the flush, staged once (a flush element would otherwise walk a chain of dependent global loads)
    const int n_cp = P.G_tot + 1, n_hp = P.n_pat * 2 * C + 1;
    int32_t *t_cp = sz + K, *t_hp = t_cp + n_cp, *t_ci = t_hp + n_hp, *t_hi = t_ci + FA.n_rows;
    for (int i = sb_tid(); i < n_bins + K; i += sb_nthr()) cnt[i] = 0;
    for (int i = sb_tid(); i < n_cp; i += sb_nthr()) t_cp[i] = sb_ldg(FA.cnt_ptr + i);
    for (int i = sb_tid(); i < n_hp; i += sb_nthr()) t_hp[i] = sb_ldg(FA.h_ptr + i);
    for (int i = sb_tid(); i < FA.n_rows; i += sb_nthr()) {       // (each list has at most n_rows entries)
        t_ci[i] = i < sb_ldg(FA.cnt_ptr + P.G_tot) ? sb_ldg(FA.cnt_idx + i) : 0;
        t_hi[i] = i < sb_ldg(FA.h_ptr + n_hp - 1) ? sb_ldg(FA.h_idx + i) : 0;
    }
    sb_cta_sync();
    const int cpr = ftw >> 4;
    const ev_bins_t bins = ev_bins(cnt);
    const uint32_t gs4 = (uint32_t)gs * 4u, base = ev_bins_base(bins);
    const uint8_t *src0 = st.src[0], *src1 = st.src[1];
    const uint32_t np = (uint32_t)P.n_pat;
    // software pipeline over the warp's trips (32 objects each): the per-object words (cluster, fine rows, committed
    // copy) of trip t + 2 and the first chunks of trip t + 1 are in flight while trip t is counted -- the source row's
    // address depends on the committed-copy byte, two dependent global loads that would otherwise be exposed per trip
    const int stride = sb_nwarps() * 32;
    const int o_first = lo_o + sb_warp() * 32 + sb_lane(), o_last = hi_o - 1;
    auto clampo = [&](int o) { return o < hi_o ? o : o_last; };
    if (lo_o + sb_warp() * 32 < hi_o) {
        int kA, kB; uint32_t fineA, fineB; uint8_t selA, selB;
        { const int o = clampo(o_first); kA = st.cluster_of[o]; fineA = sb_ldg(FA.obj_fine + o); selA = st.osel[o]; }
        { const int o = clampo(o_first + stride); kB = st.cluster_of[o]; fineB = sb_ldg(FA.obj_fine + o); selB = st.osel[o]; }
        const uint8_t *xrA, *srA;
        { const size_t off = (size_t)clampo(o_first) * P.ldx + f0; xrA = FA.xs + off; srA = (selA ? src1 : src0) + off; }
        sb_u4 xn = ev_load16(xrA), sn = ev_load16(srA);
        for (int o = o_first; o - sb_lane() < hi_o; o += stride) {
            // trip t + 1: row pointers from the words requested a trip ago, first chunks requested now; trip t + 2: words requested
            const uint8_t *xrB, *srB;
            { const size_t off = (size_t)clampo(o + stride) * P.ldx + f0; xrB = FA.xs + off; srB = (selB ? src1 : src0) + off; }
            const sb_u4 xB = ev_load16(xrB), sB = ev_load16(srB);
            int kC; uint32_t fineC; uint8_t selC;
            { const int oc = clampo(o + 2 * stride); kC = st.cluster_of[oc]; fineC = sb_ldg(FA.obj_fine + oc); selC = st.osel[oc]; }
            if (o < hi_o) {                                          // (no warp-wide operation below)
                const int k = kA;
                const uint32_t in = k >= 0 ? 1u : 0u;
                const uint32_t rc = (uint32_t)(k >= 0 ? k : K) * np + (fineA & 255u);
                const uint32_t wv = fineA + in * 0x01010100u;        // confounder rows: the in-cluster flavour is the next one
                const uint32_t glo = ((wv << 8) & 0xFFFF0000u) | (rc << 8);   // selector 1 -> cluster row, 2 / 3 -> confounder 1 / 2
                const uint32_t ghi = wv >> 24;                       // selector 4 -> confounder 3
                if (tile == 0 && k >= 0) sb_atomic_add_i32(sz + k, 1);
                sb_u4 xv, sv;
                for (int ch = 0; ch < cpr; ch++) {
                    xv = xn; sv = sn;
                    if (ch + 1 < cpr) { xn = ev_load16(xrA + (ch + 1) * 16); sn = ev_load16(srA + (ch + 1) * 16); }
                    const uint32_t t0 = base + (uint32_t)(ch * 16 * S1 * 4);
                    // component byte c (255 where the value is missing: its state byte then points at the extra bin) -> selector c + 1
                    ev_fine_word<S1, 0>(bins, xv.x, (sv.x & 0x03030303u) + 0x01010101u, glo, ghi, gs4, t0);
                    ev_fine_word<S1, 4>(bins, xv.y, (sv.y & 0x03030303u) + 0x01010101u, glo, ghi, gs4, t0);
                    ev_fine_word<S1, 8>(bins, xv.z, (sv.z & 0x03030303u) + 0x01010101u, glo, ghi, gs4, t0);
                    ev_fine_word<S1, 12>(bins, xv.w, (sv.w & 0x03030303u) + 0x01010101u, glo, ghi, gs4, t0);
                }
            }
            kA = kB; fineA = fineB; xrA = xrB; srA = srB; xn = xB; sn = sB;
            kB = kC; fineB = fineC; selB = selC;
        }
    }
    sb_cta_sync();
    // flush: every element of this tile's slice of the statistics block is a sum of bins
    const int fv = F - f0 < ftw ? F - f0 : ftw;                    // real features in the tile
    const bool own = FA.n_oblk == 1;                               // the CTA owns the slice: plain stores, nothing was zeroed
    if (fv > 0) {
        for (int g = sb_warp(); g < P.G_tot; g += sb_nwarps()) {       // warps over groups, lanes over the (feature, state) row
            const int q0 = t_cp[g], q1 = t_cp[g + 1];
            int32_t *drow = st.counts + ((size_t)g * F + f0) * S;
            for (int r = sb_lane(), fl = sb_lane() / S, s0 = sb_lane() - (sb_lane() / S) * S; r < fv * S; r += 32) {
                int v = 0;
                for (int q = q0; q < q1; q++) v += cnt[t_ci[q] * gs + fl * S1 + s0];
                if (own) drow[r] = v; else if (v) sb_atomic_add_i32(drow + r, v);
                s0 += 32 % S; fl += 32 / S; if (s0 >= S) { s0 -= S; fl++; }
            }
        }
        const int n_h = P.n_pat * 2 * fv * C;
        for (int i = sb_tid(); i < n_h; i += sb_nthr()) {
            const int pz = i / (fv * C), r = i - pz * (fv * C), fl = r / C, c = r - fl * C;
            int v = 0;
            for (int q = t_hp[pz * C + c]; q < t_hp[pz * C + c + 1]; q++) {
                const int32_t *b = cnt + t_hi[q] * gs + fl * S1;
                for (int s = 0; s < S; s++) v += b[s];
            }
            int32_t *dst = st.H + ((size_t)pz * F + f0 + fl) * C + c;
            if (own) *dst = v; else if (v) sb_atomic_add_i32(dst, v);
        }
    }
    if (tile == 0)
        for (int k = sb_tid(); k < K; k += sb_nthr()) { if (own) st.sizes[k] = sz[k]; else if (sz[k]) sb_atomic_add_i32(st.sizes + k, sz[k]); }
}

// (capping the kernel at 64 registers for a fourth CTA per SM measured much slower: 122 -> 208 us at config 4, profiles/r4m_scale.txt)
template <int S1>
SB_KERNEL sb_eval_fine_kernel(const SB_GRID_CONSTANT SbEvalArgs A, const SB_GRID_CONSTANT SbFineArgs FA) { ev_hist_fine<S1>(A, FA); }

// (forcing 64 registers for a fourth CTA per SM measured slower: profiles/r1l_eval_smem_sweep.txt)
SB_KERNEL sb_eval_hist_kernel(const SB_GRID_CONSTANT SbEvalArgs A) {
    if (A.P.obj_pack) ev_hist_fast(A);
    else if (A.P.C <= 4) ev_hist_body<4>(A);
    else ev_hist_body<8>(A);
}

// -----------------------------------------------------------------------------------------
// geo prior of one cluster: CTA = (state, cluster).  Same arithmetic and tie-breaking as
// geo_cluster of the step kernel (sb_core.cuh): members in object order, Prim from the first
// member, the first minimum wins, zero-cost edges are skipped like scipy's sparse MST does.
// -----------------------------------------------------------------------------------------
SB_DEV double ev_log_expit(double z) { double m = -z > 0 ? -z : 0.0; return -(m + log(exp(0.0 - m) + exp(-z - m))); }
SB_DEV double ev_geo_prob_fn(const SbProblem &P, double x) {
    if (P.geo_pf == 0) return -x / P.geo_rate;
    return ev_log_expit(-(x - P.geo_infl) / P.geo_rate) - ev_log_expit(P.geo_infl / P.geo_rate);
}

SB_KERNEL sb_eval_geo_kernel(const SB_GRID_CONSTANT SbEvalArgs A) {
    const SbProblem &P = A.P;
    const int k = sb_cta() % P.K, s_idx = sb_cta() / P.K;
    const SbShared st = A.states[s_idx];
    const int cap = A.geo_cap;
    double *dctl = (double *)sb_smem();
    double *best = dctl + 2;
    int *mem = (int *)(best + cap);
    int *ctl = mem + cap;            // [0] m, [1] argmin
    if (sb_tid() == 0) {
        int m = 0;
        for (int o = 0; o < P.N; o++) if (st.cluster_of[o] == k) { if (m < cap) mem[m] = o; m++; }
        ctl[0] = m;
    }
    sb_cta_sync();
    int m = ctl[0];
    if (m > cap) { *A.err = 3; m = cap; }
    double sum = 0.0, mx = 0.0; int cnt = 0;
    if (m > 1) {
        for (int i = sb_tid(); i < m; i += sb_nthr()) best[i] = i == 0 ? -1.0 : P.cost[(size_t)mem[0] * P.N + mem[i]];
        sb_cta_sync();
        for (int it = 0; it < m - 1; it++) {
            if (sb_tid() == 0) {
                int bj = -1; double bv = 0.0;
                for (int i = 0; i < m; i++) { double b = best[i]; if (b >= 0.0 && (bj < 0 || b < bv)) { bj = i; bv = b; } }
                ctl[1] = bj; dctl[0] = bv;
            }
            sb_cta_sync();
            const int j = ctl[1]; const double e = dctl[0];
            sb_cta_sync();
            if (j < 0) break;
            if (e != 0.0 && e <= 1.7976931348623157e308) { sum += e; if (cnt == 0 || e > mx) mx = e; cnt++; }
            for (int i = sb_tid(); i < m; i += sb_nthr()) {
                if (i == j) best[i] = -1.0;
                else if (best[i] >= 0.0) { double c = P.cost[(size_t)mem[j] * P.N + mem[i]]; if (c < best[i]) best[i] = c; }
            }
            sb_cta_sync();
        }
    }
    double agg = 0.0;
    if (cnt > 0) agg = P.geo_agg == 0 ? sum / (double)cnt : (P.geo_agg == 1 ? sum : mx);
    if (sb_tid() == 0) A.geo_val[(size_t)s_idx * P.K + k] = ev_geo_prob_fn(P, agg);
}

// -----------------------------------------------------------------------------------------
// posterior terms from the statistics.  CTA = (state, part); the per-part sums are combined by
// the last CTA of the state in part order, so the result does not depend on the schedule.
// -----------------------------------------------------------------------------------------
SB_DEV double ev_warp_sum(double v) {
    for (int d = 16; d > 0; d >>= 1) v += sb_shfl_f64(v, (sb_lane() + d) & 31);
    return v;
}

SB_KERNEL sb_eval_terms_kernel(const SB_GRID_CONSTANT SbEvalArgs A) {
    const SbProblem &P = A.P;
    const int F = P.F, S = P.S, C = P.C;
    const int n_parts = A.plan.n_parts;
    const int part = sb_cta() % n_parts, s_idx = sb_cta() / n_parts;
    const SbShared st = A.states[s_idx];
    const long long n_ll = (long long)P.G_tot * F, n_sp = (long long)P.n_pat * 2 * F, n_wp = P.wprior_uniform ? 0 : F;
    const long long total = n_ll + n_sp + n_wp, per = (total + n_parts - 1) / n_parts;
    const long long lo = part * per, hi = lo + per < total ? lo + per : total;
    double ll = 0.0, sp = 0.0, wp = 0.0;
    for (long long it = lo + sb_tid(); it < hi; it += sb_nthr()) {
        if (it < n_ll) {
            // Dirichlet-categorical term of one (group, feature) row, sbayes/util.py:1373-1394
            const int g = (int)(it / F), f = (int)(it % F);
            const size_t ar = (size_t)ev_arow(P, g) * F + f;
            const int32_t *cr = st.counts + (size_t)it * S;
            const double *al = P.alpha + ar * S;
            // lgamma(c + 1) and lgamma(n + integer) come from the log-factorial table (unit
            // concentrations = the uniform prior, the common case); anything else is evaluated
            int n = 0; double ser = 0.0;
            for (int s = 0; s < S; s++) {
                const int c = cr[s]; const double a = sb_ldg(al + s);
                n += c;
                if (a == 1.0 && c < P.lfact_n) ser += sb_ldg(P.lfact + c);
                else if (a > 0.0) ser += lgamma((double)c + a);
            }
            const double as = sb_ldg(P.asum + ar);
            const int ai = (int)as;
            const double lgn = ((double)ai == as && ai >= 1 && n + ai - 1 < P.lfact_n) ? sb_ldg(P.lfact + n + ai - 1) : lgamma((double)n + as);
            ll += sb_ldg(P.dcc + ar) - lgn + ser;
        } else if (it < n_ll + n_sp) {
            // SourcePrior: sum over the histogram of log normalised weights (prior.py:596-611,
            // normalize_weights likelihood.py:171-190)
            const long long i = it - n_ll;
            const int f = (int)(i % F), pz = (int)(i / F), z = pz & 1, bits = sb_ldg(P.pat_bits + (pz >> 1));
            const int32_t *h = st.H + (size_t)i * C;
            double t1 = 0.0;
            for (int c = 0; c < C; c++) {
                const double hh = c == 0 ? (double)z : (double)((bits >> c) & 1);
                t1 += hh * (double)st.W[f * C + c];
            }
            for (int c = 0; c < C; c++) {
                const int hv = h[c];
                if (hv == 0) continue;
                const double hh = c == 0 ? (double)z : (double)((bits >> c) & 1);
                const double a = t1 > 0.0 ? hh * (double)st.W[f * C + c] / t1 : 0.0;
                sp += (double)hv * log(a);
            }
        } else {
            // WeightsPrior: Dirichlet log-pdf of one weight row (prior.py:522-558, util.py:49)
            const int f = (int)(it - n_ll - n_sp);
            double t = 0.0;
            for (int c = 0; c < C; c++) {
                const double a = sb_ldg(P.alpha_w + (size_t)f * C + c);
                if (a != 1.0) t += (a - 1.0) * log((double)st.W[f * C + c]);
            }
            wp += -sb_ldg(P.wlnB + f) + t;
        }
    }
    // block reduction in a fixed order
    double *red = (double *)sb_smem();                 // [3, n_warps]
    unsigned *flag = (unsigned *)(red + 3 * 32);
    ll = ev_warp_sum(ll); sp = ev_warp_sum(sp); wp = ev_warp_sum(wp);
    if (sb_lane() == 0) { red[sb_warp()] = ll; red[32 + sb_warp()] = sp; red[64 + sb_warp()] = wp; }
    sb_cta_sync();
    if (sb_tid() == 0) {
        double a = 0.0, b = 0.0, c = 0.0;
        for (int i = 0; i < sb_nwarps(); i++) { a += red[i]; b += red[32 + i]; c += red[64 + i]; }
        double *out = A.partial + ((size_t)s_idx * n_parts + part) * 4;
        out[0] = a; out[1] = b; out[2] = c;
        sb_threadfence();
        const unsigned t = (unsigned)sb_atomic_add_i32((int *)A.ticket + s_idx, 1);
        flag[0] = t == (unsigned)(n_parts - 1);
    }
    sb_cta_sync();
    if (!flag[0] || sb_tid() != 0) return;
    sb_threadfence();
    double a = 0.0, b = 0.0, c = 0.0;
    for (int p = 0; p < n_parts; p++) {
        const volatile double *in = A.partial + ((size_t)s_idx * n_parts + p) * 4;
        a += in[0]; b += in[1]; c += in[2];
    }
    int sfp = 0;
    for (int i = 0; i < A.n_states; i++) if (A.slots[i].state == s_idx) sfp = A.slots[i].cfg.sample_from_prior;
    // ClusterSizePrior (prior.py:647-681, log_multinom util.py:1125-1173)
    double size_pr = 0.0;
    if (P.size_prior == 1) {
        long long tot = 0; double s = 0.0;
        for (int k = 0; k < P.K; k++) { tot += st.sizes[k]; s += lgamma((double)st.sizes[k] + 1.0); }
        size_pr = tot == 0 ? -0.0 : -(lgamma((double)A.n_global + 1.0) - s - lgamma((double)(A.n_global - tot) + 1.0));
    } else if (P.size_prior == 2) {
        double s = 0.0;
        for (int k = 0; k < P.K; k++) { double z = (double)st.sizes[k]; s += log(z * z); }
        size_pr = -s;
    }
    double geo = 0.0;
    if (P.geo) for (int k = 0; k < P.K; k++) geo += A.geo_val[(size_t)s_idx * P.K + k];
    double *post = A.post + (size_t)s_idx * 8;
    post[0] = sfp ? 0.0 : a; post[1] = size_pr; post[2] = geo; post[3] = c; post[4] = b; post[5] = 0.0; post[6] = 0.0; post[7] = 0.0;
}

// -----------------------------------------------------------------------------------------
// observation likelihood: what LikelihoodLogger writes per logged sample
// (sbayes/sampling/loggers.py:354-359):  lh[o,f] = sum_c w~[o,f,c] * l_c[o,f]  with the
// leave-one-out component likelihoods of likelihood_per_component_exact
// (sbayes/sampling/conditionals.py:300-367)
//     l_c[o,f] = (counts[g,f,x] + alpha[g,f,x] - d) / (sum_s(counts + alpha)[g,f] - d),  d = [src[o,f] == c]
// (0 where o is in no group of c; lh = 1 where the observation is missing).
// One streaming pass: 2 B read (state byte, source byte) + 4 B written (float32) per observation.
// CTA = (state, feature tile, object block); the tile's tables (counts + alpha, the two
// reciprocals of the row totals, the normalised weights per has-pattern) are built in shared
// memory in fp64; log(lh) is taken of products of four.  Per-object sums over the tile go to
// obj_part[state][tile][o]; sb_obs_reduce_kernel adds tiles and objects in a fixed order.
// -----------------------------------------------------------------------------------------

struct SbObsArgs {
    SbProblem P;
    const SbShared *states;
    int n_states;
    SbObsPlan plan;
    float *out;          // [n_states, N, F]
    double *obj_part;    // [n_states, n_ftiles, N]
    double *obj_ll;      // [n_states, N]
    double *sums;        // [n_states]
};

static inline int sb_obs_make_plan(const SbProblem &P, int n_states, int n_sm, SbObsPlan *pl) {
    const size_t per_f = ((size_t)P.G_tot * P.S + 2 * (size_t)P.G_tot + (size_t)P.n_pat * 2 * (P.C | 1)) * 8;
    const size_t fixed = 256 + 2 * (size_t)P.G_tot * 8;                                  // + odd-stride padding
    size_t budget = 56 * 1024;
    if (fixed + 16 * per_f > budget) budget = 110 * 1024;
    if (fixed + 16 * per_f > budget) budget = 220 * 1024;
    if (fixed + 16 * per_f > budget) return -1;
    // 16, 32, 64 or 128 features per tile (4 .. 32 lanes per row).  Narrow tiles measured fastest on
    // every BASELINE shape (profiles/r1s_obs_tile_sweep.txt): more CTAs and more rows per warp iteration
    int ft = 16;
#ifndef SB_HOSTSIM
    if (getenv("SB_OBS_FT")) { int v = atoi(getenv("SB_OBS_FT")); if ((v == 16 || v == 32 || v == 64 || v == 128) && fixed + (size_t)v * per_f <= 220 * 1024) ft = v; }   // tuning knob
#endif
    int n_ftiles = (P.ldx + ft - 1) / ft;
    // object blocks per (state, tile): every CTA builds the tile's tables first (about ten
    // operations per table entry), so minimise  waves * (observations per CTA + table cost)
    const size_t smem = fixed + (size_t)ft * per_f;
    long long per_sm = (long long)((227 * 1024) / (smem + 1024)); if (per_sm > 3) per_sm = 3; if (per_sm < 1) per_sm = 1;
    const long long slots = per_sm * n_sm, base = (long long)n_states * n_ftiles;
    long long max_blk = P.N / 128; if (max_blk < 1) max_blk = 1;
    if (max_blk > 4096) max_blk = 4096;
    long long n_oblk = 1; double best = -1.0;
    for (long long n = 1; n <= max_blk; n++) {
        const long long waves = (base * n + slots - 1) / slots;
        const double cost = (double)waves * ((double)((P.N + n - 1) / n) * ft + 10.0 * (double)ft * P.G_tot * P.S);
        if (best < 0.0 || cost < best) { best = cost; n_oblk = n; }
    }
    int rows = (int)((P.N + n_oblk - 1) / n_oblk);
    n_oblk = (P.N + rows - 1) / rows;
    pl->ft = ft; pl->n_ftiles = n_ftiles; pl->n_oblk = (int)n_oblk; pl->rows_per_blk = rows;
    pl->smem = (unsigned)smem;
    return 0;
}

// Work mapping: a lane takes one 16-byte chunk of a row (16 features of one object: one 16-byte load of the states,
// one of the sources, four 16-byte stores of floats), so the per-object set-up (cluster, group rows, weight row, row
// pointers) is paid once per 16 observations; ft/16 lanes share a row, 32 / (ft/16) rows per warp iteration.
// (Round 1 gave a lane 4 features of a row: coalesced single-byte loads, but the set-up -- about 100 instructions --
// was paid per 4 observations.)
// (EXACT: the problem has exactly UC components -- the guards `u < C` of the unrolled component loops fold away)
template <int UC, bool EXACT = false>
SB_DEV void ev_obs_body(const SbObsArgs &A) {
    const SbProblem &P = A.P;
    const SbObsPlan &pl = A.plan;
    const int F = P.F, S = P.S, C = EXACT ? UC : P.C, nconf = P.n_conf, N = P.N;
    int w = sb_cta();
    const int oblk = w % pl.n_oblk; w /= pl.n_oblk;
    const int tile = w % pl.n_ftiles; w /= pl.n_ftiles;
    const int s_idx = w;
    const SbShared st = A.states[s_idx];
    const int ft = pl.ft;                                          // 16, 32, 64 or 128
    const int f0 = tile * ft;
    const int fv = F - f0 < ft ? F - f0 : ft;                      // real features in the tile
    // state rows are dense (S doubles); the state index is rotated by the feature, (x + fl) mod S,
    // so lanes on consecutive features with equal states fall into different banks
    const int Cp = ev_odd(C);
    const int gs = ft * S, rs = ev_odd(ft), hs = ft * Cp;
    double *num = (double *)sb_smem();                             // [G_tot][ft][S]    counts + alpha (rotated)
    double *rd0 = num + (size_t)P.G_tot * gs;                      // [G_tot][rs]       1 / total
    double *rd1 = rd0 + (size_t)P.G_tot * rs;                      // [G_tot][rs]       1 / (total - 1)
    double *wt = rd1 + (size_t)P.G_tot * rs;                       // [n_pat*2][ft][Cp] normalised weights
    for (int i = sb_tid(); i < P.G_tot * ft; i += sb_nthr()) {
        const int g = i / ft, fl = i % ft;
        double tot = 0.0;
        if (fl < fv) {
            const size_t ar = ((size_t)ev_arow(P, g) * F + f0 + fl) * S, cr = ((size_t)g * F + f0 + fl) * S;
            for (int s = 0; s < S; s++) { const double v = (double)st.counts[cr + s] + sb_ldg(P.alpha + ar + s); num[(size_t)g * gs + fl * S + (s + fl) % S] = v; tot += v; }
        }
        rd0[(size_t)g * rs + fl] = tot > 0.0 ? 1.0 / tot : 0.0;
        rd1[(size_t)g * rs + fl] = tot > 1.0 ? 1.0 / (tot - 1.0) : 0.0;
    }
    for (int i = sb_tid(); i < P.n_pat * 2 * ft; i += sb_nthr()) {
        const int pz = i / ft, fl = i % ft, z = pz & 1, bits = sb_ldg(P.pat_bits + (pz >> 1));
        double t1 = 0.0;
        if (fl < fv) for (int c = 0; c < C; c++) t1 += (c == 0 ? (double)z : (double)((bits >> c) & 1)) * (double)st.W[(f0 + fl) * C + c];
        for (int c = 0; c < C; c++) {
            const double hh = c == 0 ? (double)z : (double)((bits >> c) & 1);
            wt[(size_t)pz * hs + fl * Cp + c] = (fl < fv && t1 > 0.0) ? hh * (double)st.W[(f0 + fl) * C + c] / t1 : 0.0;
        }
    }
    sb_cta_sync();
    const int lo = oblk * pl.rows_per_blk, hi = lo + pl.rows_per_blk < N ? lo + pl.rows_per_blk : N;
    const int lpr = ft >> 4, rpw = 32 / lpr;                       // lanes per row (one 16-byte chunk each), rows per warp iteration
    const int fg = sb_lane() % lpr, rw = sb_lane() / lpr;
    const int fb = fg << 4;                                         // first feature (within the tile) of this lane's chunk
    const bool has_chunk = f0 + fb < P.ldx;
    const int r0 = fb % S;
    const bool vec_out = (F & 3) == 0;                              // rows of the float output are 16-byte aligned
    const uint8_t *src0 = st.src[0], *src1 = st.src[1];
    float *out = A.out + (size_t)s_idx * N * F;
    double *opart = A.obj_part + ((size_t)s_idx * pl.n_ftiles + tile) * N;
    for (int o0 = lo + sb_warp() * rpw; o0 < hi; o0 += sb_nwarps() * rpw) {
        const int o = o0 + rw;
        const bool live = o < hi;
        int gb[UC], gr[UC];
#ifndef SB_HOSTSIM
#pragma unroll
#endif
        for (int u = 0; u < UC; u++) { gb[u] = -1; gr[u] = 0; }
        double acc = 0.0;
        if (live && has_chunk) {
            const int k = st.cluster_of[o];
            if (k >= 0) { gb[0] = k * gs; gr[0] = k * rs; }
            int pat;
            if (P.obj_pack) {
                const uint32_t pk = sb_ldg(P.obj_pack + o);
                pat = (int)(pk & 255u);
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                for (int u = 1; u < 4 && u < UC; u++) { const int g = (int)((pk >> (8 * u)) & 255u); if (g != 255) { gb[u] = g * gs; gr[u] = g * rs; } }
            } else {
                const int combo = sb_ldg(P.combo_id + o);
                const int32_t *rows = P.combo_rows + (size_t)combo * nconf;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                for (int u = 1; u < UC; u++)
                    if (u <= nconf) { const int g = sb_ldg(rows + u - 1); if (g >= 0) { gb[u] = g * gs; gr[u] = g * rs; } }
                pat = sb_ldg(P.combo_pat + combo);
            }
            const double *wrow = wt + (size_t)(pat * 2 + (k >= 0 ? 1 : 0)) * hs;
            const size_t off = (size_t)o * P.ldx + f0 + fb;
            const sb_u4 xv = ev_load16(P.x + off), sv = ev_load16((st.osel[o] ? src1 : src0) + off);
            const uint32_t xw4[4] = {xv.x, xv.y, xv.z, xv.w}, sw4[4] = {sv.x, sv.y, sv.z, sv.w};
            float *orow = out + (size_t)o * F + f0 + fb;
            int r = r0;                                              // (fb + i) mod S
#ifndef SB_HOSTSIM
#pragma unroll
#endif
            for (int half = 0; half < 2; half++) {
                double prod = 1.0;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                for (int w2 = 0; w2 < 2; w2++) {
                    const int wi = 2 * half + w2;
                    float o4[4];
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                    for (int j = 0; j < 4; j++) {
                        const int fl = fb + 4 * wi + j;
                        const int x = (int)((xw4[wi] >> (8 * j)) & 255u), cj = (int)((sw4[wi] >> (8 * j)) & 255u);
                        double lh = 1.0;
                        if (x != SB_NA_U8) {
                            lh = 0.0;
                            int xr_ = x + r; xr_ = xr_ >= S ? xr_ - S : xr_;       // (x + fl) mod S
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                            for (int u = 0; u < UC; u++) {
                                if (u < C && gb[u] >= 0) {
                                    const bool d = cj == u;
                                    const double n = num[gb[u] + fl * S + xr_] - (d ? 1.0 : 0.0);
                                    const double rr = d ? rd1[gr[u] + fl] : rd0[gr[u] + fl];
                                    lh += wrow[fl * Cp + u] * (n * rr);
                                }
                            }
                        }
                        o4[j] = (float)lh;
                        if (fl < fv) prod *= lh;                     // (padding bytes are missing values: lh = 1)
                        r = r + 1 == S ? 0 : r + 1;
                    }
                    const int flw = fb + 4 * wi;
                    if (vec_out && flw + 3 < fv) ev_store_f32x4(orow + 4 * wi, o4[0], o4[1], o4[2], o4[3]);
                    else {
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                        for (int j = 0; j < 4; j++) if (flw + j < fv) orow[4 * wi + j] = o4[j];
                    }
                }
                acc += log(prod);
            }
        }
        // sum over the lanes of the row (segments of lpr lanes, lpr a power of two)
        for (int d = lpr >> 1; d > 0; d >>= 1) acc += sb_shfl_f64(acc, sb_lane() ^ d);
        if (live && fg == 0) opart[o] = acc;
    }
}

SB_KERNEL sb_obs_lik_kernel(const SB_GRID_CONSTANT SbObsArgs A) {
    if (A.P.C == 3) ev_obs_body<3, true>(A); else if (A.P.C <= 4) ev_obs_body<4>(A); else ev_obs_body<8>(A);
}

// obj_ll[state][o] = sum over tiles (tile order); sums[state] = sum over objects.  CTA =
// (state, part of the objects); the per-part sums are combined by the last CTA of the state in
// part order, so the chain total does not depend on the schedule.
#define SB_OBS_RPARTS 32
SB_KERNEL sb_obs_reduce_kernel(const SB_GRID_CONSTANT SbObsArgs A, double *partial, unsigned *ticket) {
    const int part = sb_cta() % SB_OBS_RPARTS, s_idx = sb_cta() / SB_OBS_RPARTS, N = A.P.N, nt = A.plan.n_ftiles;
    const int per = (N + SB_OBS_RPARTS - 1) / SB_OBS_RPARTS, lo = part * per, hi = lo + per < N ? lo + per : N;
    double acc = 0.0;
    for (int o = lo + sb_tid(); o < hi; o += sb_nthr()) {
        double v = 0.0;
        for (int t = 0; t < nt; t++) v += A.obj_part[((size_t)s_idx * nt + t) * N + o];
        A.obj_ll[(size_t)s_idx * N + o] = v;
        acc += v;
    }
    double *red = (double *)sb_smem();
    unsigned *flag = (unsigned *)(red + 32);
    acc = ev_warp_sum(acc);
    if (sb_lane() == 0) red[sb_warp()] = acc;
    sb_cta_sync();
    if (sb_tid() == 0) {
        double t = 0.0;
        for (int i = 0; i < sb_nwarps(); i++) t += red[i];
        partial[(size_t)s_idx * SB_OBS_RPARTS + part] = t;
        sb_threadfence();
        flag[0] = (unsigned)sb_atomic_add_i32((int *)ticket + s_idx, 1) == (unsigned)(SB_OBS_RPARTS - 1);
        if (flag[0]) {
            sb_threadfence();
            double tot = 0.0;
            for (int p = 0; p < SB_OBS_RPARTS; p++) tot += ((const volatile double *)partial)[(size_t)s_idx * SB_OBS_RPARTS + p];
            A.sums[s_idx] = tot;
        }
    }
}

// -----------------------------------------------------------------------------------------
// full-sweep source Gibbs: GibbsSampleSource over ALL objects (sbayes/sampling/operators.py:496-574,
// the ObjectSelector.ALL path; used by the initializer, sbayes/sampling/initializers.py:285-310):
//     q[o,f,c]  ~  l_c[o,f]^(1/T) * w~[o,f,c]^(1/Tp),   l_c = plain table of the group of o in c
// for every observation of every chain in ONE streaming launch: 3 B per observation (state byte and
// old source byte read, new source byte written; SURVEY.md 8d GIBBS).  Same decomposition as the
// observation-likelihood kernel: CTA = (state, feature tile, object block), the tile's tables
// (l^(1/T) per group row, w~^(1/Tp) per has-pattern) in shared memory, fp64.  The draw of an
// pair of adjacent features is ONE Philox call addressed by (chain seed, step, SITE_GIBBS, o * ceil(F/2) + f/2):
// words 0,1 make the uniform of the even feature, words 2,3 that of the odd one (oracle/operators.py: gibbs_sweep
// mirrors the addressing); the conditional is that of source_posterior in sb_core.cuh.
//   mode 0  draw: the new sources go to the non-committed copy of the rows (sb_flip_osel_kernel commits
//           them), obj_part receives log q[new] per object and tile, n_changed the reassigned observations
//   mode 1  score the NON-committed copy (the previous sources right after a draw) under the current
//           tables: log_q_back of the proposal
//   mode 2  score the committed sources
// -----------------------------------------------------------------------------------------
struct SbGibbsArgs {
    SbProblem P;
    const SbShared *states;
    const SbSlot *slots;
    int n_states;
    SbObsPlan plan;
    double *obj_part;               // [n_states, n_ftiles, N]
    unsigned long long *n_changed;  // [n_states]
    unsigned long long step;
    int mode;
};

template <int UC, bool EXACT = false>
SB_DEV void ev_gibbs_body(const SbGibbsArgs &A) {
    const SbProblem &P = A.P;
    const SbObsPlan &pl = A.plan;
    const int F = P.F, S = P.S, C = EXACT ? UC : P.C, nconf = P.n_conf, N = P.N;
    int w = sb_cta();
    const int oblk = w % pl.n_oblk; w /= pl.n_oblk;
    const int tile = w % pl.n_ftiles; w /= pl.n_ftiles;
    const int s_idx = w;
    const SbShared st = A.states[s_idx];
    // the slot (temperatures, Philox key) that currently holds this state: the threads search the slots in parallel
    int *slot_of = (int *)sb_smem();
    if (sb_tid() == 0) *slot_of = 0;
    sb_cta_sync();
    for (int i = sb_tid(); i < A.n_states; i += sb_nthr()) if (A.slots[i].state == s_idx) *slot_of = i;
    sb_cta_sync();
    const SbSlot sl = A.slots[*slot_of];
    sb_cta_sync();                                                  // (the tables below reuse the word)
    const double T = sl.cfg.temperature, Tp = sl.cfg.prior_temperature;
    const int sfp = sl.cfg.sample_from_prior; const unsigned long long seed = sl.cfg.seed;
    const double invT = 1.0 / T, invTp = 1.0 / Tp;
    const bool unitT = T == 1.0, unitTp = Tp == 1.0;
    const int ft = pl.ft;
    const int f0 = tile * ft;
    const int fv = F - f0 < ft ? F - f0 : ft;
    const int Cp = ev_odd(C);
    const int gs = ft * S, hs = ft * Cp;
    double *lik = (double *)sb_smem();                             // [G_tot][ft][S]  l^(1/T), state index rotated by the feature
    double *wt = lik + (size_t)P.G_tot * gs;                       // [n_pat*2][ft][Cp] w~^(1/Tp)
    for (int i = sb_tid(); i < P.G_tot * ft; i += sb_nthr()) {
        const int g = i / ft, fl = i % ft;
        if (fl >= fv) continue;
        const size_t ar = (size_t)ev_arow(P, g) * F + f0 + fl, cr = ((size_t)g * F + f0 + fl) * S;
        int n = 0;
        for (int s = 0; s < S; s++) n += st.counts[cr + s];
        const double rP = 1.0 / (sb_ldg(P.asum + ar) + (double)n);
        for (int s = 0; s < S; s++) {
            double l = (sb_ldg(P.alpha + ar * S + s) + (double)st.counts[cr + s]) * rP;
            if (sfp) l = 1.0;
            else if (!unitT) l = l > 0.0 ? exp(invT * log(l)) : 0.0;
            lik[(size_t)g * gs + fl * S + (s + fl) % S] = l;
        }
    }
    for (int i = sb_tid(); i < P.n_pat * 2 * ft; i += sb_nthr()) {
        const int pz = i / ft, fl = i % ft, z = pz & 1, bits = sb_ldg(P.pat_bits + (pz >> 1));
        double t1 = 0.0;
        if (fl < fv) for (int c = 0; c < C; c++) t1 += (c == 0 ? (double)z : (double)((bits >> c) & 1)) * (double)st.W[(f0 + fl) * C + c];
        for (int c = 0; c < C; c++) {
            const double hh = c == 0 ? (double)z : (double)((bits >> c) & 1);
            double a = (fl < fv && t1 > 0.0) ? hh * (double)st.W[(f0 + fl) * C + c] / t1 : 0.0;
            if (!unitTp) a = pow(a, invTp);
            wt[(size_t)pz * hs + fl * Cp + c] = a;
        }
    }
    sb_cta_sync();
    const int lo = oblk * pl.rows_per_blk, hi = lo + pl.rows_per_blk < N ? lo + pl.rows_per_blk : N;
    // a lane takes one 16-byte chunk of a row (16 features of one object; see ev_obs_body): the per-object set-up is paid
    // once per 16 observations, states / old sources arrive in two 16-byte loads, the new sources leave in one 16-byte
    // store; one Philox call serves a PAIR of adjacent features (words 0,1 -> even feature, words 2,3 -> odd)
    const int lpr = ft >> 4, rpw = 32 / lpr;
    const int fg = sb_lane() % lpr, rw = sb_lane() / lpr;
    const int fb = fg << 4;
    const bool has_chunk = f0 + fb < P.ldx;
    const int r0 = fb % S;
    const int hF = (F + 1) >> 1;                                     // Philox calls per row
    uint8_t *src0 = st.src[0], *src1 = st.src[1];
    double *opart = A.obj_part + ((size_t)s_idx * pl.n_ftiles + tile) * N;
    const uint32_t k0 = (uint32_t)(seed & 0xFFFFFFFFull), k1 = (uint32_t)(seed >> 32);
    const uint32_t st_lo = (uint32_t)(A.step & 0xFFFFFFFFull), st_hi = (uint32_t)(A.step >> 32);
    unsigned changed = 0;
    for (int o0 = lo + sb_warp() * rpw; o0 < hi; o0 += sb_nwarps() * rpw) {
        const int o = o0 + rw;
        const bool live = o < hi;
        int gb[UC];
#ifndef SB_HOSTSIM
#pragma unroll
#endif
        for (int u = 0; u < UC; u++) gb[u] = -1;
        double acc = 0.0;
        if (live && has_chunk) {
            const int k = st.cluster_of[o];
            if (k >= 0) gb[0] = k * gs;
            int pat;
            if (P.obj_pack) {
                const uint32_t pk = sb_ldg(P.obj_pack + o);
                pat = (int)(pk & 255u);
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                for (int u = 1; u < 4 && u < UC; u++) { const int g = (int)((pk >> (8 * u)) & 255u); if (g != 255) gb[u] = g * gs; }
            } else {
                const int combo = sb_ldg(P.combo_id + o);
                const int32_t *rows = P.combo_rows + (size_t)combo * nconf;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                for (int u = 1; u < UC; u++)
                    if (u <= nconf) { const int g = sb_ldg(rows + u - 1); if (g >= 0) gb[u] = g * gs; }
                pat = sb_ldg(P.combo_pat + combo);
            }
            const double *wrow = wt + (size_t)(pat * 2 + (k >= 0 ? 1 : 0)) * hs;
            const size_t off = (size_t)o * P.ldx + f0 + fb;
            const int sel = st.osel[o];
            uint8_t *so = (sel ? src0 : src1) + off;
            const sb_u4 xv = ev_load16(P.x + off), sv = ev_load16((sel ? src1 : src0) + off);
            sb_u4 ov = sv;                                           // (padding bytes keep what the committed copy holds)
            if (A.mode == 1) ov = ev_load16(so);
            const uint32_t xw4[4] = {xv.x, xv.y, xv.z, xv.w}, sw4[4] = {sv.x, sv.y, sv.z, sv.w};
            uint32_t ow4[4] = {ov.x, ov.y, ov.z, ov.w};
            int r = r0;                                              // (fb + i) mod S
#ifndef SB_HOSTSIM
#pragma unroll
#endif
            for (int half = 0; half < 2; half++) {
                double prod = 1.0;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                for (int w2 = 0; w2 < 2; w2++) {
                    const int wi = 2 * half + w2;
                    sbp::Philox4 ph; ph.w[0] = ph.w[1] = ph.w[2] = ph.w[3] = 0;
                    uint32_t neww = ow4[wi];
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                    for (int j = 0; j < 4; j++) {
                        const int fl = fb + 4 * wi + j;
                        if (A.mode == 0 && !(j & 1) && fl < fv)      // the pair's draw: (chain seed, step, SITE_GIBBS, o * ceil(F/2) + f / 2)
                            ph = sbp::philox4x32_10((uint32_t)((size_t)o * hF + ((f0 + fl) >> 1)), (uint32_t)SITE_GIBBS, st_lo, st_hi, k0, k1);
                        const int x = (int)((xw4[wi] >> (8 * j)) & 255u), c_old = (int)((sw4[wi] >> (8 * j)) & 255u);
                        const int rr = r;
                        r = r + 1 == S ? 0 : r + 1;
                        if (fl >= fv) continue;
                        if (x == SB_NA_U8) { neww |= 255u << (8 * j); continue; }
                        int xr_ = x + rr; xr_ = xr_ >= S ? xr_ - S : xr_;
                        double q[UC], tot = 0.0;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                        for (int u = 0; u < UC; u++) {
                            q[u] = 0.0;
                            if (u < C) { const double l = gb[u] >= 0 ? lik[gb[u] + fl * S + xr_] : 0.0; q[u] = l * wrow[fl * Cp + u]; tot += q[u]; }
                        }
                        // one division per observation: q = raw * (1 / total); the inverse-CDF test u < cum / last is taken
                        // as u * last < cum (the same decision except within an ulp of a boundary)
                        const double inv = 1.0 / tot;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                        for (int u = 0; u < UC; u++) if (u < C) q[u] *= inv;
                        int pick;
                        if (A.mode == 0) {
                            const double uu = (j & 1) ? sbp::u01_double(ph.w[2], ph.w[3]) : sbp::u01_double(ph.w[0], ph.w[1]);
                            double last = 0.0, cum = 0.0;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                            for (int u = 0; u < UC; u++) if (u < C) last += q[u];
                            const double ul = uu * last;
                            pick = 0; bool got = false;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                            for (int u = 0; u < UC; u++) if (u < C) { cum += q[u]; if (!got && ul < cum) { pick = u; got = true; } }
                            neww = (neww & ~(255u << (8 * j))) | ((uint32_t)pick << (8 * j));
                            changed += pick != c_old ? 1u : 0u;
                        } else pick = A.mode == 1 ? (int)((ow4[wi] >> (8 * j)) & 255u) : c_old;
                        double qp = 0.0;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                        for (int u = 0; u < UC; u++) if (u == pick) qp = q[u];
                        prod *= qp;
                    }
                    ow4[wi] = neww;
                }
                acc += log(prod);
            }
            if (A.mode == 0) { sb_u4 nv; nv.x = ow4[0]; nv.y = ow4[1]; nv.z = ow4[2]; nv.w = ow4[3]; ev_store16(so, nv); }
        }
        for (int d = lpr >> 1; d > 0; d >>= 1) acc += sb_shfl_f64(acc, sb_lane() ^ d);
        if (live && fg == 0) opart[o] = acc;
    }
    if (A.mode == 0) {
        for (int d = 16; d > 0; d >>= 1) changed += (unsigned)sb_shfl_i32((int)changed, sb_lane() ^ d);
        if (sb_lane() == 0 && changed) sb_atomic_add_u64(A.n_changed + s_idx, (unsigned long long)changed);
    }
}

SB_KERNEL sb_gibbs_sweep_kernel(const SB_GRID_CONSTANT SbGibbsArgs A) {
    if (A.P.C == 3) ev_gibbs_body<3, true>(A); else if (A.P.C <= 4) ev_gibbs_body<4>(A); else ev_gibbs_body<8>(A);
}

// commits the rows a draw wrote: osel[o] ^= 1 for every object of every state
SB_KERNEL sb_flip_osel_kernel(const SbShared *states, int n_states, int N) {
    const long long total = (long long)n_states * N;
    for (long long i = (long long)sb_cta() * sb_nthr() + sb_tid(); i < total; i += (long long)sb_ncta() * sb_nthr()) {
        const SbShared st = states[i / N];
        st.osel[i % N] ^= 1;
    }
}

// -----------------------------------------------------------------------------------------
// Effect draws at log time: conditional_effect_sample (sbayes/sampling/conditionals.py:125-149), called by
// ParametersCSVLogger._write_sample for the cluster effect and every confounding effect of a logged sample
// (sbayes/sampling/loggers.py:160-224): p[g, f, :] ~ Dirichlet(counts[g, f, :] + alpha[g, f, :]) over the
// applicable states.  The reference first rebuilds the counts as a [groups, objects, features, states]
// boolean contraction (conditionals.py:68-86) and then loops over (group, feature) in Python; here the
// counts are the chain's committed statistics and one thread serves one (chain, group, feature).
// Gamma draws: Marsaglia-Tsang (2000), shapes below one boosted by U^(1/shape); Philox addressed by
// (chain seed, step, SITE_EFFECT, attempt sub-stream, (g*F + f)*S + s) -- oracle/effects.py draws the same.
// -----------------------------------------------------------------------------------------
#define SB_SITE_EFFECT 14
struct SbEffectArgs {
    SbProblem P;
    const SbShared *states;
    const SbSlot *slots;
    const int32_t *chains;       // [n] chain (slot) indices
    int n;
    unsigned long long step;
    double *out;                 // [n, G_tot, F, S]
};

SB_DEV double ev_effect_gamma(uint32_t k0, uint32_t k1, uint32_t st_lo, uint32_t st_hi, double shape, uint32_t idx) {
    const bool boost = shape < 1.0;
    const double a = boost ? shape + 1.0 : shape;
    const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (uint32_t t = 0;; t++) {
        const sbp::Philox4 p = sbp::philox4x32_10(idx, (uint32_t)SB_SITE_EFFECT | ((2u * t) << 8), st_lo, st_hi, k0, k1);
        const sbp::Philox4 q = sbp::philox4x32_10(idx, (uint32_t)SB_SITE_EFFECT | ((2u * t + 1u) << 8), st_lo, st_hi, k0, k1);
        const double u1 = sbp::u01_double(p.w[0], p.w[1]), u2 = sbp::u01_double(p.w[2], p.w[3]);
        const double u3 = sbp::u01_double(q.w[0], q.w[1]), u4 = sbp::u01_double(q.w[2], q.w[3]);
        const double xn = sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * 3.141592653589793 * u2);
        double v = 1.0 + c * xn;
        if (v <= 0.0) continue;
        v = v * v * v;
        if (log(1.0 - u3) < 0.5 * xn * xn + d - d * v + d * log(v)) {
            double g = d * v;
            if (boost) g *= exp(log(1.0 - u4) / shape);
            return g;
        }
    }
}

SB_KERNEL sb_effect_sample_kernel(const SB_GRID_CONSTANT SbEffectArgs A) {
    const SbProblem &P = A.P;
    const int F = P.F, S = P.S;
    const long long per = (long long)P.G_tot * F, total = per * A.n;
    for (long long it = (long long)sb_cta() * sb_nthr() + sb_tid(); it < total; it += (long long)sb_ncta() * sb_nthr()) {
        const int ci = (int)(it / per);
        const long long gf = it - (long long)ci * per;
        const int g = (int)(gf / F), f = (int)(gf - (long long)g * F);
        const SbSlot sl = A.slots[A.chains[ci]];
        const SbShared st = A.states[sl.state];
        const uint32_t k0 = (uint32_t)(sl.cfg.seed & 0xFFFFFFFFull), k1 = (uint32_t)(sl.cfg.seed >> 32);
        const uint32_t st_lo = (uint32_t)(A.step & 0xFFFFFFFFull), st_hi = (uint32_t)(A.step >> 32);
        const double *al = P.alpha + ((size_t)ev_arow(P, g) * F + f) * S;
        const int32_t *cr = st.counts + (size_t)gf * S;
        double *out = A.out + (size_t)it * S;
        double tot = 0.0; int n_app = 0;
        for (int s = 0; s < S; s++) {
            const double a = sb_ldg(al + s);
            double v = 0.0;
            if (a > 0.0) { v = ev_effect_gamma(k0, k1, st_lo, st_hi, a + (double)cr[s], (uint32_t)(gf * S + s)); n_app++; }
            out[s] = v; tot += v;
        }
        for (int s = 0; s < S; s++) {
            const bool app = sb_ldg(al + s) > 0.0;
            // (all draws underflowed -- concentrations far below one: fall back to the uniform vector)
            out[s] = tot > 0.0 ? out[s] / tot : (app ? 1.0 / (double)n_app : 0.0);
        }
    }
}

// launch geometry of the histogram pass: the widest feature tile whose bins fit `budget` bytes of
// shared memory, tiles balanced, and enough object blocks to fill the machine
static inline int sb_eval_make_plan(const SbProblem &P, int n_states, int n_sm, size_t budget, SbEvalPlan *pl) {
    const size_t bins_per_f = (size_t)P.G_tot * P.S + (size_t)P.n_pat * 2 * P.C;
    const size_t fixed = (size_t)P.K * 4 + 64 + ((size_t)P.G_tot + (size_t)P.n_pat * 2) * 4;    // + row padding, trash word
    if (budget < fixed + 16 * bins_per_f * 4) return -1;
    int ft_max = (int)((budget - fixed) / (bins_per_f * 4)) / 16 * 16;
    int n_ftiles = (P.ldx + ft_max - 1) / ft_max;
    int ft = ((P.ldx + n_ftiles - 1) / n_ftiles + 15) / 16 * 16;
    n_ftiles = (P.ldx + ft - 1) / ft;
    // object blocks per (state, tile): every CTA pays for zeroing and flushing its bins, so the
    // split that minimises  waves * (observations per CTA + 4 * bins per CTA)  is taken
    // (waves = CTAs over the resident slots: at most 3 CTAs per SM by registers)
    const size_t smem = fixed + (size_t)ft * bins_per_f * 4;
    long long per_sm = (long long)((227 * 1024) / (smem + 1024)); if (per_sm > 3) per_sm = 3; if (per_sm < 1) per_sm = 1;
    const long long slots = per_sm * n_sm, base = (long long)n_states * n_ftiles;
    long long max_blk = P.N / 128; if (max_blk < 1) max_blk = 1;         // at least 4 warps with an object group
    if (max_blk > 4096) max_blk = 4096;
    long long n_oblk = 1; double best = -1.0;
    for (long long n = 1; n <= max_blk; n++) {
        const long long waves = (base * n + slots - 1) / slots;
        const double cost = (double)waves * ((double)((P.N + n - 1) / n) * ft + 4.0 * (double)ft * bins_per_f);
        if (best < 0.0 || cost < best) { best = cost; n_oblk = n; }
    }
    int rows = (int)((P.N + n_oblk - 1) / n_oblk);
    n_oblk = (P.N + rows - 1) / rows;
    pl->ft = ft; pl->n_ftiles = n_ftiles; pl->n_oblk = (int)n_oblk; pl->rows_per_blk = rows;
    long long items = (long long)P.G_tot * P.F + (long long)P.n_pat * 2 * P.F + P.F;
    long long parts = (items + 511) / 512;               // two rows per thread: the terms are lgamma-latency bound
    pl->n_parts = (int)(parts < 1 ? 1 : (parts > SB_EVAL_MAXPARTS ? SB_EVAL_MAXPARTS : parts));
    pl->smem = (unsigned)smem;
    return 0;
}
