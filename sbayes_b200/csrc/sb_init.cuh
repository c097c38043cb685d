// sb_init.cuh -- the EM-style cluster initializer on the device
// (SbayesInitializer.generate_clusters_em, sbayes/sampling/initializers.py:93-169).
//
// A continuous relaxation of the sBayes model: z[g, o] is a soft assignment of object o to group g
// (clusters first, then every confounder's groups; a confounder group is only available to its members).
// One EM step of the reference is two contractions with the one-hot feature tensor X[o, f, s]:
//
//   counts[g, f, s] = sum_o z[g, o] X[o, f, s]                einsum("ij,jkl->ikl")          (:128)
//   p               = normalize(counts + 0.5 * applicable)                                  (:131)
//   ll[g, o]        = sum_f log sum_s p[g, f, s] X~[o, f, s]   einsum("ikl,jkl->ijk"), log   (:135-138)
//                     (X~ = X with every state set for a missing value: the term is log 1 = 0)
//   z[:, o]         = softmax_g( ll[g, o] / T_i  masked to the available groups ),  T_i = (n / (1 + i))^3   (:151-155)
//
// X is one-hot, so the first contraction is a segmented sum (the row x[o, f] selects the ONE state that
// receives z[g, o]) and the second a gather (p[g, f, x[o, f]]): evaluated as such they cost G*N*F
// multiply-adds each instead of the G*N*F*S of the dense einsum -- at config 5 (81 x 200 000 x 1000 x 8)
// 1.6e10 against 1.3e11.  Both kernels read the packed bytes x[o, f]; accumulation is fp64 (the reference
// is float32 with numpy's pairwise sums; parity tolerance in tests/test_init_em.py).
//
//   sb_em_counts_kernel   CTA = (feature tile of 32, object block, group block): the CTA stages its z rows, a warp
//                         walks the objects of the block, lane = feature; partial counts per object block are
//                         summed in block order by sb_em_tables_kernel (deterministic)
//   sb_em_tables_kernel   sums the partials, adds the prior counts, normalises, takes the log -> logp[g, f, s] f32
//   sb_em_assign_kernel   warp per object: ll[g] for all groups (lanes over features, shuffle reduction),
//                         masked tempered softmax over the groups -> z[:, o]
#pragma once
#include "sb_platform.cuh"
#include "sb_types.h"

#define SB_EM_BLOCK 256
#define SB_EM_FT 32          // features per counts CTA (one per lane)
#define SB_EM_GB 4           // groups per counts CTA

struct SbEmArgs {
    SbProblem P;
    float *z;                // [G_tot, N] current soft assignment (updated in place by the assign kernel)
    double *partial;         // [n_oblk, G_tot, F, S] partial counts
    float *logp;             // [G_tot, F, S]
    int n_oblk, rows_per_blk;
    double inv_temperature;
};

// availability of group row g for object o (clusters: every object; confounder groups: their members)
SB_DEV bool em_available(const SbProblem &P, int g, int o) {
    if (g < P.K) return true;
    int c = 1;
    while (c < P.C && g >= P.goff[c + 1]) c++;
    return sb_ldg(P.conf_gid + (size_t)(c - 1) * P.N + o) == g - P.goff[c];
}

SB_KERNEL sb_em_counts_kernel(const SB_GRID_CONSTANT SbEmArgs A) {
    const SbProblem &P = A.P;
    const int F = P.F, S = P.S, N = P.N, G = P.G_tot;
    const int n_ft = (F + SB_EM_FT - 1) / SB_EM_FT, n_gb = (G + SB_EM_GB - 1) / SB_EM_GB;
    int w = sb_cta();
    const int ft = w % n_ft; w /= n_ft;
    const int gb = w % n_gb; w /= n_gb;
    const int ob = w;
    const int f = ft * SB_EM_FT + sb_lane();
    const int g0 = gb * SB_EM_GB, ng = g0 + SB_EM_GB <= G ? SB_EM_GB : G - g0;
    const int lo = ob * A.rows_per_blk, hi = lo + A.rows_per_blk < N ? lo + A.rows_per_blk : N;
    // bins of this warp: [SB_EM_GB][S][32 lanes] doubles in shared memory (a lane's own column: no conflicts)
    double *bins = (double *)sb_smem() + (size_t)sb_warp() * SB_EM_GB * S * 32;
    for (int i = sb_lane(); i < SB_EM_GB * S * 32; i += 32) bins[i] = 0.0;
    sb_warp_sync();
    for (int o = lo + sb_warp(); o < hi; o += sb_nwarps()) {
        const int x = f < F ? (int)sb_ldg(P.x + (size_t)o * P.ldx + f) : SB_NA_U8;
        if (x == SB_NA_U8) continue;
        for (int q = 0; q < ng; q++) bins[(q * S + x) * 32 + sb_lane()] += (double)A.z[(size_t)(g0 + q) * N + o];
    }
    sb_cta_sync();
    // the warps' bins are added in warp order; thread = (group, state, feature lane)
    const int nw = sb_nwarps();
    for (int i = sb_tid(); i < ng * S * 32; i += sb_nthr()) {
        const int lane = i & 31, qs = i >> 5, q = qs / S, s = qs % S, ff = ft * SB_EM_FT + lane;
        if (ff >= F) continue;
        double v = 0.0;
        for (int ww = 0; ww < nw; ww++) v += ((double *)sb_smem())[(size_t)ww * SB_EM_GB * S * 32 + (size_t)(q * S + s) * 32 + lane];
        A.partial[(((size_t)ob * G + g0 + q) * F + ff) * S + s] = v;
    }
}

SB_KERNEL sb_em_tables_kernel(const SB_GRID_CONSTANT SbEmArgs A) {
    const SbProblem &P = A.P;
    const int F = P.F, S = P.S, G = P.G_tot;
    const size_t gf = (size_t)sb_cta() * sb_nthr() + sb_tid();
    if (gf >= (size_t)G * F) return;
    const int f = (int)(gf % F);
    double tot = 0.0;
    for (int s = 0; s < S; s++) {
        double v = 0.0;
        for (int b = 0; b < A.n_oblk; b++) v += A.partial[((size_t)b * G * F + gf) * S + s];
        // prior_counts = 0.5 * applicable states (initializers.py:117); float32 like the reference's state_counts + prior
        const double appl = P.use_mask ? (double)((sb_ldg(P.applm + f) >> s) & 1ull) : sb_ldg(P.unif + (size_t)f * S + s);
        v = (double)(float)v + 0.5 * appl;
        A.partial[gf * S + s] = v;           // block 0's slot is reused for the sums
        tot += v;
    }
    for (int s = 0; s < S; s++) {
        const double p = A.partial[gf * S + s] / tot;
        A.logp[gf * S + s] = p > 0.0 ? (float)log((double)(float)p) : -INFINITY;
    }
}

SB_KERNEL sb_em_assign_kernel(const SB_GRID_CONSTANT SbEmArgs A) {
    const SbProblem &P = A.P;
    const int F = P.F, S = P.S, N = P.N, G = P.G_tot;
    double *ll = (double *)sb_smem() + (size_t)sb_warp() * G;        // [G] per warp
    const int wpb = sb_nwarps();
    for (int o = sb_cta() * wpb + sb_warp(); o < N; o += sb_ncta() * wpb) {
        const uint8_t *xr = P.x + (size_t)o * P.ldx;
        for (int g = 0; g < G; g++) {
            double acc = 0.0;
            if (em_available(P, g, o)) {
                const float *lp = A.logp + (size_t)g * F * S;
                for (int f = sb_lane(); f < F; f += 32) {
                    const int x = (int)sb_ldg(xr + f);
                    if (x != SB_NA_U8) acc += (double)lp[(size_t)f * S + x];
                }
            }
            for (int d = 16; d > 0; d >>= 1) acc += sb_shfl_f64(acc, sb_lane() ^ d);
            if (sb_lane() == 0) ll[g] = acc;
        }
        sb_warp_sync();
        // masked tempered softmax over the groups (scipy.special.softmax of lh = ll / T, -inf where unavailable)
        double mx = -1e300;
        for (int g = sb_lane(); g < G; g += 32) if (em_available(P, g, o)) { const double v = (double)(float)((double)(float)ll[g] * A.inv_temperature); mx = v > mx ? v : mx; }
        for (int d = 16; d > 0; d >>= 1) { const double t = sb_shfl_f64(mx, sb_lane() ^ d); mx = t > mx ? t : mx; }
        double sum = 0.0;
        for (int g = sb_lane(); g < G; g += 32) if (em_available(P, g, o)) sum += exp((double)(float)((double)(float)ll[g] * A.inv_temperature) - mx);
        for (int d = 16; d > 0; d >>= 1) sum += sb_shfl_f64(sum, sb_lane() ^ d);
        for (int g = sb_lane(); g < G; g += 32)
            A.z[(size_t)g * N + o] = em_available(P, g, o) ? (float)(exp((double)(float)((double)(float)ll[g] * A.inv_temperature) - mx) / sum) : 0.0f;
        sb_warp_sync();
    }
}
