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
//   sb_em_counts_kernel   CTA = (feature tile of 32, object block, group block of 4): lane = feature, a warp walks the objects of
//                         the block for the dense rows (clusters, the universal group) and the member lists of the sparse
//                         rows (z is zero outside a confounder group), eight objects in flight; partial counts per object
//                         block are summed in block order by sb_em_tables_kernel (deterministic)
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
    // rows of the block: cluster rows and confounder groups that hold a large share of the objects (the universal group) are
    // DENSE -- every object of the block is visited; the other confounder groups are SPARSE -- z is zero outside the group, so
    // only the group's member list (P.cg_ptr / P.cg_idx) is walked
    bool dense[SB_EM_GB]; bool any_dense = false;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
    for (int q = 0; q < SB_EM_GB; q++) {
        const int g = g0 + q;
        dense[q] = q < ng && (g < P.K || 4LL * (sb_ldg(P.cg_ptr + g - P.K + 1) - sb_ldg(P.cg_ptr + g - P.K)) > (long long)N);
        any_dense = any_dense || dense[q];
    }
    // dense rows: a warp takes the objects lo + warp + k * n_warps, eight at a time: the z values (one broadcast load per
    // (object, group)) and the state bytes of all eight are requested before the first add; an object whose z is zero in every
    // dense row of the block costs no state load and no add (adding 0.0 would not change a bin)
    constexpr int U = 8;
    const int nwp = sb_nwarps();
    if (any_dense)
    for (int o0 = lo + sb_warp(); o0 < hi; o0 += nwp * U) {
        float zz[U][SB_EM_GB]; int xx[U];
#ifndef SB_HOSTSIM
#pragma unroll
#endif
        for (int u = 0; u < U; u++) {
            const int o = o0 + u * nwp;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
            for (int q = 0; q < SB_EM_GB; q++) zz[u][q] = (o < hi && dense[q]) ? A.z[(size_t)(g0 + q) * N + o] : 0.0f;
        }
#ifndef SB_HOSTSIM
#pragma unroll
#endif
        for (int u = 0; u < U; u++) {
            const int o = o0 + u * nwp;
            bool any = false;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
            for (int q = 0; q < SB_EM_GB; q++) any = any || zz[u][q] != 0.0f;
            xx[u] = (any && f < F) ? (int)sb_ldg(P.x + (size_t)o * P.ldx + f) : SB_NA_U8;
        }
#ifndef SB_HOSTSIM
#pragma unroll
#endif
        for (int u = 0; u < U; u++) {
            const int x = xx[u];
            if (x == SB_NA_U8) continue;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
            for (int q = 0; q < SB_EM_GB; q++) if (zz[u][q] != 0.0f) bins[(q * S + x) * 32 + sb_lane()] += (double)zz[u][q];
        }
    }
    // sparse rows: the members of the group that fall into this CTA's object block, eight at a time
    for (int q = 0; q < ng; q++) {
        if (dense[q]) continue;
        const int g = g0 + q;
        const int p0 = sb_ldg(P.cg_ptr + g - P.K), p1 = sb_ldg(P.cg_ptr + g - P.K + 1);
        for (int i0 = p0 + sb_warp(); i0 < p1; i0 += nwp * U) {
            float zz[U]; int xx[U];
#ifndef SB_HOSTSIM
#pragma unroll
#endif
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * nwp;
                const int o = i < p1 ? sb_ldg(P.cg_idx + i) : -1;
                const bool in = o >= lo && o < hi;
                zz[u] = in ? A.z[(size_t)g * N + o] : 0.0f;
                xx[u] = (in && f < F) ? (int)sb_ldg(P.x + (size_t)o * P.ldx + f) : SB_NA_U8;
            }
#ifndef SB_HOSTSIM
#pragma unroll
#endif
            for (int u = 0; u < U; u++) if (xx[u] != SB_NA_U8 && zz[u] != 0.0f) bins[(q * S + xx[u]) * 32 + sb_lane()] += (double)zz[u];
        }
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
        // the rows available to this object: every cluster and, per confounder, the row of its group (em_available as one
        // look-up per confounder instead of one per (group, object))
        int crow[8];
#ifndef SB_HOSTSIM
#pragma unroll
#endif
        for (int c = 1; c < 8; c++) {
            crow[c] = -1;
            if (c < P.C) { const int gid = sb_ldg(P.conf_gid + (size_t)(c - 1) * N + o); if (gid >= 0) crow[c] = P.goff[c] + gid; }
        }
        const int n_rows = P.K + P.C - 1;
        for (int r = 0; r < n_rows; r++) {
            int g = r;
            if (r >= P.K) {
                g = -1;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
                for (int c = 1; c < 8; c++) if (c == r - P.K + 1) g = crow[c];
                if (g < 0) continue;                                 // (warp-uniform)
            }
            double acc = 0.0;
            const float *lp = A.logp + (size_t)g * F * S;
            for (int f = sb_lane(); f < F; f += 32) {
                const int x = (int)sb_ldg(xr + f);
                if (x != SB_NA_U8) acc += (double)lp[(size_t)f * S + x];
            }
            for (int d = 16; d > 0; d >>= 1) acc += sb_shfl_f64(acc, sb_lane() ^ d);
            if (sb_lane() == 0) ll[g] = acc;
        }
        sb_warp_sync();
        // masked tempered softmax over the groups (scipy.special.softmax of lh = ll / T, -inf where unavailable)
        auto avail = [&](int g) {
            bool a = g < P.K;
#ifndef SB_HOSTSIM
#pragma unroll
#endif
            for (int c = 1; c < 8; c++) a = a || crow[c] == g;
            return a;
        };
        double mx = -1e300;
        for (int g = sb_lane(); g < G; g += 32) if (avail(g)) { const double v = (double)(float)((double)(float)ll[g] * A.inv_temperature); mx = v > mx ? v : mx; }
        for (int d = 16; d > 0; d >>= 1) { const double t = sb_shfl_f64(mx, sb_lane() ^ d); mx = t > mx ? t : mx; }
        double sum = 0.0;
        for (int g = sb_lane(); g < G; g += 32) if (avail(g)) sum += exp((double)(float)((double)(float)ll[g] * A.inv_temperature) - mx);
        for (int d = 16; d > 0; d >>= 1) sum += sb_shfl_f64(sum, sb_lane() ^ d);
        for (int g = sb_lane(); g < G; g += 32)
            A.z[(size_t)g * N + o] = avail(g) ? (float)(exp((double)(float)((double)(float)ll[g] * A.inv_temperature) - mx) / sum) : 0.0f;
        sb_warp_sync();
    }
}
