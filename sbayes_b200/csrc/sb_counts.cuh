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
#include "sb_plans.h"
#include "sb_eval.cuh"

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
#endif

// =========================================================================================
// The same statistics as ONE int8 tensor-core contraction over objects (tcgen05, sm_100a):
//
//   D[m, n] = sum_o  A[m, o] * B[n, o]
//     m = (feature, component c, state s)   A[m, o] = [src[o,f] == c and x[o,f] == s]   one-hot per observation
//     n = stacked group row g               B[g, o] = [o is a member of group g]
//     n = G_tot + (has-pattern, in-cluster) B[.., o] = [o has that pattern]
//   counts[g, f, s] = D[(f, comp(g), s), g]        H[pz, f, c] = sum_s D[(f, c, s), G_tot + pz]
//
// Both operands are *generated* in shared memory from the packed bytes (1 byte per observation
// and array in HBM): the A stage is zeroed with 16-byte stores and every observation writes one
// byte; the K dimension (objects) is consumed 64 objects per stage, double buffered against the
// MMAs through tcgen05.commit -> mbarrier.  The accumulators (int32) live in TMEM.
//
// Shared-memory operand layout, MN-major without swizzle (canonical UMMA "interleave" form):
//   byte(mn, k) = (mn / 16) * SBO + (mn % 16) + (k % 8) * 16 + (k / 8) * LBO,  LBO = 128, SBO = KC * 16
// i.e. 128-byte core matrices of 8 objects x 16 rows.
// =========================================================================================
#define SB_MMA_KC 64          /* objects per stage (two K = 32 MMAs per M tile) */
#define SB_MMA_BLOCK 512


static inline int sb_mma_make_plan(const SbProblem &P, int n_states, int n_sm, SbMmaPlan *mp) {
    const int cs = P.C * P.S, ncol = P.G_tot + P.n_pat * 2, npad = (ncol + 15) / 16 * 16;
    if (npad > 256 || cs > 128) return -1;
    int best = 0;
    for (int ftw = 16; ftw <= 128 && ftw <= P.ldx; ftw += 16) {
        const int tiles = (ftw * cs + 127) / 128;
        const size_t smem = 2 * ((size_t)tiles * 128 * SB_MMA_KC + (size_t)npad * SB_MMA_KC) + 256;
        if (tiles * npad <= 512 && smem <= 200 * 1024) best = ftw;
    }
    if (!best) return -1;
    int n_ftiles = (P.ldx + best - 1) / best;
    int ftw = ((P.ldx / 16 + n_ftiles - 1) / n_ftiles) * 16;          // balanced tiles
    n_ftiles = (P.ldx + ftw - 1) / ftw;
    const int tiles = (ftw * cs + 127) / 128;
    long long want = 2LL * n_sm, base = (long long)n_states * n_ftiles;
    long long n_oblk = (want + base - 1) / base;
    long long max_blk = (P.N + 4 * SB_MMA_KC - 1) / (4 * SB_MMA_KC);   // at least 4 stages per CTA
    if (n_oblk > max_blk) n_oblk = max_blk;
    if (n_oblk < 1) n_oblk = 1;
    int rows = (int)((P.N + n_oblk - 1) / n_oblk);
    rows = (rows + SB_MMA_KC - 1) / SB_MMA_KC * SB_MMA_KC;
    n_oblk = (P.N + rows - 1) / rows;
    int cols = 32;
    while (cols < tiles * npad) cols *= 2;
    mp->cs = cs; mp->npad = npad; mp->ftw = ftw; mp->n_ftiles = n_ftiles; mp->n_oblk = (int)n_oblk; mp->rows_per_blk = rows;
    mp->tiles = tiles; mp->tmem_cols = cols;
    mp->a_bytes = (unsigned)tiles * 128 * SB_MMA_KC; mp->b_bytes = (unsigned)npad * SB_MMA_KC;
    mp->smem = 2 * (mp->a_bytes + mp->b_bytes) + 256;
    return 0;
}

#ifndef SB_HOSTSIM
namespace sbmma {
SB_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
SB_DEV void mbar_init(uint32_t a, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(n) : "memory"); }
SB_DEV void mbar_wait(uint32_t a, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
SB_DEV void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
SB_DEV void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
SB_DEV void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SB_DEV void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
SB_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
SB_DEV void commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
// shared-memory matrix descriptor, no swizzle, sm_100 version field = 1
SB_DEV uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// D[tmem] (+)= A[smem] * B[smem], unsigned 8-bit operands, int32 accumulate, M = 128
SB_DEV void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
SB_DEV void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
}  // namespace sbmma

// args: the evaluation argument block (sb_eval.cuh) -- statistics go to st.counts / st.H / st.sizes
//
// Warp-specialised: 16 producer warps generate the operand stages, one extra warp issues the
// MMAs.  full[b] (one arrival per producer warp) hands a stage to the issuer, tcgen05.commit on
// empty[b] hands the buffer back; producers only meet each other at a named barrier between
// zeroing and filling a stage.
#define SB_MMA_PRODUCERS 512
#undef SB_MMA_BLOCK
#define SB_MMA_BLOCK (SB_MMA_PRODUCERS + 32)
__global__ void __launch_bounds__(SB_MMA_BLOCK, 1) sb_counts_mma_kernel(const __grid_constant__ SbEvalArgs A, const __grid_constant__ SbMmaPlan mp) {
    using namespace sbmma;
    const SbProblem &P = A.P;
    const int F = P.F, S = P.S, C = P.C, nconf = P.n_conf, KC = SB_MMA_KC;
    int w = blockIdx.x;
    const int oblk = w % mp.n_oblk; w /= mp.n_oblk;
    const int tile = w % mp.n_ftiles; w /= mp.n_ftiles;
    const SbShared st = A.states[w];
    const int f0 = tile * mp.ftw;
    const int ftw = P.ldx - f0 < mp.ftw ? P.ldx - f0 : mp.ftw;
    const int cpr = ftw >> 4;
    const int lo = oblk * mp.rows_per_blk, hi = min(lo + mp.rows_per_blk, P.N);
    const int n_stages = (hi - lo + KC - 1) / KC;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t LBO = 128, SBO = KC * 16;

    extern __shared__ __align__(1024) char sm[];
    uint64_t *full = (uint64_t *)sm, *empty = full + 2;
    uint32_t *tmem_slot = (uint32_t *)(sm + 64);
    char *bufA[2] = {sm + 256, sm + 256 + mp.a_bytes};
    char *bufB[2] = {sm + 256 + 2 * mp.a_bytes, sm + 256 + 2 * mp.a_bytes + mp.b_bytes};
    if (tid == 0) {
        mbar_init(smem_u32(full), SB_MMA_PRODUCERS / 32); mbar_init(smem_u32(full + 1), SB_MMA_PRODUCERS / 32);
        mbar_init(smem_u32(empty), 1); mbar_init(smem_u32(empty + 1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)mp.tmem_cols);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == SB_MMA_PRODUCERS / 32) {
        // ---------------------------------------------------------------- MMA issuer
        if (lane == 0) {
            // instruction descriptor: D = s32, A/B = unsigned 8 bit, both MN-major, N = npad, M = 128
            const uint32_t idesc = (2u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(mp.npad >> 3) << 17) | (8u << 24);
            const uint64_t a_d[2] = {smem_desc(smem_u32(bufA[0]), LBO, SBO), smem_desc(smem_u32(bufA[1]), LBO, SBO)};
            const uint64_t b_d[2] = {smem_desc(smem_u32(bufB[0]), LBO, SBO), smem_desc(smem_u32(bufB[1]), LBO, SBO)};
            for (int stg = 0; stg < n_stages; stg++) {
                const int b = stg & 1;
                mbar_wait(smem_u32(full + b), (uint32_t)((stg >> 1) & 1));
                fence_after();
                for (int t = 0; t < mp.tiles; t++)
#pragma unroll
                    for (int ks = 0; ks < KC / 32; ks++)       // descriptor start address is in 16-byte units
                        mma_i8(tmem + (uint32_t)(t * mp.npad), a_d[b] + (uint64_t)((t * 8 * SBO + ks * 4 * LBO) >> 4),
                               b_d[b] + (uint64_t)((ks * 4 * LBO) >> 4), idesc, (stg > 0 || ks > 0) ? 1u : 0u);
                commit(smem_u32(empty + b));
            }
        }
    } else {
        // ---------------------------------------------------------------- producers
        // this thread's 16 observations of a stage: object ol, byte columns [fc, fc + 16)
        const int ol = tid % KC, ci = tid / KC;
        const bool has_item = ci < cpr;
        const int fc = ci << 4;
        struct Item { uint4 xv, sv; int k_cl, combo; };
        auto load_stage = [&](int stg) {
            Item it;
            it.xv = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu); it.sv = it.xv; it.k_cl = -1; it.combo = -1;
            const int o = lo + stg * KC + ol;
            if (stg < n_stages && o < hi) {
                if (has_item) {
                    const size_t off = (size_t)o * P.ldx + f0 + fc;
                    it.xv = *(const uint4 *)(P.x + off);
                    it.sv = *(const uint4 *)((st.osel[o] ? st.src[1] : st.src[0]) + off);
                }
                if (ci == 0) { it.k_cl = st.cluster_of[o]; it.combo = __ldg(P.combo_id + o); }
            }
            return it;
        };
        int err = 0;
        Item cur = load_stage(0), nxt = load_stage(1);
        for (int stg = 0; stg < n_stages; stg++) {
            const int b = stg & 1;
            if (stg >= 2) mbar_wait(smem_u32(empty + b), (uint32_t)(((stg >> 1) - 1) & 1));   // MMAs of stage stg-2 have read buffer b
            {
                const uint4 z = make_uint4(0, 0, 0, 0);
                uint4 *a4 = (uint4 *)bufA[b]; uint4 *b4 = (uint4 *)bufB[b];
                for (int i = tid; i < (int)(mp.a_bytes >> 4); i += SB_MMA_PRODUCERS) a4[i] = z;
                for (int i = tid; i < (int)(mp.b_bytes >> 4); i += SB_MMA_PRODUCERS) b4[i] = z;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(SB_MMA_PRODUCERS) : "memory");
            // one byte per observation into A
            if (has_item) {
                char *base = bufA[b] + (ol & 7) * 16 + (ol >> 3) * LBO;
                const uint32_t xs[4] = {cur.xv.x, cur.xv.y, cur.xv.z, cur.xv.w}, ss[4] = {cur.sv.x, cur.sv.y, cur.sv.z, cur.sv.w};
#pragma unroll
                for (int q = 0; q < 4; q++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int x = (xs[q] >> (8 * j)) & 255, c = (ss[q] >> (8 * j)) & 255;
                        if (x == SB_NA_U8) continue;
                        if (c >= C) { err = 2; continue; }
                        const int m = (fc + q * 4 + j) * mp.cs + c * S + x;
                        base[(m >> 4) * SBO + (m & 15)] = 1;
                    }
            }
            // membership columns of the object into B
            if (ci == 0 && cur.combo >= 0) {
                char *base = bufB[b] + (ol & 7) * 16 + (ol >> 3) * LBO;
                const int k_cl = cur.k_cl;
                if (k_cl >= 0) base[(k_cl >> 4) * SBO + (k_cl & 15)] = 1;
                const int32_t *rows = P.combo_rows + (size_t)cur.combo * nconf;
                for (int c = 0; c < nconf; c++) { const int g = __ldg(rows + c); if (g >= 0) base[(g >> 4) * SBO + (g & 15)] = 1; }
                const int n = P.G_tot + __ldg(P.combo_pat + cur.combo) * 2 + (k_cl >= 0 ? 1 : 0);
                base[(n >> 4) * SBO + (n & 15)] = 1;
                if (tile == 0 && k_cl >= 0) atomicAdd(st.sizes + k_cl, 1);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(full + b)) : "memory");
            cur = nxt;
            nxt = load_stage(stg + 2);
        }
        if (err) *A.err = err;
        // wait for the last commit of each buffer, then drain TMEM
        if (n_stages > 0) {
            const int last = n_stages - 1;
            mbar_wait(smem_u32(empty + (last & 1)), (uint32_t)((last >> 1) & 1));
            if (n_stages > 1) { const int l2 = last - 1; mbar_wait(smem_u32(empty + (l2 & 1)), (uint32_t)((l2 >> 1) & 1)); }
            fence_after();
            const int fv = min(F - f0, ftw);                          // real features of this CTA
            const int quarter = warp & 3;
            for (int t = warp >> 2; t < mp.tiles; t += SB_MMA_PRODUCERS / 128) {
                const int m = t * 128 + quarter * 32 + lane;
                const int fl = m / mp.cs, code = m % mp.cs, c = code / S, s = code % S;
                const bool live = fl < fv;
                const int g_lo = P.goff[c < C ? c : 0], g_hi = P.goff[(c < C ? c : 0) + 1];
                for (int n0 = 0; n0 < mp.npad; n0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * mp.npad + n0), v);
                    if (!live) continue;
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const int n = n0 + i, val = (int)v[i];
                        if (val == 0) continue;
                        if (n < P.G_tot) { if (n >= g_lo && n < g_hi) atomicAdd(st.counts + ((size_t)n * F + f0 + fl) * S + s, val); }
                        else if (n < P.G_tot + P.n_pat * 2) atomicAdd(st.H + ((size_t)(n - P.G_tot) * F + f0 + fl) * C + c, val);
                    }
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)mp.tmem_cols);
}
#endif
