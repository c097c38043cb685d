// sb_core.cuh -- device code of the sBayes MCMC step: team primitives, draws, tables,
// the ten scheduled operators, Metropolis-Hastings accept/commit.
//
// Execution model.  One chain is advanced by a *team* of `team` CTAs of one persistent
// kernel (a thread-block cluster when team <= 16); every thread of the team executes the
// same step function.  All random draws are counter-based (Philox4x32-10 addressed by chain
// seed, step, site, sub-stream, index), so every thread can derive any draw locally and no
// draw is ever broadcast.
//
// Work that scales with objects x features (the proposal passes over candidate objects)
// is split across the team by contiguous object blocks; its sums cross CTAs through
// `team_sum` / `team_pick` / `team_compact`, which end in a team barrier and leave the
// *same* value in every thread, so all control flow is uniform across the team.
// Everything group-indexed (counts, per-(group, feature) totals, reciprocals and
// log-likelihood terms, weight tables, the source histogram) is kept as a private replica
// per CTA (shared memory where it fits): every CTA applies the same proposal to its replica
// *in place*, redundantly, and undoes it on a reject, which removes all barriers and L2
// round trips from the resampling / table-update / accept phases.  Object-indexed state in
// HBM is written by *all* CTAs with identical values, so each CTA always observes its own
// writes; barriers are only needed to keep a fast CTA's commit from overtaking a slow CTA's
// reads (one barrier between the last read and the commit).
//
// Arithmetic.  Sums of logs over features are accumulated as products with a separately
// tracked binary exponent (`Prod`): one fp64 multiply per observation instead of one `log`.
// The passes over many objects gather from per-combo tables A0 / A1 (`build_tables`): a
// combo is a distinct tuple of confounder groups, so all objects of a combo see the same
// confounder mixture sum_c p_c[f, x] * w_c; per observation that is one byte load, two
// 8-byte table loads and two multiplies.  Log-likelihood terms are updated by
// lgamma(z + d) - lgamma(z) = sum log (for the few counts a proposal changes) instead of
// re-evaluating lgamma over whole rows.
//
// What follows which reference code is cited function by function; the algorithmic
// specification is SURVEY.md Appendix A, the executable one is oracle/operators.py.
// This file is compiled once per unroll width of the mixture-component loops: the includer
// defines SB_UC (3, 4 or 8: the largest number of components the instantiation handles; with SB_EXACT_C the exact number) and
// SB_NS (the namespace the instantiation lives in); see sb_api.cu.
#include "sb_platform.cuh"
#include "sb_types.h"
#include "sb_philox.cuh"
#ifndef SB_UC
#error "define SB_UC and SB_NS before including sb_core.cuh"
#endif
namespace SB_NS {

#undef SB_FOR_C
#undef SB_FOR_C1
#undef SB_CC
// number of mixture components: a compile-time constant in an instantiation that serves problems with EXACTLY SB_UC components
// (SB_EXACT_C; sb_uc3.cu: cluster + universal + one more confounder, every BASELINE configuration) -- the guards of the
// unrolled component loops then fold away -- and the problem's own number otherwise
#ifdef SB_EXACT_C
#define SB_CC(P) SB_UC
#else
#define SB_CC(P) ((P).C)
#endif
#undef PH
#undef SB_NPH
#ifndef SB_EPS32
#define SB_EPS32 1.1920928955078125e-07   /* np.finfo(np.float32).eps, sbayes/util.py:34 */
#define SB_SMOOTH 1e-6                    /* additive smoothing, sbayes/sampling/operators.py:978 */
#define SB_W_STAY 0.15                    /* operators.py:133 */
#define SB_GEO_SCALER 2.0                 /* operators.py:1390 */
#define SB_SRC_MIN 10                     /* operators.py:440 */
#define SB_SRC_MAX_RANDOM 20              /* operators.py:185 */
#define SB_SRC_MAX_GROUPS 30              /* operators.py:194 */
#define SB_LN2 0.6931471805599453
#endif
// loops over the mixture components are unrolled to the compile-time maximum and guarded, so
// that the per-component arrays stay in registers (a runtime trip count puts them in local memory)
#ifdef SB_HOSTSIM
#define SB_FOR_C(c, C) for (int c = 0; c < (C); c++)
#define SB_FOR_C1(c, C) for (int c = 1; c < (C); c++)
#else
#define SB_FOR_C(c, C) _Pragma("unroll") for (int c = 0; c < SB_UC; c++) if (c < (C))
#define SB_FOR_C1(c, C) _Pragma("unroll") for (int c = 1; c < SB_UC; c++) if (c < (C))
#endif

#ifndef SB_INF
#ifdef SB_HOSTSIM
#define SB_INF (__builtin_inf())
#else
#define SB_INF (__longlong_as_double(0x7ff0000000000000LL))
#endif
#endif

// phase profiling (debug builds with -DSB_PHASE_PROF only): cycles of thread 0 of the rank-0
// CTA between consecutive probes, per operator
#if defined(SB_PHASE_PROF) && !defined(SB_HOSTSIM)
#define SB_NPH 48
__device__ long long g_phase[SB_N_OPS][SB_NPH];
__device__ long long g_phase_n[SB_N_OPS];
#define PH(R, id) do { if (sb_tid() == 0 && (R).rank == 0) { long long t_ = clock64(); g_phase[(R).cur_op][id] += t_ - (R).pt; (R).pt = t_; } } while (0)
#else
#define PH(R, id) do { } while (0)
#endif

using sbp::Philox4; using sbp::philox4x32_10; using sbp::u01_double; using sbp::u01_float;     // sb_philox.cuh

struct Draws {
    uint32_t k0, k1;
    unsigned long long step;
    const double *tv;     // tape values (null => Philox)
    const int32_t *toff;  // offsets of the current step's sites
    const int32_t *tlen;
    int wide_n;           // n_avail recorded for the wide operator of this step
};
SB_DEV Philox4 draw_words(const Draws &D, int site, int sub, uint32_t idx) {
    uint32_t c1 = (uint32_t)(site & 0xFF) | ((uint32_t)(sub & 0xFFFFFF) << 8);
    return philox4x32_10(idx, c1, (uint32_t)(D.step & 0xFFFFFFFFull), (uint32_t)(D.step >> 32), D.k0, D.k1);
}
SB_DEV double draw_uniform(const Draws &D, int site, uint32_t idx = 0, int sub = 0) {
    if (D.tv) return D.toff[site] >= 0 ? D.tv[D.toff[site] + idx] : 0.0;   // absent: the reference did not draw here
    Philox4 p = draw_words(D, site, sub, idx);
    return u01_double(p.w[0], p.w[1]);
}
SB_DEV double draw_uniform_f32(const Draws &D, int site, uint32_t idx, int sub) {
    if (D.tv) {
        if (site == SITE_WIDE) {   // one row of n_avail uniforms per redraw attempt (oracle/draws.py: TapeDraws)
            int n = D.wide_n > 0 ? D.wide_n : D.tlen[site];
            int n_att = D.tlen[site] / n; if (n_att < 1) n_att = 1;
            int a = sub < n_att - 1 ? sub : n_att - 1;
            return D.tv[D.toff[site] + a * n + idx];
        }
        return D.tv[D.toff[site] + idx];
    }
    Philox4 p = draw_words(D, site, sub, idx);
    return u01_float(p.w[0]);
}
SB_DEV int draw_index(const Draws &D, int site, uint32_t idx, int n) {
    if (D.tv) return (int)D.tv[D.toff[site] + idx];
    int i = (int)(draw_uniform(D, site, idx) * (double)n);
    return i < n - 1 ? i : n - 1;
}
SB_DEV void draw_pair(const Draws &D, int site, int n, int &i, int &j) {
    if (D.tv) { i = (int)D.tv[D.toff[site]]; j = (int)D.tv[D.toff[site] + 1]; return; }
    i = draw_index(D, site, 0, n);
    j = draw_index(D, site, 1, n - 1);
    if (j >= i) j++;
}
// Marsaglia-Tsang gamma, shape >= 1; attempt t uses sub-streams 4t + 2*which + {0,1} (oracle/draws.py:_gamma)
SB_DEV double draw_gamma(const Draws &D, double shape, uint32_t f, int which) {
    double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    for (int t = 0;; t++) {
        Philox4 a = draw_words(D, SITE_BETA, 4 * t + 2 * which, f);
        double u1 = u01_double(a.w[0], a.w[1]), u2 = u01_double(a.w[2], a.w[3]);
        Philox4 b = draw_words(D, SITE_BETA, 4 * t + 2 * which + 1, f);
        double u3 = u01_double(b.w[0], b.w[1]);
        double xn = sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * 3.141592653589793 * u2);
        double v = 1.0 + c * xn;
        if (v <= 0.0) continue;
        v = v * v * v;
        if (log(1.0 - u3) < 0.5 * xn * xn + d - d * v + d * log(v)) return d * v;
    }
}
SB_DEV double draw_beta(const Draws &D, double a, double b, uint32_t f) {
    if (D.tv) return D.tv[D.toff[SITE_BETA] + f];
    double gx = draw_gamma(D, a, f, 0), gy = draw_gamma(D, b, f, 1);
    return gx / (gx + gy);
}

// -----------------------------------------------------------------------------------------
// products with a separately tracked binary exponent: sum_f log a_f == log(prod a_f)
// -----------------------------------------------------------------------------------------
struct Prod { double m; int e; };
SB_DEV Prod prod_one() { Prod p; p.m = 1.0; p.e = 0; return p; }
SB_DEV void prod_renorm(Prod &p) {      // move the exponent of m into e (zero / inf / nan / denormals are left alone)
    long long b = sb_d2ll(p.m);
    int ex = (int)((b >> 52) & 0x7ff);
    if (ex != 0 && ex != 0x7ff) {
        p.e += ex - 1023;
        p.m = sb_ll2d((b & ~(0x7ffLL << 52)) | (1023LL << 52));
    }
}
SB_DEV double prod_log(const Prod &p) { return log(p.m) + (double)p.e * SB_LN2; }

// -----------------------------------------------------------------------------------------
// per-CTA replica of the group-indexed state, and the run context
// -----------------------------------------------------------------------------------------
struct Rep {
    int32_t *counts;                                  // [G_tot, F, S] current counts (candidate while a proposal is open)
    int32_t *nn;                                      // [G_tot, F]    sum_s counts
    int32_t *dcnt;                                    // [G_tot, F, S] net change of the open proposal (zero between steps)
    double *rP, *rT;                                  // [G_tot, F]    1 / (sum alpha + n),  1 / (sum tempered base + n / T)
    double *bsT;                                      // [G_tot, F]    sum_s (u + (alpha - u) / Tp)
    double *llgf;                                     // [G_tot, F]    Dirichlet-categorical term of (group, feature)
    double *bk_llgf, *bk_rP, *bk_rT;                  // backups of the dirty rows (undo)
    int32_t *H[2];                                    // [n_pat*2, F, C]; H[hcur] committed, the other one candidate
    double *w1, *wp, *wq, *lw;                        // [n_pat*2, F, C] weight tables
    SbD2 *tab;                                        // [n_combos, F, S] pass tables, the two of them interleaved (build_tables)
    double *rs;                                       // [n_combos, F, S+1] ratio rows of the tiled pair pass (HBM slice)
    const uint8_t *xblk;                              // the CTA's own block of observation rows, or null
    int xlo, xhi;
    double *lh, *bk_lh;                               // [G_tot] per-group log-likelihood
    double *geo_agg, *geo_val, *c_geo_agg, *c_geo_val;// [K]
    float *W;                                         // [F, C]
    int32_t *sizes, *ssz;                             // [K] committed / candidate cluster sizes
    uint8_t *dirty;                                   // [G_tot]
    int *dlist;                                       // [G_tot] dirty rows of the open proposal
    uint8_t *selbuf;                                  // [N] which source copy becomes the committed one, per changed object
};

enum { CHG_NONE = 0, CHG_ONE, CHG_WIDE };

// What is uniform over the CTA and constant while a chain is being advanced lives once per CTA
// in shared memory (RunC); only the few mutable scalars of the step stay in registers (Run).
// (Keeping all of it per thread spills ~2 KB per thread to local memory, which the shared
// memory carve-out leaves no L1 for.)
struct RunC {
    SbShared st;
    Rep rp;
    // problem constants and per-object state staged in shared memory where they fit (else the
    // pointers address the global copies; mcl / mcombo / mosel are null then)
    const double *alpha, *asum;  // [A_rows, F, S], [A_rows, F]
    const uint64_t *applm;       // [F] bit s set <=> state s applicable
    const int32_t *combo_rows, *combo_pat;
    int16_t *mcl;                // [N] replica of the cluster membership
    uint16_t *mcombo;            // [N]
    uint8_t *mosel;              // [N] replica of osel
    SbSlot *slot;
    double opw[SB_N_OPS];
    unsigned long long *bar;
    double *tred;      // [2][tsize][SB_REDW] team reduction scratch
    double *sred;      // smem [2][32 * SB_REDW]
    int *sint;         // smem [96]
    int *slist;        // smem [32] small list of changed objects
    int *smem_mem;     // smem member list (geo)
    double *smem_best; // smem Prim keys (geo)
    int smem_cap;
    double *tile;      // smem tile buffer of the tiled passes (null: the whole pass table is resident, rp.tab)
    int tile_cap;      // its size in doubles
    // objects sharded over GPUs (SbLaunch: xworld > 1)
    double *tredp[SB_MAX_RANKS];               // this team's reduction buffer in every rank's memory
    unsigned long long *xflagp[SB_MAX_RANKS];  // this team's flag row in every rank's memory
    unsigned long long *lrel;                  // release word of the local barrier phase
    int32_t *lcl;                              // rank-local replica of cluster_of of the current chain (or null)
    uint8_t *losel;                            // ... of osel
    int32_t *lstatp[SB_MAX_RANKS], *gstatp[SB_MAX_RANKS];   // statistics blocks of the current chain in every rank's memory
    double *valp[2][SB_MAX_RANKS];             // the chain's val arrays in every rank's memory: an object's value lives on the rank
                                               // that owns the object's block (written there, picked from there)
    int32_t *llist;                            // this rank's own copy of list_a: CTA-private compaction lists stay local
    int *xabort;
};

struct Run {
    const SbProblem *P;
    RunC *c;
    Draws D;
    double T, Tp, invT, invTp, p_grow;
    int sfp;
    int propose_only;            // operator API: commit every proposal the operator does not reject itself (no MH test)
    int tsize, rank, par, vpar, spar, use_cluster, hcur;
    int lsize, lrank, xrank, xworld;     // CTAs of the team on this GPU, index among them; GPU index / count (object sharding)
    unsigned long long xbase;
    unsigned long long epoch;
    int nd;            // number of dirty rows of the open proposal
    // candidate membership view
    int chg_mode, chg_o, chg_new, chg_k;
    // committed posterior terms: kept in registers by every thread (uniform across the team)
    double ll, pr_size, pr_geo, pr_w, pr_src;
    long long i_step;
    int err;
    long long pt; int cur_op;     // phase profiling (debug builds)
};

// layout of the always-in-shared-memory part (host and device agree through this function)
struct FixedSmem { unsigned runc, sred, best, lh, bk_lh, geo, sint, slist, sizes, ssz, mem, W, dlist, dirty, end; };
SB_HD FixedSmem sb_fixed_smem(int G, int K, int F, int C, int cap, unsigned base) {
    FixedSmem o; unsigned b = (base + 15u) & ~15u;
    o.runc = b; b += (unsigned)((sizeof(RunC) + 15) & ~(size_t)15);
    o.sred = b; b += 2 * 32 * SB_REDW * 8;
    o.best = b; b += (unsigned)cap * 8;
    o.lh = b; b += (unsigned)G * 8;
    o.bk_lh = b; b += (unsigned)G * 8;
    o.geo = b; b += (unsigned)K * 8 * 4;
    o.sint = b; b += 96 * 4;
    o.slist = b; b += 32 * 4;
    o.sizes = b; b += (unsigned)K * 4;
    o.ssz = b; b += (unsigned)K * 4;
    o.mem = b; b += (unsigned)cap * 4;
    o.W = b; b += (unsigned)F * C * 4;
    o.dlist = b; b += (unsigned)G * 4;
    o.dirty = b; b += (unsigned)G;
    o.end = (b + 15u) & ~15u;
    return o;
}

// Barrier of a team that spans several GPUs (thread 0 of every CTA).  Phase 1: L2 ticket among the CTAs of
// this GPU.  Phase 2 (the last CTA to arrive): one release store into every other rank's flag row (peer memory
// over NVLink), then an acquire spin on the own row until every rank has arrived.  Phase 3: release word for
// the CTAs of this GPU.  System-scope fences on both sides order the team's peer-memory traffic around the
// barrier.  Flags only ever grow (launch base + barrier count), so nothing has to be reset between launches.
// A wait that exceeds ~4 s sets the abort word: every later barrier falls through and the host reports an error
// instead of the GPUs hanging on each other.
SB_DEV void xteam_wait(Run &R) {
    if (sb_ld_volatile_i32(R.c->xabort)) return;
    const long long t_lim = 8000000000LL;
    sb_threadfence_system();                     // this CTA's stores (peer memory included) before its arrival
    const unsigned long long arrived = sb_atomic_add_u64(R.c->bar, 1ull) + 1ull;
    const unsigned long long ev = R.xbase + R.epoch;
    if (arrived == R.epoch * (unsigned long long)R.lsize) {
        // the fence makes everything this thread has observed (the other CTAs' arrivals and, through their fences, their
        // stores) visible system-wide before the flags; the polls are plain volatile loads, ordered by the fence after them
        sb_threadfence_system();
        for (int p = 0; p < R.xworld; p++)
            if (p != R.xrank) sb_st_volatile_u64(R.c->xflagp[p] + R.xrank, ev);
        for (int p = 0; p < R.xworld; p++) {
            if (p == R.xrank) continue;
            const long long t0 = sb_clock();
            while (sb_ld_volatile_u64(R.c->xflagp[R.xrank] + p) < ev) {
                sb_spin();
                if (sb_clock() - t0 > t_lim || sb_ld_volatile_i32(R.c->xabort)) { *R.c->xabort = 1; break; }
            }
        }
        sb_threadfence_system();
        sb_st_volatile_u64(R.c->lrel, ev);
    } else {
        const long long t0 = sb_clock();
        while (sb_ld_volatile_u64(R.c->lrel) < ev) {
            sb_spin();
            if (sb_clock() - t0 > 2 * t_lim || sb_ld_volatile_i32(R.c->xabort)) { *R.c->xabort = 1; break; }
        }
    }
    sb_threadfence_system();                     // what the other ranks stored before their arrival, before this CTA reads on
}

// a CTA's slot of the team reduction buffer (thread 0): with the objects sharded over GPUs the buffer is
// replicated and the slot is stored into every rank's copy, so that everybody reads locally after the barrier
SB_DEV void tred_put(Run &R, int par, const double *v, int K) {
    const size_t off = ((size_t)par * R.tsize + R.rank) * SB_REDW;
    if (R.xworld > 1) {
        for (int p = 0; p < R.xworld; p++)
            for (int k = 0; k < K; k++) R.c->tredp[p][off + k] = v[k];
    } else
        for (int k = 0; k < K; k++) R.c->tred[off + k] = v[k];
}

SB_DEV void team_sync(Run &R) {
#ifndef SB_HOSTSIM
    if (R.use_cluster && R.tsize > 1) {      // hardware barrier over all threads of the cluster
        asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
        return;
    }
#endif
    sb_cta_sync();
    if (R.tsize > 1) {
        R.epoch++;
        if (sb_tid() == 0) {
            if (R.xworld > 1) xteam_wait(R);
            else {
                sb_threadfence();
                sb_atomic_add_u64(R.c->bar, 1ull);
                unsigned long long target = R.epoch * (unsigned long long)R.tsize;
                while (sb_ld_acquire_u64(R.c->bar) < target) sb_spin();
                sb_threadfence();
            }
        }
        sb_cta_sync();
        if (R.xworld > 1 && sb_ld_volatile_i32(R.c->xabort)) R.err = 5;
    }
}

// butterfly all-reduce: every lane ends with the bitwise identical result (a + b == b + a)
SB_DEV double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += sb_shfl_f64(v, sb_lane() ^ off);
    return v;
}
SB_DEV double warp_prod(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v *= sb_shfl_f64(v, sb_lane() ^ off);
    return v;
}
SB_DEV double warp_min(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmin(v, sb_shfl_f64(v, sb_lane() ^ off));
    return v;
}
SB_DEV Prod warp_prod_reduce(Prod p) {     // mantissas are in [1, 2): 32 of them cannot overflow
    prod_renorm(p);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        p.m *= sb_shfl_f64(p.m, sb_lane() ^ off);
        p.e += sb_shfl_i32(p.e, sb_lane() ^ off);
    }
    prod_renorm(p);
    return p;
}

// CTA-wide sum of K values per thread; fixed tree, result in every thread.  One barrier:
// the per-warp partials alternate between two buffers.
template <int K, bool PROD = false>
SB_DEV void block_sum(Run &R, double (&v)[K]) {
    double *buf = R.c->sred + R.spar * (32 * SB_REDW);
    R.spar ^= 1;
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = PROD ? warp_prod(v[k]) : warp_sum(v[k]);
    if (sb_lane() == 0)
        for (int k = 0; k < K; k++) buf[sb_warp() * SB_REDW + k] = v[k];
    sb_cta_sync();
    const int nw = sb_nwarps();
#pragma unroll
    for (int k = 0; k < K; k++) {
        double t = sb_lane() < nw ? buf[sb_lane() * SB_REDW + k] : (PROD ? 1.0 : 0.0);
        v[k] = PROD ? warp_prod(t) : warp_sum(t);
    }
}

// Large teams (objects sharded over GPUs: more than a thousand slots): if every warp walked all slots of the reduction
// buffer, 148 CTAs x 16 warps would read it from L2 for each reduction (tens of MB and a dependent chain of L2 round trips
// per step); above this team size the warps of a CTA split the slots and combine their partials through shared memory.
// The combination order depends only on the team size and the CTA size, so the bits are the same in every CTA and in
// a single-GPU run with the same team.
#ifdef SB_HOSTSIM
SB_DEV int split_min() { const char *e = getenv("SB_HOSTSIM_SPLIT_MIN"); return e ? atoi(e) : 64; }
#else
SB_DEV int split_min() { return 64; }
#endif
SB_DEV void split_range(const Run &R, int &r_lo, int &r_hi, int &per) {
    per = (R.tsize + sb_nwarps() - 1) / sb_nwarps();
    r_lo = sb_warp() * per; if (r_lo > R.tsize) r_lo = R.tsize;
    r_hi = r_lo + per; if (r_hi > R.tsize) r_hi = R.tsize;
}

// team-wide sum (one team barrier); result identical in every thread of the team.
template <int K, bool PROD = false>
SB_DEV void team_sum(Run &R, double (&v)[K]) {
    block_sum<K, PROD>(R, v);
    if (R.tsize == 1) return;
    double *buf = R.c->tred + (size_t)R.par * R.tsize * SB_REDW;
    if (sb_tid() == 0) tred_put(R, R.par, v, K);
    R.par ^= 1;
    team_sync(R);
    if (R.tsize > split_min()) {
        int r_lo, r_hi, per; split_range(R, r_lo, r_hi, per);
#pragma unroll
        for (int k = 0; k < K; k++) {
            double t = PROD ? 1.0 : 0.0;
            for (int r = r_lo + sb_lane(); r < r_hi; r += 32) { if (PROD) t *= buf[r * SB_REDW + k]; else t += buf[r * SB_REDW + k]; }
            v[k] = t;
        }
        block_sum<K, PROD>(R, v);
        return;
    }
#pragma unroll
    for (int k = 0; k < K; k++) {
        double t = PROD ? 1.0 : 0.0;
        for (int r = sb_lane(); r < R.tsize; r += 32) { if (PROD) t *= buf[r * SB_REDW + k]; else t += buf[r * SB_REDW + k]; }
        v[k] = PROD ? warp_prod(t) : warp_sum(t);
    }
}

SB_DEV void team_range(const Run &R, int n, int &lo, int &hi) {
    int B = (n + R.tsize - 1) / R.tsize;
    lo = R.rank * B; if (lo > n) lo = n;
    hi = lo + B; if (hi > n) hi = n;
}

// Inverse-CDF draw over val[0..n): first i with cumsum(val)[i] / total > u (numpy's
// Generator.choice(p=...), oracle/draws.py:cdf_pick).  val must be fully written by its
// owners; contains one team barrier.  Returns the index (or -1 if total is not > 0).
// extra[0..1] are summed over the team on the way (free ride on the same barrier).
SB_DEV int team_pick(Run &R, const double *val, int n, double u, double &total, double *extra = nullptr, double *const *val_peers = nullptr) {
    // val_peers (objects sharded over GPUs): val of block r lives on rank r / lsize; `val` is this rank's copy
    int lo, hi; team_range(R, n, lo, hi);
    const int nthr = sb_nthr(), tid = sb_tid();
    // every thread owns a contiguous chunk of the CTA's block
    int ch = (hi - lo + nthr - 1) / nthr;
    int a = lo + tid * ch, b = a + ch; if (a > hi) a = hi; if (b > hi) b = hi;
    double cs = 0.0;
    for (int i = a; i < b; i++) cs += val[i];
    double s[3] = {cs, extra ? extra[0] : 0.0, extra ? extra[1] : 0.0};
    block_sum<3>(R, s);
    int rb = 0; double tloc;
    if (R.tsize > 1) {
        double *buf = R.c->tred + (size_t)R.par * R.tsize * SB_REDW;
        if (tid == 0) tred_put(R, R.par, s, 3);
        R.par ^= 1;
        team_sync(R);
        const int lane = sb_lane();
        int s_lo = 0, s_hi = R.tsize;                 // slots scanned for the block that holds the target
        double carry = 0.0, target;
        if (R.tsize > split_min()) {
            // the warps split the slots: range totals through shared memory, totals and the range that holds the target
            // from them in every thread, then every warp scans that one range
            int r_lo, r_hi, per; split_range(R, r_lo, r_hi, per);
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
            for (int r = r_lo + lane; r < r_hi; r += 32) { t0 += buf[r * SB_REDW]; t1 += buf[r * SB_REDW + 1]; t2 += buf[r * SB_REDW + 2]; }
            t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
            double *wt = R.c->sred + R.spar * (32 * SB_REDW);
            R.spar ^= 1;
            if (lane == 0) { wt[sb_warp() * SB_REDW] = t0; wt[sb_warp() * SB_REDW + 1] = t1; wt[sb_warp() * SB_REDW + 2] = t2; }
            sb_cta_sync();
            const int nw = sb_nwarps();
            total = 0.0; double e0 = 0.0, e1 = 0.0;
            for (int v = 0; v < nw; v++) { total += wt[v * SB_REDW]; e0 += wt[v * SB_REDW + 1]; e1 += wt[v * SB_REDW + 2]; }
            if (extra) { extra[0] = e0; extra[1] = e1; }
            if (!(total > 0.0)) return -1;
            target = u * total;
            int vs = -1; double acc = 0.0;
            for (int v = 0; v < nw; v++) {
                const double sv = wt[v * SB_REDW];
                if (sv > 0.0) { vs = v; carry = acc; if (acc + sv > target) break; }
                acc += sv;
            }
            if (vs < 0) vs = 0;
            s_lo = vs * per; s_hi = s_lo + per < R.tsize ? s_lo + per : R.tsize;
        } else {
            // every warp on its own, lanes over the team's slots: totals by lane-strided sums + butterfly
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
            for (int r = lane; r < R.tsize; r += 32) { t0 += buf[r * SB_REDW]; t1 += buf[r * SB_REDW + 1]; t2 += buf[r * SB_REDW + 2]; }
            total = warp_sum(t0);
            const double e0 = warp_sum(t1), e1 = warp_sum(t2);
            if (extra) { extra[0] = e0; extra[1] = e1; }
            if (!(total > 0.0)) return -1;
            target = u * total;
        }
        // the block that holds the target by a chunked warp scan -- the same result in every warp of every CTA
        rb = -1; double prefix = 0.0;
        for (int r0 = s_lo; r0 < s_hi; r0 += 32) {
            const int r = r0 + lane;
            const double sr = r < s_hi ? buf[r * SB_REDW] : 0.0;
            double incl = sr;
            for (int off = 1; off < 32; off <<= 1) { const double t = sb_shfl_f64(incl, lane - off); if (lane >= off) incl += t; }
            const double excl = carry + (incl - sr);
            const unsigned pos = sb_ballot(sr > 0.0), hit = sb_ballot(sr > 0.0 && excl + sr > target);
            if (hit) { const int l = sb_ffs(hit) - 1; rb = r0 + l; prefix = sb_shfl_f64(excl, l); break; }
            if (pos) { const int l = 31 - sb_clz(pos); rb = r0 + l; prefix = sb_shfl_f64(excl, l); }     // last positive block so far
            carry += sb_shfl_f64(incl, 31);
        }
        tloc = target - prefix;
        if (rb != R.rank) {       // the block that holds the target: every CTA scans it (no second barrier)
            const int B = (n + R.tsize - 1) / R.tsize;
            lo = rb * B; hi = lo + B; if (hi > n) hi = n;
            if (val_peers) val = val_peers[rb / R.lsize];
            ch = (hi - lo + nthr - 1) / nthr;
            a = lo + tid * ch; b = a + ch; if (a > hi) a = hi; if (b > hi) b = hi;
            cs = 0.0;
            for (int i = a; i < b; i++) cs += val[i];
        }
    } else {
        total = s[0];
        if (extra) { extra[0] = s[1]; extra[1] = s[2]; }
        if (!(total > 0.0)) return -1;
        tloc = u * total;
    }
    // block-wide exclusive scan of the chunk sums: warp scan, warp totals through shared memory
    const int lane = sb_lane(), w = sb_warp(), nw = sb_nwarps();
    double incl = cs;
    for (int off = 1; off < 32; off <<= 1) {
        double t = sb_shfl_f64(incl, lane - off);
        if (lane >= off) incl += t;
    }
    double *wt = R.c->sred + R.spar * (32 * SB_REDW);
    R.spar ^= 1;
    if (lane == 31) wt[w] = incl;
    sb_cta_sync();
    double wpre = 0.0;
    for (int v = 0; v < w; v++) wpre += wt[v];
    const double excl = wpre + (incl - cs);
    // the first thread whose chunk carries the running sum past the target; if rounding lets none do so,
    // the last thread with a positive chunk
    const unsigned hit = sb_ballot(cs > 0.0 && excl + cs > tloc), pos = sb_ballot(cs > 0.0);
    if (lane == 0) { R.c->sint[w] = hit ? (w << 5) + sb_ffs(hit) - 1 : -1; R.c->sint[32 + w] = pos ? (w << 5) + 31 - sb_clz(pos) : -1; }
    sb_cta_sync();
    int sel_t = -1, last_t = -1;
    for (int v = 0; v < nw; v++) { const int h = R.c->sint[v], q = R.c->sint[32 + v]; if (sel_t < 0 && h >= 0) sel_t = h; if (q >= 0) last_t = q; }
    if (sel_t < 0) sel_t = last_t;
    if (tid == (sel_t < 0 ? 0 : sel_t)) {
        double run = excl; int found = -1, last = -1;
        for (int i = a; i < b; i++) {
            const double vi = val[i];
            if (vi > 0.0) { last = i; run += vi; if (run > tloc) { found = i; break; } }
        }
        R.c->sint[64] = found >= 0 ? found : last;
    }
    sb_cta_sync();
    const int o = R.c->sint[64];
    sb_cta_sync();
    return o;
}

// Ordered compaction: list[] receives, in increasing order, all i in [0, n) with pred(i).
// Two team barriers; returns the count.  pred must be pure.
template <class Pred>
SB_DEV int team_compact(Run &R, int n, int32_t *list, Pred pred) {
    int lo, hi; team_range(R, n, lo, hi);
    double c[1] = {0.0};
    for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) c[0] += pred(i) ? 1.0 : 0.0;
    block_sum<1>(R, c);
    double *buf = R.c->tred + (size_t)R.par * R.tsize * SB_REDW;
    if (sb_tid() == 0) tred_put(R, R.par, c, 1);
    R.par ^= 1;
    team_sync(R);
    int base = 0, total = 0;
    if (R.tsize > split_min()) {
        int r_lo, r_hi, per; split_range(R, r_lo, r_hi, per);
        double bt[2] = {0.0, 0.0};      // (integers: exact in any order)
        for (int r = r_lo + sb_lane(); r < r_hi; r += 32) { const double v = buf[r * SB_REDW]; if (r < R.rank) bt[0] += v; bt[1] += v; }
        block_sum<2>(R, bt);
        base = (int)bt[0]; total = (int)bt[1];
    } else {
        double tb = 0.0, tt = 0.0;      // (integers: exact in any order)
        for (int r = sb_lane(); r < R.tsize; r += 32) { const double v = buf[r * SB_REDW]; if (r < R.rank) tb += v; tt += v; }
        base = (int)warp_sum(tb); total = (int)warp_sum(tt);
    }
    int running = base, nw = sb_nwarps();
    for (int c0 = lo; c0 < hi; c0 += sb_nthr()) {
        int i = c0 + sb_tid();
        bool p = i < hi && pred(i);
        unsigned bal = sb_ballot(p);
        if (sb_lane() == 0) R.c->sint[sb_warp()] = sb_popc(bal);
        sb_cta_sync();
        int wpre = 0, tot = 0;
        for (int w = 0; w < nw; w++) { int v = R.c->sint[w]; if (w < sb_warp()) wpre += v; tot += v; }
        if (p) list[running + wpre + sb_popc(bal & ((1u << sb_lane()) - 1u))] = i;
        running += tot;
        sb_cta_sync();
    }
    team_sync(R);
    return total;
}

// CTA-local ordered compaction of [lo, hi): list[0..n) receives, in increasing order, the i with
// pred(i); other(i) is called for the rest.  Every warp owns a contiguous segment (two passes over
// the predicate, two CTA barriers).  Returns n in every thread.  pred must be pure.
template <class Pred, class Other>
SB_DEV int cta_compact(Run &R, int lo, int hi, int32_t *list, Pred pred, Other other) {
    const int nw = sb_nwarps(), w = sb_warp(), lane = sb_lane();
    const int seg = ((hi - lo + nw - 1) / nw + 31) & ~31;
    const int a = lo + w * seg, b = a + seg < hi ? a + seg : hi;
    int cnt = 0;
    for (int i0 = a; i0 < b; i0 += 32) {
        const int i = i0 + lane;
        const bool p = i < b && pred(i);
        if (i < b && !p) other(i);
        cnt += sb_popc(sb_ballot(p));
    }
    sb_cta_sync();                              // sint may still be read from an earlier use
    if (lane == 0) R.c->sint[w] = cnt;
    sb_cta_sync();
    int base = 0, total = 0;
    for (int v = 0; v < nw; v++) { const int c = R.c->sint[v]; if (v < w) base += c; total += c; }
    for (int i0 = a; i0 < b; i0 += 32) {
        const int i = i0 + lane;
        const bool p = i < b && pred(i);
        const unsigned bal = sb_ballot(p);
        if (p) list[base + sb_popc(bal & ((1u << lane) - 1u))] = i;
        base += sb_popc(bal);
    }
    sb_cta_sync();
    return total;
}

// Items [lo, hi) of the CTA, interleaved over its warps: every item is first processed by
// its whole warp (`warp_fn`, warp-uniform, may shuffle; returns an Acc to every lane), then
// lane i finishes item i of the batch on its own (`lane_fn`, no shuffles): the
// transcendental tail of 32 items runs lane-parallel instead of 32 times.
template <class Acc, class WarpFn, class LaneFn>
SB_DEV void batch_items(int lo, int hi, WarpFn warp_fn, LaneFn lane_fn) {
    const int gw = sb_warp(), nw = sb_nwarps(), lane = sb_lane();
    for (int i0 = lo + gw; i0 < hi; i0 += nw * 32) {
        Acc mine = Acc(); bool have = false;      // value-initialised: an indeterminate `mine` lets nvcc fold the lane select away
        for (int i = 0; i < 32; i++) {
            int it = i0 + nw * i;
            if (it >= hi) break;
            Acc a = warp_fn(it);
            if (lane == i) { mine = a; have = true; }
        }
        if (have) lane_fn(i0 + nw * lane, mine);
    }
}

// -----------------------------------------------------------------------------------------
// views: committed (0) or candidate (1) membership
// -----------------------------------------------------------------------------------------
SB_DEV int m_cl(const Run &R, int o) { return R.c->mcl ? (int)R.c->mcl[o] : (R.c->lcl ? R.c->lcl[o] : R.c->st.cluster_of[o]); }
SB_DEV int m_osel(const Run &R, int o) { return R.c->mosel ? (int)R.c->mosel[o] : (R.c->losel ? (int)R.c->losel[o] : (int)R.c->st.osel[o]); }
SB_DEV int m_combo(const Run &R, int o) { return R.c->mcombo ? (int)R.c->mcombo[o] : sb_ldg(R.P->combo_id + o); }
SB_DEV int m_pat(const Run &R, int o) { return R.c->combo_pat[m_combo(R, o)]; }
// proposal values of objects (val[o]): with the objects sharded over GPUs the value of object o lives in the memory
// of the rank that owns o's block of the team split -- written there by whoever computes it, summed and scanned
// there by the pick
SB_DEV double *val_slot(const Run &R, int b, int o) {
    if (R.xworld == 1) return R.c->st.val[b] + o;
    const int B = (R.P->N + R.tsize - 1) / R.tsize;
    return R.c->valp[b][(o / B) / R.lsize] + o;
}
SB_DEV double *val_mine(const Run &R, int b) { return R.xworld == 1 ? R.c->st.val[b] : R.c->valp[b][R.xrank]; }
SB_DEV int32_t *list_mine(const Run &R) { return R.c->llist ? R.c->llist : R.c->st.list_a; }

SB_DEV int cl_of(const Run &R, int o, int view) {
    int c = m_cl(R, o);
    if (!view || R.chg_mode == CHG_NONE) return c;
    if (R.chg_mode == CHG_ONE) return o == R.chg_o ? R.chg_new : c;
    if (c < 0 || c == R.chg_k) return R.c->st.flag_a[o] ? R.chg_k : -1;      // CHG_WIDE
    return c;
}
SB_DEV int conf_group_g(const SbProblem &P, int c, int o) {   // stacked group row of object o in component c >= 1, or -1
    int gid = sb_ldg(P.conf_gid + (size_t)(c - 1) * P.N + o);
    return gid < 0 ? -1 : P.goff[c] + gid;
}
SB_DEV int conf_group(const Run &R, int c, int o) { return R.c->combo_rows[(size_t)m_combo(R, o) * R.P->n_conf + c - 1]; }
SB_DEV const uint8_t *x_row(const Run &R, int o) {
    if (R.c->rp.xblk && o >= R.c->rp.xlo && o < R.c->rp.xhi) return R.c->rp.xblk + (size_t)(o - R.c->rp.xlo) * R.P->ldx;
    return R.P->x + (size_t)o * R.P->ldx;
}
// v^(1/T): exp(log v / T) is within a few ulp of pow and a third of its instructions
SB_DEV double pow_inv(double v, double inv, bool unit) { return unit ? v : (v > 0.0 ? exp(inv * log(v)) : 0.0); }

// heat_binary_probability (sbayes/util.py:1457-1469)
SB_DEV double heat(const Run &R, double p) {
    if (R.T == 1.0) return p / (p + (1.0 - p));
    double a = pow_inv(p, R.invT, false), b = pow_inv(1.0 - p, R.invT, false);
    return a / (a + b);
}
SB_DEV double sb_expit(double z) { return z >= 0 ? 1.0 / (1.0 + exp(-z)) : exp(z) / (1.0 + exp(z)); }
SB_DEV double log_expit(double z) {   // -logaddexp(0, -z)
    double m = -z > 0 ? -z : 0.0;
    return -(m + log(exp(0.0 - m) + exp(-z - m)));
}
// GeoPrior.get_probability_function (sbayes/model/prior.py:754-767)
SB_DEV double geo_prob_fn(const SbProblem &P, double x) {
    if (P.geo_pf == 0) return -x / P.geo_rate;
    return log_expit(-(x - P.geo_infl) / P.geo_rate) - log_expit(P.geo_infl / P.geo_rate);
}

// -----------------------------------------------------------------------------------------
// effect tables as functions of the counts (sbayes/sampling/conditionals.py:105-122,174-179):
//   plain    p[g,f,s]  = (alpha + n_s) / (sum alpha + n)
//   tempered pT[g,f,s] = (u + (alpha - u)/Tp + n_s/T) / (sum_s(u + (alpha - u)/Tp) + n/T)
// The reciprocals of the denominators are kept per (group, feature) in rP / rT.
// -----------------------------------------------------------------------------------------
// the K cluster rows share one concentration row when the host found them identical (dedup)
SB_DEV int arow(const SbProblem &P, int g) { return g < P.K ? (P.dedup ? 0 : g) : g - P.arow_off; }
SB_DEV double alpha_at(const Run &R, int g, int f, int s) { return R.c->alpha[((size_t)arow(*R.P, g) * R.P->F + f) * R.P->S + s]; }
SB_DEV double asum_at(const Run &R, int g, int f) { return R.c->asum[(size_t)arow(*R.P, g) * R.P->F + f]; }
SB_DEV double unif_at(const Run &R, int f, int s) {
    return R.P->use_mask ? (double)((R.c->applm[f] >> s) & 1ull) : sb_ldg(R.P->unif + (size_t)f * R.P->S + s);
}
SB_DEV double num_P(const Run &R, int g, int f, int s) {
    return alpha_at(R, g, f, s) + (double)R.c->rp.counts[((size_t)g * R.P->F + f) * R.P->S + s];
}
SB_DEV double num_T(const Run &R, int g, int f, int s) {
    double a = alpha_at(R, g, f, s), u = unif_at(R, f, s);
    return u + (a - u) * R.invTp + (double)R.c->rp.counts[((size_t)g * R.P->F + f) * R.P->S + s] * R.invT;
}
SB_DEV double entry_P(const Run &R, int g, int f, int x) { return num_P(R, g, f, x) * R.c->rp.rP[(size_t)g * R.P->F + f]; }
SB_DEV double entry_T(const Run &R, int g, int f, int x) { return num_T(R, g, f, x) * R.c->rp.rT[(size_t)g * R.P->F + f]; }
// tempered entry with a freshly computed denominator (rT is stale while a proposal is being applied)
SB_DEV double entry_T_fresh(const Run &R, int g, int f, int x) {
    size_t gf = (size_t)g * R.P->F + f;
    return num_T(R, g, f, x) / (R.c->rp.bsT[gf] + (double)R.c->rp.nn[gf] * R.invT);
}

// lgamma(z + d) - lgamma(z) for an integer step d (z > 0, z + d > 0)
SB_DEV double lgamma_diff(double z, int d) {
    if (d == 0) return 0.0;
    if (d > 8 || d < -8) return lgamma(z + (double)d) - lgamma(z);
    double p = 1.0;
    if (d > 0) { for (int i = 0; i < d; i++) p *= z + (double)i; return log(p); }
    for (int i = 1; i <= -d; i++) p *= z - (double)i;
    return -log(p);
}

// per-(group, feature) state of whole rows from the counts: totals, reciprocals, the
// Dirichlet-categorical term (sbayes/util.py:1373-1394; sbayes/model/likelihood.py:65-101).
// CTA-local; rows g_list[0..n_rows) (null: all rows).
SB_DEV void rows_full(Run &R, const int *g_list, int n_rows) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S;
    for (int idx = sb_tid(); idx < n_rows * F; idx += sb_nthr()) {
        const int g = g_list ? g_list[idx / F] : idx / F, f = idx % F;
        const size_t gf = (size_t)g * F + f, b = gf * S;
        int n = 0; double ser = 0.0, bs = 0.0;
        for (int s = 0; s < S; s++) {
            int cnt = R.c->rp.counts[b + s];
            double a = alpha_at(R, g, f, s), u = unif_at(R, f, s);
            n += cnt; bs += u + (a - u) * R.invTp;
            if (a > 0.0) ser += lgamma((double)cnt + a);
        }
        R.c->rp.nn[gf] = n; R.c->rp.bsT[gf] = bs;
        R.c->rp.rP[gf] = 1.0 / (asum_at(R, g, f) + (double)n);
        R.c->rp.rT[gf] = 1.0 / (bs + (double)n * R.invT);
        R.c->rp.llgf[gf] = sb_ldg(P.dcc + (size_t)arow(P, g) * F + f) - lgamma((double)n + asum_at(R, g, f)) + ser;
    }
    sb_cta_sync();
}
// lh[g] = sum_f llgf[g, f] for the listed rows (warp per row, fixed tree)
SB_DEV void rows_lh(Run &R, const int *g_list, int n_rows) {
    const int F = R.P->F;
    for (int i = sb_warp(); i < n_rows; i += sb_nwarps()) {
        const int g = g_list ? g_list[i] : i;
        double s = 0.0;
        for (int f = sb_lane(); f < F; f += 32) s += R.c->rp.llgf[(size_t)g * F + f];
        s = warp_sum(s);
        if (sb_lane() == 0) R.c->rp.lh[g] = s;
    }
    sb_cta_sync();
}
SB_DEV double ll_total(const Run &R) {          // every warp on its own; identical result everywhere
    double s = 0.0;
    for (int g = sb_lane(); g < R.P->G_tot; g += 32) s += R.c->rp.lh[g];
    return warp_sum(s);
}

// weight tables for every (confounder pattern p, in-a-cluster bit z), CTA-local:
//   w1 = has*W / sum(has*W)                 normalize_weights   (sbayes/model/likelihood.py:171-190)
//   wp = w1^(1/Tp)                          source Gibbs        (operators.py:568-572,830-836)
//   wq = has*W^(1/Tp) / sum(has*W^(1/Tp))   with/without        (operators.py:1075-1095)
//   lw = log w1                             SourcePrior         (sbayes/model/prior.py:596-611)
SB_DEV void weight_tables(Run &R) {
    const SbProblem &P = *R.P;
    const int F = P.F, C = SB_CC(P), rows = P.n_pat * 2 * F;
    const bool unitTp = R.Tp == 1.0;
    for (int i = sb_tid(); i < rows; i += sb_nthr()) {
        int f = i % F, pz = i / F, z = pz & 1, bits = P.pat_bits[pz >> 1];
        double t1 = 0.0, tp = 0.0, w64[SB_UC], wpw[SB_UC];
        SB_FOR_C(c, C) {
            double h = c == 0 ? (double)z : (double)((bits >> c) & 1);
            double w = (double)R.c->rp.W[f * C + c];
            w64[c] = h * w; wpw[c] = h * (unitTp ? w : pow(w, R.invTp));
            t1 += w64[c]; tp += wpw[c];
        }
        size_t o = (size_t)i * C;
        SB_FOR_C(c, C) {
            double a = t1 > 0.0 ? w64[c] / t1 : 0.0;
            R.c->rp.w1[o + c] = a;
            R.c->rp.wp[o + c] = unitTp ? a : pow(a, R.invTp);
            R.c->rp.wq[o + c] = tp > 0.0 ? wpw[c] / tp : 0.0;
            R.c->rp.lw[o + c] = log(a);
        }
    }
    sb_cta_sync();
}

// SourcePrior (sbayes/model/prior.py:568-611) from the (pattern, feature, component)
// histogram: sum_o sum_f log w~[o,f,src] = sum H * log w1.  Whole CTA; result in every thread.
SB_DEV double source_prior_from_hist(Run &R, const int32_t *H) {
    const SbProblem &P = *R.P;
    int n = P.n_pat * 2 * P.F * SB_CC(P);
    double a[1] = {0.0};
    for (int i = sb_tid(); i < n; i += sb_nthr()) {
        int h = H[i];
        if (h != 0) a[0] += (double)h * R.c->rp.lw[i];
    }
    block_sum<1>(R, a);
    return a[0];
}

// scipy.stats.dirichlet._logpdf of one weight row (sbayes/util.py:49, prior.py:553-558)
SB_DEV double weights_prior_row(const SbProblem &P, const float *Wrow, int f) {
    double t = 0.0;
    for (int c = 0; c < SB_CC(P); c++) {
        double a = sb_ldg(P.alpha_w + (size_t)f * SB_CC(P) + c);
        if (a != 1.0) t += (a - 1.0) * log((double)Wrow[c]);
    }
    return -sb_ldg(P.wlnB + f) + t;
}
SB_DEV double weights_prior_row_reg(const SbProblem &P, const float (&Wrow)[SB_UC], int f) {
    if (P.wprior_uniform) return 0.0 - sb_ldg(P.wlnB + f);
    double t = 0.0;
    SB_FOR_C(c, SB_CC(P)) {
        double a = sb_ldg(P.alpha_w + (size_t)f * SB_CC(P) + c);
        if (a != 1.0) t += (a - 1.0) * log((double)Wrow[c]);
    }
    return -sb_ldg(P.wlnB + f) + t;
}
SB_DEV double weights_prior_total(Run &R) {   // WeightsPrior.__call__ (prior.py:522-551); whole CTA
    const SbProblem &P = *R.P;
    if (P.wprior_uniform) return 0.0;
    double a[1] = {0.0};
    for (int f = sb_tid(); f < P.F; f += sb_nthr()) a[0] += weights_prior_row(P, R.c->rp.W + (size_t)f * SB_CC(P), f);
    block_sum<1>(R, a);
    return a[0];
}

// log multinomial coefficient and ClusterSizePrior (sbayes/util.py:1125-1173, prior.py:647-681)
SB_DEV double size_prior_of(const SbProblem &P, const int *sizes) {
    if (P.size_prior == 1) {
        long long tot = 0; double s = 0.0;
        for (int k = 0; k < P.K; k++) { tot += sizes[k]; s += lgamma((double)sizes[k] + 1.0); }
        if (tot == 0) return -0.0;
        return -(lgamma((double)P.N + 1.0) - s - lgamma((double)(P.N - tot) + 1.0));
    }
    if (P.size_prior == 2) {
        double s = 0.0;
        for (int k = 0; k < P.K; k++) { double z = (double)sizes[k]; s += log(z * z); }
        return -s;
    }
    return 0.0;
}
SB_DEV double size_prior_cta(Run &R, const int *sizes) {    // whole CTA; result in every thread
    if (R.P->size_prior == 0) return 0.0;
    sb_cta_sync();
    if (sb_tid() == 0) R.c->sred[0] = size_prior_of(*R.P, sizes);
    sb_cta_sync();
    double v = R.c->sred[0];
    sb_cta_sync();
    return v;
}

// -----------------------------------------------------------------------------------------
// geo prior: minimum spanning tree of one cluster (Prim), whole CTA, CTA-local
// (GeoPrior.__call__, compute_mst_distances: sbayes/model/prior.py:769-805,950-962)
// -----------------------------------------------------------------------------------------
SB_DEV void geo_cluster(Run &R, int k, int view, double &agg_out, double &val_out) {
    const SbProblem &P = *R.P;
    int m = 0;
    sb_cta_sync();
    if (sb_tid() == 0) {
        for (int o = 0; o < P.N; o++)
            if (cl_of(R, o, view) == k) { if (m < R.c->smem_cap) R.c->smem_mem[m] = o; m++; }
        R.c->sint[65] = m;
    }
    sb_cta_sync();
    m = R.c->sint[65];
    sb_cta_sync();
    if (m > R.c->smem_cap) { R.err = 3; m = R.c->smem_cap; }
    double sum = 0.0, mx = 0.0; int cnt = 0;
    if (m > 1) {
        for (int i = sb_tid(); i < m; i += sb_nthr())
            R.c->smem_best[i] = i == 0 ? -1.0 : P.cost[(size_t)R.c->smem_mem[0] * P.N + R.c->smem_mem[i]];
        sb_cta_sync();
        for (int it = 0; it < m - 1; it++) {
            if (sb_tid() == 0) {      // clusters under a geo prior are small: serial argmin is fine
                int bj = -1; double bv = SB_INF;
                for (int i = 0; i < m; i++) { double b = R.c->smem_best[i]; if (b >= 0.0 && (bj < 0 || b < bv)) { bj = i; bv = b; } }
                R.c->sint[66] = bj;
                R.c->sred[0] = bv;
            }
            sb_cta_sync();
            int j = R.c->sint[66]; double e = R.c->sred[0];
            sb_cta_sync();
            if (j < 0) break;
            if (e != 0.0 && e < SB_INF) { sum += e; if (cnt == 0 || e > mx) mx = e; cnt++; }
            for (int i = sb_tid(); i < m; i += sb_nthr()) {
                if (i == j) R.c->smem_best[i] = -1.0;
                else if (R.c->smem_best[i] >= 0.0) {
                    double c = P.cost[(size_t)R.c->smem_mem[j] * P.N + R.c->smem_mem[i]];
                    if (c < R.c->smem_best[i]) R.c->smem_best[i] = c;
                }
            }
            sb_cta_sync();
        }
    }
    double agg = 0.0;
    if (cnt > 0) agg = P.geo_agg == 0 ? sum / (double)cnt : (P.geo_agg == 1 ? sum : mx);
    agg_out = agg;
    val_out = geo_prob_fn(P, agg);
}

// GeoPrior.get_costs_per_object (prior.py:824-852) for one object; whole warp.
SB_DEV double geo_delta(Run &R, int o, int k, int view, int m_size) {
    const SbProblem &P = *R.P;
    double mn = SB_INF;
    for (int j = sb_lane(); j < P.N; j += 32)
        if (cl_of(R, j, view) == k) mn = fmin(mn, P.cost[(size_t)j * P.N + o]);
    mn = warp_min(mn);
    double before = (view && R.c->rp.dirty[k]) ? R.c->rp.c_geo_agg[k] : R.c->rp.geo_agg[k], after;
    double m = (double)m_size;
    if (P.geo_agg == 0) after = (mn + m * before) / (1.0 + m);
    else if (P.geo_agg == 1) after = mn + before;
    else after = fmax(mn, before);
    return geo_prob_fn(P, after) - geo_prob_fn(P, before);
}

// -----------------------------------------------------------------------------------------
// cluster-membership marginals of objects against cluster k
//   s_z = sum_f log sum_c l[o,f,c] * w_z[o,f,c]   (before the 1/T power)
// compute_cluster_posterior / compute_raw_cluster_probs (operators.py:1035-1095,1420-1472).
// l[.,.,0] is the tempered table of cluster k (to the power 1/T in the wide operator), the
// other columns are the plain tables of the object's confounder groups.
// -----------------------------------------------------------------------------------------
struct Marg { Prod p0, p1; double geo; int on; };

// tables over combos: tab[j,f,x].a = sum_{c>=1} l_c w0_c,  tab[j,f,x].b = l_0 w1_0 + sum_{c>=1} l_c w1_c.  CTA-local.
SB_DEV void build_tables(Run &R, int k, bool wide) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S, C = SB_CC(P);
    const bool powT = wide && R.T != 1.0;
    for (int idx = sb_tid(); idx < P.n_combos * F; idx += sb_nthr()) {
        const int j = idx / F, f = idx % F;
        const int pat = R.c->combo_pat[j];
        const double *w0 = R.c->rp.wq + ((size_t)(pat * 2 + 0) * F + f) * C;
        const double *w1 = R.c->rp.wq + ((size_t)(pat * 2 + 1) * F + f) * C;
        int gC[SB_UC]; double rC[SB_UC];
        SB_FOR_C1(c, C) {
            gC[c] = R.c->combo_rows[(size_t)j * P.n_conf + c - 1];
            rC[c] = gC[c] >= 0 ? R.c->rp.rP[(size_t)gC[c] * F + f] : 0.0;
        }
        const double rk = R.c->rp.rT[(size_t)k * F + f], w10 = w1[0];
        SbD2 *t = R.c->rp.tab + ((size_t)j * F + f) * S;
        for (int x = 0; x < S; x++) {
            double l0 = num_T(R, k, f, x) * rk;
            if (powT) l0 = pow_inv(l0, R.invT, false);
            double a0 = 0.0, a1 = l0 * w10;
            SB_FOR_C1(c, C) {
                if (gC[c] >= 0) {
                    double l = num_P(R, gC[c], f, x) * rC[c];
                    a0 += l * w0[c]; a1 += l * w1[c];
                }
            }
            SbD2 v; v.a = a0; v.b = a1; t[x] = v;
        }
    }
    sb_cta_sync();
}

// whole warp -> every lane.  TABLE: gather from tab (must be built for this k / wide).
template <bool TABLE>
SB_DEV void obj_marginals(const Run &R, int o, int k, bool wide, Prod &p0, Prod &p1) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S, C = SB_CC(P);
    const uint32_t *xw = (const uint32_t *)x_row(R, o);
    const int nwords = P.ldx >> 2;
    p0 = prod_one(); p1 = prod_one();
    if (TABLE) {
        const int j = m_combo(R, o);
        const SbD2 *t = R.c->rp.tab + (size_t)j * F * S;
        for (int w = sb_lane(); w < nwords; w += 32) {
            uint32_t word = xw[w];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                int x = (word >> (8 * b)) & 0xFF;
                if (x == SB_NA_U8) continue;
                const SbD2 v = t[(4 * w + b) * S + x];
                p0.m *= v.a; p1.m *= v.b;
            }
            prod_renorm(p0); prod_renorm(p1);
        }
    } else {
        int gC[SB_UC];
        SB_FOR_C1(c, C) gC[c] = conf_group(R, c, o);
        const int pat = m_pat(R, o);
        const double *w0 = R.c->rp.wq + (size_t)(pat * 2 + 0) * F * C;
        const double *w1 = R.c->rp.wq + (size_t)(pat * 2 + 1) * F * C;
        const bool powT = wide && R.T != 1.0;
        for (int w = sb_lane(); w < nwords; w += 32) {
            uint32_t word = xw[w];
#pragma unroll
            for (int b = 0; b < 4; b++) {
                int x = (word >> (8 * b)) & 0xFF;
                if (x == SB_NA_U8) continue;
                int f = 4 * w + b;
                double l0 = entry_T(R, k, f, x);
                if (powT) l0 = pow_inv(l0, R.invT, false);
                double a0 = 0.0, a1 = l0 * w1[f * C];
                SB_FOR_C1(c, C) {
                    if (gC[c] >= 0) {
                        double l = entry_P(R, gC[c], f, x);
                        a0 += l * w0[f * C + c]; a1 += l * w1[f * C + c];
                    }
                }
                p0.m *= a0; p1.m *= a1;
            }
            prod_renorm(p0); prod_renorm(p1);
        }
    }
    p0 = warp_prod_reduce(p0); p1 = warp_prod_reduce(p1);
}

// one lane -> one object: the same marginals from the per-combo tables without any shuffle.  Used
// by passes with many objects per warp (large N, few CTAs per chain): 32 independent gather streams
// per warp hide the latency of tables that live in the CTA's HBM slice.
SB_DEV void obj_marginals_lane(const Run &R, int o, Prod &p0, Prod &p1) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S;
    const int j = m_combo(R, o);
    const SbD2 *t = R.c->rp.tab + (size_t)j * F * S;
    const uint32_t *xw = (const uint32_t *)x_row(R, o);
    const int nwords = P.ldx >> 2;             // a multiple of 4 (rows are padded to 16 bytes)
    p0 = prod_one(); p1 = prod_one();
    // sixteen features per iteration; the next 16 bytes of the row are requested before the
    // gathers of the current ones, so one memory round trip (not two) bounds an iteration
    uint32_t nx[4] = {xw[0], xw[1], xw[2], xw[3]};
    for (int w = 0; w < nwords; w += 4) {
        const uint32_t wd[4] = {nx[0], nx[1], nx[2], nx[3]};
        if (w + 4 < nwords) { nx[0] = xw[w + 4]; nx[1] = xw[w + 5]; nx[2] = xw[w + 6]; nx[3] = xw[w + 7]; }
#pragma unroll
        for (int half = 0; half < 2; half++) {
            // eight features, branch-free: all gathers are issued before the first multiply (a
            // missing value reads entry 0 and contributes a factor one)
            double a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int x = (wd[2 * half + (i >> 2)] >> (8 * (i & 3))) & 0xFF;
                const bool na = x == SB_NA_U8;
                const int e = na ? 0 : (4 * w + 8 * half + i) * S + x;
                const SbD2 v = t[e];                  // one 16-byte gather: both tables are interleaved
                a[i] = na ? 1.0 : v.a; b[i] = na ? 1.0 : v.b;
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                p0.m *= (a[4 * h] * a[4 * h + 1]) * (a[4 * h + 2] * a[4 * h + 3]);
                p1.m *= (b[4 * h] * b[4 * h + 1]) * (b[4 * h + 2] * b[4 * h + 3]);
                prod_renorm(p0); prod_renorm(p1);
            }
        }
    }
}

// whether a pass over about n_obj objects (split over the team) should gather from tables
// (built redundantly by every CTA); uniform across the team
SB_DEV bool pass_uses_tables(const Run &R, long long n_obj) {
    const SbProblem &P = *R.P;
    // tiled: a CTA builds only the tables of the combos its objects (in combo order) belong to
    if (R.c->tile) return P.use_tables && n_obj * 3 > (long long)(P.n_combos + R.tsize) * P.S * 2;
    return P.use_tables && n_obj * 3 > (long long)P.n_combos * P.S * 2 * R.tsize;
}

// -----------------------------------------------------------------------------------------
// Tiled passes (problems whose per-combo pass tables do not fit in shared memory as a whole:
// BASELINE configs 4 and 5).  The objects are walked in combo order (P.perm), split over the team
// by position, so a CTA meets few combos; it builds the tables of those combos in shared memory a
// tile of combos at a time and gathers from shared memory only: per observation one byte of the
// packed row (16-byte loads), one 8- or 16-byte shared-memory load and one or two multiplies.
//   TK_ALTER  one table, the ratio b / a of the two mixtures (the posterior of AlterCluster only
//             needs log m1 - log m0 = sum_f log(b_f / a_f), operators.py:1060-1073)
//   TK_WIDE   both mixtures (compute_raw_cluster_probs adds EPS to m0 + m1, operators.py:1444-1471)
//   TK_JUMP   get_jump_lh (operators.py:1679-1722): conf + w0 pT[src], conf + w0 pT[tgt]
// An object is processed by L lanes (L = 1 when the CTA has at least as many objects as threads).
// -----------------------------------------------------------------------------------------
enum { TK_ALTER = 0, TK_WIDE, TK_JUMP, TK_ALTER2 };      // TK_ALTER2: TK_ALTER on feature pairs (P.xp)
#ifdef SB_HOSTSIM
static long g_tiled_calls[4];      // emulator builds: how often each kind ran (the tests assert coverage)
#endif

SB_DEV int combo_of_pos(const SbProblem &P, int p) {
    int lo = 0, hi = P.n_combos - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (sb_ldg(P.combo_start + mid) <= p) lo = mid; else hi = mid - 1; }
    return lo;
}

struct Row16 { uint32_t w[4]; };
SB_DEV Row16 load_row16(const uint8_t *p) {      // p is 16-byte aligned
    Row16 r;
#ifdef SB_HOSTSIM
    memcpy(&r, p, 16);
#else
    const uint4 t = *(const uint4 *)p; r.w[0] = t.x; r.w[1] = t.y; r.w[2] = t.z; r.w[3] = t.w;
#endif
    return r;
}

// tables of the combos [j0, j1) into the tile buffer; tile[0..1] = 1.0 is the entry a missing value reads
template <int KIND>
SB_DEV void build_tile(Run &R, int k, int k2, int j0, int j1) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S, C = SB_CC(P);
    constexpr int W = KIND == TK_ALTER ? 1 : 2;
    double *tile = R.c->tile + 2;
    const bool powT = KIND == TK_WIDE && R.T != 1.0;
    for (int idx = sb_tid(); idx < (j1 - j0) * F; idx += sb_nthr()) {
        const int jj = idx / F, f = idx % F, j = j0 + jj;
        const int pat = R.c->combo_pat[j];
        const double *w0 = R.c->rp.wq + ((size_t)(pat * 2 + 0) * F + f) * C;
        const double *w1 = R.c->rp.wq + ((size_t)(pat * 2 + 1) * F + f) * C;
        int gC[SB_UC]; double rC[SB_UC];
        SB_FOR_C1(c, C) {
            gC[c] = R.c->combo_rows[(size_t)j * P.n_conf + c - 1];
            rC[c] = gC[c] < 0 ? 0.0 : (KIND == TK_JUMP ? R.c->rp.rT[(size_t)gC[c] * F + f] : R.c->rp.rP[(size_t)gC[c] * F + f]);
        }
        const double rk = R.c->rp.rT[(size_t)k * F + f], w10 = w1[0];
        const double rk2 = KIND == TK_JUMP ? R.c->rp.rT[(size_t)k2 * F + f] : 0.0;
        double *t = tile + ((size_t)jj * F + f) * S * W;
        for (int x = 0; x < S; x++) {
            if (KIND == TK_JUMP) {
                double conf = 0.0;
                SB_FOR_C1(c, C) if (gC[c] >= 0) conf += w1[c] * (num_T(R, gC[c], f, x) * rC[c]);
                t[2 * x] = conf + w10 * (num_T(R, k, f, x) * rk);
                t[2 * x + 1] = conf + w10 * (num_T(R, k2, f, x) * rk2);
            } else {
                double l0 = num_T(R, k, f, x) * rk;
                if (powT) l0 = pow_inv(l0, R.invT, false);
                double a0 = 0.0, a1 = l0 * w10;
                SB_FOR_C1(c, C) {
                    if (gC[c] >= 0) {
                        double l = num_P(R, gC[c], f, x) * rC[c];
                        a0 += l * w0[c]; a1 += l * w1[c];
                    }
                }
                if (KIND == TK_ALTER) t[x] = a1 / a0;
                else { t[2 * x] = a0; t[2 * x + 1] = a1; }
            }
        }
    }
    if (sb_tid() == 0) { R.c->tile[0] = 1.0; R.c->tile[1] = 1.0; }
    sb_cta_sync();
}

// sixteen features of one row against one combo's table: products of the gathered entries
template <int W>
SB_DEV void chunk_products(const double *unit, const double *tc, int S, int fbase, const Row16 &r, Prod &a, Prod &b) {
#pragma unroll
    for (int half = 0; half < 2; half++) {
        double va[8], vb[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int x = (r.w[2 * half + (i >> 2)] >> (8 * (i & 3))) & 0xFF;
            const double *ptr = x == SB_NA_U8 ? unit : tc + (size_t)((fbase + 8 * half + i) * S + x) * W;
            if (W == 1) va[i] = *ptr;
            else { const SbD2 v = *(const SbD2 *)ptr; va[i] = v.a; vb[i] = v.b; }
        }
        a.m *= ((va[0] * va[1]) * (va[2] * va[3])) * ((va[4] * va[5]) * (va[6] * va[7]));
        prod_renorm(a);
        if (W == 2) {
            b.m *= ((vb[0] * vb[1]) * (vb[2] * vb[3])) * ((vb[4] * vb[5]) * (vb[6] * vb[7]));
            prod_renorm(b);
        }
    }
}

// ratio of the two mixtures of AlterCluster for one (combo, feature): r[x] for x < S, r[S] = 1 (missing value)
template <int SM>
SB_DEV void ratio_row(const Run &R, int k, int j, int f, double (&r)[SM + 1]) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S, C = SB_CC(P);
#pragma unroll
    for (int x = 0; x <= SM; x++) r[x] = 1.0;
    if (f >= F) return;
    // Branch-free: the loads behind one component (S counts, S concentrations) are issued together, a
    // component without a group for this combo reads row 0 and is multiplied by a zero reciprocal (its
    // terms are exact zeros, so the sums equal those of build_tables)
    const int32_t *SB_RESTRICT counts = R.c->rp.counts;
    const double *SB_RESTRICT alpha = R.c->alpha;
    const int pat = R.c->combo_pat[j];
    const double *SB_RESTRICT w0 = R.c->rp.wq + ((size_t)(pat * 2 + 0) * F + f) * C;
    const double *SB_RESTRICT w1 = R.c->rp.wq + ((size_t)(pat * 2 + 1) * F + f) * C;
    double a0[SM], a1[SM];
    {
        const double rk = R.c->rp.rT[(size_t)k * F + f], w10 = w1[0];
        const size_t cb = ((size_t)k * F + f) * S, ab = ((size_t)arow(P, k) * F + f) * S;
        int cn[SM]; double al[SM];
#pragma unroll
        for (int x = 0; x < SM; x++) { const int xx = x < S ? x : S - 1; cn[x] = counts[cb + xx]; al[x] = alpha[ab + xx]; }
#pragma unroll
        for (int x = 0; x < SM; x++) {
            const double u = unif_at(R, f, x < S ? x : S - 1);
            a0[x] = 0.0;
            a1[x] = ((u + (al[x] - u) * R.invTp + (double)cn[x] * R.invT) * rk) * w10;
        }
    }
    SB_FOR_C1(c, C) {
        const int gq = R.c->combo_rows[(size_t)j * P.n_conf + c - 1], g = gq < 0 ? 0 : gq;
        const double rq = R.c->rp.rP[(size_t)g * F + f], rc = gq < 0 ? 0.0 : rq, w0c = w0[c], w1c = w1[c];
        const size_t cb = ((size_t)g * F + f) * S, ab = ((size_t)arow(P, g) * F + f) * S;
        int cn[SM]; double al[SM];
#pragma unroll
        for (int x = 0; x < SM; x++) { const int xx = x < S ? x : S - 1; cn[x] = counts[cb + xx]; al[x] = alpha[ab + xx]; }
        if (gq >= 0) {
#pragma unroll
            for (int x = 0; x < SM; x++) {
                const double l = (al[x] + (double)cn[x]) * rc;
                a0[x] += l * w0c; a1[x] += l * w1c;
            }
        }
    }
#pragma unroll
    for (int x = 0; x < SM; x++) if (x < S) r[x] = a1[x] / a0[x];
}

// ratio rows of the combos [j0, j1) into the CTA's scratch in HBM (all rows in flight at once: the loads
// of counts / reciprocals behind a row are L2 round trips), then, tile by tile, the pair tables from them:
// t[j][pair][a * (S+1) + b] = r[2 pair][a] * r[2 pair + 1][b]; pairs beyond the last feature (row padding) hold 1.0
template <int SM>
SB_DEV void build_ratio_rows(Run &R, int k, int j0, int j1) {
    const SbProblem &P = *R.P;
    const int F = P.F, S1 = P.S + 1;
    for (int idx = sb_tid(); idx < (j1 - j0) * F; idx += sb_nthr()) {
        const int j = j0 + idx / F, f = idx % F;
        double r[SM + 1];
        ratio_row<SM>(R, k, j, f, r);
        double *dst = R.c->rp.rs + ((size_t)j * F + f) * S1;
#pragma unroll
        for (int x = 0; x <= SM; x++) if (x < S1) dst[x] = x == P.S ? 1.0 : r[x];
    }
    sb_cta_sync();
}
template <int SM>
SB_DEV void build_tile_pairs(Run &R, int j0, int j1) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S, S1 = S + 1, PS = S1 * S1, ldp = P.ldp;
    double *tile = R.c->tile + 2;
    for (int idx = sb_tid(); idx < (j1 - j0) * ldp; idx += sb_nthr()) {
        const int jj = idx / ldp, pr = idx % ldp, fa = 2 * pr, fb = fa + 1;
        const double *sa = R.c->rp.rs + ((size_t)(j0 + jj) * F + fa) * S1, *sb = sa + S1;
        double ra[SM + 1], rb[SM + 1];
#pragma unroll
        for (int x = 0; x <= SM; x++) { ra[x] = (fa < F && x < S1) ? sa[x] : 1.0; rb[x] = (fb < F && x < S1) ? sb[x] : 1.0; }
        double *t = tile + ((size_t)jj * ldp + pr) * PS;
#pragma unroll
        for (int a = 0; a <= SM; a++) {
            if (a > S) continue;
#pragma unroll
            for (int b = 0; b <= SM; b++)
                if (b <= S) t[a * S1 + b] = ra[a] * rb[b];
        }
    }
    sb_cta_sync();
}

// sixteen feature pairs of one row against one combo's pair table
SB_DEV void chunk_pairs(const double *tc, int PS, int pbase, const Row16 &r, Prod &a) {
#pragma unroll
    for (int half = 0; half < 2; half++) {
        double va[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int code = (r.w[2 * half + (i >> 2)] >> (8 * (i & 3))) & 0xFF;
            va[i] = tc[(size_t)(pbase + 8 * half + i) * PS + code];
        }
        a.m *= ((va[0] * va[1]) * (va[2] * va[3])) * ((va[4] * va[5]) * (va[6] * va[7]));
        prod_renorm(a);
    }
}

// sel(o): is object o part of the pass; tail(o, on, a, b) receives, in one lane per object of the
// CTA's positions, the products over the features (a = ratio product for TK_ALTER; a, b = the two
// mixtures otherwise).  CTA-local barriers only.
template <int KIND, class Sel, class Tail>
SB_DEV void tiled_pass(Run &R, int k, int k2, Sel sel, Tail tail) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S;
    constexpr int W = (KIND == TK_ALTER || KIND == TK_ALTER2) ? 1 : 2;
    constexpr bool PAIRS = KIND == TK_ALTER2;
    int plo, phi; team_range(R, P.N, plo, phi);
#ifdef SB_HOSTSIM
    if (sb_tid() == 0 && R.rank == 0) g_tiled_calls[KIND]++;
#endif
    if (plo >= phi) return;
    const int PS = (S + 1) * (S + 1);
    const int per_combo = PAIRS ? P.ldp * PS : F * S * W;
    const int cap = (R.c->tile_cap - 2) / per_combo;              // combos per tile (>= 1: the host checks)
    const int j_lo = combo_of_pos(P, plo), j_hi = combo_of_pos(P, phi - 1);
    int L = 1;
    while (L < 32 && (phi - plo) * L * 2 <= sb_nthr()) L <<= 1;
    const int q = sb_lane() & (L - 1), gpw = 32 / L, ngroups = sb_nwarps() * gpw;
    const int nchunks = (PAIRS ? P.ldp : P.ldx) >> 4;
    const double *unit = R.c->tile;
    if (PAIRS) {
        if (S <= 4) build_ratio_rows<4>(R, k, j_lo, j_hi + 1);
        else if (S <= 6) build_ratio_rows<6>(R, k, j_lo, j_hi + 1);
        else if (S <= 8) build_ratio_rows<8>(R, k, j_lo, j_hi + 1);
        else build_ratio_rows<15>(R, k, j_lo, j_hi + 1);
        PH(R, 34);
    }
    for (int j0 = j_lo; j0 <= j_hi; j0 += cap) {
        const int j1 = j0 + cap <= j_hi + 1 ? j0 + cap : j_hi + 1;
        if (PAIRS) {
            if (S <= 4) build_tile_pairs<4>(R, j0, j1);
            else if (S <= 6) build_tile_pairs<6>(R, j0, j1);
            else if (S <= 8) build_tile_pairs<8>(R, j0, j1);
            else build_tile_pairs<15>(R, j0, j1);
        } else build_tile<PAIRS ? TK_ALTER : KIND>(R, k, k2, j0, j1);
        PH(R, 32);
        int pb = sb_ldg(P.combo_start + j0), pe = sb_ldg(P.combo_start + j1);
        if (pb < plo) pb = plo;
        if (pe > phi) pe = phi;
        // one object ahead: position -> object -> (selected?, combo, row pointer) -> first 16 bytes of the row is a chain of
        // dependent loads (L2 round trips at the large shapes); the next object's chain is in flight while this one's row is walked
        int p_n = pb + sb_warp() * gpw + sb_lane() / L;
        int o_n = p_n < pe ? sb_ldg(P.perm + p_n) : 0;
        bool on_n = p_n < pe && sel(o_n);
        const double *tc_n = unit; const uint8_t *xr_n = PAIRS ? P.xp : P.x;
        if (on_n) { tc_n = unit + 2 + (size_t)(m_combo(R, o_n) - j0) * per_combo; xr_n = PAIRS ? P.xp + (size_t)o_n * P.ldp : x_row(R, o_n); }
        Row16 nx = load_row16(xr_n + (q < nchunks ? q : 0) * 16);
        for (int p0 = pb + sb_warp() * gpw; p0 < pe; p0 += ngroups) {   // warp-uniform trip count
            const int p = p_n, o = o_n;
            const bool active = p < pe, on = on_n;
            const double *tc = tc_n; const uint8_t *xr = xr_n;
            Row16 cur0 = nx;
            p_n = p + ngroups;
            o_n = p_n < pe ? sb_ldg(P.perm + p_n) : 0;
            on_n = p_n < pe && sel(o_n);
            tc_n = unit; xr_n = PAIRS ? P.xp : P.x;
            if (on_n) { tc_n = unit + 2 + (size_t)(m_combo(R, o_n) - j0) * per_combo; xr_n = PAIRS ? P.xp + (size_t)o_n * P.ldp : x_row(R, o_n); }
            const Row16 first_n = load_row16(xr_n + (q < nchunks ? q : 0) * 16);
            Prod a = prod_one(), b = prod_one();
            if (on) {
                nx = cur0;
                for (int ch = q; ch < nchunks; ch += L) {
                    const Row16 cur = nx;
                    if (ch + L < nchunks) nx = load_row16(xr + (ch + L) * 16);   // requested before the gathers of the current chunk
                    if (PAIRS) chunk_pairs(tc, PS, ch * 16, cur, a);
                    else chunk_products<W>(unit, tc, S, ch * 16, cur, a, b);
                }
            }
            nx = first_n;
            for (int off = L >> 1; off > 0; off >>= 1) {
                a.m *= sb_shfl_f64(a.m, sb_lane() ^ off); a.e += sb_shfl_i32(a.e, sb_lane() ^ off); prod_renorm(a);
                if (W == 2) { b.m *= sb_shfl_f64(b.m, sb_lane() ^ off); b.e += sb_shfl_i32(b.e, sb_lane() ^ off); prod_renorm(b); }
            }
            if (active && q == 0) tail(o, on, a, b);
        }
        sb_cta_sync();
        PH(R, 33);
    }
}

// AlterCluster.compute_cluster_posterior (operators.py:1035-1073) from the marginals (one lane)
SB_DEV double posterior_tail(const Run &R, const Marg &a, bool flat, bool use_geo) {
    if (flat) return 0.5;
    double s0 = prod_log(a.p0), s1 = prod_log(a.p1);
    double d = s1 * R.invT - s0 * R.invT;
    if (use_geo) d = (s1 * R.invT + a.geo * R.invTp) - s0 * R.invT;
    double post = sb_expit(d);
    return (post + SB_SMOOTH) / (1.0 + 2.0 * SB_SMOOTH);
}

// get_jump_lh (operators.py:1679-1722) for one member o of k_src, whole warp: products of
// (E_conf + w_0 p_src) and (E_conf + w_0 p_tgt) over the features, tempered tables throughout
SB_DEV void jump_marginals(const Run &R, int o, int k_src, int k_tgt, Prod &ps, Prod &pj) {
    const SbProblem &P = *R.P;
    const int F = P.F, C = SB_CC(P);
    int gC[SB_UC];
    SB_FOR_C1(c, C) gC[c] = conf_group(R, c, o);
    const int pat = m_pat(R, o);
    const double *wh = R.c->rp.wq + (size_t)(pat * 2 + 1) * F * C;
    const uint32_t *xw = (const uint32_t *)x_row(R, o);
    const int nwords = P.ldx >> 2;
    ps = prod_one(); pj = prod_one();
    for (int w = sb_lane(); w < nwords; w += 32) {
        uint32_t word = xw[w];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            int x = (word >> (8 * b)) & 0xFF;
            if (x == SB_NA_U8) continue;
            int f = 4 * w + b;
            double conf = 0.0;
            SB_FOR_C1(c, C)
                if (gC[c] >= 0) conf += wh[f * C + c] * entry_T(R, gC[c], f, x);
            ps.m *= conf + wh[f * C] * entry_T(R, k_src, f, x);
            pj.m *= conf + wh[f * C] * entry_T(R, k_tgt, f, x);
        }
        prod_renorm(ps); prod_renorm(pj);
    }
    ps = warp_prod_reduce(ps); pj = warp_prod_reduce(pj);
}
SB_DEV double jump_tail(const Run &R, const Prod &ps, const Prod &pj) {
    double lh_stay = exp(prod_log(ps) * R.invT) + SB_EPS32, lh_jump = exp(prod_log(pj) * R.invT) + SB_EPS32;
    return lh_jump / (lh_jump + lh_stay);
}

// -----------------------------------------------------------------------------------------
// proposal bookkeeping: a proposal edits the replica in place (counts, totals, histogram
// copy, per-(group, feature) terms of the touched rows) and is undone on a reject
// -----------------------------------------------------------------------------------------
struct Prop {
    int changed;        // 0 => direct reject before anything was touched
    int force_reject;   // reject after the candidate was built
    int M;              // number of changed objects
    const int *S;       // their ids (shared-memory list or st.list_a)
    int src_changed;    // candidate source rows were written, H has a candidate copy
    int need_sync;      // a team barrier is still needed between the last shared read and the commit
    double lq, lqb;
    double sp_new, geo_new, wp_new;
};
SB_DEV Prop prop_init(int changed, double lq, double lqb) {
    Prop p; p.changed = changed; p.force_reject = 0; p.M = 0; p.S = nullptr; p.src_changed = 0; p.need_sync = 0;
    p.lq = lq; p.lqb = lqb; p.sp_new = 0.0; p.geo_new = 0.0; p.wp_new = 0.0;
    return p;
}

// mark the group rows object o leaves or joins
SB_DEV void mark_object(Run &R, int o) {
    const SbProblem &P = *R.P;
    int a = m_cl(R, o), b = cl_of(R, o, 1);
    if (a >= 0) R.c->rp.dirty[a] = 1;
    if (b >= 0) R.c->rp.dirty[b] = 1;
    for (int c = 1; c < SB_CC(P); c++) { int g = conf_group(R, c, o); if (g >= 0) R.c->rp.dirty[g] = 1; }
}
// group row an observation of object o assigned to component c counts towards
SB_DEV int obs_group(const Run &R, int o, int c, int view) {
    return c == 0 ? cl_of(R, o, view) : conf_group(R, c, o);
}
SB_DEV void count_add(Run &R, int g, int f, int x, int d) {
    const size_t gf = (size_t)g * R.P->F + f;
    sb_atomic_add_i32(R.c->rp.counts + gf * R.P->S + x, d);
    sb_atomic_add_i32(R.c->rp.dcnt + gf * R.P->S + x, d);
    sb_atomic_add_i32(R.c->rp.nn + gf, d);
}

// opens a proposal over the M objects S[]: candidate histogram, dirty rows + their backups.
// Ends with a barrier.  Returns the candidate histogram.
SB_DEV int32_t *proposal_open(Run &R, int M, const int *S) {
    const SbProblem &P = *R.P;
    const int nH = P.n_pat * 2 * P.F * SB_CC(P), F = P.F;
    const int32_t *h = R.c->rp.H[R.hcur];
    int32_t *hc = R.c->rp.H[R.hcur ^ 1];
    for (int i = sb_tid(); i < nH; i += sb_nthr()) hc[i] = h[i];
    for (int m = sb_tid(); m < M; m += sb_nthr()) { mark_object(R, S[m]); R.c->rp.selbuf[m] = m_osel(R, S[m]) ^ 1; }
    sb_cta_sync();
    if (sb_warp() == 0) {
        int n = 0;
        for (int b = 0; b < P.G_tot; b += 32) {
            int g = b + sb_lane();
            bool d = g < P.G_tot && R.c->rp.dirty[g];
            unsigned bal = sb_ballot(d);
            if (d) R.c->rp.dlist[n + sb_popc(bal & ((1u << sb_lane()) - 1u))] = g;
            n += sb_popc(bal);
        }
        if (sb_lane() == 0) R.c->sint[67] = n;
    }
    sb_cta_sync();
    R.nd = R.c->sint[67];
    for (int idx = sb_tid(); idx < R.nd * F; idx += sb_nthr()) {
        const size_t gf = (size_t)R.c->rp.dlist[idx / F] * F + idx % F;
        R.c->rp.bk_llgf[gf] = R.c->rp.llgf[gf]; R.c->rp.bk_rP[gf] = R.c->rp.rP[gf]; R.c->rp.bk_rT[gf] = R.c->rp.rT[gf];
    }
    for (int i = sb_tid(); i < R.nd; i += sb_nthr()) R.c->rp.bk_lh[R.c->rp.dlist[i]] = R.c->rp.lh[R.c->rp.dlist[i]];
    sb_cta_sync();
    return hc;
}

// after the counts of the dirty rows were edited: update the per-(group, feature) terms of
// the entries that changed and the per-group likelihoods.  CTA-local, ends with a barrier.
SB_DEV void proposal_rows(Run &R) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S;
    sb_cta_sync();
    for (int idx = sb_tid(); idx < R.nd * F; idx += sb_nthr()) {
        const int g = R.c->rp.dlist[idx / F], f = idx % F;
        const size_t gf = (size_t)g * F + f, b = gf * S;
        // most (group, feature) rows of a dirty group saw no change: find that out with all loads of the row in
        // flight at once (the replicas of the large shapes live in the HBM slice: a dependent load is an L2 round trip)
        int touched = 0;
#ifndef SB_HOSTSIM
#pragma unroll 8
#endif
        for (int s = 0; s < S; s++) touched |= R.c->rp.dcnt[b + s];
        if (!touched) continue;
        int D = 0; bool any = false; double delta = 0.0;
        for (int s = 0; s < S; s++) {
            int d = R.c->rp.dcnt[b + s];
            if (d == 0) continue;
            any = true; D += d;
            double a = alpha_at(R, g, f, s);
            if (a > 0.0) delta += lgamma_diff((double)(R.c->rp.counts[b + s] - d) + a, d);
        }
        if (!any) continue;
        if (D != 0) {
            const int n = R.c->rp.nn[gf];
            const double as = asum_at(R, g, f);
            delta -= lgamma_diff((double)(n - D) + as, D);
            R.c->rp.rP[gf] = 1.0 / (as + (double)n);
            R.c->rp.rT[gf] = 1.0 / (R.c->rp.bsT[gf] + (double)n * R.invT);
        }
        R.c->rp.llgf[gf] += delta;
    }
    sb_cta_sync();
    rows_lh(R, R.c->rp.dlist, R.nd);
}

// closes the proposal: keep (accept) or undo (reject) the edits of the dirty rows
SB_DEV void proposal_close(Run &R, bool accept) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S;
    sb_cta_sync();      // every warp has read the candidate likelihoods (ll_total)
    for (int idx = sb_tid(); idx < R.nd * F; idx += sb_nthr()) {
        const int g = R.c->rp.dlist[idx / F], f = idx % F;
        const size_t gf = (size_t)g * F + f, b = gf * S;
        int touched = 0;
#ifndef SB_HOSTSIM
#pragma unroll 8
#endif
        for (int s = 0; s < S; s++) touched |= R.c->rp.dcnt[b + s];
        if (!touched) continue;            // (a row whose changes cancelled keeps its terms: nothing to undo either)
        int D = 0;
        for (int s = 0; s < S; s++) {
            int d = R.c->rp.dcnt[b + s];
            if (d == 0) continue;
            D += d;
            if (!accept) R.c->rp.counts[b + s] -= d;
            R.c->rp.dcnt[b + s] = 0;
        }
        if (!accept) {
            if (D != 0) R.c->rp.nn[gf] -= D;
            R.c->rp.llgf[gf] = R.c->rp.bk_llgf[gf]; R.c->rp.rP[gf] = R.c->rp.bk_rP[gf]; R.c->rp.rT[gf] = R.c->rp.bk_rT[gf];
        }
    }
    for (int i = sb_tid(); i < R.nd; i += sb_nthr()) {
        const int g = R.c->rp.dlist[i];
        if (!accept) R.c->rp.lh[g] = R.c->rp.bk_lh[g];
        else if (P.geo && g < P.K) { R.c->rp.geo_agg[g] = R.c->rp.c_geo_agg[g]; R.c->rp.geo_val[g] = R.c->rp.c_geo_val[g]; }
        R.c->rp.dirty[g] = 0;
    }
    R.nd = 0;
    sb_cta_sync();
}

// geo prior of the candidate membership (dirty clusters re-evaluated)
SB_DEV double candidate_geo(Run &R, bool recompute) {
    const SbProblem &P = *R.P;
    if (!P.geo) return 0.0;
    double geo = 0.0;
    for (int k = 0; k < P.K; k++) {
        if (recompute && R.c->rp.dirty[k]) {
            double a, v; geo_cluster(R, k, 1, a, v);
            if (sb_tid() == 0) { R.c->rp.c_geo_agg[k] = a; R.c->rp.c_geo_val[k] = v; }
            geo += v;
        } else {
            if (R.c->rp.dirty[k] && sb_tid() == 0) { R.c->rp.c_geo_agg[k] = R.c->rp.geo_agg[k]; R.c->rp.c_geo_val[k] = R.c->rp.geo_val[k]; }
            geo += R.c->rp.geo_val[k];
        }
    }
    sb_cta_sync();
    return geo;
}

// sample_categorical (sbayes/preprocessing.py:224-256): first c with u < cumsum(q)[c] / sum(q)
SB_DEV int categorical(const double (&q)[SB_UC], int C, double u) {
    double cum = 0.0, last = 0.0;
    SB_FOR_C(c, C) last += q[c];
    int nw = 0; bool got = false;
    SB_FOR_C(c, C) { cum += q[c]; if (!got && u < cum / last) { nw = c; got = true; } }
    return nw;
}
SB_DEV double pick_c(const double (&q)[SB_UC], int C, int i) {     // q[i] without a dynamic register index
    double v = 0.0;
    SB_FOR_C(c, C) if (c == i) v = q[c];
    return v;
}

// ClusterOperator.gibbs_sample_source / ClusterJump.gibbs_sample_source_jump
// (operators.py:796-849,1775-1827) for the M changed objects S[0..M).  CTA-local: every
// CTA resamples all M objects (identical draws), edits its replica and writes the same
// candidate source rows.  Returns log_q, log_q_back of the source draw; leaves the
// candidate rows, the candidate histogram and pr.sp_new / pr.geo_new.
SB_DEV void resample_sources(Run &R, Prop &pr, int M, const int *S, int k_new, int k_old, bool jump, double &lq, double &lqb) {
    const SbProblem &P = *R.P;
    const int F = P.F, C = SB_CC(P);
    const long long items = (long long)M * F;
    PH(R, 5);
    int32_t *Hc = proposal_open(R, M, S);
    PH(R, 6);
    // R1: remove the objects' own observations; H -= old
    for (long long it = sb_tid(); it < items; it += sb_nthr()) {
        int m = (int)(it / F), f = (int)(it % F), o = S[m];
        int x = sb_ldg(P.x + (size_t)o * P.ldx + f);
        if (x == SB_NA_U8) continue;
        int c = R.c->st.src[m_osel(R, o)][(size_t)o * P.ldx + f];
        int g = obs_group(R, o, c, 0);
        int pat = m_pat(R, o), z_old = m_cl(R, o) >= 0;
        if (g >= 0) count_add(R, g, f, x, -1);
        sb_atomic_add_i32(Hc + ((size_t)(pat * 2 + z_old) * F + f) * C + c, -1);
    }
    sb_cta_sync();
    PH(R, 7);
    // R2: conditional source distribution given the unchanged observations, draw.  With many
    // flipped objects (the wide operator) the items are split over the team: the draws are
    // addressed by item, the candidate sources go to the shared state in HBM, and the team sum is
    // the barrier that publishes them; a single flipped object is done by every CTA on its own
    const bool unitT = R.T == 1.0;
    const bool split = R.tsize > 1 && items >= 4LL * sb_nthr();
    int rlo = 0, rhi = (int)items;
    if (split) team_range(R, (int)items, rlo, rhi);
    double acc[3] = {0.0, 0.0, 0.0};
    for (long long it = rlo + sb_tid(); it < rhi; it += sb_nthr()) {
        int m = (int)(it / F), f = (int)(it % F), o = S[m];
        int x = sb_ldg(P.x + (size_t)o * P.ldx + f);
        uint8_t *sn = R.c->st.src[m_osel(R, o) ^ 1] + (size_t)o * P.ldx;
        if (x == SB_NA_U8) { sn[f] = SB_NA_U8; continue; }
        int pat = m_pat(R, o), z_old = m_cl(R, o) >= 0, z_new = cl_of(R, o, 1) >= 0;
        const double *tw = jump ? R.c->rp.w1 : R.c->rp.wp;
        const double *wf = tw + ((size_t)(pat * 2 + z_new) * F + f) * C;
        const double *wb = tw + ((size_t)(pat * 2 + (jump ? z_new : z_old)) * F + f) * C;
        double q[SB_UC], qb[SB_UC], tot = 0.0, totb = 0.0;
        SB_FOR_C(c, C) {
            double l = 1.0, lo = 1.0;
            if (!R.sfp) {
                int g = c == 0 ? k_new : conf_group(R, c, o);
                l = g < 0 ? 0.0 : pow_inv(entry_T_fresh(R, g, f, x), R.invT, unitT);
                lo = l;
                if (jump && c == 0) lo = pow_inv(entry_T_fresh(R, k_old, f, x), R.invT, unitT);
            }
            q[c] = wf[c] * l; qb[c] = wb[c] * lo;
            tot += q[c]; totb += qb[c];
        }
        if (!R.sfp) { SB_FOR_C(c, C) { q[c] /= tot; qb[c] /= totb; } }
        int nw = categorical(q, C, draw_uniform(R.D, SITE_GIBBS, (uint32_t)((size_t)m * F + f)));
        int c_old = R.c->st.src[m_osel(R, o)][(size_t)o * P.ldx + f];
        acc[0] += log(pick_c(q, C, nw)); acc[1] += log(pick_c(qb, C, c_old));
        acc[2] += R.c->rp.lw[((size_t)(pat * 2 + z_new) * F + f) * C + nw] - R.c->rp.lw[((size_t)(pat * 2 + z_old) * F + f) * C + c_old];
        sn[f] = (uint8_t)nw;
    }
    if (split) team_sum<3>(R, acc); else sb_cta_sync();
    PH(R, 8);
    // R3: add the new observations; H += new
    for (long long it = sb_tid(); it < items; it += sb_nthr()) {
        int m = (int)(it / F), f = (int)(it % F), o = S[m];
        int x = sb_ldg(P.x + (size_t)o * P.ldx + f);
        if (x == SB_NA_U8) continue;
        int nw = R.c->st.src[m_osel(R, o) ^ 1][(size_t)o * P.ldx + f];
        int pat = m_pat(R, o), z_new = cl_of(R, o, 1) >= 0;
        int g = obs_group(R, o, nw, 1);
        if (g >= 0) count_add(R, g, f, x, 1);
        sb_atomic_add_i32(Hc + ((size_t)(pat * 2 + z_new) * F + f) * C + nw, 1);
    }
    if (!split) block_sum<3>(R, acc);
    lq = acc[0]; lqb = acc[1];
    pr.sp_new = R.pr_src + acc[2];
    pr.src_changed = 1;
    PH(R, 9);
    proposal_rows(R);
    PH(R, 10);
    pr.geo_new = candidate_geo(R, true);
    PH(R, 11);
}

// neighbours of a cluster on the CSR adjacency (sbayes/util.py:121-143)
SB_DEV bool touches_cluster(const Run &R, int o, int k, int view) {
    const SbProblem &P = *R.P;
    if (!P.adj_ptr) return false;
    for (int e = P.adj_ptr[o]; e < P.adj_ptr[o + 1]; e++)
        if (cl_of(R, P.adj_idx[e], view) == k) return true;
    return false;
}
SB_DEV bool touches_flag(const Run &R, int o, const uint8_t *flag) {
    const SbProblem &P = *R.P;
    if (!P.adj_ptr) return false;
    for (int e = P.adj_ptr[o]; e < P.adj_ptr[o + 1]; e++)
        if (flag[P.adj_idx[e]]) return true;
    return false;
}

// AlterCluster.grow_candidates (operators.py:1097-1107): writes flag_b[o] for the team's
// objects.  For the two-step neighbourhood one extra team barrier.
SB_DEV void grow_candidates(Run &R, int k, int nb, int view) {
    const SbProblem &P = *R.P;
    int lo, hi; team_range(R, P.N, lo, hi);
    if (nb == NB_TWOSTEP) {
        for (int o = lo + sb_tid(); o < hi; o += sb_nthr()) R.c->st.flag_a[o] = touches_cluster(R, o, k, view);
        team_sync(R);
    }
    for (int o = lo + sb_tid(); o < hi; o += sb_nthr()) {
        bool freeo = cl_of(R, o, view) < 0, c = freeo;
        if (freeo && nb == NB_DIRECT) c = touches_cluster(R, o, k, view);
        if (freeo && nb == NB_TWOSTEP) c = touches_flag(R, o, R.c->st.flag_a);
        R.c->st.flag_b[o] = c;
    }
    sb_cta_sync();
}

SB_DEV void candidate_sizes(Run &R, int k0, int v0, int k1 = -1, int v1 = 0) {
    const SbProblem &P = *R.P;
    for (int i = sb_tid(); i < P.K; i += sb_nthr()) R.c->rp.ssz[i] = i == k0 ? v0 : (i == k1 ? v1 : R.c->rp.sizes[i]);
    sb_cta_sync();
}
SB_DEV int n_free(const Run &R) {
    int occ = 0;
    for (int k = 0; k < R.P->K; k++) occ += R.c->rp.sizes[k];
    return R.P->N - occ;
}

// -----------------------------------------------------------------------------------------
// AlterCluster grow / shrink (operators.py:969-1249)
// -----------------------------------------------------------------------------------------
// posterior-derived value of the CTA's objects selected by `sel(o)`: v(o) = fn(posterior(o));
// lane_out(o, v) receives the values.  One pass; tables are built if it pays.
template <class Sel, class Out>
SB_DEV void posterior_pass(Run &R, int k, int view, bool flat, bool use_geo, int m_size, long long n_obj, Sel sel, Out lane_out) {
    const SbProblem &P = *R.P;
    int lo, hi; team_range(R, P.N, lo, hi);
    const bool geo = use_geo && P.geo && !flat;
    const bool tables = !flat && pass_uses_tables(R, n_obj) && !(R.c->tile && geo);
    if (tables && R.c->tile) {
        // tiled: objects in combo order, tables built tile by tile in shared memory (ratio form)
        auto tail = [&](int o, bool on, const Prod &a, const Prod &) {
            double post = 0.0;
            if (on) { post = sb_expit(prod_log(a) * R.invT); post = (post + SB_SMOOTH) / (1.0 + 2.0 * SB_SMOOTH); }
            lane_out(o, on, post);
        };
        if (P.xp) tiled_pass<TK_ALTER2>(R, k, 0, sel, tail);
        else tiled_pass<TK_ALTER>(R, k, 0, sel, tail);
        // the values of an object block were written by the CTAs that own its positions
        if (R.tsize > 1) team_sync(R); else sb_cta_sync();
        PH(R, 21);
        return;
    }
    if (tables) build_tables(R, k, false);
    PH(R, 20);
    if ((tables || flat) && !geo && hi - lo >= 32 * sb_nwarps()) {
        // many objects per warp: a lane per object (no shuffles, 32 gather streams in flight per
        // warp; for the flat proposals of the naive operators just the candidate predicate)
        for (int o = lo + sb_tid(); o < hi; o += sb_nthr()) {
            Marg a; a.on = sel(o) ? 1 : 0; a.geo = 0.0; a.p0 = prod_one(); a.p1 = prod_one();
            if (a.on && !flat) obj_marginals_lane(R, o, a.p0, a.p1);
            lane_out(o, a.on != 0, a.on ? posterior_tail(R, a, flat, geo) : 0.0);
        }
        sb_cta_sync();
        PH(R, 21);
        return;
    }
    if (!flat && hi - lo >= 8 * sb_nwarps()) {
        // few selected objects among many (the members of one cluster): compact them first, then a warp per
        // selected object; the others get their zero in the compaction sweep
        int32_t *list = list_mine(R) + lo;
        const int n = cta_compact(R, lo, hi, list, [&](int o) { return (bool)sel(o); }, [&](int o) { lane_out(o, false, 0.0); });
        batch_items<Marg>(0, n,
            [&](int i) {
                const int o = list[i];
                Marg a; a.on = 1; a.geo = 0.0;
                if (tables) obj_marginals<true>(R, o, k, false, a.p0, a.p1);
                else obj_marginals<false>(R, o, k, false, a.p0, a.p1);
                if (geo) a.geo = geo_delta(R, o, k, view, m_size);
                return a;
            },
            [&](int i, const Marg &a) { lane_out(list[i], true, posterior_tail(R, a, flat, geo)); });
        sb_cta_sync();
        PH(R, 21);
        return;
    }
    batch_items<Marg>(lo, hi,
        [&](int o) {
            Marg a; a.on = sel(o) ? 1 : 0; a.geo = 0.0; a.p0 = prod_one(); a.p1 = prod_one();
            if (a.on && !flat) {
                if (tables) obj_marginals<true>(R, o, k, false, a.p0, a.p1);
                else obj_marginals<false>(R, o, k, false, a.p0, a.p1);
                if (geo) a.geo = geo_delta(R, o, k, view, m_size);
            }
            return a;
        },
        [&](int o, const Marg &a) { lane_out(o, a.on != 0, a.on ? posterior_tail(R, a, flat, geo) : 0.0); });
    sb_cta_sync();
    PH(R, 21);
}

SB_DEV Prop op_alter_cluster(Run &R, bool gibbsish, bool use_geo, int nb) {
    const SbProblem &P = *R.P;
    Prop pr = prop_init(0, 0.0, -SB_INF);
    int k = draw_index(R.D, SITE_CLUSTER, 0, P.K);
    int size = R.c->rp.sizes[k];
    double extra = 0.0; bool grow;
    if (size == P.min_size) { grow = true; extra = -log(2.0); }
    else if (size == P.max_size) { grow = false; extra = -log(2.0); }
    else grow = draw_uniform(R.D, SITE_GROW) < R.p_grow;
    const bool flat = R.sfp || !gibbsish;
    double *val = val_mine(R, R.vpar);
    double *const *vpeers = R.xworld > 1 ? R.c->valp[R.vpar] : nullptr;
    const int vb = R.vpar;
    const int free_now = n_free(R);

    if (grow) {
        if (nb != NB_EVERYWHERE) grow_candidates(R, k, nb, 0);      // everywhere: every free object is a candidate
        PH(R, 2);
        double ex[2] = {0.0, 0.0};
        posterior_pass(R, k, 0, flat, use_geo, size, free_now, [&](int o) { return nb == NB_EVERYWHERE ? m_cl(R, o) < 0 : R.c->st.flag_b[o] != 0; },
                       [&](int o, bool on, double post) { *val_slot(R, vb, o) = on ? heat(R, post) : 0.0; if (on) ex[0] += 1.0;
                       });
        double total;
        int o = team_pick(R, val, P.N, draw_uniform(R.D, SITE_OBJECT), total, ex, vpeers);
        PH(R, 4);
        R.vpar ^= 1;
        if (ex[0] == 0.0 || size == P.max_size) return pr;
        if (o < 0) return pr;
        double lq_obj = log(*val_slot(R, vb, o) / total);
        pr.changed = 1; pr.M = 1;
        R.chg_mode = CHG_ONE; R.chg_o = o; R.chg_new = k;
        if (sb_tid() == 0) R.c->slist[0] = o;
        pr.S = R.c->slist;
        candidate_sizes(R, k, size + 1);
        double lqs, lqbs;
        resample_sources(R, pr, 1, R.c->slist, k, k, false, lqs, lqbs);
        // back: remove o again, p_remove ~ 1 - posterior over the members of the grown cluster
        double s[2] = {0.0, 0.0};
        posterior_pass(R, k, 1, flat, use_geo, size + 1, size + 1, [&](int j) { return cl_of(R, j, 1) == k; },
                       [&](int j, bool on, double post) { if (on) { double v = 1.0 - heat(R, post); s[0] += v; if (j == o) s[1] = v; } });
        team_sum<2>(R, s);
        PH(R, 13);
        double lqb_obj = log(s[1] / s[0]);
        pr.lq = lq_obj + lqs + log(R.p_grow);
        pr.lqb = lqb_obj + lqbs + log(1.0 - R.p_grow) + extra;
    } else {
        if (size == P.min_size) return pr;
        posterior_pass(R, k, 0, flat, use_geo, size, size, [&](int o) { return m_cl(R, o) == k; },
                       [&](int o, bool on, double post) { *val_slot(R, vb, o) = on ? 1.0 - heat(R, post) : 0.0; });
        double total;
        int o = team_pick(R, val, P.N, draw_uniform(R.D, SITE_OBJECT), total, nullptr, vpeers);
        PH(R, 4);
        R.vpar ^= 1;
        if (o < 0) return pr;
        double lq_obj = log(*val_slot(R, vb, o) / total);
        pr.changed = 1; pr.M = 1;
        R.chg_mode = CHG_ONE; R.chg_o = o; R.chg_new = -1;
        if (sb_tid() == 0) R.c->slist[0] = o;
        pr.S = R.c->slist;
        candidate_sizes(R, k, size - 1);
        double lqs, lqbs;
        resample_sources(R, pr, 1, R.c->slist, k, k, false, lqs, lqbs);
        // back: grow by o again
        if (nb != NB_EVERYWHERE) grow_candidates(R, k, nb, 1);
        double s[3] = {0.0, 0.0, 0.0};
        posterior_pass(R, k, 1, flat, use_geo, size - 1, free_now + 1, [&](int j) { return nb == NB_EVERYWHERE ? cl_of(R, j, 1) < 0 : R.c->st.flag_b[j] != 0; },
                       [&](int j, bool on, double post) { if (on) { double v = heat(R, post); s[0] += v; if (j == o) { s[1] = v; s[2] = 1.0; } } });
        team_sum<3>(R, s);
        PH(R, 13);
        if (s[2] == 0.0) { pr.force_reject = 1; return pr; }   // o is not a grow candidate of the shrunk cluster
        double lqb_obj = log(s[1] / s[0]);
        pr.lq = lq_obj + lqs + log(1.0 - R.p_grow);
        pr.lqb = lqb_obj + lqbs + log(R.p_grow) + extra;
    }
    return pr;
}

// -----------------------------------------------------------------------------------------
// AlterClusterWide (operators.py:1382-1568)
// -----------------------------------------------------------------------------------------
// compute_cluster_probs (operators.py:1399-1418) over the n_av available objects in list_b;
// p[i] is left in val[i] (compact index).  `view` selects old / new state.
SB_DEV void wide_probs(Run &R, double *p, int k, int n_av, int view, bool use_geo, int size_k) {
    const SbProblem &P = *R.P;
    const double eps = 0.01 / (double)P.N;
    int lo, hi; team_range(R, n_av, lo, hi);
    const bool geo = use_geo && P.geo && !R.sfp;
    const bool tables = !R.sfp && pass_uses_tables(R, n_av) && !(R.c->tile && geo);
    const bool tiled = tables && R.c->tile;
    if (tables && !tiled) build_tables(R, k, true);
    auto tail = [&](int i, const Marg &a) {
        double raw = 0.5;
        if (!R.sfp) {
            double m0 = exp(prod_log(a.p0) * R.invT), m1 = exp(prod_log(a.p1) * R.invT);
            if (geo) m1 = m1 * exp(a.geo * R.invTp / SB_GEO_SCALER);
            raw = m1 / (m0 + m1 + SB_EPS32);
        }
        p[i] = raw + SB_EPS32;
    };
    if (tiled) {
        // tiled: per-object values in combo order into the idle val buffer, then into list order
        double *pobj = R.c->st.val[R.vpar ^ 1];
        tiled_pass<TK_WIDE>(R, k, 0, [&](int o) { int c = m_cl(R, o); return c < 0 || c == k; },
            [&](int o, bool on, const Prod &a, const Prod &b) {
                if (!on) return;
                double m0 = exp(prod_log(a) * R.invT), m1 = exp(prod_log(b) * R.invT);
                pobj[o] = m1 / (m0 + m1 + SB_EPS32) + SB_EPS32;
            });
        if (R.tsize > 1) team_sync(R); else sb_cta_sync();
        for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) p[i] = pobj[R.c->st.list_b[i]];
    } else if (tables && !geo && hi - lo >= 32 * sb_nwarps()) {
        // many objects per warp: a lane per object (see posterior_pass)
        for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) {
            Marg a; a.on = 1; a.geo = 0.0;
            obj_marginals_lane(R, R.c->st.list_b[i], a.p0, a.p1);
            tail(i, a);
        }
    } else
    batch_items<Marg>(lo, hi,
        [&](int i) {
            Marg a; a.on = 1; a.geo = 0.0; a.p0 = prod_one(); a.p1 = prod_one();
            if (!R.sfp) {
                int o = R.c->st.list_b[i];
                if (tables) obj_marginals<true>(R, o, k, true, a.p0, a.p1);
                else obj_marginals<false>(R, o, k, true, a.p0, a.p1);
                if (geo) a.geo = geo_delta(R, o, k, view, size_k);
            }
            return a;
        },
        tail);
    sb_cta_sync();
    double s[1];
    s[0] = 0.0; for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) s[0] += p[i];
    team_sum<1>(R, s);
    double s1 = s[0];
    s[0] = 0.0;
    for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) { double v = p[i] / s1 + eps; p[i] = v; s[0] += v; }
    team_sum<1>(R, s);
    double s2 = s[0], old_size = (double)size_k;
    s[0] = 0.0;
    for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) {
        double mem = cl_of(R, R.c->st.list_b[i], view) == k ? 1.0 : 0.0;
        double v = (1.0 - SB_W_STAY) * (p[i] / s2) + SB_W_STAY * (mem / old_size);
        p[i] = v; s[0] += v;
    }
    team_sum<1>(R, s);
    double expected = s[0];
    for (int it = 0; it < 10; it++) {
        s[0] = 0.0;
        for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) {
            double v = p[i] * old_size / expected;
            v = v < eps ? eps : (v > 1.0 - eps ? 1.0 - eps : v);
            p[i] = v; s[0] += v;
        }
        team_sum<1>(R, s);
        expected = s[0];
        if (expected > 0.975 * old_size) break;
    }
}

SB_DEV Prop op_alter_cluster_wide(Run &R, bool use_geo) {
    const SbProblem &P = *R.P;
    Prop pr = prop_init(0, 0.0, -SB_INF);
    int k = draw_index(R.D, SITE_CLUSTER, 0, P.K);
    int size = R.c->rp.sizes[k];
    int n_av = team_compact(R, P.N, R.c->st.list_b, [&](int o) { int c = m_cl(R, o); return c < 0 || c == k; });
    double *p = val_mine(R, R.vpar);      // indexed by position in the list of available objects; every CTA touches its own range only
    wide_probs(R, p, k, n_av, 0, use_geo, size);
    int lo, hi; team_range(R, n_av, lo, hi);
    // redraw until the membership changes (operators.py:1515-1517)
    double st[3];
    for (int attempt = 0;; attempt++) {
        st[0] = 0.0; st[1] = 0.0; st[2] = 0.0;
        for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) {
            int o = R.c->st.list_b[i];
            bool nm = draw_uniform_f32(R.D, SITE_WIDE, (uint32_t)i, attempt) < p[i];
            bool old = m_cl(R, o) == k;
            R.c->st.flag_a[o] = nm;
            st[0] += nm != old ? 1.0 : 0.0; st[1] += nm ? 1.0 : 0.0;
        }
        team_sum<3>(R, st);
        if (st[0] > 0.0) break;
        if (attempt > 100000) { R.err = 4; break; }
    }
    int n_new = (int)st[1];
    if (n_new < P.min_size || n_new > P.max_size) return pr;
    // forward probability
    double a[1] = {0.0}, b[1] = {1.0};
    for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) {
        int o = R.c->st.list_b[i];
        bool nm = R.c->st.flag_a[o], old = m_cl(R, o) == k;
        a[0] += log(nm ? p[i] : 1.0 - p[i]);
        b[0] *= old ? p[i] : 1.0 - p[i];
    }
    team_sum<1>(R, a);
    team_sum<1, true>(R, b);
    double lq = a[0] - log(1.0 - b[0]);
    // the list of flipped objects (object ids, increasing)
    int M = team_compact(R, n_av, R.c->st.list_a, [&](int i) { int o = R.c->st.list_b[i]; return (bool)R.c->st.flag_a[o] != (m_cl(R, o) == k); });
    for (int m = R.rank * sb_nthr() + sb_tid(); m < M; m += R.tsize * sb_nthr()) R.c->st.list_a[m] = R.c->st.list_b[R.c->st.list_a[m]];
    team_sync(R);
    R.chg_mode = CHG_WIDE; R.chg_k = k;
    candidate_sizes(R, k, n_new);
    pr.changed = 1; pr.M = M; pr.S = R.c->st.list_a;
    double lqs, lqbs;
    resample_sources(R, pr, M, R.c->st.list_a, k, k, false, lqs, lqbs);
    // backward probability from the new state
    wide_probs(R, p, k, n_av, 1, use_geo, n_new);
    a[0] = 0.0; b[0] = 1.0;
    for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) {
        int o = R.c->st.list_b[i];
        bool nm = R.c->st.flag_a[o], old = m_cl(R, o) == k;
        a[0] += log(old ? p[i] : 1.0 - p[i]);
        b[0] *= nm ? p[i] : 1.0 - p[i];
    }
    team_sum<1>(R, a);
    team_sum<1, true>(R, b);
    pr.lq = lq + lqs;
    pr.lqb = (a[0] - log(1.0 - b[0])) + lqbs;
    return pr;
}

// -----------------------------------------------------------------------------------------
// ClusterJump (operators.py:1668-1827)
// -----------------------------------------------------------------------------------------
struct JumpAcc { Prod ps, pj; int on; };

SB_DEV Prop op_cluster_jump(Run &R) {
    const SbProblem &P = *R.P;
    Prop pr = prop_init(0, 0.0, -SB_INF);
    int k_src, k_tgt;
    draw_pair(R.D, SITE_JUMP, P.K, k_src, k_tgt);
    if (R.c->rp.sizes[k_src] <= P.min_size || R.c->rp.sizes[k_tgt] >= P.max_size) return pr;
    int lo, hi; team_range(R, P.N, lo, hi);
    double *val = val_mine(R, R.vpar);
    double *const *vpeers = R.xworld > 1 ? R.c->valp[R.vpar] : nullptr;
    const int vb = R.vpar;
    if (R.c->tile && pass_uses_tables(R, R.c->rp.sizes[k_src])) {
        tiled_pass<TK_JUMP>(R, k_src, k_tgt, [&](int o) { return m_cl(R, o) == k_src; },
            [&](int o, bool on, const Prod &a, const Prod &b) { *val_slot(R, vb, o) = on ? jump_tail(R, a, b) : 0.0; });
        if (R.tsize > 1) team_sync(R);
    } else {
        int32_t *list = list_mine(R) + lo;
        const int n = cta_compact(R, lo, hi, list, [&](int o) { return m_cl(R, o) == k_src; }, [&](int o) { val[o] = 0.0; });
        batch_items<JumpAcc>(0, n,
            [&](int i) { JumpAcc a; a.on = 1; jump_marginals(R, list[i], k_src, k_tgt, a.ps, a.pj); return a; },
            [&](int i, const JumpAcc &a) { val[list[i]] = jump_tail(R, a.ps, a.pj); });
    }
    sb_cta_sync();
    PH(R, 18);
    double total;
    int o = team_pick(R, val, P.N, draw_uniform(R.D, SITE_OBJECT), total, nullptr, vpeers);
    PH(R, 4);
    R.vpar ^= 1;
    if (o < 0) return pr;
    double lq_obj = log(*val_slot(R, vb, o) / total);
    pr.changed = 1; pr.M = 1;
    R.chg_mode = CHG_ONE; R.chg_o = o; R.chg_new = k_tgt;
    if (sb_tid() == 0) R.c->slist[0] = o;
    pr.S = R.c->slist;
    candidate_sizes(R, k_src, R.c->rp.sizes[k_src] - 1, k_tgt, R.c->rp.sizes[k_tgt] + 1);
    double lqs, lqbs;
    resample_sources(R, pr, 1, R.c->slist, k_tgt, k_src, true, lqs, lqbs);
    double s[2] = {0.0, 0.0};
    if (R.c->tile && pass_uses_tables(R, R.c->rp.sizes[k_tgt] + 1)) {
        tiled_pass<TK_JUMP>(R, k_tgt, k_src, [&](int j) { return cl_of(R, j, 1) == k_tgt; },
            [&](int j, bool on, const Prod &a, const Prod &b) { if (on) { double v = jump_tail(R, a, b); s[0] += v; if (j == o) s[1] = v; } });
    } else {
        int32_t *list = list_mine(R) + lo;
        const int n = cta_compact(R, lo, hi, list, [&](int j) { return cl_of(R, j, 1) == k_tgt; }, [&](int) {});
        batch_items<JumpAcc>(0, n,
            [&](int i) { JumpAcc a; a.on = 1; jump_marginals(R, list[i], k_tgt, k_src, a.ps, a.pj); return a; },
            [&](int i, const JumpAcc &a) { double v = jump_tail(R, a.ps, a.pj); s[0] += v; if (list[i] == o) s[1] = v; });
    }
    team_sum<2>(R, s);
    PH(R, 19);
    pr.lq = lq_obj + lqs;
    pr.lqb = log(s[1] / s[0]) + lqbs;
    return pr;
}

// -----------------------------------------------------------------------------------------
// GibbsSampleSource (operators.py:423-579)
// -----------------------------------------------------------------------------------------
// calculate_source_posterior (operators.py:554-574) of object o at feature f -> q[c] (normalised)
SB_DEV void source_posterior(const Run &R, int o, int f, int x, int k_o, int pat, int z, double (&q)[SB_UC]) {
    const SbProblem &P = *R.P;
    const int C = SB_CC(P);
    const double *w = R.c->rp.wp + ((size_t)(pat * 2 + z) * P.F + f) * C;
    const bool unitT = R.T == 1.0;
    double tot = 0.0;
    SB_FOR_C(c, C) {
        double l = 1.0;
        if (!R.sfp) {
            int g = c == 0 ? k_o : conf_group(R, c, o);
            l = g < 0 ? 0.0 : pow_inv(entry_P(R, g, f, x), R.invT, unitT);
        }
        q[c] = l * w[c]; tot += q[c];
    }
    SB_FOR_C(c, C) q[c] /= tot;
}

SB_DEV Prop op_gibbs_source(Run &R, bool groups_mode) {
    const SbProblem &P = *R.P;
    const int N = P.N, F = P.F, C = SB_CC(P);
    Prop pr = prop_init(1, 0.0, 0.0);
    // --- select_object_subset (operators.py:461-494) -> S[0..M)
    int M = 0;
    const int *S = R.c->slist;
    sb_cta_sync();
    if (N <= SB_SRC_MIN) {
        M = N;
        if (sb_tid() < N) R.c->slist[sb_tid()] = sb_tid();
        sb_cta_sync();
    } else {
        int pop = N, kmax = SB_SRC_MAX_RANDOM < N ? SB_SRC_MAX_RANDOM : N;
        const int32_t *popl = nullptr;      // sorted population (null: identity)
        bool identity = true;
        if (groups_mode) {
            // components are enumerated [confounders..., clusters] (sbayes/sampling/state.py:626-634)
            int ci = draw_index(R.D, SITE_COMPONENT, 0, C);
            kmax = SB_SRC_MAX_GROUPS < N ? SB_SRC_MAX_GROUPS : N;
            identity = false;
            if (ci == C - 1) {
                int g = draw_index(R.D, SITE_COMPONENT, 1, P.K);
                pop = team_compact(R, N, R.c->st.list_b, [&](int o) { return m_cl(R, o) == g; });
                popl = R.c->st.list_b;
            } else {
                int g = draw_index(R.D, SITE_COMPONENT, 1, P.conf_ng[ci]);
                int row = P.goff[ci + 1] - P.K + g;
                popl = P.cg_idx + P.cg_ptr[row];
                pop = P.cg_ptr[row + 1] - P.cg_ptr[row];
            }
        }
        if (!identity && pop <= kmax) {
            M = pop;
            if (sb_tid() < M) R.c->slist[sb_tid()] = popl[sb_tid()];
        } else {
            M = kmax;
            // uniform k-subset: Floyd's algorithm over ranks, then sort (oracle/draws.py:subset_of);
            // the M Philox draws are taken by M threads, thread 0 only runs the sequential part
            if (!R.D.tv && sb_tid() < M) R.c->sint[sb_tid()] = draw_index(R.D, SITE_SUBSET, (uint32_t)(pop - M + sb_tid()), pop - M + sb_tid() + 1);
            sb_cta_sync();
            if (sb_tid() == 0) {
                int *ch = R.c->slist;
                if (R.D.tv) {
                    for (int j = 0; j < M; j++) ch[j] = (int)R.D.tv[R.D.toff[SITE_SUBSET] + j];   // object ids
                } else {
                    int n = 0;
                    for (int j = pop - M; j < pop; j++) {
                        int t = R.c->sint[j - (pop - M)];
                        bool in = false;
                        for (int q = 0; q < n; q++) in |= ch[q] == t;
                        ch[n++] = in ? j : t;
                    }
                }
            }
            sb_cta_sync();
            // ranks -> object ids and the sort, by M threads: the look-ups are independent loads and
            // the values are distinct, so every thread places its own value at its rank
            int v = 0;
            if (sb_tid() < M) { v = R.c->slist[sb_tid()]; if (!identity && !R.D.tv) v = popl[v]; R.c->sint[32 + sb_tid()] = v; }
            sb_cta_sync();
            if (sb_tid() < M) {
                int rank = 0;
                for (int q = 0; q < M; q++) rank += R.c->sint[32 + q] < v ? 1 : 0;
                R.c->slist[rank] = v;
            }
        }
        sb_cta_sync();
    }
    pr.M = M; pr.S = S;
    R.chg_mode = CHG_NONE;
    for (int i = sb_tid(); i < P.K; i += sb_nthr()) R.c->rp.ssz[i] = R.c->rp.sizes[i];
    const long long items = (long long)M * F;
    PH(R, 14);
    int32_t *Hc = proposal_open(R, M, S);
    PH(R, 6);
    // --- G1: forward conditional from the committed tables, draw.  The (object, feature) items are
    // split over the team (the draws are addressed by item, the candidate sources go to the shared
    // state in HBM); the team sum below is also the barrier that publishes them to every CTA
    int ilo, ihi; team_range(R, (int)items, ilo, ihi);
    double acc[2] = {0.0, 0.0};
    for (long long it = ilo + sb_tid(); it < ihi; it += sb_nthr()) {
        int m = (int)(it / F), f = (int)(it % F), o = S[m];
        int x = sb_ldg(P.x + (size_t)o * P.ldx + f);
        uint8_t *sn = R.c->st.src[m_osel(R, o) ^ 1] + (size_t)o * P.ldx;
        if (x == SB_NA_U8) { sn[f] = SB_NA_U8; continue; }
        int k_o = m_cl(R, o), pat = m_pat(R, o), z = k_o >= 0;
        double q[SB_UC];
        source_posterior(R, o, f, x, k_o, pat, z, q);
        int nw = categorical(q, C, draw_uniform(R.D, SITE_GIBBS, (uint32_t)((size_t)m * F + f)));
        int c_old = R.c->st.src[m_osel(R, o)][(size_t)o * P.ldx + f];
        acc[0] += log(pick_c(q, C, nw));
        const double *lw = R.c->rp.lw + ((size_t)(pat * 2 + z) * F + f) * C;
        if (nw != c_old) acc[1] += lw[nw] - lw[c_old];
        sn[f] = (uint8_t)nw;
    }
    team_sum<2>(R, acc);
    PH(R, 15);
    // --- counts delta on the CTA's own replica (after every forward read of the tables)
    for (long long it = sb_tid(); it < items; it += sb_nthr()) {
        int m = (int)(it / F), f = (int)(it % F), o = S[m];
        int x = sb_ldg(P.x + (size_t)o * P.ldx + f);
        if (x == SB_NA_U8) continue;
        int nw = R.c->st.src[m_osel(R, o) ^ 1][(size_t)o * P.ldx + f], c_old = R.c->st.src[m_osel(R, o)][(size_t)o * P.ldx + f];
        if (nw == c_old) continue;
        int k_o = m_cl(R, o), pat = m_pat(R, o), z = k_o >= 0;
        int g0 = c_old == 0 ? k_o : conf_group(R, c_old, o), g1 = nw == 0 ? k_o : conf_group(R, nw, o);
        if (g0 >= 0) count_add(R, g0, f, x, -1);
        if (g1 >= 0) count_add(R, g1, f, x, 1);
        sb_atomic_add_i32(Hc + ((size_t)(pat * 2 + z) * F + f) * C + c_old, -1);
        sb_atomic_add_i32(Hc + ((size_t)(pat * 2 + z) * F + f) * C + nw, 1);
    }
    sb_cta_sync();
    pr.lq = acc[0];
    pr.sp_new = R.pr_src + acc[1];
    pr.src_changed = 1;
    PH(R, 16);
    proposal_rows(R);
    PH(R, 10);
    pr.geo_new = candidate_geo(R, false);
    // --- G2: backward conditional from the new tables, evaluated at the old sources
    double bq[1] = {0.0};
    for (long long it = ilo + sb_tid(); it < ihi; it += sb_nthr()) {
        int m = (int)(it / F), f = (int)(it % F), o = S[m];
        int x = sb_ldg(P.x + (size_t)o * P.ldx + f);
        if (x == SB_NA_U8) continue;
        int k_o = m_cl(R, o), pat = m_pat(R, o), z = k_o >= 0;
        double q[SB_UC];
        source_posterior(R, o, f, x, k_o, pat, z, q);
        bq[0] += log(pick_c(q, C, R.c->st.src[m_osel(R, o)][(size_t)o * P.ldx + f]));
    }
    team_sum<1>(R, bq);
    PH(R, 17);
    pr.lqb = bq[0];
    pr.need_sync = 1;
    return pr;
}

// -----------------------------------------------------------------------------------------
// GibbsSampleWeights (operators.py:582-693).  Edits W in place, always accepted.  CTA-local.
// -----------------------------------------------------------------------------------------
SB_DEV double beta_logpdf_nolnB(double x, double a, double b) {   // scipy beta.logpdf without the -betaln term
    double t1 = a == 1.0 ? 0.0 : (a - 1.0) * log(x);
    double t2 = b == 1.0 ? 0.0 : (b - 1.0) * log1p(-x);
    return t1 + t2;
}
SB_DEV Prop op_gibbs_weights(Run &R) {
    const SbProblem &P = *R.P;
    const int F = P.F, C = SB_CC(P);
    Prop pr = prop_init(1, -SB_INF, 0.0);
    int i1, i2;
    draw_pair(R.D, SITE_WPAIR, C, i1, i2);
    R.chg_mode = CHG_NONE;
    const int32_t *H = R.c->rp.H[R.hcur];
    for (int f = sb_tid(); f < F; f += sb_nthr()) {
        // counts of sources over the objects that have both components (operators.py:648-656)
        double c1 = sb_ldg(P.alpha_w + (size_t)f * C + i1), c2 = sb_ldg(P.alpha_w + (size_t)f * C + i2);
        for (int pz = 0; pz < P.n_pat * 2; pz++) {
            int bits = P.pat_bits[pz >> 1], z = pz & 1;
            bool h1 = i1 == 0 ? z : (bits >> i1) & 1, h2 = i2 == 0 ? z : (bits >> i2) & 1;
            if (h1 && h2) {
                const int32_t *h = H + ((size_t)pz * F + f) * C;
                c1 += (double)h[i1]; c2 += (double)h[i2];
            }
        }
        double a = 1.0 + c2 / R.Tp, b = 1.0 + c1 / R.Tp;
        double a2 = draw_beta(R.D, a, b, (uint32_t)f);
        // float32 arithmetic of resample_weight_for_two_components (operators.py:664-672)
        float Wo[SB_UC], Wn[SB_UC], wo1 = 0.f, wo2 = 0.f;
        SB_FOR_C(c, C) { Wo[c] = R.c->rp.W[f * C + c]; Wn[c] = Wo[c]; if (c == i1) wo1 = Wo[c]; if (c == i2) wo2 = Wo[c]; }
        float w02 = wo1 + wo2;
        SB_FOR_C(c, C) { if (c == i1) Wn[c] = (float)((1.0 - a2) * (double)w02); if (c == i2) Wn[c] = (float)(a2 * (double)w02); }
        float s = 0.0f;
        SB_FOR_C(c, C) s = s + Wn[c];
        SB_FOR_C(c, C) Wn[c] = Wn[c] / s;
        double a2_old = (double)(wo2 / w02);
        double lq = beta_logpdf_nolnB(a2, a, b), lqb = beta_logpdf_nolnB(a2_old, a, b);
        // per-feature target: source likelihood + pointwise weights prior (operators.py:617-623,678-686)
        double lp_old = 0.0, lp_new = 0.0;
        for (int pz = 0; pz < P.n_pat * 2; pz++) {
            int bits = P.pat_bits[pz >> 1], z = pz & 1;
            const int32_t *h = H + ((size_t)pz * F + f) * C;
            const double *lwo = R.c->rp.lw + ((size_t)pz * F + f) * C;
            double tot = 0.0;
            SB_FOR_C(c, C) { double hc = c == 0 ? (double)z : (double)((bits >> c) & 1); tot += hc * (double)Wn[c]; }
            SB_FOR_C(c, C) {
                if (h[c] != 0) {
                    double hc = c == 0 ? (double)z : (double)((bits >> c) & 1);
                    lp_old += (double)h[c] * lwo[c];
                    lp_new += (double)h[c] * log(hc * (double)Wn[c] / tot);
                }
            }
        }
        lp_old += weights_prior_row_reg(P, Wo, f);
        lp_new += weights_prior_row_reg(P, Wn, f);
        double p_acc = exp((lp_new - lp_old + lqb - lq) / R.Tp);
        bool acc = draw_uniform_f32(R.D, SITE_WACCEPT, (uint32_t)f, 0) < p_acc;
        if (acc) SB_FOR_C(c, C) R.c->rp.W[f * C + c] = Wn[c];
    }
    sb_cta_sync();
    PH(R, 22);
    weight_tables(R);
    PH(R, 23);
    pr.sp_new = source_prior_from_hist(R, H);
    pr.wp_new = weights_prior_total(R);
    PH(R, 24);
    return pr;
}

// -----------------------------------------------------------------------------------------
// full (uncached) evaluation of a state: count rebuild + histogram (distributed, HBM
// atomics) + per-(group, feature) terms + likelihood + priors (per replica).
// recalculate_feature_counts (sbayes/sampling/counts.py:35-52),
// Likelihood.__call__ (likelihood.py:47-63), Prior.__call__ (prior.py:74-88).
// -----------------------------------------------------------------------------------------
SB_DEV void rebuild_state(Run &R) {
    const SbProblem &P = *R.P;
    const int N = P.N, F = P.F, S = P.S, C = SB_CC(P);
    const size_t nG = (size_t)P.G_tot * F * S;
    const int nH = P.n_pat * 2 * F * C;
    // where the bins are accumulated and where the replicas are filled from: the chain's statistics block
    // (counts | sizes | H), or -- objects sharded over GPUs -- this GPU's partial block and, after the
    // exchange below, its copy of the reduced block
    int32_t *acc_counts = R.c->st.counts, *acc_sizes = R.c->st.sizes, *acc_H = R.c->st.H;
    const int32_t *fin_counts = acc_counts, *fin_sizes = acc_sizes, *fin_H = acc_H;
    if (R.xworld > 1) {
        acc_counts = R.c->lstatp[R.xrank]; acc_sizes = acc_counts + nG; acc_H = acc_sizes + P.K;
        fin_counts = R.c->gstatp[R.xrank]; fin_sizes = fin_counts + nG; fin_H = fin_sizes + P.K;
    }
    {
        // zeroed by the CTAs that will add to it (all of the team / the CTAs of this GPU)
        const int zs = (R.xworld > 1 ? R.lsize : R.tsize) * sb_nthr(), z0 = (R.xworld > 1 ? R.lrank : R.rank) * sb_nthr() + sb_tid();
        for (size_t i = z0; i < nG; i += zs) acc_counts[i] = 0;
        for (int i = z0; i < nH; i += zs) acc_H[i] = 0;
        for (int k = z0; k < P.K; k += zs) acc_sizes[k] = 0;
    }
    team_sync(R);
    // histogram pass: warp per object, lanes over features (coalesced byte reads)
    int lo, hi; team_range(R, N, lo, hi);
    for (int o = lo + sb_warp(); o < hi; o += sb_nwarps()) {
        int k = m_cl(R, o), pat = m_pat(R, o), z = k >= 0;
        int gC[SB_UC]; gC[0] = k;
        SB_FOR_C1(c, C) gC[c] = conf_group(R, c, o);
        const uint8_t *xr = P.x + (size_t)o * P.ldx;
        const uint8_t *so = R.c->st.src[m_osel(R, o)] + (size_t)o * P.ldx;
        for (int f = sb_lane(); f < F; f += 32) {
            int x = sb_ldg(xr + f);
            if (x == SB_NA_U8) continue;
            int c = so[f];
            if (c >= C) { R.err = 2; continue; }
            int g = -1;
            SB_FOR_C(cc, C) if (cc == c) g = gC[cc];
            if (g >= 0) sb_atomic_add_i32(acc_counts + ((size_t)g * F + f) * S + x, 1);
            sb_atomic_add_i32(acc_H + ((size_t)(pat * 2 + z) * F + f) * C + c, 1);
        }
        if (sb_lane() == 0 && k >= 0) sb_atomic_add_i32(acc_sizes + k, 1);
    }
    team_sync(R);
    if (R.xworld > 1) {
        // all-reduce of the partial blocks through peer memory (reduce-scatter + all-gather in one sweep):
        // the team's threads split the bins; a thread sums its bins over the ranks' partial blocks and stores
        // the totals into every rank's copy of the reduced block
        const size_t n_stat = nG + (size_t)P.K + (size_t)nH;
        const size_t stride = (size_t)R.tsize * sb_nthr();
        for (size_t i = (size_t)R.rank * sb_nthr() + sb_tid(); i < n_stat; i += stride) {
            int32_t v = 0;
            for (int p = 0; p < R.xworld; p++) v += R.c->lstatp[p][i];
            for (int p = 0; p < R.xworld; p++) R.c->gstatp[p][i] = v;
        }
        team_sync(R);
    }
    // every CTA fills its replica
    for (size_t i = sb_tid(); i < nG; i += sb_nthr()) { R.c->rp.counts[i] = fin_counts[i]; R.c->rp.dcnt[i] = 0; }
    for (int i = sb_tid(); i < nH; i += sb_nthr()) R.c->rp.H[0][i] = fin_H[i];
    for (int i = sb_tid(); i < F * C; i += sb_nthr()) R.c->rp.W[i] = R.c->st.W[i];
    for (int k = sb_tid(); k < P.K; k += sb_nthr()) { R.c->rp.sizes[k] = fin_sizes[k]; R.c->rp.ssz[k] = fin_sizes[k]; }
    for (int g = sb_tid(); g < P.G_tot; g += sb_nthr()) R.c->rp.dirty[g] = 0;
    R.hcur = 0; R.chg_mode = CHG_NONE; R.nd = 0;
    sb_cta_sync();
    rows_full(R, nullptr, P.G_tot);
    rows_lh(R, nullptr, P.G_tot);
    weight_tables(R);
    R.ll = R.sfp ? 0.0 : ll_total(R);
    R.pr_src = source_prior_from_hist(R, R.c->rp.H[0]);
    R.pr_w = weights_prior_total(R);
    R.pr_size = size_prior_cta(R, R.c->rp.sizes);
    R.pr_geo = 0.0;
    if (P.geo) {
        for (int k = 0; k < P.K; k++) {
            double a, g; geo_cluster(R, k, 0, a, g);
            if (sb_tid() == 0) { R.c->rp.geo_agg[k] = a; R.c->rp.geo_val[k] = g; }
            R.pr_geo += g;
        }
        sb_cta_sync();
    }
    team_sync(R);      // the rebuild scratch (st.counts / st.H) may be reused only after everyone copied it
}

// -----------------------------------------------------------------------------------------
// one MCMC step (MCMCChain.step, sbayes/sampling/mcmc_chain.py:186-238)
// -----------------------------------------------------------------------------------------
SB_DEV void mcmc_step(Run &R, sb_step_record *rec) {
    const SbProblem &P = *R.P;
    const long long t_begin = sb_clock();
    R.i_step++;
    R.D.step = (unsigned long long)R.i_step;
    R.chg_mode = CHG_NONE;
    // operator choice (mcmc_chain.py:179-184)
    int op;
    if (R.D.tv) op = (int)R.D.tv[R.D.toff[SITE_OP]];
    else {
        double u = draw_uniform(R.D, SITE_OP), cum = 0.0, tot = 0.0;
        for (int i = 0; i < SB_N_OPS; i++) tot += R.c->opw[i];
        op = SB_N_OPS - 1;
        for (int i = 0; i < SB_N_OPS; i++) { cum += R.c->opw[i]; if (cum / tot > u) { op = i; break; } }
    }
    Prop pr;
    const bool geo = P.geo != 0;
    R.cur_op = op; R.pt = t_begin;
    PH(R, 1);
    if (op <= OP_GIBBSISH_GEO) {
        // the five AlterCluster variants of the schedule (operators.py:50-124)
        const bool gibbsish = op >= OP_GIBBSISH;
        const bool use_geo = geo && (op == OP_NAIVE_N1_GEO || op == OP_NAIVE_N2_GEO || op == OP_GIBBSISH_GEO);
        const int nb = op == OP_NAIVE_N2_GEO ? NB_TWOSTEP : (gibbsish ? NB_EVERYWHERE : NB_DIRECT);
        pr = op_alter_cluster(R, gibbsish, use_geo, nb);
    } else if (op == OP_WIDE) pr = op_alter_cluster_wide(R, geo);
    else if (op == OP_JUMP) pr = op_cluster_jump(R);
    else if (op == OP_SRC_RANDOM || op == OP_SRC_GROUPS) pr = op_gibbs_source(R, op == OP_SRC_GROUPS);
    else pr = op_gibbs_weights(R);
    PH(R, 29);
    // candidate posterior (mcmc_chain.py:201-202)
    double ll_new = R.ll, p_size = R.pr_size, p_geo = R.pr_geo, p_w = R.pr_w, p_src = R.pr_src;
    const bool touched = pr.changed != 0;
    if (touched && op == OP_WEIGHTS) { p_src = pr.sp_new; p_w = pr.wp_new; }
    else if (touched) {
        ll_new = R.sfp ? 0.0 : ll_total(R);
        p_src = pr.sp_new;
        if (P.geo) p_geo = pr.geo_new;
        p_size = size_prior_cta(R, R.c->rp.ssz);
    }
    double prior_new = p_size + p_geo + p_w + p_src;
    double lq = pr.lq, lqb = pr.lqb;
    if (pr.force_reject) { lq = 0.0; lqb = -SB_INF; }
    // Metropolis-Hastings (mcmc_chain.py:204-217,240-255)
    bool accept; double ratio = NAN, log_u = NAN;
    if (!touched || lqb == -SB_INF) accept = false;
    else if (lq == -SB_INF) accept = true;
    else if (R.propose_only) accept = true;
    else {
        double prior_old = R.pr_size + R.pr_geo + R.pr_w + R.pr_src;
        ratio = (ll_new / R.T + prior_new / R.Tp) - (R.ll / R.T + prior_old / R.Tp) - (lq - lqb);
        double u = draw_uniform(R.D, SITE_MH);
        log_u = u > 0.0 ? log(u) : -SB_INF;
        accept = log_u < ratio;
    }
    if (rec && R.rank == 0 && sb_tid() == 0) {
        rec->op = op; rec->accept = accept; rec->changed = touched && !pr.force_reject; rec->n_changed = pr.M;
        rec->log_q = lq; rec->log_q_back = lqb;
        rec->ll_new = (touched && !pr.force_reject) ? ll_new : R.ll;
        rec->prior_new = (touched && !pr.force_reject) ? prior_new : R.pr_size + R.pr_geo + R.pr_w + R.pr_src;
        rec->mh_ratio = ratio; rec->log_u = log_u;
    }
    PH(R, 30);
    if (touched && op != OP_WEIGHTS) {
        if (pr.need_sync && accept) team_sync(R);     // a fast CTA's commit must not overtake a slow CTA's reads
        if (accept) {
            // object-indexed state in HBM: every CTA writes the same values
            // (objects sharded over GPUs: the master copy in rank 0's memory is written once, by the first CTA
            // of the team; every GPU keeps its own replica for the passes)
            const bool master = R.xworld == 1 || R.rank == 0;
            for (int m = sb_tid(); m < pr.M; m += sb_nthr()) {
                int o = pr.S[m];
                int nc = cl_of(R, o, 1);
                if (pr.src_changed) {
                    const uint8_t sel = R.c->rp.selbuf[m];
                    if (master) R.c->st.osel[o] = sel;
                    if (R.c->mosel) R.c->mosel[o] = sel;
                    if (R.c->losel) R.c->losel[o] = sel;
                }
                if (master) R.c->st.cluster_of[o] = nc;
                if (R.c->mcl) R.c->mcl[o] = (int16_t)nc;
                if (R.c->lcl) R.c->lcl[o] = nc;
            }
            for (int k = sb_tid(); k < P.K; k += sb_nthr()) R.c->rp.sizes[k] = R.c->rp.ssz[k];
            if (pr.src_changed) R.hcur ^= 1;
        }
        proposal_close(R, accept);
    }
    if (accept) {
        R.ll = ll_new; R.pr_size = p_size; R.pr_geo = p_geo; R.pr_w = p_w; R.pr_src = p_src;
        if (!(isfinite(R.ll) && isfinite(prior_new))) R.err = 1;     // mcmc_chain.py:231-236
    }
    PH(R, 31);
#if defined(SB_PHASE_PROF) && !defined(SB_HOSTSIM)
    if (R.rank == 0 && sb_tid() == 0) g_phase_n[op]++;
#endif
    if (R.rank == 0 && sb_tid() == 0) {
        if (accept) R.c->slot->accepts[op]++; else R.c->slot->rejects[op]++;
        R.c->slot->cycles[op] += sb_clock() - t_begin;
    }
}

// -----------------------------------------------------------------------------------------
// the persistent kernel: every team advances its chains by n_steps
// -----------------------------------------------------------------------------------------
SB_DEV void setup_replica(Run &R, const SbLaunch &L, char *smem, char *glob) {
    const SbProblem &P = L.P;
    const SbPlace &pl = L.place;
    FixedSmem fx = sb_fixed_smem(P.G_tot, P.K, P.F, SB_CC(P), L.smem_members, pl.fixed_off);
    RunC *c = (RunC *)(smem + fx.runc);
    R.c = c;
    int lo, hi; team_range(R, P.N, lo, hi);
    if (sb_tid() == 0) {
        auto at = [&](int a) -> char * { return pl.in_smem[a] ? smem + pl.smem_off[a] : glob + pl.glob_off[a]; };
        Rep &rp = c->rp;
        rp.tab = pl.tiled ? nullptr : (SbD2 *)at(A_TAB0);
        rp.rs = pl.tiled ? (double *)(glob + pl.glob_off[A_RS]) : nullptr;
        c->tile = pl.tiled ? (double *)(smem + pl.smem_off[A_TAB0]) : nullptr;
        c->tile_cap = pl.tiled ? (int)pl.tile_doubles : 0;
        rp.wq = (double *)at(A_WQ); rp.lw = (double *)at(A_LW);
        rp.wp = pl.alias_w ? rp.wq : (double *)at(A_WP); rp.w1 = pl.alias_w ? rp.wq : (double *)at(A_W1);
        c->alpha = pl.in_smem[A_ALPHA] ? (const double *)(smem + pl.smem_off[A_ALPHA]) : P.alpha;
        c->asum = pl.in_smem[A_ASUM] ? (const double *)(smem + pl.smem_off[A_ASUM]) : P.asum;
        c->applm = pl.in_smem[A_APPLM] ? (const uint64_t *)(smem + pl.smem_off[A_APPLM]) : P.applm;
        c->combo_rows = pl.in_smem[A_COMBOT] ? (const int32_t *)(smem + pl.smem_off[A_COMBOT]) : P.combo_rows;
        c->combo_pat = pl.in_smem[A_COMBOT] ? (const int32_t *)(smem + pl.smem_off[A_COMBOT]) + (size_t)P.n_combos * P.n_conf : P.combo_pat;
        c->mcl = pl.in_smem[A_MCL] ? (int16_t *)(smem + pl.smem_off[A_MCL]) : nullptr;
        c->mcombo = pl.in_smem[A_MCOMBO] ? (uint16_t *)(smem + pl.smem_off[A_MCOMBO]) : nullptr;
        c->mosel = pl.in_smem[A_MOSEL] ? (uint8_t *)(smem + pl.smem_off[A_MOSEL]) : nullptr;
        rp.counts = (int32_t *)at(A_COUNTS); rp.nn = (int32_t *)at(A_NN); rp.dcnt = (int32_t *)at(A_DCNT);
        rp.rT = (double *)at(A_RT); rp.rP = (double *)at(A_RP); rp.bsT = (double *)at(A_BST);
        rp.llgf = (double *)at(A_LLGF);
        rp.bk_llgf = (double *)at(A_BK_LLGF); rp.bk_rP = (double *)at(A_BK_RP); rp.bk_rT = (double *)at(A_BK_RT);
        rp.H[0] = (int32_t *)at(A_H0); rp.H[1] = (int32_t *)at(A_H1);
        rp.selbuf = (uint8_t *)(glob + pl.sel_off);
        c->sred = (double *)(smem + fx.sred); c->smem_best = (double *)(smem + fx.best);
        rp.lh = (double *)(smem + fx.lh); rp.bk_lh = (double *)(smem + fx.bk_lh);
        double *geo = (double *)(smem + fx.geo);
        rp.geo_agg = geo; rp.geo_val = geo + P.K; rp.c_geo_agg = geo + 2 * P.K; rp.c_geo_val = geo + 3 * P.K;
        c->sint = (int *)(smem + fx.sint); c->slist = (int *)(smem + fx.slist);
        rp.sizes = (int32_t *)(smem + fx.sizes); rp.ssz = (int32_t *)(smem + fx.ssz);
        c->smem_mem = (int *)(smem + fx.mem);
        rp.W = (float *)(smem + fx.W); rp.dlist = (int *)(smem + fx.dlist); rp.dirty = (uint8_t *)(smem + fx.dirty);
        c->smem_cap = L.smem_members;
        const bool xb = pl.xblk_rows > 0 && pl.in_smem[A_XBLK];
        rp.xblk = xb ? (const uint8_t *)at(A_XBLK) : nullptr; rp.xlo = xb ? lo : 0; rp.xhi = xb ? hi : 0;
    }
    sb_cta_sync();
    // stage the problem constants that were given room in shared memory
    if (pl.in_smem[A_ALPHA]) { double *d = (double *)c->alpha; for (size_t i = sb_tid(); i < (size_t)P.A_rows * P.F * P.S; i += sb_nthr()) d[i] = sb_ldg(P.alpha + i); }
    if (pl.in_smem[A_ASUM]) { double *d = (double *)c->asum; for (size_t i = sb_tid(); i < (size_t)P.A_rows * P.F; i += sb_nthr()) d[i] = sb_ldg(P.asum + i); }
    if (pl.in_smem[A_APPLM]) { uint64_t *d = (uint64_t *)c->applm; for (int i = sb_tid(); i < P.F; i += sb_nthr()) d[i] = sb_ldg(P.applm + i); }
    if (pl.in_smem[A_COMBOT]) {
        int32_t *d = (int32_t *)c->combo_rows;
        const int nr = P.n_combos * P.n_conf;
        for (int i = sb_tid(); i < nr; i += sb_nthr()) d[i] = sb_ldg(P.combo_rows + i);
        for (int i = sb_tid(); i < P.n_combos; i += sb_nthr()) d[nr + i] = sb_ldg(P.combo_pat + i);
    }
    if (c->mcombo) for (int o = sb_tid(); o < P.N; o += sb_nthr()) c->mcombo[o] = (uint16_t)sb_ldg(P.combo_id + o);
    sb_cta_sync();
    // the CTA's own block of observation rows
    if (c->rp.xblk) {
        uint32_t *dst = (uint32_t *)c->rp.xblk;
        const uint32_t *srcw = (const uint32_t *)(P.x + (size_t)lo * P.ldx);
        const int nwords = (hi - lo) * (P.ldx >> 2);
        for (int i = sb_tid(); i < nwords; i += sb_nthr()) dst[i] = srcw[i];
        sb_cta_sync();
    }
}

SB_KERNEL_BIG sb_run_kernel(const SB_GRID_CONSTANT SbLaunch L) {
    Run R;
    R.P = &L.P;
    // CTAs of this launch: [cta_base, cta_base + lgrid) of the grid (the whole grid except in emulated multi-rank launches)
    const int lcta = sb_cta() - L.cta_base, lncta = L.lgrid > 0 ? L.lgrid : sb_ncta();
    R.xworld = L.xworld > 1 ? L.xworld : 1; R.xrank = R.xworld > 1 ? L.xrank : 0; R.xbase = L.xepoch_base;
    R.lsize = L.team; R.lrank = lcta % L.team;
    R.tsize = L.team * R.xworld; R.rank = R.xrank * L.team + R.lrank;
    R.par = 0; R.vpar = 0; R.spar = 0; R.epoch = 0; R.use_cluster = L.use_cluster;
    R.hcur = 0; R.nd = 0; R.pt = 0; R.cur_op = 0;
    const int team_id = lcta / L.team, n_teams = lncta / L.team;
    setup_replica(R, L, sb_smem(), L.rep_base + (size_t)lcta * L.rep_stride);
    if (sb_tid() == 0) {
        const size_t tred_off = (size_t)team_id * 2 * L.team * SB_MAX_RANKS * SB_REDW;
        R.c->bar = L.bars + team_id;
        R.c->tred = L.tred + tred_off;
        R.c->lcl = nullptr; R.c->losel = nullptr; R.c->llist = nullptr; R.c->xabort = L.xabort; R.c->lrel = L.lrelease ? L.lrelease + team_id : nullptr;
        for (int p = 0; p < SB_MAX_RANKS; p++) {
            R.c->tredp[p] = p < R.xworld && L.tred_peer[p] ? L.tred_peer[p] + tred_off : nullptr;
            R.c->xflagp[p] = p < R.xworld && L.xflag_peer[p] ? L.xflag_peer[p] + (size_t)team_id * SB_MAX_RANKS : nullptr;
            R.c->lstatp[p] = nullptr; R.c->gstatp[p] = nullptr;
        }
    }
    for (int s = team_id; s < L.n_slots; s += n_teams) {
        SbSlot *slot = L.slots + s;
        if (!slot->active) continue;
        sb_cta_sync();
        if (sb_tid() == 0) {
            R.c->slot = slot;
            R.c->st = L.states[slot->state];
            for (int i = 0; i < SB_N_OPS; i++) R.c->opw[i] = slot->cfg.op_weights[i];
            if (R.xworld > 1) {
                R.c->lcl = L.lcl + (size_t)slot->state * L.P.N; R.c->losel = L.losel + (size_t)slot->state * L.P.N;
                for (int p = 0; p < R.xworld; p++) {
                    R.c->lstatp[p] = L.lstat_peer[p] + (size_t)slot->state * L.stats_stride;
                    R.c->gstatp[p] = L.gstat_peer[p] + (size_t)slot->state * L.stats_stride;
                    for (int b = 0; b < 2; b++) R.c->valp[b][p] = (double *)(L.arena_peer[p] + ((char *)R.c->st.val[b] - L.arena_peer[0]));
                }
                R.c->llist = (int32_t *)(L.arena_peer[R.xrank] + ((char *)R.c->st.list_a - L.arena_peer[0]));
            }
        }
        sb_cta_sync();
        if (R.c->mcl) for (int o = sb_tid(); o < L.P.N; o += sb_nthr()) R.c->mcl[o] = (int16_t)R.c->st.cluster_of[o];
        if (R.c->mosel) for (int o = sb_tid(); o < L.P.N; o += sb_nthr()) R.c->mosel[o] = R.c->st.osel[o];
        if (R.c->lcl) {
            // this GPU's replica of the membership / committed-copy selector (read from the master in rank 0's
            // memory by the CTAs of this GPU, each its share; the first team barrier of rebuild_state publishes it)
            for (int o = R.lrank * sb_nthr() + sb_tid(); o < L.P.N; o += R.lsize * sb_nthr()) { R.c->lcl[o] = R.c->st.cluster_of[o]; R.c->losel[o] = R.c->st.osel[o]; }
        }
        sb_cta_sync();
        const sb_chain_cfg &cfg = slot->cfg;
        R.T = cfg.temperature; R.Tp = cfg.prior_temperature; R.invT = 1.0 / R.T; R.invTp = 1.0 / R.Tp;
        R.p_grow = cfg.p_grow; R.sfp = cfg.sample_from_prior; R.propose_only = cfg.flags & SB_CFG_PROPOSE_ONLY;
        R.D.k0 = (uint32_t)(cfg.seed & 0xFFFFFFFFull); R.D.k1 = (uint32_t)(cfg.seed >> 32);
        R.D.tv = L.tape.vals; R.D.toff = nullptr; R.D.tlen = nullptr; R.D.wide_n = 0; R.D.step = 0;
        R.err = 0;
        R.i_step = (long long)R.c->st.post[5];
        rebuild_state(R);
        for (long long it = 0; it < L.n_steps; it++) {
            if (L.tape.vals) {
                R.D.toff = L.tape.off + it * SB_N_SITES; R.D.tlen = L.tape.len + it * SB_N_SITES;
                R.D.wide_n = L.tape.wide_rows ? L.tape.wide_rows[it] : 0;
            }
            mcmc_step(R, L.records ? L.records + (size_t)s * L.n_steps + it : nullptr);
        }
        team_sync(R);
        if (R.rank == 0) {
            // export copies of the replica for the host
            const size_t nG = (size_t)L.P.G_tot * L.P.F * L.P.S;
            for (size_t i = sb_tid(); i < nG; i += sb_nthr()) R.c->st.counts[i] = R.c->rp.counts[i];
            for (int i = sb_tid(); i < L.P.F * SB_CC(L.P); i += sb_nthr()) R.c->st.W[i] = R.c->rp.W[i];
            for (int g = sb_tid(); g < L.P.G_tot; g += sb_nthr()) R.c->st.glh[g] = R.sfp ? 0.0 : R.c->rp.lh[g];
            for (int k = sb_tid(); k < L.P.K; k += sb_nthr()) R.c->st.geo_out[k] = L.P.geo ? R.c->rp.geo_val[k] : 0.0;
            if (sb_tid() == 0) {
                R.c->st.post[0] = R.ll; R.c->st.post[1] = R.pr_size; R.c->st.post[2] = R.pr_geo; R.c->st.post[3] = R.pr_w;
                R.c->st.post[4] = R.pr_src; R.c->st.post[5] = (double)R.i_step;
            }
        }
        // an error seen by ANY CTA of the team (e.g. a bad source byte in its block of objects) is reported
        if (R.err && sb_tid() == 0) slot->err = R.err;
        team_sync(R);
    }
}

// Batched sample export (sb_export_samples): the committed state of n chains gathered into staging buffers
// in the host layout -- source rows unpadded and taken from whichever copy is committed -- plus the
// per-object source prior (SourcePrior's cache node, sbayes/model/prior.py:596-611:
// sum_f log w~[o, f, src[o,f]] with w~ = has * W / sum(has * W), normalize_weights likelihood.py:171-190).
// grid = (blocks, n); one warp per object for the rows.
SB_KERNEL sb_export_samples_kernel(SbProblem P, const SbShared *states, SbExport E) {
    const int N = P.N, F = P.F, C = SB_CC(P);
#ifdef SB_HOSTSIM
    const int j = sb_cta() % E.n, blk = sb_cta() / E.n, nblk = sb_ncta() / E.n;
#else
    const int j = blockIdx.y, blk = blockIdx.x, nblk = gridDim.x;
#endif
    const SbShared st = states[E.state_idx[j]];
    const int t0 = blk * sb_nthr() + sb_tid(), stride = nblk * sb_nthr();
    const size_t nG = (size_t)P.G_tot * F * P.S;
    if (E.want & 1u) for (int o = t0; o < N; o += stride) E.cluster_of[(size_t)j * N + o] = st.cluster_of[o];
    if (E.want & 2u) for (int i = t0; i < F * C; i += stride) E.W[(size_t)j * F * C + i] = st.W[i];
    if (E.want & 8u) for (size_t i = t0; i < nG; i += stride) E.counts[(size_t)j * nG + i] = st.counts[i];
    if (E.want & 16u) for (int i = t0; i < 8; i += stride) E.post[(size_t)j * 8 + i] = st.post[i];
    if (E.want & 32u) for (int g = t0; g < P.G_tot; g += stride) E.glh[(size_t)j * P.G_tot + g] = st.glh[g];
    if (E.want & 64u) for (int k = t0; k < P.K; k += stride) E.geo[(size_t)j * P.K + k] = st.geo_out[k];
    if (!(E.want & (4u | 128u))) return;
    const int wpb = sb_nwarps();
    for (int o = blk * wpb + sb_warp(); o < N; o += nblk * wpb) {
        const uint8_t *row = st.src[st.osel[o]] + (size_t)o * P.ldx;
        const int bits = sb_ldg(P.pat_bits + sb_ldg(P.pat_id + o)), z = st.cluster_of[o] >= 0;
        double acc = 0.0;
        for (int f = sb_lane(); f < F; f += 32) {
            const int c = row[f];
            if (E.want & 4u) E.src[((size_t)j * N + o) * F + f] = (uint8_t)c;
            if ((E.want & 128u) && c != SB_NA_U8) {
                double t1 = 0.0, wc = 0.0;
                for (int cc = 0; cc < C; cc++) {
                    const double hw = (cc == 0 ? (double)z : (double)((bits >> cc) & 1)) * (double)st.W[f * C + cc];
                    t1 += hw; if (cc == c) wc = hw;
                }
                acc += log(wc / t1);
            }
        }
        if (E.want & 128u) {
            acc = warp_sum(acc);
            if (sb_lane() == 0) E.sp_obj[(size_t)j * N + o] = acc;
        }
    }
}

// export of the committed source rows (they sit in either copy)
SB_KERNEL sb_export_kernel(SbProblem P, SbShared st, uint8_t *src_out) {
    const size_t nS = (size_t)P.N * P.ldx;
    const size_t t0 = (size_t)sb_cta() * sb_nthr() + sb_tid(), stride = (size_t)sb_ncta() * sb_nthr();
    for (size_t i = t0; i < nS; i += stride) { int o = (int)(i / P.ldx); src_out[i] = st.src[st.osel[o]][i]; }
}
}  // namespace SB_NS
