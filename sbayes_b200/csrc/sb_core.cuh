// sb_core.cuh -- device code of the sBayes MCMC step: team primitives, draws, tables,
// the ten scheduled operators, Metropolis-Hastings accept/commit.
//
// Execution model.  One chain is advanced by a *team* of `team` CTAs of one persistent
// kernel; every thread of the team executes the same step function.  All random draws
// are counter-based (Philox4x32-10 addressed by chain seed, step, site, sub-stream,
// index), so every thread can derive any draw locally and no draw is ever broadcast.
// Anything that depends on a reduction over objects goes through `team_sum` /
// `team_pick` / `team_compact`, which end in a team barrier and leave the *same* value in
// every thread, so all control flow below is uniform across the team.
//
// What follows which reference code is cited function by function; the algorithmic
// specification is SURVEY.md Appendix A, the executable one is oracle/operators.py.
#pragma once
#include "sb_platform.cuh"
#include "sb_types.h"

#define SB_EPS32 1.1920928955078125e-07   /* np.finfo(np.float32).eps, sbayes/util.py:34 */
#define SB_SMOOTH 1e-6                    /* additive smoothing, sbayes/sampling/operators.py:978 */
#define SB_W_STAY 0.15                    /* operators.py:133 */
#define SB_GEO_SCALER 2.0                 /* operators.py:1390 */
#define SB_SRC_MIN 10                     /* operators.py:440 */
#define SB_SRC_MAX_RANDOM 20              /* operators.py:185 */
#define SB_SRC_MAX_GROUPS 30              /* operators.py:194 */

#ifdef SB_HOSTSIM
#define SB_INF (__builtin_inf())
#else
#define SB_INF (__longlong_as_double(0x7ff0000000000000LL))
#endif

// scalars published in st.sc[] during a step
enum { SC_LQ = 0, SC_LQB, SC_SP_NEW, SC_GEO_NEW, SC_WP_NEW, SC_AUX0, SC_AUX1, SC_AUX2, SC_AUX3 };

// -----------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011); stream addressing identical to oracle/philox.py
// -----------------------------------------------------------------------------------------
struct Philox4 { uint32_t w[4]; };

SB_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Philox4 o; o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3; return o;
}
SB_HD double u01_double(uint32_t w0, uint32_t w1) {
    return ((double)(w0 >> 5) * 67108864.0 + (double)(w1 >> 6)) / 9007199254740992.0;
}
SB_DEV double u01_float(uint32_t w0) { return (double)(w0 >> 8) / 16777216.0; }

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
    if (D.tv) return D.tv[D.toff[site] + idx];
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
// per-team run context
// -----------------------------------------------------------------------------------------
struct Run {
    const SbProblem *P;
    SbState st;
    SbSlot *slot;
    Draws D;
    double T, Tp, invT, invTp, p_grow;
    int sfp;
    int tsize, rank, par;
    unsigned long long *bar;
    unsigned long long epoch;
    double *tred;      // [2][tsize][SB_REDW] team reduction scratch
    double *sred;      // smem [32 * SB_REDW]
    int *sint;         // smem [96]
    int *ssz;          // smem [K] candidate cluster sizes
    int *smem_mem;     // smem member list (geo)
    double *smem_best; // smem Prim keys (geo)
    int smem_cap;
    // committed posterior terms: kept in registers by every thread (uniform across the team)
    double ll, pr_size, pr_geo, pr_w, pr_src;
    long long i_step;
    int err;
};

SB_HD size_t sb_smem_bytes(int K, int cap) {
    size_t b = 32 * SB_REDW * sizeof(double);     // sred
    b += (size_t)cap * sizeof(double);             // Prim keys
    b += 96 * sizeof(int);                         // sint
    b += (size_t)K * sizeof(int);                  // ssz
    b += (size_t)cap * sizeof(int);                // member list
    return (b + 15) & ~(size_t)15;
}

SB_DEV void team_sync(Run &R) {
    sb_cta_sync();
    if (R.tsize > 1) {
        R.epoch++;
        if (sb_tid() == 0) {
            sb_threadfence();
            sb_atomic_add_u64(R.bar, 1ull);
            unsigned long long target = R.epoch * (unsigned long long)R.tsize;
            while (sb_ld_acquire_u64(R.bar) < target) sb_spin();
            sb_threadfence();
        }
        sb_cta_sync();
    }
}

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

// CTA-wide sum of K values per thread; fixed order, result in every thread.
template <int K, bool PROD = false>
SB_DEV void block_sum(Run &R, double (&v)[K]) {
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = PROD ? warp_prod(v[k]) : warp_sum(v[k]);
    if (sb_lane() == 0)
        for (int k = 0; k < K; k++) R.sred[sb_warp() * SB_REDW + k] = v[k];
    sb_cta_sync();
    int nw = sb_nwarps();
    for (int k = 0; k < K; k++) {
        double t = PROD ? 1.0 : 0.0;
        for (int w = 0; w < nw; w++) { if (PROD) t *= R.sred[w * SB_REDW + k]; else t += R.sred[w * SB_REDW + k]; }
        v[k] = t;
    }
    sb_cta_sync();
}

// team-wide sum (one barrier); result identical in every thread of the team.
template <int K, bool PROD = false>
SB_DEV void team_sum(Run &R, double (&v)[K]) {
    block_sum<K, PROD>(R, v);
    if (R.tsize == 1) return;
    double *buf = R.tred + (size_t)R.par * R.tsize * SB_REDW;
    R.par ^= 1;
    if (sb_tid() == 0)
        for (int k = 0; k < K; k++) buf[R.rank * SB_REDW + k] = v[k];
    team_sync(R);
    for (int k = 0; k < K; k++) {
        double t = PROD ? 1.0 : 0.0;
        for (int r = 0; r < R.tsize; r++) { if (PROD) t *= buf[r * SB_REDW + k]; else t += buf[r * SB_REDW + k]; }
        v[k] = t;
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
SB_DEV int team_pick(Run &R, const double *val, int n, double u, double &total) {
    int lo, hi; team_range(R, n, lo, hi);
    double s[1] = {0.0};
    for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) s[0] += val[i];
    block_sum<1>(R, s);
    double *buf = R.tred + (size_t)R.par * R.tsize * SB_REDW;
    R.par ^= 1;
    if (sb_tid() == 0) buf[R.rank * SB_REDW] = s[0];
    team_sync(R);
    total = 0.0;
    for (int r = 0; r < R.tsize; r++) total += buf[r * SB_REDW];
    if (!(total > 0.0)) return -1;
    double target = u * total;
    // block containing the target
    int B = (n + R.tsize - 1) / R.tsize;
    int rb = -1; double prefix = 0.0, acc = 0.0;
    for (int r = 0; r < R.tsize; r++) {
        double sr = buf[r * SB_REDW];
        if (sr > 0.0) { rb = r; prefix = acc; if (acc + sr > target) break; }
        acc += sr;
    }
    int blo = rb * B, bhi = blo + B; if (bhi > n) bhi = n;
    double tloc = target - prefix;
    if (sb_warp() == 0) {
        int len = bhi - blo, ch = (len + 31) / 32, lane = sb_lane();
        int a = blo + lane * ch, b = a + ch; if (a > bhi) a = bhi; if (b > bhi) b = bhi;
        double cs = 0.0;
        for (int i = a; i < b; i++) cs += val[i];
        double incl = cs;
        for (int off = 1; off < 32; off <<= 1) {
            double t = sb_shfl_f64(incl, lane - off);
            if (lane >= off) incl += t;
        }
        unsigned hit = sb_ballot(incl > tloc && cs > 0.0);
        unsigned pos = sb_ballot(cs > 0.0);
        int sel_lane = hit ? (sb_ffs(hit) - 1) : (pos ? 31 - sb_clz(pos) : 0);
        int found = -1;
        if (lane == sel_lane) {
            double run = incl - cs; int last = -1;
            for (int i = a; i < b; i++) {
                double vi = val[i];
                if (vi > 0.0) { last = i; run += vi; if (run > tloc) { found = i; break; } }
            }
            if (found < 0) found = last;
        }
        found = sb_shfl_i32(found, sel_lane);
        if (lane == 0) R.sint[64] = found;
    }
    sb_cta_sync();
    int o = R.sint[64];
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
    double *buf = R.tred + (size_t)R.par * R.tsize * SB_REDW;
    R.par ^= 1;
    if (sb_tid() == 0) buf[R.rank * SB_REDW] = c[0];
    team_sync(R);
    int base = 0, total = 0;
    for (int r = 0; r < R.tsize; r++) { int v = (int)buf[r * SB_REDW]; if (r < R.rank) base += v; total += v; }
    int running = base, nw = sb_nwarps();
    for (int c0 = lo; c0 < hi; c0 += sb_nthr()) {
        int i = c0 + sb_tid();
        bool p = i < hi && pred(i);
        unsigned bal = sb_ballot(p);
        if (sb_lane() == 0) R.sint[sb_warp()] = sb_popc(bal);
        sb_cta_sync();
        int wpre = 0, tot = 0;
        for (int w = 0; w < nw; w++) { int v = R.sint[w]; if (w < sb_warp()) wpre += v; tot += v; }
        if (p) list[running + wpre + sb_popc(bal & ((1u << sb_lane()) - 1u))] = i;
        running += tot;
        sb_cta_sync();
    }
    team_sync(R);
    return total;
}

// -----------------------------------------------------------------------------------------
// views: committed (0) or candidate (1) copy of a group row / object row
// -----------------------------------------------------------------------------------------
SB_DEV int gsel_of(const Run &R, int g, int view) { return R.st.gsel[g] ^ (view & R.st.dirty[g]); }
SB_DEV int osel_of(const Run &R, int o, int view) { return R.st.osel[o] ^ (view & R.st.ochg[o]); }
SB_DEV int cl_of(const Run &R, int o, int view) { return view ? R.st.cl_new[o] : R.st.cluster_of[o]; }
SB_DEV int conf_group(const SbProblem &P, int c, int o) {   // stacked group row of object o in component c >= 1, or -1
    int gid = sb_ldg(P.conf_gid + (size_t)(c - 1) * P.N + o);
    return gid < 0 ? -1 : P.goff[c] + gid;
}
SB_DEV double pow_inv(double v, double inv, bool unit) { return unit ? v : pow(v, inv); }

// heat_binary_probability (sbayes/util.py:1457-1469)
SB_DEV double heat(const Run &R, double p) {
    double a = pow_inv(p, R.invT, R.T == 1.0), b = pow_inv(1.0 - p, R.invT, R.T == 1.0);
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
// group rows: counts -> plain / tempered table, Dirichlet-categorical log-likelihood
// (sbayes/sampling/conditionals.py:105-122,174-179; sbayes/util.py:1373-1394;
//  sbayes/model/likelihood.py:65-101).  One CTA per group row; the row is written into
// copy `dst`.  If `apply_delta` the new counts are committed - dneg + dpos (and the delta
// rows are cleared), else the counts already sit in copy `dst`.
// -----------------------------------------------------------------------------------------
SB_DEV void group_row(Run &R, int g, int dst, bool apply_delta) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S;
    const size_t base = (size_t)g * F * S;
    int cur = R.st.gsel[g];
    double acc[1] = {0.0};
    for (int f = sb_tid(); f < F; f += sb_nthr()) {
        const size_t b = base + (size_t)f * S;
        double n = 0.0, sp = 0.0, stt = 0.0, ser = 0.0;
        for (int s = 0; s < S; s++) {
            int cnt;
            if (apply_delta) {
                cnt = R.st.counts[cur][b + s] - R.st.dneg[b + s] + R.st.dpos[b + s];
                R.st.counts[dst][b + s] = cnt;
                R.st.dneg[b + s] = 0; R.st.dpos[b + s] = 0;
            } else cnt = R.st.counts[dst][b + s];
            double a = sb_ldg(P.alpha + b + s), u = sb_ldg(P.unif + (size_t)f * S + s);
            double vp = (double)cnt + a;
            double vt = u + (a - u) / R.Tp + (double)cnt / R.T;
            n += (double)cnt; sp += vp; stt += vt;
            if (a > 0.0) ser += lgamma((double)cnt + a);
        }
        for (int s = 0; s < S; s++) {
            int cnt = R.st.counts[dst][b + s];
            double a = sb_ldg(P.alpha + b + s), u = sb_ldg(P.unif + (size_t)f * S + s);
            R.st.tabP[dst][b + s] = ((double)cnt + a) / sp;
            R.st.tabT[dst][b + s] = (u + (a - u) / R.Tp + (double)cnt / R.T) / stt;
        }
        size_t gf = (size_t)g * F + f;
        acc[0] += sb_ldg(P.dcc + gf) - lgamma(n + sb_ldg(P.asum + gf)) + ser;
    }
    block_sum<1>(R, acc);
    if (sb_tid() == 0) R.st.lh[dst][g] = acc[0];
}

// weight tables for every (confounder pattern p, in-a-cluster bit z):
//   w1 = has*W / sum(has*W)                 normalize_weights   (sbayes/model/likelihood.py:171-190)
//   wp = w1^(1/Tp)                          source Gibbs        (operators.py:568-572,830-836)
//   wq = has*W^(1/Tp) / sum(has*W^(1/Tp))   with/without        (operators.py:1075-1095)
//   lw = log w1                             SourcePrior         (sbayes/model/prior.py:596-611)
SB_DEV void weight_tables(Run &R) {
    const SbProblem &P = *R.P;
    const int F = P.F, C = P.C, rows = P.n_pat * 2 * F;
    for (int i = R.rank * sb_nthr() + sb_tid(); i < rows; i += R.tsize * sb_nthr()) {
        int f = i % F, pz = i / F, z = pz & 1, bits = P.pat_bits[pz >> 1];
        double t1 = 0.0, tp = 0.0, w64[SB_MAXC], wpw[SB_MAXC];
        for (int c = 0; c < C; c++) {
            double h = c == 0 ? (double)z : (double)((bits >> c) & 1);
            double w = (double)R.st.W[f * C + c];
            w64[c] = h * w; wpw[c] = h * pow_inv(w, R.invTp, R.Tp == 1.0);
            t1 += w64[c]; tp += wpw[c];
        }
        size_t o = (size_t)i * C;
        for (int c = 0; c < C; c++) {
            double a = t1 > 0.0 ? w64[c] / t1 : 0.0;
            R.st.w1[o + c] = a;
            R.st.wp[o + c] = pow_inv(a, R.invTp, R.Tp == 1.0);
            R.st.wq[o + c] = tp > 0.0 ? wpw[c] / tp : 0.0;
            R.st.lw[o + c] = log(a);
        }
    }
}

// SourcePrior (sbayes/model/prior.py:568-611) from the (pattern, feature, component)
// histogram: sum_o sum_f log w~[o,f,src] = sum H * log w1.  Whole CTA; result in every thread.
SB_DEV double source_prior_from_hist(Run &R) {
    const SbProblem &P = *R.P;
    int n = P.n_pat * 2 * P.F * P.C;
    double a[1] = {0.0};
    for (int i = sb_tid(); i < n; i += sb_nthr()) {
        int h = R.st.H[i];
        if (h != 0) a[0] += (double)h * R.st.lw[i];
    }
    block_sum<1>(R, a);
    return a[0];
}

// scipy.stats.dirichlet._logpdf of one weight row (sbayes/util.py:49, prior.py:553-558)
SB_DEV double weights_prior_row(const SbProblem &P, const float *Wrow, int f) {
    double t = 0.0;
    for (int c = 0; c < P.C; c++) {
        double a = sb_ldg(P.alpha_w + (size_t)f * P.C + c);
        if (a != 1.0) t += (a - 1.0) * log((double)Wrow[c]);
    }
    return -sb_ldg(P.wlnB + f) + t;
}
SB_DEV double weights_prior_total(Run &R) {   // WeightsPrior.__call__ (prior.py:522-551); whole CTA
    const SbProblem &P = *R.P;
    if (P.wprior_uniform) return 0.0;
    double a[1] = {0.0};
    for (int f = sb_tid(); f < P.F; f += sb_nthr()) a[0] += weights_prior_row(P, R.st.W + (size_t)f * P.C, f);
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

// -----------------------------------------------------------------------------------------
// geo prior: minimum spanning tree of one cluster (Prim), whole CTA
// (GeoPrior.__call__, compute_mst_distances: sbayes/model/prior.py:769-805,950-962)
// -----------------------------------------------------------------------------------------
SB_DEV void geo_cluster(Run &R, int k, int view, double &agg_out, double &val_out) {
    const SbProblem &P = *R.P;
    // member list in shared memory (ordered)
    int m = 0;
    if (sb_tid() == 0) {
        for (int o = 0; o < P.N; o++)
            if (cl_of(R, o, view) == k) { if (m < R.smem_cap) R.smem_mem[m] = o; m++; }
        R.sint[65] = m;
    }
    sb_cta_sync();
    m = R.sint[65];
    sb_cta_sync();
    if (m > R.smem_cap) { R.err = 3; m = R.smem_cap; }
    double sum = 0.0, mx = 0.0; int cnt = 0;
    if (m > 1) {
        // Prim from member 0; keys in shared memory, key < 0 marks "in tree"
        for (int i = sb_tid(); i < m; i += sb_nthr())
            R.smem_best[i] = i == 0 ? -1.0 : P.cost[(size_t)R.smem_mem[0] * P.N + R.smem_mem[i]];
        sb_cta_sync();
        for (int it = 0; it < m - 1; it++) {
            if (sb_tid() == 0) {      // clusters under a geo prior are small: serial argmin is fine
                int bj = -1; double bv = SB_INF;
                for (int i = 0; i < m; i++) { double b = R.smem_best[i]; if (b >= 0.0 && (bj < 0 || b < bv)) { bj = i; bv = b; } }
                R.sint[66] = bj;
                R.sred[0] = bv;
            }
            sb_cta_sync();
            int j = R.sint[66]; double e = R.sred[0];
            sb_cta_sync();
            if (j < 0) break;
            if (e != 0.0 && e < SB_INF) { sum += e; if (cnt == 0 || e > mx) mx = e; cnt++; }
            for (int i = sb_tid(); i < m; i += sb_nthr()) {
                if (i == j) R.smem_best[i] = -1.0;
                else if (R.smem_best[i] >= 0.0) {
                    double c = P.cost[(size_t)R.smem_mem[j] * P.N + R.smem_mem[i]];
                    if (c < R.smem_best[i]) R.smem_best[i] = c;
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
    double before = R.st.geo_agg[gsel_of(R, k, view)][k], after;
    double m = (double)m_size;
    if (P.geo_agg == 0) after = (mn + m * before) / (1.0 + m);
    else if (P.geo_agg == 1) after = mn + before;
    else after = fmax(mn, before);
    return geo_prob_fn(P, after) - geo_prob_fn(P, before);
}

// -----------------------------------------------------------------------------------------
// cluster-membership marginals of one object against cluster k, whole warp:
//   s_z = sum_f log sum_c l[o,f,c] * w_z[o,f,c]   (before the 1/T power)
// compute_cluster_posterior / compute_raw_cluster_probs (operators.py:1035-1095,1420-1472)
// -----------------------------------------------------------------------------------------
SB_DEV void cluster_marginals(Run &R, int o, int k, int view, bool wide, double &s0, double &s1) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S, C = P.C;
    const size_t FS = (size_t)F * S;
    const double *rowK = R.st.tabT[gsel_of(R, k, view)] + (size_t)k * FS;
    const double *rowC[SB_MAXC];
    for (int c = 1; c < C; c++) {
        int g = conf_group(P, c, o);
        rowC[c] = g < 0 ? nullptr : R.st.tabP[gsel_of(R, g, view)] + (size_t)g * FS;
    }
    int pat = sb_ldg(P.pat_id + o);
    const double *w0 = R.st.wq + (size_t)(pat * 2 + 0) * F * C;
    const double *w1 = R.st.wq + (size_t)(pat * 2 + 1) * F * C;
    const uint8_t *xr = P.x + (size_t)o * P.ldx;
    const bool powT = wide && R.T != 1.0;
    double a0s = 0.0, a1s = 0.0;
    for (int f = sb_lane(); f < F; f += 32) {
        int x = sb_ldg(xr + f);
        if (x == SB_NA_U8) continue;
        double l0 = rowK[(size_t)f * S + x];
        if (powT) l0 = pow(l0, R.invT);
        double a0 = 0.0, a1 = l0 * w1[f * C];
        for (int c = 1; c < C; c++) {
            if (!rowC[c]) continue;
            double l = rowC[c][(size_t)f * S + x];
            a0 += l * w0[f * C + c];
            a1 += l * w1[f * C + c];
        }
        a0s += log(a0); a1s += log(a1);
    }
    s0 = warp_sum(a0s); s1 = warp_sum(a1s);
}

// AlterCluster.compute_cluster_posterior (operators.py:1035-1073), whole warp -> every lane
SB_DEV double cluster_posterior(Run &R, int o, int k, int view, bool gibbsish, bool use_geo, int m_size) {
    if (R.sfp || !gibbsish) return 0.5;
    double s0, s1;
    cluster_marginals(R, o, k, view, false, s0, s1);
    double d = s1 * R.invT - s0 * R.invT;
    if (use_geo && R.P->geo) d = (s1 * R.invT + geo_delta(R, o, k, view, m_size) * R.invTp) - s0 * R.invT;
    double post = sb_expit(d);
    return (post + SB_SMOOTH) / (1.0 + 2.0 * SB_SMOOTH);
}

// get_jump_lh (operators.py:1679-1722) for one member o of k_src, whole warp
SB_DEV double jump_prob(Run &R, int o, int k_src, int k_tgt, int view) {
    const SbProblem &P = *R.P;
    const int F = P.F, S = P.S, C = P.C;
    const size_t FS = (size_t)F * S;
    const double *rowS = R.st.tabT[gsel_of(R, k_src, view)] + (size_t)k_src * FS;
    const double *rowG = R.st.tabT[gsel_of(R, k_tgt, view)] + (size_t)k_tgt * FS;
    const double *rowC[SB_MAXC];
    for (int c = 1; c < C; c++) {
        int g = conf_group(P, c, o);
        rowC[c] = g < 0 ? nullptr : R.st.tabT[gsel_of(R, g, view)] + (size_t)g * FS;
    }
    int pat = sb_ldg(P.pat_id + o);
    const double *wh = R.st.wq + (size_t)(pat * 2 + 1) * F * C;
    const uint8_t *xr = P.x + (size_t)o * P.ldx;
    double ls = 0.0, lj = 0.0;
    for (int f = sb_lane(); f < F; f += 32) {
        int x = sb_ldg(xr + f);
        if (x == SB_NA_U8) continue;
        double conf = 0.0;
        for (int c = 1; c < C; c++)
            if (rowC[c]) conf += wh[f * C + c] * rowC[c][(size_t)f * S + x];
        ls += log(conf + wh[f * C] * rowS[(size_t)f * S + x]);
        lj += log(conf + wh[f * C] * rowG[(size_t)f * S + x]);
    }
    ls = warp_sum(ls); lj = warp_sum(lj);
    double lh_stay = exp(ls * R.invT) + SB_EPS32, lh_jump = exp(lj * R.invT) + SB_EPS32;
    return lh_jump / (lh_jump + lh_stay);
}

// -----------------------------------------------------------------------------------------
// proposal bookkeeping
// -----------------------------------------------------------------------------------------
struct Prop {
    int changed;        // 0 => direct reject before anything was touched
    int force_reject;   // reject after the candidate was built (rollback needed)
    int M;              // number of changed objects (st.list_a[0..M))
    double lq, lqb;
    double sp_new, wp_new;   // weights operator: recomputed source / weights prior
};

// mark object o of the proposal: dirty rows of all groups it leaves or joins
SB_DEV void mark_object(Run &R, int o) {
    const SbProblem &P = *R.P;
    R.st.ochg[o] = 1;
    int a = R.st.cluster_of[o], b = R.st.cl_new[o];
    if (a >= 0) R.st.dirty[a] = 1;
    if (b >= 0) R.st.dirty[b] = 1;
    for (int c = 1; c < P.C; c++) { int g = conf_group(P, c, o); if (g >= 0) R.st.dirty[g] = 1; }
}

// group row an observation of object o assigned to component c counts towards
SB_DEV int obs_group(const Run &R, int o, int c, int view) {
    return c == 0 ? cl_of(R, o, view) : conf_group(*R.P, c, o);
}

// tempered table entry of row g at (f, x) with the proposal's own observations removed
// (component_likelihood_given_unchanged, operators.py:863-928)
SB_DEV double unchanged_entry(const Run &R, int g, int f, int x) {
    const SbProblem &P = *R.P;
    const size_t b = ((size_t)g * P.F + f) * P.S;
    const int32_t *cnt = R.st.counts[R.st.gsel[g]] + b;
    double sum = 0.0, at = 0.0;
    for (int s = 0; s < P.S; s++) {
        double a = sb_ldg(P.alpha + b + s), u = sb_ldg(P.unif + (size_t)f * P.S + s);
        double v = u + (a - u) / R.Tp + (double)(cnt[s] - R.st.dneg[b + s]) / R.T;
        sum += v; if (s == x) at = v;
    }
    return at / sum;
}

// ClusterOperator.gibbs_sample_source / ClusterJump.gibbs_sample_source_jump
// (operators.py:796-849,1775-1827) for the M changed objects in st.list_a.  cl_new holds
// the candidate membership.  Three team barriers.  Returns log_q, log_q_back of the source
// draw, leaves candidate rows (counts, tables, lh, geo) and the updated histogram H.
SB_DEV void resample_sources(Run &R, int M, int k_new, int k_old, bool jump, double &lq, double &lqb) {
    const SbProblem &P = *R.P;
    const int F = P.F, C = P.C;
    // R1: remove the objects' own observations (dneg), mark dirty rows, H -= old
    for (int m = R.rank; m < M; m += R.tsize) {
        int o = R.st.list_a[m];
        int pat = sb_ldg(P.pat_id + o), z_old = R.st.cluster_of[o] >= 0;
        const uint8_t *xr = P.x + (size_t)o * P.ldx;
        const uint8_t *so = R.st.src[R.st.osel[o]] + (size_t)o * P.ldx;
        for (int f = sb_tid(); f < F; f += sb_nthr()) {
            int x = sb_ldg(xr + f);
            if (x == SB_NA_U8) continue;
            int c = so[f];
            int g = obs_group(R, o, c, 0);
            if (g >= 0) sb_atomic_add_i32(R.st.dneg + ((size_t)g * F + f) * P.S + x, 1);
            sb_atomic_add_i32(R.st.H + ((size_t)(pat * 2 + z_old) * F + f) * C + c, -1);
        }
        if (sb_tid() == 0) mark_object(R, o);
    }
    team_sync(R);
    // R2: conditional source distribution, draw, add the new observations (dpos), H += new
    const bool unitT = R.T == 1.0;
    for (int m = R.rank; m < M; m += R.tsize) {
        int o = R.st.list_a[m];
        int pat = sb_ldg(P.pat_id + o), z_old = R.st.cluster_of[o] >= 0, z_new = R.st.cl_new[o] >= 0;
        const uint8_t *xr = P.x + (size_t)o * P.ldx;
        const uint8_t *so = R.st.src[R.st.osel[o]] + (size_t)o * P.ldx;
        uint8_t *sn = R.st.src[R.st.osel[o] ^ 1] + (size_t)o * P.ldx;
        const double *tw = jump ? R.st.w1 : R.st.wp;
        const double *wf = tw + (size_t)(pat * 2 + z_new) * F * C;
        const double *wb = tw + (size_t)(pat * 2 + (jump ? z_new : z_old)) * F * C;
        int gC[SB_MAXC];
        for (int c = 1; c < C; c++) gC[c] = conf_group(P, c, o);
        double acc[2] = {0.0, 0.0};
        for (int f = sb_tid(); f < F; f += sb_nthr()) {
            int x = sb_ldg(xr + f);
            if (x == SB_NA_U8) { sn[f] = SB_NA_U8; continue; }
            double q[SB_MAXC], qb[SB_MAXC], tot = 0.0, totb = 0.0;
            for (int c = 0; c < C; c++) {
                double l = 1.0, lo = 1.0;
                if (!R.sfp) {
                    int g = c == 0 ? k_new : gC[c];
                    l = g < 0 ? 0.0 : pow_inv(unchanged_entry(R, g, f, x), R.invT, unitT);
                    lo = l;
                    if (jump && c == 0) lo = pow_inv(unchanged_entry(R, k_old, f, x), R.invT, unitT);
                }
                q[c] = wf[f * C + c] * l; qb[c] = wb[f * C + c] * lo;
                tot += q[c]; totb += qb[c];
            }
            if (!R.sfp) { for (int c = 0; c < C; c++) { q[c] /= tot; qb[c] /= totb; } }
            // sample_categorical (sbayes/preprocessing.py:224-256)
            double u = draw_uniform(R.D, SITE_GIBBS, (uint32_t)((size_t)m * F + f));
            double cum = 0.0, last = 0.0;
            for (int c = 0; c < C; c++) last += q[c];
            int nw = 0; bool got = false;
            for (int c = 0; c < C; c++) { cum += q[c]; if (!got && u < cum / last) { nw = c; got = true; } }
            int c_old = so[f];
            acc[0] += log(q[nw]); acc[1] += log(qb[c_old]);
            sn[f] = (uint8_t)nw;
            int g = obs_group(R, o, nw, 1);
            if (g >= 0) sb_atomic_add_i32(R.st.dpos + ((size_t)g * F + f) * P.S + x, 1);
            sb_atomic_add_i32(R.st.H + ((size_t)(pat * 2 + z_new) * F + f) * C + nw, 1);
        }
        block_sum<2>(R, acc);
        if (sb_tid() == 0) { R.st.red_a[m] = acc[0]; R.st.red_b[m] = acc[1]; }
    }
    team_sync(R);
    // (the caller runs candidate_rows() next, then sums red_a / red_b)
    lq = 0.0; lqb = 0.0;
}

// R3: candidate rows of every dirty group (one CTA per row), source prior of the candidate
// histogram and geo prior of the dirty clusters (last CTA of the team).  One team barrier.
SB_DEV void candidate_rows(Run &R, bool geo_clusters) {
    const SbProblem &P = *R.P;
    int j = 0;
    for (int g = 0; g < P.G_tot; g++) {
        if (!R.st.dirty[g]) continue;
        if (j % R.tsize == R.rank) group_row(R, g, R.st.gsel[g] ^ 1, true);
        j++;
    }
    if (R.rank == R.tsize - 1) {
        double sp = source_prior_from_hist(R);
        double geo = 0.0;
        if (P.geo) {
            for (int k = 0; k < P.K; k++) {
                if (geo_clusters && R.st.dirty[k]) {
                    double a, v; geo_cluster(R, k, 1, a, v);
                    if (sb_tid() == 0) { R.st.geo_agg[R.st.gsel[k] ^ 1][k] = a; R.st.geo_val[R.st.gsel[k] ^ 1][k] = v; }
                    geo += v;
                } else {
                    if (R.st.dirty[k] && sb_tid() == 0) {     // membership unchanged: carry the committed value over
                        R.st.geo_agg[R.st.gsel[k] ^ 1][k] = R.st.geo_agg[R.st.gsel[k]][k];
                        R.st.geo_val[R.st.gsel[k] ^ 1][k] = R.st.geo_val[R.st.gsel[k]][k];
                    }
                    geo += R.st.geo_val[R.st.gsel[k]][k];
                }
            }
        }
        if (sb_tid() == 0) { R.st.sc[SC_SP_NEW] = sp; R.st.sc[SC_GEO_NEW] = geo; }
    }
    team_sync(R);
}

// sum of red_a / red_b over the M changed objects (fixed order), whole CTA
SB_DEV void sum_object_partials(Run &R, int M, double &a, double &b) {
    double v[2] = {0.0, 0.0};
    for (int m = sb_tid(); m < M; m += sb_nthr()) { v[0] += R.st.red_a[m]; v[1] += R.st.red_b[m]; }
    block_sum<2>(R, v);
    a = v[0]; b = v[1];
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
        for (int o = lo + sb_tid(); o < hi; o += sb_nthr()) R.st.flag_a[o] = touches_cluster(R, o, k, view);
        team_sync(R);
    }
    for (int o = lo + sb_tid(); o < hi; o += sb_nthr()) {
        bool freeo = cl_of(R, o, view) < 0, c = freeo;
        if (freeo && nb == NB_DIRECT) c = touches_cluster(R, o, k, view);
        if (freeo && nb == NB_TWOSTEP) c = touches_flag(R, o, R.st.flag_a);
        R.st.flag_b[o] = c;
    }
    sb_cta_sync();
}

// -----------------------------------------------------------------------------------------
// AlterCluster grow / shrink (operators.py:969-1249)
// -----------------------------------------------------------------------------------------
SB_DEV Prop op_alter_cluster(Run &R, bool gibbsish, bool use_geo, int nb) {
    const SbProblem &P = *R.P;
    Prop pr; pr.changed = 0; pr.force_reject = 0; pr.M = 0; pr.lq = 0.0; pr.lqb = -SB_INF; pr.sp_new = 0.0; pr.wp_new = 0.0;
    int k = draw_index(R.D, SITE_CLUSTER, 0, P.K);
    int size = R.st.sizes[k];
    double extra = 0.0; bool grow;
    if (size == P.min_size) { grow = true; extra = -log(2.0); }
    else if (size == P.max_size) { grow = false; extra = -log(2.0); }
    else grow = draw_uniform(R.D, SITE_GROW) < R.p_grow;
    int lo, hi; team_range(R, P.N, lo, hi);
    const int gw = sb_warp(), nw = sb_nwarps();

    if (grow) {
        grow_candidates(R, k, nb, 0);
        double cnt[1] = {0.0};
        for (int o = lo + gw; o < hi; o += nw) {
            double v = 0.0;
            if (R.st.flag_b[o]) { v = heat(R, cluster_posterior(R, o, k, 0, gibbsish, use_geo, size)); if (sb_lane() == 0) cnt[0] += 1.0; }
            if (sb_lane() == 0) R.st.val[o] = v;
        }
        team_sum<1>(R, cnt);
        if (cnt[0] == 0.0 || size == P.max_size) return pr;
        double total;
        int o = team_pick(R, R.st.val, P.N, draw_uniform(R.D, SITE_OBJECT), total);
        if (o < 0) return pr;
        double lq_obj = log(R.st.val[o] / total);
        pr.changed = 1; pr.M = 1;
        if (R.rank == 0 && sb_tid() == 0) { R.st.cl_new[o] = k; R.st.list_a[0] = o; }
        if (sb_tid() == 0) { for (int i = 0; i < P.K; i++) R.ssz[i] = R.st.sizes[i]; R.ssz[k] = size + 1; }
        team_sync(R);
        double lqs, lqbs;
        resample_sources(R, 1, k, k, false, lqs, lqbs);
        candidate_rows(R, true);
        sum_object_partials(R, 1, lqs, lqbs);
        // back: remove o again, p_remove ~ 1 - posterior over the members of the grown cluster
        double s[1] = {0.0};
        for (int j = lo + gw; j < hi; j += nw) {
            if (R.st.cl_new[j] != k) continue;
            double v = 1.0 - heat(R, cluster_posterior(R, j, k, 1, gibbsish, use_geo, size + 1));
            if (sb_lane() == 0) { s[0] += v; if (j == o) R.st.sc[SC_AUX0] = v; }
        }
        team_sum<1>(R, s);
        double lqb_obj = log(R.st.sc[SC_AUX0] / s[0]);
        pr.lq = lq_obj + lqs + log(R.p_grow);
        pr.lqb = lqb_obj + lqbs + log(1.0 - R.p_grow) + extra;
    } else {
        if (size == P.min_size) return pr;
        for (int o = lo + gw; o < hi; o += nw) {
            double v = 0.0;
            if (R.st.cluster_of[o] == k) v = 1.0 - heat(R, cluster_posterior(R, o, k, 0, gibbsish, use_geo, size));
            if (sb_lane() == 0) R.st.val[o] = v;
        }
        sb_cta_sync();
        double total;
        int o = team_pick(R, R.st.val, P.N, draw_uniform(R.D, SITE_OBJECT), total);
        if (o < 0) return pr;
        double lq_obj = log(R.st.val[o] / total);
        pr.changed = 1; pr.M = 1;
        if (R.rank == 0 && sb_tid() == 0) { R.st.cl_new[o] = -1; R.st.list_a[0] = o; }
        if (sb_tid() == 0) { for (int i = 0; i < P.K; i++) R.ssz[i] = R.st.sizes[i]; R.ssz[k] = size - 1; }
        team_sync(R);
        double lqs, lqbs;
        resample_sources(R, 1, k, k, false, lqs, lqbs);
        candidate_rows(R, true);
        sum_object_partials(R, 1, lqs, lqbs);
        // back: grow by o again
        grow_candidates(R, k, nb, 1);
        double s[2] = {0.0, 0.0};
        for (int j = lo + gw; j < hi; j += nw) {
            if (!R.st.flag_b[j]) continue;
            double v = heat(R, cluster_posterior(R, j, k, 1, gibbsish, use_geo, size - 1));
            if (sb_lane() == 0) { s[0] += v; if (j == o) { R.st.sc[SC_AUX0] = v; s[1] = 1.0; } }
        }
        team_sum<2>(R, s);
        if (s[1] == 0.0) { pr.force_reject = 1; return pr; }   // o is not a grow candidate of the shrunk cluster
        double lqb_obj = log(R.st.sc[SC_AUX0] / s[0]);
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
SB_DEV void wide_probs(Run &R, int k, int n_av, int view, bool use_geo, int size_k) {
    const SbProblem &P = *R.P;
    const double eps = 0.01 / (double)P.N;
    int lo, hi; team_range(R, n_av, lo, hi);
    const int gw = sb_warp(), nw = sb_nwarps();
    double *p = R.st.val;
    for (int i = lo + gw; i < hi; i += nw) {
        int o = R.st.list_b[i];
        double raw = 0.5;
        if (!R.sfp) {
            double s0, s1;
            cluster_marginals(R, o, k, view, true, s0, s1);
            double m0 = exp(s0 * R.invT), m1 = exp(s1 * R.invT);
            if (use_geo && P.geo) m1 = m1 * exp(geo_delta(R, o, k, view, size_k) * R.invTp / SB_GEO_SCALER);
            raw = m1 / (m0 + m1 + SB_EPS32);
        }
        if (sb_lane() == 0) p[i] = raw + SB_EPS32;
    }
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
        double mem = cl_of(R, R.st.list_b[i], view) == k ? 1.0 : 0.0;
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
    Prop pr; pr.changed = 0; pr.force_reject = 0; pr.M = 0; pr.lq = 0.0; pr.lqb = -SB_INF; pr.sp_new = 0.0; pr.wp_new = 0.0;
    int k = draw_index(R.D, SITE_CLUSTER, 0, P.K);
    int size = R.st.sizes[k];
    int n_av = team_compact(R, P.N, R.st.list_b, [&](int o) { int c = R.st.cluster_of[o]; return c < 0 || c == k; });
    wide_probs(R, k, n_av, 0, use_geo, size);
    int lo, hi; team_range(R, n_av, lo, hi);
    double *p = R.st.val;
    // redraw until the membership changes (operators.py:1515-1517)
    double st[3];
    for (int attempt = 0;; attempt++) {
        st[0] = 0.0; st[1] = 0.0; st[2] = 0.0;
        for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) {
            int o = R.st.list_b[i];
            bool nm = draw_uniform_f32(R.D, SITE_WIDE, (uint32_t)i, attempt) < p[i];
            bool old = R.st.cluster_of[o] == k;
            R.st.flag_a[o] = nm;
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
        int o = R.st.list_b[i];
        bool nm = R.st.flag_a[o], old = R.st.cluster_of[o] == k;
        a[0] += log(nm ? p[i] : 1.0 - p[i]);
        b[0] *= old ? p[i] : 1.0 - p[i];
    }
    team_sum<1>(R, a);
    team_sum<1, true>(R, b);
    double lq = a[0] - log(1.0 - b[0]);
    // candidate membership and the list of flipped objects
    for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) {
        int o = R.st.list_b[i];
        R.st.cl_new[o] = R.st.flag_a[o] ? k : -1;
    }
    int M = team_compact(R, n_av, R.st.list_a, [&](int i) { int o = R.st.list_b[i]; return (bool)R.st.flag_a[o] != (R.st.cluster_of[o] == k); });
    for (int m = R.rank * sb_nthr() + sb_tid(); m < M; m += R.tsize * sb_nthr()) R.st.list_a[m] = R.st.list_b[R.st.list_a[m]];
    if (sb_tid() == 0) { for (int i = 0; i < P.K; i++) R.ssz[i] = R.st.sizes[i]; R.ssz[k] = n_new; }
    team_sync(R);
    pr.changed = 1; pr.M = M;
    double lqs, lqbs;
    resample_sources(R, M, k, k, false, lqs, lqbs);
    candidate_rows(R, true);
    sum_object_partials(R, M, lqs, lqbs);
    // backward probability from the new state
    wide_probs(R, k, n_av, 1, use_geo, n_new);
    a[0] = 0.0; b[0] = 1.0;
    for (int i = lo + sb_tid(); i < hi; i += sb_nthr()) {
        int o = R.st.list_b[i];
        bool nm = R.st.flag_a[o], old = R.st.cluster_of[o] == k;
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
SB_DEV Prop op_cluster_jump(Run &R) {
    const SbProblem &P = *R.P;
    Prop pr; pr.changed = 0; pr.force_reject = 0; pr.M = 0; pr.lq = 0.0; pr.lqb = -SB_INF; pr.sp_new = 0.0; pr.wp_new = 0.0;
    int k_src, k_tgt;
    draw_pair(R.D, SITE_JUMP, P.K, k_src, k_tgt);
    if (R.st.sizes[k_src] <= P.min_size || R.st.sizes[k_tgt] >= P.max_size) return pr;
    int lo, hi; team_range(R, P.N, lo, hi);
    const int gw = sb_warp(), nw = sb_nwarps();
    for (int o = lo + gw; o < hi; o += nw) {
        double v = 0.0;
        if (R.st.cluster_of[o] == k_src) v = jump_prob(R, o, k_src, k_tgt, 0);
        if (sb_lane() == 0) R.st.val[o] = v;
    }
    sb_cta_sync();
    double total;
    int o = team_pick(R, R.st.val, P.N, draw_uniform(R.D, SITE_OBJECT), total);
    if (o < 0) return pr;
    double lq_obj = log(R.st.val[o] / total);
    pr.changed = 1; pr.M = 1;
    if (R.rank == 0 && sb_tid() == 0) { R.st.cl_new[o] = k_tgt; R.st.list_a[0] = o; }
    if (sb_tid() == 0) {
        for (int i = 0; i < P.K; i++) R.ssz[i] = R.st.sizes[i];
        R.ssz[k_src] -= 1; R.ssz[k_tgt] += 1;
    }
    team_sync(R);
    double lqs, lqbs;
    resample_sources(R, 1, k_tgt, k_src, true, lqs, lqbs);
    candidate_rows(R, true);
    sum_object_partials(R, 1, lqs, lqbs);
    double s[1] = {0.0};
    for (int j = lo + gw; j < hi; j += nw) {
        if (R.st.cl_new[j] != k_tgt) continue;
        double v = jump_prob(R, j, k_tgt, k_src, 1);
        if (sb_lane() == 0) { s[0] += v; if (j == o) R.st.sc[SC_AUX0] = v; }
    }
    team_sum<1>(R, s);
    pr.lq = lq_obj + lqs;
    pr.lqb = log(R.st.sc[SC_AUX0] / s[0]) + lqbs;
    return pr;
}

// -----------------------------------------------------------------------------------------
// GibbsSampleSource (operators.py:423-579)
// -----------------------------------------------------------------------------------------
// calculate_source_posterior (operators.py:554-574) of object o at feature f -> q[c] (normalised)
SB_DEV void source_posterior(const Run &R, int o, int f, int x, int view, const int *gC, int pat, int z, double *q) {
    const SbProblem &P = *R.P;
    const int C = P.C;
    const double *w = R.st.wp + ((size_t)(pat * 2 + z) * P.F + f) * C;
    const bool unitT = R.T == 1.0;
    double tot = 0.0;
    for (int c = 0; c < C; c++) {
        double l = 1.0;
        if (!R.sfp) {
            int g = gC[c];
            l = g < 0 ? 0.0 : pow_inv(R.st.tabP[gsel_of(R, g, view)][((size_t)g * P.F + f) * P.S + x], R.invT, unitT);
        }
        q[c] = l * w[c]; tot += q[c];
    }
    for (int c = 0; c < C; c++) q[c] /= tot;
}

SB_DEV Prop op_gibbs_source(Run &R, bool groups_mode) {
    const SbProblem &P = *R.P;
    const int N = P.N, F = P.F, C = P.C;
    Prop pr; pr.changed = 1; pr.force_reject = 0; pr.M = 0; pr.lq = 0.0; pr.lqb = 0.0; pr.sp_new = 0.0; pr.wp_new = 0.0;
    // --- select_object_subset (operators.py:461-494) -> list_a[0..M)
    int M = 0;
    if (N <= SB_SRC_MIN) {
        M = N;
        for (int i = R.rank * sb_nthr() + sb_tid(); i < N; i += R.tsize * sb_nthr()) R.st.list_a[i] = i;
        team_sync(R);
    } else {
        int pop = N, kmax = SB_SRC_MAX_RANDOM < N ? SB_SRC_MAX_RANDOM : N;
        bool identity = true;
        if (groups_mode) {
            // components are enumerated [confounders..., clusters] (sbayes/sampling/state.py:626-634)
            int ci = draw_index(R.D, SITE_COMPONENT, 0, C);
            kmax = SB_SRC_MAX_GROUPS < N ? SB_SRC_MAX_GROUPS : N;
            if (ci == C - 1) {
                int g = draw_index(R.D, SITE_COMPONENT, 1, P.K);
                pop = team_compact(R, N, R.st.list_b, [&](int o) { return R.st.cluster_of[o] == g; });
            } else {
                int g = draw_index(R.D, SITE_COMPONENT, 1, P.conf_ng[ci]);
                const int32_t *gid = P.conf_gid + (size_t)ci * N;
                pop = team_compact(R, N, R.st.list_b, [&](int o) { return gid[o] == g; });
            }
            identity = false;
        }
        if (!identity && pop <= kmax) {
            M = pop;
            for (int i = R.rank * sb_nthr() + sb_tid(); i < M; i += R.tsize * sb_nthr()) R.st.list_a[i] = R.st.list_b[i];
        } else {
            M = kmax;
            // uniform k-subset: Floyd's algorithm over ranks, then sort (oracle/draws.py:subset_of)
            if (R.rank == 0 && sb_tid() == 0) {
                int ch[32];
                if (R.D.tv) {
                    for (int j = 0; j < M; j++) ch[j] = (int)R.D.tv[R.D.toff[SITE_SUBSET] + j];   // object ids
                } else {
                    int n = 0;
                    for (int j = pop - M; j < pop; j++) {
                        int t = draw_index(R.D, SITE_SUBSET, (uint32_t)j, j + 1);
                        bool in = false;
                        for (int q = 0; q < n; q++) in |= ch[q] == t;
                        ch[n++] = in ? j : t;
                    }
                    if (!identity) for (int j = 0; j < M; j++) ch[j] = R.st.list_b[ch[j]];
                }
                for (int a = 1; a < M; a++) { int v = ch[a], b = a - 1; while (b >= 0 && ch[b] > v) { ch[b + 1] = ch[b]; b--; } ch[b + 1] = v; }
                for (int j = 0; j < M; j++) R.st.list_a[j] = ch[j];
            }
        }
        team_sync(R);
    }
    pr.M = M;
    if (sb_tid() == 0) for (int i = 0; i < P.K; i++) R.ssz[i] = R.st.sizes[i];
    // --- G1: forward conditional, draw, counts delta
    for (int m = R.rank; m < M; m += R.tsize) {
        int o = R.st.list_a[m];
        int pat = sb_ldg(P.pat_id + o), z = R.st.cluster_of[o] >= 0;
        const uint8_t *xr = P.x + (size_t)o * P.ldx;
        const uint8_t *so = R.st.src[R.st.osel[o]] + (size_t)o * P.ldx;
        uint8_t *sn = R.st.src[R.st.osel[o] ^ 1] + (size_t)o * P.ldx;
        int gC[SB_MAXC];
        gC[0] = R.st.cluster_of[o];
        for (int c = 1; c < C; c++) gC[c] = conf_group(P, c, o);
        double acc[1] = {0.0};
        for (int f = sb_tid(); f < F; f += sb_nthr()) {
            int x = sb_ldg(xr + f);
            if (x == SB_NA_U8) { sn[f] = SB_NA_U8; continue; }
            double q[SB_MAXC];
            source_posterior(R, o, f, x, 0, gC, pat, z, q);
            double u = draw_uniform(R.D, SITE_GIBBS, (uint32_t)((size_t)m * F + f));
            double cum = 0.0, last = 0.0;
            for (int c = 0; c < C; c++) last += q[c];
            int nw = 0; bool got = false;
            for (int c = 0; c < C; c++) { cum += q[c]; if (!got && u < cum / last) { nw = c; got = true; } }
            int c_old = so[f];
            acc[0] += log(q[nw]);
            sn[f] = (uint8_t)nw;
            if (nw != c_old) {
                int g0 = gC[c_old], g1 = gC[nw];
                if (g0 >= 0) sb_atomic_add_i32(R.st.dneg + ((size_t)g0 * F + f) * P.S + x, 1);
                if (g1 >= 0) sb_atomic_add_i32(R.st.dpos + ((size_t)g1 * F + f) * P.S + x, 1);
                sb_atomic_add_i32(R.st.H + ((size_t)(pat * 2 + z) * F + f) * C + c_old, -1);
                sb_atomic_add_i32(R.st.H + ((size_t)(pat * 2 + z) * F + f) * C + nw, 1);
            }
        }
        block_sum<1>(R, acc);
        if (sb_tid() == 0) { R.st.red_a[m] = acc[0]; mark_object(R, o); }
    }
    team_sync(R);
    candidate_rows(R, false);
    // --- G2: backward conditional from the new tables, evaluated at the old sources
    for (int m = R.rank; m < M; m += R.tsize) {
        int o = R.st.list_a[m];
        int pat = sb_ldg(P.pat_id + o), z = R.st.cluster_of[o] >= 0;
        const uint8_t *xr = P.x + (size_t)o * P.ldx;
        const uint8_t *so = R.st.src[R.st.osel[o]] + (size_t)o * P.ldx;
        int gC[SB_MAXC];
        gC[0] = R.st.cluster_of[o];
        for (int c = 1; c < C; c++) gC[c] = conf_group(P, c, o);
        double acc[1] = {0.0};
        for (int f = sb_tid(); f < F; f += sb_nthr()) {
            int x = sb_ldg(xr + f);
            if (x == SB_NA_U8) continue;
            double q[SB_MAXC];
            source_posterior(R, o, f, x, R.sfp ? 0 : 1, gC, pat, z, q);
            acc[0] += log(q[so[f]]);
        }
        block_sum<1>(R, acc);
        if (sb_tid() == 0) R.st.red_b[m] = acc[0];
    }
    team_sync(R);
    sum_object_partials(R, M, pr.lq, pr.lqb);
    return pr;
}

// -----------------------------------------------------------------------------------------
// GibbsSampleWeights (operators.py:582-693).  Edits W in place, always accepted.
// -----------------------------------------------------------------------------------------
SB_DEV double beta_logpdf_nolnB(double x, double a, double b) {   // scipy beta.logpdf without the -betaln term
    double t1 = a == 1.0 ? 0.0 : (a - 1.0) * log(x);
    double t2 = b == 1.0 ? 0.0 : (b - 1.0) * log1p(-x);
    return t1 + t2;
}
SB_DEV Prop op_gibbs_weights(Run &R) {
    const SbProblem &P = *R.P;
    const int F = P.F, C = P.C;
    Prop pr; pr.changed = 1; pr.force_reject = 0; pr.M = 0; pr.lq = -SB_INF; pr.lqb = 0.0; pr.sp_new = 0.0; pr.wp_new = 0.0;
    int i1, i2;
    draw_pair(R.D, SITE_WPAIR, C, i1, i2);
    if (sb_tid() == 0) for (int i = 0; i < P.K; i++) R.ssz[i] = R.st.sizes[i];
    for (int f = R.rank * sb_nthr() + sb_tid(); f < F; f += R.tsize * sb_nthr()) {
        // counts of sources over the objects that have both components (operators.py:648-656)
        double c1 = sb_ldg(P.alpha_w + (size_t)f * C + i1), c2 = sb_ldg(P.alpha_w + (size_t)f * C + i2);
        for (int pz = 0; pz < P.n_pat * 2; pz++) {
            int bits = P.pat_bits[pz >> 1], z = pz & 1;
            bool h1 = i1 == 0 ? z : (bits >> i1) & 1, h2 = i2 == 0 ? z : (bits >> i2) & 1;
            if (h1 && h2) {
                const int32_t *h = R.st.H + ((size_t)pz * F + f) * C;
                c1 += (double)h[i1]; c2 += (double)h[i2];
            }
        }
        double a = 1.0 + c2 / R.Tp, b = 1.0 + c1 / R.Tp;
        double a2 = draw_beta(R.D, a, b, (uint32_t)f);
        // float32 arithmetic of resample_weight_for_two_components (operators.py:664-672)
        float Wo[SB_MAXC], Wn[SB_MAXC];
        for (int c = 0; c < C; c++) { Wo[c] = R.st.W[f * C + c]; Wn[c] = Wo[c]; }
        float w02 = Wo[i1] + Wo[i2];
        Wn[i1] = (float)((1.0 - a2) * (double)w02);
        Wn[i2] = (float)(a2 * (double)w02);
        float s = 0.0f;
        for (int c = 0; c < C; c++) s = s + Wn[c];
        for (int c = 0; c < C; c++) Wn[c] = Wn[c] / s;
        double a2_old = (double)(Wo[i2] / w02);
        double lq = beta_logpdf_nolnB(a2, a, b), lqb = beta_logpdf_nolnB(a2_old, a, b);
        // per-feature target: source likelihood + pointwise weights prior (operators.py:617-623,678-686)
        double lp_old = 0.0, lp_new = 0.0;
        for (int pz = 0; pz < P.n_pat * 2; pz++) {
            int bits = P.pat_bits[pz >> 1], z = pz & 1;
            const int32_t *h = R.st.H + ((size_t)pz * F + f) * C;
            const double *lwo = R.st.lw + ((size_t)pz * F + f) * C;
            double tot = 0.0;
            for (int c = 0; c < C; c++) { double hc = c == 0 ? (double)z : (double)((bits >> c) & 1); tot += hc * (double)Wn[c]; }
            for (int c = 0; c < C; c++) {
                if (h[c] == 0) continue;
                double hc = c == 0 ? (double)z : (double)((bits >> c) & 1);
                lp_old += (double)h[c] * lwo[c];
                lp_new += (double)h[c] * log(hc * (double)Wn[c] / tot);
            }
        }
        lp_old += weights_prior_row(P, Wo, f);
        lp_new += weights_prior_row(P, Wn, f);
        double p_acc = exp((lp_new - lp_old + lqb - lq) / R.Tp);
        bool acc = draw_uniform_f32(R.D, SITE_WACCEPT, (uint32_t)f, 0) < p_acc;
        if (acc) for (int c = 0; c < C; c++) R.st.W[f * C + c] = Wn[c];
    }
    team_sync(R);
    weight_tables(R);
    team_sync(R);
    double sp = source_prior_from_hist(R);
    double wp = weights_prior_total(R);
    pr.sp_new = sp; pr.wp_new = wp;     // every CTA computed the same values
    return pr;
}

// -----------------------------------------------------------------------------------------
// full (uncached) evaluation of a state: count rebuild + histogram + tables + likelihood +
// priors.  recalculate_feature_counts (sbayes/sampling/counts.py:35-52),
// Likelihood.__call__ (likelihood.py:47-63), Prior.__call__ (prior.py:74-88).
// -----------------------------------------------------------------------------------------
SB_DEV void rebuild_state(Run &R) {
    const SbProblem &P = *R.P;
    const int N = P.N, F = P.F, S = P.S, C = P.C;
    const size_t nG = (size_t)P.G_tot * F * S;
    const int nH = P.n_pat * 2 * F * C;
    const int stride = R.tsize * sb_nthr(), t0 = R.rank * sb_nthr() + sb_tid();
    for (size_t i = t0; i < nG; i += stride) { R.st.counts[0][i] = 0; R.st.dneg[i] = 0; R.st.dpos[i] = 0; }
    for (int i = t0; i < nH; i += stride) R.st.H[i] = 0;
    for (int g = t0; g < P.G_tot; g += stride) { R.st.gsel[g] = 0; R.st.dirty[g] = 0; }
    for (int k = t0; k < P.K; k += stride) R.st.sizes[k] = 0;
    for (int o = t0; o < N; o += stride) { R.st.cl_new[o] = R.st.cluster_of[o]; R.st.ochg[o] = 0; }
    team_sync(R);
    // histogram pass: warp per object, lanes over features (coalesced byte reads)
    int lo, hi; team_range(R, N, lo, hi);
    for (int o = lo + sb_warp(); o < hi; o += sb_nwarps()) {
        int k = R.st.cluster_of[o], pat = sb_ldg(P.pat_id + o), z = k >= 0;
        int gC[SB_MAXC]; gC[0] = k;
        for (int c = 1; c < C; c++) gC[c] = conf_group(P, c, o);
        const uint8_t *xr = P.x + (size_t)o * P.ldx;
        const uint8_t *so = R.st.src[R.st.osel[o]] + (size_t)o * P.ldx;
        for (int f = sb_lane(); f < F; f += 32) {
            int x = sb_ldg(xr + f);
            if (x == SB_NA_U8) continue;
            int c = so[f];
            if (c >= C) { R.err = 2; continue; }
            int g = gC[c];
            if (g >= 0) sb_atomic_add_i32(R.st.counts[0] + ((size_t)g * F + f) * S + x, 1);
            sb_atomic_add_i32(R.st.H + ((size_t)(pat * 2 + z) * F + f) * C + c, 1);
        }
        if (sb_lane() == 0 && k >= 0) sb_atomic_add_i32(R.st.sizes + k, 1);
    }
    team_sync(R);
    for (int g = R.rank; g < P.G_tot; g += R.tsize) group_row(R, g, 0, false);
    weight_tables(R);
    team_sync(R);
    double v[1] = {0.0};
    for (int g = sb_tid(); g < P.G_tot; g += sb_nthr()) v[0] += R.st.lh[0][g];
    block_sum<1>(R, v);
    R.ll = R.sfp ? 0.0 : v[0];
    R.pr_src = source_prior_from_hist(R);
    R.pr_w = weights_prior_total(R);
    if (sb_tid() == 0) { for (int k = 0; k < P.K; k++) R.ssz[k] = R.st.sizes[k]; R.sred[0] = size_prior_of(P, R.ssz); }
    sb_cta_sync();
    R.pr_size = R.sred[0];
    sb_cta_sync();
    R.pr_geo = 0.0;
    if (P.geo) {
        // every CTA computes all clusters (identical results); rank 0 stores them
        for (int k = 0; k < P.K; k++) {
            double a, g; geo_cluster(R, k, 0, a, g);
            if (R.rank == 0 && sb_tid() == 0) { R.st.geo_agg[0][k] = a; R.st.geo_val[0][k] = g; }
            R.pr_geo += g;
        }
    }
    team_sync(R);
}

// -----------------------------------------------------------------------------------------
// one MCMC step (MCMCChain.step, sbayes/sampling/mcmc_chain.py:186-238)
// -----------------------------------------------------------------------------------------
SB_DEV void mcmc_step(Run &R, sb_step_record *rec) {
    const SbProblem &P = *R.P;
    const long long t_begin = sb_clock();
    R.i_step++;
    R.D.step = (unsigned long long)R.i_step;
    // operator choice (mcmc_chain.py:179-184)
    int op;
    if (R.D.tv) op = (int)R.D.tv[R.D.toff[SITE_OP]];
    else {
        double u = draw_uniform(R.D, SITE_OP), cum = 0.0, tot = 0.0;
        for (int i = 0; i < SB_N_OPS; i++) tot += R.slot->cfg.op_weights[i];
        op = SB_N_OPS - 1;
        for (int i = 0; i < SB_N_OPS; i++) { cum += R.slot->cfg.op_weights[i]; if (cum / tot > u) { op = i; break; } }
    }
    Prop pr;
    const bool geo = P.geo != 0;
    switch (op) {
        case OP_NAIVE_N1: pr = op_alter_cluster(R, false, false, NB_DIRECT); break;
        case OP_NAIVE_N1_GEO: pr = op_alter_cluster(R, false, geo, NB_DIRECT); break;
        case OP_NAIVE_N2_GEO: pr = op_alter_cluster(R, false, geo, NB_TWOSTEP); break;
        case OP_GIBBSISH: pr = op_alter_cluster(R, true, false, NB_EVERYWHERE); break;
        case OP_GIBBSISH_GEO: pr = op_alter_cluster(R, true, geo, NB_EVERYWHERE); break;
        case OP_WIDE: pr = op_alter_cluster_wide(R, geo); break;
        case OP_JUMP: pr = op_cluster_jump(R); break;
        case OP_SRC_RANDOM: pr = op_gibbs_source(R, false); break;
        case OP_SRC_GROUPS: pr = op_gibbs_source(R, true); break;
        default: pr = op_gibbs_weights(R); break;
    }
    // candidate posterior (mcmc_chain.py:201-202)
    double ll_new = R.ll, p_size = R.pr_size, p_geo = R.pr_geo, p_w = R.pr_w, p_src = R.pr_src;
    bool touched = pr.changed != 0;
    if (touched && op == OP_WEIGHTS) { p_src = pr.sp_new; p_w = pr.wp_new; }
    else if (touched) {
        double v[1] = {0.0};
        for (int g = sb_tid(); g < P.G_tot; g += sb_nthr()) v[0] += R.st.lh[gsel_of(R, g, 1)][g];
        block_sum<1>(R, v);
        ll_new = R.sfp ? 0.0 : v[0];
        p_src = R.st.sc[SC_SP_NEW];
        if (P.geo) p_geo = R.st.sc[SC_GEO_NEW];
        if (sb_tid() == 0) R.sred[0] = size_prior_of(P, R.ssz);
        sb_cta_sync();
        p_size = R.sred[0];
        sb_cta_sync();
    }
    double prior_new = p_size + p_geo + p_w + p_src;
    double lq = pr.lq, lqb = pr.lqb;
    if (pr.force_reject) { lq = 0.0; lqb = -SB_INF; }
    // Metropolis-Hastings (mcmc_chain.py:204-217,240-255)
    bool accept; double ratio = NAN, log_u = NAN;
    if (!touched || lqb == -SB_INF) accept = false;
    else if (lq == -SB_INF) accept = true;
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
    if (R.rank == 0 && sb_tid() == 0) {
        if (accept) R.slot->accepts[op]++; else R.slot->rejects[op]++;
        R.slot->cycles[op] += sb_clock() - t_begin;
    }
    if (touched && op != OP_WEIGHTS) {
        team_sync(R);     // every CTA has finished reading the candidate view
        const int F = P.F, C = P.C;
        for (int m = R.rank; m < pr.M; m += R.tsize) {
            int o = R.st.list_a[m];
            if (!accept) {
                // undo the histogram update
                int pat = sb_ldg(P.pat_id + o), z_old = R.st.cluster_of[o] >= 0, z_new = R.st.cl_new[o] >= 0;
                const uint8_t *xr = P.x + (size_t)o * P.ldx;
                const uint8_t *so = R.st.src[R.st.osel[o]] + (size_t)o * P.ldx;
                const uint8_t *sn = R.st.src[R.st.osel[o] ^ 1] + (size_t)o * P.ldx;
                for (int f = sb_tid(); f < F; f += sb_nthr()) {
                    if (sb_ldg(xr + f) == SB_NA_U8) continue;
                    if (so[f] == sn[f] && z_old == z_new) continue;
                    sb_atomic_add_i32(R.st.H + ((size_t)(pat * 2 + z_new) * F + f) * C + sn[f], -1);
                    sb_atomic_add_i32(R.st.H + ((size_t)(pat * 2 + z_old) * F + f) * C + so[f], 1);
                }
            }
            sb_cta_sync();
            if (sb_tid() == 0) {
                if (accept) { R.st.osel[o] ^= 1; R.st.cluster_of[o] = R.st.cl_new[o]; }
                else R.st.cl_new[o] = R.st.cluster_of[o];
                R.st.ochg[o] = 0;
            }
        }
        if (R.rank == 0) {
            for (int g = sb_tid(); g < P.G_tot; g += sb_nthr())
                if (R.st.dirty[g]) { if (accept) R.st.gsel[g] ^= 1; R.st.dirty[g] = 0; }
            if (accept) for (int k = sb_tid(); k < P.K; k += sb_nthr()) R.st.sizes[k] = R.ssz[k];
        }
    }
    if (accept) {
        R.ll = ll_new; R.pr_size = p_size; R.pr_geo = p_geo; R.pr_w = p_w; R.pr_src = p_src;
        if (!(isfinite(R.ll) && isfinite(prior_new))) R.err = 1;     // mcmc_chain.py:231-236
    }
    team_sync(R);
}

// -----------------------------------------------------------------------------------------
// the persistent kernel: every team advances its chains by n_steps
// -----------------------------------------------------------------------------------------
SB_KERNEL sb_run_kernel(SbLaunch L) {
    Run R;
    R.P = &L.P;
    R.tsize = L.team; R.rank = sb_cta() % L.team; R.par = 0; R.epoch = 0;
    const int team_id = sb_cta() / L.team, n_teams = sb_ncta() / L.team;
    R.bar = L.bars + team_id;
    R.tred = L.tred + (size_t)team_id * 2 * L.team * SB_REDW;
    char *sm = sb_smem();
    R.sred = (double *)sm; sm += 32 * SB_REDW * sizeof(double);
    R.smem_best = (double *)sm; sm += (size_t)L.smem_members * sizeof(double);
    R.sint = (int *)sm; sm += 96 * sizeof(int);
    R.ssz = (int *)sm; sm += (size_t)L.P.K * sizeof(int);
    R.smem_mem = (int *)sm;
    R.smem_cap = L.smem_members;
    for (int s = team_id; s < L.n_slots; s += n_teams) {
        R.slot = L.slots + s;
        if (!R.slot->active) continue;
        R.st = L.states[R.slot->state];
        const sb_chain_cfg &cfg = R.slot->cfg;
        R.T = cfg.temperature; R.Tp = cfg.prior_temperature; R.invT = 1.0 / R.T; R.invTp = 1.0 / R.Tp;
        R.p_grow = cfg.p_grow; R.sfp = cfg.sample_from_prior;
        R.D.k0 = (uint32_t)(cfg.seed & 0xFFFFFFFFull); R.D.k1 = (uint32_t)(cfg.seed >> 32);
        R.D.tv = L.tape.vals; R.D.toff = nullptr; R.D.tlen = nullptr; R.D.wide_n = 0; R.D.step = 0;
        R.err = 0;
        R.i_step = (long long)R.st.post[5];
        rebuild_state(R);
        for (long long it = 0; it < L.n_steps; it++) {
            if (L.tape.vals) {
                R.D.toff = L.tape.off + it * SB_N_SITES; R.D.tlen = L.tape.len + it * SB_N_SITES;
                R.D.wide_n = L.tape.wide_rows ? L.tape.wide_rows[it] : 0;
            }
            mcmc_step(R, L.records ? L.records + (size_t)s * L.n_steps + it : nullptr);
        }
        if (R.rank == 0 && sb_tid() == 0) {
            R.st.post[0] = R.ll; R.st.post[1] = R.pr_size; R.st.post[2] = R.pr_geo; R.st.post[3] = R.pr_w;
            R.st.post[4] = R.pr_src; R.st.post[5] = (double)R.i_step;
            if (R.err) R.slot->err = R.err;
        }
        team_sync(R);
    }
}

// export of the committed state (source rows and counts sit in either copy)
SB_KERNEL sb_export_kernel(SbProblem P, SbState st, uint8_t *src_out, int32_t *counts_out) {
    const size_t nS = (size_t)P.N * P.ldx, nG = (size_t)P.G_tot * P.F * P.S;
    const size_t t0 = (size_t)sb_cta() * sb_nthr() + sb_tid(), stride = (size_t)sb_ncta() * sb_nthr();
    for (size_t i = t0; i < nS; i += stride) { int o = (int)(i / P.ldx); src_out[i] = st.src[st.osel[o]][i]; }
    for (size_t i = t0; i < nG; i += stride) { int g = (int)(i / ((size_t)P.F * P.S)); counts_out[i] = st.counts[st.gsel[g]][i]; }
}
