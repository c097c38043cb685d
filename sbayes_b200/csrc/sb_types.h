// sb_types.h -- device-side image of the packed problem and of one chain's state in HBM.
// Layout rationale: DESIGN.md "Data layout in HBM".
#pragma once
#include <stdint.h>
#include "../../include/sbayes_b200.h"

#define SB_MAXC SB_MAX_COMPONENTS
#define SB_MAXK 1024         // clusters (sizes are staged in shared memory)
#define SB_REDW 8            // doubles per CTA in a team reduction
#define SB_NA_U8 255

// draw sites (oracle/draws.py)
enum {
    SITE_OP = 0, SITE_MH, SITE_CLUSTER, SITE_GROW, SITE_OBJECT, SITE_SUBSET, SITE_COMPONENT, SITE_GIBBS,
    SITE_WIDE, SITE_JUMP, SITE_WPAIR, SITE_BETA, SITE_WACCEPT, SITE_SWAP
};
// operators (sbayes/sampling/operators.py:50-205)
enum {
    OP_NAIVE_N1 = 0, OP_NAIVE_N1_GEO, OP_NAIVE_N2_GEO, OP_GIBBSISH, OP_GIBBSISH_GEO, OP_WIDE, OP_JUMP,
    OP_SRC_RANDOM, OP_SRC_GROUPS, OP_WEIGHTS
};
enum { NB_EVERYWHERE = 0, NB_DIRECT = 1, NB_TWOSTEP = 2 };

struct SbProblem {
    int N, F, S, K, C, n_conf, G_tot, ldx, n_pat;
    int goff[SB_MAXC + 1];       // first stacked group row of every component
    int conf_ng[SB_MAXC];        // groups per confounder (index c-1)
    int min_size, max_size, size_prior, wprior_uniform;
    int geo, geo_agg, geo_pf;
    double geo_rate, geo_infl;
    const uint8_t *x;            // [N, ldx]  state index, 255 = NA
    const int32_t *conf_gid;     // [n_conf, N]
    const uint8_t *pat_id;       // [N] index of the object's confounder has-pattern
    const int32_t *pat_bits;     // [n_pat] bit c set <=> object is in a group of component c (c >= 1)
    const double *alpha;         // [G_tot, F, S]
    const double *asum;          // [G_tot, F]  sum_s alpha
    const double *dcc;           // [G_tot, F]  lgamma(sum alpha) - sum_{alpha>0} lgamma(alpha)
    const double *unif;          // [F, S] 1.0 on applicable states
    const double *alpha_w;       // [F, C]
    const double *wlnB;          // [F] log multivariate beta of alpha_w[f, :]
    const double *cost;          // [N, N] or null
    const int32_t *adj_ptr;      // [N+1] or null
    const int32_t *adj_idx;
};

// Everything one chain state owns in HBM.  Group-indexed rows (counts, tables, lh, geo) and
// object-indexed rows (src) exist twice; `gsel[g]` / `osel[o]` say which copy is the
// committed one, the other copy receives the candidate of the running proposal and a
// commit is a flip of the selector (no copy, free rollback).
struct SbState {
    int32_t *cluster_of;         // [N] committed membership, -1 = none
    int32_t *cl_new;             // [N] candidate membership (== cluster_of outside a proposal)
    int32_t *sizes;              // [K] committed cluster sizes
    float *W;                    // [F, C]
    uint8_t *src[2];             // [N, ldx]
    uint8_t *osel;               // [N]
    uint8_t *ochg;               // [N] object is part of the running proposal
    int32_t *counts[2];          // [G_tot, F, S]
    double *tabP[2];             // [G_tot, F, S] plain tables   norm(counts + alpha)
    double *tabT[2];             // [G_tot, F, S] tempered tables norm(u + (alpha-u)/Tp + counts/T)
    double *lh[2];               // [G_tot] per-group log-likelihood
    double *geo_agg[2];          // [K] aggregated MST cost of every cluster
    double *geo_val[2];          // [K] geo prior of every cluster
    uint8_t *gsel;               // [G_tot]
    uint8_t *dirty;              // [G_tot] group row is being re-proposed
    int32_t *dneg, *dpos;        // [G_tot, F, S] observations removed / added by the proposal
    int32_t *H;                  // [n_pat*2, F, C] #objects with has-pattern (p, z) whose source at f is c
    double *w1, *wp, *wq, *lw;   // [n_pat*2, F, C] weight tables (see sb_core.cuh: weight_tables)
    double *val;                 // [N] proposal probabilities
    double *red_a, *red_b;       // [N] per-object partial sums
    int32_t *list_a, *list_b;    // [N] compacted object lists
    uint8_t *flag_a, *flag_b;    // [N]
    double *sc;                  // [64] scalars published inside a step
    double *post;                // [8] committed posterior terms {ll, size, geo, weights, source, i_step}
};

struct SbSlot {
    sb_chain_cfg cfg;
    long long accepts[SB_N_OPS], rejects[SB_N_OPS];
    long long cycles[SB_N_OPS];  // SM clock cycles spent in each operator's steps (rank-0 CTA)
    int state;                   // which SbState currently sits at this temperature
    int err;
    int active;                  // a state has been uploaded
    int pad_;
};

struct SbLaunch {
    SbProblem P;
    SbState *states;
    SbSlot *slots;
    int n_slots, team;
    long long n_steps;
    unsigned long long *bars;    // [n_teams] barrier counters (zeroed before launch)
    double *tred;                // [n_teams, 2, team, SB_REDW]
    sb_tape tape;                // device pointers; vals == null => Philox
    sb_step_record *records;     // [n_slots, n_steps] or null
    int smem_members;            // capacity of the shared-memory member list (geo prior)
};
