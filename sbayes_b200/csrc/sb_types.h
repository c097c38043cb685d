// sb_types.h -- device-side image of the packed problem and of one chain's state in HBM.
// Layout rationale: DESIGN.md "Data layout in HBM".
#pragma once
#include <stdint.h>
#include "../../include/sbayes_b200.h"

#ifndef SB_MAXC
#define SB_MAXC SB_MAX_COMPONENTS
#endif
#define SB_MAXK 1024         // clusters (sizes are staged in shared memory)
#define SB_MAX_RANKS 8       // GPUs one chain's objects can be sharded over
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
    // concentration rows: row of group g is arow(g); with dedup the K cluster rows (identical
    // by construction, sbayes/model/prior.py:438-455) are stored once
    int dedup, arow_off, A_rows, use_mask;
    const double *alpha;         // [A_rows, F, S]
    const double *asum;          // [A_rows, F]  sum_s alpha
    const double *dcc;           // [A_rows, F]  lgamma(sum alpha) - sum_{alpha>0} lgamma(alpha)
    const double *unif;          // [F, S] 1.0 on applicable states
    const uint64_t *applm;       // [F] the same as a bit mask (S <= 64)
    const double *alpha_w;       // [F, C]
    const double *wlnB;          // [F] log multivariate beta of alpha_w[f, :]
    const double *cost;          // [N, N] or null
    const int32_t *adj_ptr;      // [N+1] or null
    const int32_t *adj_idx;
    const int32_t *cg_ptr;       // [sum conf groups + 1] CSR: members of every confounder group (sorted)
    const int32_t *cg_idx;
    // distinct tuples of confounder groups ("combos"): objects of one combo share every
    // confounder table row, so proposal passes can gather from per-combo tables
    int n_combos, use_tables;
    const int32_t *combo_id;     // [N]
    const int32_t *combo_rows;   // [n_combos, n_conf] stacked group row or -1
    const int32_t *combo_pat;    // [n_combos] has-pattern index
    // one word per object for the streaming kernels (null unless C <= 4, G_tot < 255, n_pat < 256):
    // byte 0 = has-pattern index, bytes 1..3 = stacked group row of components 1..3 (255 = none)
    const uint32_t *obj_pack;
    // objects sorted by combo (stable): the tiled proposal passes walk `perm` so that the objects a CTA
    // processes share few combos and only those combos' pass tables are built, tile by tile, in shared memory
    const int32_t *perm;         // [N]
    const int32_t *combo_start;  // [n_combos + 1] first position of every combo in perm
    // two features per byte for the ratio pass of AlterCluster: xp[o, i] = x[o, 2i] * (S+1) + x[o, 2i+1] with
    // a missing value coded as state S (null unless (S+1)^2 <= 256); rows padded to 16 bytes with the
    // code of two missing values
    const uint8_t *xp;           // [N, ldp]
    int ldp;
    // lfact[i] = lgamma(i + 1) for i < lfact_n: with unit concentrations (uniform priors, the default)
    // the Dirichlet-categorical terms of the evaluation are table look-ups instead of lgamma calls
    const double *lfact;
    int lfact_n;
};

// ---------------------------------------------------------------------------------------
// State of one chain.
//
// (1) SbShared: what the whole team of CTAs that advances the chain shares, in HBM: the
//     object-indexed arrays.  Source rows exist twice; osel[o] says which copy is committed,
//     the other one receives the candidate of the running proposal (commit = set osel).
// (2) Rep (sb_core.cuh): every CTA of the team keeps a *private replica* of the
//     group-indexed state (counts, per-(group, feature) totals / reciprocals /
//     log-likelihood terms, weight tables, the (pattern, feature, component) histogram,
//     cluster sizes, the per-combo pass tables).  Replicas live in shared memory where they
//     fit and in a private slice of HBM otherwise (SbPlace, decided by the host).  All CTAs
//     apply the same proposal to their replica *in place* and undo it on a reject, so
//     nothing group-indexed ever has to be exchanged inside a step; only sums over objects
//     cross CTAs.
// ---------------------------------------------------------------------------------------
struct SbShared {
    int32_t *cluster_of;         // [N] committed membership, -1 = none
    uint8_t *src[2];             // [N, ldx]
    uint8_t *osel;               // [N]
    float *W;                    // [F, C]   import / export copy of the weights
    int32_t *counts;             // [G_tot, F, S] export copy of the committed counts; rebuild scratch
    int32_t *H;                  // [n_pat*2, F, C] rebuild scratch of the histogram
    int32_t *sizes;              // [K] rebuild scratch
    double *val[2];              // [N] proposal probabilities (alternating per pick)
    int32_t *list_a, *list_b;    // [N] compacted object lists
    uint8_t *flag_a, *flag_b;    // [N]
    double *post;                // [8] committed posterior terms {ll, size, geo, weights, source, i_step}
    double *glh;                 // [G_tot] export copy of the per-group log-likelihoods (ModelCache.group_likelihoods)
    double *geo_out;             // [K] export copy of the per-cluster geo prior (ModelCache.geo_prior)
};

// staging buffers of sb_export_samples: the samples of n chains in the layout of the host buffers
struct SbExport {
    int32_t *cluster_of;         // [n, N]
    float *W;                    // [n, F, C]
    uint8_t *src;                // [n, N, F] committed sources, rows unpadded
    int32_t *counts;             // [n, G_tot, F, S]
    double *post;                // [n, 8]
    double *glh;                 // [n, G_tot]
    double *geo;                 // [n, K]
    double *sp_obj;              // [n, N] per-object source prior
    const int32_t *state_idx;    // [n] which state each requested chain runs
    int n;
    unsigned want;               // bit mask of the arrays to fill (enum order above: bit 0 = cluster_of ...)
};

// replica arrays that may live in shared memory (placement priority = enum order)
// (the small per-object lookups first: every item of every pass reads them)
enum { A_MCL = 0, A_MCOMBO, A_MOSEL, A_COMBOT, A_APPLM, A_TAB0, A_TAB1, A_COUNTS, A_NN, A_RT, A_RP, A_WQ, A_ALPHA, A_XBLK,
       A_WP, A_W1, A_LW, A_H0, A_H1, A_DCNT, A_LLGF, A_ASUM, A_BST, A_BK_LLGF, A_BK_RP, A_BK_RT, A_RS, A_N };

// a pair of doubles read with one 16-byte access (the two pass tables, interleaved)
struct alignas(16) SbD2 { double a, b; };

struct SbPlace {
    int in_smem[A_N];
    unsigned smem_off[A_N];      // byte offset inside the dynamic shared memory (if in_smem)
    size_t glob_off[A_N];        // byte offset inside the CTA's private HBM slice
    size_t sel_off;              // selbuf [N]
    size_t rep_bytes;            // size of one CTA's private HBM slice
    unsigned fixed_off;          // start of the always-in-smem small arrays
    unsigned smem_bytes;         // total dynamic shared memory
    int xblk_rows;               // objects per CTA block cached in A_XBLK (0 = not cached)
    int alias_w;                 // every chain has prior temperature 1: wp == w1 == wq, stored once
    // tiled passes: the per-combo pass tables of the whole problem do not fit in shared memory; A_TAB0 is
    // then a tile buffer of `tile_doubles` doubles in shared memory that holds the tables of a few combos
    // at a time (sb_core.cuh: tiled_pass), and nothing of it lives in the HBM slice
    int tiled;
    unsigned tile_doubles;
};

struct SbSlot {
    sb_chain_cfg cfg;
    long long accepts[SB_N_OPS], rejects[SB_N_OPS];
    long long cycles[SB_N_OPS];  // SM clock cycles spent in each operator's steps (rank-0 CTA)
    int state;                   // which SbShared currently sits at this temperature
    int err;
    int active;                  // a state has been uploaded
    int pad_;
};

struct SbLaunch {
    SbProblem P;
    SbShared *states;
    SbSlot *slots;
    int n_slots, team;
    long long n_steps;
    unsigned long long *bars;    // [n_teams] barrier counters (zeroed before launch)
    double *tred;                // [n_teams, 2, team, SB_REDW]
    char *rep_base;              // [grid] private HBM slices of the CTAs
    size_t rep_stride;
    SbPlace place;
    sb_tape tape;                // device pointers; vals == null => Philox
    sb_step_record *records;     // [n_slots, n_steps] or null
    int smem_members;            // capacity of the shared-memory member list (geo prior)
    int use_cluster;             // team == thread-block cluster: hardware cluster barrier
    int cta_base, lgrid;         // first CTA and number of CTAs of this launch (cta_base != 0 only in emulated multi-rank launches)
    // ---- objects of a chain sharded over the GPUs of one box (sb_shard_attach; xworld == 1: off) ----
    // A chain is advanced by xworld * team CTAs: `team` on every GPU.  The object-indexed state of the chain
    // (states[]) lives in rank 0's memory and is reached by the other ranks through peer mappings (NVLink);
    // the reduction buffer is replicated (every CTA stores its slot into every rank's copy), the team barrier
    // is hierarchical (L2 ticket per GPU, then one flag per rank pair in peer memory), the bulk statistics
    // are reduced rank-local first and exchanged through peer memory.
    int xrank, xworld;
    unsigned long long xepoch_base;                 // flag values of this launch start above it (flags are never reset)
    double *tred_peer[SB_MAX_RANKS];                // every rank's copy of tred
    unsigned long long *xflag_peer[SB_MAX_RANKS];   // every rank's flag words [n_teams][SB_MAX_RANKS]
    unsigned long long *lrelease;                   // [n_teams] release word of the local barrier phase
    char *arena_peer[SB_MAX_RANKS];                 // every rank's arena (same layout everywhere: an array is found at the same offset)
    int32_t *lstat_peer[SB_MAX_RANKS];              // every rank's partial statistics [n_chains][stats_stride]
    int32_t *gstat_peer[SB_MAX_RANKS];              // every rank's copy of the reduced statistics
    size_t stats_stride;
    int32_t *lcl;                                   // [n_chains][N] rank-local replica of cluster_of
    uint8_t *losel;                                 // [n_chains][N] rank-local replica of osel
    int *xabort;                                    // set when a cross-GPU wait timed out: every later barrier falls through
};
