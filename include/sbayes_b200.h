/* sbayes_b200.h -- C ABI of the B200 backend for the sBayes MCMC inner loop.
 *
 * One `sb_ctx` owns one GPU: the packed problem (shared by all chains), the per-chain
 * state buffers in HBM, one CUDA stream and the persistent step kernel.  Every entry
 * point returns 0 on success and a negative code on failure (`sb_last_error` gives the
 * text); nothing throws across the boundary.  The caller owns all host buffers; the ctx
 * owns all device memory.  One host thread per ctx.
 *
 * Which reference interface each entry point stands in for (paths are into the sBayes
 * reference tree, v1.1):
 *
 *   sb_create            Data + Model construction as seen by the chain
 *                        (sbayes/load_data.py:85-256, sbayes/model/model.py:24-45,
 *                         sbayes/model/prior.py:155-255,325-354,438-455,704-853)
 *   sb_chain_config      MCMCChain.__init__ (sbayes/sampling/mcmc_chain.py:30-72) and
 *                        get_operator_schedule (sbayes/sampling/operators.py:28-209)
 *   sb_set_state         Sample.from_numpy_arrays + recalculate_feature_counts
 *                        (sbayes/sampling/state.py:492-552, sbayes/sampling/counts.py:35-52)
 *   sb_get_state         reading Sample.clusters/weights/source/feature_counts
 *                        (sbayes/sampling/loggers.py:157-301,426-442)
 *   sb_export_samples    the same for several chains in one call, plus the values of the ModelCache nodes the
 *                        loggers read: cluster_size_prior, geo_prior[K], source_prior[N], weights_prior,
 *                        group_likelihoods (sbayes/sampling/state.py:396-449, loggers.py:249-252)
 *   sb_eval              Model.likelihood(sample, caching=False) + Model.prior(sample,
 *                        caching=False) (sbayes/model/model.py:47-51,
 *                        sbayes/model/likelihood.py:47-101, sbayes/model/prior.py:74-88)
 *   sb_eval_batch/_stats/_terms  the same for every chain at once as stand-alone kernels
 *                        (recalculate_feature_counts sbayes/sampling/counts.py:35-52,
 *                         Likelihood.__call__ likelihood.py:47-101, Prior.__call__ prior.py:74-88)
 *   sb_run               MCMCChain.run / MCMCChain.step for every chain of the ctx
 *                        (sbayes/sampling/mcmc_chain.py:128-255) including the ten scheduled
 *                        operators (sbayes/sampling/operators.py:423-1827)
 *   sb_posterior         MCMCChain._ll / _prior and Sample.last_lh / last_prior
 *                        (sbayes/sampling/mcmc_chain.py:87,108)
 *   sb_posteriors / sb_set_temperatures   the scalars swap_chains reads and the temperatures it exchanges,
 *                        for all chains of the ctx in one call (sbayes/mcmc_setup.py:263-269,389-403)
 *   sb_gibbs_sweep       GibbsSampleSource._propose with ObjectSelector.ALL (sbayes/sampling/operators.py:496-574)
 *   sb_init_em           SbayesInitializer.generate_clusters_em (sbayes/sampling/initializers.py:93-169)
 *   sb_op_stats          Operator.accepts / rejects / step_times (sbayes/sampling/operators.py:278-293)
 *   sb_shard_handle / sb_shard_attach / sb_run_group   one chain advanced by several GPUs, objects sharded
 *                        (the sums over objects of the proposals cross GPUs inside the step kernel)
 *   sb_swap              swap_chains (sbayes/mcmc_setup.py:365-418) over the chains of one ctx
 *   sb_swap_scalars      the same decision procedure on host scalars (no ctx): used when the
 *                        temperature ladder spans several GPUs / processes, and to replay the
 *                        reference's recorded pairs and uniforms (tests/golden/mc3_swap.npz)
 *   sb_counts_rebuild_mma  recalculate_feature_counts as an int8 tensor-core contraction
 *                        (sbayes/sampling/counts.py:10-52)
 */
#ifndef SBAYES_B200_H
#define SBAYES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_N_OPS 10      /* scheduled operators, order of sbayes/sampling/operators.py:50-205 */
#define SB_N_SITES 14    /* random-draw sites of one step (see sb_tape) */
#define SB_MAX_COMPONENTS 8
#define SB_NA 255

typedef struct sb_ctx sb_ctx;

/* Packed problem.  All pointers are host memory and are copied by sb_create. */
typedef struct sb_problem_desc {
    int32_t n_objects, n_features, n_states, n_clusters, n_confounders;
    const uint8_t *x;             /* [N, F]   state index, 255 = missing            */
    const uint8_t *applicable;    /* [F, S]   1 where the state exists for f        */
    const int32_t *conf_gid;      /* [n_conf, N] group of each object, -1 = none    */
    const int32_t *conf_ngroups;  /* [n_conf]                                        */
    const double *alpha;          /* [G_tot, F, S] Dirichlet concentration, rows =   *
                                   *   K cluster rows, then every confounder's groups */
    const double *alpha_w;        /* [F, C] weights-prior concentration              */
    int32_t min_size, max_size;
    int32_t size_prior;           /* 0 uniform_area, 1 uniform_size, 2 quadratic     */
    int32_t weights_prior_uniform;
    int32_t geo_enabled;          /* cost_based geo prior with MST skeleton          */
    int32_t geo_aggregation;      /* 0 mean, 1 sum, 2 max                            */
    int32_t geo_prob_fn;          /* 0 exponential, 1 sigmoid                        */
    double geo_rate, geo_inflection;
    const double *geo_cost;       /* [N, N] or NULL                                  */
    const int32_t *adj_indptr;    /* [N+1] CSR adjacency (may be NULL: no neighbours)*/
    const int32_t *adj_indices;
} sb_problem_desc;

/* sb_chain_cfg.flags.  SB_CFG_PROPOSE_ONLY: the operator API of the reference
 * (Operator.function(sample) -> (sample', log_q, log_q_back), sbayes/sampling/operators.py:269-276):
 * every proposal the operator does not reject itself is committed without the Metropolis-Hastings
 * test, so one step with a one-hot op_weights vector and a step record is one operator call. */
#define SB_CFG_PROPOSE_ONLY 1

/* Per-chain run parameters. */
typedef struct sb_chain_cfg {
    double temperature, prior_temperature;
    int32_t sample_from_prior;
    int32_t flags;                /* SB_CFG_* */
    double op_weights[SB_N_OPS];  /* normalised operator probabilities               */
    double p_grow;
    uint64_t seed;                /* Philox key of the chain                         */
} sb_chain_cfg;

/* Draw tape for replay mode: the draws of every step recorded from the reference
 * (layout of oracle/fixtures.py:flatten_tapes).  NULL tape = Philox streams. */
typedef struct sb_tape {
    const double *vals;           /* [total]                                         */
    const int32_t *off;           /* [n_steps, SB_N_SITES] offset into vals, -1 = absent */
    const int32_t *len;           /* [n_steps, SB_N_SITES]                           */
    const int32_t *wide_rows;     /* [n_steps] n_avail of the wide operator's step   */
    int64_t n_vals;
} sb_tape;

/* One record per chain and step (optional output of sb_run). */
typedef struct sb_step_record {
    int32_t op, accept, changed, n_changed;
    double log_q, log_q_back, ll_new, prior_new, mh_ratio, log_u;
} sb_step_record;

/* posterior terms of a chain, un-tempered */
typedef struct sb_posterior {
    double ll, prior, size_prior, geo_prior, weights_prior, source_prior;
} sb_posterior_t;

/* host buffers of sb_export_samples for n chains; any pointer may be NULL */
typedef struct sb_sample_out {
    int32_t *cluster_of;      /* [n, N] */
    float *W;                 /* [n, F, C] */
    uint8_t *src;             /* [n, N, F] */
    int32_t *counts;          /* [n, G_tot, F, S] */
    double *post;             /* [n, 8] {ll, size prior, geo prior, weights prior, source prior, i_step, 0, 0} */
    double *group_ll;         /* [n, G_tot] ModelCache.group_likelihoods, stacked like the counts */
    double *geo_prior;        /* [n, K]     ModelCache.geo_prior */
    double *source_prior;     /* [n, N]     ModelCache.source_prior */
} sb_sample_out;

int sb_create(const sb_problem_desc *desc, int device, int n_chains, int team_ctas, sb_ctx **out);
void sb_destroy(sb_ctx *ctx);
const char *sb_last_error(const sb_ctx *ctx);       /* ctx may be NULL: error of the last failed sb_create */

int sb_chain_config(sb_ctx *ctx, int chain, const sb_chain_cfg *cfg);
int sb_set_state(sb_ctx *ctx, int chain, const int32_t *cluster_of, const float *W, const uint8_t *src,
                 int64_t i_step);
int sb_get_state(sb_ctx *ctx, int chain, int32_t *cluster_of, float *W, uint8_t *src, int32_t *counts);
int sb_posteriors(sb_ctx *ctx, sb_posterior_t *out /* [n_chains] */);   /* sb_posterior of every chain, one copy */
int sb_set_temperatures(sb_ctx *ctx, const double *temps, const double *prior_temps);   /* [n_chains] each */
int sb_export_samples(sb_ctx *ctx, int32_t n, const int32_t *chains, const sb_sample_out *out);   /* batched sb_get_state + cache-node values */
int sb_eval(sb_ctx *ctx, int chain, sb_posterior_t *out);          /* full uncached evaluation */
int sb_posterior(sb_ctx *ctx, int chain, sb_posterior_t *out);     /* cached values of the current state */
/* Stand-alone full evaluation of EVERY chain of the ctx (HBM-bound statistics pass + terms):
 * sb_eval_batch = sb_eval_stats + sb_eval_terms.  out[n_chains]; ms2 (may be NULL) receives the
 * device time in ms of {statistics pass, terms}.  With the objects of a large matrix sharded over
 * GPUs, every rank runs sb_eval_stats on its shard, all-reduces the leading counts | sizes part of the int32 block
 * returned by sb_stats_buffer (per state: counts | sizes | histogram; layout3 = {n_counts, n_hist, n_sizes}), and calls sb_eval_terms(n_global = objects of the whole problem); the source
 * prior it returns is the shard's partial sum (sbayes_b200/shard.py). */
int sb_eval_batch(sb_ctx *ctx, sb_posterior_t *out, float *ms2);
int sb_eval_stats(sb_ctx *ctx, float *ms);
int sb_eval_set_mode(sb_ctx *ctx, int mode);   /* statistics pass: 0 shared-memory histogram, 1 int8 tcgen05 contraction */
int sb_eval_terms(sb_ctx *ctx, int64_t n_global, sb_posterior_t *out, float *ms);
/* Observation likelihood of every chain: lh_out [n_chains, N, F] float32 = what LikelihoodLogger
 * appends per sample (sbayes/sampling/loggers.py:354-359, leave-one-out component likelihoods of
 * sbayes/sampling/conditionals.py:300-367); obj_ll_out [n_chains, N] / sums_out [n_chains] = sums
 * of log lh per object / chain.  Any output may be NULL. */
int sb_obs_loglik(sb_ctx *ctx, float *lh_out, double *obj_ll_out, double *sums_out, float *ms);
/* Full-sweep source Gibbs of every chain (GibbsSampleSource over ALL objects, sbayes/sampling/operators.py:
 * 496-574; initializer: sbayes/sampling/initializers.py:285-310).  mode 0: resample and commit the source of every
 * observation (Philox draws (chain seed, step)), log_q_out[n_chains] = sum log q[new], n_changed_out[n_chains];
 * mode 1: log_q_out = sum log q[previous sources] under the current counts (log_q_back after a mode-0 sweep);
 * mode 2: the same for the committed sources.  Outputs may be NULL. */
int sb_gibbs_sweep(sb_ctx *ctx, uint64_t step, int mode, double *log_q_out, int64_t *n_changed_out, float *ms);
/* EM-style cluster initializer (SbayesInitializer.generate_clusters_em, sbayes/sampling/initializers.py:93-169):
 * n_steps EM steps on the soft assignment z0 [G_tot, N] float32 (rows: clusters, then every confounder's groups);
 * z_out [G_tot, N] receives the final assignment; z0 must be zero outside a confounder group's members (the reference's
 * start is: initializers.py:108-119).  Uniform geo prior only. */
int sb_init_em(sb_ctx *ctx, int32_t n_steps, const float *z0, float *z_out, float *ms);
/* Which statistics pass sb_eval_stats runs and its launch geometry: out8 = {pass (0 two-increment histogram, 1 int8
 * tcgen05 contraction, 2 fine-row histogram), feature-tile width, feature tiles, object blocks, dynamic shared memory,
 * fine rows, CTAs, block size}. */
int sb_eval_info(sb_ctx *ctx, int32_t *out8);
/* Effect draws at log time (conditional_effect_sample, sbayes/sampling/conditionals.py:125-149, as called by
 * ParametersCSVLogger._write_sample, sbayes/sampling/loggers.py:160-224): for each listed chain
 * out[i, g, f, :] ~ Dirichlet(counts[g, f, :] + concentration[g, f, :]) over the applicable states (zero elsewhere), counts =
 * the chain's committed count tensors; Philox draws (chain seed, step).  out: float64 [n, G_tot, F, S]. */
int sb_effect_sample(sb_ctx *ctx, int32_t n, const int32_t *chains, uint64_t step, double *out, float *ms);
int sb_stats_buffer(sb_ctx *ctx, uint64_t *dev_ptr, int64_t *n_int32, int64_t *stride, int64_t *layout3);
int sb_run(sb_ctx *ctx, int64_t n_steps, const sb_tape *tape, sb_step_record *records);
int sb_run_async(sb_ctx *ctx, int64_t n_steps);                     /* enqueue only; sb_sync waits */
int sb_sync(sb_ctx *ctx);
int sb_op_stats(sb_ctx *ctx, int chain, int64_t *accepts, int64_t *rejects, int64_t *cycles /* may be NULL */);
int sb_swap(sb_ctx *ctx, int attempts, int only_neighbours, uint64_t seed, int64_t round,
            int32_t *state_of_slot, int32_t *n_accepted);
/* ll / prior [n]: un-tempered posterior terms of state s; temps / prior_temps [n]: the ladder (slot i);
 * state_at_slot [n] in/out; tape_pairs [attempts,2] + tape_u [attempts] replay recorded draws (NULL: Philox
 * stream (seed, round)); accepted_pairs [attempts,2] (may be NULL) receives the accepted (from, to) slots. */
int sb_swap_scalars(int32_t n, const double *ll, const double *prior, const double *temps, const double *prior_temps,
                    int32_t attempts, int32_t only_neighbours, uint64_t seed, int64_t round, const int32_t *tape_pairs,
                    const double *tape_u, int32_t *state_at_slot, int32_t *accepted_pairs, int32_t *n_accepted);
/* Objects of a chain sharded over the GPUs of one box (one ctx per GPU, each created with the SAME whole
 * problem, chain count and team size): sb_shard_handle gives the 128-byte handle of a ctx's arena (the ranks
 * exchange them, e.g. torch.distributed all_gather); sb_shard_attach(rank, world, handles[world * 128]) maps the
 * peers' arenas (CUDA IPC between processes, peer access inside one process).  Afterwards rank 0's chain state IS
 * the state of the chains (sb_set_state / sb_get_state / sb_export_samples on rank 0; ranks > 0 call
 * sb_set_state(chain, NULL, NULL, NULL, 0) to mark the chain active), and sb_run / sb_run_async / sb_eval are
 * COLLECTIVE: every rank calls them with the same arguments and the kernels meet on the device (team barriers and
 * reductions through NVLink peer memory, no host collective per step).  sb_run_group launches all ranks of one
 * process together.  Replaces the per-step sums over objects of sbayes/sampling/operators.py:1035-1073,496-552
 * and sbayes/model/prior.py:596-611 for matrices too large for one GPU's pass. */
int sb_shard_handle(sb_ctx *ctx, void *out128);
int sb_shard_attach(sb_ctx *ctx, int rank, int world, const void *handles);
int sb_run_group(sb_ctx **ctxs, int n, int64_t n_steps, sb_step_record *records_of_rank0);
int sb_get_slots(sb_ctx *ctx, int32_t *state_of_slot);
int sb_set_slots(sb_ctx *ctx, const int32_t *state_of_slot);
int sb_counts_rebuild_mma(sb_ctx *ctx, int chain, int32_t *counts_out, float *ms);
int sb_counts_rebuild_hist(sb_ctx *ctx, int chain, int32_t *counts_out, float *ms);
int sb_info(const sb_ctx *ctx, int32_t *out8);     /* {device, n_chains, team, grid, block, smem, n_sm, launches} */
uint64_t sb_stream(const sb_ctx *ctx);              /* cudaStream_t of the ctx (for event timing) */
int64_t sb_clock_khz(const sb_ctx *ctx);            /* SM clock rate (kHz) behind the cycle counters of sb_op_stats */

#ifdef __cplusplus
}
#endif
#endif
