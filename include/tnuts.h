/*
 * tnuts.h -- C ABI of the B200-native many-chain HMC/NUTS engine (libtnuts.so, sm_100a).
 *
 * This is the drop-in boundary for the ONE hot path of Turing.jl that the engine replaces:
 * the leapfrog integrator + log-density gradient + NUTS tree + warm-up adaptation that
 * `sample(model, NUTS(delta), MCMCThreads(), N, nchains)` runs once per chain on a CPU task.
 * Julia binds these symbols with `ccall` from ext/TuringCUDANUTSExt.jl;
 * the Python host mirror in turing.jl_b200/ binds the very same symbols with ctypes.
 *
 * Conventions
 *   - plain C99 types only; every pointer is a HOST pointer unless its name ends in _dev;
 *   - the caller owns every buffer it passes; the library copies model data to the device in
 *     tnuts_set_model and never keeps a host pointer; outputs are caller-allocated;
 *   - every entry returns 0 on success, <0 on failure (TNUTS_E_*); the message is available
 *     from tnuts_last_error(); no C++ exception crosses this boundary;
 *   - one engine handle == one GPU == one caller thread (not re-entrant).
 *   - there is NO CPU fallback: tnuts_create fails with TNUTS_E_CUDA when no sm_100 device
 *     is usable.
 *
 * Reference interfaces replaced (paths into the Turing.jl v0.46.1 tree):
 *   tnuts_create / tnuts_config ..... NUTS/HMC/HMCDA structs + ctors   src/mcmc/hmc.jl:59-80,285-327,352-402
 *                                     make_ahmc_kernel                 src/mcmc/hmc.jl:425-441
 *                                     AHMCAdaptor                      src/mcmc/hmc.jl:447-469
 *   tnuts_set_model ................. DynamicPPL.LogDensityFunction(model, getlogjoint_internal,
 *                                     LinkAll(); adtype)               src/mcmc/hmc.jl:170-176
 *   tnuts_init ...................... init step: find_initial_params_ldf, phasepoint,
 *                                     find_good_stepsize               src/mcmc/hmc.jl:156-208,
 *                                                                      src/mcmc/abstractmcmc.jl:42-70
 *   tnuts_run ....................... the mcmcsample loop over the iterate step:
 *                                     AHMC.transition + AHMC.adapt! + ParamsWithStats
 *                                                                      src/mcmc/hmc.jl:210-252, :85-154
 *   tnuts_fetch ..................... bundle_samples input (params, logprior/loglikelihood/
 *                                     logjoint, sampler stats)         src/mcmc/hmc.jl:247
 *   tnuts_logdensity_and_gradient ... LogDensityProblems.logdensity_and_gradient(ldf, theta)
 *                                                                      src/mcmc/hmc.jl:181
 *   tnuts_logdensity ................ LogDensityProblems.logdensity    src/mcmc/hmc.jl:180
 *   tnuts_dimension ................. LogDensityProblems.dimension     src/mcmc/hmc.jl:179
 *   tnuts_get_stepsize .............. getstepsize                      src/mcmc/hmc.jl:412-418
 *   tnuts_get_state / set_state ..... HMCState + loadstate / initial_state
 *                                                                      src/mcmc/hmc.jl:10-23,135-152
 *   tnuts_num_divergences ........... post_sample_hook                 src/mcmc/hmc.jl:476-486
 *   tnuts_diagnostics ............... ess(chain; kind=:bulk), rhat(chain) of MCMCChains
 */
#ifndef TNUTS_H
#define TNUTS_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNUTS_ABI_VERSION 1

/* status codes */
#define TNUTS_OK 0
#define TNUTS_E_ARG (-1)      /* bad argument / call order                         */
#define TNUTS_E_CUDA (-2)     /* CUDA runtime failure or no sm_100 device          */
#define TNUTS_E_INIT (-3)     /* no finite initial point in 1000 tries             */
#define TNUTS_E_UNSUPPORTED (-4)
#define TNUTS_E_NCCL (-5)

/* registered model families */
#define TNUTS_GDEMO 0          /* test/test_utils/models.jl:15-21                    */
#define TNUTS_LINREG 1         /* README.md:38-47                                    */
#define TNUTS_EIGHT_SCHOOLS 2  /* non-centred eight schools                          */
#define TNUTS_LOGISTIC 3       /* Bayesian logistic regression                       */
#define TNUTS_HIER_LINREG 4    /* hierarchical (varying intercept+slope) regression  */
#define TNUTS_STD_NORMAL 5     /* x ~ Normal() iid                                   */
#define TNUTS_BETA_BERNOULLI 6 /* test/mcmc/hmc.jl:24-43: p ~ Beta(a, b); obs[i] ~ Bernoulli(p);
                                  theta = logit(p); y[N] = observations, hyper = {a, b}      */
#define TNUTS_DIRICHLET_CATEGORICAL 7 /* test/mcmc/hmc.jl:45-62: ps ~ Dirichlet(K1, al1);
                                  pd ~ Dirichlet(K2, al2); obs[i] ~ Categorical(ps).  Each simplex
                                  links by stick breaking: D = K1 - 1 + K2 - 1 linked dimensions,
                                  tnuts_num_params = K1 + K2 chain columns.  y[N] = categories 1..K1,
                                  hyper = {K1, al1, K2, al2}; (K1, K2) in {(2,4), (3,0), (2,0)}   */

/* algorithms (src/mcmc/hmc.jl:425-441) */
#define TNUTS_ALG_NUTS 0
#define TNUTS_ALG_HMC 1
#define TNUTS_ALG_HMCDA 2
/* stochastic-gradient samplers on the batched gradient kernel (src/mcmc/sghmc.jl:79-105, 223-249):
 * no accept step, no adaptation, one gradient evaluation per transition.  They reuse the config
 * fields: SGHMC  init_eps = learning_rate, lambda = momentum_decay;
 *         SGLD   init_eps = a, delta = b, lambda = gamma of PolynomialStepsize a / (t + b)^gamma.
 * Stats columns: n_steps = 1, log_density, step_size (= SGLD_stepsize); the others are NaN. */
/* metricT (src/mcmc/hmc.jl:59-80, 285-327, 352-402): NUTS defaults to DiagEuclideanMetric, HMC and
 * HMCDA to UnitEuclideanMetric.  With the unit metric the StanHMCAdaptor keeps its windows (the
 * dual-averaging restarts at their ends: the NaiveHMCAdaptor branch of hmc.jl:456 is dead code, it
 * compares an instance with a type) but the mass matrix stays the identity.  DenseEuclideanMetric:
 * M^-1 [D][D] per chain, momentum through its Cholesky factor, Welford covariance with the Stan
 * regulariser at the window ends. */
#define TNUTS_METRIC_DIAG 0
#define TNUTS_METRIC_UNIT 1
#define TNUTS_METRIC_DENSE 2   /* NUTS / HMC / HMCDA; thread-per-chain strategy (gdemo, README regression,
                                  eight schools, small standard normal) and, for D <= 256, the pump
                                  strategy (every family; logistic regression incl. "logistic_tc"):
                                  there the state machine runs the unit-metric dynamics in whitened
                                  coordinates phi = U^-T theta, M^-1 = U'U */
#define TNUTS_ALG_SGHMC 3
#define TNUTS_ALG_SGLD 4

typedef struct tnuts_engine tnuts_engine;

/* Model descriptor: what the Julia shim extracts from a recognised DynamicPPL.Model.
 * theta lives in linked (unconstrained) space, variables in model statement order. */
typedef struct {
  int32_t family;
  int32_t D;            /* LogDensityProblems.dimension                              */
  int64_t N;            /* observations (rows)                                       */
  int32_t G;            /* groups (TNUTS_HIER_LINREG)                                */
  const double *X;      /* LOGISTIC: row-major [N][D]; LINREG/HIER: x[N]             */
  const double *y;      /* [N]                                                       */
  const double *sigma;  /* EIGHT_SCHOOLS: [N]                                        */
  const int32_t *group; /* HIER: [N], 0-based                                        */
  double hyper[8];      /* prior hyper-parameters, family specific       */
} tnuts_model;

typedef struct {
  int32_t algorithm;    /* TNUTS_ALG_*                                               */
  double delta;         /* target acceptance                                         */
  int32_t max_depth;    /* default 10                                                */
  double delta_max;     /* default 1000.0                                            */
  double init_eps;      /* 0 => find_good_stepsize                                   */
  int64_t n_adapts;     /* resolved by the host: src/mcmc/hmc.jl:106-110             */
  int32_t n_leapfrog;   /* HMC                                                       */
  double lambda;        /* HMCDA                                                     */
  uint64_t seed;        /* one seed; per-chain Philox streams derive from it         */
  int32_t n_chains;     /* chains resident on this engine                            */
  int32_t chain_offset; /* global id of local chain 0 (chain sharding across GPUs)   */
  int32_t device;       /* CUDA device ordinal                                       */
  int32_t row_shards;   /* >1: data rows sharded over ranks (needs tnuts_comm_init)  */
  int32_t row_rank;
  int32_t strategy;     /* 0 auto, 1 one-thread-per-chain, 2 lock-step pump, 3 one-CTA-per-chain */
  int32_t metric;       /* TNUTS_METRIC_*: metricT of the sampler (src/mcmc/hmc.jl:178-179)  */
  int32_t reserved[4];
} tnuts_config;

#define TNUTS_NSTAT 9
/* sampler stats columns (AdvancedHMC names): n_steps, acceptance_rate, log_density,
 * hamiltonian_energy, hamiltonian_energy_error, max_hamiltonian_energy_error, tree_depth,
 * numerical_error, step_size.  is_accept == true and nom_step_size == step_size for NUTS. */

int tnuts_abi_version(void);
const char *tnuts_last_error(const tnuts_engine *e); /* e may be NULL: last create error */

int tnuts_create(const tnuts_config *cfg, tnuts_engine **out);
int tnuts_destroy(tnuts_engine *e);
int tnuts_set_model(tnuts_engine *e, const tnuts_model *model);
/* same, but X / y / group are DEVICE pointers on the engine's device; they are adopted, not
 * copied (a 12.8 GB row shard is not duplicated): the caller keeps them alive until
 * tnuts_destroy.  Data-parallel families only (logistic, hierarchical). */
int tnuts_set_model_device(tnuts_engine *e, const tnuts_model *model_dev);
int tnuts_dimension(const tnuts_engine *e);
int tnuts_num_params(const tnuts_engine *e);

/* init step for all chains.  theta0: [n_chains][D] row-major linked-space initial
 * parameters, or NULL for InitFromUniform U(-2,2) with <=1000 tries per chain.
 * On TNUTS_E_INIT the message contains the reference's initial-parameters URL. */
int tnuts_init(tnuts_engine *e, const double *theta0);

/* run n_transitions for every chain; keep transitions first_keep, first_keep+thin, ...
 * (1-based within this call; first_keep<=0 keeps nothing).  Kept draws stay on the
 * device until tnuts_fetch.  Returns the number of kept draws in *n_kept (may be NULL). */
int tnuts_run(tnuts_engine *e, int64_t n_transitions, int64_t first_keep, int64_t thin,
              int64_t *n_kept);

/* tnuts_run + the device->host copy of the kept-draw array ([n_kept][ncol][n_chains], see
 * tnuts_fetch_draws) overlapped with the computation: the run is cut into segments of
 * chunk_draws kept draws and every finished segment is copied on a second stream while the
 * next one computes.  host_draws should be page-locked (tnuts_host_alloc) for the overlap. */
int tnuts_run_streamed(tnuts_engine *e, int64_t n_transitions, int64_t first_keep, int64_t thin,
                       double *host_draws, int64_t chunk_draws, int64_t *n_kept);

/* tnuts_run_streamed into a slab of a wider chain array: the host array is
 * [n_kept][ncol][host_chain_stride] and host_draws points at this engine's first chain in it
 * (host_chain_stride >= n_chains).  Several engines (one per GPU) fill ONE
 * (iterations x names x chains) array, the layout `bundle_samples` produces for an ensemble
 * (src/mcmc/abstractmcmc.jl:133-166, AbstractMCMC.mcmcsample with MCMCThreads), each with its
 * own overlapped device->host copies and no host-side reshuffle. */
int tnuts_run_streamed_strided(tnuts_engine *e, int64_t n_transitions, int64_t first_keep, int64_t thin,
                               double *host_draws, int64_t host_chain_stride, int64_t chunk_draws,
                               int64_t *n_kept);

/* copy the draws kept by the last tnuts_run to the host.  Any pointer may be NULL.
 *   params [n_kept][P][n_chains]   constrained (user-space) parameters
 *   logp   [n_kept][3][n_chains]   logprior, loglikelihood, logjoint (no Jacobian)
 *   stats  [n_kept][TNUTS_NSTAT][n_chains]
 *   theta  [n_kept][D][n_chains]   linked-space draws */
int tnuts_fetch(tnuts_engine *e, double *params, double *logp, double *stats, double *theta);

/* the whole kept-draw array in one contiguous copy: [n_kept][ncol][n_chains] with
 * ncol = tnuts_num_columns() = P + 3 + TNUTS_NSTAT (params | logp | stats): this IS the
 * (iterations x names x chains) array of MCMCChains.Chains in C order. */
int tnuts_num_columns(const tnuts_engine *e);
int tnuts_fetch_draws(tnuts_engine *e, double *draws);

/* engine options:
 *   "keep_theta"  (0/1, default 0) keeps linked-space draws for tnuts_fetch
 *   "sort_chains" (0/1, default 1) thread-per-chain strategy: re-sort chains by step size after warm-up
 *   "logistic_tc" (0/1, default 0) TNUTS_LOGISTIC (D <= 256, the family's limit): evaluate the likelihood and its
 *                 gradient (the X*B contraction across chains, src/mcmc/hmc.jl:181) on the tcgen05
 *                 tensor cores with bf16x3 operand planes -- fp32-grade dot products, fp64
 *                 accumulation of the log-likelihood -- instead of the fp64 DMMA kernel.
 *                 Stated bounds (tests/test_gpu_tc.py): |d lp| <= 3e-7 |lp|,
 *                 |d grad| <= 1e-4 |grad|.  Set before tnuts_init.  128 < D <= 256 runs a cluster of
 *                 two CTAs per tile (one half of the features each, partial X*B exchanged through
 *                 distributed shared memory).
 *   "time_kernels" (0/1, default 0) see tnuts_kernel_times
 *   "persist"     (0/1, default 1) "logistic_tc" on one GPU, D <= 128: once <= 512 chains are unfinished, the
 *                 rest of a tnuts_run is ONE persistent cooperative kernel (gradient, reduction and
 *                 NUTS state machine per leapfrog, grid barriers instead of launches); bit-identical
 *                 to the host-pumped path */
int tnuts_set_option(tnuts_engine *e, const char *name, double value);

/* current linked-space position of every chain, [n_chains][D] row-major */
int tnuts_get_theta(tnuts_engine *e, double *theta);
/* constrained parameters + user-space log-probs of arbitrary points: theta [n][D],
 * params [n][P], logp [n][3] */
int tnuts_constrain(tnuts_engine *e, const double *theta, int64_t n, double *params, double *logp);
int tnuts_get_stepsize(tnuts_engine *e, double *eps /* [n_chains] */);
int tnuts_get_inv_metric(tnuts_engine *e, double *minv /* [n_chains][D] */);
/* TNUTS_METRIC_DENSE: the full M^-1 of every chain, [n_chains][D][D] row-major (tnuts_get_inv_metric
 * reports its diagonal) */
int tnuts_get_dense_inv_metric(tnuts_engine *e, double *minv /* [n_chains][D][D] */);

/* LogDensityProblems interface on arbitrary points: theta [n][D] row-major, lp [n],
 * grad [n][D] (grad may be NULL) */
int tnuts_logdensity_and_gradient(tnuts_engine *e, const double *theta, int64_t n, double *lp,
                                  double *grad);
int tnuts_logdensity(tnuts_engine *e, const double *theta, int64_t n, double *lp);

/* fixed-momentum L-step leapfrog trajectories from n start points (parity check against
 * AdvancedHMC's step(Leapfrog)): theta,r [n][D]; minv [D]; outputs [n][L+1][D], H [n][L+1] */
int tnuts_leapfrog(tnuts_engine *e, const double *theta, const double *r, const double *minv,
                   double eps, int32_t L, int64_t n, double *traj_theta, double *traj_r,
                   double *traj_H);

/* sampler-state blob (HMCState analogue): size query with blob == NULL */
int tnuts_get_state(tnuts_engine *e, void *blob, size_t *nbytes);
int tnuts_set_state(tnuts_engine *e, const void *blob, size_t nbytes);

/* rank-normalised split R-hat and bulk ESS of the kept draws, computed on the device;
 * rhat, ess_bulk: [P].  With chains sharded over GPUs, call tnuts_comm_init first. */
int tnuts_diagnostics(tnuts_engine *e, double *rhat, double *ess_bulk);

/* counters */
int64_t tnuts_num_grad_evals(const tnuts_engine *e);    /* summed over local chains     */
int64_t tnuts_num_divergences(const tnuts_engine *e);   /* kept draws with numerical_error */
int64_t tnuts_num_kernel_launches(const tnuts_engine *e);
double tnuts_device_seconds(const tnuts_engine *e);     /* CUDA-event time of last tnuts_run */
double tnuts_init_device_seconds(const tnuts_engine *e);/* CUDA-event time of last tnuts_init */
int tnuts_init_tries_max(const tnuts_engine *e);        /* worst chain's init attempts  */

/* measurement aid (option "time_kernels" = 1, set before tnuts_run): CUDA-event time, on the engine's
 * stream, of the launches of the dominant kernel inside the last tnuts_run -- lock-step strategy:
 * the family's batched gradient kernel (k_logistic_tc / k_logistic_mma / ...); the other strategies
 * run ONE kernel per call, so they report the run time.  *seconds is the total over the run, estimated
 * from a systematic sample of *sampled of the *launches launches; *tile_launches is the number of
 * 128-chain x row-chunk-set tiles those launches covered (tcgen05 path).  *persist_seconds /
 * *persist_pumps: time of, and pumps executed by, the persistent tail kernel (option "persist"),
 * which fuses gradient, reduction and state machine and is not part of *seconds.  Any pointer may
 * be NULL. */
int tnuts_kernel_times(const tnuts_engine *e, double *seconds, int64_t *launches, int64_t *sampled,
                       int64_t *tile_launches, double *persist_seconds, int64_t *persist_pumps);

/* measurement aid: time `reps` gradient evaluations of all chains at their current positions
 * with CUDA events on the engine's stream (lock-step strategy) */
int tnuts_bench_gradient(tnuts_engine *e, int32_t reps, double *seconds);

/* multi-GPU: NCCL communicator for row sharding and cross-GPU diagnostics.
 * unique_id: 128-byte ncclUniqueId produced by tnuts_comm_unique_id on rank 0 and
 * broadcast by the host (torch.distributed / MPI / Distributed.jl). */
int tnuts_comm_unique_id(void *unique_id_128);
int tnuts_comm_init(tnuts_engine *e, const void *unique_id_128, int32_t world, int32_t rank);
/* sum a small host vector over all ranks of the communicator (diagnostics statistics) */
int tnuts_comm_allreduce_host(tnuts_engine *e, double *values, int64_t n);

/* pinned host staging for the caller's output arrays (optional, speeds up tnuts_fetch) */
int tnuts_host_alloc(void **p, size_t nbytes);
int tnuts_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* TNUTS_H */
