/*
 * tnuts_oracle.h -- CPU restatement of the Turing.jl HMC/NUTS hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (turing.jl_b200/,
 * include/) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker or
 * as the reported CPU baseline -- never as the thing shipped.
 *
 * PARITY STATUS: "parity unpinned" at the leapfrog-trajectory / gradient-vector
 * level.  The reference (Turing.jl v0.46.1, /root/reference) holds no golden
 * trajectory or gradient vectors and its arithmetic lives in
 * un-vendored Julia dependencies (AdvancedHMC 0.8.3, DynamicPPL 0.42, Bijectors,
 * AbstractMCMC 5.13) that cannot be executed here (no Julia toolchain).  This
 * file restates those packages' published algorithms and is pinned against
 * everything the reference's tests DO hold for this path: the gdemo log-joint
 * surface (MAP 49/54, 7/6: test/optimisation/Optimisation.jl:299-301; posterior
 * means 49/24, 7/6: test/test_utils/numerical_tests.jl:21-25), logdensity of a
 * standard normal (test/main.jl:43-45), user-space logprior/loglikelihood/
 * logjoint identities (test/test_utils/sampler.jl:14-29) and the BinomialLogit
 * arithmetic (src/stdlib/distributions.jl:70-93).
 */
#ifndef TNUTS_ORACLE_H
#define TNUTS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* model families (same ids as include/tnuts.h) */
enum {
  ORC_GDEMO = 0,          /* test/test_utils/models.jl:15-21 */
  ORC_LINREG = 1,         /* README.md:38-47 */
  ORC_EIGHT_SCHOOLS = 2,  /* BASELINE.json configs[1] */
  ORC_LOGISTIC = 3,       /* BASELINE.json configs[2], configs[4] */
  ORC_HIER_LINREG = 4,    /* BASELINE.json configs[3] */
  ORC_STD_NORMAL = 5,     /* test/main.jl:39-48, x ~ Normal() iid, D free */
  ORC_BETA_BERNOULLI = 6, /* test/mcmc/hmc.jl:24-43: p ~ Beta(a, b); obs ~ Bernoulli(p); logit link */
  ORC_DIRICHLET_CATEGORICAL = 7 /* test/mcmc/hmc.jl:45-62: ps ~ Dirichlet(K1, al1), pd ~ Dirichlet(K2, al2),
                                   obs ~ Categorical(ps); stick-breaking simplex link */
};

typedef struct {
  int family;
  int D;              /* dimension in linked (unconstrained) space */
  int64_t N;          /* observations */
  int G;              /* groups (hier) */
  const double *X;    /* logistic: row-major [N][D]; linreg/hier: x[N] */
  const double *y;    /* [N] (gdemo: 2 values) */
  const double *sigma;/* eight schools: [N] */
  const int32_t *group; /* hier: [N], 0-based */
  double hyper[8];    /* family specific prior hyper-parameters */
} orc_model;

/* number of user-space (constrained) parameters: D, except K per K-simplex (K - 1 linked dims) */
int orc_num_params(const orc_model *m);

/* (a9) lp and gradient of getlogjoint_internal in linked space
 *      src/mcmc/hmc.jl:170-182 */
double orc_logdensity_and_gradient(const orc_model *m, const double *theta, double *grad);

/* (a16) constrained parameters and user-space logprior / loglikelihood / logjoint
 *       (no Jacobian) -- src/mcmc/hmc.jl:202,247; test/test_utils/sampler.jl:24-28 */
void orc_constrain_and_logp(const orc_model *m, const double *theta, double *params,
                            double *logp3);

/* ---- counter-based RNG shared BY SPECIFICATION with the CUDA engine ---- */
enum {
  ORC_SLOT_INIT = 0,      /* sub = try*ceil(D/2) + pair */
  ORC_SLOT_MOMENTUM = 1,  /* sub = pair index */
  ORC_SLOT_DIRECTION = 2, /* sub = depth j */
  ORC_SLOT_MH = 3,        /* sub = depth j */
  ORC_SLOT_LEAF = 4,      /* sub = (1<<j) + n   (streaming candidate choice) */
  ORC_SLOT_MERGE = 5,     /* sub = node id      (AdvancedHMC-style merge choice) */
  ORC_SLOT_EPS_MOMENTUM = 6 /* find_good_stepsize momentum, sub = pair index */
};
void orc_philox4x32_10(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                       uint32_t out[4]);
/* two uniforms in (0,1) / two normals from one Philox block */
void orc_uniform2(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t slot, uint32_t sub,
                  double u[2]);
void orc_normal2(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t slot, uint32_t sub,
                 double z[2]);

typedef struct {
  double delta;       /* target acceptance, NUTS(delta): src/mcmc/hmc.jl:352-402 */
  int max_depth;      /* default 10 */
  double delta_max;   /* default 1000.0 */
  double init_eps;    /* 0 => find_good_stepsize, src/mcmc/hmc.jl:189-194 */
  int64_t n_adapts;   /* resolved by the caller: src/mcmc/hmc.jl:106-110 */
  int sampler_mode;   /* 0 = AdvancedHMC merge sampling (recursive uniform progressive),
                         1 = streaming (reservoir) form used by the CUDA engine; same law */
  uint64_t seed;
  int algorithm;      /* 0 = NUTS, 1 = HMC (fixed n_leapfrog), 2 = HMCDA (fixed lambda) */
  int n_leapfrog;     /* HMC */
  double lambda;      /* HMCDA */
  int metric;         /* 0 = DiagEuclideanMetric (NUTS default), 2 = DenseEuclideanMetric,
                         1 = UnitEuclideanMetric (HMC / HMCDA
                         default, src/mcmc/hmc.jl:76,308,323): windows and dual-averaging restarts
                         as usual, the mass matrix stays the identity */
} orc_config;

#define ORC_NSTAT 9
/* stats columns: n_steps, acceptance_rate, log_density, hamiltonian_energy,
 * hamiltonian_energy_error, max_hamiltonian_energy_error, tree_depth,
 * numerical_error, step_size  (is_accept == 1 and nom_step_size == step_size) */

typedef struct {
  /* phase point */
  double *theta, *grad, *minv; /* [D] each, owned */
  double lp;
  double eps;          /* step size the next transition will use */
  /* dual averaging (AdvancedHMC NesterovDualAveraging) */
  double da_mu, da_xbar, da_hbar, da_eps; int64_t da_m;
  /* Welford variance estimator */
  int64_t wf_n; double *wf_mean, *wf_m2; /* [D] */
  /* DenseEuclideanMetric (cfg->metric == 2): M^-1, its upper Cholesky factor, Welford M; [D][D] */
  double *minv_dense, *chol, *wf_cov;
  /* iteration counter i of src/mcmc/hmc.jl:228 */
  int64_t iter;
  int64_t n_grad_evals;
  int init_tries;
} orc_chain;

orc_chain *orc_chain_new(int D);
void orc_chain_free(orc_chain *c);

/* (a4,a5,a14,a15) init step: initial params (theta0 or U(-2,2) with <=1000 tries),
 * phase point, find_good_stepsize, adaptor initialisation.
 * returns 0, or -1 when no finite initial point was found. */
int orc_chain_init(const orc_model *m, const orc_config *cfg, uint32_t chain_id,
                   const double *theta0, orc_chain *c);

/* (a7,a8,a10) one transition + adaptation; fills stats[ORC_NSTAT] */
void orc_chain_step(const orc_model *m, const orc_config *cfg, uint32_t chain_id, orc_chain *c,
                    double *stats);

/* drive n_transitions; keep transitions first_keep, first_keep+thin, ...
 * outputs may be NULL.  Layouts: [keep][D], [keep][P], [keep][3], [keep][ORC_NSTAT] */
int64_t orc_chain_run(const orc_model *m, const orc_config *cfg, uint32_t chain_id, orc_chain *c,
                      int64_t n_transitions, int64_t first_keep, int64_t thin,
                      double *out_theta, double *out_params, double *out_logp, double *out_stats);

/* whole job, one chain per OpenMP thread (the MCMCThreads fan-out,
 * src/mcmc/abstractmcmc.jl:133-166).  Outputs [chain][keep][...]. theta0 may be NULL
 * or [chains][D].  Returns total gradient evaluations, or -1 on init failure. */
int64_t orc_sample_chains(const orc_model *m, const orc_config *cfg, int n_chains, int n_threads,
                          const double *theta0, int64_t n_transitions, int64_t first_keep,
                          int64_t thin, double *out_theta, double *out_params, double *out_logp,
                          double *out_stats, double *out_final_eps, double *out_final_minv);

/* the same for an arbitrary set of global chain ids (the ids address the random streams): lets a
 * test compare a few chains of a 65 536-chain engine run without running the others */
int64_t orc_sample_chain_ids(const orc_model *m, const orc_config *cfg, int n_chains, int n_threads,
                             const uint32_t *chain_ids, const double *theta0, int64_t n_transitions,
                             int64_t first_keep, int64_t thin, double *out_theta, double *out_params,
                             double *out_logp, double *out_stats, double *out_final_eps,
                             double *out_final_minv);

/* continue adapted chains from (theta, eps, diagonal M^-1, iteration counter): the resume path of
 * src/mcmc/hmc.jl:135-152 (nadapts = 0 when cfg->n_adapts <= iter0).  Layouts as above. */
int64_t orc_resume_chains(const orc_model *m, const orc_config *cfg, int n_chains, int n_threads,
                          uint32_t chain_id0, const double *theta, const double *eps,
                          const double *minv, int64_t iter0, int64_t n_transitions,
                          int64_t first_keep, int64_t thin, double *out_theta, double *out_params,
                          double *out_logp, double *out_stats, double *out_final_theta);

/* building blocks exposed for parity tests */
/* fixed-momentum L-step leapfrog trajectory; traj_theta/traj_r [L+1][D] (incl. start),
 * traj_H [L+1].  AdvancedHMC step(Leapfrog) restated. */
void orc_leapfrog_trajectory(const orc_model *m, const double *theta, const double *r,
                             const double *minv, double eps, int L, double *traj_theta,
                             double *traj_r, double *traj_H);
/* Stan window schedule; returns number of splits written (1-based ends) */
int orc_window_splits(int64_t n_adapts, int64_t *window_start, int64_t *window_end,
                      int64_t *splits, int max_splits);
double orc_find_good_stepsize(const orc_model *m, const orc_config *cfg, uint32_t chain_id,
                              const double *theta, const double *minv, int64_t *n_grad);

#ifdef __cplusplus
}
#endif
#endif
