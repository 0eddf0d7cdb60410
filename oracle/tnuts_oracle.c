/*
 * tnuts_oracle.c -- CPU restatement (plain C99, fp64) of the Turing.jl HMC/NUTS hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see tnuts_oracle.h).  "parity unpinned" at the
 * trajectory/gradient-vector level: the reference holds no such golden vectors and its
 * arithmetic is in un-vendored Julia packages; see the header for what pins this file.
 *
 * What is restated, and from where (paths are into /root/reference unless marked [dep]):
 *   - schedule / wiring ......... src/mcmc/hmc.jl:85-154 (nadapts, discard), :156-208 (init
 *                                 step), :210-252 (iterate step), :425-469 (kernel, adaptor)
 *   - initial parameters ........ src/mcmc/abstractmcmc.jl:42-70 (<=1000 tries, finite lp+grad),
 *                                 src/mcmc/hmc.jl:82 (InitFromUniform: U(-2,2) in linked space)
 *   - leapfrog, NUTS transition . [dep AdvancedHMC 0.8.3] Leapfrog step, Trajectory{MultinomialTS}
 *                                 with GeneralisedNoUTurn; call site src/mcmc/hmc.jl:225
 *   - adaptation ................ [dep AdvancedHMC 0.8.3] StanHMCAdaptor(MassMatrixAdaptor,
 *                                 StepSizeAdaptor); call site src/mcmc/hmc.jl:229-241
 *   - find_good_stepsize ........ [dep AdvancedHMC 0.8.3]; call site src/mcmc/hmc.jl:190
 *   - log densities ............. [dep DynamicPPL 0.42 / Bijectors / Distributions]
 *                                 getlogjoint_internal in linked space, src/mcmc/hmc.jl:170-176;
 *                                 logit arithmetic as src/stdlib/distributions.jl:70-93
 * Random numbers: the reference uses Julia's task-local Xoshiro; streams cannot match, so
 * this oracle and the CUDA engine share a Philox4x32-10 addressing SPEC instead (header).
 */
#include "tnuts_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define LOG_2PI 1.8378770664093454835606594728112
#define LOG_2 0.69314718055994530941723212145818
#define LOG_PI 1.1447298858494001741434273513531

/* ------------------------------------------------------------------------------------------
 * Log densities.  theta is in linked space; positive variables link by log.
 * ------------------------------------------------------------------------------------------ */
int orc_num_params(const orc_model *m) {
  /* a K-simplex has K - 1 linked dimensions (stick breaking) and K user-space parameters */
  if (m->family == ORC_DIRICHLET_CATEGORICAL) return (int)m->hyper[0] + (int)m->hyper[2];
  return m->D;
}

static double log1pexp(double x) {
  /* StatsFuns.log1pexp, used by BinomialLogit: src/stdlib/distributions.jl:77 */
  if (x < -37.0) return exp(x);
  if (x < 18.0) return log1p(exp(x));
  if (x < 33.3) return x + exp(-x);
  return x;
}

/* gdemo: s ~ InverseGamma(2,3); m ~ Normal(0,sqrt(s)); x_i ~ Normal(m,sqrt(s))
 * test/test_utils/models.jl:15-21.  theta = (log s, m). hyper = {alpha, beta}. */
static double lp_gdemo(const orc_model *md, const double *th, double *g) {
  const double al = md->hyper[0], be = md->hyper[1];
  const double a = th[0], m = th[1], s = exp(a);
  double Q = m * m;
  double sres = 0.0;
  for (int64_t i = 0; i < md->N; ++i) {
    double d = md->y[i] - m;
    Q += d * d;
    sres += d;
  }
  const double n1 = (double)(md->N + 1);
  double lp = al * log(be) - lgamma(al) - (al + 1.0) * a - be / s /* InverseGamma */
              - 0.5 * n1 * LOG_2PI - 0.5 * n1 * a - Q / (2.0 * s)     /* normals */
              + a;                                                   /* log|ds/da| */
  if (g) {
    g[0] = -(al + 1.0) + be / s - 0.5 * n1 + Q / (2.0 * s) + 1.0;
    g[1] = (-m + sres) / s;
  }
  return lp;
}

/* README.md:38-47: alpha,beta ~ Normal(0,1); sigma2 ~ truncated(Cauchy(0,3); lower=0);
 * y ~ MvNormal(alpha .+ beta .* x, sigma2*I).  theta = (alpha, beta, log sigma2).
 * hyper = {prior sd of alpha/beta, cauchy scale}. */
static double lp_linreg(const orc_model *md, const double *th, double *g) {
  const double sd = md->hyper[0], c = md->hyper[1];
  const double al = th[0], be = th[1], t = th[2], v = exp(t);
  double ssr = 0, sr = 0, srx = 0;
  for (int64_t i = 0; i < md->N; ++i) {
    double res = md->y[i] - al - be * md->X[i];
    ssr += res * res;
    sr += res;
    srx += res * md->X[i];
  }
  const double n = (double)md->N, w = (v / c) * (v / c);
  double lp = -LOG_2PI - 2.0 * log(sd) - 0.5 * (al * al + be * be) / (sd * sd)
              + LOG_2 - LOG_PI - log(c) - log1p(w) + t
              - 0.5 * n * (LOG_2PI + t) - ssr / (2.0 * v);
  if (g) {
    g[0] = -al / (sd * sd) + sr / v;
    g[1] = -be / (sd * sd) + srx / v;
    g[2] = -2.0 * w / (1.0 + w) + 1.0 - 0.5 * n + ssr / (2.0 * v);
  }
  return lp;
}

/* eight schools, non-centred: mu ~ Normal(0,5); tau ~ truncated(Cauchy(0,5); lower=0);
 * thetat ~ MvNormal(0, I); y_j ~ Normal(mu + tau*thetat_j, sigma_j).
 * theta = (mu, log tau, thetat_1..J).  hyper = {mu sd, cauchy scale}. */
static double lp_eight_schools(const orc_model *md, const double *th, double *g) {
  const double sd = md->hyper[0], c = md->hyper[1];
  const int J = (int)md->N;
  const double mu = th[0], lt = th[1], tau = exp(lt), w = (tau / c) * (tau / c);
  double lp = -0.5 * LOG_2PI - log(sd) - 0.5 * mu * mu / (sd * sd)
              + LOG_2 - LOG_PI - log(c) - log1p(w) + lt;
  double sz = 0, szt = 0;
  for (int j = 0; j < J; ++j) {
    const double tj = th[2 + j], sg = md->sigma[j];
    const double res = md->y[j] - mu - tau * tj;
    const double z = res / (sg * sg);
    lp += -0.5 * LOG_2PI - 0.5 * tj * tj - 0.5 * LOG_2PI - log(sg) - 0.5 * res * z;
    sz += z;
    szt += z * tj;
    if (g) g[2 + j] = -tj + tau * z;
  }
  if (g) {
    g[0] = -mu / (sd * sd) + sz;
    g[1] = -2.0 * w / (1.0 + w) + 1.0 + tau * szt;
  }
  return lp;
}

/* Bayesian logistic regression: beta ~ MvNormal(0, sd^2 I); y_i ~ BernoulliLogit(x_i . beta).
 * hyper = {sd}.  X row-major [N][D].
 * One pass over the rows per gradient (eta_i = x_i . beta, then g += r_i x_i), which is what one
 * task of the reference does per leapfrog.  The dot product runs in ORC_LANES interleaved partial
 * sums with a fixed combination order, so the value does not depend on the vector ISA the
 * compiler picks (the clones below only differ in how many lanes one instruction carries);
 * log1pexp and logistic share one exp:  e = exp(-|eta|),  log1pexp = max(eta, 0) + log1p(e),
 * logistic = 1 / (1 + e) or e / (1 + e)  (StatsFuns' branches agree with this to 1 ulp). */
#define ORC_LANES 8
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("default", "avx", "avx2", "avx512f")))
#endif
static double lp_logistic_rows(const double *X, const double *y, int64_t N, int D, const double *th,
                               double *g) {
  double lp = 0.0;
  const int Dv = D & ~(ORC_LANES - 1);
  for (int64_t i = 0; i < N; ++i) {
    const double *x = X + i * (int64_t)D;
    double acc[ORC_LANES];
    for (int l = 0; l < ORC_LANES; ++l) acc[l] = 0.0;
    for (int d = 0; d < Dv; d += ORC_LANES) {
#pragma omp simd
      for (int l = 0; l < ORC_LANES; ++l) acc[l] += x[d + l] * th[d + l];
    }
    double eta = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    for (int d = Dv; d < D; ++d) eta += x[d] * th[d];
    const double e = exp(-fabs(eta));
    lp += y[i] * eta - ((eta > 0.0 ? eta : 0.0) + log1p(e));
    if (g) {
      const double r = y[i] - (eta >= 0.0 ? 1.0 / (1.0 + e) : e / (1.0 + e));
#pragma omp simd
      for (int d = 0; d < D; ++d) g[d] += r * x[d];
    }
  }
  return lp;
}

static double lp_logistic(const orc_model *md, const double *th, double *g) {
  const double sd = md->hyper[0];
  const int D = md->D;
  double lp = -0.5 * D * LOG_2PI - D * log(sd);
  for (int d = 0; d < D; ++d) {
    lp -= 0.5 * th[d] * th[d] / (sd * sd);
    if (g) g[d] = -th[d] / (sd * sd);
  }
  return lp + lp_logistic_rows(md->X, md->y, md->N, D, th, g);
}

/* hierarchical linear regression (centred): mu_a, mu_b ~ Normal(0, sdm);
 * sigma_a, sigma_b, sigma_y ~ truncated(Normal(0, s0); lower=0);
 * a_g ~ Normal(mu_a, sigma_a); b_g ~ Normal(mu_b, sigma_b);
 * y_i ~ Normal(a_{g(i)} + b_{g(i)} x_i, sigma_y).
 * theta = (mu_a, mu_b, log sigma_a, log sigma_b, log sigma_y, a_1..G, b_1..G).
 * hyper = {sdm, s0}. */
static double lp_hier(const orc_model *md, const double *th, double *g) {
  const double sdm = md->hyper[0], s0 = md->hyper[1];
  const int G = md->G;
  const double mua = th[0], mub = th[1], la = th[2], lb = th[3], ly = th[4];
  const double sa = exp(la), sb = exp(lb), sy = exp(ly);
  const double *a = th + 5, *b = th + 5 + G;
  double lp = 2.0 * (-0.5 * LOG_2PI - log(sdm)) - 0.5 * (mua * mua + mub * mub) / (sdm * sdm);
  /* half-normal priors + Jacobians */
  lp += 3.0 * (LOG_2 - 0.5 * LOG_2PI - log(s0)) - 0.5 * (sa * sa + sb * sb + sy * sy) / (s0 * s0)
        + la + lb + ly;
  double qa = 0, qb = 0, da = 0, db = 0;
  if (g) memset(g, 0, sizeof(double) * (size_t)md->D);
  for (int k = 0; k < G; ++k) {
    const double ea = a[k] - mua, eb = b[k] - mub;
    qa += ea * ea;
    qb += eb * eb;
    da += ea;
    db += eb;
    if (g) {
      g[5 + k] = -ea / (sa * sa);
      g[5 + G + k] = -eb / (sb * sb);
    }
  }
  lp += -(double)G * LOG_2PI - G * la - G * lb - 0.5 * qa / (sa * sa) - 0.5 * qb / (sb * sb);
  double ssr = 0;
  const double isy2 = 1.0 / (sy * sy);
  for (int64_t i = 0; i < md->N; ++i) {
    const int k = md->group[i];
    const double res = md->y[i] - a[k] - b[k] * md->X[i];
    ssr += res * res;
    if (g) {
      g[5 + k] += res * isy2;
      g[5 + G + k] += res * md->X[i] * isy2;
    }
  }
  const double n = (double)md->N;
  lp += -0.5 * n * LOG_2PI - n * ly - 0.5 * ssr * isy2;
  if (g) {
    g[0] = -mua / (sdm * sdm) + da / (sa * sa);
    g[1] = -mub / (sdm * sdm) + db / (sb * sb);
    g[2] = -sa * sa / (s0 * s0) + 1.0 - G + qa / (sa * sa);
    g[3] = -sb * sb / (s0 * s0) + 1.0 - G + qb / (sb * sb);
    g[4] = -sy * sy / (s0 * s0) + 1.0 - n + ssr * isy2;
  }
  return lp;
}

/* ---- the two constrained-support models of the reference's HMC tests --------------------------
 * test/mcmc/hmc.jl:24-43  "constrained bounded":  p ~ Beta(a, b); obs[i] ~ Bernoulli(p)
 *   theta = logit(p) (Bijectors' Logit link on (0, 1)); log|det J_invlink| = log p + log(1 - p).
 *   lp = (a + s) log p + (b + f) log(1 - p) - log B(a, b),  s / f = number of ones / zeros.
 *   hyper = {a, b}; y[N] = observations (0 / 1).
 * test/mcmc/hmc.jl:45-62  "constrained simplex":  ps ~ Dirichlet(K1, al1); pd ~ Dirichlet(K2, al2);
 *   obs[i] ~ Categorical(ps)
 *   Each simplex x of size K links by stick breaking (Bijectors' SimplexBijector, the Stan
 *   transform):  z_k = logistic(y_k - log(K - k)), x_k = z_k rem_k, rem_1 = 1,
 *   rem_{k+1} = rem_k (1 - z_k), x_K = rem_K;  log|det J_invlink| = sum_k log z_k + log(1 - z_k) +
 *   log rem_k.  hyper = {K1, al1, K2, al2}; y[N] = observed categories 1..K1.  D = K1 - 1 + K2 - 1.
 * Both are sums of a log sigma(u) + b log sigma(-u) terms, u = theta_d - shift_d. */
static double log_sigmoid(double u) { /* log logistic(u), stable in both tails */
  return u >= 0 ? -log1p(exp(-u)) : u - log1p(exp(u));
}
static double sigmoid(double u) {
  if (u >= 0) return 1.0 / (1.0 + exp(-u));
  const double e = exp(u);
  return e / (1.0 + e);
}

static double lp_beta_bernoulli(const orc_model *md, const double *th, double *g) {
  const double a = md->hyper[0], b = md->hyper[1];
  double s = 0, f = 0;
  for (int64_t i = 0; i < md->N; ++i) {
    if (md->y[i] != 0.0) s += 1.0;
    else f += 1.0;
  }
  const double u = th[0], p = sigmoid(u);
  if (g) g[0] = (a + s) * (1.0 - p) - (b + f) * p;
  return (a + s) * log_sigmoid(u) + (b + f) * log_sigmoid(-u) - (lgamma(a) + lgamma(b) - lgamma(a + b));
}

/* one Dirichlet(K, al) block with category counts cnt[K] (NULL: no observations): adds its lp,
 * writes K - 1 gradient entries */
static double lp_simplex_block(int K, double al, const double *cnt, const double *y, double *g) {
  double lp = lgamma(K * al) - K * lgamma(al);
  double tail = (al - 1.0) + (cnt ? cnt[K - 1] : 0.0); /* sum of c_k over k > j, c_k = al - 1 + n_k */
  for (int j = K - 2; j >= 0; --j) {                  /* stick j (0-based): K - 1 - j categories after it */
    const double cj = (al - 1.0) + (cnt ? cnt[j] : 0.0);
    const double aj = cj + 1.0, bj = tail + 1.0 + (double)(K - 2 - j);
    const double u = y[j] - log((double)(K - 1 - j));
    const double z = sigmoid(u);
    lp += aj * log_sigmoid(u) + bj * log_sigmoid(-u);
    if (g) g[j] = aj * (1.0 - z) - bj * z;
    tail += cj;
  }
  return lp;
}

static double lp_dirichlet_categorical(const orc_model *md, const double *th, double *g) {
  const int K1 = (int)md->hyper[0], K2 = (int)md->hyper[2];
  const double al1 = md->hyper[1], al2 = md->hyper[3];
  double cnt[64];
  for (int k = 0; k < K1; ++k) cnt[k] = 0.0;
  for (int64_t i = 0; i < md->N; ++i) cnt[(int)md->y[i] - 1] += 1.0;
  double lp = lp_simplex_block(K1, al1, cnt, th, g);
  if (K2 > 0) lp += lp_simplex_block(K2, al2, NULL, th + (K1 - 1), g ? g + (K1 - 1) : NULL);
  return lp;
}

static void simplex_invlink(int K, const double *y, double *x) {
  double rem = 1.0;
  for (int j = 0; j < K - 1; ++j) {
    const double z = sigmoid(y[j] - log((double)(K - 1 - j)));
    x[j] = z * rem;
    rem *= 1.0 - z;
  }
  x[K - 1] = rem;
}

static double lp_std_normal(const orc_model *md, const double *th, double *g) {
  double lp = -0.5 * md->D * LOG_2PI;
  for (int d = 0; d < md->D; ++d) {
    lp -= 0.5 * th[d] * th[d];
    if (g) g[d] = -th[d];
  }
  return lp;
}

double orc_logdensity_and_gradient(const orc_model *m, const double *theta, double *grad) {
  switch (m->family) {
    case ORC_GDEMO: return lp_gdemo(m, theta, grad);
    case ORC_LINREG: return lp_linreg(m, theta, grad);
    case ORC_EIGHT_SCHOOLS: return lp_eight_schools(m, theta, grad);
    case ORC_LOGISTIC: return lp_logistic(m, theta, grad);
    case ORC_HIER_LINREG: return lp_hier(m, theta, grad);
    case ORC_STD_NORMAL: return lp_std_normal(m, theta, grad);
    case ORC_BETA_BERNOULLI: return lp_beta_bernoulli(m, theta, grad);
    case ORC_DIRICHLET_CATEGORICAL: return lp_dirichlet_categorical(m, theta, grad);
  }
  return NAN;
}

/* user-space output of DynamicPPL.ParamsWithStats (src/mcmc/hmc.jl:202,247): constrained
 * parameters in model order and logprior / loglikelihood / logjoint WITHOUT the Jacobian
 * (test/test_utils/sampler.jl:24-28). */
void orc_constrain_and_logp(const orc_model *md, const double *th, double *p, double *lp3) {
  double prior = 0, lik = 0;
  switch (md->family) {
    case ORC_GDEMO: {
      const double al = md->hyper[0], be = md->hyper[1];
      const double a = th[0], m = th[1], s = exp(a);
      p[0] = s;
      p[1] = m;
      prior = al * log(be) - lgamma(al) - (al + 1.0) * a - be / s - 0.5 * LOG_2PI - 0.5 * a -
              m * m / (2.0 * s);
      for (int64_t i = 0; i < md->N; ++i) {
        double d = md->y[i] - m;
        lik += -0.5 * LOG_2PI - 0.5 * a - d * d / (2.0 * s);
      }
    } break;
    case ORC_LINREG: {
      const double sd = md->hyper[0], c = md->hyper[1];
      const double v = exp(th[2]), w = (v / c) * (v / c);
      p[0] = th[0];
      p[1] = th[1];
      p[2] = v;
      prior = -LOG_2PI - 2.0 * log(sd) - 0.5 * (th[0] * th[0] + th[1] * th[1]) / (sd * sd) + LOG_2 -
              LOG_PI - log(c) - log1p(w);
      double ssr = 0;
      for (int64_t i = 0; i < md->N; ++i) {
        double res = md->y[i] - th[0] - th[1] * md->X[i];
        ssr += res * res;
      }
      lik = -0.5 * md->N * (LOG_2PI + th[2]) - ssr / (2.0 * v);
    } break;
    case ORC_EIGHT_SCHOOLS: {
      const double sd = md->hyper[0], c = md->hyper[1];
      const double mu = th[0], tau = exp(th[1]), w = (tau / c) * (tau / c);
      p[0] = mu;
      p[1] = tau;
      prior = -0.5 * LOG_2PI - log(sd) - 0.5 * mu * mu / (sd * sd) + LOG_2 - LOG_PI - log(c) -
              log1p(w);
      for (int j = 0; j < (int)md->N; ++j) {
        const double tj = th[2 + j], sg = md->sigma[j];
        const double res = md->y[j] - mu - tau * tj;
        p[2 + j] = tj;
        prior += -0.5 * LOG_2PI - 0.5 * tj * tj;
        lik += -0.5 * LOG_2PI - log(sg) - 0.5 * res * res / (sg * sg);
      }
    } break;
    case ORC_LOGISTIC: {
      const double sd = md->hyper[0];
      prior = -0.5 * md->D * LOG_2PI - md->D * log(sd);
      for (int d = 0; d < md->D; ++d) {
        p[d] = th[d];
        prior -= 0.5 * th[d] * th[d] / (sd * sd);
      }
      for (int64_t i = 0; i < md->N; ++i) {
        const double *x = md->X + i * (int64_t)md->D;
        double eta = 0;
        for (int d = 0; d < md->D; ++d) eta += x[d] * th[d];
        lik += md->y[i] * eta - log1pexp(eta);
      }
    } break;
    case ORC_HIER_LINREG: {
      const double sdm = md->hyper[0], s0 = md->hyper[1];
      const int G = md->G;
      const double sa = exp(th[2]), sb = exp(th[3]), sy = exp(th[4]);
      p[0] = th[0];
      p[1] = th[1];
      p[2] = sa;
      p[3] = sb;
      p[4] = sy;
      prior = 2.0 * (-0.5 * LOG_2PI - log(sdm)) - 0.5 * (th[0] * th[0] + th[1] * th[1]) / (sdm * sdm) +
              3.0 * (LOG_2 - 0.5 * LOG_2PI - log(s0)) - 0.5 * (sa * sa + sb * sb + sy * sy) / (s0 * s0);
      double qa = 0, qb = 0;
      for (int k = 0; k < G; ++k) {
        p[5 + k] = th[5 + k];
        p[5 + G + k] = th[5 + G + k];
        double ea = th[5 + k] - th[0], eb = th[5 + G + k] - th[1];
        qa += ea * ea;
        qb += eb * eb;
      }
      prior += -(double)G * LOG_2PI - G * th[2] - G * th[3] - 0.5 * qa / (sa * sa) - 0.5 * qb / (sb * sb);
      double ssr = 0;
      for (int64_t i = 0; i < md->N; ++i) {
        const int k = md->group[i];
        double res = md->y[i] - th[5 + k] - th[5 + G + k] * md->X[i];
        ssr += res * res;
      }
      lik = -0.5 * md->N * LOG_2PI - md->N * th[4] - 0.5 * ssr / (sy * sy);
    } break;
    case ORC_STD_NORMAL: {
      prior = -0.5 * md->D * LOG_2PI;
      for (int d = 0; d < md->D; ++d) {
        p[d] = th[d];
        prior -= 0.5 * th[d] * th[d];
      }
    } break;
    case ORC_BETA_BERNOULLI: {
      const double a = md->hyper[0], b = md->hyper[1];
      const double lpp = log_sigmoid(th[0]), lq = log_sigmoid(-th[0]);
      p[0] = sigmoid(th[0]);
      prior = (a - 1.0) * lpp + (b - 1.0) * lq - (lgamma(a) + lgamma(b) - lgamma(a + b));
      for (int64_t i = 0; i < md->N; ++i) lik += md->y[i] != 0.0 ? lpp : lq;
    } break;
    case ORC_DIRICHLET_CATEGORICAL: {
      const int K1 = (int)md->hyper[0], K2 = (int)md->hyper[2];
      const double al1 = md->hyper[1], al2 = md->hyper[3];
      simplex_invlink(K1, th, p);
      prior = lgamma(K1 * al1) - K1 * lgamma(al1);
      for (int k = 0; k < K1; ++k) prior += (al1 - 1.0) * log(p[k]);
      if (K2 > 0) {
        simplex_invlink(K2, th + (K1 - 1), p + K1);
        prior += lgamma(K2 * al2) - K2 * lgamma(al2);
        for (int k = 0; k < K2; ++k) prior += (al2 - 1.0) * log(p[K1 + k]);
      }
      for (int64_t i = 0; i < md->N; ++i) lik += log(p[(int)md->y[i] - 1]);
    } break;
  }
  lp3[0] = prior;
  lp3[1] = lik;
  lp3[2] = prior + lik;
}

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al. 2011).  Addressing spec: key = seed, ctr = (chain, iter,
 * slot, sub); see tnuts_oracle.h.
 * ------------------------------------------------------------------------------------------ */
void orc_philox4x32_10(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                       uint32_t out[4]) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static double u53(uint32_t hi, uint32_t lo) {
  uint64_t b = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)b + 0.5) * (1.0 / 9007199254740992.0);
}
void orc_uniform2(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t slot, uint32_t sub,
                  double u[2]) {
  uint32_t o[4];
  orc_philox4x32_10(seed, chain, iter, slot, sub, o);
  u[0] = u53(o[0], o[1]);
  u[1] = u53(o[2], o[3]);
}
void orc_normal2(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t slot, uint32_t sub,
                 double z[2]) {
  double u[2];
  orc_uniform2(seed, chain, iter, slot, sub, u);
  const double rad = sqrt(-2.0 * log(u[0]));
  const double ang = 6.283185307179586476925286766559 * u[1];
  z[0] = rad * cos(ang);
  z[1] = rad * sin(ang);
}

/* ------------------------------------------------------------------------------------------
 * Hamiltonian pieces ([dep AdvancedHMC] Hamiltonian/PhasePoint)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  double *theta, *r, *grad; /* grad = d log pi / d theta */
  double lp, H;
} phase;

static void phase_alloc(phase *z, int D) {
  z->theta = (double *)malloc(sizeof(double) * 3 * (size_t)D);
  z->r = z->theta + D;
  z->grad = z->r + D;
  z->lp = 0;
  z->H = 0;
}
static void phase_free(phase *z) { free(z->theta); }
static void phase_copy(phase *dst, const phase *src, int D) {
  memcpy(dst->theta, src->theta, sizeof(double) * 3 * (size_t)D);
  dst->lp = src->lp;
  dst->H = src->H;
}

/* DenseEuclideanMetric(M^-1) ([dep AdvancedHMC], restated from memory like App. A): the metric
 * stores M^-1 and the upper Cholesky factor U of M^-1 (M^-1 = U'U); momentum r = U \ randn(D)
 * (so cov(r) = M), kinetic energy r' M^-1 r / 2, dH/dr = M^-1 r, the U-turn products use M^-1 r.
 * The active chain's dense metric sits in a thread-local context (one chain per thread); when it is
 * set, the `minv` argument of the functions below is ignored.  Row-major [D][D]. */
static __thread const double *tl_dense_minv = NULL, *tl_dense_chol = NULL;

static void dense_apply(int D, const double *A, const double *r, double *out) {
  for (int i = 0; i < D; ++i) {
    double s = 0;
    for (int j = 0; j < D; ++j) s += A[(size_t)i * D + j] * r[j];
    out[i] = s;
  }
}
/* upper Cholesky factor U of the symmetric positive definite A (A = U'U); 0 on success */
static int dense_chol_upper(int D, const double *A, double *U) {
  memset(U, 0, sizeof(double) * (size_t)D * D);
  for (int j = 0; j < D; ++j) {
    double s = A[(size_t)j * D + j];
    for (int k = 0; k < j; ++k) s -= U[(size_t)k * D + j] * U[(size_t)k * D + j];
    if (!(s > 0)) return -1;
    const double ujj = sqrt(s);
    U[(size_t)j * D + j] = ujj;
    for (int i = j + 1; i < D; ++i) {
      double t = A[(size_t)j * D + i];
      for (int k = 0; k < j; ++k) t -= U[(size_t)k * D + j] * U[(size_t)k * D + i];
      U[(size_t)j * D + i] = t / ujj;
    }
  }
  return 0;
}

/* energy with the non-finite rule of AdvancedHMC's PhasePoint constructor: any non-finite
 * lp, gradient or kinetic energy makes H = +Inf (App. A.1) */
static double energy(int D, double lp, const double *grad, const double *r, const double *minv) {
  double k = 0;
  int ok = isfinite(lp);
  if (tl_dense_minv) {
    for (int i = 0; i < D; ++i) {
      double s = 0;
      for (int j = 0; j < D; ++j) s += tl_dense_minv[(size_t)i * D + j] * r[j];
      k += r[i] * s;
      ok = ok && isfinite(grad[i]);
    }
  } else {
    for (int d = 0; d < D; ++d) {
      k += minv[d] * r[d] * r[d];
      ok = ok && isfinite(grad[d]);
    }
  }
  k *= 0.5;
  if (!ok || !isfinite(k)) return INFINITY;
  return -lp + k;
}

/* one leapfrog step of signed size eps (App. A.2): half kick, drift, gradient, half kick */
static void leapfrog(const orc_model *m, phase *z, const double *minv, double eps,
                     int64_t *n_grad) {
  const int D = m->D;
  const double half = 0.5 * eps;
  for (int d = 0; d < D; ++d) z->r[d] = z->r[d] + half * z->grad[d];
  if (tl_dense_minv) {
    double *v = (double *)malloc(sizeof(double) * (size_t)D);
    dense_apply(D, tl_dense_minv, z->r, v);
    for (int d = 0; d < D; ++d) z->theta[d] = z->theta[d] + eps * v[d];
    free(v);
  } else {
    for (int d = 0; d < D; ++d) z->theta[d] = z->theta[d] + eps * (minv[d] * z->r[d]);
  }
  z->lp = orc_logdensity_and_gradient(m, z->theta, z->grad);
  if (n_grad) ++*n_grad;
  for (int d = 0; d < D; ++d) z->r[d] = z->r[d] + half * z->grad[d];
  z->H = energy(D, z->lp, z->grad, z->r, minv);
}

void orc_leapfrog_trajectory(const orc_model *m, const double *theta, const double *r,
                             const double *minv, double eps, int L, double *traj_theta,
                             double *traj_r, double *traj_H) {
  const int D = m->D;
  phase z;
  phase_alloc(&z, D);
  memcpy(z.theta, theta, sizeof(double) * (size_t)D);
  memcpy(z.r, r, sizeof(double) * (size_t)D);
  z.lp = orc_logdensity_and_gradient(m, z.theta, z.grad);
  z.H = energy(D, z.lp, z.grad, z.r, minv);
  for (int l = 0; l <= L; ++l) {
    memcpy(traj_theta + (size_t)l * D, z.theta, sizeof(double) * (size_t)D);
    memcpy(traj_r + (size_t)l * D, z.r, sizeof(double) * (size_t)D);
    traj_H[l] = z.H;
    if (l < L) leapfrog(m, &z, minv, eps, NULL);
  }
  phase_free(&z);
}

static double logaddexp(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  double mx = a > b ? a : b;
  return mx + log1p(exp(-fabs(a - b)));
}

static void draw_momentum(const orc_config *cfg, uint32_t chain, uint32_t iter, uint32_t slot,
                          int D, const double *minv, double *r) {
  /* r = randn(D) ./ sqrt.(M^-1)  (App. A.1) */
  if (tl_dense_chol) { /* dense: solve U r = z by back substitution */
    for (int p = 0; 2 * p < D; ++p) {
      double z[2];
      orc_normal2(cfg->seed, chain, iter, slot, (uint32_t)p, z);
      r[2 * p] = z[0];
      if (2 * p + 1 < D) r[2 * p + 1] = z[1];
    }
    for (int d = D - 1; d >= 0; --d) {
      double t = r[d];
      for (int e = d + 1; e < D; ++e) t -= tl_dense_chol[(size_t)d * D + e] * r[e];
      r[d] = t / tl_dense_chol[(size_t)d * D + d];
    }
    return;
  }
  for (int p = 0; 2 * p < D; ++p) {
    double z[2];
    orc_normal2(cfg->seed, chain, iter, slot, (uint32_t)p, z);
    r[2 * p] = z[0] / sqrt(minv[2 * p]);
    if (2 * p + 1 < D) r[2 * p + 1] = z[1] / sqrt(minv[2 * p + 1]);
  }
}


/* ------------------------------------------------------------------------------------------
 * NUTS transition: Trajectory{MultinomialTS}(Leapfrog, GeneralisedNoUTurn)  (App. A.3)
 * Recursive build_tree exactly as AdvancedHMC states it.  Candidate selection inside a
 * doubling has two equivalent-in-law forms, chosen by cfg->sampler_mode:
 *   0: AdvancedHMC's own: at every merge of two sub-trees keep the first candidate w.p.
 *      exp(lw1 - logaddexp(lw1,lw2))  (one uniform per merge, ORC_SLOT_MERGE);
 *   1: streaming: leaf n replaces the running candidate w.p. w_n / sum_{k<=n} w_k
 *      (one uniform per leaf, ORC_SLOT_LEAF).  Both give P(leaf n) = w_n / sum_k w_k and
 *      are independent of the trajectory, hence of termination.  The CUDA engine uses 1.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  phase zl, zr; /* leftmost / rightmost phase points */
  phase zc;     /* candidate (mode 0) */
  double *rho;  /* sum of momenta */
  double sum_a; /* sum of min(1,exp(-dH)) */
  int64_t n_a;
  double dH_max;
  double lw; /* log weight */
  int t_dyn, t_num;
} subtree;

static void subtree_alloc(subtree *t, int D) {
  phase_alloc(&t->zl, D);
  phase_alloc(&t->zr, D);
  phase_alloc(&t->zc, D);
  t->rho = (double *)malloc(sizeof(double) * (size_t)D);
}
static void subtree_free(subtree *t) {
  phase_free(&t->zl);
  phase_free(&t->zr);
  phase_free(&t->zc);
  free(t->rho);
}

#define ORC_MAX_DEPTH 24
typedef struct {
  const orc_model *m;
  const orc_config *cfg;
  const double *minv;
  uint32_t chain, iter;
  double eps, H0;
  int jtop; /* depth of the doubling being built */
  int64_t *n_grad;
  subtree pool[ORC_MAX_DEPTH]; /* pool[j] = scratch second half for a depth-(j+1) call */
  int pool_n;
  /* streaming-candidate state (sampler_mode 1) */
  double res_lw;
  phase res_c;
} tree_ctx;

static int uturn(int D, const double *rho, const double *rl, const double *rr, const double *minv) {
  /* GeneralisedNoUTurn: dot(rho, M^-1 r_left) <= 0 || dot(rho, M^-1 r_right) <= 0 */
  double a = 0, b = 0;
  if (tl_dense_minv) {
    for (int i = 0; i < D; ++i) {
      double sl = 0, sr = 0;
      for (int j = 0; j < D; ++j) {
        sl += tl_dense_minv[(size_t)i * D + j] * rl[j];
        sr += tl_dense_minv[(size_t)i * D + j] * rr[j];
      }
      a += rho[i] * sl;
      b += rho[i] * sr;
    }
    return (a <= 0) || (b <= 0);
  }
  for (int d = 0; d < D; ++d) {
    a += rho[d] * (minv[d] * rl[d]);
    b += rho[d] * (minv[d] * rr[d]);
  }
  return (a <= 0) || (b <= 0);
}

/* build_tree(z, v, j); n0 = index of this sub-tree's first leaf inside the doubling */
static void build_tree(tree_ctx *c, const phase *z, int v, int j, uint32_t n0, subtree *out) {
  const int D = c->m->D;
  if (j == 0) {
    phase_copy(&out->zl, z, D);
    leapfrog(c->m, &out->zl, c->minv, v * c->eps, c->n_grad);
    phase_copy(&out->zr, &out->zl, D);
    const double dH = out->zl.H - c->H0;
    memcpy(out->rho, out->zl.r, sizeof(double) * (size_t)D);
    const double ma = -dH;
    out->sum_a = (dH != dH) ? 0.0 : exp(ma < 0 ? ma : 0.0); /* exp(min(0,-dH)) */
    out->n_a = 1;
    out->dH_max = dH;
    out->lw = -dH;
    out->t_dyn = 0;
    out->t_num = !(dH < c->cfg->delta_max); /* NaN or >= Delta_max -> divergent */
    if (c->cfg->sampler_mode == 0) {
      phase_copy(&out->zc, &out->zl, D);
    } else {
      const double lwn = logaddexp(c->res_lw, out->lw);
      const double p = (out->lw == -INFINITY) ? 0.0 : exp(out->lw - lwn);
      double u[2];
      orc_uniform2(c->cfg->seed, c->chain, c->iter, ORC_SLOT_LEAF, (1u << c->jtop) + n0, u);
      if (u[0] < p) phase_copy(&c->res_c, &out->zl, D);
      c->res_lw = lwn;
    }
    return;
  }
  build_tree(c, z, v, j - 1, n0, out);
  if (out->t_dyn || out->t_num) return;
  subtree *s2 = &c->pool[j - 1];
  build_tree(c, v == -1 ? &out->zl : &out->zr, v, j - 1, n0 + (1u << (j - 1)), s2);
  /* combine trees */
  if (v == -1) phase_copy(&out->zl, &s2->zl, D);
  else phase_copy(&out->zr, &s2->zr, D);
  for (int d = 0; d < D; ++d) out->rho[d] += s2->rho[d];
  out->sum_a += s2->sum_a;
  out->n_a += s2->n_a;
  if (fabs(s2->dH_max) > fabs(out->dH_max)) out->dH_max = s2->dH_max;
  /* combine samplers */
  const double lw = logaddexp(out->lw, s2->lw);
  if (c->cfg->sampler_mode == 0) {
    double u[2];
    orc_uniform2(c->cfg->seed, c->chain, c->iter, ORC_SLOT_MERGE,
                 ((uint32_t)c->jtop << 25) | ((uint32_t)j << 20) | n0, u);
    if (!(u[0] < exp(out->lw - lw))) phase_copy(&out->zc, &s2->zc, D);
  }
  out->lw = lw;
  /* termination' * termination'' * isterminated(h, tree') */
  out->t_dyn = s2->t_dyn || uturn(D, out->rho, out->zl.r, out->zr.r, c->minv);
  out->t_num = s2->t_num;
}

/* stats layout */
enum { ST_NSTEPS, ST_ACC, ST_LP, ST_H, ST_HERR, ST_MAXHERR, ST_DEPTH, ST_NUMERR, ST_EPS };

/* one NUTS transition from (theta, lp, grad) cached in the chain; momentum refreshed here.
 * AdvancedHMC re-evaluates lp/grad at theta on refresh; the value is identical to the cached
 * one, so we reuse it and do not count a gradient evaluation for it. */
static void nuts_transition(const orc_model *m, const orc_config *cfg, uint32_t chain,
                            orc_chain *c, double *stats) {
  const int D = m->D;
  tree_ctx ctx;
  ctx.m = m;
  ctx.cfg = cfg;
  ctx.minv = c->minv;
  ctx.chain = chain;
  ctx.iter = (uint32_t)(c->iter + 1);
  ctx.eps = c->eps;
  ctx.n_grad = &c->n_grad_evals;
  const int md = cfg->max_depth < ORC_MAX_DEPTH ? cfg->max_depth : ORC_MAX_DEPTH;
  ctx.pool_n = md;
  for (int j = 0; j < md; ++j) subtree_alloc(&ctx.pool[j], D);
  phase_alloc(&ctx.res_c, D);

  subtree tree, sub;
  subtree_alloc(&tree, D);
  subtree_alloc(&sub, D);
  phase z0;
  phase_alloc(&z0, D);
  memcpy(z0.theta, c->theta, sizeof(double) * (size_t)D);
  memcpy(z0.grad, c->grad, sizeof(double) * (size_t)D);
  z0.lp = c->lp;
  draw_momentum(cfg, chain, ctx.iter, ORC_SLOT_MOMENTUM, D, c->minv, z0.r);
  z0.H = energy(D, z0.lp, z0.grad, z0.r, c->minv);
  const double H0 = z0.H;
  ctx.H0 = H0;

  phase_copy(&tree.zl, &z0, D);
  phase_copy(&tree.zr, &z0, D);
  memcpy(tree.rho, z0.r, sizeof(double) * (size_t)D);
  tree.sum_a = 0;
  tree.n_a = 0;
  tree.dH_max = 0;
  double lw = 0.0; /* MultinomialTS(rng, z0): log weight 0 */
  phase zcand;
  phase_alloc(&zcand, D);
  phase_copy(&zcand, &z0, D);
  int t_dyn = 0, t_num = 0, j = 0;
  while (!(t_dyn || t_num) && j < cfg->max_depth) {
    double u[2];
    orc_uniform2(cfg->seed, chain, ctx.iter, ORC_SLOT_DIRECTION, (uint32_t)j, u);
    const int v = (u[0] < 0.5) ? -1 : 1;
    ctx.jtop = j;
    ctx.res_lw = -INFINITY;
    build_tree(&ctx, v == -1 ? &tree.zl : &tree.zr, v, j, 0u, &sub);
    const int term_s = sub.t_dyn || sub.t_num;
    if (!term_s) {
      ++j; /* depth increments only for a non-terminated doubling */
      orc_uniform2(cfg->seed, chain, ctx.iter, ORC_SLOT_MH, (uint32_t)(j - 1), u);
      const double dl = sub.lw - lw;
      if (u[0] < (dl < 0 ? exp(dl) : 1.0)) /* biased progressive: min(1, exp(lw' - lw)) */
        phase_copy(&zcand, cfg->sampler_mode == 0 ? &sub.zc : &ctx.res_c, D);
    }
    /* combine (no matter terminated or not) */
    if (v == -1) phase_copy(&tree.zl, &sub.zl, D);
    else phase_copy(&tree.zr, &sub.zr, D);
    for (int d = 0; d < D; ++d) tree.rho[d] += sub.rho[d];
    tree.sum_a += sub.sum_a;
    tree.n_a += sub.n_a;
    if (fabs(sub.dH_max) > fabs(tree.dH_max)) tree.dH_max = sub.dH_max;
    lw = logaddexp(lw, sub.lw);
    t_dyn = t_dyn || sub.t_dyn || uturn(D, tree.rho, tree.zl.r, tree.zr.r, c->minv);
    t_num = t_num || sub.t_num;
  }
  stats[ST_NSTEPS] = (double)tree.n_a;
  stats[ST_ACC] = tree.sum_a / (double)tree.n_a;
  stats[ST_LP] = zcand.lp;
  stats[ST_H] = zcand.H;
  stats[ST_HERR] = zcand.H - H0;
  stats[ST_MAXHERR] = tree.dH_max;
  stats[ST_DEPTH] = (double)j;
  stats[ST_NUMERR] = (double)t_num;
  stats[ST_EPS] = c->eps;
  memcpy(c->theta, zcand.theta, sizeof(double) * (size_t)D);
  memcpy(c->grad, zcand.grad, sizeof(double) * (size_t)D);
  c->lp = zcand.lp;

  phase_free(&zcand);
  phase_free(&z0);
  subtree_free(&tree);
  subtree_free(&sub);
  phase_free(&ctx.res_c);
  for (int k = 0; k < md; ++k) subtree_free(&ctx.pool[k]);
}

/* HMC / HMCDA: Trajectory{EndPointTS}(Leapfrog, FixedNSteps | FixedIntegrationTime)
 * src/mcmc/hmc.jl:425-434; [dep AdvancedHMC] static trajectory + Metropolis accept. */
static void hmc_transition(const orc_model *m, const orc_config *cfg, uint32_t chain,
                           orc_chain *c, double *stats) {
  const int D = m->D;
  const uint32_t iter = (uint32_t)(c->iter + 1);
  int L = cfg->n_leapfrog;
  if (cfg->algorithm == 2) {
    L = (int)floor(cfg->lambda / c->eps);
    if (L < 1) L = 1;
  }
  phase z0, z;
  phase_alloc(&z0, D);
  phase_alloc(&z, D);
  memcpy(z0.theta, c->theta, sizeof(double) * (size_t)D);
  memcpy(z0.grad, c->grad, sizeof(double) * (size_t)D);
  z0.lp = c->lp;
  draw_momentum(cfg, chain, iter, ORC_SLOT_MOMENTUM, D, c->minv, z0.r);
  z0.H = energy(D, z0.lp, z0.grad, z0.r, c->minv);
  phase_copy(&z, &z0, D);
  int nsteps = 0;
  for (int l = 0; l < L; ++l) {
    leapfrog(m, &z, c->minv, c->eps, &c->n_grad_evals);
    ++nsteps;
    if (!isfinite(z.H)) break;
  }
  const double dH = z.H - z0.H;
  double alpha = (dH != dH) ? 0.0 : exp(-dH < 0 ? -dH : 0.0);
  double u[2];
  orc_uniform2(cfg->seed, chain, iter, ORC_SLOT_MH, 0u, u);
  const int accept = u[0] < alpha;
  const phase *zn = accept ? &z : &z0;
  stats[ST_NSTEPS] = (double)nsteps;
  stats[ST_ACC] = alpha;
  stats[ST_LP] = zn->lp;
  stats[ST_H] = zn->H;
  stats[ST_HERR] = zn->H - z0.H;
  stats[ST_MAXHERR] = dH;
  stats[ST_DEPTH] = (double)accept; /* HMC has no tree depth; column carries is_accept */
  stats[ST_NUMERR] = (double)(!isfinite(z.H));
  stats[ST_EPS] = c->eps;
  memcpy(c->theta, zn->theta, sizeof(double) * (size_t)D);
  memcpy(c->grad, zn->grad, sizeof(double) * (size_t)D);
  c->lp = zn->lp;
  phase_free(&z0);
  phase_free(&z);
}

/* ------------------------------------------------------------------------------------------
 * Adaptation (App. A.5-A.7): StanHMCAdaptor(MassMatrixAdaptor(Diag), StepSizeAdaptor(delta,eps))
 * ------------------------------------------------------------------------------------------ */
int orc_window_splits(int64_t n, int64_t *window_start, int64_t *window_end, int64_t *splits,
                      int max_splits) {
  int64_t init_buffer = 75, term_buffer = 50, base = 25;
  if (n < 20) { /* too short for slow windows: no mass-matrix adaptation */
    *window_start = n + 1;
    *window_end = n;
    return 0;
  }
  if (init_buffer + base + term_buffer > n) {
    init_buffer = (int64_t)(0.15 * (double)n);
    term_buffer = (int64_t)(0.1 * (double)n);
    base = n - init_buffer - term_buffer;
  }
  *window_start = init_buffer + 1; /* 1-based iteration numbers */
  *window_end = n - term_buffer;
  int k = 0;
  int64_t next = init_buffer + base; /* end of the first slow window */
  while (k < max_splits) {
    if (next >= *window_end) { /* last window always ends at window_end */
      splits[k++] = *window_end;
      break;
    }
    splits[k++] = next;
    base *= 2;
    int64_t nn = next + base;
    if (nn != *window_end && nn + 2 * base > *window_end) nn = *window_end; /* stretch the last */
    next = nn;
  }
  return k;
}

static int is_window_end(int64_t i, const int64_t *splits, int ns) {
  for (int k = 0; k < ns; ++k)
    if (splits[k] == i) return 1;
  return 0;
}

static void da_reset(orc_chain *c) {
  c->da_mu = log(10.0 * c->da_eps);
  c->da_m = 0;
  c->da_xbar = 0;
  c->da_hbar = 0;
}

static void da_adapt(orc_chain *c, double delta, double alpha) {
  /* NesterovDualAveraging: gamma=0.05, t0=10, kappa=0.75 */
  const double gamma = 0.05, t0 = 10.0, kappa = 0.75;
  if (alpha > 1.0) alpha = 1.0;
  const int64_t mm = c->da_m + 1;
  const double eta_h = 1.0 / ((double)mm + t0);
  const double hbar = (1.0 - eta_h) * c->da_hbar + eta_h * (delta - alpha);
  const double x = c->da_mu - hbar * sqrt((double)mm) / gamma;
  const double eta_x = pow((double)mm, -kappa);
  const double xbar = (1.0 - eta_x) * c->da_xbar + eta_x * x;
  const double eps = exp(x);
  if (isfinite(eps)) { /* non-finite eps: keep the previous m, eps, x_bar, H_bar (adapt_stepsize!
                          reverts the whole state) */
    c->da_m = mm;
    c->da_hbar = hbar;
    c->da_xbar = xbar;
    c->da_eps = eps;
  }
}

static void welford_push(orc_chain *c, int D, const double *theta) {
  c->wf_n += 1;
  for (int d = 0; d < D; ++d) {
    const double delta = theta[d] - c->wf_mean[d];
    c->wf_mean[d] += delta / (double)c->wf_n;
    c->wf_m2[d] += delta * (theta[d] - c->wf_mean[d]);
  }
}

/* WelfordCov ([dep AdvancedHMC]): delta = s - mu; mu += delta / n; M += (s - mu_new) delta' */
static void welford_cov_push(orc_chain *c, int D, const double *theta) {
  c->wf_n += 1;
  double *delta = (double *)malloc(sizeof(double) * (size_t)D);
  for (int d = 0; d < D; ++d) {
    delta[d] = theta[d] - c->wf_mean[d];
    c->wf_mean[d] += delta[d] / (double)c->wf_n;
  }
  for (int i = 0; i < D; ++i)
    for (int j = 0; j < D; ++j) c->wf_cov[(size_t)i * D + j] += (theta[i] - c->wf_mean[i]) * delta[j];
  free(delta);
}

static void adapt(const orc_config *cfg, int D, orc_chain *c, double alpha) {
  /* AHMC.adapt!(h, kernel, adaptor, i, nadapts, theta, alpha): src/mcmc/hmc.jl:229-241 */
  const int64_t i = c->iter;
  if (i > cfg->n_adapts) return;
  int64_t ws, we, splits[64];
  const int ns = orc_window_splits(cfg->n_adapts, &ws, &we, splits, 64);
  da_adapt(c, cfg->delta, alpha);
  const int wend = is_window_end(i, splits, ns);
  if (i >= ws && i <= we && cfg->metric == 2) {
    /* MassMatrixAdaptor(DenseEuclideanMetric): regularised covariance
     * (n / (n + 5)) M / (n - 1) + 1e-3 (5 / (n + 5)) I, then the metric is renewed (new Cholesky) */
    welford_cov_push(c, D, c->theta);
    if (wend && c->wf_n >= 10) {
      const double n = (double)c->wf_n;
      for (int a = 0; a < D; ++a)
        for (int b = 0; b < D; ++b)
          c->minv_dense[(size_t)a * D + b] = n / ((n + 5.0) * (n - 1.0)) * c->wf_cov[(size_t)a * D + b] +
                                             (a == b ? 1e-3 * (5.0 / (n + 5.0)) : 0.0);
      dense_chol_upper(D, c->minv_dense, c->chol);
      for (int d = 0; d < D; ++d) c->minv[d] = c->minv_dense[(size_t)d * D + d];
    }
  } else if (i >= ws && i <= we) {
    welford_push(c, D, c->theta);
    if (wend && c->wf_n >= 10 && cfg->metric != 1) { /* n_min = 10; UnitMassMatrix: no update */
      const double n = (double)c->wf_n;
      for (int d = 0; d < D; ++d)
        c->minv[d] = n / ((n + 5.0) * (n - 1.0)) * c->wf_m2[d] + 1e-3 * (5.0 / (n + 5.0));
    }
  }
  if (wend) {
    da_reset(c);
    c->wf_n = 0;
    memset(c->wf_mean, 0, sizeof(double) * (size_t)D);
    memset(c->wf_m2, 0, sizeof(double) * (size_t)D);
    if (c->wf_cov) memset(c->wf_cov, 0, sizeof(double) * (size_t)D * D);
  }
  if (i == cfg->n_adapts) c->da_eps = exp(c->da_xbar); /* finalize! */
  c->eps = c->da_eps;
}

/* ------------------------------------------------------------------------------------------
 * find_good_stepsize (App. A.8) ; call site src/mcmc/hmc.jl:190
 * ------------------------------------------------------------------------------------------ */
double orc_find_good_stepsize(const orc_model *m, const orc_config *cfg, uint32_t chain,
                              const double *theta, const double *minv, int64_t *n_grad) {
  const int D = m->D;
  const double a_min = 0.25, a_cross = 0.5, a_max = 0.75, dd = 2.0;
  double eps = 0.1, eps_p = 0.1;
  phase z, zp;
  phase_alloc(&z, D);
  phase_alloc(&zp, D);
  memcpy(z.theta, theta, sizeof(double) * (size_t)D);
  z.lp = orc_logdensity_and_gradient(m, z.theta, z.grad);
  if (n_grad) ++*n_grad;
  draw_momentum(cfg, chain, 0u, ORC_SLOT_EPS_MOMENTUM, D, minv, z.r);
  z.H = energy(D, z.lp, z.grad, z.r, minv);
  const double H = z.H;
  phase_copy(&zp, &z, D);
  leapfrog(m, &zp, minv, eps, n_grad);
  double dH = H - zp.H;
  const int direction = dH > log(a_cross) ? 1 : -1;
  for (int it = 0; it < 100; ++it) {
    eps_p = direction == 1 ? dd * eps : 1.0 / dd * eps;
    phase_copy(&zp, &z, D);
    leapfrog(m, &zp, minv, eps, n_grad);
    dH = H - zp.H;
    if (direction == 1 && !(dH > log(a_cross))) break;
    else if (direction == -1 && !(dH < log(a_cross))) break;
    else eps = eps_p;
  }
  if (eps > eps_p) {
    double t = eps;
    eps = eps_p;
    eps_p = t;
  }
  for (int it = 0; it < 100; ++it) {
    const double mid = 0.5 * (eps + eps_p);
    phase_copy(&zp, &z, D);
    leapfrog(m, &zp, minv, mid, n_grad);
    dH = H - zp.H;
    if (exp(dH) > a_max) eps = mid;
    else if (exp(dH) < a_min) eps_p = mid;
    else {
      eps = mid;
      break;
    }
  }
  phase_free(&z);
  phase_free(&zp);
  return eps;
}

/* ------------------------------------------------------------------------------------------
 * Chain driver (App. C: AbstractMCMC.mcmcsample; src/mcmc/hmc.jl:156-252)
 * ------------------------------------------------------------------------------------------ */
orc_chain *orc_chain_new(int D) {
  orc_chain *c = (orc_chain *)calloc(1, sizeof(orc_chain));
  c->theta = (double *)calloc((size_t)D * 5, sizeof(double));
  c->grad = c->theta + D;
  c->minv = c->grad + D;
  c->wf_mean = c->minv + D;
  c->wf_m2 = c->wf_mean + D;
  return c;
}
void orc_chain_free(orc_chain *c) {
  if (!c) return;
  free(c->theta);
  free(c->minv_dense); /* the three dense arrays are one block */
  free(c);
}

int orc_chain_init(const orc_model *m, const orc_config *cfg, uint32_t chain,
                   const double *theta0, orc_chain *c) {
  const int D = m->D;
  for (int d = 0; d < D; ++d) c->minv[d] = 1.0; /* DiagEuclideanMetric(D): ones */
  tl_dense_minv = tl_dense_chol = NULL;
  if (cfg->metric == 2) { /* DenseEuclideanMetric(D): identity */
    if (!c->minv_dense) {
      c->minv_dense = (double *)calloc((size_t)D * D * 3, sizeof(double));
      c->chol = c->minv_dense + (size_t)D * D;
      c->wf_cov = c->chol + (size_t)D * D;
    }
    memset(c->minv_dense, 0, sizeof(double) * (size_t)D * D * 3);
    for (int d = 0; d < D; ++d) c->minv_dense[(size_t)d * D + d] = c->chol[(size_t)d * D + d] = 1.0;
    tl_dense_minv = c->minv_dense;
    tl_dense_chol = c->chol;
  }
  c->iter = 0;
  c->n_grad_evals = 0;
  c->wf_n = 0;
  int ok = 0;
  const int npair = (D + 1) / 2;
  if (theta0) {
    /* user-supplied initial_params: evaluated once, must be finite (abstractmcmc.jl:42-70
     * with a deterministic strategy retries the same point, so one try decides) */
    memcpy(c->theta, theta0, sizeof(double) * (size_t)D);
    c->lp = orc_logdensity_and_gradient(m, c->theta, c->grad);
    ++c->n_grad_evals;
    ok = isfinite(c->lp);
    for (int d = 0; d < D; ++d) ok = ok && isfinite(c->grad[d]);
    c->init_tries = 1;
  } else {
    for (int t = 0; t < 1000 && !ok; ++t) { /* max_attempts = 1000 */
      for (int p = 0; p < npair; ++p) {
        double u[2];
        orc_uniform2(cfg->seed, chain, 0u, ORC_SLOT_INIT, (uint32_t)(t * npair + p), u);
        c->theta[2 * p] = 4.0 * u[0] - 2.0; /* InitFromUniform: U(-2,2) */
        if (2 * p + 1 < D) c->theta[2 * p + 1] = 4.0 * u[1] - 2.0;
      }
      c->lp = orc_logdensity_and_gradient(m, c->theta, c->grad);
      ++c->n_grad_evals;
      ok = isfinite(c->lp);
      for (int d = 0; d < D; ++d) ok = ok && isfinite(c->grad[d]);
      c->init_tries = t + 1;
    }
  }
  if (!ok) {
    tl_dense_minv = tl_dense_chol = NULL;
    return -1;
  }
  if (cfg->algorithm >= 3) { /* SGHMC / SGLD: velocity zero, no step-size search */
    for (int d = 0; d < D; ++d) c->wf_mean[d] = 0.0;
    c->eps = cfg->init_eps;
  } else if (cfg->init_eps == 0.0) /* iszero(spl.eps): every Hamiltonian sampler, hmc.jl:189-194 */
    c->eps = orc_find_good_stepsize(m, cfg, chain, c->theta, c->minv, &c->n_grad_evals);
  else
    c->eps = cfg->init_eps;
  c->da_eps = c->eps;
  da_reset(c); /* StepSizeAdaptor(delta, eps): mu = log(10 eps) */
  tl_dense_minv = tl_dense_chol = NULL;
  return 0;
}

/* SGHMC (algorithm 3) and SGLD (algorithm 4), src/mcmc/sghmc.jl:79-105 and :223-249.
 * Both take the gradient at the current position (cached from the previous step / the init
 * step), move, and re-evaluate; no accept step, no adaptation.
 *   SGHMC: theta += v;  v = (1 - alpha) v + eta grad + sqrt(2 eta alpha) z   (eq. 15 of Chen et al.)
 *          eta = cfg->init_eps (learning_rate), alpha = cfg->lambda (momentum_decay);
 *          the velocity lives in c->wf_mean (unused otherwise: no mass-matrix adaptation)
 *   SGLD:  eps_t = a / (t + b)^gamma, t = 1, 2, ...  (PolynomialStepsize, sghmc.jl:124-161)
 *          theta += eps_t / 2 grad + sqrt(eps_t) z
 *          a = cfg->init_eps, b = cfg->delta, gamma = cfg->lambda
 * z: the MOMENTUM slot of the addressing specification, iteration = transition number. */
static void sg_transition(const orc_model *m, const orc_config *cfg, uint32_t chain, orc_chain *c,
                          double *stats) {
  const int D = m->D;
  const uint32_t it = (uint32_t)(c->iter + 1);
  double eps;
  if (cfg->algorithm == 3) {
    const double eta = cfg->init_eps, alpha = cfg->lambda, sd = sqrt(2.0 * eta * alpha);
    double *v = c->wf_mean;
    for (int d = 0; d < D; ++d) {
      double z[2];
      orc_normal2(cfg->seed, chain, it, ORC_SLOT_MOMENTUM, (uint32_t)(d >> 1), z);
      c->theta[d] += v[d];
      v[d] = (1.0 - alpha) * v[d] + eta * c->grad[d] + sd * z[d & 1];
    }
    eps = eta;
  } else {
    eps = cfg->init_eps / pow((double)(c->iter + 1) + cfg->delta, cfg->lambda);
    const double sd = sqrt(eps);
    for (int d = 0; d < D; ++d) {
      double z[2];
      orc_normal2(cfg->seed, chain, it, ORC_SLOT_MOMENTUM, (uint32_t)(d >> 1), z);
      c->theta[d] += 0.5 * eps * c->grad[d] + sd * z[d & 1];
    }
  }
  c->lp = orc_logdensity_and_gradient(m, c->theta, c->grad);
  ++c->n_grad_evals;
  for (int q = 0; q < ORC_NSTAT; ++q) stats[q] = NAN;
  stats[ST_NSTEPS] = 1.0;
  stats[ST_LP] = c->lp;
  stats[ST_EPS] = eps; /* SGLD_stepsize (sghmc.jl:238) */
}

void orc_chain_step(const orc_model *m, const orc_config *cfg, uint32_t chain, orc_chain *c,
                    double *stats) {
  if (cfg->algorithm >= 3) {
    sg_transition(m, cfg, chain, c, stats);
    c->iter += 1;
    return;
  }
  if (cfg->metric == 2) {
    tl_dense_minv = c->minv_dense;
    tl_dense_chol = c->chol;
  }
  if (cfg->algorithm == 0) nuts_transition(m, cfg, chain, c, stats);
  else hmc_transition(m, cfg, chain, c, stats);
  tl_dense_minv = tl_dense_chol = NULL;
  c->iter += 1;
  if (cfg->algorithm != 1) adapt(cfg, m->D, c, stats[ST_ACC]);
}

int64_t orc_chain_run(const orc_model *m, const orc_config *cfg, uint32_t chain, orc_chain *c,
                      int64_t n_transitions, int64_t first_keep, int64_t thin, double *out_theta,
                      double *out_params, double *out_logp, double *out_stats) {
  const int D = m->D, P = orc_num_params(m);
  int64_t k = 0;
  double stats[ORC_NSTAT];
  for (int64_t t = 1; t <= n_transitions; ++t) {
    orc_chain_step(m, cfg, chain, c, stats);
    if (t >= first_keep && (t - first_keep) % thin == 0) {
      if (out_theta) memcpy(out_theta + k * D, c->theta, sizeof(double) * (size_t)D);
      if (out_params || out_logp) {
        double lp3[3];
        double *pp = out_params ? out_params + k * P : (double *)malloc(sizeof(double) * (size_t)P);
        orc_constrain_and_logp(m, c->theta, pp, lp3);
        if (out_logp) memcpy(out_logp + k * 3, lp3, sizeof(lp3));
        if (!out_params) free(pp);
      }
      if (out_stats) memcpy(out_stats + k * ORC_NSTAT, stats, sizeof(stats));
      ++k;
    }
  }
  return k;
}

int64_t orc_sample_chain_ids(const orc_model *m, const orc_config *cfg, int n_chains, int n_threads,
                             const uint32_t *chain_ids, const double *theta0, int64_t n_transitions,
                             int64_t first_keep, int64_t thin, double *out_theta, double *out_params,
                             double *out_logp, double *out_stats, double *out_final_eps,
                             double *out_final_minv) {
  const int D = m->D, P = orc_num_params(m);
  int64_t nkeep = 0;
  if (n_transitions >= first_keep && first_keep >= 1) nkeep = (n_transitions - first_keep) / thin + 1;
  int64_t total = 0;
  int fail = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total) reduction(| : fail)
  for (int ch = 0; ch < n_chains; ++ch) {
    const uint32_t id = chain_ids ? chain_ids[ch] : (uint32_t)ch; /* global chain id: RNG address */
    orc_chain *c = orc_chain_new(D);
    if (orc_chain_init(m, cfg, id, theta0 ? theta0 + (size_t)ch * D : NULL, c) != 0) {
      fail |= 1;
    } else {
      orc_chain_run(m, cfg, id, c, n_transitions, first_keep, thin,
                    out_theta ? out_theta + (size_t)ch * nkeep * D : NULL,
                    out_params ? out_params + (size_t)ch * nkeep * P : NULL,
                    out_logp ? out_logp + (size_t)ch * nkeep * 3 : NULL,
                    out_stats ? out_stats + (size_t)ch * nkeep * ORC_NSTAT : NULL);
      if (out_final_eps) out_final_eps[ch] = c->eps;
      if (out_final_minv) memcpy(out_final_minv + (size_t)ch * D, c->minv, sizeof(double) * (size_t)D);
      total += c->n_grad_evals;
    }
    orc_chain_free(c);
  }
  return fail ? -1 : total;
}

int64_t orc_sample_chains(const orc_model *m, const orc_config *cfg, int n_chains, int n_threads,
                          const double *theta0, int64_t n_transitions, int64_t first_keep,
                          int64_t thin, double *out_theta, double *out_params, double *out_logp,
                          double *out_stats, double *out_final_eps, double *out_final_minv) {
  return orc_sample_chain_ids(m, cfg, n_chains, n_threads, NULL, theta0, n_transitions, first_keep, thin,
                              out_theta, out_params, out_logp, out_stats, out_final_eps, out_final_minv);
}

/* Continue already adapted chains (the HMCState of src/mcmc/hmc.jl:10-23 restored, nadapts = 0 as in
 * hmc.jl:135-152): chain ch starts at theta[ch], with step size eps[ch], diagonal M^-1 minv[ch] and
 * iteration counter iter0 (the Philox iteration index of its next transition is iter0 + 1).
 * Used by bench.py's CPU legs to time a bounded block of post-warm-up transitions; global chain
 * ids are chain_id0 + ch.  out_final_theta [chains][D] may be NULL.  Returns the gradient
 * evaluations of this call (one per chain for the restored phase point included). */
int64_t orc_resume_chains(const orc_model *m, const orc_config *cfg, int n_chains, int n_threads,
                          uint32_t chain_id0, const double *theta, const double *eps,
                          const double *minv, int64_t iter0, int64_t n_transitions,
                          int64_t first_keep, int64_t thin, double *out_theta, double *out_params,
                          double *out_logp, double *out_stats, double *out_final_theta) {
  const int D = m->D, P = orc_num_params(m);
  int64_t nkeep = 0;
  if (n_transitions >= first_keep && first_keep >= 1) nkeep = (n_transitions - first_keep) / thin + 1;
  int64_t total = 0;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
  for (int ch = 0; ch < n_chains; ++ch) {
    orc_chain *c = orc_chain_new(D);
    memcpy(c->theta, theta + (size_t)ch * D, sizeof(double) * (size_t)D);
    memcpy(c->minv, minv + (size_t)ch * D, sizeof(double) * (size_t)D);
    c->lp = orc_logdensity_and_gradient(m, c->theta, c->grad);
    c->n_grad_evals = 1;
    c->eps = c->da_eps = eps[ch];
    da_reset(c);
    c->iter = iter0;
    orc_chain_run(m, cfg, chain_id0 + (uint32_t)ch, c, n_transitions, first_keep, thin,
                  out_theta ? out_theta + (size_t)ch * nkeep * D : NULL,
                  out_params ? out_params + (size_t)ch * nkeep * P : NULL,
                  out_logp ? out_logp + (size_t)ch * nkeep * 3 : NULL,
                  out_stats ? out_stats + (size_t)ch * nkeep * ORC_NSTAT : NULL);
    if (out_final_theta) memcpy(out_final_theta + (size_t)ch * D, c->theta, sizeof(double) * (size_t)D);
    total += c->n_grad_evals;
    orc_chain_free(c);
  }
  return total;
}
