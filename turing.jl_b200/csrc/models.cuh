// models.cuh -- log joint density (linked space) + gradient for the small registered families,
// evaluated by ONE thread per chain with the whole parameter vector in registers.
//
// This is the device replacement of `LogDensityProblems.logdensity_and_gradient(ldf, theta)`
// for `ldf = LogDensityFunction(model, getlogjoint_internal, LinkAll())`
// (src/mcmc/hmc.jl:170-182): log prior + log likelihood + log|det J| of the inverse link.
// Gradients are analytic (the reference uses ForwardDiff duals, src/Turing.jl:25).
//
// Families with a data-parallel likelihood (logistic, hierarchical) live in lpgrad_*.cu.
#pragma once
#include "common.cuh"

namespace tnuts {

// Small per-model constants and data carried in kernel-parameter (constant-bank) space so
// that uniform reads cost no memory traffic.
struct SmallData {
  int n;           // observations
  double a[16];    // y
  double b[16];    // x (LINREG) or 1/sigma^2 (EIGHT_SCHOOLS)
  double cst;      // sum of theta-independent terms of lp
  double cst_lik;  // theta-independent part of the likelihood (user-space output)
  double h[4];     // hyper-parameters
  double c[16];    // logit-term families: shift of dimension d (u_d = theta_d - c_d)
  double cnt[16];  // DIRICHLET_CATEGORICAL: observation count of category k
};

// ---- gdemo: s ~ InverseGamma(al, be); m ~ Normal(0, sqrt(s)); x_i ~ Normal(m, sqrt(s)) ----
// theta = (log s, m).  test/test_utils/models.jl:15-21
struct GDemo {
  static constexpr int D = 2;
  static constexpr int P = D;   // user-space parameters
  static constexpr int family = TNUTS_GDEMO;
  __device__ static double lpgrad(const SmallData &md, const double (&th)[2], double (&g)[2]) {
    const double al = md.h[0], be = md.h[1];
    const double a = th[0], m = th[1], is = exp(-a);
    double Q = m * m, sres = 0.0;
    for (int i = 0; i < md.n; ++i) {
      const double d = md.a[i] - m;
      Q += d * d;
      sres += d;
    }
    const double n1 = (double)(md.n + 1);
    const double lp = md.cst - (al + 1.0) * a - be * is - 0.5 * n1 * a - 0.5 * Q * is + a;
    g[0] = -(al + 1.0) + be * is - 0.5 * n1 + 0.5 * Q * is + 1.0;
    g[1] = (sres - m) * is;
    return lp;
  }
  __device__ static void constrain(const SmallData &md, const double (&th)[2], double (&p)[2],
                                   double &prior, double &lik) {
    const double al = md.h[0], be = md.h[1];
    const double a = th[0], m = th[1], s = exp(a);
    p[0] = s;
    p[1] = m;
    // cst = al*log(be) - lgamma(al) - 0.5*(n+1)*log(2pi)
    prior = (md.cst + 0.5 * md.n * kLog2Pi) - (al + 1.0) * a - be / s - 0.5 * a - m * m / (2.0 * s);
    double q = 0;
    for (int i = 0; i < md.n; ++i) {
      const double d = md.a[i] - m;
      q += d * d;
    }
    lik = -0.5 * md.n * kLog2Pi - 0.5 * md.n * a - q / (2.0 * s);
  }
};

// ---- README linear regression (README.md:38-47) ----
// alpha, beta ~ Normal(0, sd); sigma2 ~ truncated(Cauchy(0, c); lower=0);
// y ~ MvNormal(alpha .+ beta .* x, sigma2 * I).  theta = (alpha, beta, log sigma2)
struct LinReg {
  static constexpr int D = 3;
  static constexpr int P = D;   // user-space parameters
  static constexpr int family = TNUTS_LINREG;
  __device__ static double lpgrad(const SmallData &md, const double (&th)[3], double (&g)[3]) {
    const double isd2 = md.h[2], ic = md.h[3];
    const double al = th[0], be = th[1], t = th[2], v = exp(t), iv = 1.0 / v;
    double ssr = 0, sr = 0, srx = 0;
    for (int i = 0; i < md.n; ++i) {
      const double res = md.a[i] - al - be * md.b[i];
      ssr += res * res;
      sr += res;
      srx += res * md.b[i];
    }
    const double n = (double)md.n, vc = v * ic, w = vc * vc;
    const double lp = md.cst - 0.5 * (al * al + be * be) * isd2 - log1p(w) + t - 0.5 * n * t -
                      0.5 * ssr * iv;
    g[0] = -al * isd2 + sr * iv;
    g[1] = -be * isd2 + srx * iv;
    g[2] = -2.0 * w / (1.0 + w) + 1.0 - 0.5 * n + 0.5 * ssr * iv;
    return lp;
  }
  __device__ static void constrain(const SmallData &md, const double (&th)[3], double (&p)[3],
                                   double &prior, double &lik) {
    const double isd2 = md.h[2], ic = md.h[3];
    const double v = exp(th[2]), vc = v * ic, w = vc * vc;
    p[0] = th[0];
    p[1] = th[1];
    p[2] = v;
    // cst = -log(2pi) - 2 log(sd) + log 2 - log pi - log c - 0.5 n log(2pi)
    prior = (md.cst - md.cst_lik) - 0.5 * (th[0] * th[0] + th[1] * th[1]) * isd2 - log1p(w);
    double ssr = 0;
    for (int i = 0; i < md.n; ++i) {
      const double res = md.a[i] - th[0] - th[1] * md.b[i];
      ssr += res * res;
    }
    lik = md.cst_lik - 0.5 * md.n * th[2] - ssr / (2.0 * v);
  }
};

// ---- eight schools, non-centred ----
// mu ~ Normal(0, sd); tau ~ truncated(Cauchy(0, c); lower=0); thetat ~ MvNormal(0, I);
// y_j ~ Normal(mu + tau*thetat_j, sigma_j).   theta = (mu, log tau, thetat_1..J)
template <int J>
struct EightSchools {
  static constexpr int D = J + 2;
  static constexpr int P = D;   // user-space parameters
  static constexpr int family = TNUTS_EIGHT_SCHOOLS;
  __device__ static double lpgrad(const SmallData &md, const double (&th)[D], double (&g)[D]) {
    const double isd2 = md.h[2], ic = md.h[3];
    const double mu = th[0], lt = th[1], tau = exp(lt), tc = tau * ic, w = tc * tc;
    double lp = md.cst - 0.5 * mu * mu * isd2 - log1p(w) + lt;
    double sz = 0, szt = 0;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const double tj = th[2 + j];
      const double res = md.a[j] - mu - tau * tj;
      const double z = res * md.b[j];
      lp += -0.5 * tj * tj - 0.5 * res * z;
      sz += z;
      szt += z * tj;
      g[2 + j] = -tj + tau * z;
    }
    g[0] = -mu * isd2 + sz;
    g[1] = -2.0 * w / (1.0 + w) + 1.0 + tau * szt;
    return lp;
  }
  __device__ static void constrain(const SmallData &md, const double (&th)[D], double (&p)[D],
                                   double &prior, double &lik) {
    const double isd2 = md.h[2], ic = md.h[3];
    const double mu = th[0], tau = exp(th[1]), tc = tau * ic, w = tc * tc;
    p[0] = mu;
    p[1] = tau;
    prior = (md.cst - md.cst_lik) - 0.5 * mu * mu * isd2 - log1p(w);
    lik = md.cst_lik;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const double tj = th[2 + j];
      const double res = md.a[j] - mu - tau * tj;
      p[2 + j] = tj;
      prior += -0.5 * tj * tj;
      lik += -0.5 * res * res * md.b[j];
    }
  }
};

// ---- x ~ Normal() iid (test/main.jl:39-48) ----
template <int D_>
struct StdNormal {
  static constexpr int D = D_;
  static constexpr int P = D;   // user-space parameters
  static constexpr int family = TNUTS_STD_NORMAL;
  __device__ static double lpgrad(const SmallData &md, const double (&th)[D], double (&g)[D]) {
    double lp = md.cst;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      lp -= 0.5 * th[d] * th[d];
      g[d] = -th[d];
    }
    return lp;
  }
  __device__ static void constrain(const SmallData &md, const double (&th)[D], double (&p)[D],
                                   double &prior, double &lik) {
    prior = md.cst;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      p[d] = th[d];
      prior -= 0.5 * th[d] * th[d];
    }
    lik = 0.0;
  }
};

// ---- constrained-support models of the reference's HMC tests (test/mcmc/hmc.jl:24-62) ----
// Both log joints are sums of  A_d log sigma(u_d) + B_d log sigma(-u_d),  u_d = theta_d - c_d
// (logit link for p in (0, 1); stick breaking for a simplex: see oracle/tnuts_oracle.c and
// tests/golden/make_golden.py, which derives the same densities from the distribution definitions).
__device__ __forceinline__ double log_sigmoid(double u) {
  return u >= 0.0 ? -log1p(exp(-u)) : u - log1p(exp(u));
}
__device__ __forceinline__ double sigmoid(double u) {
  if (u >= 0.0) return 1.0 / (1.0 + exp(-u));
  const double e = exp(u);
  return e / (1.0 + e);
}

// p ~ Beta(a, b); obs[i] ~ Bernoulli(p).  theta = logit(p).  md.a[0] = a + #ones, md.b[0] = b + #zeros,
// md.h = {a, b, #ones, #zeros}, md.cst = -log B(a, b)
struct BetaBernoulli {
  static constexpr int D = 1;
  static constexpr int P = 1;
  static constexpr int family = TNUTS_BETA_BERNOULLI;
  __device__ static double lpgrad(const SmallData &md, const double (&th)[1], double (&g)[1]) {
    const double u = th[0], p = sigmoid(u);
    g[0] = md.a[0] * (1.0 - p) - md.b[0] * p;
    return md.a[0] * log_sigmoid(u) + md.b[0] * log_sigmoid(-u) + md.cst;
  }
  __device__ static void constrain(const SmallData &md, const double (&th)[1], double (&p)[1], double &prior,
                                   double &lik) {
    const double lpp = log_sigmoid(th[0]), lq = log_sigmoid(-th[0]);
    p[0] = sigmoid(th[0]);
    prior = (md.h[0] - 1.0) * lpp + (md.h[1] - 1.0) * lq + md.cst;
    lik = md.h[2] * lpp + md.h[3] * lq;
  }
};

// ps ~ Dirichlet(K1, al1); pd ~ Dirichlet(K2, al2); obs[i] ~ Categorical(ps).  theta = stick-breaking
// coordinates of ps, then of pd.  md.a / md.b / md.c per linked dimension, md.cnt[k] = count of
// category k, md.h = {al1, al2}, md.cst = sum over the blocks of lgamma(K al) - K lgamma(al)
template <int K1, int K2>
struct DirichletCat {
  static constexpr int D = K1 - 1 + (K2 > 0 ? K2 - 1 : 0);
  static constexpr int P = K1 + K2;
  static constexpr int family = TNUTS_DIRICHLET_CATEGORICAL;
  __device__ static double lpgrad(const SmallData &md, const double (&th)[D], double (&g)[D]) {
    double lp = md.cst;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double u = th[d] - md.c[d], z = sigmoid(u);
      lp += md.a[d] * log_sigmoid(u) + md.b[d] * log_sigmoid(-u);
      g[d] = md.a[d] * (1.0 - z) - md.b[d] * z;
    }
    return lp;
  }
  template <int K>
  __device__ static void invlink(const SmallData &md, const double *y, const double *shift, double *x) {
    double rem = 1.0;
#pragma unroll
    for (int j = 0; j < K - 1; ++j) {
      const double z = sigmoid(y[j] - shift[j]);
      x[j] = z * rem;
      rem *= 1.0 - z;
    }
    x[K - 1] = rem;
  }
  __device__ static void constrain(const SmallData &md, const double (&th)[D], double (&p)[P], double &prior,
                                   double &lik) {
    invlink<K1>(md, th, md.c, p);
    prior = md.cst;
    lik = 0.0;
#pragma unroll
    for (int k = 0; k < K1; ++k) {
      const double l = log(p[k]);
      prior += (md.h[0] - 1.0) * l;
      lik += md.cnt[k] * l;
    }
    if (K2 > 0) {
      invlink<(K2 > 0 ? K2 : 1)>(md, th + (K1 - 1), md.c + (K1 - 1), p + K1);
#pragma unroll
      for (int k = 0; k < K2; ++k) prior += (md.h[1] - 1.0) * log(p[K1 + k]);
    }
  }
};

using DirichletCat24 = DirichletCat<2, 4>;   // test/mcmc/hmc.jl:45-62
using DirichletCat30 = DirichletCat<3, 0>;
using DirichletCat20 = DirichletCat<2, 0>;

}  // namespace tnuts
