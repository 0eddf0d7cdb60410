// chain_kernels.cuh -- "one thread per chain" execution strategy (tiny-D families).
//
// A whole NUTS transition -- momentum refresh, tree doubling, leapfrogs, gradient,
// multinomial sampling, generalised U-turn, divergence check, dual averaging, Welford
// mass-matrix update, constrained output -- runs inside ONE thread with the position,
// momentum and gradient in registers; a single launch advances every chain by many
// transitions.  Chains are laid out chain-fastest ([D][C]) so every state load/store and
// every output row is a coalesced 256-byte warp transaction.
//
// Replaces, per chain (paths into Turing.jl v0.46.1; [dep] = AdvancedHMC 0.8.3 restated in
// SURVEY.md App. A):
//   AbstractMCMC.step(rng, model, spl, state)       src/mcmc/hmc.jl:210-252
//   AHMC.transition(rng, h, kernel, z)        [dep] call site src/mcmc/hmc.jl:225
//   AHMC.adapt!(h, kernel, adaptor, i, n, th, a) [dep] call site src/mcmc/hmc.jl:229-241
//   DynamicPPL.ParamsWithStats(theta, ldf, stat)    src/mcmc/hmc.jl:247
//   AbstractMCMC.step(rng, model, spl; ...)  (init) src/mcmc/hmc.jl:156-208
//   find_initial_params_ldf                         src/mcmc/abstractmcmc.jl:42-70
//   AHMC.find_good_stepsize                   [dep] call site src/mcmc/hmc.jl:190
//
// The tree is built iteratively (no recursion, no per-level candidate copies):
//   * sub-tree U-turn checks use per-level checkpoints of (r, cumulative rho): leaf n stores
//     at level popc(n>>1) when even and checks its ctz(~n) enclosing aligned sub-trees when
//     odd -- exactly the set of combine() checks of the recursive build_tree;
//   * the candidate inside a doubling is chosen by streaming (reservoir) multinomial
//     sampling -- leaf n replaces the candidate w.p. w_n / sum_{k<=n} w_k -- which has the
//     same law as AdvancedHMC's per-merge uniform progressive sampling;
//   * the top-level step is AdvancedHMC's biased progressive sampling, min(1, exp(lw'-lw)).
#pragma once
#include "common.cuh"
#include "models.cuh"
#include "philox.cuh"

namespace tnuts {

constexpr int kChainBlock = 128;
constexpr int kMaxCk = 12;  // checkpoint levels: max_depth <= kMaxCk + 1 for this strategy
// the lowest checkpoint levels are the busy ones (level l is written every 2^(l+1) leaves and read
// about as often): they live in shared memory, [level][r | rho][d][thread], the others in local
constexpr int kCkSm = 4;
template <int D>
constexpr size_t chain_ck_smem_bytes() { return sizeof(double) * kCkSm * 2 * D * 128; }

__device__ __forceinline__ double logaddexp_d(double a, double b) {
  if (a == -CUDART_INF) return b;
  if (b == -CUDART_INF) return a;
  const double mx = fmax(a, b);
  return mx + log1p(exp(-fabs(a - b)));
}

template <int D>
__device__ __forceinline__ double energy_d(double lp, const double (&g)[D], const double (&r)[D],
                                           const double (&minv)[D]) {
  double k = 0;
  bool ok = isfinite(lp);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    k += minv[d] * r[d] * r[d];
    ok = ok && isfinite(g[d]);
  }
  k *= 0.5;
  return (ok && isfinite(k)) ? (-lp + k) : CUDART_INF;
}

// one leapfrog step of signed size eps: half kick, drift, gradient, half kick
template <class M>
__device__ __forceinline__ void leapfrog_d(const SmallData &md, double (&th)[M::D],
                                           double (&r)[M::D], double (&g)[M::D],
                                           const double (&minv)[M::D], double eps, double &lp,
                                           double &H) {
  constexpr int D = M::D;
  const double half = 0.5 * eps;
#pragma unroll
  for (int d = 0; d < D; ++d) r[d] = r[d] + half * g[d];
#pragma unroll
  for (int d = 0; d < D; ++d) th[d] = th[d] + eps * (minv[d] * r[d]);
  lp = M::lpgrad(md, th, g);
#pragma unroll
  for (int d = 0; d < D; ++d) r[d] = r[d] + half * g[d];
  H = energy_d<D>(lp, g, r, minv);
}

template <int D>
__device__ __forceinline__ void draw_momentum_d(unsigned long long seed, uint32_t chain,
                                                uint32_t iter, uint32_t slot,
                                                const double (&minv)[D], double (&r)[D]) {
#pragma unroll
  for (int p = 0; 2 * p < D; ++p) {
    double z0, z1;
    normal2(seed, chain, iter, slot, (uint32_t)p, z0, z1);
    r[2 * p] = z0 / sqrt(minv[2 * p]);
    if (2 * p + 1 < D) r[2 * p + 1] = z1 / sqrt(minv[2 * p + 1]);
  }
}

template <int D>
__device__ __forceinline__ bool uturn_d(const double (&rho)[D], const double *rl, const double *rr,
                                        const double (&minv)[D]) {
  double a = 0, b = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    a += rho[d] * (minv[d] * rl[d]);
    b += rho[d] * (minv[d] * rr[d]);
  }
  return (a <= 0) || (b <= 0);
}

// ------------------------------------------------------------------------------------------
// DenseEuclideanMetric (metricT of NUTS/HMC/HMCDA, src/mcmc/hmc.jl:178-179; AdvancedHMC semantics
// restated in the oracle): M^-1 and the upper Cholesky factor U of M^-1 (M^-1 = U'U), row-major,
// per chain.  The functions below overload the diagonal ones on the metric argument, so the
// transition code is written once (template parameter MT).
// ------------------------------------------------------------------------------------------
template <int D>
struct DenseMetric {
  double minv[D * D];
  double U[D * D];
};

template <int D>
__device__ __forceinline__ void dense_apply_d(const DenseMetric<D> &met, const double *r, double (&out)[D]) {
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double s = 0;
#pragma unroll
    for (int j = 0; j < D; ++j) s += met.minv[i * D + j] * r[j];
    out[i] = s;
  }
}

template <int D>
__device__ __forceinline__ double energy_d(double lp, const double (&g)[D], const double (&r)[D],
                                           const DenseMetric<D> &met) {
  double k = 0;
  bool ok = isfinite(lp);
  double v[D];
  dense_apply_d<D>(met, r, v);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    k += r[d] * v[d];
    ok = ok && isfinite(g[d]);
  }
  k *= 0.5;
  return (ok && isfinite(k)) ? (-lp + k) : CUDART_INF;
}

template <class M>
__device__ __forceinline__ void leapfrog_d(const SmallData &md, double (&th)[M::D],
                                           double (&r)[M::D], double (&g)[M::D],
                                           const DenseMetric<M::D> &met, double eps, double &lp,
                                           double &H) {
  constexpr int D = M::D;
  const double half = 0.5 * eps;
#pragma unroll
  for (int d = 0; d < D; ++d) r[d] = r[d] + half * g[d];
  double v[D];
  dense_apply_d<D>(met, r, v);
#pragma unroll
  for (int d = 0; d < D; ++d) th[d] = th[d] + eps * v[d];
  lp = M::lpgrad(md, th, g);
#pragma unroll
  for (int d = 0; d < D; ++d) r[d] = r[d] + half * g[d];
  H = energy_d<D>(lp, g, r, met);
}

// r = U \ randn(D): back substitution
template <int D>
__device__ __forceinline__ void draw_momentum_d(unsigned long long seed, uint32_t chain,
                                                uint32_t iter, uint32_t slot,
                                                const DenseMetric<D> &met, double (&r)[D]) {
#pragma unroll
  for (int p = 0; 2 * p < D; ++p) {
    double z0, z1;
    normal2(seed, chain, iter, slot, (uint32_t)p, z0, z1);
    r[2 * p] = z0;
    if (2 * p + 1 < D) r[2 * p + 1] = z1;
  }
#pragma unroll
  for (int d = D - 1; d >= 0; --d) {
    double t = r[d];
#pragma unroll
    for (int e = d + 1; e < D; ++e) t -= met.U[d * D + e] * r[e];
    r[d] = t / met.U[d * D + d];
  }
}

template <int D>
__device__ __forceinline__ bool uturn_d(const double (&rho)[D], const double *rl, const double *rr,
                                        const DenseMetric<D> &met) {
  double a = 0, b = 0;
  double vl[D], vr[D];
  dense_apply_d<D>(met, rl, vl);
  dense_apply_d<D>(met, rr, vr);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    a += rho[d] * vl[d];
    b += rho[d] * vr[d];
  }
  return (a <= 0) || (b <= 0);
}

// upper Cholesky factor of met.minv into met.U (same loop order as the oracle's dense_chol_upper)
template <int D>
__device__ __forceinline__ void chol_upper_d(DenseMetric<D> &met) {
  for (int i = 0; i < D * D; ++i) met.U[i] = 0.0;
  for (int j = 0; j < D; ++j) {
    double s = met.minv[j * D + j];
    for (int k = 0; k < j; ++k) s -= met.U[k * D + j] * met.U[k * D + j];
    const double ujj = sqrt(s);
    met.U[j * D + j] = ujj;
    for (int i = j + 1; i < D; ++i) {
      double t = met.minv[j * D + i];
      for (int k = 0; k < j; ++k) t -= met.U[k * D + j] * met.U[k * D + i];
      met.U[j * D + i] = t / ujj;
    }
  }
}

// ------------------------------------------------------------------------------------------
// NUTS transition.  On entry th/g/lp = current point; on exit = the selected candidate.
// ------------------------------------------------------------------------------------------
template <class M, class MT>
__device__ __forceinline__ void nuts_transition_d(const SmallData &md, const RunParams &rp,
                                                  uint32_t gchain, uint32_t it,
                                                  double (&th)[M::D], double (&g)[M::D],
                                                  double &lp, const MT &minv,
                                                  double eps, long long &ngrad,
                                                  double (&stats)[TNUTS_NSTAT],
                                                  double *__restrict__ ck_sm /* smem + tid, or null */) {
  constexpr int D = M::D;
  double r[D];
  draw_momentum_d<D>(rp.seed, gchain, it, SLOT_MOMENTUM, minv, r);
  const double H0 = energy_d<D>(lp, g, r, minv);

  // tree edges: [0] = left (v = -1), [1] = right (v = +1); each holds theta | r | grad
  double edge[2][3 * D];
  double elp[2];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    edge[0][d] = edge[1][d] = th[d];
    edge[0][D + d] = edge[1][D + d] = r[d];
    edge[0][2 * D + d] = edge[1][2 * D + d] = g[d];
  }
  elp[0] = elp[1] = lp;

  double rho[D];
#pragma unroll
  for (int d = 0; d < D; ++d) rho[d] = r[d];
  double sum_a = 0.0, dHmax = 0.0, lw = 0.0;
  long long n_a = 0;
  // candidate of the whole tree (starts at z0)
  double thC[D], gC[D], lpC = lp, HC = H0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    thC[d] = th[d];
    gC[d] = g[d];
  }
  // per-level checkpoints for the sub-tree U-turn checks
  double ck_r[kMaxCk][D], ck_rho[kMaxCk][D];
  const int sm_levels = ck_sm ? kCkSm : 0;

  int j = 0;
  bool term = false, numerr = false;
  while (!term && j < rp.max_depth) {
    const int v = (uniform1(rp.seed, gchain, it, SLOT_DIRECTION, (uint32_t)j) < 0.5) ? 0 : 1;
    const double se = v ? eps : -eps;
    double zt[D], zr[D], zg[D], zlp = elp[v], zH;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      zt[d] = edge[v][d];
      zr[d] = edge[v][D + d];
      zg[d] = edge[v][2 * D + d];
    }
    double rho_s[D];
#pragma unroll
    for (int d = 0; d < D; ++d) rho_s[d] = 0.0;
    double lw_s = -CUDART_INF;
    double thS[D], gS[D], lpS = 0.0, HS = 0.0;
    bool term_s = false;
    const int nleaf = 1 << j;
    for (int n = 0; n < nleaf; ++n) {
      leapfrog_d<M>(md, zt, zr, zg, minv, se, zlp, zH);
      ++ngrad;
      const double dH = zH - H0;
      const double a = exp(fmin(0.0, -dH));
      sum_a += a;
      ++n_a;
      if (fabs(dH) > fabs(dHmax)) dHmax = dH;
      // streaming multinomial candidate
      const double lw_leaf = -dH;
      const double lw_new = logaddexp_d(lw_s, lw_leaf);
      const double p = (lw_leaf == -CUDART_INF) ? 0.0 : exp(lw_leaf - lw_new);
      const double u = uniform1(rp.seed, gchain, it, SLOT_LEAF, (1u << j) + (uint32_t)n);
      if (u < p) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          thS[d] = zt[d];
          gS[d] = zg[d];
        }
        lpS = zlp;
        HS = zH;
      }
      lw_s = lw_new;
#pragma unroll
      for (int d = 0; d < D; ++d) rho_s[d] += zr[d];
      if (!(dH < rp.delta_max)) {  // divergence (NaN-safe)
        numerr = true;
        term_s = true;
        break;
      }
      const int idx_max = __popc((unsigned)n >> 1);
      if ((n & 1) == 0) {
        if (idx_max < sm_levels) {
#pragma unroll
          for (int d = 0; d < D; ++d) {
            ck_sm[((idx_max * 2 + 0) * D + d) * kChainBlock] = zr[d];
            ck_sm[((idx_max * 2 + 1) * D + d) * kChainBlock] = rho_s[d];
          }
        } else {
#pragma unroll
          for (int d = 0; d < D; ++d) {
            ck_r[idx_max][d] = zr[d];
            ck_rho[idx_max][d] = rho_s[d];
          }
        }
      } else {
        const int ntrail = __ffs(~(unsigned)n) - 1;  // trailing ones of n
        const int idx_min = idx_max - ntrail + 1;
        for (int i = idx_max; i >= idx_min; --i) {
          double sr[D], cr[D];
          if (i < sm_levels) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
              cr[d] = ck_sm[((i * 2 + 0) * D + d) * kChainBlock];
              sr[d] = rho_s[d] - ck_sm[((i * 2 + 1) * D + d) * kChainBlock] + cr[d];
            }
          } else {
#pragma unroll
            for (int d = 0; d < D; ++d) {
              cr[d] = ck_r[i][d];
              sr[d] = rho_s[d] - ck_rho[i][d] + cr[d];
            }
          }
          if (uturn_d<D>(sr, cr, zr, minv)) {
            term_s = true;
            break;
          }
        }
        if (term_s) break;
      }
    }
    // the new sub-tree becomes the tree's edge on side v (combine)
#pragma unroll
    for (int d = 0; d < D; ++d) {
      edge[v][d] = zt[d];
      edge[v][D + d] = zr[d];
      edge[v][2 * D + d] = zg[d];
    }
    elp[v] = zlp;
    if (!term_s) {
      const double u = uniform1(rp.seed, gchain, it, SLOT_MH, (uint32_t)j);
      ++j;
      const double dl = lw_s - lw;
      if (u < (dl < 0.0 ? exp(dl) : 1.0)) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
          thC[d] = thS[d];
          gC[d] = gS[d];
        }
        lpC = lpS;
        HC = HS;
      }
    }
#pragma unroll
    for (int d = 0; d < D; ++d) rho[d] += rho_s[d];
    lw = logaddexp_d(lw, lw_s);
    term = term_s || uturn_d<D>(rho, &edge[0][D], &edge[1][D], minv);
  }
  stats[ST_NSTEPS] = (double)n_a;
  stats[ST_ACC] = sum_a / (double)n_a;
  stats[ST_LP] = lpC;
  stats[ST_H] = HC;
  stats[ST_HERR] = HC - H0;
  stats[ST_MAXHERR] = dHmax;
  stats[ST_DEPTH] = (double)j;
  stats[ST_NUMERR] = numerr ? 1.0 : 0.0;
  stats[ST_EPS] = eps;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    th[d] = thC[d];
    g[d] = gC[d];
  }
  lp = lpC;
}

// HMC / HMCDA: static trajectory + Metropolis accept (src/mcmc/hmc.jl:425-434)
template <class M, class MT>
__device__ __forceinline__ void hmc_transition_d(const SmallData &md, const RunParams &rp,
                                                 uint32_t gchain, uint32_t it, double (&th)[M::D],
                                                 double (&g)[M::D], double &lp,
                                                 const MT &minv, double eps,
                                                 long long &ngrad, double (&stats)[TNUTS_NSTAT]) {
  constexpr int D = M::D;
  int L = rp.n_leapfrog;
  if (rp.algorithm == TNUTS_ALG_HMCDA) {
    L = (int)floor(rp.lambda / eps);
    if (L < 1) L = 1;
  }
  double r[D];
  draw_momentum_d<D>(rp.seed, gchain, it, SLOT_MOMENTUM, minv, r);
  const double H0 = energy_d<D>(lp, g, r, minv);
  double zt[D], zg[D], zlp = lp, zH = H0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    zt[d] = th[d];
    zg[d] = g[d];
  }
  int nsteps = 0;
  for (int l = 0; l < L; ++l) {
    leapfrog_d<M>(md, zt, r, zg, minv, eps, zlp, zH);
    ++ngrad;
    ++nsteps;
    if (!isfinite(zH)) break;
  }
  const double dH = zH - H0;
  const double alpha = exp(fmin(0.0, -dH));
  const double u = uniform1(rp.seed, gchain, it, SLOT_MH, 0u);
  const bool accept = u < alpha;
  if (accept) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      th[d] = zt[d];
      g[d] = zg[d];
    }
    lp = zlp;
  }
  stats[ST_NSTEPS] = (double)nsteps;
  stats[ST_ACC] = alpha;
  stats[ST_LP] = lp;
  stats[ST_H] = accept ? zH : H0;
  stats[ST_HERR] = accept ? dH : 0.0;
  stats[ST_MAXHERR] = dH;
  stats[ST_DEPTH] = accept ? 1.0 : 0.0;
  stats[ST_NUMERR] = isfinite(zH) ? 0.0 : 1.0;
  stats[ST_EPS] = eps;
}

// ------------------------------------------------------------------------------------------
// run kernel: n_transitions for every chain
// ------------------------------------------------------------------------------------------
// transitions (t_begin, t_end] of this launch for the chain in `slot`; state is loaded from and
// stored back to the ChainState arrays, so a launch can be cut into ranges that run on different
// SMs (k_chain_run_dyn).  State loads bypass L1 (ld.cg): another SM may have written them.
template <class M, bool IS_NUTS>
__device__ __forceinline__ void chain_run_range(const SmallData &md, const ChainState &st,
                                                const DrawBuffers &out, const RunParams &rp,
                                                const int *__restrict__ perm, int slot,
                                                long long t_begin, long long t_end,
                                                double *__restrict__ ck_sm) {
  constexpr int D = M::D;
  const int C = st.C;
  if (slot >= C) return;
  // lane -> chain map: after warm-up the chains are sorted by step size so that the 32 chains
  // of a warp build trees of similar length (a chain's own result does not depend on its lane:
  // every random draw is addressed by the chain id)
  const int c = perm ? perm[slot] : slot;
  if (st.status[c] != 0) return;
  const uint32_t gchain = (uint32_t)(rp.chain_offset + c);

  double th[D], g[D], minv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    th[d] = __ldcg(st.theta + (size_t)d * C + c);
    g[d] = __ldcg(st.grad + (size_t)d * C + c);
    minv[d] = __ldcg(st.minv + (size_t)d * C + c);
  }
  double lp = __ldcg(st.lp + c), eps = __ldcg(st.eps + c);
  long long iter = __ldcg(st.iter + c), ngrad = __ldcg(st.n_grad + c);
  double da_mu = __ldcg(st.da_mu + c), da_xbar = __ldcg(st.da_xbar + c), da_hbar = __ldcg(st.da_hbar + c),
         da_eps = __ldcg(st.da_eps + c);
  long long da_m = __ldcg(st.da_m + c), wf_n = __ldcg(st.wf_n + c);

  for (long long t = t_begin + 1; t <= t_end; ++t) {
    const uint32_t it = (uint32_t)(iter + 1);
    double stats[TNUTS_NSTAT];
    if (IS_NUTS)  // compile-time: the NUTS kernel carries no HMC code and vice versa
      nuts_transition_d<M>(md, rp, gchain, it, th, g, lp, minv, eps, ngrad, stats, ck_sm);
    else
      hmc_transition_d<M>(md, rp, gchain, it, th, g, lp, minv, eps, ngrad, stats);
    iter += 1;

    // ---- AHMC.adapt!: StanHMCAdaptor(MassMatrixAdaptor, StepSizeAdaptor) ----
    if (rp.algorithm != TNUTS_ALG_HMC && iter <= rp.n_adapts) {
      {  // Nesterov dual averaging: gamma = 0.05, t0 = 10, kappa = 0.75
        const double alpha = fmin(stats[ST_ACC], 1.0);
        const long long mm = da_m + 1;
        const double eta_h = 1.0 / ((double)mm + 10.0);
        const double hbar = (1.0 - eta_h) * da_hbar + eta_h * (rp.delta - alpha);
        const double x = da_mu - hbar * sqrt((double)mm) / 0.05;
        const double eta_x = pow((double)mm, -0.75);
        const double xbar = (1.0 - eta_x) * da_xbar + eta_x * x;
        const double e = exp(x);
        if (isfinite(e)) {  // a non-finite step size reverts m, x_bar, H_bar and eps (AdvancedHMC adapt_stepsize!)
          da_m = mm;
          da_hbar = hbar;
          da_xbar = xbar;
          da_eps = e;
        }
      }
      bool wend = false;
      for (int k = 0; k < rp.n_splits; ++k) wend = wend || (rp.splits[k] == iter);
      if (iter >= rp.window_start && iter <= rp.window_end) {
        wf_n += 1;
        const double n = (double)wf_n;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const size_t o = (size_t)d * C + c;
          double mean = __ldcg(st.wf_mean + o), m2 = __ldcg(st.wf_m2 + o);
          const double delta = th[d] - mean;
          mean += delta / n;
          m2 += delta * (th[d] - mean);
          if (wend) {
            if (wf_n >= 10 && !rp.unit_metric) minv[d] = n / ((n + 5.0) * (n - 1.0)) * m2 + 1e-3 * (5.0 / (n + 5.0));
            mean = 0.0;
            m2 = 0.0;
          }
          st.wf_mean[o] = mean;
          st.wf_m2[o] = m2;
        }
      }
      if (wend) {
        da_mu = log(10.0 * da_eps);
        da_m = 0;
        da_xbar = 0.0;
        da_hbar = 0.0;
        wf_n = 0;
      }
      if (iter == rp.n_adapts) da_eps = exp(da_xbar);
      eps = da_eps;
    }

    // ---- record (ParamsWithStats): constrained params, user-space log-probs, stats ----
    if (rp.first_keep >= 1 && t >= rp.first_keep && (t - rp.first_keep) % rp.thin == 0) {
      const long long k = (t - rp.first_keep) / rp.thin;
      constexpr int P = M::P;   // user-space parameters (a K-simplex: K columns for K - 1 linked dims)
      double p[P], prior, lik;
      M::constrain(md, th, p, prior, lik);
      double *row = out.draws + (size_t)k * out.ncol * C + c;
#pragma unroll
      for (int d = 0; d < P; ++d) row[(size_t)d * C] = p[d];
      row[(size_t)(P + 0) * C] = prior;
      row[(size_t)(P + 1) * C] = lik;
      row[(size_t)(P + 2) * C] = prior + lik;
#pragma unroll
      for (int s = 0; s < TNUTS_NSTAT; ++s) row[(size_t)(P + 3 + s) * C] = stats[s];
      if (out.theta != nullptr) {
#pragma unroll
        for (int d = 0; d < D; ++d) out.theta[((size_t)k * D + d) * C + c] = th[d];
      }
    }
  }

#pragma unroll
  for (int d = 0; d < D; ++d) {
    st.theta[(size_t)d * C + c] = th[d];
    st.grad[(size_t)d * C + c] = g[d];
    st.minv[(size_t)d * C + c] = minv[d];
  }
  st.lp[c] = lp;
  st.eps[c] = eps;
  st.iter[c] = iter;
  st.n_grad[c] = ngrad;
  st.da_mu[c] = da_mu;
  st.da_xbar[c] = da_xbar;
  st.da_hbar[c] = da_hbar;
  st.da_eps[c] = da_eps;
  st.da_m[c] = da_m;
  st.wf_n[c] = wf_n;
}

template <class M, int MINB, bool IS_NUTS>
__global__ void __launch_bounds__(kChainBlock, MINB) k_chain_run(SmallData md, ChainState st,
                                                                DrawBuffers out, RunParams rp,
                                                                const int *__restrict__ perm,
                                                                int use_ck_sm) {
  extern __shared__ double ck_smem[];
  chain_run_range<M, IS_NUTS>(md, st, out, rp, perm, blockIdx.x * blockDim.x + threadIdx.x, 0,
                              rp.n_transitions, (IS_NUTS && use_ck_sm) ? ck_smem + threadIdx.x : nullptr);
}

// The same run with dynamic scheduling: 65 536 chains are 512 groups of 128 chains, but only
// 2 x 148 CTAs are resident, so a static grid runs 1.73 waves (the second one 73 % full) and a
// group with long trees finishes last.  Here one resident CTA per slot pulls tasks
// (sub-range s of the launch's transitions, chain group g) from a counter in order of s; task
// (s, g) waits for (s - 1, g), which an earlier-claiming -- hence running -- CTA owns, so there is
// no deadlock.  A chain's state travels through the ChainState arrays; its result does not depend
// on where or when a task runs.
template <class M, int MINB, bool IS_NUTS>
__global__ void __launch_bounds__(kChainBlock, MINB)
k_chain_run_dyn(SmallData md, ChainState st, DrawBuffers out, RunParams rp, const int *__restrict__ perm,
                int n_groups, int n_sub, long long sub_len, int *counter, int *done /* [n_groups] */,
                int use_ck_sm) {
  __shared__ int s_task;
  extern __shared__ double ck_smem[];
  for (;;) {
    if (threadIdx.x == 0) s_task = atomicAdd(counter, 1);
    __syncthreads();
    const int task = s_task;
    __syncthreads();
    if (task >= n_groups * n_sub) return;
    const int sub = task / n_groups, grp = task - sub * n_groups;
    if (sub > 0) {
      if (threadIdx.x == 0) {
        while (atomicAdd(done + grp, 0) < sub) __nanosleep(200);
        __threadfence();
      }
      __syncthreads();
    }
    const long long t0 = (long long)sub * sub_len;
    const long long t1 = (t0 + sub_len < rp.n_transitions) ? t0 + sub_len : rp.n_transitions;
    chain_run_range<M, IS_NUTS>(md, st, out, rp, perm, grp * kChainBlock + threadIdx.x, t0, t1,
                                (IS_NUTS && use_ck_sm) ? ck_smem + threadIdx.x : nullptr);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicExch(done + grp, sub + 1);
  }
}

// ------------------------------------------------------------------------------------------
// init kernel: initial parameters, phase point, find_good_stepsize, adaptor state
// ------------------------------------------------------------------------------------------
template <class M, class MT>
__device__ __forceinline__ double find_good_stepsize_d(const SmallData &md, const RunParams &rp,
                                                       uint32_t gchain, const double (&th)[M::D],
                                                       const double (&g)[M::D], double lp,
                                                       const MT &minv,
                                                       long long &ngrad) {
  constexpr int D = M::D;
  const double a_min = 0.25, a_cross = 0.5, a_max = 0.75, dd = 2.0;
  double eps = 0.1, eps_p = 0.1;
  double r0[D];
  draw_momentum_d<D>(rp.seed, gchain, 0u, SLOT_EPS_MOMENTUM, minv, r0);
  const double H = energy_d<D>(lp, g, r0, minv);
  double zt[D], zr[D], zg[D], zlp, zH;
  auto trial = [&](double e) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      zt[d] = th[d];
      zr[d] = r0[d];
      zg[d] = g[d];
    }
    leapfrog_d<M>(md, zt, zr, zg, minv, e, zlp, zH);
    ++ngrad;
    return H - zH;
  };
  double dH = trial(eps);
  const double lc = log(a_cross);
  const int direction = dH > lc ? 1 : -1;
  for (int i = 0; i < 100; ++i) {
    eps_p = direction == 1 ? dd * eps : 1.0 / dd * eps;
    dH = trial(eps);
    if (direction == 1 && !(dH > lc)) break;
    else if (direction == -1 && !(dH < lc)) break;
    else eps = eps_p;
  }
  if (eps > eps_p) {
    const double t = eps;
    eps = eps_p;
    eps_p = t;
  }
  for (int i = 0; i < 100; ++i) {
    const double mid = 0.5 * (eps + eps_p);
    dH = trial(mid);
    const double ea = exp(dH);
    if (ea > a_max) eps = mid;
    else if (ea < a_min) eps_p = mid;
    else {
      eps = mid;
      break;
    }
  }
  return eps;
}

template <class M>
__global__ void __launch_bounds__(kChainBlock) k_chain_init(SmallData md, ChainState st,
                                                           RunParams rp, const double *theta0,
                                                           double init_eps) {
  constexpr int D = M::D;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int C = st.C;
  if (c >= C) return;
  const uint32_t gchain = (uint32_t)(rp.chain_offset + c);
  double th[D], g[D], minv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) minv[d] = 1.0;  // DiagEuclideanMetric(D)
  double lp = 0.0;
  long long ngrad = 0;
  bool ok = false;
  int tries = 0;
  constexpr int npair = (D + 1) / 2;
  if (theta0 != nullptr) {
#pragma unroll
    for (int d = 0; d < D; ++d) th[d] = theta0[(size_t)c * D + d];
    lp = M::lpgrad(md, th, g);
    ++ngrad;
    ok = isfinite(lp);
#pragma unroll
    for (int d = 0; d < D; ++d) ok = ok && isfinite(g[d]);
    tries = 1;
  } else {
    for (int t = 0; t < 1000 && !ok; ++t) {  // max_attempts = 1000
#pragma unroll
      for (int p = 0; p < npair; ++p) {
        double u0, u1;
        uniform2(rp.seed, gchain, 0u, SLOT_INIT, (uint32_t)(t * npair + p), u0, u1);
        th[2 * p] = 4.0 * u0 - 2.0;  // InitFromUniform: U(-2, 2) in linked space
        if (2 * p + 1 < D) th[2 * p + 1] = 4.0 * u1 - 2.0;
      }
      lp = M::lpgrad(md, th, g);
      ++ngrad;
      ok = isfinite(lp);
#pragma unroll
      for (int d = 0; d < D; ++d) ok = ok && isfinite(g[d]);
      tries = t + 1;
    }
  }
  st.init_tries[c] = tries;
  st.status[c] = ok ? 0 : 1;
  double eps = init_eps;
  // iszero(spl.eps) => AHMC.find_good_stepsize for every Hamiltonian sampler, HMC included
  // (src/mcmc/hmc.jl:189-194)
  if (ok && init_eps == 0.0) eps = find_good_stepsize_d<M>(md, rp, gchain, th, g, lp, minv, ngrad);
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const size_t o = (size_t)d * C + c;
    st.theta[o] = th[d];
    st.grad[o] = g[d];
    st.minv[o] = minv[d];
    st.wf_mean[o] = 0.0;
    st.wf_m2[o] = 0.0;
  }
  st.lp[c] = lp;
  st.eps[c] = eps;
  st.da_eps[c] = eps;
  st.da_mu[c] = log(10.0 * eps);  // StepSizeAdaptor(delta, eps)
  st.da_xbar[c] = 0.0;
  st.da_hbar[c] = 0.0;
  st.da_m[c] = 0;
  st.wf_n[c] = 0;
  st.iter[c] = 0;
  st.n_grad[c] = ngrad;
}

// ------------------------------------------------------------------------------------------
// DenseEuclideanMetric variants of the run / init kernels (TNUTS_METRIC_DENSE).  Same loop as
// chain_run_range -- kept apart on purpose: the diagonal kernel is the headline kernel and sits at
// 255 registers -- with the mass-matrix part replaced: Welford covariance in st.wf_cov [D*D][C],
// M^-1 <- (n / (n + 5)) cov + 1e-3 (5 / (n + 5)) I at the window ends, new Cholesky factor.
// ------------------------------------------------------------------------------------------
template <class M, int MINB, bool IS_NUTS>
__global__ void __launch_bounds__(kChainBlock, MINB) k_chain_run_dense(SmallData md, ChainState st,
                                                                      DrawBuffers out, RunParams rp) {
  constexpr int D = M::D;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int C = st.C;
  if (c >= C || st.status[c] != 0) return;
  const uint32_t gchain = (uint32_t)(rp.chain_offset + c);

  double th[D], g[D];
  DenseMetric<D> met;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    th[d] = st.theta[(size_t)d * C + c];
    g[d] = st.grad[(size_t)d * C + c];
  }
  for (int i = 0; i < D * D; ++i) {
    met.minv[i] = st.minv_dense[(size_t)i * C + c];
    met.U[i] = st.chol[(size_t)i * C + c];
  }
  double lp = st.lp[c], eps = st.eps[c];
  long long iter = st.iter[c], ngrad = st.n_grad[c];
  double da_mu = st.da_mu[c], da_xbar = st.da_xbar[c], da_hbar = st.da_hbar[c], da_eps = st.da_eps[c];
  long long da_m = st.da_m[c], wf_n = st.wf_n[c];

  for (long long t = 1; t <= rp.n_transitions; ++t) {
    const uint32_t it = (uint32_t)(iter + 1);
    double stats[TNUTS_NSTAT];
    if (IS_NUTS) nuts_transition_d<M>(md, rp, gchain, it, th, g, lp, met, eps, ngrad, stats, nullptr);
    else hmc_transition_d<M>(md, rp, gchain, it, th, g, lp, met, eps, ngrad, stats);
    iter += 1;

    if (rp.algorithm != TNUTS_ALG_HMC && iter <= rp.n_adapts) {
      {  // Nesterov dual averaging: gamma = 0.05, t0 = 10, kappa = 0.75
        const double alpha = fmin(stats[ST_ACC], 1.0);
        const long long mm = da_m + 1;
        const double eta_h = 1.0 / ((double)mm + 10.0);
        const double hbar = (1.0 - eta_h) * da_hbar + eta_h * (rp.delta - alpha);
        const double x = da_mu - hbar * sqrt((double)mm) / 0.05;
        const double eta_x = pow((double)mm, -0.75);
        const double xbar = (1.0 - eta_x) * da_xbar + eta_x * x;
        const double e = exp(x);
        if (isfinite(e)) {  // a non-finite step size reverts m, x_bar, H_bar and eps (AdvancedHMC adapt_stepsize!)
          da_m = mm;
          da_hbar = hbar;
          da_xbar = xbar;
          da_eps = e;
        }
      }
      bool wend = false;
      for (int k = 0; k < rp.n_splits; ++k) wend = wend || (rp.splits[k] == iter);
      if (iter >= rp.window_start && iter <= rp.window_end) {
        wf_n += 1;
        const double n = (double)wf_n;
        double delta[D], dm[D];   // s - mu_old, s - mu_new
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const size_t o = (size_t)d * C + c;
          double mean = st.wf_mean[o];
          delta[d] = th[d] - mean;
          mean += delta[d] / n;
          dm[d] = th[d] - mean;
          st.wf_mean[o] = wend ? 0.0 : mean;
        }
        const bool refresh = wend && wf_n >= 10;
        for (int a = 0; a < D; ++a)
          for (int b = 0; b < D; ++b) {
            const size_t o = (size_t)(a * D + b) * C + c;
            const double cov = st.wf_cov[o] + dm[a] * delta[b];
            if (refresh)
              met.minv[a * D + b] = n / ((n + 5.0) * (n - 1.0)) * cov + (a == b ? 1e-3 * (5.0 / (n + 5.0)) : 0.0);
            st.wf_cov[o] = wend ? 0.0 : cov;
          }
        if (refresh) chol_upper_d<D>(met);
      }
      if (wend) {
        da_mu = log(10.0 * da_eps);
        da_m = 0;
        da_xbar = 0.0;
        da_hbar = 0.0;
        wf_n = 0;
      }
      if (iter == rp.n_adapts) da_eps = exp(da_xbar);
      eps = da_eps;
    }

    if (rp.first_keep >= 1 && t >= rp.first_keep && (t - rp.first_keep) % rp.thin == 0) {
      const long long k = (t - rp.first_keep) / rp.thin;
      constexpr int P = M::P;   // user-space parameters (a K-simplex: K columns for K - 1 linked dims)
      double p[P], prior, lik;
      M::constrain(md, th, p, prior, lik);
      double *row = out.draws + (size_t)k * out.ncol * C + c;
#pragma unroll
      for (int d = 0; d < P; ++d) row[(size_t)d * C] = p[d];
      row[(size_t)(P + 0) * C] = prior;
      row[(size_t)(P + 1) * C] = lik;
      row[(size_t)(P + 2) * C] = prior + lik;
#pragma unroll
      for (int s = 0; s < TNUTS_NSTAT; ++s) row[(size_t)(P + 3 + s) * C] = stats[s];
      if (out.theta != nullptr) {
#pragma unroll
        for (int d = 0; d < D; ++d) out.theta[((size_t)k * D + d) * C + c] = th[d];
      }
    }
  }

#pragma unroll
  for (int d = 0; d < D; ++d) {
    st.theta[(size_t)d * C + c] = th[d];
    st.grad[(size_t)d * C + c] = g[d];
    st.minv[(size_t)d * C + c] = met.minv[d * D + d];   // tnuts_get_inv_metric reports the diagonal
  }
  for (int i = 0; i < D * D; ++i) {
    st.minv_dense[(size_t)i * C + c] = met.minv[i];
    st.chol[(size_t)i * C + c] = met.U[i];
  }
  st.lp[c] = lp;
  st.eps[c] = eps;
  st.iter[c] = iter;
  st.n_grad[c] = ngrad;
  st.da_mu[c] = da_mu;
  st.da_xbar[c] = da_xbar;
  st.da_hbar[c] = da_hbar;
  st.da_eps[c] = da_eps;
  st.da_m[c] = da_m;
  st.wf_n[c] = wf_n;
}

// after k_chain_init: identity dense metric, and the step-size search repeated with it (identical
// to the diagonal search while M^-1 = I, so only the arrays need initialising)
__global__ void k_chain_init_dense(ChainState st) {
  const int C = st.C, D = st.D;
  const long long tot = (long long)D * D * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot;
       i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i / C);
    const double v = (e / D == e % D) ? 1.0 : 0.0;
    st.minv_dense[i] = v;
    st.chol[i] = v;
    st.wf_cov[i] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------
// point-wise helpers (LogDensityProblems interface, constrain, leapfrog trajectories)
// ------------------------------------------------------------------------------------------
template <class M>
__global__ void k_point_lpgrad(SmallData md, const double *theta, long long n, double *lp,
                               double *grad) {
  constexpr int D = M::D;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double th[D], g[D];
#pragma unroll
  for (int d = 0; d < D; ++d) th[d] = theta[i * D + d];
  lp[i] = M::lpgrad(md, th, g);
  if (grad != nullptr) {
#pragma unroll
    for (int d = 0; d < D; ++d) grad[i * D + d] = g[d];
  }
}

template <class M>
__global__ void k_point_constrain(SmallData md, const double *theta, long long n, double *params,
                                  double *logp) {
  constexpr int D = M::D;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int P = M::P;
  double th[D], p[P], prior, lik;
#pragma unroll
  for (int d = 0; d < D; ++d) th[d] = theta[i * D + d];
  M::constrain(md, th, p, prior, lik);
#pragma unroll
  for (int d = 0; d < P; ++d) params[i * P + d] = p[d];
  logp[i * 3 + 0] = prior;
  logp[i * 3 + 1] = lik;
  logp[i * 3 + 2] = prior + lik;
}

template <class M>
__global__ void k_point_leapfrog(SmallData md, const double *theta, const double *r0,
                                 const double *minv_in, double eps, int L, long long n,
                                 double *traj_theta, double *traj_r, double *traj_H) {
  constexpr int D = M::D;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double th[D], r[D], g[D], minv[D], lp, H;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    th[d] = theta[i * D + d];
    r[d] = r0[i * D + d];
    minv[d] = minv_in[d];
  }
  lp = M::lpgrad(md, th, g);
  H = energy_d<D>(lp, g, r, minv);
  for (int l = 0; l <= L; ++l) {
#pragma unroll
    for (int d = 0; d < D; ++d) {
      traj_theta[(i * (L + 1) + l) * D + d] = th[d];
      traj_r[(i * (L + 1) + l) * D + d] = r[d];
    }
    traj_H[i * (L + 1) + l] = H;
    if (l < L) leapfrog_d<M>(md, th, r, g, minv, eps, lp, H);
  }
}

}  // namespace tnuts
