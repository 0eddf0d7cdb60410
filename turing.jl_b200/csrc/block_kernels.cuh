// block_kernels.cuh -- "one CTA per chain" execution strategy.
//
// For families whose gradient is chain-local (no pass over shared data rows) but whose
// dimension is far too large for one thread -- the hierarchical regression of
// BASELINE.json configs[3], D = 2005 -- a whole chain lives inside ONE CTA for the whole run:
// thread t owns the dimensions d = t, t + 256, ... (KD per thread) in registers, reductions
// over d (kinetic energy, U-turn dot products, the model's sufficient sums) are block
// reductions, and the complete NUTS state machine of chain_kernels.cuh (same algorithm, same
// Philox addressing) runs without ever returning to the host.  Chains with long trees or a
// small step size therefore cost only their own SM slot: nobody waits for them.
//
// Kept draws are written chain-major ([C][keep][ncol], coalesced per CTA) and transposed to
// the engine's [keep][ncol][C] layout by k_block_transpose_draws at the end of the run.
#pragma once
#include "lockstep_kernels.cuh"

namespace tnuts {

#ifndef TNUTS_BLOCK_THREADS
#define TNUTS_BLOCK_THREADS 256
#endif
constexpr int kBlockThreads = TNUTS_BLOCK_THREADS;
constexpr int kBlockWarps = kBlockThreads / 32;
constexpr int kBlockCk = 12;  // checkpoint levels (max_depth <= 13)

// block-wide sum of NR values; result valid in every thread.  red: [NR][kBlockWarps] doubles
template <int NR>
__device__ __forceinline__ void block_reduce(double (&v)[NR], double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[r] += __shfl_xor_sync(0xffffffffu, v[r], o);
  }
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < NR; ++r) red[r * kBlockWarps + warp] = v[r];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kBlockWarps; ++w) s += red[r * kBlockWarps + w];
    v[r] = s;
  }
}

// runtime-count variant for the U-turn dot products (nv <= 2 + 2 * kBlockCk)
__device__ __forceinline__ void block_reduce_n(double *v, int nv, double *red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = 0; r < nv; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[r] += __shfl_xor_sync(0xffffffffu, v[r], o);
  }
  __syncthreads();
  if (lane == 0)
    for (int r = 0; r < nv; ++r) red[r * kBlockWarps + warp] = v[r];
  __syncthreads();
  for (int r = 0; r < nv; ++r) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kBlockWarps; ++w) s += red[r * kBlockWarps + w];
    v[r] = s;
  }
}

struct HierStatsDev {
  const double *n, *xbar, *ybar, *sxx, *sxy, *syy;
};

// ---- model: hierarchical linear regression from centred per-group sufficient statistics ----
// (same arithmetic as k_hier_tile in lpgrad_kernels.cuh)
template <int KD>
struct HierBlock {
  __device__ static double lpgrad(const ModelDev &m, const HierStatsDev &hs, double *sth /* smem [D] */,
                                  double *red, const double (&th)[KD], double (&g)[KD]) {
    const int D = m.D, G = m.G, tid = threadIdx.x;
    __syncthreads();  // previous readers of sth are done
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      const int d = tid + k * kBlockThreads;
      if (d < D) sth[d] = th[k];
    }
    __syncthreads();
    const double mua = sth[0], mub = sth[1], la = sth[2], lb = sth[3], ly = sth[4];
    const double isa2 = exp(-2.0 * la), isb2 = exp(-2.0 * lb), isy2 = exp(-2.0 * ly);
    double acc[5] = {0, 0, 0, 0, 0};  // qa, qb, da, db, ssr
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      const int d = tid + k * kBlockThreads;
      g[k] = 0.0;
      if (d >= 5 && d < D) {
        const bool is_a = d < 5 + G;
        const int gi = is_a ? d - 5 : d - 5 - G;
        const double a = is_a ? th[k] : sth[5 + gi];
        const double b = is_a ? sth[5 + G + gi] : th[k];
        const double ng = hs.n[gi], xb = hs.xbar[gi];
        const double mres = hs.ybar[gi] - a - b * xb;
        const double sres = ng * mres;
        if (is_a) {
          const double ea = a - mua;
          const double sxy = hs.sxy[gi], sxx = hs.sxx[gi];
          acc[0] += ea * ea;
          acc[2] += ea;
          acc[4] += hs.syy[gi] - 2.0 * b * sxy + b * b * sxx + sres * mres;
          g[k] = -ea * isa2 + sres * isy2;
        } else {
          const double eb = b - mub;
          const double sresx = (hs.sxy[gi] - b * hs.sxx[gi]) + xb * sres;
          acc[1] += eb * eb;
          acc[3] += eb;
          g[k] = -eb * isb2 + sresx * isy2;
        }
      }
    }
    block_reduce<5>(acc, red);
    const double sdm = m.hyper[0], s0 = m.hyper[1];
    const double sa2 = exp(2.0 * la), sb2 = exp(2.0 * lb), sy2 = exp(2.0 * ly);
    const double nn = (double)m.N, Gd = (double)G;
    double lp = 2.0 * (-0.5 * kLog2Pi - log(sdm)) - 0.5 * (mua * mua + mub * mub) / (sdm * sdm);
    lp += 3.0 * (kLog2 - 0.5 * kLog2Pi - log(s0)) - 0.5 * (sa2 + sb2 + sy2) / (s0 * s0) + la + lb + ly;
    lp += -Gd * kLog2Pi - Gd * la - Gd * lb - 0.5 * acc[0] * isa2 - 0.5 * acc[1] * isb2;
    lp += -0.5 * nn * kLog2Pi - nn * ly - 0.5 * acc[4] * isy2;
    if (tid < 5) {  // hyper-parameters live in slot k = 0 of threads 0..4
      double gh;
      if (tid == 0) gh = -mua / (sdm * sdm) + acc[2] * isa2;
      else if (tid == 1) gh = -mub / (sdm * sdm) + acc[3] * isb2;
      else if (tid == 2) gh = -sa2 / (s0 * s0) + 1.0 - Gd + acc[0] * isa2;
      else if (tid == 3) gh = -sb2 / (s0 * s0) + 1.0 - Gd + acc[1] * isb2;
      else gh = -sy2 / (s0 * s0) + 1.0 - nn + acc[4] * isy2;
      g[0] = gh;
    }
    return lp;
  }
  // user-space output: constrained value of dimension d, prior sums
  __device__ static void user_logp(const ModelDev &m, double *sth, double *red, const double (&th)[KD],
                                   double lp, double &prior, double &lik) {
    const int D = m.D, G = m.G, tid = threadIdx.x;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      const int d = tid + k * kBlockThreads;
      if (d < D) sth[d] = th[k];
    }
    __syncthreads();
    const double mua = sth[0], mub = sth[1], la = sth[2], lb = sth[3], ly = sth[4];
    double q[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      const int d = tid + k * kBlockThreads;
      if (d >= 5 && d < D) {
        if (d < 5 + G) {
          const double ea = th[k] - mua;
          q[0] += ea * ea;
        } else {
          const double eb = th[k] - mub;
          q[1] += eb * eb;
        }
      }
    }
    block_reduce<2>(q, red);
    const double sdm = m.hyper[0], s0 = m.hyper[1];
    const double sa = exp(la), sb = exp(lb), sy = exp(ly);
    prior = 2.0 * (-0.5 * kLog2Pi - log(sdm)) - 0.5 * (mua * mua + mub * mub) / (sdm * sdm) +
            3.0 * (kLog2 - 0.5 * kLog2Pi - log(s0)) - 0.5 * (sa * sa + sb * sb + sy * sy) / (s0 * s0) -
            (double)G * kLog2Pi - G * la - G * lb - 0.5 * q[0] / (sa * sa) - 0.5 * q[1] / (sb * sb);
    lik = lp - prior - (la + lb + ly);
  }
};

// ---- model: iid standard normal (any D); exercises the strategy at arbitrary dimension ----
template <int KD>
struct StdNormalBlock {
  __device__ static double lpgrad(const ModelDev &m, const HierStatsDev &, double *, double *red,
                                  const double (&th)[KD], double (&g)[KD]) {
    double q[1] = {0.0};
#pragma unroll
    for (int k = 0; k < KD; ++k) {
      const int d = threadIdx.x + k * kBlockThreads;
      g[k] = 0.0;
      if (d < m.D) {
        q[0] += th[k] * th[k];
        g[k] = -th[k];
      }
    }
    block_reduce<1>(q, red);
    return -0.5 * m.D * kLog2Pi - 0.5 * q[0];
  }
  __device__ static void user_logp(const ModelDev &, double *, double *, const double (&)[KD], double lp,
                                   double &prior, double &lik) {
    prior = lp;
    lik = 0.0;
  }
};

template <int KD>
struct BlockVec {
  double v[KD];
};

// energy pieces: sum minv r^2 and count of non-finite gradient entries
template <int KD>
__device__ __forceinline__ void kin_parts(const double (&r)[KD], const double (&g)[KD],
                                          const double (&minv)[KD], int D, double &k2, double &bad) {
  k2 = 0.0;
  bad = 0.0;
#pragma unroll
  for (int k = 0; k < KD; ++k) {
    const int d = threadIdx.x + k * kBlockThreads;
    if (d < D) {
      k2 += minv[k] * r[k] * r[k];
      bad += isfinite(g[k]) ? 0.0 : 1.0;
    }
  }
}

template <int KD>
__device__ __forceinline__ void draw_momentum_b(const RunParams &rp, uint32_t gchain, uint32_t it,
                                                uint32_t slot, const double (&minv)[KD], int D,
                                                double (&r)[KD]) {
#pragma unroll
  for (int k = 0; k < KD; ++k) {
    const int d = threadIdx.x + k * kBlockThreads;
    r[k] = 0.0;
    if (d < D) {
      double z0, z1;
      normal2(rp.seed, gchain, it, slot, (uint32_t)(d >> 1), z0, z1);
      r[k] = ((d & 1) ? z1 : z0) / sqrt(minv[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// run kernel: CTA c = chain c, all transitions of this call
// ------------------------------------------------------------------------------------------
template <class M, int KD>
__global__ void __launch_bounds__(kBlockThreads)
k_block_run(ModelDev m, HierStatsDev hs, ChainState st, RunParams rp, double *draws_cm /* [C][keep][ncol] */,
            long long keep, int ncol, double *theta_cm /* [C][keep][D] or null */) {
  extern __shared__ double smem[];
  double *sth = smem;             // [D]
  double *red = smem + m.D;       // [(2 + 2*kBlockCk)][kBlockWarps]
  const int c = blockIdx.x, C = st.C, D = m.D, tid = threadIdx.x;
  if (st.status[c] != 0) return;
  const uint32_t gchain = (uint32_t)(rp.chain_offset + c);
  double th[KD], g[KD], minv[KD];
#pragma unroll
  for (int k = 0; k < KD; ++k) {
    const int d = tid + k * kBlockThreads;
    th[k] = g[k] = 0.0;
    minv[k] = 1.0;
    if (d < D) {
      th[k] = st.theta[(size_t)d * C + c];
      g[k] = st.grad[(size_t)d * C + c];
      minv[k] = st.minv[(size_t)d * C + c];
    }
  }
  double lp = st.lp[c], eps = st.eps[c];
  long long iter = st.iter[c], ngrad = 0;
  double da_mu = st.da_mu[c], da_xbar = st.da_xbar[c], da_hbar = st.da_hbar[c], da_eps = st.da_eps[c];
  long long da_m = st.da_m[c], wf_n = st.wf_n[c];

  for (long long tt = 1; tt <= rp.n_transitions; ++tt) {
    const uint32_t it = (uint32_t)(iter + 1);
    // ---- momentum refresh, H0 ----
    double r[KD];
    draw_momentum_b<KD>(rp, gchain, it, SLOT_MOMENTUM, minv, D, r);
    double e0[2];
    kin_parts<KD>(r, g, minv, D, e0[0], e0[1]);
    block_reduce<2>(e0, red);
    const double K0 = 0.5 * e0[0];
    const double H0 = (isfinite(lp) && e0[1] == 0.0 && isfinite(K0)) ? (-lp + K0) : CUDART_INF;
    double stats[TNUTS_NSTAT];
    if (rp.algorithm != TNUTS_ALG_NUTS) {
      // ---- HMC / HMCDA: static trajectory + Metropolis accept (src/mcmc/hmc.jl:425-434) ----
      int L = rp.n_leapfrog;
      if (rp.algorithm == TNUTS_ALG_HMCDA) {
        L = (int)floor(rp.lambda / eps);
        if (L < 1) L = 1;
      }
      double zt[KD], zg[KD], zlp = lp, zH = H0;
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        zt[k] = th[k];
        zg[k] = g[k];
      }
      const double half = 0.5 * eps;
      int nsteps = 0;
      for (int l = 0; l < L; ++l) {
#pragma unroll
        for (int k = 0; k < KD; ++k) {
          r[k] = r[k] + half * zg[k];
          zt[k] = zt[k] + eps * (minv[k] * r[k]);
        }
        zlp = M::lpgrad(m, hs, sth, red, zt, zg);
        ++ngrad;
#pragma unroll
        for (int k = 0; k < KD; ++k) r[k] = r[k] + half * zg[k];
        double ev[2];
        kin_parts<KD>(r, zg, minv, D, ev[0], ev[1]);
        block_reduce<2>(ev, red);
        const double K = 0.5 * ev[0];
        zH = (isfinite(zlp) && ev[1] == 0.0 && isfinite(K)) ? (-zlp + K) : CUDART_INF;
        ++nsteps;
        if (!isfinite(zH)) break;
      }
      const double dH = zH - H0;
      const double alpha = exp(fmin(0.0, -dH));
      const double u = uniform1(rp.seed, gchain, it, SLOT_MH, 0u);
      const bool accept = u < alpha;
      if (accept) {
#pragma unroll
        for (int k = 0; k < KD; ++k) {
          th[k] = zt[k];
          g[k] = zg[k];
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
    } else {
    // ---- tree ----
      double edge[2][3 * KD], elp[2];
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        edge[0][k] = edge[1][k] = th[k];
        edge[0][KD + k] = edge[1][KD + k] = r[k];
        edge[0][2 * KD + k] = edge[1][2 * KD + k] = g[k];
      }
      elp[0] = elp[1] = lp;
      double rho[KD], thC[KD], gC[KD];
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        rho[k] = r[k];
        thC[k] = th[k];
        gC[k] = g[k];
      }
      double sum_a = 0.0, dHmax = 0.0, lw = 0.0, lpC = lp, HC = H0;
      long long n_a = 0;
      double ck_r[kBlockCk][KD], ck_rho[kBlockCk][KD];
      int j = 0;
      bool term = false, numerr = false;
      while (!term && j < rp.max_depth) {
        const int v = (uniform1(rp.seed, gchain, it, SLOT_DIRECTION, (uint32_t)j) < 0.5) ? 0 : 1;
        const double se = v ? eps : -eps, half = 0.5 * se;
        double zt[KD], zr[KD], zg[KD], zlp = elp[v], zH = 0.0;
#pragma unroll
        for (int k = 0; k < KD; ++k) {
          zt[k] = edge[v][k];
          zr[k] = edge[v][KD + k];
          zg[k] = edge[v][2 * KD + k];
        }
        double rho_s[KD], thS[KD], gS[KD], lpS = 0.0, HS = 0.0;
#pragma unroll
        for (int k = 0; k < KD; ++k) rho_s[k] = 0.0;
        double lw_s = -CUDART_INF;
        bool term_s = false;
        const int nleaf = 1 << j;
        for (int n = 0; n < nleaf; ++n) {
          // leapfrog
#pragma unroll
          for (int k = 0; k < KD; ++k) {
            zr[k] = zr[k] + half * zg[k];
            zt[k] = zt[k] + se * (minv[k] * zr[k]);
          }
          zlp = M::lpgrad(m, hs, sth, red, zt, zg);
          ++ngrad;
#pragma unroll
          for (int k = 0; k < KD; ++k) zr[k] = zr[k] + half * zg[k];
          // one block reduction: kinetic energy, finiteness, and the sub-tree U-turn dots
          const int idx_max = __popc((unsigned)n >> 1);
          const bool odd = (n & 1) != 0;
          const int ntrail = __ffs(~(unsigned)n) - 1;
          const int idx_min = idx_max - ntrail + 1;
          const int nlev = odd ? (idx_max - idx_min + 1) : 0;
          double rv[2 + 2 * kBlockCk];
          kin_parts<KD>(zr, zg, minv, D, rv[0], rv[1]);
#pragma unroll
          for (int k = 0; k < KD; ++k) rho_s[k] += zr[k];
          for (int q = 0; q < nlev; ++q) {
            const int i = idx_max - q;
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int k = 0; k < KD; ++k) {
              const double cr = ck_r[i][k];
              const double sr = rho_s[k] - ck_rho[i][k] + cr;
              a += sr * (minv[k] * cr);
              b += sr * (minv[k] * zr[k]);
            }
            rv[2 + 2 * q] = a;
            rv[3 + 2 * q] = b;
          }
          block_reduce_n(rv, 2 + 2 * nlev, red);
          const double K = 0.5 * rv[0];
          const bool ok = isfinite(zlp) && rv[1] == 0.0 && isfinite(K);
          zH = ok ? (-zlp + K) : CUDART_INF;
          const double dH = zH - H0;
          const double a = exp(fmin(0.0, -dH));
          sum_a += a;
          ++n_a;
          if (fabs(dH) > fabs(dHmax)) dHmax = dH;
          const double lw_leaf = -dH;
          const double lw_new = ls_logaddexp(lw_s, lw_leaf);
          const double p = (lw_leaf == -CUDART_INF) ? 0.0 : exp(lw_leaf - lw_new);
          const double u = uniform1(rp.seed, gchain, it, SLOT_LEAF, (1u << j) + (uint32_t)n);
          if (u < p) {
#pragma unroll
            for (int k = 0; k < KD; ++k) {
              thS[k] = zt[k];
              gS[k] = zg[k];
            }
            lpS = zlp;
            HS = zH;
          }
          lw_s = lw_new;
          if (!(dH < rp.delta_max)) {
            numerr = true;
            term_s = true;
            break;
          }
          if (!odd) {
#pragma unroll
            for (int k = 0; k < KD; ++k) {
              ck_r[idx_max][k] = zr[k];
              ck_rho[idx_max][k] = rho_s[k];
            }
          } else {
            for (int q = 0; q < nlev; ++q)
              if ((rv[2 + 2 * q] <= 0.0) || (rv[3 + 2 * q] <= 0.0)) term_s = true;
            if (term_s) break;
          }
        }
#pragma unroll
        for (int k = 0; k < KD; ++k) {
          edge[v][k] = zt[k];
          edge[v][KD + k] = zr[k];
          edge[v][2 * KD + k] = zg[k];
        }
        elp[v] = zlp;
        if (!term_s) {
          const double u = uniform1(rp.seed, gchain, it, SLOT_MH, (uint32_t)j);
          ++j;
          const double dl = lw_s - lw;
          if (u < (dl < 0.0 ? exp(dl) : 1.0)) {
#pragma unroll
            for (int k = 0; k < KD; ++k) {
              thC[k] = thS[k];
              gC[k] = gS[k];
            }
            lpC = lpS;
            HC = HS;
          }
        }
        double d2[2] = {0.0, 0.0};
#pragma unroll
        for (int k = 0; k < KD; ++k) {
          rho[k] += rho_s[k];
          d2[0] += rho[k] * (minv[k] * edge[0][KD + k]);
          d2[1] += rho[k] * (minv[k] * edge[1][KD + k]);
        }
        block_reduce<2>(d2, red);
        lw = ls_logaddexp(lw, lw_s);
        term = term_s || (d2[0] <= 0.0) || (d2[1] <= 0.0);
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
      for (int k = 0; k < KD; ++k) {
        th[k] = thC[k];
        g[k] = gC[k];
      }
      lp = lpC;
    }
    iter += 1;

    // ---- AHMC.adapt! ----
    if (rp.algorithm != TNUTS_ALG_HMC && iter <= rp.n_adapts) {
      {
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
      for (int q = 0; q < rp.n_splits; ++q) wend = wend || (rp.splits[q] == iter);
      if (iter >= rp.window_start && iter <= rp.window_end) {
        wf_n += 1;
        const double nw = (double)wf_n;
#pragma unroll
        for (int k = 0; k < KD; ++k) {
          const int d = tid + k * kBlockThreads;
          if (d < D) {
            const size_t o = (size_t)d * C + c;
            double mean = st.wf_mean[o], m2 = st.wf_m2[o];
            const double delta = th[k] - mean;
            mean += delta / nw;
            m2 += delta * (th[k] - mean);
            if (wend) {
              if (wf_n >= 10 && !rp.unit_metric) minv[k] = nw / ((nw + 5.0) * (nw - 1.0)) * m2 + 1e-3 * (5.0 / (nw + 5.0));
              mean = 0.0;
              m2 = 0.0;
            }
            st.wf_mean[o] = mean;
            st.wf_m2[o] = m2;
          }
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

    // ---- kept draw (chain-major staging) ----
    if (rp.first_keep >= 1 && tt >= rp.first_keep && (tt - rp.first_keep) % rp.thin == 0) {
      const long long ks = (tt - rp.first_keep) / rp.thin;
      double prior, lik;
      M::user_logp(m, sth, red, th, lp, prior, lik);
      double *row = draws_cm + ((size_t)c * keep + ks) * ncol;
#pragma unroll
      for (int k = 0; k < KD; ++k) {
        const int d = tid + k * kBlockThreads;
        if (d < D) {
          row[d] = ls_is_positive(m, d) ? exp(th[k]) : th[k];
          if (theta_cm) theta_cm[((size_t)c * keep + ks) * D + d] = th[k];
        }
      }
      if (tid == 0) {
        row[D + 0] = prior;
        row[D + 1] = lik;
        row[D + 2] = prior + lik;
#pragma unroll
        for (int q = 0; q < TNUTS_NSTAT; ++q) row[D + 3 + q] = stats[q];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KD; ++k) {
    const int d = tid + k * kBlockThreads;
    if (d < D) {
      st.theta[(size_t)d * C + c] = th[k];
      st.grad[(size_t)d * C + c] = g[k];
      st.minv[(size_t)d * C + c] = minv[k];
    }
  }
  if (tid == 0) {
    st.lp[c] = lp;
    st.eps[c] = eps;
    st.iter[c] = iter;
    st.n_grad[c] += ngrad;
    st.da_mu[c] = da_mu;
    st.da_xbar[c] = da_xbar;
    st.da_hbar[c] = da_hbar;
    st.da_eps[c] = da_eps;
    st.da_m[c] = da_m;
    st.wf_n[c] = wf_n;
  }
}

// [C][keep][w] -> [keep][w][C]
__global__ void k_block_transpose_draws(const double *cm, double *out, long long keep, int w, int C) {
  __shared__ double tile[32][33];
  // grid.x: (draw, col) flattened rows, grid.y: chains
  const long long rows = keep * w;
  const long long r0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long r = r0 + threadIdx.x;
    if (c < C && r < rows) tile[i][threadIdx.x] = cm[(size_t)c * rows + r];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long r = r0 + i;
    const int c = c0 + threadIdx.x;
    if (c < C && r < rows) out[(size_t)r * C + c] = tile[threadIdx.x][i];
  }
}

}  // namespace tnuts
