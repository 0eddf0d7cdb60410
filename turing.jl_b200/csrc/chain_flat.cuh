// chain_flat.cuh -- flat-scheduled variant of the one-thread-per-chain NUTS kernel.
//
// k_chain_run (chain_kernels.cuh) nests the loops transition > doubling > leaf, so the 32
// chains of a warp re-converge at the end of every transition: a warp runs at the speed of its
// deepest tree (ncu: 12.2 of 32 lanes active).  Here every lane is its own state machine over
// ONE flat loop whose body is a single leapfrog:
//   * a lane whose doubling ends combines it and starts the next one right away;
//   * a lane whose tree ends WAITS; as soon as kFlatBatch lanes wait (or nobody runs) the warp
//     executes the transition boundary -- stats, adaptation, kept draw, momentum refresh, tree
//     init -- for all waiting lanes at once, so that this expensive code keeps a good lane
//     occupancy, and those lanes rejoin the leapfrog stream at their next transition.
// Lanes are therefore at different transitions; a lane only idles while it waits for the
// batch.  Arithmetic and random-number addressing per chain are identical to k_chain_run
// (results are bit-identical), only the interleaving across lanes changes.
#pragma once
#include "chain_kernels.cuh"

namespace tnuts {

constexpr int kFlatBatch = 12;  // lanes that must wait before the warp runs a transition boundary

template <class M, int MINB>
__global__ void __launch_bounds__(kChainBlock, MINB)
k_chain_run_flat(SmallData md, ChainState st, DrawBuffers out, RunParams rp,
                 const int *__restrict__ perm) {
  constexpr int D = M::D;
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  const int C = st.C;
  const bool live = slot < C;
  const int c = live ? (perm ? perm[slot] : slot) : 0;
  const bool ok = live && st.status[c] == 0;
  const uint32_t gchain = (uint32_t)(rp.chain_offset + c);

  // ---- chain state ----
  double th[D], g[D], minv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    th[d] = ok ? st.theta[(size_t)d * C + c] : 0.0;
    g[d] = ok ? st.grad[(size_t)d * C + c] : 0.0;
    minv[d] = ok ? st.minv[(size_t)d * C + c] : 1.0;
  }
  double lp = ok ? st.lp[c] : 0.0, eps = ok ? st.eps[c] : 0.1;
  long long iter = ok ? st.iter[c] : 0, ngrad = ok ? st.n_grad[c] : 0;
  double da_mu = 0, da_xbar = 0, da_hbar = 0, da_eps = 0;
  long long da_m = 0, wf_n = 0;
  if (ok) {
    da_mu = st.da_mu[c]; da_xbar = st.da_xbar[c]; da_hbar = st.da_hbar[c]; da_eps = st.da_eps[c];
    da_m = st.da_m[c]; wf_n = st.wf_n[c];
  }
  // ---- tree state ----
  double edge[2][3 * D], elp[2] = {0.0, 0.0};
  double rho[D], thC[D], gC[D], lpC = 0.0, HC = 0.0;
  double sum_a = 0.0, dHmax = 0.0, lw = 0.0, H0 = 0.0;
  long long n_a = 0;
  double ck_r[kMaxCk][D], ck_rho[kMaxCk][D];
  // ---- doubling state ----
  double zt[D], zr[D], zg[D], zlp = 0.0, rho_s[D], thS[D], gS[D], lpS = 0.0, HS = 0.0;
  double lw_s = 0.0, se = 0.0;
  int v = 0, j = 0, n = 0;
  bool numerr = false;
  uint32_t it = 0;

  int phase = ok ? 0 : 2;  // 0 wait at a transition boundary, 1 running, 2 done
  bool have_finished = false;
  long long t = 0;

  while (true) {
    const unsigned run_mask = __ballot_sync(0xffffffffu, phase == 1);
    const unsigned wait_mask = __ballot_sync(0xffffffffu, phase == 0);
    if (run_mask == 0u && wait_mask == 0u) break;
    if (wait_mask != 0u && (run_mask == 0u || __popc(wait_mask) >= rp.flat_batch)) {
      // ================= transition boundary for every waiting lane =================
      if (phase == 0) {
        if (have_finished) {
          double stats[TNUTS_NSTAT];
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
          iter += 1;
          // ---- AHMC.adapt! ----
          if (iter <= rp.n_adapts) {
            {
              const double alpha = fmin(stats[ST_ACC], 1.0);
              const long long mm = da_m + 1;
              const double eta_h = 1.0 / ((double)mm + 10.0);
              const double hbar = (1.0 - eta_h) * da_hbar + eta_h * (rp.delta - alpha);
              const double x = da_mu - hbar * sqrt((double)mm) / 0.05;
              const double eta_x = pow((double)mm, -0.75);
              const double xbar = (1.0 - eta_x) * da_xbar + eta_x * x;
              const double e = exp(x);
              da_m = mm;
              if (isfinite(e)) {
                da_hbar = hbar;
                da_xbar = xbar;
                da_eps = e;
              }
            }
            bool wend = false;
            for (int k = 0; k < rp.n_splits; ++k) wend = wend || (rp.splits[k] == iter);
            if (iter >= rp.window_start && iter <= rp.window_end) {
              wf_n += 1;
              const double nw = (double)wf_n;
#pragma unroll
              for (int d = 0; d < D; ++d) {
                const size_t o = (size_t)d * C + c;
                double mean = st.wf_mean[o], m2 = st.wf_m2[o];
                const double delta = th[d] - mean;
                mean += delta / nw;
                m2 += delta * (th[d] - mean);
                if (wend) {
                  if (wf_n >= 10 && !rp.unit_metric) minv[d] = nw / ((nw + 5.0) * (nw - 1.0)) * m2 + 1e-3 * (5.0 / (nw + 5.0));
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
          // ---- kept draw ----
          if (rp.first_keep >= 1 && t >= rp.first_keep && (t - rp.first_keep) % rp.thin == 0) {
            const long long k = (t - rp.first_keep) / rp.thin;
            double p[D], prior, lik;
            M::constrain(md, th, p, prior, lik);
            double *row = out.draws + (size_t)k * out.ncol * C + c;
#pragma unroll
            for (int d = 0; d < D; ++d) row[(size_t)d * C] = p[d];
            row[(size_t)(D + 0) * C] = prior;
            row[(size_t)(D + 1) * C] = lik;
            row[(size_t)(D + 2) * C] = prior + lik;
#pragma unroll
            for (int s = 0; s < TNUTS_NSTAT; ++s) row[(size_t)(D + 3 + s) * C] = stats[s];
            if (out.theta != nullptr) {
#pragma unroll
              for (int d = 0; d < D; ++d) out.theta[((size_t)k * D + d) * C + c] = th[d];
            }
          }
          have_finished = false;
        }
        if (t >= rp.n_transitions) {
          phase = 2;
        } else {
          // ---- start the next transition: momentum refresh, H0, tree init, doubling 0 ----
          ++t;
          it = (uint32_t)(iter + 1);
          double r[D];
          draw_momentum_d<D>(rp.seed, gchain, it, SLOT_MOMENTUM, minv, r);
          H0 = energy_d<D>(lp, g, r, minv);
#pragma unroll
          for (int d = 0; d < D; ++d) {
            edge[0][d] = edge[1][d] = th[d];
            edge[0][D + d] = edge[1][D + d] = r[d];
            edge[0][2 * D + d] = edge[1][2 * D + d] = g[d];
            rho[d] = r[d];
            thC[d] = th[d];
            gC[d] = g[d];
            zt[d] = th[d];
            zr[d] = r[d];
            zg[d] = g[d];
            rho_s[d] = 0.0;
          }
          elp[0] = elp[1] = lp;
          sum_a = 0.0;
          n_a = 0;
          dHmax = 0.0;
          lw = 0.0;
          lpC = lp;
          HC = H0;
          numerr = false;
          j = 0;
          n = 0;
          v = (uniform1(rp.seed, gchain, it, SLOT_DIRECTION, 0u) < 0.5) ? 0 : 1;
          se = v ? eps : -eps;
          zlp = lp;
          lw_s = -CUDART_INF;
          phase = 1;
        }
      }
      continue;
    }
    // ================= one leapfrog for every running lane =================
    if (phase == 1) {
      double zH;
      leapfrog_d<M>(md, zt, zr, zg, minv, se, zlp, zH);
      ++ngrad;
      const double dH = zH - H0;
      const double a = exp(fmin(0.0, -dH));
      sum_a += a;
      ++n_a;
      if (fabs(dH) > fabs(dHmax)) dHmax = dH;
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
      bool term_s = false;
      if (!(dH < rp.delta_max)) {
        numerr = true;
        term_s = true;
      } else {
        const int idx_max = __popc((unsigned)n >> 1);
        if ((n & 1) == 0) {
#pragma unroll
          for (int d = 0; d < D; ++d) {
            ck_r[idx_max][d] = zr[d];
            ck_rho[idx_max][d] = rho_s[d];
          }
        } else {
          const int ntrail = __ffs(~(unsigned)n) - 1;
          const int idx_min = idx_max - ntrail + 1;
          for (int i = idx_max; i >= idx_min; --i) {
            double sr[D];
#pragma unroll
            for (int d = 0; d < D; ++d) sr[d] = rho_s[d] - ck_rho[i][d] + ck_r[i][d];
            if (uturn_d<D>(sr, ck_r[i], zr, minv)) {
              term_s = true;
              break;
            }
          }
        }
      }
      if (term_s || n + 1 == (1 << j)) {
        // ---- end of the doubling: combine, biased progressive sampling, tree U-turn ----
#pragma unroll
        for (int d = 0; d < D; ++d) {
          edge[v][d] = zt[d];
          edge[v][D + d] = zr[d];
          edge[v][2 * D + d] = zg[d];
        }
        elp[v] = zlp;
        if (!term_s) {
          const double um = uniform1(rp.seed, gchain, it, SLOT_MH, (uint32_t)j);
          ++j;
          const double dl = lw_s - lw;
          if (um < (dl < 0.0 ? exp(dl) : 1.0)) {
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
        const bool term = term_s || uturn_d<D>(rho, &edge[0][D], &edge[1][D], minv);
        if (!term && j < rp.max_depth) {
          // ---- next doubling ----
          v = (uniform1(rp.seed, gchain, it, SLOT_DIRECTION, (uint32_t)j) < 0.5) ? 0 : 1;
          se = v ? eps : -eps;
#pragma unroll
          for (int d = 0; d < D; ++d) {
            zt[d] = edge[v][d];
            zr[d] = edge[v][D + d];
            zg[d] = edge[v][2 * D + d];
            rho_s[d] = 0.0;
          }
          zlp = elp[v];
          lw_s = -CUDART_INF;
          n = 0;
        } else {
          phase = 0;
          have_finished = true;
        }
      } else {
        ++n;
      }
    }
  }

  if (ok) {
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
}

}  // namespace tnuts
