// dense_pump.cuh -- DenseEuclideanMetric on the pump (lock-step) strategy: the kernels around the
// whitened state machine (lockstep_kernels.cuh, ls_pump_body<TM, DENSE = true>).
//
// metricT = DenseEuclideanMetric (src/mcmc/hmc.jl:377-386) keeps, per chain, M^-1 (st.minv_dense), the
// upper Cholesky factor U of M^-1 = U'U (st.chol) and the Welford covariance accumulator (st.wf_cov),
// [D*D][C] each, element (i, j) of chain c at (i * D + j) * C + c.  The state machine runs the
// unit-metric dynamics on phi = U^-T theta; per pump
//     k_dense_post   after the state machine: Welford covariance push of the chains that ended a
//                    transition inside an adaptation window; at a window end the new metric
//                    (n / ((n + 5)(n - 1)) cov + 1e-3 * 5 / (n + 5) I, AdvancedHMC's regulariser), its
//                    Cholesky factor, phi and grad_phi re-expressed in the new factor, the paused chain
//                    released; then theta = U'phi of every chain's drifted position (ls.wT) for the
//                    family's gradient kernel
//     (gradient kernel: ls.wT -> ls.wG, zlp)
//     k_dense_grad   grad_phi = U grad_theta (ls.zg) for the chains that wait for a leaf
// Same tile mapping as the other lock-step kernels: 8 chains x 32 d-lanes per CTA.
#pragma once
#include "lockstep_kernels.cuh"

namespace tnuts {

constexpr int kDenseMaxD = 256;              // lanes keep D / 32 elements of a vector in registers
constexpr int kDenseOwn = kDenseMaxD / kTileD;

// theta_d = sum_{e <= d} U[e][d] phi_e for this lane's d
// (no __restrict__ / read-only loads here: k_dense_post reads the factor it has just written)
__device__ __forceinline__ void dense_ut_apply(const double *U, const double *phi, double *out, int D, int C,
                                               const Tile &t) {
  for (int d = t.dl; d < D; d += kTileD) {
    double a0 = 0.0, a1 = 0.0;
    int e = 0;
    for (; e + 1 <= d; e += 2) {
      a0 = fma(U[((size_t)e * D + d) * C + t.c], phi[(size_t)e * C + t.c], a0);
      a1 = fma(U[((size_t)(e + 1) * D + d) * C + t.c], phi[(size_t)(e + 1) * C + t.c], a1);
    }
    if (e <= d) a0 = fma(U[((size_t)e * D + d) * C + t.c], phi[(size_t)e * C + t.c], a0);
    out[(size_t)d * C + t.c] = a0 + a1;
  }
}
// out_e = sum_{d >= e} U[e][d] g_d for this lane's e
__device__ __forceinline__ void dense_u_apply(const double *U, const double *g, double *out, int D, int C,
                                              const Tile &t) {
  for (int e = t.dl; e < D; e += kTileD) {
    double a0 = 0.0, a1 = 0.0;
    int d = e;
    for (; d + 1 < D; d += 2) {
      a0 = fma(U[((size_t)e * D + d) * C + t.c], g[(size_t)d * C + t.c], a0);
      a1 = fma(U[((size_t)e * D + d + 1) * C + t.c], g[(size_t)(d + 1) * C + t.c], a1);
    }
    if (d < D) a0 = fma(U[((size_t)e * D + d) * C + t.c], g[(size_t)d * C + t.c], a0);
    out[(size_t)e * C + t.c] = a0 + a1;
  }
}

// identity metric at init; phi = theta
static __global__ void k_dense_init(ChainState st) {
  const size_t C = (size_t)st.C, D = (size_t)st.D;
  const size_t tot = D * D * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (size_t)gridDim.x * blockDim.x) {
    const size_t ij = i / C;
    const double v = (ij / D == ij % D) ? 1.0 : 0.0;
    st.minv_dense[i] = v;
    st.chol[i] = v;
    st.wf_cov[i] = 0.0;
    if (i < D * C) st.wf_m2[i] = st.theta[i];
  }
}

static __global__ void __launch_bounds__(kTileThreads)
k_dense_grad(ChainState st, LockstepState ls) {
  const Tile t;
  if (t.c >= st.C || ls.ts.phase[t.c] != 1) return;
  dense_u_apply(st.chol, ls.wG, ls.zg, st.D, st.C, t);
}

static __global__ void __launch_bounds__(kTileThreads)
k_dense_post(ChainState st, LockstepState ls) {
  __shared__ double bc[2][kTileC];
  const Tile t;
  const int C = st.C, D = st.D;
  const bool in = t.c < C;
  const int req = in ? ls.dn_req[t.c] : 0;
  double *const U = st.chol;
  double *const A = st.minv_dense;
  double *const cov = st.wf_cov;
  if (__syncthreads_or(req != 0)) {
    const double n = in ? (double)ls.dn_n[t.c] : 1.0;
    double *const sv = ls.wG;   // scratch column: last pump's gradient has been consumed
    // ---- Welford covariance push (oracle: welford_cov_push): delta = theta - mu, mu += delta / n,
    //      cov[i][j] += (theta_i - mu_new_i) delta_j
    const bool push = (req & 1) != 0;
    double own[kDenseOwn];
    if (push) {
      int k = 0;
      for (int d = t.dl; d < D; d += kTileD, ++k) {
        const size_t o = (size_t)d * C + t.c;
        const double th = st.theta[o], mu = st.wf_mean[o];
        const double delta = th - mu, mu2 = mu + delta / n;
        st.wf_mean[o] = mu2;
        sv[o] = delta;
        own[k] = th - mu2;
      }
    }
    __syncthreads();
    if (push) {
      int k = 0;
      for (int i = t.dl; i < D; i += kTileD, ++k) {
        const double a = own[k];
        for (int j = 0; j < D; ++j) {
          const size_t o = ((size_t)i * D + j) * C + t.c;
          cov[o] = fma(a, sv[(size_t)j * C + t.c], cov[o]);
        }
      }
    }
    // ---- window end: new metric
    const bool wend = (req & 2) != 0;
    const bool renew = wend && n >= 10.0;
    if (__syncthreads_or(renew)) {
      // (1) grad_theta = U_old^-1 grad_phi by back substitution, column by column; kept in sv
      {
        double x[kDenseOwn];
        int k = 0;
        for (int e = t.dl; e < D; e += kTileD, ++k) x[k] = renew ? st.grad[(size_t)e * C + t.c] : 0.0;
        for (int d = D - 1; d >= 0; --d) {
          const int slot = (D - 1 - d) & 1;
          if (renew && (d % kTileD) == t.dl) bc[slot][t.cl] = x[d / kTileD] / U[((size_t)d * D + d) * C + t.c];
          __syncthreads();
          if (renew) {
            const double gd = bc[slot][t.cl];
            if ((d % kTileD) == t.dl) sv[(size_t)d * C + t.c] = gd;
            k = 0;
            for (int e = t.dl; e < d; e += kTileD, ++k) x[k] = fma(-U[((size_t)e * D + d) * C + t.c], gd, x[k]);
          }
        }
      }
      // (2) M^-1 = n / ((n + 5)(n - 1)) cov + 1e-3 * 5 / (n + 5) I
      if (renew) {
        const double sc = n / ((n + 5.0) * (n - 1.0)), reg = 1e-3 * (5.0 / (n + 5.0));
        for (int i = t.dl; i < D; i += kTileD)
          for (int j = 0; j < D; ++j) {
            const size_t o = ((size_t)i * D + j) * C + t.c;
            A[o] = sc * cov[o] + (i == j ? reg : 0.0);
          }
      }
      __syncthreads();   // row j of M^-1 is written by the lane that owns row j and read by every column owner
      // (3) upper Cholesky factor, row by row (oracle: dense_chol_upper): the lane that owns column i
      //     computes U[j][i], every lane recomputes the pivot of the row
      for (int j = 0; j < D; ++j) {
        if (renew) {
          double s = A[((size_t)j * D + j) * C + t.c];
          for (int k = 0; k < j; ++k) {
            const double u = U[((size_t)k * D + j) * C + t.c];
            s = fma(-u, u, s);
          }
          const double ujj = s > 0.0 ? sqrt(s) : 1.0;   // the regulariser keeps the matrix positive definite
          for (int i = t.dl; i < D; i += kTileD) {
            const size_t o = ((size_t)j * D + i) * C + t.c;
            if (i < j) {
              U[o] = 0.0;   // rows below j have not been read yet; entries left of the diagonal are zero
            } else if (i == j) {
              U[o] = ujj;
            } else {
              double tt = A[o];
              for (int k = 0; k < j; ++k)
                tt = fma(-U[((size_t)k * D + j) * C + t.c], U[((size_t)k * D + i) * C + t.c], tt);
              U[o] = tt / ujj;
            }
          }
        }
        __syncthreads();
      }
      // (4) phi = U^-T theta by forward substitution, column by column; (5) grad_phi = U grad_theta
      {
        double x[kDenseOwn];
        int k = 0;
        for (int d = t.dl; d < D; d += kTileD, ++k) x[k] = renew ? st.theta[(size_t)d * C + t.c] : 0.0;
        for (int e = 0; e < D; ++e) {
          const int slot = e & 1;
          if (renew && (e % kTileD) == t.dl) bc[slot][t.cl] = x[e / kTileD] / U[((size_t)e * D + e) * C + t.c];
          __syncthreads();
          if (renew) {
            const double pe = bc[slot][t.cl];
            if ((e % kTileD) == t.dl) st.wf_m2[(size_t)e * C + t.c] = pe;
            k = 0;
            for (int d = t.dl; d < D; d += kTileD, ++k)
              if (d > e) x[k] = fma(-U[((size_t)e * D + d) * C + t.c], pe, x[k]);
          }
        }
      }
      if (renew) dense_u_apply(U, sv, st.grad, D, C, t);
    }
    // every window end: the Welford accumulators restart
    if (wend) {
      for (int i = t.dl; i < D; i += kTileD) {
        st.wf_mean[(size_t)i * C + t.c] = 0.0;
        for (int j = 0; j < D; ++j) cov[((size_t)i * D + j) * C + t.c] = 0.0;
      }
    }
    __syncthreads();
    if (in && t.dl == 0 && req != 0) {
      ls.dn_req[t.c] = 0;
      if (ls.ts.phase[t.c] == 3) ls.ts.phase[t.c] = 0;
    }
  }
  // ---- the position of the pending leaf in theta space, for the family's gradient kernel
  if (in && ls.ts.phase[t.c] == 1) dense_ut_apply(U, ls.zt, ls.wT, D, C, t);
}

}  // namespace tnuts
