// lpgrad_kernels.cuh -- batched log-density + gradient kernels of the lock-step strategy.
//
// Every kernel evaluates `LogDensityProblems.logdensity_and_gradient(ldf, theta)`
// (src/mcmc/hmc.jl:181) for MANY chains at once: theta is [D][C] (chain fastest), the work
// list names the chains that still need a gradient (`list[k]`, `*count` entries; NULL = all).
//
//   k_lpgrad_small ........ gdemo / linreg / eight schools / std-normal at any size, one thread
//                           per chain (these are O(N + D) per chain)
//   k_logistic_tile ....... Bayesian logistic regression.  The per-chain likelihood over N rows
//                           becomes a dense contraction across chains:
//                               eta = X * B        [rows x D] * [D x chains]
//                               R   = y - sigma(eta), lp += y*eta - log1pexp(eta)
//                               G  += X^T * R      [D x rows] * [rows x chains]
//                           fused per row tile so X is read ONCE per gradient; row chunks go to
//                           different CTAs and are summed in a fixed order (deterministic).
//   k_hier_tile ........... hierarchical regression from per-group centred sufficient
//                           statistics: O(D) per chain, streaming over [D][C].
#pragma once
#include "lockstep_kernels.cuh"

namespace tnuts {

__device__ __forceinline__ double d_log1pexp(double x) {
  // StatsFuns.log1pexp (src/stdlib/distributions.jl:77)
  if (x < -37.0) return exp(x);
  if (x < 18.0) return log1p(exp(x));
  if (x < 33.3) return x + exp(-x);
  return x;
}
__device__ __forceinline__ double d_logistic(double x) {
  if (x >= 0) return 1.0 / (1.0 + exp(-x));
  const double e = exp(x);
  return e / (1.0 + e);
}

// ------------------------------------------------------------------------------------------
__global__ void k_lpgrad_small(ModelDev m, const double *th, int C, double *lp_out, double *g,
                               const int *list, const int *count, int n_all) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = count ? *count : n_all;
  if (k >= n) return;
  const int c = list ? list[k] : k;
  auto T = [&](int d) { return th[(size_t)d * C + c]; };
  auto Gd = [&](int d) -> double & { return g[(size_t)d * C + c]; };
  double lp = 0.0;
  switch (m.family) {
    case TNUTS_GDEMO: {
      const double al = m.hyper[0], be = m.hyper[1], a = T(0), mm = T(1), s = exp(a);
      double Q = mm * mm, sres = 0.0;
      for (long long i = 0; i < m.N; ++i) {
        const double d = m.y[i] - mm;
        Q += d * d;
        sres += d;
      }
      const double n1 = (double)(m.N + 1);
      lp = al * log(be) - lgamma(al) - (al + 1.0) * a - be / s - 0.5 * n1 * kLog2Pi - 0.5 * n1 * a -
           Q / (2.0 * s) + a;
      Gd(0) = -(al + 1.0) + be / s - 0.5 * n1 + Q / (2.0 * s) + 1.0;
      Gd(1) = (-mm + sres) / s;
    } break;
    case TNUTS_LINREG: {
      const double sd = m.hyper[0], cc = m.hyper[1];
      const double al = T(0), be = T(1), t = T(2), v = exp(t);
      double ssr = 0, sr = 0, srx = 0;
      for (long long i = 0; i < m.N; ++i) {
        const double x = m.X[i];
        const double res = m.y[i] - al - be * x;
        ssr += res * res;
        sr += res;
        srx += res * x;
      }
      const double nn = (double)m.N, w = (v / cc) * (v / cc);
      lp = -kLog2Pi - 2.0 * log(sd) - 0.5 * (al * al + be * be) / (sd * sd) + kLog2 - kLogPi - log(cc) -
           log1p(w) + t - 0.5 * nn * (kLog2Pi + t) - ssr / (2.0 * v);
      Gd(0) = -al / (sd * sd) + sr / v;
      Gd(1) = -be / (sd * sd) + srx / v;
      Gd(2) = -2.0 * w / (1.0 + w) + 1.0 - 0.5 * nn + ssr / (2.0 * v);
    } break;
    case TNUTS_EIGHT_SCHOOLS: {
      const double sd = m.hyper[0], cc = m.hyper[1];
      const int J = (int)m.N;
      const double mu = T(0), lt = T(1), tau = exp(lt), w = (tau / cc) * (tau / cc);
      lp = -0.5 * kLog2Pi - log(sd) - 0.5 * mu * mu / (sd * sd) + kLog2 - kLogPi - log(cc) - log1p(w) + lt;
      double sz = 0, szt = 0;
      for (int jn = 0; jn < J; ++jn) {
        const double tj = T(2 + jn), sg = m.sigma[jn];
        const double res = m.y[jn] - mu - tau * tj;
        const double z = res / (sg * sg);
        lp += -0.5 * kLog2Pi - 0.5 * tj * tj - 0.5 * kLog2Pi - log(sg) - 0.5 * res * z;
        sz += z;
        szt += z * tj;
        Gd(2 + jn) = -tj + tau * z;
      }
      Gd(0) = -mu / (sd * sd) + sz;
      Gd(1) = -2.0 * w / (1.0 + w) + 1.0 + tau * szt;
    } break;
    case TNUTS_STD_NORMAL: {
      lp = -0.5 * m.D * kLog2Pi;
      for (int d = 0; d < m.D; ++d) {
        const double x = T(d);
        lp -= 0.5 * x * x;
        Gd(d) = -x;
      }
    } break;
    default:
      lp = CUDART_NAN;
  }
  lp_out[c] = lp;
}

// ------------------------------------------------------------------------------------------
// logistic regression: fused eta / residual / X^T R over one row chunk and one chain tile.
//   grid = (chain tiles, row chunks), 256 threads.
//   TR rows x TCH chains per tile; thread (ty = tid / 16, tx = tid % 16) owns the
//   (TR/16) x (TCH/16) block of eta and, for G, the dims {ty + 16 i} x (TCH/16) chains.
// ------------------------------------------------------------------------------------------
template <int TR, int TCH, int ND>
__global__ void __launch_bounds__(256)
k_logistic_tile(const double *__restrict__ X, const double *__restrict__ y, long long row0,
                long long nrows, int D, const double *__restrict__ th, int C,
                const int *__restrict__ list, const int *__restrict__ count, int n_all,
                int rows_per_chunk, double *__restrict__ part /* [chunks][D+1][cap] */, int cap) {
  constexpr int RM = TR / 16, CM = TCH / 16;
  extern __shared__ double sm[];
  double *Bs = sm;                       // [D][TCH]
  double *Xs = Bs + (size_t)D * TCH;     // [TR][D]
  double *Rs = Xs + (size_t)TR * D;      // [TR][TCH]
  double *Ys = Rs + (size_t)TR * TCH;    // [TR]
  const int n = count ? *count : n_all;
  const int k0 = blockIdx.x * TCH;
  if (k0 >= n) return;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int chunk = blockIdx.y;
  // B tile (gathered through the work list)
  for (int i = tid; i < D * TCH; i += 256) {
    const int d = i / TCH, cc = i % TCH;
    const int k = k0 + cc;
    double v = 0.0;
    if (k < n) {
      const int c = list ? list[k] : k;
      v = th[(size_t)d * C + c];
    }
    Bs[i] = v;
  }
  double gacc[ND][CM];
#pragma unroll
  for (int i = 0; i < ND; ++i)
#pragma unroll
    for (int q = 0; q < CM; ++q) gacc[i][q] = 0.0;
  double lpacc[CM];
#pragma unroll
  for (int q = 0; q < CM; ++q) lpacc[q] = 0.0;

  const long long rbeg = (long long)chunk * rows_per_chunk;
  long long rend = rbeg + rows_per_chunk;
  if (rend > nrows) rend = nrows;
  for (long long r0 = rbeg; r0 < rend; r0 += TR) {
    const int nr = (int)((rend - r0) < TR ? (rend - r0) : TR);
    __syncthreads();  // previous tile fully consumed (also covers the B tile on the 1st pass)
    // X tile: TR*D contiguous doubles
    const double *xg = X + (size_t)(row0 + r0) * D;
    for (int i = tid; i < TR * D; i += 256) Xs[i] = (i < nr * D) ? xg[i] : 0.0;
    if (tid < TR) Ys[tid] = (tid < nr) ? y[row0 + r0 + tid] : 0.0;
    __syncthreads();
    // eta micro-tile
    double eta[RM][CM];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
      for (int q = 0; q < CM; ++q) eta[i][q] = 0.0;
    for (int k = 0; k < D; ++k) {
      double a[RM], b[CM];
#pragma unroll
      for (int i = 0; i < RM; ++i) a[i] = Xs[(ty * RM + i) * D + k];
#pragma unroll
      for (int q = 0; q < CM; ++q) b[q] = Bs[k * TCH + tx * CM + q];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int q = 0; q < CM; ++q) eta[i][q] += a[i] * b[q];
    }
    // residuals + log-likelihood terms
#pragma unroll
    for (int i = 0; i < RM; ++i) {
      const int row = ty * RM + i;
      const double yy = Ys[row];
      const bool live = row < nr;
#pragma unroll
      for (int q = 0; q < CM; ++q) {
        const double e = eta[i][q];
        double r = 0.0;
        if (live) {
          r = yy - d_logistic(e);
          lpacc[q] += yy * e - d_log1pexp(e);
        }
        Rs[row * TCH + tx * CM + q] = r;
      }
    }
    __syncthreads();
    // G += X^T R
    for (int row = 0; row < TR; ++row) {
      double r[CM];
#pragma unroll
      for (int q = 0; q < CM; ++q) r[q] = Rs[row * TCH + tx * CM + q];
#pragma unroll
      for (int i = 0; i < ND; ++i) {
        const int d = ty + 16 * i;
        const double x = (d < D) ? Xs[row * D + d] : 0.0;
#pragma unroll
        for (int q = 0; q < CM; ++q) gacc[i][q] += x * r[q];
      }
    }
  }
  // epilogue: partial gradient of this row chunk, and the lp partial (fixed-order over ty)
  double *pc = part + (size_t)chunk * (D + 1) * cap;
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    const int d = ty + 16 * i;
    if (d < D) {
#pragma unroll
      for (int q = 0; q < CM; ++q) {
        const int k = k0 + tx * CM + q;
        if (k < n) pc[(size_t)d * cap + k] = gacc[i][q];
      }
    }
  }
  __syncthreads();
  double *Ls = Rs;  // [16][TCH]
#pragma unroll
  for (int q = 0; q < CM; ++q) Ls[ty * TCH + tx * CM + q] = lpacc[q];
  __syncthreads();
  if (tid < TCH) {
    double s = 0.0;
    for (int w = 0; w < 16; ++w) s += Ls[w * TCH + tid];
    const int k = k0 + tid;
    if (k < n) pc[(size_t)D * cap + k] = s;
  }
}

// sum the row-chunk partials in a fixed order -> lik[D+1][cap] (compact, by work-list slot)
__global__ void k_logistic_reduce(const double *part, int chunks, int D, int cap, const int *count,
                                  int n_all, double *lik) {
  const int n = count ? *count : n_all;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int d = blockIdx.y;
  if (k >= n) return;
  double s = 0.0;
  for (int ch = 0; ch < chunks; ++ch) s += part[((size_t)ch * (D + 1) + d) * cap + k];
  lik[(size_t)d * cap + k] = s;
}

// add the prior and scatter to the chain-indexed outputs
__global__ void k_logistic_finish(ModelDev m, const double *lik, int cap, const double *th, int C,
                                  const int *list, const int *count, int n_all, double prior_scale,
                                  double *lp_out, double *g) {
  const int n = count ? *count : n_all;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int c = list ? list[k] : k;
  const int D = m.D;
  const double sd = m.hyper[0], is2 = 1.0 / (sd * sd);
  double q = 0.0;
  for (int d = 0; d < D; ++d) {
    const double t = th[(size_t)d * C + c];
    q += t * t;
    g[(size_t)d * C + c] = lik[(size_t)d * cap + k] - prior_scale * t * is2;
  }
  lp_out[c] = lik[(size_t)D * cap + k] + prior_scale * (-0.5 * D * kLog2Pi - D * log(sd) - 0.5 * q * is2);
}

// ------------------------------------------------------------------------------------------
// hierarchical linear regression from centred per-group sufficient statistics:
//   n_g, xbar_g, ybar_g, Sxx_g = sum (x - xbar)^2, Sxy_g, Syy_g
//   sum_i res_i        = n_g (ybar - a - b xbar)                       =: n_g m
//   sum_i res_i x_i    = (Sxy - b Sxx) + xbar n_g m
//   sum_i res_i^2      = Syy - 2 b Sxy + b^2 Sxx + n_g m^2
// tile kernel over the work list: 8 chains x 32 group-lanes per CTA.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads)
k_hier_tile(ModelDev m, LockstepState ls, const double *th, int C, const int *list, const int *count,
            int n_all, double *lp_out, double *g) {
  __shared__ double red[5 * 64];
  const int n = count ? *count : n_all;
  const int cl = threadIdx.x & 7, dl = threadIdx.x >> 3;
  const int k = blockIdx.x * kTileC + cl;
  if (blockIdx.x * kTileC >= n) return;
  const bool in = k < n;
  const int c = in ? (list ? list[k] : k) : 0;
  const int G = m.G;
  double mua = 0, mub = 0, la = 0, lb = 0, ly = 0, isa2 = 0, isb2 = 0, isy2 = 0;
  if (in) {
    mua = th[c];
    mub = th[(size_t)C + c];
    la = th[(size_t)2 * C + c];
    lb = th[(size_t)3 * C + c];
    ly = th[(size_t)4 * C + c];
    isa2 = exp(-2.0 * la);
    isb2 = exp(-2.0 * lb);
    isy2 = exp(-2.0 * ly);
  }
  double acc[5] = {0, 0, 0, 0, 0};  // qa, qb, da, db, ssr
  if (in) {
    for (int gi = dl; gi < G; gi += kTileD) {
      const size_t oa = (size_t)(5 + gi) * C + c, ob = (size_t)(5 + G + gi) * C + c;
      const double a = th[oa], b = th[ob];
      const double ea = a - mua, eb = b - mub;
      const double ng = ls.hs_n[gi], xb = ls.hs_xbar[gi];
      const double mres = ls.hs_ybar[gi] - a - b * xb;
      const double sres = ng * mres;
      const double sxy = ls.hs_sxy[gi], sxx = ls.hs_sxx[gi];
      const double sresx = (sxy - b * sxx) + xb * sres;
      acc[0] += ea * ea;
      acc[1] += eb * eb;
      acc[2] += ea;
      acc[3] += eb;
      acc[4] += ls.hs_syy[gi] - 2.0 * b * sxy + b * b * sxx + sres * mres;
      g[oa] = -ea * isa2 + sres * isy2;
      g[ob] = -eb * isb2 + sresx * isy2;
    }
  }
  tile_reduce<5>(acc, red);
  if (in && dl == 0) {
    const double sdm = m.hyper[0], s0 = m.hyper[1];
    const double sa2 = exp(2.0 * la), sb2 = exp(2.0 * lb), sy2 = exp(2.0 * ly);
    const double nn = (double)m.N, Gd = (double)G;
    double lp = 2.0 * (-0.5 * kLog2Pi - log(sdm)) - 0.5 * (mua * mua + mub * mub) / (sdm * sdm);
    lp += 3.0 * (kLog2 - 0.5 * kLog2Pi - log(s0)) - 0.5 * (sa2 + sb2 + sy2) / (s0 * s0) + la + lb + ly;
    lp += -Gd * kLog2Pi - Gd * la - Gd * lb - 0.5 * acc[0] * isa2 - 0.5 * acc[1] * isb2;
    lp += -0.5 * nn * kLog2Pi - nn * ly - 0.5 * acc[4] * isy2;
    lp_out[c] = lp;
    g[c] = -mua / (sdm * sdm) + acc[2] * isa2;
    g[(size_t)C + c] = -mub / (sdm * sdm) + acc[3] * isb2;
    g[(size_t)2 * C + c] = -sa2 / (s0 * s0) + 1.0 - Gd + acc[0] * isa2;
    g[(size_t)3 * C + c] = -sb2 / (s0 * s0) + 1.0 - Gd + acc[1] * isb2;
    g[(size_t)4 * C + c] = -sy2 / (s0 * s0) + 1.0 - nn + acc[4] * isy2;
  }
}

}  // namespace tnuts
