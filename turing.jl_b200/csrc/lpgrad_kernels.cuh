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
#include "logistic_finish.cuh"
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

// sigma(eta) and log1pexp(eta) from ONE exponential: e = exp(-|eta|),
//   log1pexp = max(eta, 0) + log1p(e),  sigma = eta >= 0 ? 1/(1+e) : e/(1+e)
// (agrees with StatsFuns.log1pexp / logistic to rounding on the whole real line)
__device__ __forceinline__ void d_sigmoid_log1pexp(double x, double &sig, double &l1pe) {
  const double e = exp(-fabs(x));
  const double t = 1.0 / (1.0 + e);
  sig = x >= 0.0 ? t : e * t;
  l1pe = fmax(x, 0.0) + log1p(e);
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
// logistic regression: fused eta / residual / X^T R over one row chunk and one chain tile on
// the fp64 tensor path (DMMA, mma.sync.m8n8k4.f64).  tcgen05 has no f64 kind, and measured on
// this B200 DMMA and DFMA both peak at 37.2 TFLOP/s, so DMMA is used for its 8x lower issue
// and register-bandwidth pressure, not for a higher ceiling.
//
//   grid = (chain tiles, row chunks); 256 threads = 8 warps; one CTA per SM.
//   per row tile of TR rows (X tile double-buffered in shared memory with cp.async):
//     GEMM1  eta[TR x TCH]  = Xs[TR x D] * Bs[D x TCH]          warp grid WM1 x WN1
//     elementwise: r = y - sigma(eta), lp += y*eta - log1pexp(eta)  -> Rs[TR x TCH]
//     GEMM2  G[D x TCH]    += Xs^T[D x TR] * Rs[TR x TCH]       warp grid NWM x NWN,
//            accumulators (MT m-tiles x 2 doubles) stay in registers for the whole row chunk
//   Shared-memory leading dimensions are == 4 (mod 16) doubles, which makes every DMMA
//   fragment load (8 rows x 4 k or 4 k x 8 cols per warp) bank-conflict free.
//
// Fragment layout of mma.m8n8k4.f64 (lane l): a0 = A[l>>2][l&3], b0 = B[l&3][l>>2],
// c0,c1 = C[l>>2][2*(l&3) + {0,1}].
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__host__ __device__ inline int ld_pad(int n) {  // smallest m >= n with m == 4 (mod 16)
  int m = n + ((4 - (n % 16)) + 16) % 16;
  return m;
}

template <int NWN, int NWM, int MT, int TR, int NG>
__global__ void __launch_bounds__(256 * NG, 1)
k_logistic_mma(const double *__restrict__ X, const double *__restrict__ y, long long nrows, int D,
               const double *__restrict__ th, int C, const int *__restrict__ list,
               const int *__restrict__ count, int n_all, int rows_per_chunk,
               double *__restrict__ part /* [chunks][D+1][cap] */, int cap) {
  constexpr int TCH = 8 * NWN;                 // chains per CTA
  constexpr int WN1 = (NWN >= 8) ? 4 : 2;      // GEMM1 warp grid
  constexpr int WM1 = 8 / WN1;
  constexpr int MT1 = (TR / 8) / WM1;          // m-tiles per warp (rows)
  constexpr int NT1 = NWN / WN1;               // n-tiles per warp (chains)
  static_assert(NWN * NWM == 8 && MT1 >= 1 && NT1 >= 1, "bad tiling");
  constexpr int LDB = TCH + 4, LDR = TCH + 4;
  extern __shared__ double sm[];
  const int Kp = (D + 3) & ~3;
  const int LDX = ld_pad(Kp);
  // NG independent groups of 8 warps share the B tile and work on different row chunks; their
  // GEMM / elementwise phases drift apart, so one group's DMMAs overlap the other's exp/log
  const int grp = threadIdx.x >> 8;
  double *Bs = sm;                              // [Kp][LDB]                (shared by the groups)
  double *Xs = Bs + (size_t)Kp * LDB + (size_t)grp * ((size_t)2 * TR * LDX + (size_t)TR * LDR + 2 * TR);
  double *Rs = Xs + (size_t)2 * TR * LDX;       // [TR][LDR]
  double *Ys = Rs + (size_t)TR * LDR;           // [2][TR]
  const int n = count ? *count : n_all;
  const int k0 = blockIdx.x * TCH;
  if (k0 >= n) return;
  const int tid = threadIdx.x & 255, lane = tid & 31, warp = tid >> 5;
  const int lr = lane >> 2, lc = lane & 3;
  const int chunk = blockIdx.y * NG + grp;
  auto group_sync = [&]() {
    if (NG == 1) __syncthreads();
    else asm volatile("bar.sync %0, 256;\n" ::"r"(1 + grp));
  };
  long long rbeg = (long long)chunk * rows_per_chunk;
  if (rbeg > nrows) rbeg = nrows;
  long long rend = rbeg + rows_per_chunk;
  if (rend > nrows) rend = nrows;
  const int ntile = (int)((rend - rbeg + TR - 1) / TR);

  auto prefetch = [&](int t, int buf) {
    const long long r0 = rbeg + (long long)t * TR;
    const int nr = (int)((rend - r0) < TR ? (rend - r0) : TR);
    double *xb = Xs + (size_t)buf * TR * LDX;
    const double *xg = X + (size_t)r0 * D;
    // 8 rows per pass: thread (tid >> 5) picks the row, the 32 lanes stride along it
    for (int r = tid >> 5; r < TR; r += 8) {
      if (r < nr) {
        for (int d = lane; d < D; d += 32) cp_async8(xb + r * LDX + d, xg + (size_t)r * D + d);
      } else {
        for (int d = lane; d < D; d += 32) xb[r * LDX + d] = 0.0;
      }
    }
    if (tid < TR) {
      if (tid < nr) cp_async8(Ys + buf * TR + tid, y + r0 + tid);
      else Ys[buf * TR + tid] = 0.0;
    }
  };

  // zero the K padding of both X buffers once, load the B tile through the work list
  for (int i = tid; i < 2 * TR * (LDX - D); i += 256) {
    const int r = i / (LDX - D), d = D + i % (LDX - D);
    Xs[(size_t)r * LDX + d] = 0.0;
  }
  for (int i = threadIdx.x; i < Kp * TCH; i += 256 * NG) {
    const int d = i / TCH, cc = i - d * TCH;
    const int k = k0 + cc;
    double v = 0.0;
    if (k < n && d < D) {
      const int c = list ? list[k] : k;
      v = th[(size_t)d * C + c];
    }
    Bs[d * LDB + cc] = v;
  }
  if (ntile > 0) prefetch(0, 0);
  cp_async_commit();
  if (NG > 1) __syncthreads();  // the B tile is complete before the groups go their own way

  // GEMM2 accumulators: this warp's n-tile (wn2) and m-tiles {wm2 + NWM * i}
  const int wn2 = warp % NWN, wm2 = warp / NWN;
  double gacc[MT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i) gacc[i][0] = gacc[i][1] = 0.0;
  // GEMM1 position
  const int wm1 = warp / WN1, wn1 = warp % WN1;
  double lpacc[NT1][2];
#pragma unroll
  for (int q = 0; q < NT1; ++q) lpacc[q][0] = lpacc[q][1] = 0.0;

  for (int t = 0; t < ntile; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntile) prefetch(t + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    group_sync();
    const double *xb = Xs + (size_t)buf * TR * LDX;
    const long long r0 = rbeg + (long long)t * TR;
    const int nr = (int)((rend - r0) < TR ? (rend - r0) : TR);
    // ---- GEMM1: eta = X * B ----
    double eta[MT1][NT1][2];
#pragma unroll
    for (int i = 0; i < MT1; ++i)
#pragma unroll
      for (int q = 0; q < NT1; ++q) eta[i][q][0] = eta[i][q][1] = 0.0;
    for (int kk = 0; kk < Kp; kk += 4) {
      double a[MT1], b[NT1];
#pragma unroll
      for (int i = 0; i < MT1; ++i) a[i] = xb[((wm1 * MT1 + i) * 8 + lr) * LDX + kk + lc];
#pragma unroll
      for (int q = 0; q < NT1; ++q) b[q] = Bs[(kk + lc) * LDB + (wn1 * NT1 + q) * 8 + lr];
#pragma unroll
      for (int i = 0; i < MT1; ++i)
#pragma unroll
        for (int q = 0; q < NT1; ++q) dmma884(eta[i][q][0], eta[i][q][1], a[i], b[q]);
    }
    // ---- residuals / log-likelihood ----
#pragma unroll
    for (int i = 0; i < MT1; ++i) {
      const int row = (wm1 * MT1 + i) * 8 + lr;
      const double yy = Ys[buf * TR + row];
      const bool live = row < nr;
#pragma unroll
      for (int q = 0; q < NT1; ++q) {
        double r2[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const double e = eta[i][q][h];
          double r = 0.0;
          if (live) {
            double sg, l1;
            d_sigmoid_log1pexp(e, sg, l1);
            r = yy - sg;
            lpacc[q][h] += yy * e - l1;
          }
          r2[h] = r;
        }
        double *dst = Rs + row * LDR + (wn1 * NT1 + q) * 8 + 2 * lc;
        dst[0] = r2[0];
        dst[1] = r2[1];
      }
    }
    group_sync();
    // ---- GEMM2: G += X^T * R ----
    for (int kk = 0; kk < TR; kk += 4) {
      const double b = Rs[(kk + lc) * LDR + wn2 * 8 + lr];
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int d = (wm2 + NWM * i) * 8 + lr;
        const double a = (d < D) ? xb[(kk + lc) * LDX + d] : 0.0;
        dmma884(gacc[i][0], gacc[i][1], a, b);
      }
    }
    group_sync();  // Rs and this X buffer are free for the next iteration
  }
  cp_async_wait<0>();
  // ---- epilogue ----
  if (rbeg >= nrows && chunk > 0) return;  // a group without rows (its chunk slot does not exist)
  double *pc = part + (size_t)chunk * (D + 1) * cap;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int d = (wm2 + NWM * i) * 8 + lr;
    if (d < D) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = k0 + wn2 * 8 + 2 * lc + h;
        if (k < n) pc[(size_t)d * cap + k] = gacc[i][h];
      }
    }
  }
  // lp: sum over the 8 row-lanes of the warp, then over the WM1 warps sharing the columns
#pragma unroll
  for (int q = 0; q < NT1; ++q)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double v = lpacc[q][h];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      lpacc[q][h] = v;
    }
  double *Ls = Rs;  // [WM1][TCH]
  if (lr == 0) {
#pragma unroll
    for (int q = 0; q < NT1; ++q)
#pragma unroll
      for (int h = 0; h < 2; ++h) Ls[wm1 * TCH + (wn1 * NT1 + q) * 8 + 2 * lc + h] = lpacc[q][h];
  }
  group_sync();
  if (tid < TCH) {
    double s = 0.0;
    for (int w = 0; w < WM1; ++w) s += Ls[w * TCH + tid];
    const int k = k0 + tid;
    if (k < n) pc[(size_t)D * cap + k] = s;
  }
}

// sum the row-chunk partials in a fixed order -> lik[D+1][cap] (compact, by work-list slot)
__global__ void k_logistic_reduce(const double *part, int chunks, int D, int stride, const int *count,
                                  int n_all, double *lik, int cap) {
  const int n = count ? *count : n_all;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int d = blockIdx.y;
  if (k >= n) return;
  double s = 0.0;
  for (int ch = 0; ch < chunks; ++ch) s += part[((size_t)ch * (D + 1) + d) * stride + k];
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

// k_logistic_reduce + k_logistic_finish in one launch (single-GPU rows): block = 32 work-list
// slots x 8 rows of the (D+1)-row partial block, one row per thread.  The sum over the row chunks
// runs in the same fixed order (chunk 0, 1, ...), so the result is bit-identical to the
// two-kernel path; the thread that owns the lp row also evaluates the prior.
__global__ void __launch_bounds__(256)
k_logistic_reduce_finish(ModelDev m, const double *__restrict__ part, int chunks, int stride,
                         const double *__restrict__ th, int C, const int *__restrict__ list,
                         const int *__restrict__ count, int n_all, double *__restrict__ lp_out,
                         double *__restrict__ g) {
  const int n = count ? *count : n_all;
  const int k = blockIdx.x * 32 + (threadIdx.x & 31);
  const int d = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int D = m.D;
  if (k >= n || d > D) return;
  const int c = list ? list[k] : k;
  const size_t cs = (size_t)(D + 1) * stride;
  const double *p = part + (size_t)d * stride + k;
  double s = 0.0;
  int ch = 0;
  for (; ch + 4 <= chunks; ch += 4) {  // four loads in flight, summed in chunk order
    const double a0 = p[(size_t)ch * cs], a1 = p[(size_t)(ch + 1) * cs], a2 = p[(size_t)(ch + 2) * cs],
                 a3 = p[(size_t)(ch + 3) * cs];
    s += a0;
    s += a1;
    s += a2;
    s += a3;
  }
  for (; ch < chunks; ++ch) s += p[(size_t)ch * cs];
  const double sd = m.hyper[0], is2 = 1.0 / (sd * sd);
  if (d < D) {
    g[(size_t)d * C + c] = logistic_finish_grad(s, th[(size_t)d * C + c], is2);
  } else {
    double q = 0.0;
    for (int j = 0; j < D; ++j) {
      const double t = th[(size_t)j * C + c];
      q = __fma_rn(t, t, q);
    }
    lp_out[c] = logistic_finish_lp(s, q, is2, D, sd);
  }
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
