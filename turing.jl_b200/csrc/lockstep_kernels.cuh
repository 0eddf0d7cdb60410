// lockstep_kernels.cuh -- streaming kernels of the lock-step strategy.
//
// Thread mapping of every "tile" kernel: a CTA of 256 threads owns kTileC = 8 consecutive
// chains; thread t handles chain (t & 7) and the dimensions d = (t >> 3), +32, +64, ...
// For a fixed d the 8 chains are 64 contiguous bytes (two full sectors), a warp touches four
// such rows.  Reductions over d (kinetic energy, U-turn dot products, finiteness) go
// warp-shuffle -> shared memory -> fixed-order sum, so every result is deterministic.
//
// Algorithmic content (per chain) is identical to chain_kernels.cuh; see there for the
// mapping onto the reference's functions.
#pragma once
#include "lockstep.cuh"
#include "philox.cuh"

namespace tnuts {

__device__ __forceinline__ double ls_logaddexp(double a, double b) {
  if (a == -CUDART_INF) return b;
  if (b == -CUDART_INF) return a;
  const double mx = fmax(a, b);
  return mx + log1p(exp(-fabs(a - b)));
}

struct Tile {
  int c, cl, dl;
  __device__ Tile() {
    cl = threadIdx.x & (kTileC - 1);
    dl = threadIdx.x >> 3;
    c = blockIdx.x * kTileC + cl;
  }
};

// sum over the 32 d-lanes of each chain for NR values at once; result valid in every thread
template <int NR>
__device__ __forceinline__ void tile_reduce(double (&v)[NR], double *smem /* [NR][8][8] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cl = threadIdx.x & 7;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    v[r] += __shfl_xor_sync(0xffffffffu, v[r], 8);
    v[r] += __shfl_xor_sync(0xffffffffu, v[r], 16);
  }
  __syncthreads();  // protect smem reuse between consecutive calls
  if (lane < 8) {
#pragma unroll
    for (int r = 0; r < NR; ++r) smem[(r * 8 + warp) * 8 + lane] = v[r];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += smem[(r * 8 + w) * 8 + cl];
    v[r] = s;
  }
}

// ------------------------------------------------------------------------------------------
// transition start: momentum refresh, H0, tree initialisation
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads)
k_ls_transition_start(ChainState st, LockstepState ls, RunParams rp) {
  __shared__ double red[1 * 64];
  const Tile t;
  const int C = st.C, D = st.D;
  const bool in = t.c < C && st.status[t.c] == 0;
  const uint32_t gchain = (uint32_t)(rp.chain_offset + t.c);
  const uint32_t it = in ? (uint32_t)(st.iter[t.c] + 1) : 0u;
  double acc[1] = {0.0};
  if (in) {
    const int npair = (D + 1) / 2;
    for (int p = t.dl; p < npair; p += kTileD) {
      double z0, z1;
      normal2(rp.seed, gchain, it, SLOT_MOMENTUM, (uint32_t)p, z0, z1);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int d = 2 * p + q;
        if (d < D) {
          const size_t o = (size_t)d * C + t.c;
          const double mi = st.minv[o];
          const double r = (q ? z1 : z0) / sqrt(mi);
          const double th = st.theta[o], g = st.grad[o];
          acc[0] += mi * r * r;
          ls.rho[o] = r;
          ls.thC[o] = th;
          ls.gC[o] = g;
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            double *e = ls.edge[s];
            e[o] = th;
            e[(size_t)D * C + o] = r;
            e[2 * (size_t)D * C + o] = g;
          }
        }
      }
    }
  }
  tile_reduce<1>(acc, red);
  if (in && t.dl == 0) {
    const double lp = st.lp[t.c];
    const double K = 0.5 * acc[0];
    const double H0 = (isfinite(lp) && isfinite(K)) ? (-lp + K) : CUDART_INF;
    TreeScalars &s = ls.ts;
    s.H0[t.c] = H0;
    s.sum_a[t.c] = 0.0;
    s.n_a[t.c] = 0;
    s.dHmax[t.c] = 0.0;
    s.lw[t.c] = 0.0;
    s.lpC[t.c] = lp;
    s.HC[t.c] = H0;
    s.elp[t.c] = lp;
    s.elp[C + t.c] = lp;
    s.depth[t.c] = 0;
    s.active[t.c] = 1;
    s.numerr[t.c] = 0;
  }
  if (t.c < C && !in && t.dl == 0) ls.ts.active[t.c] = 0;
}

// ------------------------------------------------------------------------------------------
// doubling start: direction, load the edge, first half kick + drift of leaf 0
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads)
k_ls_doubling_start(ChainState st, LockstepState ls, RunParams rp, int j) {
  const Tile t;
  const int C = st.C, D = st.D;
  if (t.c >= C) return;
  TreeScalars &s = ls.ts;
  const bool act = s.active[t.c] != 0;
  if (!act) {
    if (t.dl == 0) s.sub_active[t.c] = 0;
    return;
  }
  const uint32_t gchain = (uint32_t)(rp.chain_offset + t.c);
  const uint32_t it = (uint32_t)(st.iter[t.c] + 1);
  const int v = (uniform1(rp.seed, gchain, it, SLOT_DIRECTION, (uint32_t)j) < 0.5) ? 0 : 1;
  const double eps = st.eps[t.c];
  const double se = v ? eps : -eps, half = 0.5 * se;
  const double *e = ls.edge[v];
  const size_t DC = (size_t)D * C;
  for (int d = t.dl; d < D; d += kTileD) {
    const size_t o = (size_t)d * C + t.c;
    const double g = e[2 * DC + o];
    const double r = e[DC + o] + half * g;
    ls.zr[o] = r;
    ls.zg[o] = g;
    ls.zt[o] = e[o] + se * (st.minv[o] * r);
    ls.rho_s[o] = 0.0;
  }
  if (t.dl == 0) {
    s.v[t.c] = v;
    s.sub_active[t.c] = 1;
    s.term_s[t.c] = 0;
    s.lw_s[t.c] = -CUDART_INF;
    s.zlp[t.c] = s.elp[v * C + t.c];
  }
}

// ------------------------------------------------------------------------------------------
// leaf: finish leapfrog n of doubling j (second half kick, energy), leaf statistics,
// streaming multinomial candidate, checkpointed sub-tree U-turn checks; then, when the
// doubling continues, the first half kick + drift of leaf n+1.
// The gradient at the drifted position was written to zg / zlp by the family's kernel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads)
k_ls_leaf(ChainState st, LockstepState ls, RunParams rp, int j, int n, int nleaf) {
  __shared__ double red[2 * 64];
  const Tile t;
  const int C = st.C, D = st.D;
  TreeScalars &s = ls.ts;
  const bool in = t.c < C && s.sub_active[t.c] != 0;
  double eps = 0, se = 0, half = 0;
  if (in) {
    eps = st.eps[t.c];
    se = s.v[t.c] ? eps : -eps;
    half = 0.5 * se;
  }
  // pass 1: second half kick, kinetic energy, finiteness of the gradient
  double a1[2] = {0.0, 0.0};
  if (in) {
    for (int d = t.dl; d < D; d += kTileD) {
      const size_t o = (size_t)d * C + t.c;
      const double g = ls.zg[o];
      const double r = ls.zr[o] + half * g;
      ls.zr[o] = r;
      a1[0] += st.minv[o] * r * r;
      a1[1] += isfinite(g) ? 0.0 : 1.0;
    }
  }
  tile_reduce<2>(a1, red);
  // per-chain leaf logic, computed identically by the 32 threads of the chain
  bool accept = false, divergent = false;
  if (in) {
    const uint32_t gchain = (uint32_t)(rp.chain_offset + t.c);
    const uint32_t it = (uint32_t)(st.iter[t.c] + 1);
    const double lp = s.zlp[t.c];
    const double K = 0.5 * a1[0];
    const bool ok = isfinite(lp) && a1[1] == 0.0 && isfinite(K);
    const double H = ok ? (-lp + K) : CUDART_INF;
    const double dH = H - s.H0[t.c];
    const double a = exp(fmin(0.0, -dH));
    const double lw_leaf = -dH;
    const double lw_s = s.lw_s[t.c];
    const double lw_new = ls_logaddexp(lw_s, lw_leaf);
    const double p = (lw_leaf == -CUDART_INF) ? 0.0 : exp(lw_leaf - lw_new);
    const double u = uniform1(rp.seed, gchain, it, SLOT_LEAF, (1u << j) + (uint32_t)n);
    accept = u < p;
    divergent = !(dH < rp.delta_max);
    if (t.dl == 0) {
      s.sum_a[t.c] += a;
      s.n_a[t.c] += 1;
      if (fabs(dH) > fabs(s.dHmax[t.c])) s.dHmax[t.c] = dH;
      s.lw_s[t.c] = lw_new;
      s.zH[t.c] = H;
      if (accept) {
        s.lpS[t.c] = lp;
        s.HS[t.c] = H;
      }
      if (divergent) {
        s.numerr[t.c] = 1;
        s.term_s[t.c] = 1;
        s.sub_active[t.c] = 0;
      }
      st.n_grad[t.c] += 1;
    }
  }
  // pass 2: rho_s, candidate copy, checkpoints / U-turn partial dot products
  const int idx_max = __popc((unsigned)n >> 1);
  const bool odd = (n & 1) != 0;
  const int ntrail = __ffs(~(unsigned)n) - 1;
  const int idx_min = idx_max - ntrail + 1;
  const int nlev = odd ? (idx_max - idx_min + 1) : 0;
  double dots[2 * kMaxLevels];
#pragma unroll
  for (int q = 0; q < 2 * kMaxLevels; ++q) dots[q] = 0.0;
  const size_t DC = (size_t)D * C;
  if (in) {
    for (int d = t.dl; d < D; d += kTileD) {
      const size_t o = (size_t)d * C + t.c;
      const double r = ls.zr[o];
      const double rs = ls.rho_s[o] + r;
      ls.rho_s[o] = rs;
      if (accept) {
        ls.thS[o] = ls.zt[o];
        ls.gS[o] = ls.zg[o];
      }
      if (!divergent) {
        if (!odd) {
          ls.ck_r[idx_max * DC + o] = r;
          ls.ck_rho[idx_max * DC + o] = rs;
        } else {
          const double mi = st.minv[o];
          for (int q = 0; q < nlev; ++q) {
            const int i = idx_max - q;
            const double cr = ls.ck_r[i * DC + o];
            const double sr = rs - ls.ck_rho[i * DC + o] + cr;
            dots[2 * q] += sr * (mi * cr);
            dots[2 * q + 1] += sr * (mi * r);
          }
        }
      }
    }
  }
  bool term_s = divergent;
  if (odd) {  // uniform across the grid: every CTA takes part in the reductions
    // reduce only the levels in use (nlev is grid-uniform)
    for (int q = 0; q < nlev; ++q) {
      double pr[2] = {dots[2 * q], dots[2 * q + 1]};
      tile_reduce<2>(pr, red);
      if (in && !divergent && ((pr[0] <= 0.0) || (pr[1] <= 0.0))) term_s = true;
    }
    if (in && term_s && !divergent && t.dl == 0) {
      s.term_s[t.c] = 1;
      s.sub_active[t.c] = 0;
    }
  }
  // pass 3: first half kick + drift of the next leaf
  if (in && !term_s && n + 1 < nleaf) {
    for (int d = t.dl; d < D; d += kTileD) {
      const size_t o = (size_t)d * C + t.c;
      const double r = ls.zr[o] + half * ls.zg[o];
      ls.zr[o] = r;
      ls.zt[o] = ls.zt[o] + se * (st.minv[o] * r);
    }
  }
}

// ------------------------------------------------------------------------------------------
// doubling end: combine the new sub-tree into the tree, biased progressive sampling,
// tree-level U-turn, depth cap; counts the chains whose transition continues.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTileThreads)
k_ls_doubling_end(ChainState st, LockstepState ls, RunParams rp, int j, int *n_active) {
  __shared__ double red[2 * 64];
  const Tile t;
  const int C = st.C, D = st.D;
  TreeScalars &s = ls.ts;
  const bool in = t.c < C && s.active[t.c] != 0;
  const size_t DC = (size_t)D * C;
  bool take = false, term_s = false;
  int v = 0;
  if (in) {
    v = s.v[t.c];
    term_s = s.term_s[t.c] != 0;
    if (!term_s) {
      const uint32_t gchain = (uint32_t)(rp.chain_offset + t.c);
      const uint32_t it = (uint32_t)(st.iter[t.c] + 1);
      const double u = uniform1(rp.seed, gchain, it, SLOT_MH, (uint32_t)j);
      const double dl = s.lw_s[t.c] - s.lw[t.c];
      take = u < (dl < 0.0 ? exp(dl) : 1.0);
    }
  }
  double dots[2] = {0.0, 0.0};
  if (in) {
    double *e = ls.edge[v];
    const double *eo = ls.edge[1 - v];
    for (int d = t.dl; d < D; d += kTileD) {
      const size_t o = (size_t)d * C + t.c;
      const double r = ls.zr[o];
      e[o] = ls.zt[o];
      e[DC + o] = r;
      e[2 * DC + o] = ls.zg[o];
      if (take) {
        ls.thC[o] = ls.thS[o];
        ls.gC[o] = ls.gS[o];
      }
      const double rho = ls.rho[o] + ls.rho_s[o];
      ls.rho[o] = rho;
      const double mi = st.minv[o];
      const double ro = eo[DC + o];
      // left edge is edge[0], right edge is edge[1]
      dots[v] += rho * (mi * r);
      dots[1 - v] += rho * (mi * ro);
    }
  }
  tile_reduce<2>(dots, red);
  if (in && t.dl == 0) {
    s.elp[v * C + t.c] = s.zlp[t.c];
    int depth = s.depth[t.c];
    if (!term_s) {
      depth += 1;
      s.depth[t.c] = depth;
      if (take) {
        s.lpC[t.c] = s.lpS[t.c];
        s.HC[t.c] = s.HS[t.c];
      }
    }
    s.lw[t.c] = ls_logaddexp(s.lw[t.c], s.lw_s[t.c]);
    const bool uturn = (dots[0] <= 0.0) || (dots[1] <= 0.0);
    const bool cont = !(term_s || uturn) && depth < rp.max_depth;
    s.active[t.c] = cont ? 1 : 0;
    if (cont) atomicAdd(n_active, 1);
  }
}

// ------------------------------------------------------------------------------------------
// transition end: stats, move to the candidate, AHMC.adapt! (dual averaging, Welford,
// window ends), and -- for kept transitions -- constrained parameters / log-probs / stats
// ------------------------------------------------------------------------------------------
struct RecordInfo {
  long long k;  // draw slot, or -1 when the transition is not kept
};

// family-specific user-space pieces ------------------------------------------------------------
// positive-support parameters (linked by log) per family
__device__ __forceinline__ bool ls_is_positive(const ModelDev &m, int d) {
  switch (m.family) {
    case TNUTS_GDEMO: return d == 0;
    case TNUTS_LINREG: return d == 2;
    case TNUTS_EIGHT_SCHOOLS: return d == 1;
    case TNUTS_HIER_LINREG: return d >= 2 && d <= 4;
    default: return false;
  }
}

// log prior (user space) and log|det J| at theta column c; small families also return the
// likelihood directly (lik_direct = true); the others derive it as lp - prior - logjac.
// Called by the dl == 0 thread for the small families; the D-parallel sums for logistic /
// hierarchical are passed in (q0..q2).
__device__ void ls_prior_small(const ModelDev &m, const double *th, int C, int c, double &prior,
                               double &lik) {
  auto T = [&](int d) { return th[(size_t)d * C + c]; };
  switch (m.family) {
    case TNUTS_GDEMO: {
      const double al = m.hyper[0], be = m.hyper[1], a = T(0), mm = T(1), s = exp(a);
      prior = al * log(be) - lgamma(al) - (al + 1.0) * a - be / s - 0.5 * kLog2Pi - 0.5 * a -
              mm * mm / (2.0 * s);
      double q = 0;
      for (long long i = 0; i < m.N; ++i) {
        const double d = m.y[i] - mm;
        q += d * d;
      }
      lik = -0.5 * m.N * kLog2Pi - 0.5 * m.N * a - q / (2.0 * s);
    } break;
    case TNUTS_LINREG: {
      const double sd = m.hyper[0], cc = m.hyper[1];
      const double v = exp(T(2)), w = (v / cc) * (v / cc);
      prior = -kLog2Pi - 2.0 * log(sd) - 0.5 * (T(0) * T(0) + T(1) * T(1)) / (sd * sd) + kLog2 -
              kLogPi - log(cc) - log1p(w);
      double ssr = 0;
      for (long long i = 0; i < m.N; ++i) {
        const double res = m.y[i] - T(0) - T(1) * m.X[i];
        ssr += res * res;
      }
      lik = -0.5 * m.N * (kLog2Pi + T(2)) - ssr / (2.0 * v);
    } break;
    case TNUTS_EIGHT_SCHOOLS: {
      const double sd = m.hyper[0], cc = m.hyper[1];
      const double mu = T(0), tau = exp(T(1)), w = (tau / cc) * (tau / cc);
      prior = -0.5 * kLog2Pi - log(sd) - 0.5 * mu * mu / (sd * sd) + kLog2 - kLogPi - log(cc) -
              log1p(w);
      lik = 0;
      for (int jn = 0; jn < (int)m.N; ++jn) {
        const double tj = T(2 + jn), sg = m.sigma[jn];
        const double res = m.y[jn] - mu - tau * tj;
        prior += -0.5 * kLog2Pi - 0.5 * tj * tj;
        lik += -0.5 * kLog2Pi - log(sg) - 0.5 * res * res / (sg * sg);
      }
    } break;
    case TNUTS_STD_NORMAL: {
      prior = -0.5 * m.D * kLog2Pi;
      for (int d = 0; d < m.D; ++d) prior -= 0.5 * T(d) * T(d);
      lik = 0.0;
    } break;
    default:
      prior = lik = 0.0;
  }
}

// tile-level user-space log-probs for any family; th = [D][C] column c; lp = linked log joint
__device__ __forceinline__ void ls_user_logp(const ModelDev &m, const double *th, int C, const Tile &t,
                                             bool in, double lp, double *red, double &prior,
                                             double &lik) {
  prior = lik = 0.0;
  if (m.family == TNUTS_LOGISTIC) {
    double q[1] = {0.0};
    if (in)
      for (int d = t.dl; d < m.D; d += kTileD) {
        const double x = th[(size_t)d * C + t.c];
        q[0] += x * x;
      }
    tile_reduce<1>(q, red);
    const double sd = m.hyper[0];
    prior = -0.5 * m.D * kLog2Pi - m.D * log(sd) - 0.5 * q[0] / (sd * sd);
    lik = lp - prior;
  } else if (m.family == TNUTS_HIER_LINREG) {
    const int G = m.G;
    double q[2] = {0.0, 0.0};
    double mua = 0, mub = 0;
    if (in) {
      mua = th[t.c];
      mub = th[(size_t)C + t.c];
      for (int g = t.dl; g < G; g += kTileD) {
        const double ea = th[(size_t)(5 + g) * C + t.c] - mua;
        const double eb = th[(size_t)(5 + G + g) * C + t.c] - mub;
        q[0] += ea * ea;
        q[1] += eb * eb;
      }
    }
    tile_reduce<2>(q, red);
    if (in) {
      const double sdm = m.hyper[0], s0 = m.hyper[1];
      const double la = th[(size_t)2 * C + t.c], lb = th[(size_t)3 * C + t.c], ly = th[(size_t)4 * C + t.c];
      const double sa = exp(la), sb = exp(lb), sy = exp(ly);
      prior = 2.0 * (-0.5 * kLog2Pi - log(sdm)) - 0.5 * (mua * mua + mub * mub) / (sdm * sdm) +
              3.0 * (kLog2 - 0.5 * kLog2Pi - log(s0)) - 0.5 * (sa * sa + sb * sb + sy * sy) / (s0 * s0) -
              (double)G * kLog2Pi - G * la - G * lb - 0.5 * q[0] / (sa * sa) - 0.5 * q[1] / (sb * sb);
      lik = lp - prior - (la + lb + ly);
    }
  } else {
    if (in && t.dl == 0) ls_prior_small(m, th, C, t.c, prior, lik);
  }
}

__global__ void __launch_bounds__(kTileThreads)
k_ls_transition_end(ChainState st, LockstepState ls, RunParams rp, ModelDev m, DrawBuffers out,
                    long long k_slot) {
  __shared__ double red[2 * 64];
  const Tile t;
  const int C = st.C, D = st.D;
  TreeScalars &s = ls.ts;
  const bool in = t.c < C && st.status[t.c] == 0;
  long long iter = 0, wf_n = 0;
  bool adapt = false, in_window = false, wend = false;
  double eps_used = 0.0;
  if (in) {
    iter = st.iter[t.c] + 1;
    wf_n = st.wf_n[t.c];
    eps_used = st.eps[t.c];
    adapt = rp.algorithm != TNUTS_ALG_HMC && iter <= rp.n_adapts;
    if (adapt) {
      for (int q = 0; q < rp.n_splits; ++q) wend = wend || (rp.splits[q] == iter);
      in_window = iter >= rp.window_start && iter <= rp.window_end;
    }
  }
  // move to the candidate; Welford push; mass-matrix refresh at a window end
  if (in) {
    const double n = (double)(wf_n + 1);
    for (int d = t.dl; d < D; d += kTileD) {
      const size_t o = (size_t)d * C + t.c;
      const double th = ls.thC[o];
      st.theta[o] = th;
      st.grad[o] = ls.gC[o];
      if (in_window) {
        double mean = st.wf_mean[o], m2 = st.wf_m2[o];
        const double delta = th - mean;
        mean += delta / n;
        m2 += delta * (th - mean);
        if (wend) {
          if (wf_n + 1 >= 10)
            st.minv[o] = n / ((n + 5.0) * (n - 1.0)) * m2 + 1e-3 * (5.0 / (n + 5.0));
          mean = 0.0;
          m2 = 0.0;
        }
        st.wf_mean[o] = mean;
        st.wf_m2[o] = m2;
      }
    }
  }
  double stats[TNUTS_NSTAT];
  if (in) {
    const double na = (double)s.n_a[t.c];
    stats[ST_NSTEPS] = na;
    stats[ST_ACC] = s.sum_a[t.c] / na;
    stats[ST_LP] = s.lpC[t.c];
    stats[ST_H] = s.HC[t.c];
    stats[ST_HERR] = s.HC[t.c] - s.H0[t.c];
    stats[ST_MAXHERR] = s.dHmax[t.c];
    stats[ST_DEPTH] = (double)s.depth[t.c];
    stats[ST_NUMERR] = (double)s.numerr[t.c];
    stats[ST_EPS] = eps_used;
  }
  // kept draw
  if (k_slot >= 0) {
    double prior, lik;
    ls_user_logp(m, ls.thC, C, t, in, in ? s.lpC[t.c] : 0.0, red, prior, lik);
    if (in) {
      double *row = out.draws + (size_t)k_slot * out.ncol * C + t.c;
      for (int d = t.dl; d < D; d += kTileD) {
        const double th = ls.thC[(size_t)d * C + t.c];
        row[(size_t)d * C] = ls_is_positive(m, d) ? exp(th) : th;
        if (out.theta) out.theta[((size_t)k_slot * D + d) * C + t.c] = th;
      }
      if (t.dl == 0) {
        row[(size_t)(D + 0) * C] = prior;
        row[(size_t)(D + 1) * C] = lik;
        row[(size_t)(D + 2) * C] = prior + lik;
#pragma unroll
        for (int q = 0; q < TNUTS_NSTAT; ++q) row[(size_t)(D + 3 + q) * C] = stats[q];
      }
    }
  }
  if (in && t.dl == 0) {
    st.lp[t.c] = s.lpC[t.c];
    st.iter[t.c] = iter;
    if (adapt) {
      double da_mu = st.da_mu[t.c], da_xbar = st.da_xbar[t.c], da_hbar = st.da_hbar[t.c],
             da_eps = st.da_eps[t.c];
      long long da_m = st.da_m[t.c];
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
      long long wn = in_window ? wf_n + 1 : wf_n;
      if (wend) {
        da_mu = log(10.0 * da_eps);
        da_m = 0;
        da_xbar = 0.0;
        da_hbar = 0.0;
        wn = 0;
      }
      if (iter == rp.n_adapts) da_eps = exp(da_xbar);
      st.da_mu[t.c] = da_mu;
      st.da_xbar[t.c] = da_xbar;
      st.da_hbar[t.c] = da_hbar;
      st.da_eps[t.c] = da_eps;
      st.da_m[t.c] = da_m;
      st.wf_n[t.c] = wn;
      st.eps[t.c] = da_eps;
    }
  }
}

// ------------------------------------------------------------------------------------------
// compaction of the active chains for the gradient kernels (ordered, deterministic)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_ls_compact(const int *flag, int C, int *list, int *count) {
  __shared__ int wsum[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c0 = 0; c0 < C; c0 += 1024) {
    const int c = c0 + threadIdx.x;
    const int f = (c < C && flag[c] != 0) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int pre = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    if (f) list[off + pre] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += wsum[w];
      base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base;
}

}  // namespace tnuts
