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

// Memory-level parallelism of the state machine's passes over d.  Every pass reads a few [D][C]
// arrays at this lane's dimensions d = dl, dl + 32, ... and writes some of them back.  Written as one
// loop ("load, compute, store" per d) the compiler keeps a store between the loads of consecutive
// iterations (same array, addresses it cannot tell apart) and sinks every load next to its use: a pass
// became 8 - 12 dependent L2 round trips, and the whole pump was bound by that latency (k_ls_pump
// 0.11 ms for 50 MB of traffic; 20 - 40 us per leapfrog of a single chain in the persistent kernel).
// The passes therefore load a batch of kNB dimensions of all their arrays first (index clamped to
// D - 1 for the lanes past the end, whose results are masked), gate the first use on the last loads
// having arrived (ls_gate), and only then compute and store.  Same operations per element in the same
// order: results are unchanged.
constexpr int kNB = 4;
// x, once `last` (the value of a batch's last load) has arrived; the bit pattern compared with is a
// signalling NaN that no computation produces
__device__ __forceinline__ double ls_gate(double x, double last) {
  return __double_as_longlong(last) == 0x7ff4dead0000beefLL ? last : x;
}
#define LS_BATCH_OFFSETS(o, d0)                                                                    \
  size_t o[kNB];                                                                                   \
  _Pragma("unroll") for (int i_ = 0; i_ < kNB; ++i_) o[i_] = (size_t)min((d0) + i_ * kTileD, D - 1) * C + t.c
#define LS_BATCH_LOAD(dst, src, o)                                                                 \
  double dst[kNB];                                                                                 \
  _Pragma("unroll") for (int i_ = 0; i_ < kNB; ++i_) dst[i_] = (src)[o[i_]]

// A "team" is the 256 threads that own one tile of 8 chains.  In the stand-alone kernels it is the
// whole CTA; inside the persistent pump kernel (tc_persist.cuh) it is 8 warps of a larger CTA that
// synchronise on a named barrier.  tid() is the thread's index inside its team.
struct BlockTeam {
  static constexpr bool kWarp = false;
  static constexpr bool kContiguous = true;   // the tile's 8 chains are consecutive (64 B per d)
  __device__ static __forceinline__ int tid() { return threadIdx.x; }
  __device__ static __forceinline__ void sync() { __syncthreads(); }
  __device__ static __forceinline__ int sync_or(int p) { return __syncthreads_or(p); }
  __device__ static __forceinline__ int sync_and(int p) { return __syncthreads_and(p); }
};
template <int BAR, int TID0>
struct NamedTeam {
  static constexpr bool kWarp = false;
  static constexpr bool kContiguous = false;  // chains picked from the work list
  __device__ static __forceinline__ int tid() { return (int)threadIdx.x - TID0; }
  __device__ static __forceinline__ void sync() {
    asm volatile("bar.sync %0, %1;" ::"r"(BAR), "r"(kTileThreads) : "memory");
  }
  __device__ static __forceinline__ int sync_or(int p) {
    int r;
    asm volatile(
        "{\n\t.reg .pred q, r;\n\t"
        "setp.ne.s32 q, %1, 0;\n\t"
        "bar.red.or.pred r, %2, %3, q;\n\t"
        "selp.s32 %0, 1, 0, r;\n\t}"
        : "=r"(r)
        : "r"(p), "r"(BAR), "r"(kTileThreads)
        : "memory");
    return r;
  }
  __device__ static __forceinline__ int sync_and(int p) {
    int r;
    asm volatile(
        "{\n\t.reg .pred q, r;\n\t"
        "setp.ne.s32 q, %1, 0;\n\t"
        "bar.red.and.pred r, %2, %3, q;\n\t"
        "selp.s32 %0, 1, 0, r;\n\t}"
        : "=r"(r)
        : "r"(p), "r"(BAR), "r"(kTileThreads)
        : "memory");
    return r;
  }
};

// One warp = ONE chain, lane = d-lane: the team of the persistent pump kernel's tail, where there are
// more warps than unfinished chains.  No CTA barrier anywhere (a chain only walks the sections of the
// state machine it needs, and never waits for seven neighbours), and the reductions over d are
// shuffles -- in the same association order as the 8-chain tile's, so the results are bit-identical.
struct WarpTeam {
  static constexpr bool kWarp = true;
  static constexpr bool kContiguous = false;
  __device__ static __forceinline__ int tid() { return (int)(threadIdx.x & 31); }
  __device__ static __forceinline__ void sync() { __syncwarp(); }
  __device__ static __forceinline__ int sync_or(int p) { return __any_sync(0xffffffffu, p); }
  __device__ static __forceinline__ int sync_and(int p) { return __all_sync(0xffffffffu, p); }
};

struct Tile {
  int c, cl, dl;
  __device__ Tile() {
    cl = threadIdx.x & (kTileC - 1);
    dl = threadIdx.x >> 3;
    c = blockIdx.x * kTileC + cl;
  }
  // thread `tid` of a team; its chain id is given (any chain: the tile need not be contiguous)
  __device__ Tile(int tid, int chain) {
    cl = tid & (kTileC - 1);
    dl = tid >> 3;
    c = chain;
  }
  // lane `lane` of a WarpTeam
  __device__ static Tile of_warp(int lane, int chain) {
    Tile t(0, chain);
    t.cl = 0;
    t.dl = lane;
    return t;
  }
};

// sum over the 32 d-lanes of each chain for NR values at once; result valid in every thread
template <int NR, class TM = BlockTeam>
__device__ __forceinline__ void tile_reduce(double (&v)[NR], double *smem /* [NR][8][8] */) {
  if constexpr (TM::kWarp) {
    // the tile's order: d-lanes {4w .. 4w+3} pairwise ((a0 + a1) + (a2 + a3)), then the eight groups
    // w = 0 .. 7 one after the other onto 0.0
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      double x = v[r];
      x += __shfl_xor_sync(0xffffffffu, x, 1);
      x += __shfl_xor_sync(0xffffffffu, x, 2);
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += __shfl_sync(0xffffffffu, x, 4 * w);
      v[r] = s;
    }
    return;
  }
  const int tid = TM::tid();
  const int lane = tid & 31, warp = tid >> 5, cl = tid & 7;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    v[r] += __shfl_xor_sync(0xffffffffu, v[r], 8);
    v[r] += __shfl_xor_sync(0xffffffffu, v[r], 16);
  }
  TM::sync();  // protect smem reuse between consecutive calls
  if (lane < 8) {
#pragma unroll
    for (int r = 0; r < NR; ++r) smem[(r * 8 + warp) * 8 + lane] = v[r];
  }
  TM::sync();
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += smem[(r * 8 + w) * 8 + cl];
    v[r] = s;
  }
}

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
static __device__ void ls_prior_small(const ModelDev &m, const double *th, int C, int c, double &prior,
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
template <class TM = BlockTeam>
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
    tile_reduce<1, TM>(q, red);
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
    tile_reduce<2, TM>(q, red);
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

// ------------------------------------------------------------------------------------------
// k_sg_step -- SGHMC / SGLD (src/mcmc/sghmc.jl:79-105, 223-249) on the batched gradient kernels:
// all chains are in lock step, one launch per transition.  The launch first records the current
// position as kept draw k_slot (its lp / gradient were just written by the family's gradient
// kernel), then moves every chain:
//   SGHMC  theta += v;  v = (1 - alpha) v + eta grad + sqrt(2 eta alpha) z     v in st.wf_mean
//   SGLD   eps = a / (t + b)^gamma;  theta += eps / 2 grad + sqrt(eps) z       t = iteration, from 1
// z from the MOMENTUM slot of the RNG addressing specification (the oracle does the same).
// ------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(kTileThreads)
k_sg_step(ChainState st, RunParams rp, ModelDev m, DrawBuffers out, long long k_slot, int do_update,
          double sg_a, double sg_b, double sg_c) {
  __shared__ double red[2 * 64];
  const Tile t;
  const int C = st.C, D = st.D;
  const bool in = t.c < C && st.status[t.c] == 0;
  if (k_slot >= 0) {  // launch-uniform
    double prior, lik;
    ls_user_logp(m, st.theta, C, t, in, in ? st.lp[t.c] : 0.0, red, prior, lik);
    if (in) {
      double *row = out.draws + (size_t)k_slot * out.ncol * C + t.c;
      for (int d = t.dl; d < D; d += kTileD) {
        const double th = st.theta[(size_t)d * C + t.c];
        row[(size_t)d * C] = ls_is_positive(m, d) ? exp(th) : th;
        if (out.theta) out.theta[((size_t)k_slot * D + d) * C + t.c] = th;
      }
      if (t.dl == 0) {
        row[(size_t)(D + 0) * C] = prior;
        row[(size_t)(D + 1) * C] = lik;
        row[(size_t)(D + 2) * C] = prior + lik;
#pragma unroll
        for (int q = 0; q < TNUTS_NSTAT; ++q) row[(size_t)(D + 3 + q) * C] = CUDART_NAN;
        row[(size_t)(D + 3 + ST_NSTEPS) * C] = 1.0;
        row[(size_t)(D + 3 + ST_LP) * C] = st.lp[t.c];
        row[(size_t)(D + 3 + ST_EPS) * C] = st.eps[t.c];
      }
    }
  }
  __syncthreads();  // the d-lane-0 thread reads the whole theta column for the user-space log-probs
  if (!do_update || !in) return;
  const long long iter = st.iter[t.c];
  const uint32_t it = (uint32_t)(iter + 1);
  const uint32_t gchain = (uint32_t)(rp.chain_offset + t.c);
  double eps;
  if (rp.algorithm == TNUTS_ALG_SGHMC) {
    const double eta = sg_a, alpha = sg_c, sd = sqrt(2.0 * eta * alpha);
    eps = eta;
    for (int d = t.dl; d < D; d += kTileD) {
      const size_t o = (size_t)d * C + t.c;
      double z0, z1;
      normal2(rp.seed, gchain, it, SLOT_MOMENTUM, (uint32_t)(d >> 1), z0, z1);
      const double v = st.wf_mean[o];
      st.theta[o] += v;
      st.wf_mean[o] = (1.0 - alpha) * v + eta * st.grad[o] + sd * ((d & 1) ? z1 : z0);
    }
  } else {
    eps = sg_a / pow((double)(iter + 1) + sg_b, sg_c);
    const double sd = sqrt(eps);
    for (int d = t.dl; d < D; d += kTileD) {
      const size_t o = (size_t)d * C + t.c;
      double z0, z1;
      normal2(rp.seed, gchain, it, SLOT_MOMENTUM, (uint32_t)(d >> 1), z0, z1);
      st.theta[o] += 0.5 * eps * st.grad[o] + sd * ((d & 1) ? z1 : z0);
    }
  }
  __syncthreads();  // every d-lane has read iter before it moves on
  if (t.dl == 0) {
    st.iter[t.c] = iter + 1;
    st.n_grad[t.c] += 1;
    st.eps[t.c] = eps;
  }
}

// ------------------------------------------------------------------------------------------
// k_ls_pump -- the whole per-chain NUTS state machine as ONE kernel that the host "pumps":
// every launch advances EVERY unfinished chain by exactly one leapfrog (the gradient at the
// drifted position was written to zg / zlp by the family's kernel just before).  Chains are
// NOT in lock step with each other: each carries its own (transition, doubling j, leaf n),
// so a chain whose tree ends starts its next transition in the same launch while its
// neighbours keep extending theirs -- no chain ever waits for the deepest tree of the batch.
//
//   phase 0: start a transition (momentum refresh, H0, tree init, doubling 0)
//   phase 1: a leaf is pending: [A] finish the leapfrog, leaf statistics, streaming
//            multinomial candidate, checkpointed sub-tree U-turn checks;
//            [B] at the end of a doubling: combine, biased progressive sampling, tree U-turn;
//            [T] at the end of a transition: stats, move to the candidate, AHMC.adapt!,
//                kept draw (constrained parameters, user-space log-probs, stats);
//            [S] start of the next transition;  [D] first half kick + drift of the next leaf.
//   phase 2: the chain has done all its transitions of this run.
// Per-chain scalar logic is computed redundantly by the chain's 32 threads (identical inputs
// => identical results); every reduction over d is executed CTA-uniformly.
// ------------------------------------------------------------------------------------------
// The per-chain state machine for one tile of 8 chains, executed by a team of 256 threads.
// `t` carries the thread's chain id (t.c >= st.C: no chain in this lane) and d-lane;
// `red` is [2 * 64] doubles and `s_maxlev_p` one int of shared memory private to the team.
//
// DENSE (DenseEuclideanMetric on this strategy, metricT of src/mcmc/hmc.jl:377-386): the "whitened
// pump".  With M^-1 = U'U (U upper triangular, st.chol) the dense-metric dynamics in (theta, r) IS the
// unit-metric dynamics in phi = U^-T theta, p = U r: the drift theta += eps M^-1 r becomes phi += eps p,
// the kick uses grad_phi = U grad_theta, the kinetic energy r'M^-1 r / 2 is p'p / 2, the generalised
// U-turn products rho . M^-1 r are rho_p . p, and the momentum draw r = U \ z is p = z -- the very
// normals the unit-metric state machine draws.  So this state machine runs UNCHANGED on (phi, p) with
// minv = 1; around it k_dense_post hands the family's gradient kernel theta = U'phi (ls.wT) and
// k_dense_grad turns its gradient (ls.wG) into grad_phi = U grad_theta (ls.zg).  What changes here:
// at the end of a transition the candidate is mapped back (theta = U' phi_C) for st.theta, the kept
// draw and the Welford covariance; the chain's phi-space state for the next transition lives in
// st.wf_m2 (phi; the diagonal Welford M2 is unused with a dense metric) and st.grad (grad_phi); the
// Welford covariance push and, at a window end, the new metric (regularised covariance, Cholesky,
// phi and grad_phi re-expressed in the new factor) are left to k_dense_post through ls.dn_req, and
// a chain whose window just ended pauses (phase 3) until that kernel has renewed its metric.
template <class TM, bool DENSE = false>
__device__ __forceinline__ void ls_pump_body(const Tile &t, ChainState st, LockstepState ls, RunParams rp,
                                             ModelDev m, DrawBuffers out, int *n_done, double *red,
                                             int *s_maxlev_p, long long *tprof = nullptr) {
  int &s_maxlev = *s_maxlev_p;
  // profiling aid of the persistent kernel's trace (tprof != nullptr for one warp only): ns per section
  long long tp_mark = 0;
  auto tp_lap = [&](int i) {
    if (tprof && TM::tid() == 0) {
      unsigned long long tnow;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tnow)::"memory");
      if (i >= 0) tprof[i] += (long long)tnow - tp_mark;
      tp_mark = (long long)tnow;
    }
  };
  tp_lap(-1);
  const int C = st.C, D = st.D;
  TreeScalars &s = ls.ts;
  const size_t DC = (size_t)D * C;
  // distinct arrays never alias: tell the compiler so that it can batch the loads of the
  // unrolled d-loops (memory-level parallelism is what these streaming passes live on)
  double *__restrict__ const zt_ = ls.zt, *__restrict__ const zr_ = ls.zr, *__restrict__ const zg_ = ls.zg;
  double *__restrict__ const rhos_ = ls.rho_s, *__restrict__ const rho_ = ls.rho;
  double *__restrict__ const thS_ = ls.thS, *__restrict__ const gS_ = ls.gS;
  double *__restrict__ const thC_ = ls.thC, *__restrict__ const gC_ = ls.gC;
  double *__restrict__ const ckr_ = ls.ck_r, *__restrict__ const ckrho_ = ls.ck_rho;
  // start state of the next transition: (theta, grad); DENSE: (phi, grad_phi) in (st.wf_m2, st.grad)
  double *__restrict__ const minv_ = st.minv, *__restrict__ const theta_ = DENSE ? st.wf_m2 : st.theta,
                              *__restrict__ const grad_ = st.grad;
  const bool in = t.c < C && st.status[t.c] == 0;
  int phase = in ? s.phase[t.c] : 2;
  // a CTA whose 8 chains have all finished has nothing to do (the tail of a run is mostly such CTAs)
  if (TM::sync_and(phase == 2)) return;
  // The kernel is a chain of dependent passes over this chain's columns with a block reduction
  // between them, and the gradient kernel that ran just before has pushed those columns out of L2
  // (ncu: long-scoreboard stall 21 per issue, L2 hit rate 45 %).  Start all the misses at once:
  // one prefetch per 32-byte sector (4 chains) of every array the common leaf path reads.
  if (TM::kContiguous) {
    if (phase == 1 && (t.cl & 3) == 0) {
      for (int d = t.dl; d < D; d += kTileD) {
        const size_t o = (size_t)d * C + t.c;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(zg_ + o));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(zr_ + o));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(minv_ + o));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(rhos_ + o));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(zt_ + o));
      }
    }
  }
  // engine-wide chain index: Philox address and column of the draw arrays (compacted state: gid)
  const int gcol = (out.gid && t.c < C) ? out.gid[t.c] : t.c;
  const int Cout = out.gid ? out.Cout : C;
  const uint32_t gchain = (uint32_t)(rp.chain_offset + gcol);
  // chain scalars (registers, redundantly in the chain's 32 threads)
  double eps = 0, H0 = 0, lw = 0, lw_s = 0, sum_a = 0, dHmax = 0, lpC = 0, HC = 0, lpS = 0, HS = 0,
         zlp = 0, elp0 = 0, elp1 = 0;
  long long n_a = 0, iter = 0;
  int v = 0, j = 0, n = 0, numerr = 0;
  if (phase != 2) {
    eps = st.eps[t.c];
    iter = st.iter[t.c];
  }
  if (phase == 1) {
    H0 = s.H0[t.c]; lw = s.lw[t.c]; lw_s = s.lw_s[t.c]; sum_a = s.sum_a[t.c]; dHmax = s.dHmax[t.c];
    lpC = s.lpC[t.c]; HC = s.HC[t.c]; lpS = s.lpS[t.c]; HS = s.HS[t.c]; zlp = s.zlp[t.c];
    elp0 = s.elp[t.c]; elp1 = s.elp[C + t.c]; n_a = s.n_a[t.c];
    v = s.v[t.c]; j = s.depth[t.c]; n = s.n[t.c]; numerr = s.numerr[t.c];
  }
  tp_lap(0);   // chain scalars
  const bool leaf = phase == 1;
  bool start_transition = phase == 0;
  uint32_t it = (uint32_t)(iter + 1);

  // ================= [A] leaf finish =================
  double se = v ? eps : -eps, half = 0.5 * se;
  double a1[2] = {0.0, 0.0};
  if (leaf) {
    for (int d0 = t.dl; d0 < D; d0 += kNB * kTileD) {
      LS_BATCH_OFFSETS(o, d0);
      LS_BATCH_LOAD(g4, zg_, o);
      LS_BATCH_LOAD(r4, zr_, o);
      LS_BATCH_LOAD(m4, minv_, o);
      g4[0] = ls_gate(ls_gate(ls_gate(g4[0], g4[kNB - 1]), r4[kNB - 1]), m4[kNB - 1]);
#pragma unroll
      for (int i = 0; i < kNB; ++i)
        if (d0 + i * kTileD < D) {
          const double g = g4[i];
          const double r = r4[i] + half * g;
          zr_[o[i]] = r;
          a1[0] += m4[i] * r * r;
          a1[1] += isfinite(g) ? 0.0 : 1.0;
        }
    }
  }
  tile_reduce<2, TM>(a1, red);
  tp_lap(1);   // [A] second half kick + kinetic energy
  const bool hmc = rp.algorithm != TNUTS_ALG_NUTS;
  bool accept = false, divergent = false, term_s = false, sub_done = false;
  bool trans_end = false;
  double zH = 0.0;
  if (leaf && hmc) {
    // ---- HMC / HMCDA: static trajectory of j (= L) leapfrogs, Metropolis accept at its end ----
    const double K = 0.5 * a1[0];
    const bool ok = isfinite(zlp) && a1[1] == 0.0 && isfinite(K);
    zH = ok ? (-zlp + K) : CUDART_INF;
    n += 1;
    n_a += 1;
    if (n >= j || !ok) {
      const double dH = zH - H0;
      const double alpha = exp(fmin(0.0, -dH));
      const double u = uniform1(rp.seed, gchain, it, SLOT_MH, 0u);
      accept = u < alpha;
      sum_a = alpha * (double)n_a;  // acceptance_rate column = alpha
      dHmax = dH;
      numerr = ok ? 0 : 1;
      if (accept) {
        for (int d0 = t.dl; d0 < D; d0 += kNB * kTileD) {
          LS_BATCH_OFFSETS(o, d0);
          LS_BATCH_LOAD(t4, zt_, o);
          LS_BATCH_LOAD(g4, zg_, o);
#pragma unroll
          for (int i = 0; i < kNB; ++i)
            if (d0 + i * kTileD < D) {
              thC_[o[i]] = t4[i];
              gC_[o[i]] = g4[i];
            }
        }
        lpC = zlp;
        HC = zH;
      }
      j = accept ? 1 : 0;  // the tree_depth column carries is_accept for HMC
      trans_end = true;
    }
  }
  if (leaf && !hmc) {
    const double K = 0.5 * a1[0];
    const bool ok = isfinite(zlp) && a1[1] == 0.0 && isfinite(K);
    zH = ok ? (-zlp + K) : CUDART_INF;
    const double dH = zH - H0;
    const double a = exp(fmin(0.0, -dH));
    const double lw_leaf = -dH;
    const double lw_new = ls_logaddexp(lw_s, lw_leaf);
    const double p = (lw_leaf == -CUDART_INF) ? 0.0 : exp(lw_leaf - lw_new);
    const double u = uniform1(rp.seed, gchain, it, SLOT_LEAF, (1u << j) + (uint32_t)n);
    accept = u < p;
    divergent = !(dH < rp.delta_max);
    sum_a += a;
    n_a += 1;
    if (fabs(dH) > fabs(dHmax)) dHmax = dH;
    lw_s = lw_new;
    if (accept) {
      lpS = zlp;
      HS = zH;
    }
    if (divergent) {
      numerr = 1;
      term_s = true;
    }
  }
  tp_lap(2);   // leaf weights
  // checkpoints / sub-tree U-turn partial dot products (leaf index n is per chain)
  const int idx_max = __popc((unsigned)n >> 1);
  const bool odd = (n & 1) != 0;
  const int ntrail = __ffs(~(unsigned)n) - 1;
  const int idx_min = idx_max - ntrail + 1;
  const int nlev = (leaf && !hmc && odd && !divergent) ? (idx_max - idx_min + 1) : 0;
  if (TM::tid() == 0) s_maxlev = 0;
  TM::sync();
  if (t.dl == 0 && nlev > 0) atomicMax(&s_maxlev, nlev);
  TM::sync();
  const int maxlev = s_maxlev;
  double dots[2 * kMaxLevels];
#pragma unroll
  for (int q = 0; q < 2 * kMaxLevels; ++q) dots[q] = 0.0;
  if (leaf && !hmc) {
    const bool ck_store = !divergent && !odd, ck_dots = !divergent && odd;
    for (int d0 = t.dl; d0 < D; d0 += kNB * kTileD) {
      LS_BATCH_OFFSETS(o, d0);
      LS_BATCH_LOAD(r4, zr_, o);
      LS_BATCH_LOAD(rs4, rhos_, o);
      double t4[kNB], g4[kNB], m4[kNB];
      if (accept) {
#pragma unroll
        for (int i = 0; i < kNB; ++i) {
          t4[i] = zt_[o[i]];
          g4[i] = zg_[o[i]];
        }
      }
      if (ck_dots) {
#pragma unroll
        for (int i = 0; i < kNB; ++i) m4[i] = minv_[o[i]];
      }
      r4[0] = ls_gate(ls_gate(r4[0], r4[kNB - 1]), rs4[kNB - 1]);
#pragma unroll
      for (int i = 0; i < kNB; ++i) {
        rs4[i] = rs4[i] + r4[i];
        if (d0 + i * kTileD < D) {
          rhos_[o[i]] = rs4[i];
          if (accept) {
            thS_[o[i]] = t4[i];
            gS_[o[i]] = g4[i];
          }
          if (ck_store) {
            ckr_[idx_max * DC + o[i]] = r4[i];
            ckrho_[idx_max * DC + o[i]] = rs4[i];
          }
        }
      }
      if (ck_dots) {
        for (int q = 0; q < nlev; ++q) {
          const size_t lv = (size_t)(idx_max - q) * DC;
          double cr4[kNB], ch4[kNB];
#pragma unroll
          for (int i = 0; i < kNB; ++i) {
            cr4[i] = ckr_[lv + o[i]];
            ch4[i] = ckrho_[lv + o[i]];
          }
          cr4[0] = ls_gate(ls_gate(cr4[0], cr4[kNB - 1]), ch4[kNB - 1]);
#pragma unroll
          for (int i = 0; i < kNB; ++i)
            if (d0 + i * kTileD < D) {
              const double cr = cr4[i];
              const double sr = rs4[i] - ch4[i] + cr;
              dots[2 * q] += sr * (m4[i] * cr);
              dots[2 * q + 1] += sr * (m4[i] * r4[i]);
            }
        }
      }
    }
  }
  for (int q = 0; q < maxlev; ++q) {  // CTA-uniform trip count
    double pr[2] = {dots[2 * q], dots[2 * q + 1]};
    tile_reduce<2, TM>(pr, red);
    if (q < nlev && ((pr[0] <= 0.0) || (pr[1] <= 0.0))) term_s = true;
  }
  if (leaf && !hmc) {
    sub_done = term_s || (n + 1 == (1 << j));
    if (!sub_done) n += 1;
  }

  tp_lap(3);   // checkpoints + sub-tree U-turns
  // ================= [B] end of a doubling =================
  if (TM::sync_or(leaf && sub_done)) {
    const bool me = leaf && sub_done;
    bool take = false;
    if (me && !term_s) {
      const double u = uniform1(rp.seed, gchain, it, SLOT_MH, (uint32_t)j);
      const double dl = lw_s - lw;
      take = u < (dl < 0.0 ? exp(dl) : 1.0);
    }
    double d2[2] = {0.0, 0.0};
    if (me) {
      double *e = ls.edge[v];
      const double *eo = ls.edge[1 - v];
      for (int d0 = t.dl; d0 < D; d0 += kNB * kTileD) {
        LS_BATCH_OFFSETS(o, d0);
        LS_BATCH_LOAD(r4, zr_, o);
        LS_BATCH_LOAD(t4, zt_, o);
        LS_BATCH_LOAD(g4, zg_, o);
        LS_BATCH_LOAD(rh4, rho_, o);
        LS_BATCH_LOAD(rs4, rhos_, o);
        LS_BATCH_LOAD(m4, minv_, o);
        double eo4[kNB], ts4[kNB], gs4[kNB];
#pragma unroll
        for (int i = 0; i < kNB; ++i) eo4[i] = eo[DC + o[i]];
        if (take) {
#pragma unroll
          for (int i = 0; i < kNB; ++i) {
            ts4[i] = thS_[o[i]];
            gs4[i] = gS_[o[i]];
          }
        }
        r4[0] = ls_gate(ls_gate(ls_gate(r4[0], rh4[kNB - 1]), m4[kNB - 1]), eo4[kNB - 1]);
#pragma unroll
        for (int i = 0; i < kNB; ++i)
          if (d0 + i * kTileD < D) {
            const double r = r4[i];
            e[o[i]] = t4[i];
            e[DC + o[i]] = r;
            e[2 * DC + o[i]] = g4[i];
            if (take) {
              thC_[o[i]] = ts4[i];
              gC_[o[i]] = gs4[i];
            }
            const double rho = rh4[i] + rs4[i];
            rho_[o[i]] = rho;
            const double mi = m4[i];
            d2[v] += rho * (mi * r);
            d2[1 - v] += rho * (mi * eo4[i]);
          }
      }
    }
    tile_reduce<2, TM>(d2, red);
    if (me) {
      if (v) elp1 = zlp; else elp0 = zlp;
      if (!term_s) {
        j += 1;
        if (take) {
          lpC = lpS;
          HC = HS;
        }
      }
      lw = ls_logaddexp(lw, lw_s);
      const bool uturn = (d2[0] <= 0.0) || (d2[1] <= 0.0);
      const bool cont = !(term_s || uturn) && j < rp.max_depth;
      if (cont) {  // next doubling
        v = (uniform1(rp.seed, gchain, it, SLOT_DIRECTION, (uint32_t)j) < 0.5) ? 0 : 1;
        const double *en = ls.edge[v];
        for (int d0 = t.dl; d0 < D; d0 += kNB * kTileD) {
          LS_BATCH_OFFSETS(o, d0);
          double et4[kNB], er4[kNB], eg4[kNB];
#pragma unroll
          for (int i = 0; i < kNB; ++i) {
            et4[i] = en[o[i]];
            er4[i] = en[DC + o[i]];
            eg4[i] = en[2 * DC + o[i]];
          }
#pragma unroll
          for (int i = 0; i < kNB; ++i)
            if (d0 + i * kTileD < D) {
              zt_[o[i]] = et4[i];
              zr_[o[i]] = er4[i];
              zg_[o[i]] = eg4[i];
              rhos_[o[i]] = 0.0;
            }
        }
        zlp = v ? elp1 : elp0;
        lw_s = -CUDART_INF;
        n = 0;
      } else {
        trans_end = true;
      }
    }
  }

  tp_lap(4);   // [B]
  // ================= [T] end of a transition =================
  if (TM::sync_or(trans_end)) {
    long long k_slot = -1;
    bool adapt = false, in_window = false, wend = false;
    long long wf_n = 0;
    double stats[TNUTS_NSTAT];
    if (trans_end) {
      iter += 1;
      const long long tt = iter - s.iter0[t.c];  // 1-based transition number inside this run
      if (rp.first_keep >= 1 && tt >= rp.first_keep && (tt - rp.first_keep) % rp.thin == 0)
        k_slot = (tt - rp.first_keep) / rp.thin;
      wf_n = st.wf_n[t.c];
      adapt = rp.algorithm != TNUTS_ALG_HMC && iter <= rp.n_adapts;
      if (adapt) {
        for (int q = 0; q < rp.n_splits; ++q) wend = wend || (rp.splits[q] == iter);
        in_window = iter >= rp.window_start && iter <= rp.window_end;
      }
      if (DENSE && t.dl == 0) {   // Welford covariance / new metric: k_dense_post, right after this kernel
        ls.dn_req[t.c] = (in_window ? 1 : 0) | (wend ? 2 : 0);
        ls.dn_n[t.c] = (int)(wf_n + 1);
      }
      const double nw = (double)(wf_n + 1);
      for (int d0 = t.dl; d0 < D; d0 += kNB * kTileD) {
        LS_BATCH_OFFSETS(o, d0);
        LS_BATCH_LOAD(tc4, thC_, o);
        LS_BATCH_LOAD(gc4, gC_, o);
        double wm4[kNB], w24[kNB];
        if (!DENSE && in_window) {
#pragma unroll
          for (int i = 0; i < kNB; ++i) {
            wm4[i] = st.wf_mean[o[i]];
            w24[i] = st.wf_m2[o[i]];
          }
        }
#pragma unroll
        for (int i = 0; i < kNB; ++i)
          if (d0 + i * kTileD < D) {
            const double th = tc4[i];
            theta_[o[i]] = th;
            grad_[o[i]] = gc4[i];
            if (!DENSE && in_window) {
              double mean = wm4[i], m2 = w24[i];
              const double delta = th - mean;
              mean += delta / nw;
              m2 += delta * (th - mean);
              if (wend) {
                if (wf_n + 1 >= 10 && !rp.unit_metric)
                  minv_[o[i]] = nw / ((nw + 5.0) * (nw - 1.0)) * m2 + 1e-3 * (5.0 / (nw + 5.0));
                mean = 0.0;
                m2 = 0.0;
              }
              st.wf_mean[o[i]] = mean;
              st.wf_m2[o[i]] = m2;
            }
          }
      }
      const double na = (double)n_a;
      stats[ST_NSTEPS] = na;
      stats[ST_ACC] = sum_a / na;
      stats[ST_LP] = lpC;
      stats[ST_H] = HC;
      stats[ST_HERR] = HC - H0;
      stats[ST_MAXHERR] = dHmax;
      stats[ST_DEPTH] = (double)j;
      stats[ST_NUMERR] = (double)numerr;
      stats[ST_EPS] = eps;
    }
    // the position in theta space: the candidate itself, DENSE: theta = U' phi_C into st.theta
    const double *thU = thC_;
    if constexpr (DENSE) {
      TM::sync();   // phi_C may have been written by this chain's other lanes a moment ago ([B], HMC accept)
      if (trans_end) {
        const double *__restrict__ const U = st.chol;
        for (int d = t.dl; d < D; d += kTileD) {
          double acc0 = 0.0, acc1 = 0.0;   // theta_d = sum_{e <= d} U[e][d] phi_e
          int e = 0;
          for (; e + 1 <= d; e += 2) {
            acc0 = fma(U[((size_t)e * D + d) * C + t.c], thC_[(size_t)e * C + t.c], acc0);
            acc1 = fma(U[((size_t)(e + 1) * D + d) * C + t.c], thC_[(size_t)(e + 1) * C + t.c], acc1);
          }
          if (e <= d) acc0 = fma(U[((size_t)e * D + d) * C + t.c], thC_[(size_t)e * C + t.c], acc0);
          st.theta[(size_t)d * C + t.c] = acc0 + acc1;
        }
      }
      TM::sync();   // the record block below reads the whole theta column
      thU = st.theta;
    }
    if (TM::sync_or(trans_end && k_slot >= 0)) {
      const bool rec = trans_end && k_slot >= 0;
      double prior, lik;
      ls_user_logp<TM>(m, thU, C, t, rec, rec ? lpC : 0.0, red, prior, lik);
      if (rec) {
        double *row = out.draws + (size_t)k_slot * out.ncol * Cout + gcol;
        for (int d0 = t.dl; d0 < D; d0 += kNB * kTileD) {
          LS_BATCH_OFFSETS(o, d0);
          LS_BATCH_LOAD(tc4, thU, o);
#pragma unroll
          for (int i = 0; i < kNB; ++i) {
            const int d = d0 + i * kTileD;
            if (d < D) {
              const double th = tc4[i];
              row[(size_t)d * Cout] = ls_is_positive(m, d) ? exp(th) : th;
              if (out.theta) out.theta[((size_t)k_slot * D + d) * Cout + gcol] = th;
            }
          }
        }
        if (t.dl == 0) {
          row[(size_t)(D + 0) * Cout] = prior;
          row[(size_t)(D + 1) * Cout] = lik;
          row[(size_t)(D + 2) * Cout] = prior + lik;
#pragma unroll
          for (int q = 0; q < TNUTS_NSTAT; ++q) row[(size_t)(D + 3 + q) * Cout] = stats[q];
        }
      }
    }
    if (trans_end) {
      if (adapt) {  // Nesterov dual averaging + window bookkeeping (all 32 threads agree)
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
        if (isfinite(e)) {  // a non-finite step size reverts m, x_bar, H_bar and eps (AdvancedHMC adapt_stepsize!)
          da_m = mm;
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
        eps = da_eps;
        if (t.dl == 0) {
          st.da_mu[t.c] = da_mu;
          st.da_xbar[t.c] = da_xbar;
          st.da_hbar[t.c] = da_hbar;
          st.da_eps[t.c] = da_eps;
          st.da_m[t.c] = da_m;
          st.wf_n[t.c] = wn;
          st.eps[t.c] = da_eps;
        }
      }
      if (t.dl == 0) {
        st.lp[t.c] = lpC;
        st.iter[t.c] = iter;
      }
      if (iter >= s.iter_end[t.c]) {
        phase = 2;
        if (t.dl == 0) atomicAdd(n_done, 1);
      } else if (DENSE && wend) {
        phase = 3;   // wait for the new metric (k_dense_post sets phase 0)
      } else {
        start_transition = true;
        it = (uint32_t)(iter + 1);
      }
    }
  }

  // ================= [S] start of a transition =================
  if (TM::sync_or(start_transition)) {
    double acc[1] = {0.0};
    double lp0 = 0.0;
    if (start_transition) {
      lp0 = st.lp[t.c];
      if (trans_end) lp0 = lpC;  // just written by the dl == 0 thread of this chain
    }
    // One Philox / Box-Muller evaluation gives the normals of dimensions 2p and 2p + 1: the lane of
    // the even dimension evaluates it and hands z1 to its neighbour (the d-lane next to it: 1 lane
    // away in a one-warp team, 8 lanes away in a tile of 8 chains).  Every lane of the team runs the
    // loop (the shuffles need whole warps); lanes whose chain does not start a transition only pass.
    for (int b0 = 0; b0 < D; b0 += kNB * kTileD) {
      const int d0 = b0 + t.dl;
      LS_BATCH_OFFSETS(o, d0);
      double m4[kNB], th4[kNB], g4[kNB], z4[kNB];
      if (start_transition) {
#pragma unroll
        for (int i = 0; i < kNB; ++i) {
          m4[i] = minv_[o[i]];
          th4[i] = theta_[o[i]];
          g4[i] = grad_[o[i]];
        }
      }
#pragma unroll
      for (int i = 0; i < kNB; ++i) {
        const int d = d0 + i * kTileD;
        double z0 = 0.0, z1 = 0.0;
        if (start_transition && !(d & 1) && d < D)
          normal2(rp.seed, gchain, it, SLOT_MOMENTUM, (uint32_t)(d >> 1), z0, z1);
        const double zp = __shfl_xor_sync(0xffffffffu, z1, TM::kWarp ? 1 : kTileC);
        z4[i] = (d & 1) ? zp : z0;
      }
      if (start_transition) {
#pragma unroll
        for (int i = 0; i < kNB; ++i)
          if (d0 + i * kTileD < D) {
            const double mi = m4[i];
            const double r = z4[i] / sqrt(mi);
            const double th = th4[i], g = g4[i];
            acc[0] += mi * r * r;
            rho_[o[i]] = r;
            thC_[o[i]] = th;
            gC_[o[i]] = g;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              double *e = ls.edge[q];
              e[o[i]] = th;
              e[DC + o[i]] = r;
              e[2 * DC + o[i]] = g;
            }
            zt_[o[i]] = th;
            zr_[o[i]] = r;
            zg_[o[i]] = g;
            rhos_[o[i]] = 0.0;
          }
      }
    }
    tile_reduce<1, TM>(acc, red);
    if (start_transition) {
      const double K = 0.5 * acc[0];
      H0 = (isfinite(lp0) && isfinite(K)) ? (-lp0 + K) : CUDART_INF;
      sum_a = 0.0;
      n_a = 0;
      dHmax = 0.0;
      lw = 0.0;
      lpC = lp0;
      HC = H0;
      elp0 = elp1 = lp0;
      j = 0;
      numerr = 0;
      if (hmc) {
        v = 1;
        j = rp.n_leapfrog;
        if (rp.algorithm == TNUTS_ALG_HMCDA) {
          j = (int)floor(rp.lambda / eps);
          if (j < 1) j = 1;
        }
      } else {
        v = (uniform1(rp.seed, gchain, it, SLOT_DIRECTION, 0u) < 0.5) ? 0 : 1;
      }
      zlp = lp0;
      lw_s = -CUDART_INF;
      n = 0;
      phase = 1;
    }
  }

  tp_lap(5);   // [T] + [S]
  // ================= [D] first half kick + drift of the next leaf =================
  if (phase == 1) {
    se = v ? eps : -eps;
    half = 0.5 * se;
    for (int d0 = t.dl; d0 < D; d0 += kNB * kTileD) {
      LS_BATCH_OFFSETS(o, d0);
      LS_BATCH_LOAD(r4, zr_, o);
      LS_BATCH_LOAD(g4, zg_, o);
      LS_BATCH_LOAD(t4, zt_, o);
      LS_BATCH_LOAD(m4, minv_, o);
      r4[0] = ls_gate(ls_gate(ls_gate(r4[0], g4[kNB - 1]), t4[kNB - 1]), m4[kNB - 1]);
#pragma unroll
      for (int i = 0; i < kNB; ++i)
        if (d0 + i * kTileD < D) {
          const double r = r4[i] + half * g4[i];
          zr_[o[i]] = r;
          zt_[o[i]] = t4[i] + se * (m4[i] * r);
        }
    }
  }
  if (in && t.dl == 0) {
    s.phase[t.c] = phase;
    if (phase == 1) {
      s.H0[t.c] = H0; s.lw[t.c] = lw; s.lw_s[t.c] = lw_s; s.sum_a[t.c] = sum_a; s.dHmax[t.c] = dHmax;
      s.lpC[t.c] = lpC; s.HC[t.c] = HC; s.lpS[t.c] = lpS; s.HS[t.c] = HS; s.zlp[t.c] = zlp;
      s.elp[t.c] = elp0; s.elp[C + t.c] = elp1; s.n_a[t.c] = n_a;
      s.v[t.c] = v; s.depth[t.c] = j; s.n[t.c] = n; s.numerr[t.c] = numerr;
    }
    if (leaf) st.n_grad[t.c] += 1;
  }
  tp_lap(6);   // [D] + scalars out
}

static __global__ void __launch_bounds__(kTileThreads, 2)
k_ls_pump(ChainState st, LockstepState ls, RunParams rp, ModelDev m, DrawBuffers out, int *n_done,
          int *h_done /* mapped pinned host word */) {
  __shared__ double red[2 * 64];
  __shared__ int s_maxlev;
  const Tile t;
  if (blockIdx.x == 0 && threadIdx.x == 0) *(volatile int *)h_done = *(volatile int *)n_done;
  ls_pump_body<BlockTeam>(t, st, ls, rp, m, out, n_done, red, &s_maxlev);
}
static __global__ void __launch_bounds__(kTileThreads, 2)
k_ls_pump_dense(ChainState st, LockstepState ls, RunParams rp, ModelDev m, DrawBuffers out, int *n_done,
                int *h_done /* mapped pinned host word */) {
  __shared__ double red[2 * 64];
  __shared__ int s_maxlev;
  const Tile t;
  if (blockIdx.x == 0 && threadIdx.x == 0) *(volatile int *)h_done = *(volatile int *)n_done;
  ls_pump_body<BlockTeam, true>(t, st, ls, rp, m, out, n_done, red, &s_maxlev);
}

// arm a run: every chain gets its transition window [iter0, iter_end) and phase 0
static __global__ void k_ls_begin_run(ChainState st, LockstepState ls, long long n_transitions, int *n_done) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= st.C) return;
  const long long it0 = st.iter[c];
  ls.ts.iter0[c] = it0;
  ls.ts.iter_end[c] = it0 + n_transitions;
  const bool idle = st.status[c] != 0 || n_transitions <= 0;
  ls.ts.phase[c] = idle ? 2 : 0;
  if (idle) atomicAdd(n_done, 1);
}

// work-list flag for the gradient kernels: chains that still run
static __global__ void k_ls_running_flags(LockstepState ls, int C, int *flag) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) flag[c] = ls.ts.phase[c] != 2;
}

// ------------------------------------------------------------------------------------------
// compaction of the active chains for the gradient kernels (ordered, deterministic)
// ------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(1024) k_ls_compact(const int *flag, int C, int *list, int *count) {
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
