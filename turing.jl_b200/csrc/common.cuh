// common.cuh -- shared device-side definitions of the tnuts engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/tnuts.h"

namespace tnuts {

constexpr double kLog2Pi = 1.8378770664093454835606594728112;
constexpr double kLog2 = 0.69314718055994530941723212145818;
constexpr double kLogPi = 1.1447298858494001741434273513531;

// stats columns (include/tnuts.h TNUTS_NSTAT)
enum { ST_NSTEPS, ST_ACC, ST_LP, ST_H, ST_HERR, ST_MAXHERR, ST_DEPTH, ST_NUMERR, ST_EPS };

// RNG slots: the addressing spec shared (by specification, not by code) with the oracle
enum {
  SLOT_INIT = 0,
  SLOT_MOMENTUM = 1,
  SLOT_DIRECTION = 2,
  SLOT_MH = 3,
  SLOT_LEAF = 4,
  SLOT_MERGE = 5,  // unused by the engine (AdvancedHMC-style merge sampling lives in the oracle)
  SLOT_EPS_MOMENTUM = 6
};

// Device view of a registered model.  All pointers are device pointers.
struct ModelDev {
  int family;
  int D;
  int G;
  long long N;
  const double *X;      // LOGISTIC: row-major [N][D] ; LINREG/HIER: x[N]
  const double *y;      // [N]
  const double *sigma;  // EIGHT_SCHOOLS [N]
  const int *group;     // HIER [N]
  double hyper[8];
};

// Per-chain persistent sampler state, structure-of-arrays with the chain index fastest
// (element (d, c) of a [D][C] array is at d*C + c) so that a warp of chains coalesces.
struct ChainState {
  int C;          // chains on this engine
  int D;
  double *theta;  // [D][C]  current position (linked space)
  double *grad;   // [D][C]  d log pi / d theta at theta
  double *lp;     // [C]
  double *minv;   // [D][C]  diagonal inverse metric
  double *eps;    // [C]     step size the next transition uses
  // Nesterov dual averaging
  double *da_mu, *da_xbar, *da_hbar, *da_eps;  // [C]
  long long *da_m;                              // [C]
  // Welford
  long long *wf_n;           // [C]
  double *wf_mean, *wf_m2;   // [D][C]
  // DenseEuclideanMetric (TNUTS_METRIC_DENSE, thread-per-chain strategy): M^-1, its upper Cholesky
  // factor and the Welford covariance accumulator, [D*D][C]; null otherwise
  double *minv_dense, *chol, *wf_cov;
  long long *iter;           // [C] iteration counter i (src/mcmc/hmc.jl:228)
  long long *n_grad;         // [C] gradient evaluations
  int *init_tries;           // [C]
  int *status;               // [C] 0 ok, 1 init failed
};

// Kept-draw output buffer on the device: ONE array [keep][ncol][C] whose columns are
// P constrained parameters | logprior, loglikelihood, logjoint | NSTAT sampler stats --
// exactly the (iterations x names x chains) array MCMCChains wants, so the host fetch is a
// single contiguous copy.  theta (linked-space draws) is optional.
struct DrawBuffers {
  long long capacity;  // draws
  int P;
  int ncol;        // P + 3 + TNUTS_NSTAT
  double *draws;   // [keep][ncol][C]
  double *theta;   // [keep][D][C] or nullptr
  // Compacted state (lockstep.cu, "progressive compaction"): the kernel works on a narrower copy of
  // the per-chain arrays that holds only the unfinished chains; column k of that copy is chain gid[k]
  // of the engine (its Philox address and its column of the draw arrays, which stay Cout wide).
  // gid == nullptr: the state arrays are the engine's own, column = chain.
  const int *gid = nullptr;
  int Cout = 0;
};

struct RunParams {
  int algorithm;
  double delta;
  int max_depth;
  double delta_max;
  long long n_adapts;
  int n_leapfrog;
  double lambda;
  unsigned long long seed;
  int chain_offset;
  // window schedule (1-based), computed on the host exactly once per engine
  long long window_start, window_end;
  int n_splits;
  long long splits[24];
  // this call
  long long n_transitions, first_keep, thin;
  int unit_metric; // TNUTS_METRIC_UNIT: the adaptation windows run, the mass matrix stays the identity
};

}  // namespace tnuts
