// lockstep.cuh -- state of the lock-step execution strategy (big-D families).
//
// All chains take the same leapfrog at the same time: one gradient launch serves every
// chain (that is where the X·B contraction of logistic regression becomes a dense GEMM), the
// streaming work is laid out [D][C] with the chain index fastest, and the NUTS tree
// bookkeeping is per-chain scalar state, masked so that finished chains idle.
#pragma once
#include "common.cuh"

namespace tnuts {

constexpr int kTileC = 8;    // chains per CTA in the streaming kernels
constexpr int kTileD = 32;   // d-lanes per chain
constexpr int kTileThreads = kTileC * kTileD;
constexpr int kMaxLevels = 24;

struct TreeScalars {  // [C] each
  double *H0, *sum_a, *dHmax, *lw, *lw_s, *lpC, *HC, *lpS, *HS, *zlp, *zH;
  double *elp;        // [2][C] lp at the left / right edge
  long long *n_a;
  int *depth;         // j of the chain (completed doublings)
  int *active;        // transition still running
  int *sub_active;    // current doubling still running
  int *term_s;        // current doubling terminated (U-turn or divergence)
  int *numerr;        // divergence seen in this transition
  int *v;             // direction of the current doubling: 0 = left, 1 = right
  int *n;             // leaf index inside the current doubling
  int *phase;         // 0 start a transition, 1 leaf pending, 2 done with this run
  long long *iter0, *iter_end;  // transition window of the current run
};

struct LockstepState {
  int C = 0, D = 0, levels = 0;
  // moving state and tree storage, [D][C] each
  double *zt = nullptr, *zr = nullptr, *zg = nullptr;
  double *edge[2] = {nullptr, nullptr};  // [3][D][C]: theta | r | grad
  double *rho = nullptr, *rho_s = nullptr;
  double *thC = nullptr, *gC = nullptr, *thS = nullptr, *gS = nullptr;
  double *ck_r = nullptr, *ck_rho = nullptr;  // [levels][D][C]
  TreeScalars ts{};
  // active-chain list for the gradient kernels
  int *list = nullptr;      // [C]
  int *count = nullptr;     // [1] (+ [1] active transitions)
  int *h_count = nullptr;   // pinned [2]
  // gradient workspaces
  double *gpart = nullptr;  // logistic: [chunks][D+1][C]
  size_t gpart_elems = 0;
  int chunks = 0;
  // hierarchical model: per-group centred sufficient statistics [G] each
  double *hs_n = nullptr, *hs_xbar = nullptr, *hs_ybar = nullptr, *hs_sxx = nullptr,
         *hs_sxy = nullptr, *hs_syy = nullptr;
  // step-size search scratch [C]
  double *fs_eps = nullptr, *fs_epsp = nullptr, *fs_H = nullptr;
  int *fs_dir = nullptr, *fs_phase = nullptr, *fs_iter = nullptr;
  double *r0 = nullptr;  // [D][C] momentum of the search
  // DenseEuclideanMetric on this strategy (dense_pump.cuh; null otherwise)
  double *wT = nullptr, *wG = nullptr;  // [D][C] theta-space position handed to / gradient returned by the family's kernel
  int *dn_req = nullptr;                // [C] bit 0: Welford covariance push pending, bit 1: window end (new metric)
  int *dn_n = nullptr;                  // [C] Welford count including the pending sample
};

// the state the persistent pump kernel works on (persist.cu): the engine's own arrays or a compacted
// copy of them (lockstep.cu, progressive compaction)
struct PersistView {
  const ChainState *st;
  const LockstepState *ls;
  const DrawBuffers *out;
  int c_total;             // chains of the engine (the finished-chain counter runs up to this)
  int exit_below;          // return for a re-compaction once the unfinished chains drop to this many (0: never)
};

}  // namespace tnuts
