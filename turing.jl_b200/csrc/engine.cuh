// engine.cuh -- host-side engine object behind the C ABI (include/tnuts.h).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "models.cuh"

struct ncclComm;

namespace tnuts {

struct LockstepState;  // lockstep.cu
struct DiagWork;       // diagnostics.cu

enum Strategy { STRAT_AUTO = 0, STRAT_THREAD = 1, STRAT_LOCKSTEP = 2, STRAT_BLOCK = 3 };

struct Engine {
  tnuts_config cfg{};
  std::string err;
  int device = 0;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool has_model = false, inited = false;
  int strategy = STRAT_THREAD;

  // model
  ModelDev model{};        // device pointers
  SmallData small{};       // constant-bank copy for tiny families
  std::vector<void *> model_allocs;
  int D = 0, P = 0;

  // chain state
  ChainState st{};
  std::vector<void *> state_allocs;
  RunParams rp{};

  // kept draws
  DrawBuffers out{};
  long long n_kept = 0;

  // counters
  long long launches = 0;
  double last_run_seconds = 0.0, last_init_seconds = 0.0;
  bool keep_theta = false;
  // lane -> chain permutation of the thread-per-chain strategy (chains sorted by step size)
  int *perm = nullptr;
  double *perm_keys = nullptr;
  void *perm_tmp = nullptr;
  size_t perm_tmp_bytes = 0;
  bool perm_valid = false, sort_chains = true;
  long long host_iter = 0;  // transitions done since init (identical for every chain)
  bool dyn_sched = true;   // thread-per-chain: task queue over (transition sub-range, chain group)
  int *dyn_work = nullptr; // [1 + groups]: task counter, per-group completed sub-ranges
  bool ck_smem = true;     // thread-per-chain NUTS: lowest checkpoint levels in shared memory
  int occupancy = 2;  // CTAs of 128 threads per SM the thread-per-chain kernel is compiled for
  // logistic regression: likelihood gradient on the tcgen05 tensor cores (bf16x3 planes, fp32-grade
  // dot products, logistic_tc.cuh) instead of the fp64 DMMA kernel; opt-in
  bool logistic_tc = false;
  // measurement aid (option "time_kernels"): CUDA-event time of the dominant kernel's launches inside
  // tnuts_run (lock-step strategy: the family's batched gradient kernel), a systematic sample
  bool time_kernels = false;
  double kernel_seconds = 0.0;      // estimated total over the last run (sampled sum x stride)
  long long kernel_launches = 0;    // gradient-kernel launches of the last run
  long long kernel_sampled = 0;     // launches actually bracketed by events
  long long kernel_tile_launches = 0;  // sum over launches of 128-chain tiles (tcgen05 path)
  // the tail of a lock-step run as one persistent cooperative kernel (tc_persist.cuh); option "persist"
  bool persist = true;
  // lock-step runs gather the unfinished chains into narrower arrays as the others finish (lockstep.cu);
  // option "compact"
  bool compact = true;
  double persist_seconds = 0.0;     // CUDA-event time of the persistent launches of the last run
  long long persist_pumps = 0;      // pumps (leapfrogs of every unfinished chain) they executed

  LockstepState *ls = nullptr;
  DiagWork *diag = nullptr;

  // multi-GPU
  ncclComm *comm = nullptr;
  int world = 1, rank = 0;
};

// error helpers ---------------------------------------------------------------------------
int set_error(Engine *e, int code, const std::string &msg);
#define TN_CUDA(e, call)                                                                      \
  do {                                                                                        \
    cudaError_t _s = (call);                                                                  \
    if (_s != cudaSuccess)                                                                    \
      return ::tnuts::set_error((e), TNUTS_E_CUDA,                                            \
                                std::string(#call) + ": " + cudaGetErrorString(_s));          \
  } while (0)

// small-family (thread-per-chain) entry points, chain_kernels.cu
bool thread_strategy_supported(const Engine *e);
int thread_init(Engine *e, const double *theta0_dev);
int thread_run(Engine *e);
int thread_sort_chains(Engine *e);
int thread_point_lpgrad(Engine *e, const double *theta_dev, long long n, double *lp_dev,
                        double *grad_dev);
int thread_point_constrain(Engine *e, const double *theta_dev, long long n, double *params_dev,
                           double *logp_dev);
int thread_point_leapfrog(Engine *e, const double *theta_dev, const double *r_dev,
                          const double *minv_dev, double eps, int L, long long n, double *tt_dev,
                          double *tr_dev, double *tH_dev);

// one-CTA-per-chain strategy, block.cu
bool block_strategy_supported(const Engine *e);
int block_run(Engine *e);

// lock-step strategy, lockstep.cu
int lockstep_create(Engine *e);
void lockstep_destroy(Engine *e);
int lockstep_init(Engine *e, const double *theta0_dev);
int lockstep_run(Engine *e);
int lockstep_bench_gradient(Engine *e, int reps, double *seconds);
int lockstep_point_lpgrad(Engine *e, const double *theta_dev, long long n, double *lp_dev,
                          double *grad_dev);
int lockstep_point_constrain(Engine *e, const double *theta_dev, long long n, double *params_dev,
                             double *logp_dev);
int lockstep_point_leapfrog(Engine *e, const double *theta_dev, const double *r_dev,
                            const double *minv_dev, double eps, int L, long long n,
                            double *tt_dev, double *tr_dev, double *tH_dev);

// comm.cu
int comm_allreduce_sum(Engine *e, double *buf, size_t n);
void comm_destroy(Engine *e);

// diagnostics.cu
int diagnostics_run(Engine *e, double *rhat_host, double *ess_host);
void diagnostics_destroy(Engine *e);

}  // namespace tnuts
