// chain_kernels.cu -- launchers of the one-thread-per-chain strategy (chain_kernels.cuh).
#include <cub/cub.cuh>

#include <algorithm>

#include "chain_kernels.cuh"
#include "engine.cuh"

namespace tnuts {

// family / dimension dispatch: every instantiation is a separate fully-unrolled kernel
#define TN_DISPATCH_SMALL(e, CALL)                                                     \
  do {                                                                                 \
    const int _f = (e)->model.family, _D = (e)->D;                                     \
    if (_f == TNUTS_GDEMO && _D == 2) { CALL(GDemo); }                                 \
    else if (_f == TNUTS_LINREG && _D == 3) { CALL(LinReg); }                          \
    else if (_f == TNUTS_EIGHT_SCHOOLS && _D == 10) { CALL(EightSchools<8>); }         \
    else if (_f == TNUTS_STD_NORMAL && _D == 1) { CALL(StdNormal<1>); }                \
    else if (_f == TNUTS_STD_NORMAL && _D == 2) { CALL(StdNormal<2>); }                \
    else if (_f == TNUTS_STD_NORMAL && _D == 4) { CALL(StdNormal<4>); }                \
    else if (_f == TNUTS_STD_NORMAL && _D == 10) { CALL(StdNormal<10>); }              \
    else if (_f == TNUTS_BETA_BERNOULLI && _D == 1) { CALL(BetaBernoulli); }           \
    else if (_f == TNUTS_DIRICHLET_CATEGORICAL && _D == 4 && (e)->P == 6) { CALL(DirichletCat24); } \
    else if (_f == TNUTS_DIRICHLET_CATEGORICAL && _D == 2 && (e)->P == 3) { CALL(DirichletCat30); } \
    else if (_f == TNUTS_DIRICHLET_CATEGORICAL && _D == 1 && (e)->P == 2) { CALL(DirichletCat20); } \
    else return set_error((e), TNUTS_E_UNSUPPORTED, "no thread-per-chain kernel for this family/D"); \
  } while (0)

bool thread_strategy_supported(const Engine *e) {
  const int f = e->model.family, D = e->D;
  const long long N = e->model.N;
  if (e->cfg.max_depth > kMaxCk + 1) return false;
  if (f == TNUTS_GDEMO) return D == 2 && N <= 16;
  if (f == TNUTS_LINREG) return D == 3 && N <= 16;
  if (f == TNUTS_EIGHT_SCHOOLS) return D == 10 && N == 8;
  if (f == TNUTS_STD_NORMAL) return D == 1 || D == 2 || D == 4 || D == 10;
  if (f == TNUTS_BETA_BERNOULLI) return D == 1;
  if (f == TNUTS_DIRICHLET_CATEGORICAL) return (D == 4 && e->P == 6) || (D == 2 && e->P == 3) || (D == 1 && e->P == 2);
  return false;
}

static inline int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }

__global__ void k_iota(int *p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

int thread_init(Engine *e, const double *theta0_dev) {
#define CALL(M)                                                                          \
  k_chain_init<M><<<grid_for(e->st.C, kChainBlock), kChainBlock, 0, e->stream>>>(        \
      e->small, e->st, e->rp, theta0_dev, e->cfg.init_eps)
  TN_DISPATCH_SMALL(e, CALL);
#undef CALL
  e->launches += 1;
  if (e->st.minv_dense) {  // DenseEuclideanMetric(D): identity (the step-size search above is the same)
    k_chain_init_dense<<<148 * 4, 256, 0, e->stream>>>(e->st);
    e->launches += 1;
  }
  TN_CUDA(e, cudaGetLastError());
  return TNUTS_OK;
}

static int thread_run_dense(Engine *e) {
#define CALL(M)                                                                              \
  do {                                                                                       \
    const int g_ = grid_for(e->st.C, kChainBlock);                                           \
    if (e->cfg.algorithm != TNUTS_ALG_NUTS)                                                  \
      k_chain_run_dense<M, 1, false><<<g_, kChainBlock, 0, e->stream>>>(e->small, e->st, e->out, e->rp); \
    else                                                                                     \
      k_chain_run_dense<M, 1, true><<<g_, kChainBlock, 0, e->stream>>>(e->small, e->st, e->out, e->rp);  \
  } while (0)
  TN_DISPATCH_SMALL(e, CALL);
#undef CALL
  e->launches += 1;
  TN_CUDA(e, cudaGetLastError());
  return TNUTS_OK;
}

int thread_run(Engine *e) {
  if (e->st.minv_dense) return thread_run_dense(e);
  int n_sm = 0;
  TN_CUDA(e, cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, e->device));
  const bool dyn_ok = e->dyn_sched;
#define CALL(M)                                                                              \
  do {                                                                                       \
    const int g_ = grid_for(e->st.C, kChainBlock);                                           \
    const int *pm_ = e->perm_valid ? e->perm : nullptr;                                      \
    const size_t sm_ = e->ck_smem ? chain_ck_smem_bytes<M::D>() : 0;                         \
    if (e->cfg.algorithm != TNUTS_ALG_NUTS)                                               \
      k_chain_run<M, 2, false><<<g_, kChainBlock, 0, e->stream>>>(e->small, e->st, e->out, e->rp, pm_, 0); \
    else if (dyn_ok) {                                                                       \
      int per_sm_ = 0;                                                                       \
      TN_CUDA(e, cudaFuncSetAttribute(k_chain_run_dyn<M, 2, true>,                           \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_ck_smem_bytes<M::D>())); \
      TN_CUDA(e, cudaFuncSetAttribute(k_chain_run<M, 2, true>,                               \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chain_ck_smem_bytes<M::D>())); \
      TN_CUDA(e, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_, k_chain_run_dyn<M, 2, true>, \
                                                               kChainBlock, sm_));           \
      const int slots_ = per_sm_ * n_sm;                                                     \
      if (slots_ > 0 && g_ > slots_ && e->rp.n_transitions >= 32) {                          \
        /* about ten tasks per slot, sub-ranges of at least 16 transitions */                \
        long long n_sub_ = (10LL * slots_ + g_ - 1) / g_;                                    \
        n_sub_ = std::max(1LL, std::min(n_sub_, e->rp.n_transitions / 16));                  \
        const long long len_ = (e->rp.n_transitions + n_sub_ - 1) / n_sub_;                  \
        n_sub_ = (e->rp.n_transitions + len_ - 1) / len_;                                    \
        if (!e->dyn_work) TN_CUDA(e, cudaMallocAsync((void **)&e->dyn_work, sizeof(int) * (size_t)(g_ + 1), e->stream)); \
        TN_CUDA(e, cudaMemsetAsync(e->dyn_work, 0, sizeof(int) * (size_t)(g_ + 1), e->stream)); \
        k_chain_run_dyn<M, 2, true><<<slots_, kChainBlock, sm_, e->stream>>>(                \
            e->small, e->st, e->out, e->rp, pm_, g_, (int)n_sub_, len_, e->dyn_work, e->dyn_work + 1, (int)(sm_ > 0)); \
      } else {                                                                               \
        k_chain_run<M, 2, true><<<g_, kChainBlock, sm_, e->stream>>>(e->small, e->st, e->out, e->rp, pm_, (int)(sm_ > 0)); \
      }                                                                                      \
    } else                                                                                   \
      k_chain_run<M, 2, true><<<g_, kChainBlock, 0, e->stream>>>(e->small, e->st, e->out, e->rp, pm_, 0); \
  } while (0)
  TN_DISPATCH_SMALL(e, CALL);
#undef CALL
  e->launches += 1;
  TN_CUDA(e, cudaGetLastError());
  return TNUTS_OK;
}

// sort chains by step size (ascending) -> e->perm; cub radix sort on the engine's stream
int thread_sort_chains(Engine *e) {
  const int C = e->st.C;
  if (!e->perm) {
    // stream-ordered like the draw buffer: a synchronous cudaMalloc/cudaFree next to the pooled
    // 11.5 GB draw buffer was measured to cost up to 0.5 s in tnuts_destroy every other job
    TN_CUDA(e, cudaMallocAsync((void **)&e->perm, sizeof(int) * (size_t)C * 2, e->stream));
    TN_CUDA(e, cudaMallocAsync((void **)&e->perm_keys, sizeof(double) * (size_t)C, e->stream));
    e->perm_tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, e->perm_tmp_bytes, (const double *)nullptr, e->perm_keys,
                                    (const int *)nullptr, e->perm, C, 0, 64, e->stream);
    TN_CUDA(e, cudaMallocAsync(&e->perm_tmp, e->perm_tmp_bytes, e->stream));
  }
  int *iota = e->perm + C;
  k_iota<<<grid_for(C, 256), 256, 0, e->stream>>>(iota, C);
  TN_CUDA(e, cub::DeviceRadixSort::SortPairs(e->perm_tmp, e->perm_tmp_bytes, e->st.eps, e->perm_keys, iota,
                                             e->perm, C, 0, 64, e->stream));
  e->launches += 2;
  e->perm_valid = true;
  return TNUTS_OK;
}

int thread_point_lpgrad(Engine *e, const double *theta_dev, long long n, double *lp_dev,
                        double *grad_dev) {
#define CALL(M) \
  k_point_lpgrad<M><<<grid_for(n, 128), 128, 0, e->stream>>>(e->small, theta_dev, n, lp_dev, grad_dev)
  TN_DISPATCH_SMALL(e, CALL);
#undef CALL
  e->launches += 1;
  TN_CUDA(e, cudaGetLastError());
  return TNUTS_OK;
}

int thread_point_constrain(Engine *e, const double *theta_dev, long long n, double *params_dev,
                           double *logp_dev) {
#define CALL(M)                                                                              \
  k_point_constrain<M><<<grid_for(n, 128), 128, 0, e->stream>>>(e->small, theta_dev, n,      \
                                                                params_dev, logp_dev)
  TN_DISPATCH_SMALL(e, CALL);
#undef CALL
  e->launches += 1;
  TN_CUDA(e, cudaGetLastError());
  return TNUTS_OK;
}

int thread_point_leapfrog(Engine *e, const double *theta_dev, const double *r_dev,
                          const double *minv_dev, double eps, int L, long long n, double *tt_dev,
                          double *tr_dev, double *tH_dev) {
#define CALL(M)                                                                               \
  k_point_leapfrog<M><<<grid_for(n, 128), 128, 0, e->stream>>>(e->small, theta_dev, r_dev,    \
                                                               minv_dev, eps, L, n, tt_dev,   \
                                                               tr_dev, tH_dev)
  TN_DISPATCH_SMALL(e, CALL);
#undef CALL
  e->launches += 1;
  TN_CUDA(e, cudaGetLastError());
  return TNUTS_OK;
}

}  // namespace tnuts
