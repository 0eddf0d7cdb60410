// block.cu -- launchers of the one-CTA-per-chain strategy (block_kernels.cuh).
#include "block_kernels.cuh"
#include "engine.cuh"

namespace tnuts {

bool block_strategy_supported(const Engine *e) {
  const int f = e->model.family, D = e->D;
  if (e->cfg.max_depth > kBlockCk + 1) return false;
  if (D > 8 * kBlockThreads) return false;
  return f == TNUTS_HIER_LINREG || f == TNUTS_STD_NORMAL;
}

int block_run(Engine *e) {
  LockstepState *ls = e->ls;
  const ModelDev &m = e->model;
  const int C = e->st.C, D = e->D, ncol = e->out.ncol;
  long long keep = 0;
  const RunParams &rp = e->rp;
  if (rp.first_keep >= 1 && rp.n_transitions >= rp.first_keep)
    keep = (rp.n_transitions - rp.first_keep) / rp.thin + 1;
  double *cm = nullptr, *tcm = nullptr;
  if (keep > 0) {
    TN_CUDA(e, cudaMalloc((void **)&cm, 8 * (size_t)C * keep * ncol));
    if (e->out.theta) TN_CUDA(e, cudaMalloc((void **)&tcm, 8 * (size_t)C * keep * D));
  }
  HierStatsDev hs{ls->hs_n, ls->hs_xbar, ls->hs_ybar, ls->hs_sxx, ls->hs_sxy, ls->hs_syy};
  const size_t smem = sizeof(double) * ((size_t)D + (2 + 2 * kBlockCk) * kBlockWarps);
  const int kd = (D + kBlockThreads - 1) / kBlockThreads;
#define TN_BLK(M, KD)                                                                              \
  do {                                                                                             \
    auto kfn = k_block_run<M<KD>, KD>;                                                             \
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
    kfn<<<C, kBlockThreads, smem, e->stream>>>(m, hs, e->st, rp, cm, keep, ncol, tcm);             \
  } while (0)
#define TN_BLK_KD(M)                  \
  do {                                \
    if (kd <= 1) TN_BLK(M, 1);        \
    else if (kd <= 2) TN_BLK(M, 2);   \
    else if (kd <= 4) TN_BLK(M, 4);   \
    else TN_BLK(M, 8);                \
  } while (0)
  if (m.family == TNUTS_HIER_LINREG) TN_BLK_KD(HierBlock);
  else TN_BLK_KD(StdNormalBlock);
#undef TN_BLK_KD
#undef TN_BLK
  e->launches += 1;
  cudaError_t s = cudaGetLastError();
  if (s == cudaSuccess && keep > 0) {
    const long long rows = keep * ncol;
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((C + 31) / 32));
    k_block_transpose_draws<<<grid, dim3(32, 8), 0, e->stream>>>(cm, e->out.draws, keep, ncol, C);
    if (tcm) {
      const long long trows = keep * D;
      dim3 tg((unsigned)((trows + 31) / 32), (unsigned)((C + 31) / 32));
      k_block_transpose_draws<<<tg, dim3(32, 8), 0, e->stream>>>(tcm, e->out.theta, keep, D, C);
    }
    e->launches += tcm ? 2 : 1;
    s = cudaGetLastError();
  }
  if (s == cudaSuccess) s = cudaStreamSynchronize(e->stream);
  if (cm) cudaFree(cm);
  if (tcm) cudaFree(tcm);
  if (s != cudaSuccess) return set_error(e, TNUTS_E_CUDA, std::string("block_run: ") + cudaGetErrorString(s));
  return TNUTS_OK;
}

}  // namespace tnuts
