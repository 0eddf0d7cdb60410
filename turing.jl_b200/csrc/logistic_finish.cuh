// logistic_finish.cuh -- the prior terms added to the reduced likelihood sums of the logistic model,
// shared by k_logistic_reduce_finish (lpgrad_kernels.cuh) and the persistent pump kernel
// (tc_persist.cuh).
#pragma once
#include "common.cuh"

namespace tnuts {

// prior terms of the logistic model added to the reduced likelihood sums.  Explicit fused
// multiply-adds: the persistent pump kernel (tc_persist.cuh) evaluates the very same expressions and
// must produce the very same bits, whatever contraction the compiler would pick in either context.
__device__ __forceinline__ double logistic_finish_grad(double s, double th, double is2) {
  return __fma_rn(-th, is2, s);
}
__device__ __forceinline__ double logistic_finish_lp(double s, double q, double is2, int D, double sd) {
  const double c0 = __dsub_rn(__dmul_rn(-0.5 * D, kLog2Pi), __dmul_rn((double)D, log(sd)));
  return __dadd_rn(s, __fma_rn(__dmul_rn(-0.5, q), is2, c0));
}

}  // namespace tnuts
