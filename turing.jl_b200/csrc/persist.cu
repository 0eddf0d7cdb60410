// persist.cu -- host side of the persistent pump kernel (tc_persist.cuh), in its own translation unit
// (the kernel instantiates the whole tcgen05 core and the NUTS state machine: it compiles in parallel
// with lockstep.cu).
#include "tc_persist.cuh"

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <string>
#include <vector>

#include "engine.cuh"

namespace tnuts {

// Runs pumps [pump0, ...) of the current lock-step run until every chain has finished (or the pump
// budget is spent).  The work list must have been refreshed by the caller for `listed_done`, and
// n_max = C - listed_done must fit kPersistMaxTiles chain tiles.  Synchronises the stream.
// *pumps: pumps executed; *n_done: finished chains at the end.
// *reason: 1 / 2 every chain finished, 0 pump budget spent, 3 returned at a list refresh with at most
// pv.exit_below chains unfinished (the caller gathers them into a narrower view and calls again with
// listed_done = *sched_done)
int tc_persist_run(Engine *e, const tc::Data &td, const PersistView &pv, double *part, int stride,
                   int *count_done_dev, int *h_done, long long pump0, long long max_pumps, int listed_done, int n_max,
                   long long *pumps, int *n_done, int *reason, int *sched_done) {
  const ModelDev &m = e->model;
  int n_sm = 0;
  TN_CUDA(e, cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, e->device));
  if (n_sm < tc::kSMs) return set_error(e, TNUTS_E_UNSUPPORTED, "persistent pump: fewer SMs than the row cut assumes");
  tc::PersistCut cut{};
  for (int nt = 1; nt <= tc::kPersistMaxTiles; ++nt) tc::cut_for(m.N, nt, &cut.chunks[nt], &cut.rpc[nt]);
  // barrier counter + out_state in one small allocation
  unsigned long long *scratch = nullptr;
  const size_t n_scratch = 32 + 5 * (size_t)tc::kSMs;
  TN_CUDA(e, cudaMallocAsync((void **)&scratch, 8 * n_scratch, e->stream));
  TN_CUDA(e, cudaMemsetAsync(scratch, 0, 8 * n_scratch, e->stream));
  tc::PersistArgs pa{};
  pa.pump0 = pump0;
  pa.max_pumps = max_pumps;
  pa.listed_done = listed_done;
  pa.n_max = n_max;
  pa.c_total = pv.c_total;
  pa.exit_below = pv.exit_below;
  static const int bar_mode = [] { const char *v = getenv("TNUTS_BAR_MODE"); return v ? atoi(v) : 0; }();
  pa.bar_mode = bar_mode;
  pa.gbar = scratch;
  pa.n_done = count_done_dev;
  pa.h_done = h_done;
  pa.out_state = (long long *)(scratch + 1);
  static const bool trace = getenv("TNUTS_PUMP_TRACE") != nullptr;
  pa.cta_times = trace ? (long long *)(scratch + 32) : nullptr;
  long long nrows = m.N;
  int cap = stride;
  const float *yf = td.yf;
  void *args[] = {(void *)&td.map, (void *)&yf, (void *)&nrows, (void *)&part, (void *)&cap, (void *)pv.st,
                  (void *)pv.ls,   (void *)&e->rp, (void *)&e->model, (void *)pv.out, (void *)&cut, (void *)&pa};
  const void *fn = td.pl.kp == 128 ? (const void *)tc::k_tc_persist<128> : (const void *)tc::k_tc_persist<64>;
  TN_CUDA(e, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)td.pl.smem));
  TN_CUDA(e, cudaLaunchCooperativeKernel(fn, dim3(tc::kSMs), dim3(tc::kThreads), args, td.pl.smem, e->stream));
  e->launches += 1;
  long long st4[22] = {0};
  std::vector<long long> ct(5 * (size_t)tc::kSMs, 0);
  TN_CUDA(e, cudaMemcpyAsync(st4, pa.out_state, sizeof(st4), cudaMemcpyDeviceToHost, e->stream));
  if (trace) TN_CUDA(e, cudaMemcpyAsync(ct.data(), scratch + 32, 8 * ct.size(), cudaMemcpyDeviceToHost, e->stream));
  cudaError_t s = cudaStreamSynchronize(e->stream);
  cudaFreeAsync(scratch, e->stream);
  if (s != cudaSuccess)
    return set_error(e, TNUTS_E_CUDA, std::string("persistent pump kernel: ") + cudaGetErrorString(s));
  *pumps = st4[0];
  *n_done = (int)st4[2];
  *reason = (int)st4[3];
  *sched_done = (int)st4[11];
  if (trace && st4[0] > 0) {
    long long mx[3] = {0, 0, 0}, sm[3] = {0, 0, 0};
    for (int b = 0; b < tc::kSMs; ++b)
      for (int i = 0; i < 3; ++i) {
        mx[i] = std::max(mx[i], ct[5 * b + i]);
        sm[i] += ct[5 * b + i];
      }
    fprintf(stderr, "[tnuts persist] us per pump inside the phases, busiest CTA / mean CTA: gradient %.1f / %.1f | reduce %.1f / %.1f | "
            "state machine %.1f / %.1f\n", mx[0] * 1e-3 / st4[0], sm[0] * 1e-3 / st4[0] / tc::kSMs, mx[1] * 1e-3 / st4[0],
            sm[1] * 1e-3 / st4[0] / tc::kSMs, mx[2] * 1e-3 / st4[0], sm[2] * 1e-3 / st4[0] / tc::kSMs);
    {
      std::vector<long long> gs;
      for (int b = 0; b < tc::kSMs; ++b) gs.push_back(ct[5 * b]);
      fprintf(stderr, "[tnuts persist] gradient phase, us per pump by CTA:");
      for (int b = 0; b < tc::kSMs; ++b) fprintf(stderr, " %.0f", gs[b] * 1e-3 / st4[0]);
      fprintf(stderr, "\n[tnuts persist] reduce phase, us per pump by CTA:");
      for (int b = 0; b < tc::kSMs; ++b) fprintf(stderr, " %.0f", ct[5 * b + 1] * 1e-3 / st4[0]);
      fprintf(stderr, "\n[tnuts persist] slowest state-machine warp of the CTA, us per pump by CTA:");
      for (int b = 0; b < tc::kSMs; ++b) fprintf(stderr, " %.0f", ct[5 * b + 3] * 1e-3 / st4[0]);
      fprintf(stderr, "\n[tnuts persist] state machine phase, us per pump by CTA:");
      for (int b = 0; b < tc::kSMs; ++b) fprintf(stderr, " %.0f", ct[5 * b + 2] * 1e-3 / st4[0]);
      fprintf(stderr, "\n");
    }
    const double k = 1e-3 / (double)st4[0];
    fprintf(stderr,
            "[tnuts persist] %lld pumps, us per pump at CTA 0: gradient %.1f | barrier %.1f | reduce %.1f | barrier %.1f | "
            "state machine %.1f | barrier %.1f | list refreshes %.1f\n",
            st4[0], st4[4] * k, st4[5] * k, st4[6] * k, st4[7] * k, st4[8] * k, st4[9] * k, st4[10] * k);
    fprintf(stderr, "[tnuts persist] inside the gradient phase at CTA 0: positions -> TMEM %.1f | row tiles %.1f | drain + partial sums %.1f\n",
            st4[12] * k, st4[13] * k, st4[14] * k);
    fprintf(stderr, "[tnuts persist] inside the state machine, first warp of CTA 0: scalars in %.1f | [A] %.1f | leaf weights %.1f | "
            "checkpoints + U-turns %.1f | [B] %.1f | [T]+[S] %.1f | [D] + scalars out %.1f\n",
            st4[15] * k, st4[16] * k, st4[17] * k, st4[18] * k, st4[19] * k, st4[20] * k, st4[21] * k);
  }
  return TNUTS_OK;
}

}  // namespace tnuts
