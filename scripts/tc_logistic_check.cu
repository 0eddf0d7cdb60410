// Stand-alone check + timing of the tcgen05 logistic kernel (turing.jl_b200/csrc/logistic_tc.cuh)
// against a plain fp64 CPU evaluation.  Development tool; the parity tests proper go through the
// C ABI (tests/test_gpu_tc.py).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DTNUTS_TC_EXPERIMENTS -o scripts/tc_logistic_check scripts/tc_logistic_check.cu
//   scripts/tc_logistic_check 1          # correctness cases + config-3 timing
//   TC_XMODES=1 scripts/tc_logistic_check 2   # timing with pieces switched off (xmode bits: 1 no epilogue
//                                              # arithmetic, 2 one MMA pass, 4 one X plane)
//   TNUTS_TC_CHUNKS=n scripts/tc_logistic_check 3   # small chain counts with a forced row cut
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../turing.jl_b200/csrc/logistic_tc.cuh"

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t s_ = (x);                                                                  \
    if (s_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(s_)); \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

using namespace tnuts;

static int run_case(long long N, int D, int C, int ncheck, bool timing, bool use_list) {
  printf("=== N=%lld D=%d C=%d list=%d\n", N, D, C, (int)use_list);
  std::mt19937_64 rng(12345 + N + D);
  std::normal_distribution<double> nd(0.0, 1.0);
  std::uniform_real_distribution<double> ud(0.0, 1.0);
  std::vector<double> X((size_t)N * D), y((size_t)N), th((size_t)D * C), bstar(D);
  for (auto &b : bstar) b = nd(rng) * 2.0 / std::sqrt((double)D);
  for (long long i = 0; i < N; ++i) {
    double eta = 0;
    for (int d = 0; d < D; ++d) {
      X[(size_t)i * D + d] = nd(rng);
      eta += X[(size_t)i * D + d] * bstar[d];
    }
    y[(size_t)i] = ud(rng) < 1.0 / (1.0 + std::exp(-eta)) ? 1.0 : 0.0;
  }
  for (int d = 0; d < D; ++d)
    for (int c = 0; c < C; ++c) th[(size_t)d * C + c] = bstar[d] + 0.3 * nd(rng);
  std::vector<int> list(C);
  for (int k = 0; k < C; ++k) list[k] = use_list ? (C - 1 - k) : k;

  double *dX, *dy, *dth, *dpart;
  int *dlist;
  CK(cudaMalloc(&dX, 8 * X.size()));
  CK(cudaMalloc(&dy, 8 * y.size()));
  CK(cudaMalloc(&dth, 8 * th.size()));
  CK(cudaMalloc(&dlist, 4 * C));
  CK(cudaMemcpy(dX, X.data(), 8 * X.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy, y.data(), 8 * y.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dth, th.data(), 8 * th.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dlist, list.data(), 4 * C, cudaMemcpyHostToDevice));
  tc::Data data;
  const char *msg = tc::prepare(dX, dy, N, D, 0, &data);
  if (msg) {
    printf("prepare failed: %s\n", msg);
    return 1;
  }
  CK(cudaDeviceSynchronize());
  const tc::Plan &p = data.pl;
  printf("plan: KP=%d T=%d chunks=%d rows/chunk=%d smem=%zu\n", p.kp, p.t, p.chunks, p.rpc, p.smem);
  const int cap = ((C + 127) / 128) * 128;
  int chunks = 0, rpc = 0;
  tc::cut_for(N, cap / 128, &chunks, &rpc);   // the cut the launch below will use
  printf("cut for %d chain tiles: chunks=%d rows/chunk=%d\n", cap / 128, chunks, rpc);
  const size_t part_elems = (size_t)chunks * (D + 1) * cap;
  CK(cudaMalloc(&dpart, 8 * part_elems));
  CK(cudaMemset(dpart, 0xff, 8 * part_elems));
  float *ddbg;
  const size_t dbg_elems = 2 * 128 * (size_t)p.t;
  CK(cudaMalloc(&ddbg, 4 * dbg_elems));
  CK(cudaMemset(ddbg, 0, 4 * dbg_elems));
  CK(tc::launch(data, N, D, dth, C, use_list ? dlist : nullptr, nullptr, C, C, dpart, cap, 0, ddbg));
  CK(cudaDeviceSynchronize());
  std::vector<double> part(part_elems);
  std::vector<float> dbg(dbg_elems);
  CK(cudaMemcpy(part.data(), dpart, 8 * part_elems, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(dbg.data(), ddbg, 4 * dbg_elems, cudaMemcpyDeviceToHost));

  // ---- tile-0 eta / r of CTA (0,0) ----
  int bad = 0;
  {
    double max_eta = 0, max_r = 0;
    const int nk = C < 128 ? C : 128;
    for (int kl = 0; kl < nk; ++kl) {
      const int c = list[kl];
      for (int j = 0; j < p.t && j < N; ++j) {
        double eta = 0, ab = 0;
        for (int d = 0; d < D; ++d) {
          eta += X[(size_t)j * D + d] * th[(size_t)d * C + c];
          ab += std::fabs(X[(size_t)j * D + d] * th[(size_t)d * C + c]);
        }
        const double e1 = std::fabs(dbg[(size_t)kl * p.t + j] - eta) / ab;
        const double r = y[j] - 1.0 / (1.0 + std::exp(-eta));
        const double e2 = std::fabs(dbg[(size_t)128 * p.t + (size_t)kl * p.t + j] - r);
        if (e1 > max_eta) max_eta = e1;
        if (e2 > max_r) max_r = e2;
        if ((e1 > 2e-6 || e2 > 1e-5) && bad < 8) {
          printf("  tile0 mismatch kl=%d row=%d eta gpu=%.8g ref=%.8g  r gpu=%.8g ref=%.8g\n", kl, j,
                 dbg[(size_t)kl * p.t + j], eta, dbg[(size_t)128 * p.t + (size_t)kl * p.t + j], r);
          ++bad;
        }
      }
    }
    printf("tile0: max |d eta| / sum|x theta| = %.3e   max |d r| = %.3e\n", max_eta, max_r);
  }
  // ---- full lp / gradient for the first ncheck and last chains ----
  double max_lp_abs = 0, max_g_rel = 0, max_g_scaled = 0;
  std::vector<int> ks;
  for (int k = 0; k < ncheck && k < C; ++k) ks.push_back(k);
  if (C > ncheck) ks.push_back(C - 1);
  if (C > 130) ks.push_back(129);
  for (int k : ks) {
    const int c = list[k];
    std::vector<double> g(D, 0.0), gabs(D, 0.0);
    double lp = 0;
    for (long long i = 0; i < N; ++i) {
      double eta = 0;
      for (int d = 0; d < D; ++d) eta += X[(size_t)i * D + d] * th[(size_t)d * C + c];
      const double l1 = (eta > 0 ? eta : 0) + std::log1p(std::exp(-std::fabs(eta)));
      lp += y[(size_t)i] * eta - l1;
      const double r = y[(size_t)i] - 1.0 / (1.0 + std::exp(-eta));
      for (int d = 0; d < D; ++d) {
        g[d] += X[(size_t)i * D + d] * r;
        gabs[d] += std::fabs(X[(size_t)i * D + d] * r);
      }
    }
    double lp_gpu = 0;
    std::vector<double> g_gpu(D, 0.0);
    for (int ch = 0; ch < chunks; ++ch) {
      lp_gpu += part[((size_t)ch * (D + 1) + D) * cap + k];
      for (int d = 0; d < D; ++d) g_gpu[d] += part[((size_t)ch * (D + 1) + d) * cap + k];
    }
    double gn = 0, dn = 0;
    for (int d = 0; d < D; ++d) {
      gn += g[d] * g[d];
      dn += (g[d] - g_gpu[d]) * (g[d] - g_gpu[d]);
      const double sc = std::fabs(g[d] - g_gpu[d]) / gabs[d];
      if (sc > max_g_scaled) max_g_scaled = sc;
    }
    const double rel = std::sqrt(dn / gn);
    if (rel > max_g_rel) max_g_rel = rel;
    if (std::fabs(lp - lp_gpu) > max_lp_abs) max_lp_abs = std::fabs(lp - lp_gpu);
    if (k < 2 || rel > 1e-4)
      printf("  chain slot %d: lp gpu=%.10f ref=%.10f (|d|=%.2e)  |dg|/|g|=%.3e  g0 gpu=%.8f ref=%.8f\n", k, lp_gpu, lp,
             std::fabs(lp - lp_gpu), rel, g_gpu[0], g[0]);
  }
  printf("RESULT N=%lld D=%d C=%d: max |d lp| = %.3e   max |dg|/|g| = %.3e   max |dg_d|/sum|x r| = %.3e\n", N, D, C,
         max_lp_abs, max_g_rel, max_g_scaled);
  if (!(max_g_rel < 1e-4) || !(max_lp_abs < 1e-2 * (1 + N / 1e5))) bad = 1;
  if (timing) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int nx = getenv("TC_XMODES") ? 8 : 1;
    for (int xm = 0; xm < nx; ++xm) {
      for (int i = 0; i < 3; ++i) CK(tc::launch(data, N, D, dth, C, nullptr, nullptr, C, C, dpart, cap, 0, nullptr, xm));
      CK(cudaEventRecord(e0));
      const int reps = 20;
      for (int i = 0; i < reps; ++i) CK(tc::launch(data, N, D, dth, C, nullptr, nullptr, C, C, dpart, cap, 0, nullptr, xm));
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      ms /= reps;
      printf("TIMING xmode=%d N=%lld D=%d C=%d: %.3f ms per gradient of all chains = %.1f TFLOP/s-equivalent (4ND per chain)\n",
             xm, N, D, C, ms, 4.0 * N * D * C / (ms * 1e-3) / 1e12);
    }
  }
  tc::release(&data, 0);
  cudaFree(dX);
  cudaFree(dy);
  cudaFree(dth);
  cudaFree(dlist);
  cudaFree(dpart);
  cudaFree(ddbg);
  return bad;
}

int main(int argc, char **argv) {
  int bad = 0;
  const bool big = argc > 1 && atoi(argv[1]) > 0;
  if (argc > 1 && atoi(argv[1]) == 2) return run_case(100000, 100, 4096, 1, true, false);
  if (argc > 1 && atoi(argv[1]) == 3) return run_case(100000, 100, 128, 1, true, false) + run_case(100000, 100, 512, 1, true, false);
  bad += run_case(70, 100, 128, 4, false, false);
  bad += run_case(5000 + 17, 100, 200, 4, false, true);
  bad += run_case(3001, 128, 256, 3, false, false);
  bad += run_case(2000, 64, 130, 3, false, false);
  bad += run_case(999, 20, 64, 3, false, true);
  bad += run_case(1500, 96, 128, 3, false, false);
  if (big) bad += run_case(100000, 100, 4096, 3, true, false);
  printf(bad ? "FAILED (%d)\n" : "ALL OK\n", bad);
  return bad ? 1 : 0;
}
