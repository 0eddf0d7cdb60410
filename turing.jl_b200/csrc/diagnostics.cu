// diagnostics.cu -- on-device rank-normalised split R-hat and bulk ESS of the kept draws
// (Vehtari et al. 2021; what `ess(chain; kind=:bulk)` and `rhat(chain)` of MCMCChains /
// MCMCDiagnosticTools compute).  The draws never leave HBM: 65 536 chains
// x 1000 draws x 10 parameters is 5.2 GB, far too much to diagnose on the host per run.
//
// Per parameter:
//   1. gather the split half-chains (n = K/2 draws each, 2C half-chains) into one key array;
//   2. radix-sort (cub) with the element index as payload; tied keys get the average rank;
//   3. z = Phi^-1((rank - 3/8) / (S + 1/4)) scattered back to [n][2C] (half-chain fastest);
//   4. per-half-chain mean and biased autocovariance at the lags of a batch; the sum over
//      half-chains is reduced deterministically (warp shuffle -> per-block partials -> host);
//   5. the host applies Geyer's initial-positive / monotone sequence to the [lags] vector and
//      asks for another lag batch only when the sequence has not been truncated yet.
//   R-hat is max(bulk, folded) like rhat(kind=:rank).
#include <cub/cub.cuh>

#include <cmath>
#include <vector>

#include "engine.cuh"

namespace tnuts {

struct DiagWork {
  size_t S = 0;
  double *keys = nullptr, *keys_alt = nullptr, *z = nullptr;
  uint32_t *idx = nullptr, *idx_alt = nullptr;
  void *cub_tmp = nullptr;
  size_t cub_bytes = 0;
  double *mean_h = nullptr, *var_h = nullptr;  // [2C]
  double *partial = nullptr;                   // [lags][blocks]
  size_t partial_cap = 0;
};

__global__ void k_diag_gather(const double *draws, int ncol, int C, long long K, long long n, int p,
                              double *keys, uint32_t *idx, double fold_center, int fold) {
  const long long S = 2LL * C * n;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < S;
       j += (long long)gridDim.x * blockDim.x) {
    const long long i = j / (2LL * C);
    const int h = (int)(j % (2LL * C));
    const int half = h / C, c = h % C;
    const long long k = i + (half ? (K - n) : 0);
    double v = draws[((size_t)k * ncol + p) * C + c];
    if (fold) v = fabs(v - fold_center);
    keys[j] = v;
    idx[j] = (uint32_t)j;
  }
}

__global__ void k_diag_rank(const double *keys, const uint32_t *idx, long long S, double *z) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < S;
       j += (long long)gridDim.x * blockDim.x) {
    const double kj = keys[j];
    if (j > 0 && keys[j - 1] == kj) continue;  // not a run start
    long long e = j;
    while (e + 1 < S && keys[e + 1] == kj) ++e;
    const double avg = 0.5 * (double)(j + e) + 1.0;  // 1-based average rank of the tie run
    const double zz = normcdfinv((avg - 0.375) / ((double)S + 0.25));
    for (long long q = j; q <= e; ++q) z[idx[q]] = zz;
  }
}

// per half-chain mean and biased variance; z is [n][H]
__global__ void k_diag_moments(const double *z, long long n, int H, double *mean_h, double *var_h) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  double s = 0;
  for (long long i = 0; i < n; ++i) s += z[i * H + h];
  const double m = s / (double)n;
  double q = 0;
  for (long long i = 0; i < n; ++i) {
    const double d = z[i * H + h] - m;
    q += d * d;
  }
  mean_h[h] = m;
  var_h[h] = q / (double)n;
}

// biased autocovariance at lags t0 + threadIdx.y ..., summed over the 32 half-chains of a warp
// block = (32, LT); grid = (ceil(H/32)); partial[(t - t0) * gridDim.x + blockIdx.x]
__global__ void k_diag_acov(const double *z, const double *mean_h, long long n, int H, int t0,
                            int nlag, double *partial) {
  const int h = blockIdx.x * 32 + threadIdx.x;
  for (int tt = threadIdx.y; tt < nlag; tt += blockDim.y) {
    const int t = t0 + tt;
    double acc = 0.0;
    if (h < H && t < n) {
      const double m = mean_h[h];
      for (long long i = 0; i + t < n; ++i) acc += (z[i * H + h] - m) * (z[(i + t) * H + h] - m);
      acc /= (double)n;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) partial[(size_t)tt * gridDim.x + blockIdx.x] = acc;
  }
}

static void free_work(DiagWork *w) {
  for (void *p : {(void *)w->keys, (void *)w->keys_alt, (void *)w->z, (void *)w->idx,
                  (void *)w->idx_alt, w->cub_tmp, (void *)w->mean_h, (void *)w->var_h,
                  (void *)w->partial})
    if (p) cudaFree(p);
  *w = DiagWork();
}

void diagnostics_destroy(Engine *e) {
  if (e->diag) {
    free_work(e->diag);
    delete e->diag;
    e->diag = nullptr;
  }
}

// host side of one parameter, given per-half-chain means/vars and a lag oracle
struct LagSource {
  Engine *e;
  DiagWork *w;
  long long n;
  int H;
  std::vector<double> acov;  // mean over half-chains of biased acov, by lag
  int rc = TNUTS_OK;
  bool extend(int upto) {  // make acov valid for lags < upto
    while ((int)acov.size() < upto && (long long)acov.size() < n) {
      const int t0 = (int)acov.size();
      int nlag = t0 == 0 ? 32 : t0;  // 32, 32, 64, 128, ...
      if (t0 + nlag > n) nlag = (int)(n - t0);
      const int blocks = (H + 31) / 32;
      const size_t need = (size_t)nlag * blocks;
      if (need > w->partial_cap) {
        if (w->partial) cudaFree(w->partial);
        if (cudaMalloc((void **)&w->partial, 8 * need) != cudaSuccess) {
          rc = set_error(e, TNUTS_E_CUDA, "diagnostics: cudaMalloc(partial) failed");
          return false;
        }
        w->partial_cap = need;
      }
      k_diag_acov<<<blocks, dim3(32, 8), 0, e->stream>>>(w->z, w->mean_h, n, H, t0, nlag, w->partial);
      e->launches += 1;
      std::vector<double> part(need);
      if (cudaMemcpyAsync(part.data(), w->partial, 8 * need, cudaMemcpyDeviceToHost, e->stream) != cudaSuccess ||
          cudaStreamSynchronize(e->stream) != cudaSuccess) {
        rc = set_error(e, TNUTS_E_CUDA, "diagnostics: acov D2H failed");
        return false;
      }
      for (int tt = 0; tt < nlag; ++tt) {
        double s = 0;
        for (int b = 0; b < blocks; ++b) s += part[(size_t)tt * blocks + b];
        acov.push_back(s / (double)H);
      }
    }
    return (int)acov.size() >= upto;
  }
};

struct DiagSrc {  // where the draws of one parameter live: draws[(k * ncol + col) * C + c]
  const double *draws;
  int ncol, col, C;
};

static int rank_normalise(Engine *e, DiagWork *w, const DiagSrc &src, long long K, long long n,
                          bool fold, double center, double *median_out) {
  const int C = src.C;
  const long long S = 2LL * C * n;
  const int grid = 148 * 8;
  k_diag_gather<<<grid, 256, 0, e->stream>>>(src.draws, src.ncol, C, K, n, src.col, w->keys, w->idx,
                                             center, fold ? 1 : 0);
  cub::DoubleBuffer<double> dk(w->keys, w->keys_alt);
  cub::DoubleBuffer<uint32_t> dv(w->idx, w->idx_alt);
  TN_CUDA(e, cub::DeviceRadixSort::SortPairs(w->cub_tmp, w->cub_bytes, dk, dv, (int)S, 0, 64, e->stream));
  if (median_out) {
    double m2[2];
    TN_CUDA(e, cudaMemcpyAsync(m2, dk.Current() + (S - 1) / 2, 8, cudaMemcpyDeviceToHost, e->stream));
    TN_CUDA(e, cudaMemcpyAsync(m2 + 1, dk.Current() + S / 2, 8, cudaMemcpyDeviceToHost, e->stream));
    TN_CUDA(e, cudaStreamSynchronize(e->stream));
    *median_out = 0.5 * (m2[0] + m2[1]);
  }
  k_diag_rank<<<grid, 256, 0, e->stream>>>(dk.Current(), dv.Current(), S, w->z);
  k_diag_moments<<<(2 * C + 127) / 128, 128, 0, e->stream>>>(w->z, n, 2 * C, w->mean_h, w->var_h);
  e->launches += 4;
  TN_CUDA(e, cudaGetLastError());
  return TNUTS_OK;
}

// between/within variances from per-half-chain moments (host, fixed summation order)
static void within_between(const std::vector<double> &mean_h, const std::vector<double> &var_h,
                           long long n, double *mean_var, double *var_plus) {
  const size_t H = mean_h.size();
  double w = 0, mm = 0;
  for (size_t h = 0; h < H; ++h) {
    w += var_h[h] * (double)n / (double)(n - 1);
    mm += mean_h[h];
  }
  w /= (double)H;
  mm /= (double)H;
  double b = 0;
  for (size_t h = 0; h < H; ++h) b += (mean_h[h] - mm) * (mean_h[h] - mm);
  b /= (double)(H - 1);
  *mean_var = w;
  *var_plus = w * (double)(n - 1) / (double)n + (H > 1 ? b : 0.0);
}

static int ensure_work(Engine *e, int C, long long K) {
  const int H = 2 * C;
  const long long n = K / 2;
  if (n < 2) return set_error(e, TNUTS_E_ARG, "diagnostics: need at least 4 kept draws");
  const long long S = (long long)H * n;
  if (S >= (1LL << 31)) return set_error(e, TNUTS_E_UNSUPPORTED, "diagnostics: more than 2^31 draws per parameter");
  if (!e->diag) e->diag = new DiagWork();
  DiagWork *w = e->diag;
  if (w->S != (size_t)S) {
    free_work(w);
    TN_CUDA(e, cudaMalloc((void **)&w->keys, 8 * S));
    TN_CUDA(e, cudaMalloc((void **)&w->keys_alt, 8 * S));
    TN_CUDA(e, cudaMalloc((void **)&w->z, 8 * S));
    TN_CUDA(e, cudaMalloc((void **)&w->idx, 4 * S));
    TN_CUDA(e, cudaMalloc((void **)&w->idx_alt, 4 * S));
    TN_CUDA(e, cudaMalloc((void **)&w->mean_h, 8 * (size_t)H));
    TN_CUDA(e, cudaMalloc((void **)&w->var_h, 8 * (size_t)H));
    cub::DoubleBuffer<double> dk(w->keys, w->keys_alt);
    cub::DoubleBuffer<uint32_t> dv(w->idx, w->idx_alt);
    TN_CUDA(e, cub::DeviceRadixSort::SortPairs(nullptr, w->cub_bytes, dk, dv, (int)S, 0, 64, e->stream));
    TN_CUDA(e, cudaMalloc(&w->cub_tmp, w->cub_bytes));
    w->S = (size_t)S;
  }
  return TNUTS_OK;
}

// rank-normalised split R-hat (max of bulk and folded) and bulk ESS of ONE parameter
static int diag_one_param(Engine *e, const DiagSrc &src, long long K, double *rhat_out, double *ess_out) {
  DiagWork *w = e->diag;
  const int C = src.C, H = 2 * C;
  const long long n = K / 2;
  const long long S = (long long)H * n;
  std::vector<double> mean_h((size_t)H), var_h((size_t)H);
  {
    double median = 0.0;
    int rc = rank_normalise(e, w, src, K, n, false, 0.0, &median);
    if (rc) return rc;
    TN_CUDA(e, cudaMemcpyAsync(mean_h.data(), w->mean_h, 8 * (size_t)H, cudaMemcpyDeviceToHost, e->stream));
    TN_CUDA(e, cudaMemcpyAsync(var_h.data(), w->var_h, 8 * (size_t)H, cudaMemcpyDeviceToHost, e->stream));
    TN_CUDA(e, cudaStreamSynchronize(e->stream));
    double mean_var, var_plus;
    within_between(mean_h, var_h, n, &mean_var, &var_plus);
    const double rhat_bulk = std::sqrt(var_plus / mean_var);

    // Geyer initial monotone sequence (same control flow as Stan / ArviZ)
    LagSource ls{e, w, n, H, {}};
    auto rho_at = [&](long long t) -> double {
      ls.extend((int)t + 1);
      return 1.0 - (mean_var - ls.acov[(size_t)t]) / var_plus;
    };
    std::vector<double> rho((size_t)n, 0.0);
    rho[0] = 1.0;
    double rho_even = 1.0, rho_odd = rho_at(1);
    if (ls.rc) return ls.rc;
    rho[1] = rho_odd;
    long long t = 1;
    while (t < n - 3 && (rho_even + rho_odd) > 0) {
      rho_even = rho_at(t + 1);
      rho_odd = rho_at(t + 2);
      if (ls.rc) return ls.rc;
      if (rho_even + rho_odd >= 0) {
        rho[(size_t)t + 1] = rho_even;
        rho[(size_t)t + 2] = rho_odd;
      }
      t += 2;
    }
    const long long max_t = t;
    if (rho_even > 0) rho[(size_t)max_t + 1] = rho_even;
    t = 1;
    while (t <= max_t - 2) {
      if (rho[(size_t)t + 1] + rho[(size_t)t + 2] > rho[(size_t)t - 1] + rho[(size_t)t]) {
        rho[(size_t)t + 1] = (rho[(size_t)t - 1] + rho[(size_t)t]) / 2.0;
        rho[(size_t)t + 2] = rho[(size_t)t + 1];
      }
      t += 2;
    }
    double tau = -1.0;
    for (long long q = 0; q <= max_t; ++q) tau += 2.0 * rho[(size_t)q];
    if (rho_even > 0) tau += rho[(size_t)max_t + 1];
    const double floor_tau = 1.0 / std::log10((double)S);
    if (tau < floor_tau) tau = floor_tau;
    *ess_out = (double)S / tau;

    // folded (tail) R-hat
    rc = rank_normalise(e, w, src, K, n, true, median, nullptr);
    if (rc) return rc;
    TN_CUDA(e, cudaMemcpyAsync(mean_h.data(), w->mean_h, 8 * (size_t)H, cudaMemcpyDeviceToHost, e->stream));
    TN_CUDA(e, cudaMemcpyAsync(var_h.data(), w->var_h, 8 * (size_t)H, cudaMemcpyDeviceToHost, e->stream));
    TN_CUDA(e, cudaStreamSynchronize(e->stream));
    double mv2, vp2;
    within_between(mean_h, var_h, n, &mv2, &vp2);
    const double rhat_tail = std::sqrt(vp2 / mv2);
    *rhat_out = rhat_bulk > rhat_tail ? rhat_bulk : rhat_tail;
  }
  return TNUTS_OK;
}

int diagnostics_run_distributed(Engine *e, double *rhat_host, double *ess_host);  // below

int diagnostics_run(Engine *e, double *rhat_host, double *ess_host) {
  if (e->comm && e->world > 1 && e->cfg.row_shards <= 1) return diagnostics_run_distributed(e, rhat_host, ess_host);
  const int C = e->st.C, P = e->P;
  const long long K = e->n_kept;
  int rc = ensure_work(e, C, K);
  if (rc) return rc;
  for (int p = 0; p < P; ++p) {
    DiagSrc src{e->out.draws, e->out.ncol, p, C};
    if ((rc = diag_one_param(e, src, K, &rhat_host[p], &ess_host[p]))) return rc;
  }
  return TNUTS_OK;
}

// ------------------------------------------------------------------------------------------
// chains sharded over GPUs: the rank normalisation needs the pooled draws of ALL chains, so
// the parameters are dealt out to the ranks (parameter p is diagnosed by rank p mod W):
// every rank packs its [K][C_local] slab of each parameter and ncclSend's it to the owner,
// the owner assembles [K][C_total] and runs the single-GPU diagnostic on it; the [P] results
// are combined with one small all-reduce.  Traffic: the kept parameter draws cross NVLink once.
// ------------------------------------------------------------------------------------------
__global__ void k_diag_pack(const double *draws, int ncol, int col, int C, long long K, double *dst) {
  const long long tot = K * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot;
       i += (long long)gridDim.x * blockDim.x) {
    const long long k = i / C;
    const int c = (int)(i % C);
    dst[i] = draws[((size_t)k * ncol + col) * C + c];
  }
}

// stacked [q][K][Cq] -> [K][Ctot]
__global__ void k_diag_unstack(const double *src, long long K, int Cq, int off, int Ctot, double *dst) {
  const long long tot = K * Cq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot;
       i += (long long)gridDim.x * blockDim.x) {
    const long long k = i / Cq;
    const int c = (int)(i % Cq);
    dst[(size_t)k * Ctot + off + c] = src[i];
  }
}

int comm_exchange_slabs(Engine *e, const double *const *send, const size_t *send_elems, bool do_recv,
                        double *recv, const size_t *recv_off, const size_t *recv_elems);  // comm.cu

int diagnostics_run_distributed(Engine *e, double *rhat_host, double *ess_host) {
  const int W = e->world, R = e->rank, C = e->st.C, P = e->P;
  const long long K = e->n_kept;
  int rc;
  // chain counts of every rank
  std::vector<double> cnt((size_t)W, 0.0);
  cnt[(size_t)R] = (double)C;
  {
    double *d = nullptr;
    TN_CUDA(e, cudaMalloc((void **)&d, 8 * (size_t)W));
    TN_CUDA(e, cudaMemcpyAsync(d, cnt.data(), 8 * (size_t)W, cudaMemcpyHostToDevice, e->stream));
    rc = comm_allreduce_sum(e, d, (size_t)W);
    if (!rc) {
      cudaMemcpyAsync(cnt.data(), d, 8 * (size_t)W, cudaMemcpyDeviceToHost, e->stream);
      cudaStreamSynchronize(e->stream);
    }
    cudaFree(d);
    if (rc) return rc;
  }
  std::vector<size_t> off((size_t)W + 1, 0), relems((size_t)W), roff((size_t)W), selems((size_t)W);
  for (int q = 0; q < W; ++q) off[(size_t)q + 1] = off[(size_t)q] + (size_t)cnt[(size_t)q];
  const int Ctot = (int)off[(size_t)W];
  for (int q = 0; q < W; ++q) {
    relems[(size_t)q] = (size_t)K * (size_t)cnt[(size_t)q];
    roff[(size_t)q] = (size_t)K * off[(size_t)q];
  }
  if ((rc = ensure_work(e, Ctot, K))) return rc;
  double *sendbuf = nullptr, *stack = nullptr, *pooled = nullptr;
  TN_CUDA(e, cudaMalloc((void **)&sendbuf, 8 * (size_t)W * K * C));
  TN_CUDA(e, cudaMalloc((void **)&stack, 8 * (size_t)K * Ctot));
  TN_CUDA(e, cudaMalloc((void **)&pooled, 8 * (size_t)K * Ctot));
  std::vector<double> res(2 * (size_t)P, 0.0);
  rc = TNUTS_OK;
  for (int r0 = 0; r0 < P && !rc; r0 += W) {
    std::vector<const double *> sp((size_t)W, nullptr);
    for (int q = 0; q < W; ++q) {
      const int p = r0 + q;
      selems[(size_t)q] = 0;
      if (p < P) {
        double *dst = sendbuf + (size_t)q * K * C;
        k_diag_pack<<<148 * 4, 256, 0, e->stream>>>(e->out.draws, e->out.ncol, p, C, K, dst);
        sp[(size_t)q] = dst;
        selems[(size_t)q] = (size_t)K * C;
        e->launches += 1;
      }
    }
    const int mine = r0 + R;
    rc = comm_exchange_slabs(e, sp.data(), selems.data(), mine < P, stack, roff.data(), relems.data());
    if (rc) break;
    if (mine < P) {
      for (int q = 0; q < W; ++q)
        k_diag_unstack<<<148 * 4, 256, 0, e->stream>>>(stack + roff[(size_t)q], K, (int)cnt[(size_t)q],
                                                       (int)off[(size_t)q], Ctot, pooled);
      e->launches += W;
      DiagSrc src{pooled, 1, 0, Ctot};
      rc = diag_one_param(e, src, K, &res[(size_t)mine], &res[(size_t)P + mine]);
    }
  }
  cudaStreamSynchronize(e->stream);
  cudaFree(sendbuf);
  cudaFree(stack);
  cudaFree(pooled);
  if (rc) return rc;
  {
    double *d = nullptr;
    TN_CUDA(e, cudaMalloc((void **)&d, 8 * res.size()));
    TN_CUDA(e, cudaMemcpyAsync(d, res.data(), 8 * res.size(), cudaMemcpyHostToDevice, e->stream));
    rc = comm_allreduce_sum(e, d, res.size());
    if (!rc) {
      cudaMemcpyAsync(res.data(), d, 8 * res.size(), cudaMemcpyDeviceToHost, e->stream);
      cudaStreamSynchronize(e->stream);
    }
    cudaFree(d);
    if (rc) return rc;
  }
  for (int p = 0; p < P; ++p) {
    rhat_host[p] = res[(size_t)p];
    ess_host[p] = res[(size_t)P + p];
  }
  return TNUTS_OK;
}

}  // namespace tnuts
