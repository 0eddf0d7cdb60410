// engine.cu -- C ABI of the tnuts engine (include/tnuts.h).  Host-side plumbing only:
// device memory, model upload, strategy selection, launches, D2H of kept draws.
#include "engine.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace tnuts {

static std::string g_create_error;
static std::mutex g_create_mu;

int set_error(Engine *e, int code, const std::string &msg) {
  if (e) e->err = msg;
  else {
    std::lock_guard<std::mutex> lk(g_create_mu);
    g_create_error = msg;
  }
  return code;
}

template <class T>
static int dev_alloc(Engine *e, std::vector<void *> &pool, T **p, size_t n) {
  // From the device's stream-ordered pool, like the draw buffer: the ~25 state arrays of an engine
  // are recycled by the next engine of the process.  With cudaMalloc / cudaFree, tnuts_destroy
  // stalled for 40-500 ms in one job out of three (the driver unmapping them one by one).
  void *q = nullptr;
  TN_CUDA(e, cudaMallocAsync(&q, (n ? n : 1) * sizeof(T), e->stream));
  TN_CUDA(e, cudaMemsetAsync(q, 0, (n ? n : 1) * sizeof(T), e->stream));
  pool.push_back(q);
  *p = (T *)q;
  return TNUTS_OK;
}
static void free_pool(std::vector<void *> &pool, cudaStream_t stream) {
  for (void *p : pool) cudaFreeAsync(p, stream);
  pool.clear();
}

// Stan's windowed-adaptation schedule as used by AdvancedHMC's StanHMCAdaptor
// (init_buffer 75, term_buffer 50, window_size 25), 1-based.
static void window_schedule(long long n, RunParams &rp) {
  long long init_buffer = 75, term_buffer = 50, base = 25;
  rp.n_splits = 0;
  if (n < 20) {
    rp.window_start = n + 1;
    rp.window_end = n;
    return;
  }
  if (init_buffer + base + term_buffer > n) {
    init_buffer = (long long)(0.15 * (double)n);
    term_buffer = (long long)(0.1 * (double)n);
    base = n - init_buffer - term_buffer;
  }
  rp.window_start = init_buffer + 1;
  rp.window_end = n - term_buffer;
  long long next = init_buffer + base;
  while (rp.n_splits < 24) {
    if (next >= rp.window_end) {
      rp.splits[rp.n_splits++] = rp.window_end;
      break;
    }
    rp.splits[rp.n_splits++] = next;
    base *= 2;
    long long nn = next + base;
    if (nn != rp.window_end && nn + 2 * base > rp.window_end) nn = rp.window_end;
    next = nn;
  }
}

static int ensure_draw_capacity(Engine *e, long long keep) {
  const bool want_theta = e->keep_theta;
  if (keep <= e->out.capacity && (!want_theta || e->out.theta)) return TNUTS_OK;
  // stream-ordered allocations from the device's default pool (release threshold raised in
  // tnuts_create): a 10+ GB draw buffer is recycled by the next engine instead of paying
  // cudaMalloc/cudaFree (hundreds of ms, and a device-wide synchronisation) per sampling job
  for (double **p : {&e->out.draws, &e->out.theta}) {
    if (*p) cudaFreeAsync(*p, e->stream);
    *p = nullptr;
  }
  const size_t C = (size_t)e->st.C;
  e->out.capacity = 0;
  e->out.P = e->P;
  e->out.ncol = e->P + 3 + TNUTS_NSTAT;
  TN_CUDA(e, cudaMallocAsync((void **)&e->out.draws, sizeof(double) * (size_t)keep * e->out.ncol * C, e->stream));
  if (want_theta)
    TN_CUDA(e, cudaMallocAsync((void **)&e->out.theta, sizeof(double) * (size_t)keep * e->D * C, e->stream));
  e->out.capacity = keep;
  return TNUTS_OK;
}

static int build_small(Engine *e, const tnuts_model *m) {
  SmallData &s = e->small;
  std::memset(&s, 0, sizeof(s));
  s.n = (int)m->N;
  const int n = s.n;
  switch (m->family) {
    case TNUTS_GDEMO: {
      if (n > 16) return TNUTS_OK;
      for (int i = 0; i < n; ++i) s.a[i] = m->y[i];
      const double al = m->hyper[0], be = m->hyper[1];
      s.h[0] = al;
      s.h[1] = be;
      s.cst = al * std::log(be) - std::lgamma(al) - 0.5 * (n + 1) * kLog2Pi;
    } break;
    case TNUTS_LINREG: {
      if (n > 16) return TNUTS_OK;
      for (int i = 0; i < n; ++i) {
        s.a[i] = m->y[i];
        s.b[i] = m->X[i];
      }
      const double sd = m->hyper[0], c = m->hyper[1];
      s.h[0] = sd;
      s.h[1] = c;
      s.h[2] = 1.0 / (sd * sd);
      s.h[3] = 1.0 / c;
      s.cst_lik = -0.5 * n * kLog2Pi;
      s.cst = -kLog2Pi - 2.0 * std::log(sd) + kLog2 - kLogPi - std::log(c) + s.cst_lik;
    } break;
    case TNUTS_EIGHT_SCHOOLS: {
      if (n > 16) return TNUTS_OK;
      const double sd = m->hyper[0], c = m->hyper[1];
      s.h[0] = sd;
      s.h[1] = c;
      s.h[2] = 1.0 / (sd * sd);
      s.h[3] = 1.0 / c;
      s.cst_lik = 0.0;
      for (int j = 0; j < n; ++j) {
        s.a[j] = m->y[j];
        s.b[j] = 1.0 / (m->sigma[j] * m->sigma[j]);
        s.cst_lik += -0.5 * kLog2Pi - std::log(m->sigma[j]);
      }
      s.cst = -0.5 * kLog2Pi - std::log(sd) + kLog2 - kLogPi - std::log(c) - 0.5 * n * kLog2Pi +
              s.cst_lik;
    } break;
    case TNUTS_STD_NORMAL:
      s.cst = -0.5 * m->D * kLog2Pi;
      break;
    case TNUTS_BETA_BERNOULLI: {
      const double a = m->hyper[0], b = m->hyper[1];
      double ones = 0, zeros = 0;
      for (long long i = 0; i < m->N; ++i) (m->y[i] != 0.0 ? ones : zeros) += 1.0;
      s.n = 0;
      s.a[0] = a + ones;
      s.b[0] = b + zeros;
      s.h[0] = a;
      s.h[1] = b;
      s.h[2] = ones;
      s.h[3] = zeros;
      s.cst = -(std::lgamma(a) + std::lgamma(b) - std::lgamma(a + b));
    } break;
    case TNUTS_DIRICHLET_CATEGORICAL: {
      // per linked dimension: lp = sum_d A_d log sigma(u_d) + B_d log sigma(-u_d), u_d = theta_d - c_d
      // (stick breaking; derivation in oracle/tnuts_oracle.c)
      const int K[2] = {(int)m->hyper[0], (int)m->hyper[2]};
      const double al[2] = {m->hyper[1], m->hyper[3]};
      s.n = 0;
      for (long long i = 0; i < m->N; ++i) s.cnt[(int)m->y[i] - 1] += 1.0;
      s.h[0] = al[0];
      s.h[1] = al[1];
      s.cst = 0.0;
      int d0 = 0;
      for (int blk = 0; blk < 2; ++blk) {
        const int Kb = K[blk];
        if (Kb <= 0) continue;
        s.cst += std::lgamma(Kb * al[blk]) - Kb * std::lgamma(al[blk]);
        double tail = (al[blk] - 1.0) + (blk == 0 ? s.cnt[Kb - 1] : 0.0);
        for (int j = Kb - 2; j >= 0; --j) {
          const double cj = (al[blk] - 1.0) + (blk == 0 ? s.cnt[j] : 0.0);
          s.a[d0 + j] = cj + 1.0;
          s.b[d0 + j] = tail + 1.0 + (double)(Kb - 2 - j);
          s.c[d0 + j] = std::log((double)(Kb - 1 - j));
          tail += cj;
        }
        d0 += Kb - 1;
      }
    } break;
    default:
      break;
  }
  return TNUTS_OK;
}

}  // namespace tnuts

using namespace tnuts;

extern "C" {

int tnuts_abi_version(void) { return TNUTS_ABI_VERSION; }

const char *tnuts_last_error(const tnuts_engine *h) {
  if (h) return ((const Engine *)h)->err.c_str();
  return g_create_error.c_str();
}

int tnuts_create(const tnuts_config *cfg, tnuts_engine **out) {
  if (!cfg || !out) return set_error(nullptr, TNUTS_E_ARG, "tnuts_create: null argument");
  *out = nullptr;
  if (cfg->n_chains <= 0) return set_error(nullptr, TNUTS_E_ARG, "tnuts_create: n_chains must be > 0");
  if (cfg->max_depth < 1 || cfg->max_depth > 24)
    return set_error(nullptr, TNUTS_E_ARG, "tnuts_create: max_depth must be in 1..24");
  int ndev = 0;
  cudaError_t s = cudaGetDeviceCount(&ndev);
  if (s != cudaSuccess || ndev == 0)
    return set_error(nullptr, TNUTS_E_CUDA,
                     std::string("tnuts_create: no CUDA device (this engine has no CPU fallback): ") +
                         cudaGetErrorString(s));
  if (cfg->device < 0 || cfg->device >= ndev)
    return set_error(nullptr, TNUTS_E_ARG, "tnuts_create: bad device ordinal");
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, cfg->device);
  if (prop.major != 10)
    return set_error(nullptr, TNUTS_E_CUDA,
                     std::string("tnuts_create: libtnuts is built for sm_100a only; device is ") +
                         prop.name);
  if (cfg->metric != TNUTS_METRIC_DIAG && cfg->metric != TNUTS_METRIC_UNIT && cfg->metric != TNUTS_METRIC_DENSE)
    return set_error(nullptr, TNUTS_E_UNSUPPORTED, "tnuts_create: unknown metric");
  Engine *e = new Engine();
  e->cfg = *cfg;
  e->device = cfg->device;
  if (cudaSetDevice(e->device) != cudaSuccess || cudaStreamCreate(&e->stream) != cudaSuccess ||
      cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess) {
    delete e;
    return set_error(nullptr, TNUTS_E_CUDA, "tnuts_create: stream/event creation failed");
  }
  {  // keep freed pool memory cached for the next engine of this process
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, e->device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  RunParams &rp = e->rp;
  rp.algorithm = cfg->algorithm;
  rp.delta = cfg->delta;
  rp.max_depth = cfg->max_depth;
  rp.delta_max = cfg->delta_max;
  rp.n_adapts = cfg->n_adapts;
  rp.n_leapfrog = cfg->n_leapfrog;
  rp.lambda = cfg->lambda;
  rp.seed = cfg->seed;
  rp.chain_offset = cfg->chain_offset;
  rp.unit_metric = cfg->metric == TNUTS_METRIC_UNIT ? 1 : 0;
  window_schedule(cfg->n_adapts, rp);
  *out = (tnuts_engine *)e;
  return TNUTS_OK;
}

int tnuts_destroy(tnuts_engine *h) {
  if (!h) return TNUTS_OK;
  Engine *e = (Engine *)h;
  cudaSetDevice(e->device);
  static const bool trace = getenv("TNUTS_DESTROY_TRACE") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  const auto t0 = now();
  cudaStreamSynchronize(e->stream);
  const auto t1 = now();
  lockstep_destroy(e);
  diagnostics_destroy(e);
  comm_destroy(e);
  const auto t2 = now();
  free_pool(e->model_allocs, e->stream);
  const auto t3 = now();
  free_pool(e->state_allocs, e->stream);
  const auto t4 = now();
  if (trace)
    fprintf(stderr, "[tnuts destroy] sync %.1f ms, strategy/diag/comm %.1f ms, model frees %.1f ms, state frees %.1f ms\n",
            ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4));
  for (double *p : {e->out.draws, e->out.theta})
    if (p) cudaFreeAsync(p, e->stream);
  if (e->dyn_work) cudaFreeAsync(e->dyn_work, e->stream);
  if (e->perm) cudaFreeAsync(e->perm, e->stream);
  if (e->perm_keys) cudaFreeAsync(e->perm_keys, e->stream);
  if (e->perm_tmp) cudaFreeAsync(e->perm_tmp, e->stream);
  cudaStreamSynchronize(e->stream);
  if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
  cudaEventDestroy(e->ev0);
  cudaEventDestroy(e->ev1);
  cudaStreamDestroy(e->stream);
  delete e;
  return TNUTS_OK;
}

static int set_model_impl(tnuts_engine *h, const tnuts_model *m, bool device_ptrs);

int tnuts_set_model(tnuts_engine *h, const tnuts_model *m) { return set_model_impl(h, m, false); }

int tnuts_set_model_device(tnuts_engine *h, const tnuts_model *m) {
  Engine *e = (Engine *)h;
  if (m && m->family != TNUTS_LOGISTIC && m->family != TNUTS_HIER_LINREG)
    return set_error(e, TNUTS_E_UNSUPPORTED, "tnuts_set_model_device: data-parallel families only");
  return set_model_impl(h, m, true);
}

static int set_model_impl(tnuts_engine *h, const tnuts_model *m, bool device_ptrs) {
  Engine *e = (Engine *)h;
  if (!e || !m) return set_error(e, TNUTS_E_ARG, "tnuts_set_model: null argument");
  if (e->has_model) return set_error(e, TNUTS_E_ARG, "tnuts_set_model: model already set");
  if (m->D <= 0) return set_error(e, TNUTS_E_ARG, "tnuts_set_model: D must be > 0");
  TN_CUDA(e, cudaSetDevice(e->device));
  e->D = m->D;
  e->P = m->D;
  if (m->family == TNUTS_DIRICHLET_CATEGORICAL) {   // a K-simplex: K - 1 linked dimensions, K chain columns
    const int K1 = (int)m->hyper[0], K2 = (int)m->hyper[2];
    if (K1 < 2 || K2 < 0 || K2 == 1 || K1 + K2 > 16 || m->D != K1 - 1 + (K2 > 0 ? K2 - 1 : 0))
      return set_error(e, TNUTS_E_ARG, "dirichlet-categorical: hyper = {K1, al1, K2, al2}, D = K1 - 1 + K2 - 1");
    e->P = K1 + K2;
  }
  ModelDev &md = e->model;
  md.family = m->family;
  md.D = m->D;
  md.G = m->G;
  md.N = m->N;
  std::memcpy(md.hyper, m->hyper, sizeof(md.hyper));
  auto upload = [&](const void *src, size_t bytes, const void **dst) -> int {
    if (!src || bytes == 0) return TNUTS_OK;
    if (device_ptrs) {  // adopted, not copied: the caller keeps the device arrays alive
      *dst = src;
      return TNUTS_OK;
    }
    void *q = nullptr;
    TN_CUDA(e, cudaMallocAsync(&q, bytes, e->stream));
    e->model_allocs.push_back(q);
    TN_CUDA(e, cudaMemcpyAsync(q, src, bytes, cudaMemcpyHostToDevice, e->stream));
    *dst = q;
    return TNUTS_OK;
  };
  int rc = TNUTS_OK;
  const size_t N = (size_t)m->N;
  switch (m->family) {
    case TNUTS_GDEMO:
      if (!m->y) return set_error(e, TNUTS_E_ARG, "gdemo needs y");
      rc = upload(m->y, N * 8, (const void **)&md.y);
      break;
    case TNUTS_LINREG:
      if (!m->y || !m->X || m->D != 3) return set_error(e, TNUTS_E_ARG, "linreg needs x, y, D=3");
      rc = upload(m->y, N * 8, (const void **)&md.y);
      if (!rc) rc = upload(m->X, N * 8, (const void **)&md.X);
      break;
    case TNUTS_EIGHT_SCHOOLS:
      if (!m->y || !m->sigma || m->D != m->N + 2)
        return set_error(e, TNUTS_E_ARG, "eight schools needs y, sigma, D=N+2");
      rc = upload(m->y, N * 8, (const void **)&md.y);
      if (!rc) rc = upload(m->sigma, N * 8, (const void **)&md.sigma);
      break;
    case TNUTS_LOGISTIC:
      if (!m->y || !m->X) return set_error(e, TNUTS_E_ARG, "logistic needs X, y");
      rc = upload(m->y, N * 8, (const void **)&md.y);
      if (!rc) rc = upload(m->X, N * (size_t)m->D * 8, (const void **)&md.X);
      break;
    case TNUTS_HIER_LINREG:
      if (!m->y || !m->X || !m->group || m->D != 5 + 2 * m->G)
        return set_error(e, TNUTS_E_ARG, "hier needs x, y, group, D=5+2G");
      rc = upload(m->y, N * 8, (const void **)&md.y);
      if (!rc) rc = upload(m->X, N * 8, (const void **)&md.X);
      if (!rc) rc = upload(m->group, N * 4, (const void **)&md.group);
      break;
    case TNUTS_STD_NORMAL:
      break;
    case TNUTS_BETA_BERNOULLI:
      if (!m->y || m->D != 1 || !(m->hyper[0] > 0) || !(m->hyper[1] > 0))
        return set_error(e, TNUTS_E_ARG, "beta-bernoulli needs y, D=1, hyper = {a, b} > 0");
      rc = upload(m->y, N * 8, (const void **)&md.y);
      break;
    case TNUTS_DIRICHLET_CATEGORICAL:
      if (!m->y || !(m->hyper[1] > 0) || (m->hyper[2] > 0 && !(m->hyper[3] > 0)))
        return set_error(e, TNUTS_E_ARG, "dirichlet-categorical needs y and positive concentrations");
      for (long long i = 0; i < m->N; ++i)
        if (!(m->y[i] >= 1.0 && m->y[i] <= m->hyper[0]))
          return set_error(e, TNUTS_E_ARG, "dirichlet-categorical: observed categories must be in 1..K1");
      rc = upload(m->y, N * 8, (const void **)&md.y);
      break;
    default:
      return set_error(e, TNUTS_E_ARG, "tnuts_set_model: unknown family");
  }
  if (rc) return rc;
  build_small(e, m);
  // strategy
  const bool thread_ok = thread_strategy_supported(e);
  if (e->cfg.strategy == STRAT_THREAD && !thread_ok)
    return set_error(e, TNUTS_E_UNSUPPORTED, "thread-per-chain strategy not available for this model");
  const bool block_ok = block_strategy_supported(e);
  if (e->cfg.strategy == STRAT_BLOCK && !block_ok)
    return set_error(e, TNUTS_E_UNSUPPORTED, "one-CTA-per-chain strategy not available for this model");
  if ((m->family == TNUTS_BETA_BERNOULLI || m->family == TNUTS_DIRICHLET_CATEGORICAL) &&
      (!thread_ok || (e->cfg.strategy != STRAT_AUTO && e->cfg.strategy != STRAT_THREAD) ||
       e->cfg.algorithm == TNUTS_ALG_SGHMC || e->cfg.algorithm == TNUTS_ALG_SGLD))
    return set_error(e, TNUTS_E_UNSUPPORTED,
                     "the constrained-support test families run on the thread-per-chain strategy (NUTS / HMC / HMCDA); "
                     "simplex sizes (K1, K2) in {(2,4), (3,0), (2,0)}");
  const bool sg = e->cfg.algorithm == TNUTS_ALG_SGHMC || e->cfg.algorithm == TNUTS_ALG_SGLD;
  if (sg && (e->cfg.strategy == STRAT_THREAD || e->cfg.strategy == STRAT_BLOCK))
    return set_error(e, TNUTS_E_UNSUPPORTED, "SGHMC / SGLD run on the batched (lock-step) strategy only");
  const bool dense = e->cfg.metric == TNUTS_METRIC_DENSE;
  // DenseEuclideanMetric: the thread-per-chain strategy (tiny D) or the pump strategy (whitened state
  // machine, dense_pump.cuh: [D*D][C] arrays, D <= 256); the CTA-per-chain strategy keeps a chain's
  // vectors in registers and has no dense form
  if (dense && (sg || e->cfg.strategy == STRAT_BLOCK ||
                (!(thread_ok && e->cfg.strategy != STRAT_LOCKSTEP) && m->D > 256)))
    return set_error(e, TNUTS_E_UNSUPPORTED,
                     "DenseEuclideanMetric is implemented for NUTS / HMC / HMCDA on the thread-per-chain strategy "
                     "and, for D <= 256, on the pump strategy");
  if (sg) e->strategy = STRAT_LOCKSTEP;
  else if (e->cfg.strategy == STRAT_THREAD || (e->cfg.strategy == STRAT_AUTO && thread_ok)) e->strategy = STRAT_THREAD;
  else if (!dense && (e->cfg.strategy == STRAT_BLOCK || (e->cfg.strategy == STRAT_AUTO && block_ok))) e->strategy = STRAT_BLOCK;
  else e->strategy = STRAT_LOCKSTEP;
  // dense metric on the pump: the state machine runs the unit-metric dynamics in whitened coordinates
  if (dense && e->strategy == STRAT_LOCKSTEP) e->rp.unit_metric = 1;

  // chain state
  const size_t C = (size_t)e->cfg.n_chains, D = (size_t)m->D;
  ChainState &st = e->st;
  st.C = (int)C;
  st.D = (int)D;
  auto &pool = e->state_allocs;
  if ((rc = dev_alloc(e, pool, &st.theta, D * C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.grad, D * C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.minv, D * C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.wf_mean, D * C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.wf_m2, D * C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.lp, C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.eps, C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.da_mu, C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.da_xbar, C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.da_hbar, C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.da_eps, C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.da_m, C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.wf_n, C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.iter, C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.n_grad, C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.init_tries, C))) return rc;
  if ((rc = dev_alloc(e, pool, &st.status, C))) return rc;
  st.minv_dense = st.chol = st.wf_cov = nullptr;
  if (dense) {
    if ((rc = dev_alloc(e, pool, &st.minv_dense, D * D * C))) return rc;
    if ((rc = dev_alloc(e, pool, &st.chol, D * D * C))) return rc;
    if ((rc = dev_alloc(e, pool, &st.wf_cov, D * D * C))) return rc;
  }
  if (e->strategy != STRAT_THREAD) {  // the block strategy shares init / point-wise kernels
    if ((rc = lockstep_create(e))) return rc;
  }
  TN_CUDA(e, cudaStreamSynchronize(e->stream));
  e->has_model = true;
  return TNUTS_OK;
}

int tnuts_dimension(const tnuts_engine *h) { return h ? ((const Engine *)h)->D : TNUTS_E_ARG; }
int tnuts_num_params(const tnuts_engine *h) { return h ? ((const Engine *)h)->P : TNUTS_E_ARG; }

static const char *kInitFailMsg =
    "failed to find valid initial parameters in 1000 tries. See "
    "https://turinglang.org/docs/uri/initial-parameters for common causes and solutions. If the "
    "issue persists, please open an issue at https://github.com/TuringLang/Turing.jl/issues";

int tnuts_init(tnuts_engine *h, const double *theta0) {
  Engine *e = (Engine *)h;
  if (!e || !e->has_model) return set_error(e, TNUTS_E_ARG, "tnuts_init: set a model first");
  TN_CUDA(e, cudaSetDevice(e->device));
  double *t0 = nullptr;
  const size_t bytes = sizeof(double) * (size_t)e->st.C * e->D;
  if (theta0) {
    TN_CUDA(e, cudaMalloc((void **)&t0, bytes));
    TN_CUDA(e, cudaMemcpyAsync(t0, theta0, bytes, cudaMemcpyHostToDevice, e->stream));
  }
  cudaEventRecord(e->ev0, e->stream);
  int rc = (e->strategy == STRAT_THREAD) ? thread_init(e, t0) : lockstep_init(e, t0);
  cudaEventRecord(e->ev1, e->stream);
  cudaError_t s = cudaStreamSynchronize(e->stream);
  {
    float ms = 0;
    if (s == cudaSuccess && cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess)
      e->last_init_seconds = ms * 1e-3;
  }
  if (t0) cudaFree(t0);
  if (rc) return rc;
  if (s != cudaSuccess) return set_error(e, TNUTS_E_CUDA, std::string("tnuts_init: ") + cudaGetErrorString(s));
  std::vector<int> status((size_t)e->st.C);
  TN_CUDA(e, cudaMemcpy(status.data(), e->st.status, sizeof(int) * status.size(), cudaMemcpyDeviceToHost));
  for (int v : status)
    if (v != 0) return set_error(e, TNUTS_E_INIT, kInitFailMsg);
  e->inited = true;
  e->n_kept = 0;
  e->host_iter = 0;
  e->perm_valid = false;
  return TNUTS_OK;
}

// Runs the transitions in segments that end on kept-draw boundaries.  With host_draws != NULL
// every finished segment's draws are copied device->host on a second stream while the next
// segment computes (the draw index is the slowest dimension, so a segment is contiguous).
static int run_segments(Engine *e, int64_t n_transitions, int64_t first_keep, int64_t thin,
                        double *host_draws, int64_t chunk_draws, int64_t *n_kept, int64_t host_stride = 0) {
  if (!e || !e->inited) return set_error(e, TNUTS_E_ARG, "tnuts_run: call tnuts_init first");
  if (n_transitions < 0 || thin < 1) return set_error(e, TNUTS_E_ARG, "tnuts_run: bad n_transitions/thin");
  TN_CUDA(e, cudaSetDevice(e->device));
  long long keep = 0;
  if (first_keep >= 1 && n_transitions >= first_keep) keep = (n_transitions - first_keep) / thin + 1;
  int rc = ensure_draw_capacity(e, keep > 0 ? keep : 1);
  if (rc) return rc;
  if (host_draws && !e->copy_stream) TN_CUDA(e, cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
  const size_t C = (size_t)e->st.C, row = (size_t)e->out.ncol * C;
  // host layout [keep][ncol][hstride]: hstride > C when this engine's chains are a slab of a wider
  // chain array (several GPUs filling one (iterations x names x chains) array)
  const size_t hstride = host_stride > 0 ? (size_t)host_stride : C;
  if (host_draws && hstride < C) return set_error(e, TNUTS_E_ARG, "tnuts_run_streamed: host stride < n_chains");
  double *const base = e->out.draws, *const tbase = e->out.theta;
  // the lock-step strategy advances chains asynchronously (each at its own transition), so a
  // prefix of the draws is not complete before the end: one segment, one copy
  if (chunk_draws <= 0 || !host_draws || e->strategy != STRAT_THREAD) chunk_draws = keep > 0 ? keep : 1;
  TN_CUDA(e, cudaEventRecord(e->ev0, e->stream));
  long long t_done = 0, k_done = 0;
  std::vector<cudaEvent_t> evs;
  while (t_done < n_transitions) {
    long long k_end = std::min<long long>(keep, k_done + chunk_draws);
    // last transition of this segment: the one that produces draw k_end - 1 (or the very end)
    long long t_end = (keep == 0 || k_end >= keep) ? n_transitions : first_keep + (k_end - 1) * thin;
    if (k_done >= keep) t_end = n_transitions;
    if (e->strategy == STRAT_THREAD && e->sort_chains && e->cfg.algorithm == TNUTS_ALG_NUTS && !e->st.minv_dense) {
      // thread-per-chain: once the step sizes are final, sort the chains by step size so
      // that a warp's 32 chains grow trees of similar length; cut the segment at that point
      const long long to_adapt_end = e->cfg.n_adapts - (e->host_iter + t_done);
      if (!e->perm_valid && to_adapt_end <= 0) {
        if ((rc = thread_sort_chains(e))) break;
      } else if (!e->perm_valid && t_done + to_adapt_end < t_end) {
        t_end = t_done + to_adapt_end;
        // draws kept up to and including t_end
        k_end = (t_end >= first_keep && keep > 0) ? std::min<long long>(keep, (t_end - first_keep) / thin + 1) : k_done;
      }
    }
    e->rp.n_transitions = t_end - t_done;
    e->rp.thin = thin;
    e->rp.first_keep = (k_done < keep) ? (first_keep + k_done * thin - t_done) : 0;
    e->out.draws = base + (size_t)k_done * row;
    e->out.theta = tbase ? tbase + (size_t)k_done * e->D * C : nullptr;
    rc = (e->strategy == STRAT_THREAD) ? thread_run(e)
         : (e->strategy == STRAT_BLOCK) ? block_run(e) : lockstep_run(e);
    if (rc) break;
    if (host_draws && k_end > k_done) {
      cudaEvent_t ev;
      if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventRecord(ev, e->stream) != cudaSuccess ||
          cudaStreamWaitEvent(e->copy_stream, ev, 0) != cudaSuccess ||
          (hstride == C
               ? cudaMemcpyAsync(host_draws + (size_t)k_done * row, base + (size_t)k_done * row,
                                 8 * (size_t)(k_end - k_done) * row, cudaMemcpyDeviceToHost, e->copy_stream)
               : cudaMemcpy2DAsync(host_draws + (size_t)k_done * e->out.ncol * hstride, 8 * hstride,
                                   base + (size_t)k_done * row, 8 * C, 8 * C,
                                   (size_t)(k_end - k_done) * e->out.ncol, cudaMemcpyDeviceToHost,
                                   e->copy_stream)) != cudaSuccess) {
        rc = set_error(e, TNUTS_E_CUDA, "tnuts_run_streamed: D2H enqueue failed");
        break;
      }
      evs.push_back(ev);
    }
    t_done = t_end;
    k_done = k_end > k_done ? k_end : k_done;
  }
  e->out.draws = base;
  e->out.theta = tbase;
  if (!rc) {
    cudaError_t s = cudaEventRecord(e->ev1, e->stream);
    if (s == cudaSuccess) s = cudaStreamSynchronize(e->stream);
    if (s == cudaSuccess && e->copy_stream) s = cudaStreamSynchronize(e->copy_stream);
    if (s != cudaSuccess) rc = set_error(e, TNUTS_E_CUDA, std::string("tnuts_run: ") + cudaGetErrorString(s));
  }
  for (cudaEvent_t ev : evs) cudaEventDestroy(ev);
  if (rc) return rc;
  float ms = 0;
  TN_CUDA(e, cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  e->last_run_seconds = ms * 1e-3;
  e->n_kept = keep;
  e->host_iter += n_transitions;
  if (n_kept) *n_kept = keep;
  return TNUTS_OK;
}

int tnuts_run(tnuts_engine *h, int64_t n_transitions, int64_t first_keep, int64_t thin,
              int64_t *n_kept) {
  return run_segments((Engine *)h, n_transitions, first_keep, thin, nullptr, 0, n_kept);
}

int tnuts_run_streamed(tnuts_engine *h, int64_t n_transitions, int64_t first_keep, int64_t thin,
                       double *host_draws, int64_t chunk_draws, int64_t *n_kept) {
  Engine *e = (Engine *)h;
  if (!host_draws) return set_error(e, TNUTS_E_ARG, "tnuts_run_streamed: host_draws is NULL");
  return run_segments(e, n_transitions, first_keep, thin, host_draws, chunk_draws, n_kept);
}

int tnuts_run_streamed_strided(tnuts_engine *h, int64_t n_transitions, int64_t first_keep, int64_t thin,
                               double *host_draws, int64_t host_chain_stride, int64_t chunk_draws,
                               int64_t *n_kept) {
  Engine *e = (Engine *)h;
  if (!host_draws) return set_error(e, TNUTS_E_ARG, "tnuts_run_streamed_strided: host_draws is NULL");
  return run_segments(e, n_transitions, first_keep, thin, host_draws, chunk_draws, n_kept, host_chain_stride);
}

int tnuts_num_columns(const tnuts_engine *h) {
  const Engine *e = (const Engine *)h;
  return e ? e->P + 3 + TNUTS_NSTAT : TNUTS_E_ARG;
}

int tnuts_set_option(tnuts_engine *h, const char *name, double value) {
  Engine *e = (Engine *)h;
  if (!e || !name) return set_error(e, TNUTS_E_ARG, "tnuts_set_option: null argument");
  const std::string n(name);
  if (n == "keep_theta") e->keep_theta = value != 0.0;
  else if (n == "sort_chains") e->sort_chains = value != 0.0;
  else if (n == "dyn_sched") e->dyn_sched = value != 0.0;
  else if (n == "ck_smem") e->ck_smem = value != 0.0;
  else if (n == "logistic_tc") e->logistic_tc = value != 0.0;
  else if (n == "time_kernels") e->time_kernels = value != 0.0;
  else if (n == "persist") e->persist = value != 0.0;
  else if (n == "compact") e->compact = value != 0.0;
  else return set_error(e, TNUTS_E_ARG, "tnuts_set_option: unknown option " + n);
  return TNUTS_OK;
}

int tnuts_fetch_draws(tnuts_engine *h, double *draws) {
  Engine *e = (Engine *)h;
  if (!e || !e->inited || !draws) return set_error(e, TNUTS_E_ARG, "tnuts_fetch_draws: bad call");
  TN_CUDA(e, cudaSetDevice(e->device));
  const size_t K = (size_t)e->n_kept, C = (size_t)e->st.C;
  if (K == 0) return TNUTS_OK;
  TN_CUDA(e, cudaMemcpyAsync(draws, e->out.draws, 8 * K * e->out.ncol * C, cudaMemcpyDeviceToHost, e->stream));
  TN_CUDA(e, cudaStreamSynchronize(e->stream));
  return TNUTS_OK;
}

int tnuts_fetch(tnuts_engine *h, double *params, double *logp, double *stats, double *theta) {
  Engine *e = (Engine *)h;
  if (!e || !e->inited) return set_error(e, TNUTS_E_ARG, "tnuts_fetch: nothing to fetch");
  TN_CUDA(e, cudaSetDevice(e->device));
  const size_t K = (size_t)e->n_kept, C = (size_t)e->st.C, P = (size_t)e->P;
  if (K == 0) return TNUTS_OK;
  const size_t pitch = 8 * (size_t)e->out.ncol * C;  // one draw
  auto cols = [&](double *dst, size_t col0, size_t ncols) -> int {
    if (!dst) return TNUTS_OK;
    TN_CUDA(e, cudaMemcpy2DAsync(dst, 8 * ncols * C, e->out.draws + col0 * C, pitch, 8 * ncols * C, K,
                                 cudaMemcpyDeviceToHost, e->stream));
    return TNUTS_OK;
  };
  int rc;
  if ((rc = cols(params, 0, P))) return rc;
  if ((rc = cols(logp, P, 3))) return rc;
  if ((rc = cols(stats, P + 3, TNUTS_NSTAT))) return rc;
  if (theta) {
    if (!e->out.theta) return set_error(e, TNUTS_E_ARG, "tnuts_fetch: theta was not kept (set option keep_theta)");
    TN_CUDA(e, cudaMemcpyAsync(theta, e->out.theta, 8 * K * e->D * C, cudaMemcpyDeviceToHost, e->stream));
  }
  TN_CUDA(e, cudaStreamSynchronize(e->stream));
  return TNUTS_OK;
}

// [D][C] device -> [C][D] host
static int fetch_dc(Engine *e, const double *dev, double *host) {
  const size_t C = (size_t)e->st.C, D = (size_t)e->D;
  std::vector<double> tmp(C * D);
  TN_CUDA(e, cudaMemcpy(tmp.data(), dev, 8 * C * D, cudaMemcpyDeviceToHost));
  for (size_t d = 0; d < D; ++d)
    for (size_t c = 0; c < C; ++c) host[c * D + d] = tmp[d * C + c];
  return TNUTS_OK;
}

int tnuts_get_theta(tnuts_engine *h, double *theta) {
  Engine *e = (Engine *)h;
  if (!e || !e->inited || !theta) return set_error(e, TNUTS_E_ARG, "tnuts_get_theta: bad call");
  TN_CUDA(e, cudaSetDevice(e->device));
  return fetch_dc(e, e->st.theta, theta);
}

int tnuts_get_inv_metric(tnuts_engine *h, double *minv) {
  Engine *e = (Engine *)h;
  if (!e || !e->inited || !minv) return set_error(e, TNUTS_E_ARG, "tnuts_get_inv_metric: bad call");
  TN_CUDA(e, cudaSetDevice(e->device));
  if (e->st.minv_dense && e->strategy == STRAT_LOCKSTEP) {
    // whitened pump: st.minv is the identity of the phi space; report the diagonal of M^-1
    const size_t C = (size_t)e->st.C, D = (size_t)e->D;
    std::vector<double> tmp(D * C);
    for (size_t d = 0; d < D; ++d)
      TN_CUDA(e, cudaMemcpy(tmp.data() + d * C, e->st.minv_dense + (d * D + d) * C, 8 * C, cudaMemcpyDeviceToHost));
    for (size_t c = 0; c < C; ++c)
      for (size_t d = 0; d < D; ++d) minv[c * D + d] = tmp[d * C + c];
    return TNUTS_OK;
  }
  return fetch_dc(e, e->st.minv, minv);
}

int tnuts_get_dense_inv_metric(tnuts_engine *h, double *minv) {
  Engine *e = (Engine *)h;
  if (!e || !e->inited || !minv) return set_error(e, TNUTS_E_ARG, "tnuts_get_dense_inv_metric: bad call");
  if (!e->st.minv_dense) return set_error(e, TNUTS_E_ARG, "tnuts_get_dense_inv_metric: the metric is not dense");
  TN_CUDA(e, cudaSetDevice(e->device));
  const size_t C = (size_t)e->st.C, DD = (size_t)e->D * e->D;
  std::vector<double> tmp(DD * C);   // device layout [D*D][C] -> [C][D][D]
  TN_CUDA(e, cudaMemcpy(tmp.data(), e->st.minv_dense, 8 * DD * C, cudaMemcpyDeviceToHost));
  for (size_t c = 0; c < C; ++c)
    for (size_t i = 0; i < DD; ++i) minv[c * DD + i] = tmp[i * C + c];
  return TNUTS_OK;
}

int tnuts_get_stepsize(tnuts_engine *h, double *eps) {
  Engine *e = (Engine *)h;
  if (!e || !e->inited || !eps) return set_error(e, TNUTS_E_ARG, "tnuts_get_stepsize: bad call");
  TN_CUDA(e, cudaSetDevice(e->device));
  TN_CUDA(e, cudaMemcpy(eps, e->st.eps, 8 * (size_t)e->st.C, cudaMemcpyDeviceToHost));
  return TNUTS_OK;
}

// scratch helper: host -> device copies for the point-wise entry points
struct Scratch {
  std::vector<void *> p;
  ~Scratch() {
    for (void *q : p) cudaFree(q);
  }
  double *put(Engine *e, const double *src, size_t n, int *rc) {
    double *q = nullptr;
    if (cudaMalloc((void **)&q, 8 * (n ? n : 1)) != cudaSuccess) {
      *rc = set_error(e, TNUTS_E_CUDA, "scratch cudaMalloc failed");
      return nullptr;
    }
    p.push_back(q);
    if (src && cudaMemcpyAsync(q, src, 8 * n, cudaMemcpyHostToDevice, e->stream) != cudaSuccess) {
      *rc = set_error(e, TNUTS_E_CUDA, "scratch H2D failed");
      return nullptr;
    }
    return q;
  }
};

int tnuts_logdensity_and_gradient(tnuts_engine *h, const double *theta, int64_t n, double *lp,
                                  double *grad) {
  Engine *e = (Engine *)h;
  if (!e || !e->has_model || !theta || !lp || n < 0)
    return set_error(e, TNUTS_E_ARG, "tnuts_logdensity_and_gradient: bad call");
  if (n == 0) return TNUTS_OK;
  TN_CUDA(e, cudaSetDevice(e->device));
  Scratch s;
  int rc = TNUTS_OK;
  const size_t D = (size_t)e->D;
  double *th = s.put(e, theta, (size_t)n * D, &rc);
  double *dlp = s.put(e, nullptr, (size_t)n, &rc);
  double *dg = grad ? s.put(e, nullptr, (size_t)n * D, &rc) : nullptr;
  if (rc) return rc;
  rc = (e->strategy == STRAT_THREAD) ? thread_point_lpgrad(e, th, n, dlp, dg)
                                     : lockstep_point_lpgrad(e, th, n, dlp, dg);
  if (rc) return rc;
  TN_CUDA(e, cudaMemcpyAsync(lp, dlp, 8 * (size_t)n, cudaMemcpyDeviceToHost, e->stream));
  if (grad) TN_CUDA(e, cudaMemcpyAsync(grad, dg, 8 * (size_t)n * D, cudaMemcpyDeviceToHost, e->stream));
  TN_CUDA(e, cudaStreamSynchronize(e->stream));
  return TNUTS_OK;
}

int tnuts_logdensity(tnuts_engine *h, const double *theta, int64_t n, double *lp) {
  return tnuts_logdensity_and_gradient(h, theta, n, lp, nullptr);
}

int tnuts_constrain(tnuts_engine *h, const double *theta, int64_t n, double *params, double *logp) {
  Engine *e = (Engine *)h;
  if (!e || !e->has_model || !theta || !params || !logp || n < 0)
    return set_error(e, TNUTS_E_ARG, "tnuts_constrain: bad call");
  if (n == 0) return TNUTS_OK;
  TN_CUDA(e, cudaSetDevice(e->device));
  Scratch s;
  int rc = TNUTS_OK;
  const size_t D = (size_t)e->D;
  double *th = s.put(e, theta, (size_t)n * D, &rc);
  double *dp = s.put(e, nullptr, (size_t)n * e->P, &rc);
  double *dl = s.put(e, nullptr, (size_t)n * 3, &rc);
  if (rc) return rc;
  rc = (e->strategy == STRAT_THREAD) ? thread_point_constrain(e, th, n, dp, dl)
                                     : lockstep_point_constrain(e, th, n, dp, dl);
  if (rc) return rc;
  TN_CUDA(e, cudaMemcpyAsync(params, dp, 8 * (size_t)n * e->P, cudaMemcpyDeviceToHost, e->stream));
  TN_CUDA(e, cudaMemcpyAsync(logp, dl, 8 * (size_t)n * 3, cudaMemcpyDeviceToHost, e->stream));
  TN_CUDA(e, cudaStreamSynchronize(e->stream));
  return TNUTS_OK;
}

int tnuts_leapfrog(tnuts_engine *h, const double *theta, const double *r, const double *minv,
                   double eps, int32_t L, int64_t n, double *traj_theta, double *traj_r,
                   double *traj_H) {
  Engine *e = (Engine *)h;
  if (!e || !e->has_model || !theta || !r || !minv || L < 0 || n < 0 || !traj_theta || !traj_r || !traj_H)
    return set_error(e, TNUTS_E_ARG, "tnuts_leapfrog: bad call");
  if (n == 0) return TNUTS_OK;
  TN_CUDA(e, cudaSetDevice(e->device));
  Scratch s;
  int rc = TNUTS_OK;
  const size_t D = (size_t)e->D, T = (size_t)n * (L + 1);
  double *th = s.put(e, theta, (size_t)n * D, &rc);
  double *dr = s.put(e, r, (size_t)n * D, &rc);
  double *dm = s.put(e, minv, D, &rc);
  double *tt = s.put(e, nullptr, T * D, &rc);
  double *tr = s.put(e, nullptr, T * D, &rc);
  double *tH = s.put(e, nullptr, T, &rc);
  if (rc) return rc;
  rc = (e->strategy == STRAT_THREAD) ? thread_point_leapfrog(e, th, dr, dm, eps, L, n, tt, tr, tH)
                                     : lockstep_point_leapfrog(e, th, dr, dm, eps, L, n, tt, tr, tH);
  if (rc) return rc;
  TN_CUDA(e, cudaMemcpyAsync(traj_theta, tt, 8 * T * D, cudaMemcpyDeviceToHost, e->stream));
  TN_CUDA(e, cudaMemcpyAsync(traj_r, tr, 8 * T * D, cudaMemcpyDeviceToHost, e->stream));
  TN_CUDA(e, cudaMemcpyAsync(traj_H, tH, 8 * T, cudaMemcpyDeviceToHost, e->stream));
  TN_CUDA(e, cudaStreamSynchronize(e->stream));
  return TNUTS_OK;
}

// ---- state blob: header + every ChainState array, in a fixed order ----
struct BlobHeader {
  uint32_t magic, version;
  int32_t C, D, family, algorithm;
  int64_t n_adapts;
  uint64_t seed;
  int32_t metric, chain_offset;
};
static constexpr uint32_t kBlobVersion = 2u;
static constexpr uint32_t kBlobMagic = 0x544e5554u;  // "TUNT"

// the metric word of the header: a dense-metric blob also names the strategy that wrote it (the pump
// strategy keeps the whitened start state phi / grad_phi where the thread-per-chain strategy keeps
// the unused diagonal Welford M2 / grad_theta)
static int32_t blob_metric_word(const Engine *e) {
  return e->cfg.metric | (e->st.minv_dense ? ((int32_t)e->strategy << 8) : 0);
}

static size_t blob_bytes(const Engine *e) {
  const size_t C = (size_t)e->st.C, D = (size_t)e->D;
  const size_t dense = e->st.minv_dense ? 8 * 3 * D * D * C : 0;
  return sizeof(BlobHeader) + 8 * (5 * D * C + 6 * C) + 8 * 5 * C + dense;
}

int tnuts_get_state(tnuts_engine *h, void *blob, size_t *nbytes) {
  Engine *e = (Engine *)h;
  if (!e || !e->inited || !nbytes) return set_error(e, TNUTS_E_ARG, "tnuts_get_state: bad call");
  const size_t need = blob_bytes(e);
  if (!blob) {
    *nbytes = need;
    return TNUTS_OK;
  }
  if (*nbytes < need) return set_error(e, TNUTS_E_ARG, "tnuts_get_state: buffer too small");
  TN_CUDA(e, cudaSetDevice(e->device));
  const size_t C = (size_t)e->st.C, D = (size_t)e->D;
  BlobHeader hd{kBlobMagic, kBlobVersion, (int32_t)C, (int32_t)D, e->model.family, e->cfg.algorithm,
                e->cfg.n_adapts, e->cfg.seed, blob_metric_word(e), e->cfg.chain_offset};
  char *p = (char *)blob;
  std::memcpy(p, &hd, sizeof(hd));
  p += sizeof(hd);
  const ChainState &st = e->st;
  const void *arrs[] = {st.theta, st.grad, st.minv, st.wf_mean, st.wf_m2};
  for (const void *a : arrs) {
    TN_CUDA(e, cudaMemcpy(p, a, 8 * D * C, cudaMemcpyDeviceToHost));
    p += 8 * D * C;
  }
  const void *sc[] = {st.lp, st.eps, st.da_mu, st.da_xbar, st.da_hbar, st.da_eps,
                      st.da_m, st.wf_n, st.iter, st.n_grad};
  for (const void *a : sc) {
    TN_CUDA(e, cudaMemcpy(p, a, 8 * C, cudaMemcpyDeviceToHost));
    p += 8 * C;
  }
  // 11th scalar slot: status/init_tries are not part of the resumable state
  std::memset(p, 0, 8 * C);
  p += 8 * C;
  if (st.minv_dense) {
    const void *dn[] = {st.minv_dense, st.chol, st.wf_cov};
    for (const void *a : dn) {
      TN_CUDA(e, cudaMemcpy(p, a, 8 * D * D * C, cudaMemcpyDeviceToHost));
      p += 8 * D * D * C;
    }
  }
  *nbytes = need;
  return TNUTS_OK;
}

int tnuts_set_state(tnuts_engine *h, const void *blob, size_t nbytes) {
  Engine *e = (Engine *)h;
  if (!e || !e->has_model || !blob) return set_error(e, TNUTS_E_ARG, "tnuts_set_state: bad call");
  if (nbytes < blob_bytes(e)) return set_error(e, TNUTS_E_ARG, "tnuts_set_state: blob too small");
  BlobHeader hd;
  std::memcpy(&hd, blob, sizeof(hd));
  if (hd.magic != kBlobMagic || hd.C != e->st.C || hd.D != e->D || hd.family != e->model.family)
    return set_error(e, TNUTS_E_ARG, "tnuts_set_state: blob does not match this engine");
  if (hd.version != kBlobVersion)
    return set_error(e, TNUTS_E_ARG, "tnuts_set_state: blob was written by another version of the library");
  // The arrays mean different things per sampler (SGHMC keeps its velocity where NUTS keeps the
  // Welford mean) and per metric, and the random streams are addressed by (seed, global chain id):
  // a blob from another sampler / metric / seed / shard would continue a different chain silently.
  if (hd.algorithm != e->cfg.algorithm || hd.metric != blob_metric_word(e))
    return set_error(e, TNUTS_E_ARG, "tnuts_set_state: blob comes from another sampler or metric (or, dense "
                                     "metric, from another execution strategy)");
  if (hd.seed != e->cfg.seed || hd.chain_offset != e->cfg.chain_offset)
    return set_error(e, TNUTS_E_ARG, "tnuts_set_state: blob comes from another seed or chain shard "
                                     "(create the engine with the blob's seed and chain_offset)");
  TN_CUDA(e, cudaSetDevice(e->device));
  const size_t C = (size_t)e->st.C, D = (size_t)e->D;
  const char *p = (const char *)blob + sizeof(hd);
  ChainState &st = e->st;
  void *arrs[] = {st.theta, st.grad, st.minv, st.wf_mean, st.wf_m2};
  for (void *a : arrs) {
    TN_CUDA(e, cudaMemcpy(a, p, 8 * D * C, cudaMemcpyHostToDevice));
    p += 8 * D * C;
  }
  void *sc[] = {st.lp, st.eps, st.da_mu, st.da_xbar, st.da_hbar, st.da_eps,
                st.da_m, st.wf_n, st.iter, st.n_grad};
  for (void *a : sc) {
    TN_CUDA(e, cudaMemcpy(a, p, 8 * C, cudaMemcpyHostToDevice));
    p += 8 * C;
  }
  if (st.minv_dense) {
    p += 8 * C;  // the unused 11th scalar slot
    void *dn[] = {st.minv_dense, st.chol, st.wf_cov};
    for (void *a : dn) {
      TN_CUDA(e, cudaMemcpy(a, p, 8 * D * D * C, cudaMemcpyHostToDevice));
      p += 8 * D * D * C;
    }
  }
  TN_CUDA(e, cudaMemset(st.status, 0, sizeof(int) * C));
  {
    long long it0 = 0;
    TN_CUDA(e, cudaMemcpy(&it0, st.iter, sizeof(long long), cudaMemcpyDeviceToHost));
    e->host_iter = it0;
    e->perm_valid = false;
  }
  e->inited = true;
  e->n_kept = 0;
  return TNUTS_OK;
}

int tnuts_diagnostics(tnuts_engine *h, double *rhat, double *ess_bulk) {
  Engine *e = (Engine *)h;
  if (!e || !e->inited || e->n_kept < 4)
    return set_error(e, TNUTS_E_ARG, "tnuts_diagnostics: need >= 4 kept draws");
  TN_CUDA(e, cudaSetDevice(e->device));
  return diagnostics_run(e, rhat, ess_bulk);
}

static int64_t sum_ll(const Engine *e, const long long *dev) {
  std::vector<long long> v((size_t)e->st.C);
  if (cudaMemcpy(v.data(), dev, 8 * v.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  long long s = 0;
  for (long long x : v) s += x;
  return s;
}

int64_t tnuts_num_grad_evals(const tnuts_engine *h) {
  const Engine *e = (const Engine *)h;
  if (!e || !e->has_model) return -1;
  cudaSetDevice(e->device);
  return sum_ll(e, e->st.n_grad);
}

__global__ void k_count_divergent(const double *draws, long long K, int ncol, int col, int C,
                                  unsigned long long *out) {
  unsigned long long n = 0;
  const long long tot = K * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot;
       i += (long long)gridDim.x * blockDim.x) {
    const long long k = i / C, c = i % C;
    n += draws[((size_t)k * ncol + col) * C + c] != 0.0;
  }
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(out, n);
}

int64_t tnuts_num_divergences(const tnuts_engine *h) {
  Engine *e = (Engine *)h;
  if (!e || !e->inited || e->n_kept == 0) return 0;
  cudaSetDevice(e->device);
  unsigned long long *d = nullptr, hres = 0;
  if (cudaMalloc((void **)&d, 8) != cudaSuccess) return -1;
  cudaMemsetAsync(d, 0, 8, e->stream);
  k_count_divergent<<<148 * 4, 256, 0, e->stream>>>(e->out.draws, e->n_kept, e->out.ncol,
                                                   e->P + 3 + ST_NUMERR, e->st.C, d);
  e->launches += 1;
  cudaMemcpyAsync(&hres, d, 8, cudaMemcpyDeviceToHost, e->stream);
  const cudaError_t s = cudaStreamSynchronize(e->stream);
  cudaFree(d);
  return s == cudaSuccess ? (int64_t)hres : -1;
}

int64_t tnuts_num_kernel_launches(const tnuts_engine *h) {
  return h ? ((const Engine *)h)->launches : -1;
}
double tnuts_device_seconds(const tnuts_engine *h) {
  return h ? ((const Engine *)h)->last_run_seconds : -1.0;
}
double tnuts_init_device_seconds(const tnuts_engine *h) {
  return h ? ((const Engine *)h)->last_init_seconds : -1.0;
}
int tnuts_init_tries_max(const tnuts_engine *h) {
  const Engine *e = (const Engine *)h;
  if (!e || !e->has_model) return -1;
  cudaSetDevice(e->device);
  std::vector<int> v((size_t)e->st.C);
  if (cudaMemcpy(v.data(), e->st.init_tries, 4 * v.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  int m = 0;
  for (int x : v) m = x > m ? x : m;
  return m;
}

int tnuts_kernel_times(const tnuts_engine *h, double *seconds, int64_t *launches, int64_t *sampled,
                       int64_t *tile_launches, double *persist_seconds, int64_t *persist_pumps) {
  const Engine *e = (const Engine *)h;
  if (!e) return TNUTS_E_ARG;
  if (seconds) *seconds = e->strategy == STRAT_LOCKSTEP ? e->kernel_seconds : e->last_run_seconds;
  if (launches) *launches = e->strategy == STRAT_LOCKSTEP ? e->kernel_launches : 1;
  if (sampled) *sampled = e->strategy == STRAT_LOCKSTEP ? e->kernel_sampled : 1;
  if (tile_launches) *tile_launches = e->kernel_tile_launches;
  if (persist_seconds) *persist_seconds = e->persist_seconds;
  if (persist_pumps) *persist_pumps = e->persist_pumps;
  return TNUTS_OK;
}

int tnuts_bench_gradient(tnuts_engine *h, int32_t reps, double *seconds) {
  Engine *e = (Engine *)h;
  if (!e || !e->inited || reps < 1 || !seconds) return set_error(e, TNUTS_E_ARG, "tnuts_bench_gradient: bad call");
  if (e->strategy == STRAT_THREAD)
    return set_error(e, TNUTS_E_UNSUPPORTED, "tnuts_bench_gradient: not available for the thread-per-chain strategy");
  TN_CUDA(e, cudaSetDevice(e->device));
  return lockstep_bench_gradient(e, reps, seconds);
}

int tnuts_host_alloc(void **p, size_t nbytes) {
  if (!p) return TNUTS_E_ARG;
  return cudaMallocHost(p, nbytes) == cudaSuccess ? TNUTS_OK : TNUTS_E_CUDA;
}
int tnuts_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? TNUTS_OK : TNUTS_E_CUDA; }

}  // extern "C"
