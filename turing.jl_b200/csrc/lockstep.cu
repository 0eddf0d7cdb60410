// lockstep.cu -- host driver of the lock-step strategy: one launch per leapfrog phase for all
// chains, the NUTS doubling schedule driven from the host (one small D2H per doubling tells it
// whether any chain still runs), gradient kernels per family, init / step-size search,
// point-wise LogDensityProblems entry points.
#include "lockstep.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "engine.cuh"
#include "lockstep_kernels.cuh"
#include "logistic_tc.cuh"
#include "dense_pump.cuh"
#include "lpgrad_kernels.cuh"

#ifdef TNUTS_WITH_NCCL
#include <nccl.h>
#endif

namespace tnuts {

int comm_allreduce_sum(Engine *e, double *buf, size_t n);  // comm.cu (no-op when world == 1)
namespace tc { struct Data; }
// persist.cu: the tail of a run as one persistent cooperative kernel (tc_persist.cuh)
int tc_persist_run(Engine *e, const tc::Data &td, const PersistView &pv, double *part, int stride,
                   int *count_done_dev, int *h_done, long long pump0, long long max_pumps, int listed_done, int n_max,
                   long long *pumps, int *n_done, int *reason, int *sched_done);

template <class T>
static int ls_alloc(Engine *e, std::vector<void *> *pool, T **p, size_t n) {
  void *q = nullptr;
  TN_CUDA(e, cudaMallocAsync(&q, (n ? n : 1) * sizeof(T), e->stream));   // stream-ordered pool, like engine.cu
  TN_CUDA(e, cudaMemsetAsync(q, 0, (n ? n : 1) * sizeof(T), e->stream));
  pool->push_back(q);
  *p = (T *)q;
  return TNUTS_OK;
}

static inline int tiles(int C) { return (C + kTileC - 1) / kTileC; }

// ------------------------------------------------------------------------------------------
// gradient dispatch: lp[c], g[d][c] for the chains in the work list (NULL = all n_all columns)
// ------------------------------------------------------------------------------------------
struct GradWork {  // workspaces sized for `cap` columns
  double *part = nullptr, *lik = nullptr;
  size_t part_elems = 0;
  int cap = 0, chunks = 0, rows_per_chunk = 0, stride = 0;
};

// tuning knob for experiments: TNUTS_LOGISTIC_GROUPS=1 selects the single-group kernel
int g_logistic_groups = [] { const char *v = getenv("TNUTS_LOGISTIC_GROUPS"); return v ? atoi(v) : 2; }();

struct LogiPlan {
  int nwn, nwm, mt, tr, tch, chunks, rpc, ng;
  size_t smem;
};

static int logistic_plan(const ModelDev &m, int cap, LogiPlan *p) {
  const int D = m.D;
  const int mtiles = (D + 7) / 8;
  if (mtiles <= 16) {
    p->nwn = 8;
    p->nwm = 1;
    p->mt = mtiles <= 1 ? 1 : mtiles <= 2 ? 2 : mtiles <= 4 ? 4 : mtiles <= 8 ? 8 : mtiles <= 13 ? 13 : 16;
    p->tr = 64;
  } else if (mtiles <= 32) {
    p->nwn = 4;
    p->nwm = 2;
    p->mt = 16;
    p->tr = 32;
  } else {
    return -1;
  }
  p->tch = 8 * p->nwn;
  const int Kp = (D + 3) & ~3, LDX = ld_pad(Kp);
  auto smem_of = [&](int tr, int ng) {
    return sizeof(double) * ((size_t)Kp * (p->tch + 4) +
                             (size_t)ng * ((size_t)2 * tr * LDX + (size_t)tr * (p->tch + 4) + 2 * tr));
  };
  // preferred: two ping-pong groups of 8 warps with 32-row tiles (16 warps per SM)
  p->ng = 2;
  p->tr = 32;
  p->smem = smem_of(32, 2);
  if (p->smem > 227 * 1024 || p->nwm == 2 || p->mt > 13 || g_logistic_groups == 1) {
    p->ng = 1;
    p->tr = (p->nwm == 2) ? 32 : 64;
    p->smem = smem_of(p->tr, 1);
    if (p->smem > 227 * 1024 && p->tr == 64) {
      p->tr = 32;
      p->smem = smem_of(32, 1);
    }
  }
  // row chunks: a FIXED cut of the rows (a function of N and the tile height only), so the
  // summation order of a chain's gradient never depends on how many chains run beside it:
  // same seed => bit-identical chains for any chain count / sharding.  296 chunks = 148 CTAs
  // of two groups: a single chain tile already fills the chip, 64 tiles make 64 full waves.
  (void)cap;
  const long long maxchunks = std::max<long long>(1, (m.N + p->tr - 1) / p->tr);
  int want = (int)std::min<long long>(maxchunks, 296);
  long long r = (m.N + want - 1) / want;
  r = ((r + p->tr - 1) / p->tr) * p->tr;  // whole row tiles per chunk
  p->rpc = (int)r;
  p->chunks = (int)((m.N + r - 1) / r);
  return 0;
}

static tc::Data *tc_data_of(Engine *e);

// Option "time_kernels": CUDA events around a systematic sample of the gradient-kernel launches of
// a run (every stride-th launch; when the event pool is full, every other sample is dropped and the
// stride doubles), so that bench.py can state the kernel's time inside its own timed region.
struct KernelTimer {
  static constexpr int kCap = 2048;
  std::vector<cudaEvent_t> ev;   // 2 per sample
  std::vector<long long> idx;    // launch index of every sample
  long long stride = 1, launches = 0, tiles = 0;
  bool armed = false;            // this launch is bracketed
  void begin_run() {
    idx.clear();
    stride = 1;
    launches = tiles = 0;
  }
  void before(cudaStream_t s) {
    armed = false;
    if (launches % stride == 0) {
      if ((int)idx.size() == kCap) {  // thin out: keep the samples on the doubled stride
        stride *= 2;
        size_t w = 0;
        for (size_t r = 0; r < idx.size(); ++r)
          if (idx[r] % stride == 0) {
            std::swap(ev[2 * w], ev[2 * r]);
            std::swap(ev[2 * w + 1], ev[2 * r + 1]);
            idx[w++] = idx[r];
          }
        idx.resize(w);
      }
      if (launches % stride == 0) {
        const size_t k = idx.size();
        while (ev.size() < 2 * (k + 1)) {
          cudaEvent_t x;
          if (cudaEventCreate(&x) != cudaSuccess) return;
          ev.push_back(x);
        }
        idx.push_back(launches);
        cudaEventRecord(ev[2 * k], s);
        armed = true;
      }
    }
  }
  void after(cudaStream_t s, long long n_tiles) {
    if (armed) cudaEventRecord(ev[2 * idx.size() - 1], s);
    armed = false;
    launches += 1;
    tiles += n_tiles;
  }
  // call after the stream has been synchronised
  void finish(Engine *e) {
    double tot = 0.0;
    for (size_t k = 0; k < idx.size(); ++k) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, ev[2 * k], ev[2 * k + 1]) == cudaSuccess) tot += ms * 1e-3;
    }
    e->kernel_sampled = (long long)idx.size();
    e->kernel_launches = launches;
    e->kernel_tile_launches = tiles;
    e->kernel_seconds = idx.empty() ? 0.0 : tot * ((double)launches / (double)idx.size());
  }
  ~KernelTimer() {
    for (cudaEvent_t x : ev) cudaEventDestroy(x);
  }
};
static KernelTimer *timer_of(Engine *e);

static int grad_launch(Engine *e, const double *th, int C, double *lp, double *g, const int *list,
                       const int *count, int n_all, int n_max, GradWork *gw) {
  const ModelDev &m = e->model;
  LockstepState *ls = e->ls;
  if (n_max <= 0) return TNUTS_OK;
  switch (m.family) {
    case TNUTS_GDEMO:
    case TNUTS_LINREG:
    case TNUTS_EIGHT_SCHOOLS:
    case TNUTS_STD_NORMAL:
      k_lpgrad_small<<<(n_max + 127) / 128, 128, 0, e->stream>>>(m, th, C, lp, g, list, count, n_all);
      e->launches += 1;
      break;
    case TNUTS_HIER_LINREG:
      k_hier_tile<<<tiles(n_max), kTileThreads, 0, e->stream>>>(m, *ls, th, C, list, count, n_all, lp, g);
      e->launches += 1;
      break;
    case TNUTS_LOGISTIC: {
      LogiPlan lp_;
      if (logistic_plan(m, gw->cap, &lp_)) return set_error(e, TNUTS_E_UNSUPPORTED, "logistic regression: D > 256 not supported");
      const int stride = gw->stride;
      const int D = m.D;
      int chunks = lp_.chunks;
      if (e->logistic_tc) {
        // tcgen05 path: same partial-sum layout, its own (fixed) row cut
        tc::Data *td = tc_data_of(e);
        if (!td->ready) {
          const char *msg = tc::prepare(m.X, m.y, m.N, D, e->stream, td);
          if (msg) return set_error(e, TNUTS_E_UNSUPPORTED, msg);
          e->launches += 2;
        }
        int rpc_;
        tc::cut_for(m.N, (n_max + tc::kChainsPerCta - 1) / tc::kChainsPerCta, &chunks, &rpc_, td->pl.wide ? 2 : 1);
        if ((size_t)chunks * stride * ((size_t)D + 1) > gw->part_elems)
          return set_error(e, TNUTS_E_UNSUPPORTED, "logistic_tc: partial buffer too small");
        KernelTimer *kt = e->time_kernels ? timer_of(e) : nullptr;
        if (kt) kt->before(e->stream);
        TN_CUDA(e, tc::launch(*td, m.N, D, th, C, list, count, n_all, n_max, gw->part, stride, e->stream));
        if (kt) kt->after(e->stream, (n_max + tc::kChainsPerCta - 1) / tc::kChainsPerCta);
      } else {
      KernelTimer *kt = e->time_kernels ? timer_of(e) : nullptr;
      if (kt) kt->before(e->stream);
      const dim3 grid((n_max + lp_.tch - 1) / lp_.tch, (chunks + lp_.ng - 1) / lp_.ng);
#define TN_LOGI(NWN, NWM, MT, TR, NG)                                                                 \
  do {                                                                                                \
    auto kfn = k_logistic_mma<NWN, NWM, MT, TR, NG>;                                                  \
    TN_CUDA(e, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp_.smem)); \
    kfn<<<grid, 256 * NG, lp_.smem, e->stream>>>(m.X, m.y, m.N, D, th, C, list, count, n_all,         \
                                                 lp_.rpc, gw->part, stride);                          \
  } while (0)
      if (lp_.ng == 2) {
        if (lp_.mt == 1) TN_LOGI(8, 1, 1, 32, 2);
        else if (lp_.mt == 2) TN_LOGI(8, 1, 2, 32, 2);
        else if (lp_.mt == 4) TN_LOGI(8, 1, 4, 32, 2);
        else if (lp_.mt == 8) TN_LOGI(8, 1, 8, 32, 2);
        else TN_LOGI(8, 1, 13, 32, 2);
      } else if (lp_.nwm == 2) TN_LOGI(4, 2, 16, 32, 1);
      else if (lp_.tr == 32) {
        if (lp_.mt == 16) TN_LOGI(8, 1, 16, 32, 1);
        else TN_LOGI(8, 1, 13, 32, 1);
      } else if (lp_.mt == 1) TN_LOGI(8, 1, 1, 64, 1);
      else if (lp_.mt == 2) TN_LOGI(8, 1, 2, 64, 1);
      else if (lp_.mt == 4) TN_LOGI(8, 1, 4, 64, 1);
      else if (lp_.mt == 8) TN_LOGI(8, 1, 8, 64, 1);
      else if (lp_.mt == 13) TN_LOGI(8, 1, 13, 64, 1);
      else TN_LOGI(8, 1, 16, 64, 1);
#undef TN_LOGI
      if (kt) kt->after(e->stream, (n_max + lp_.tch - 1) / lp_.tch);
      }
      if (e->world > 1 && e->cfg.row_shards > 1) {
        const dim3 rg((n_max + 127) / 128, D + 1);
        k_logistic_reduce<<<rg, 128, 0, e->stream>>>(gw->part, chunks, D, stride, count, n_all, gw->lik, gw->cap);
        int rc = comm_allreduce_sum(e, gw->lik, (size_t)(D + 1) * gw->cap);
        if (rc) return rc;
        k_logistic_finish<<<(n_max + 127) / 128, 128, 0, e->stream>>>(m, gw->lik, gw->cap, th, C, list,
                                                                     count, n_all, 1.0, lp, g);
        e->launches += 3;
      } else {
        const dim3 rg((n_max + 31) / 32, (D + 1 + 7) / 8);
        k_logistic_reduce_finish<<<rg, 256, 0, e->stream>>>(m, gw->part, chunks, stride, th, C, list, count, n_all,
                                                            lp, g);
        e->launches += 2;
      }
    } break;
    default:
      return set_error(e, TNUTS_E_UNSUPPORTED, "lock-step: unknown family");
  }
  TN_CUDA(e, cudaGetLastError());
  return TNUTS_OK;
}

static int gradwork_alloc(Engine *e, GradWork *gw, int cap, std::vector<void *> *pool) {
  gw->cap = cap;
  if (e->model.family != TNUTS_LOGISTIC) return TNUTS_OK;
  LogiPlan lp_;
  if (logistic_plan(e->model, cap, &lp_)) return set_error(e, TNUTS_E_UNSUPPORTED, "logistic regression: D > 256 not supported");
  gw->chunks = lp_.chunks;
  gw->rows_per_chunk = lp_.rpc;
  gw->stride = ((cap + lp_.tch - 1) / lp_.tch) * lp_.tch;
  int max_chunks = lp_.chunks;
  tc::Plan tp;
  if (tc::plan(e->model.N, e->model.D, &tp)) {
    if (tc::kSMs > max_chunks) max_chunks = tc::kSMs;  // the finest cut of tc::cut_for
  }
  gw->part_elems = (size_t)max_chunks * gw->stride * ((size_t)e->model.D + 1);
  const size_t D1 = (size_t)e->model.D + 1;
  void *p = nullptr;
  TN_CUDA(e, cudaMallocAsync(&p, 8 * gw->part_elems, e->stream));
  pool->push_back(p);
  gw->part = (double *)p;
  TN_CUDA(e, cudaMallocAsync(&p, 8 * D1 * cap, e->stream));
  TN_CUDA(e, cudaMemsetAsync(p, 0, 8 * D1 * cap, e->stream));
  pool->push_back(p);
  gw->lik = (double *)p;
  return TNUTS_OK;
}

struct LockstepExt {
  LockstepState ls;  // first member: Engine::ls points here
  GradWork gw;
  tc::Data tcd;      // logistic regression on tcgen05 (option "logistic_tc"), built on first use
  KernelTimer timer; // option "time_kernels"
  std::vector<void *> allocs;
};
static LockstepExt *ext_of(Engine *e) { return reinterpret_cast<LockstepExt *>(e->ls); }
static tc::Data *tc_data_of(Engine *e) { return &ext_of(e)->tcd; }
static KernelTimer *timer_of(Engine *e) { return &ext_of(e)->timer; }

// ------------------------------------------------------------------------------------------
int lockstep_create(Engine *e) {
  if (e->cfg.max_depth > kMaxLevels)
    return set_error(e, TNUTS_E_UNSUPPORTED, "lock-step: max_depth too large");
  LockstepExt *x = new LockstepExt();
  LockstepState *ls = &x->ls;
  e->ls = ls;
  const size_t C = (size_t)e->st.C, D = (size_t)e->D, DC = D * C;
  ls->C = (int)C;
  ls->D = (int)D;
  ls->levels = std::max(1, e->cfg.max_depth - 1);
  int rc;
#define A(p, n) if ((rc = ls_alloc(e, &x->allocs, &(p), (n)))) return rc
  A(ls->zt, DC); A(ls->zr, DC); A(ls->zg, DC);
  A(ls->edge[0], 3 * DC); A(ls->edge[1], 3 * DC);
  A(ls->rho, DC); A(ls->rho_s, DC);
  A(ls->thC, DC); A(ls->gC, DC); A(ls->thS, DC); A(ls->gS, DC);
  A(ls->ck_r, ls->levels * DC); A(ls->ck_rho, ls->levels * DC);
  TreeScalars &s = ls->ts;
  A(s.H0, C); A(s.sum_a, C); A(s.dHmax, C); A(s.lw, C); A(s.lw_s, C); A(s.lpC, C); A(s.HC, C);
  A(s.lpS, C); A(s.HS, C); A(s.zlp, C); A(s.zH, C); A(s.elp, 2 * C); A(s.n_a, C);
  A(s.depth, C); A(s.active, C); A(s.sub_active, C); A(s.term_s, C); A(s.numerr, C); A(s.v, C);
  A(s.n, C); A(s.phase, C); A(s.iter0, C); A(s.iter_end, C);
  A(ls->list, C); A(ls->count, 2);
  A(ls->fs_eps, C); A(ls->fs_epsp, C); A(ls->fs_H, C); A(ls->fs_dir, C); A(ls->fs_phase, C);
  A(ls->fs_iter, C); A(ls->r0, DC);
  if (e->st.minv_dense) {   // DenseEuclideanMetric: the whitened pump (dense_pump.cuh)
    A(ls->wT, DC); A(ls->wG, DC); A(ls->dn_req, C); A(ls->dn_n, C);
  }
  TN_CUDA(e, cudaMallocHost((void **)&ls->h_count, 4 * sizeof(int)));
  if ((rc = gradwork_alloc(e, &x->gw, (int)C, &x->allocs))) return rc;
  // hierarchical model: centred sufficient statistics per group, computed once on the host
  if (e->model.family == TNUTS_HIER_LINREG) {
    const int G = e->model.G;
    const long long N = e->model.N;
    std::vector<double> hx((size_t)N), hy((size_t)N);
    std::vector<int> grp((size_t)N);
    TN_CUDA(e, cudaMemcpy(hx.data(), e->model.X, 8 * (size_t)N, cudaMemcpyDeviceToHost));
    TN_CUDA(e, cudaMemcpy(hy.data(), e->model.y, 8 * (size_t)N, cudaMemcpyDeviceToHost));
    TN_CUDA(e, cudaMemcpy(grp.data(), e->model.group, 4 * (size_t)N, cudaMemcpyDeviceToHost));
    std::vector<double> n(G, 0.0), xb(G, 0.0), yb(G, 0.0), sxx(G, 0.0), sxy(G, 0.0), syy(G, 0.0);
    for (long long i = 0; i < N; ++i) {
      const int g = grp[(size_t)i];
      if (g < 0 || g >= G) return set_error(e, TNUTS_E_ARG, "hier: group index out of range");
      n[g] += 1.0;
      xb[g] += hx[(size_t)i];
      yb[g] += hy[(size_t)i];
    }
    for (int g = 0; g < G; ++g)
      if (n[g] > 0) {
        xb[g] /= n[g];
        yb[g] /= n[g];
      }
    for (long long i = 0; i < N; ++i) {
      const int g = grp[(size_t)i];
      const double dx = hx[(size_t)i] - xb[g], dy = hy[(size_t)i] - yb[g];
      sxx[g] += dx * dx;
      sxy[g] += dx * dy;
      syy[g] += dy * dy;
    }
    auto up = [&](double **dst, const std::vector<double> &v) -> int {
      int r2 = ls_alloc(e, &x->allocs, dst, v.size());
      if (r2) return r2;
      TN_CUDA(e, cudaMemcpy(*dst, v.data(), 8 * v.size(), cudaMemcpyHostToDevice));
      return TNUTS_OK;
    };
    if ((rc = up(&ls->hs_n, n)) || (rc = up(&ls->hs_xbar, xb)) || (rc = up(&ls->hs_ybar, yb)) ||
        (rc = up(&ls->hs_sxx, sxx)) || (rc = up(&ls->hs_sxy, sxy)) || (rc = up(&ls->hs_syy, syy)))
      return rc;
  }
#undef A
  return TNUTS_OK;
}

void lockstep_destroy(Engine *e) {
  if (!e->ls) return;
  LockstepExt *x = ext_of(e);
  for (void *p : x->allocs) cudaFreeAsync(p, e->stream);
  tc::release(&x->tcd, e->stream);
  if (x->ls.h_count) cudaFreeHost(x->ls.h_count);
  delete x;
  e->ls = nullptr;
}

static GradWork *gw_of(Engine *e) { return &ext_of(e)->gw; }

// ------------------------------------------------------------------------------------------
// init: initial parameters (<= 1000 tries), step-size search
// ------------------------------------------------------------------------------------------
__global__ void k_ls_init_draw(ChainState st, RunParams rp, int attempt) {
  const int C = st.C, D = st.D;
  const int npair = (D + 1) / 2;
  const long long tot = (long long)npair * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot;
       i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i / C), c = (int)(i % C);
    if (st.status[c] == 0) continue;  // already has a finite point
    double u0, u1;
    uniform2(rp.seed, (uint32_t)(rp.chain_offset + c), 0u, SLOT_INIT, (uint32_t)(attempt * npair + p), u0, u1);
    st.theta[(size_t)(2 * p) * C + c] = 4.0 * u0 - 2.0;
    if (2 * p + 1 < D) st.theta[(size_t)(2 * p + 1) * C + c] = 4.0 * u1 - 2.0;
  }
}

__global__ void k_ls_set_theta0(ChainState st, const double *theta0 /* [C][D] */) {
  const int C = st.C, D = st.D;
  const long long tot = (long long)D * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot;
       i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i / C), c = (int)(i % C);
    st.theta[i] = theta0[(size_t)c * D + d];
  }
}

// status: 1 = needs a point; after the gradient: finite -> 0
__global__ void __launch_bounds__(kTileThreads) k_ls_init_check(ChainState st, int attempt, int *n_bad) {
  __shared__ double red[64];
  const Tile t;
  const int C = st.C, D = st.D;
  const bool in = t.c < C && st.status[t.c] != 0;
  double bad[1] = {0.0};
  if (in)
    for (int d = t.dl; d < D; d += kTileD) bad[0] += isfinite(st.grad[(size_t)d * C + t.c]) ? 0.0 : 1.0;
  tile_reduce<1>(bad, red);
  if (in && t.dl == 0) {
    const bool ok = isfinite(st.lp[t.c]) && bad[0] == 0.0;
    st.init_tries[t.c] = attempt + 1;
    st.n_grad[t.c] += 1;
    if (ok) st.status[t.c] = 0;
    else atomicAdd(n_bad, 1);
  }
}

__global__ void k_ls_init_state(ChainState st, double init_eps, int algorithm) {
  const int C = st.C, D = st.D;
  const long long tot = (long long)D * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot;
       i += (long long)gridDim.x * blockDim.x) {
    st.minv[i] = 1.0;
    st.wf_mean[i] = 0.0;
    st.wf_m2[i] = 0.0;
    if (i < C) {
      st.status[i] = 1;
      st.iter[i] = 0;
      st.n_grad[i] = 0;
      st.wf_n[i] = 0;
      st.da_m[i] = 0;
      st.da_xbar[i] = 0.0;
      st.da_hbar[i] = 0.0;
      st.eps[i] = init_eps;
      st.init_tries[i] = 0;
    }
  }
}

// step-size search, lock-step: phase 0 first trial, 1 crossing, 2 bisection, 3 done
__global__ void __launch_bounds__(kTileThreads)
k_ls_fs_begin(ChainState st, LockstepState ls, RunParams rp) {
  __shared__ double red[64];
  const Tile t;
  const int C = st.C, D = st.D;
  const bool in = t.c < C;
  double acc[1] = {0.0};
  if (in) {
    const int npair = (D + 1) / 2;
    const uint32_t gchain = (uint32_t)(rp.chain_offset + t.c);
    for (int p = t.dl; p < npair; p += kTileD) {
      double z0, z1;
      normal2(rp.seed, gchain, 0u, SLOT_EPS_MOMENTUM, (uint32_t)p, z0, z1);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int d = 2 * p + q;
        if (d < D) {
          const size_t o = (size_t)d * C + t.c;
          const double mi = st.minv[o];
          const double r = (q ? z1 : z0) / sqrt(mi);
          ls.r0[o] = r;
          acc[0] += mi * r * r;
        }
      }
    }
  }
  tile_reduce<1>(acc, red);
  if (in && t.dl == 0) {
    const double lp = st.lp[t.c], K = 0.5 * acc[0];
    ls.fs_H[t.c] = (isfinite(lp) && isfinite(K)) ? (-lp + K) : CUDART_INF;
    ls.fs_eps[t.c] = 0.1;
    ls.fs_epsp[t.c] = 0.1;
    ls.fs_phase[t.c] = 0;
    ls.fs_iter[t.c] = 0;
    ls.ts.sub_active[t.c] = 1;
  }
}

// trial leapfrog, first half: z = (theta, r0, grad) -> kick + drift with the chain's trial eps
__global__ void __launch_bounds__(kTileThreads)
k_ls_fs_trial(ChainState st, LockstepState ls) {
  const Tile t;
  const int C = st.C, D = st.D;
  if (t.c >= C || ls.fs_phase[t.c] >= 3) return;
  const int ph = ls.fs_phase[t.c];
  const double eps = ph == 2 ? 0.5 * (ls.fs_eps[t.c] + ls.fs_epsp[t.c]) : ls.fs_eps[t.c];
  const double half = 0.5 * eps;
  for (int d = t.dl; d < D; d += kTileD) {
    const size_t o = (size_t)d * C + t.c;
    const double r = ls.r0[o] + half * st.grad[o];
    ls.zr[o] = r;
    ls.zt[o] = st.theta[o] + eps * (st.minv[o] * r);
  }
}

__global__ void __launch_bounds__(kTileThreads)
k_ls_fs_decide(ChainState st, LockstepState ls, int *n_left) {
  __shared__ double red[2 * 64];
  const Tile t;
  const int C = st.C, D = st.D;
  const bool in = t.c < C && ls.fs_phase[t.c] < 3;
  double a[2] = {0.0, 0.0};
  int ph = 3;
  double eps = 0, epsp = 0, tried = 0;
  if (in) {
    ph = ls.fs_phase[t.c];
    eps = ls.fs_eps[t.c];
    epsp = ls.fs_epsp[t.c];
    tried = ph == 2 ? 0.5 * (eps + epsp) : eps;
    const double half = 0.5 * tried;
    for (int d = t.dl; d < D; d += kTileD) {
      const size_t o = (size_t)d * C + t.c;
      const double g = ls.zg[o];
      const double r = ls.zr[o] + half * g;
      a[0] += st.minv[o] * r * r;
      a[1] += isfinite(g) ? 0.0 : 1.0;
    }
  }
  tile_reduce<2>(a, red);
  if (in && t.dl == 0) {
    const double lp = ls.ts.zlp[t.c], K = 0.5 * a[0];
    const double Hn = (isfinite(lp) && a[1] == 0.0 && isfinite(K)) ? (-lp + K) : CUDART_INF;
    const double dH = ls.fs_H[t.c] - Hn;
    const double lc = log(0.5);
    int iter = ls.fs_iter[t.c];
    st.n_grad[t.c] += 1;
    if (ph == 0) {
      ls.fs_dir[t.c] = dH > lc ? 1 : -1;
      ph = 1;
      iter = 0;
      epsp = ls.fs_dir[t.c] == 1 ? 2.0 * eps : 0.5 * eps;  // eps' of the first crossing iteration
    } else if (ph == 1) {
      const int dir = ls.fs_dir[t.c];
      bool brk = (dir == 1 && !(dH > lc)) || (dir == -1 && !(dH < lc));
      if (!brk) {
        eps = epsp;
        iter += 1;
        if (iter >= 100) brk = true;
        else epsp = dir == 1 ? 2.0 * eps : 0.5 * eps;
      }
      if (brk) {
        if (eps > epsp) {
          const double tmp = eps;
          eps = epsp;
          epsp = tmp;
        }
        ph = 2;
        iter = 0;
      }
    } else {
      const double ea = exp(dH);
      if (ea > 0.75) eps = tried;
      else if (ea < 0.25) epsp = tried;
      else {
        eps = tried;
        ph = 3;
      }
      iter += 1;
      if (iter >= 100) ph = 3;
    }
    ls.fs_eps[t.c] = eps;
    ls.fs_epsp[t.c] = epsp;
    ls.fs_phase[t.c] = ph;
    ls.fs_iter[t.c] = iter;
    ls.ts.sub_active[t.c] = ph < 3 ? 1 : 0;
    if (ph < 3) atomicAdd(n_left, 1);
    else st.eps[t.c] = eps;
  }
}

__global__ void k_ls_set_eps(ChainState st, double eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < st.C) st.eps[c] = eps;
}

__global__ void k_ls_finish_init(ChainState st) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= st.C) return;
  const double eps = st.eps[c];
  st.da_eps[c] = eps;
  st.da_mu[c] = log(10.0 * eps);
}

static int read_count(Engine *e, int *dev, int *host_slot) {
  TN_CUDA(e, cudaMemcpyAsync(host_slot, dev, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  TN_CUDA(e, cudaStreamSynchronize(e->stream));
  return TNUTS_OK;
}

int lockstep_init(Engine *e, const double *theta0_dev) {
  LockstepState *ls = e->ls;
  ChainState &st = e->st;
  const int C = st.C;
  int rc;
  k_ls_init_state<<<148 * 4, 256, 0, e->stream>>>(st, e->cfg.init_eps, e->cfg.algorithm);
  e->launches += 1;
  const int max_tries = theta0_dev ? 1 : 1000;
  int n_bad = C;
  for (int attempt = 0; attempt < max_tries && n_bad > 0; ++attempt) {
    if (theta0_dev) k_ls_set_theta0<<<148 * 4, 256, 0, e->stream>>>(st, theta0_dev);
    else k_ls_init_draw<<<148 * 4, 256, 0, e->stream>>>(st, e->rp, attempt);
    k_ls_compact<<<1, 1024, 0, e->stream>>>(st.status, C, ls->list, ls->count);
    if ((rc = grad_launch(e, st.theta, C, st.lp, st.grad, ls->list, ls->count, C, n_bad, gw_of(e)))) return rc;
    TN_CUDA(e, cudaMemsetAsync(ls->count + 1, 0, sizeof(int), e->stream));
    k_ls_init_check<<<tiles(C), kTileThreads, 0, e->stream>>>(st, attempt, ls->count + 1);
    e->launches += 3;
    if ((rc = read_count(e, ls->count + 1, ls->h_count))) return rc;
    n_bad = ls->h_count[0];
  }
  if (n_bad > 0) return TNUTS_OK;  // tnuts_init reports TNUTS_E_INIT from the status array
  if (e->cfg.algorithm == TNUTS_ALG_SGHMC || e->cfg.algorithm == TNUTS_ALG_SGLD) {
    // no step-size search; the velocity (st.wf_mean) was zeroed by k_ls_init_state
    k_ls_set_eps<<<(C + 255) / 256, 256, 0, e->stream>>>(st, e->cfg.init_eps);
    e->launches += 1;
  } else if (e->cfg.init_eps == 0.0) {
    // find_good_stepsize for every chain (iszero(spl.eps), HMC included: src/mcmc/hmc.jl:189-194)
    k_ls_fs_begin<<<tiles(C), kTileThreads, 0, e->stream>>>(st, *ls, e->rp);
    e->launches += 1;
    int left = C;
    for (int iter = 0; iter < 1 + 100 + 100 + 2 && left > 0; ++iter) {
      k_ls_fs_trial<<<tiles(C), kTileThreads, 0, e->stream>>>(st, *ls);
      k_ls_compact<<<1, 1024, 0, e->stream>>>(ls->ts.sub_active, C, ls->list, ls->count);
      if ((rc = grad_launch(e, ls->zt, C, ls->ts.zlp, ls->zg, ls->list, ls->count, C, left, gw_of(e)))) return rc;
      TN_CUDA(e, cudaMemsetAsync(ls->count + 1, 0, sizeof(int), e->stream));
      k_ls_fs_decide<<<tiles(C), kTileThreads, 0, e->stream>>>(st, *ls, ls->count + 1);
      e->launches += 3;
      if ((rc = read_count(e, ls->count + 1, ls->h_count))) return rc;
      left = ls->h_count[0];
    }
  }
  k_ls_finish_init<<<(C + 255) / 256, 256, 0, e->stream>>>(st);
  e->launches += 1;
  if (st.minv_dense) {   // DenseEuclideanMetric(D): identity, phi = theta
    k_dense_init<<<148 * 8, 256, 0, e->stream>>>(st);
    e->launches += 1;
  }
  TN_CUDA(e, cudaGetLastError());
  return TNUTS_OK;
}

// ------------------------------------------------------------------------------------------
// run: the host pumps "one leapfrog for every unfinished chain" until all chains have done
// their transitions.  No host<->device synchronisation on the way: the pump kernel mirrors its
// finished-chain counter into mapped pinned memory, the host polls that word and only bounds
// how far it runs ahead of the GPU (an event every 64 pumps).
// ------------------------------------------------------------------------------------------
// SGHMC / SGLD: every chain takes exactly one gradient per transition, so the "pump" is a plain
// loop: move (and record the previous kept draw), gradient at the new position, repeat.
static int lockstep_run_sg(Engine *e) {
  ChainState &st = e->st;
  const RunParams &rp = e->rp;
  const int C = st.C;
  const int T = tiles(C);
  const double a = e->cfg.init_eps, b = e->cfg.delta, c = e->cfg.lambda;
  auto slot = [&](long long tt) -> long long {
    if (tt >= 1 && rp.first_keep >= 1 && tt >= rp.first_keep && (tt - rp.first_keep) % rp.thin == 0)
      return (tt - rp.first_keep) / rp.thin;
    return -1;
  };
  int rc;
  for (long long tt = 1; tt <= rp.n_transitions; ++tt) {
    k_sg_step<<<T, kTileThreads, 0, e->stream>>>(st, rp, e->model, e->out, slot(tt - 1), 1, a, b, c);
    e->launches += 1;
    if ((rc = grad_launch(e, st.theta, C, st.lp, st.grad, nullptr, nullptr, C, C, gw_of(e)))) return rc;
  }
  const long long last = slot(rp.n_transitions);
  if (last >= 0) {
    k_sg_step<<<T, kTileThreads, 0, e->stream>>>(st, rp, e->model, e->out, last, 0, a, b, c);
    e->launches += 1;
  }
  TN_CUDA(e, cudaGetLastError());
  return TNUTS_OK;
}

// ------------------------------------------------------------------------------------------
// Progressive compaction.  Chains finish a run at very different pump counts (config 3: the median
// chain needs 27 000 leapfrogs, the slowest 1 % need 90 000 - 200 000), so most pumps serve a small
// and shrinking subset of the chains -- scattered over [rows][C] arrays in which every touched 8-byte
// element then sits in a cache line of its own (4400 lines = 560 KB per chain: 128 unfinished chains
// alone use up L2 next to the X planes, and the state machine runs at DRAM latency).  Whenever the
// unfinished chains fit half of the current width, they are gathered into arrays of that width
// ("view"), all kernels of the pump work on the view, and column k of the view is chain gid[k] of
// the engine for the two things that name a chain globally: its Philox address and its column of the
// draw arrays (DrawBuffers::gid).  Chains are scattered back before the next gather and at the end of
// the run.  Per-chain arithmetic does not depend on the column a chain lives in, and the compaction
// points are a function of the (deterministic) refresh schedule, so results are bit-identical with
// and without it (option "compact", tests/test_gpu_sampling.py).
// ------------------------------------------------------------------------------------------
struct ColArray {
  char *home, *view;
  int rows, esz;
};

template <class F>
static void for_each_chain_array(ChainState &st, LockstepState &ls, F &&f) {
  const int D = st.D;
#define COL(p, rows) f(reinterpret_cast<void **>(&(p)), (int)(rows), (int)sizeof(*(p)))
  COL(st.theta, D); COL(st.grad, D); COL(st.lp, 1); COL(st.minv, D); COL(st.eps, 1);
  COL(st.da_mu, 1); COL(st.da_xbar, 1); COL(st.da_hbar, 1); COL(st.da_eps, 1); COL(st.da_m, 1);
  COL(st.wf_n, 1); COL(st.wf_mean, D); COL(st.wf_m2, D); COL(st.iter, 1); COL(st.n_grad, 1);
  COL(st.init_tries, 1); COL(st.status, 1);
  COL(ls.zt, D); COL(ls.zr, D); COL(ls.zg, D); COL(ls.edge[0], 3 * D); COL(ls.edge[1], 3 * D);
  COL(ls.rho, D); COL(ls.rho_s, D); COL(ls.thC, D); COL(ls.gC, D); COL(ls.thS, D); COL(ls.gS, D);
  COL(ls.ck_r, ls.levels * D); COL(ls.ck_rho, ls.levels * D);
  TreeScalars &s = ls.ts;
  COL(s.H0, 1); COL(s.sum_a, 1); COL(s.dHmax, 1); COL(s.lw, 1); COL(s.lw_s, 1); COL(s.lpC, 1); COL(s.HC, 1);
  COL(s.lpS, 1); COL(s.HS, 1); COL(s.zlp, 1); COL(s.zH, 1); COL(s.elp, 2); COL(s.n_a, 1);
  COL(s.depth, 1); COL(s.active, 1); COL(s.sub_active, 1); COL(s.term_s, 1); COL(s.numerr, 1); COL(s.v, 1);
  COL(s.n, 1); COL(s.phase, 1); COL(s.iter0, 1); COL(s.iter_end, 1);
#undef COL
}

// to_view: view[r][k] = home[r][gid[k]] for k < n; else the other way round
__global__ void k_cols_move(const ColArray *tab, const int *__restrict__ gid, int n, int Chome, int Cview,
                            int to_view) {
  const ColArray a = tab[blockIdx.y];
  const long long tot = (long long)a.rows * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n;
    const int k = (int)(i - r * n);
    const size_t oh = (size_t)r * Chome + gid[k], ov = (size_t)r * Cview + k;
    if (a.esz == 8) {
      double *h = reinterpret_cast<double *>(a.home), *v = reinterpret_cast<double *>(a.view);
      if (to_view) v[ov] = h[oh];
      else h[oh] = v[ov];
    } else {
      int *h = reinterpret_cast<int *>(a.home), *v = reinterpret_cast<int *>(a.view);
      if (to_view) v[ov] = h[oh];
      else h[oh] = v[ov];
    }
  }
}

// columns [n, Cview) of a view hold no chain: finished, not ok; the work list is the identity
__global__ void k_view_finish(int *phase, int *status, int *list, int n, int Cview) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Cview) return;
  if (k >= n) {
    phase[k] = 2;
    status[k] = 1;
  }
  list[k] = k < n ? k : 0;
}

struct Compactor {
  Engine *e = nullptr;
  ChainState st;          // current view (== the engine's arrays while cap == C)
  LockstepState ls;
  DrawBuffers out;
  int cap = 0, n_cols = 0;     // width of the view, columns that hold a chain
  bool compacted = false;
  char *pool = nullptr;
  size_t pool_bytes = 0;
  ColArray *tab_dev = nullptr;
  int *gid[2] = {nullptr, nullptr};
  int cur = 0, n_tab = 0, max_rows = 1;

  void begin(Engine *e_) {
    e = e_;
    st = e->st;
    ls = *e->ls;
    out = e->out;
    cap = e->st.C;
    n_cols = cap;
  }
  static int want_cap(int n) { return std::max(128, ((n + 127) / 128) * 128); }
  // the largest number of unfinished chains at which a view of width cap_ is worth narrowing
  static int narrow_below(int cap_) { return cap_ >= 256 ? 128 * (cap_ / 256) : 0; }

  int scatter_back() {
    if (!compacted) return TNUTS_OK;
    const dim3 g((unsigned)std::min<long long>(4096, ((long long)max_rows * n_cols + 255) / 256), n_tab);
    if (n_cols > 0) k_cols_move<<<g, 256, 0, e->stream>>>(tab_dev, gid[cur], n_cols, e->st.C, cap, 0);
    e->launches += 1;
    TN_CUDA(e, cudaGetLastError());
    return TNUTS_OK;
  }

  // gather the unfinished chains (n_max is the schedule's bound on their number) into a view of
  // width want_cap(n_max); on return the view's work list is the identity and ls->count[0] = n
  int compact(int n_max) {
    LockstepState *home = e->ls;
    const int C = e->st.C, D = e->st.D;
    const int ncap = want_cap(n_max);
    int rc;
    if ((rc = scatter_back())) return rc;
    if (!pool) {
      // sized once, for the first (widest) view
      size_t bytes = 0;
      n_tab = 0;
      ChainState s0 = e->st;
      LockstepState l0 = *home;
      for_each_chain_array(s0, l0, [&](void **p, int rows, int esz) {
        if (!*p) return;
        bytes += (((size_t)rows * ncap * esz + 255) / 256) * 256;
        max_rows = std::max(max_rows, rows);
        ++n_tab;
      });
      pool_bytes = bytes + (((size_t)ncap * 4 + 255) / 256) * 256;   // + the view's work list
      TN_CUDA(e, cudaMallocAsync((void **)&pool, pool_bytes, e->stream));
      TN_CUDA(e, cudaMallocAsync((void **)&tab_dev, sizeof(ColArray) * n_tab, e->stream));
      TN_CUDA(e, cudaMallocAsync((void **)&gid[0], sizeof(int) * C, e->stream));
      TN_CUDA(e, cudaMallocAsync((void **)&gid[1], sizeof(int) * C, e->stream));
    }
    // the unfinished chains of the engine, in chain order -> the new gid
    const int nxt = compacted ? 1 - cur : 0;
    k_ls_running_flags<<<(C + 255) / 256, 256, 0, e->stream>>>(*home, C, home->ts.sub_active);
    k_ls_compact<<<1, 1024, 0, e->stream>>>(home->ts.sub_active, C, gid[nxt], home->count);
    e->launches += 2;
    TN_CUDA(e, cudaMemcpyAsync(home->h_count + 2, home->count, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    TN_CUDA(e, cudaStreamSynchronize(e->stream));
    const int n = home->h_count[2];
    if (n > ncap) return set_error(e, TNUTS_E_CUDA, "compaction: more unfinished chains than the schedule's bound");
    // lay the view out in the pool and build the move table
    std::vector<ColArray> tab;
    tab.reserve(n_tab);
    st = e->st;
    ls = *home;
    size_t off = 0;
    for_each_chain_array(st, ls, [&](void **p, int rows, int esz) {
      if (!*p) return;
      ColArray a{reinterpret_cast<char *>(*p), pool + off, rows, esz};
      off += (((size_t)rows * ncap * esz + 255) / 256) * 256;
      *p = a.view;
      tab.push_back(a);
    });
    ls.list = reinterpret_cast<int *>(pool + off);
    st.C = ncap;
    ls.C = ncap;
    (void)D;
    TN_CUDA(e, cudaMemcpyAsync(tab_dev, tab.data(), sizeof(ColArray) * tab.size(), cudaMemcpyHostToDevice, e->stream));
    TN_CUDA(e, cudaStreamSynchronize(e->stream));   // `tab` is a stack object
    if (n > 0) {
      const dim3 g((unsigned)std::min<long long>(4096, ((long long)max_rows * n + 255) / 256), n_tab);
      k_cols_move<<<g, 256, 0, e->stream>>>(tab_dev, gid[nxt], n, C, ncap, 1);
    }
    k_view_finish<<<(ncap + 255) / 256, 256, 0, e->stream>>>(ls.ts.phase, st.status, ls.list, n, ncap);
    e->launches += 2;
    TN_CUDA(e, cudaGetLastError());
    out = e->out;
    out.gid = gid[nxt];
    out.Cout = C;
    cur = nxt;
    cap = ncap;
    n_cols = n;
    compacted = true;
    return TNUTS_OK;
  }

  // end of the run (also on the error paths): chains back home, pool released
  int finish() {
    int rc = scatter_back();
    compacted = false;
    if (pool) cudaFreeAsync(pool, e->stream);
    if (tab_dev) cudaFreeAsync(tab_dev, e->stream);
    for (int i = 0; i < 2; ++i)
      if (gid[i]) cudaFreeAsync(gid[i], e->stream);
    pool = nullptr;
    tab_dev = nullptr;
    gid[0] = gid[1] = nullptr;
    return rc;
  }
};

int lockstep_run(Engine *e) {
  if (e->cfg.algorithm == TNUTS_ALG_SGHMC || e->cfg.algorithm == TNUTS_ALG_SGLD) return lockstep_run_sg(e);
  LockstepState *ls = e->ls;
  const RunParams &rp = e->rp;
  const int C = e->st.C;
  int rc;
  volatile int *h_done = ls->h_count;
  TN_CUDA(e, cudaStreamSynchronize(e->stream));  // no pump of an earlier run may still write the mirror
  *h_done = 0;
  TN_CUDA(e, cudaMemsetAsync(ls->count + 1, 0, sizeof(int), e->stream));
  k_ls_begin_run<<<(C + 255) / 256, 256, 0, e->stream>>>(e->st, *ls, rp.n_transitions, ls->count + 1);
  e->launches += 1;
  cudaEvent_t ev[2];
  TN_CUDA(e, cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
  TN_CUDA(e, cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
  cudaEvent_t ev_cnt;
  TN_CUDA(e, cudaEventCreateWithFlags(&ev_cnt, cudaEventDisableTiming));
  const long long max_pumps = rp.n_transitions * ((1LL << rp.max_depth) + 2 + (rp.algorithm != TNUTS_ALG_NUTS ? 100000 : 0)) + 64;
  int listed_done = -1, n_max = C;
  rc = TNUTS_OK;
  // Rows sharded over ranks: every pump contains an NCCL all-reduce, so all ranks must issue
  // exactly the same launch sequence.  The chains are bit-identical on every rank, hence so is
  // the finished-chain counter after a given pump: read it with a stream synchronisation at
  // fixed pump numbers instead of polling the (timing-dependent) mirror.
  // The tcgen05 logistic kernel cuts the rows by the number of listed chains, so that number has
  // to be a deterministic function of the pump index as well.
  const bool lockstep_ranks = (e->world > 1 && e->cfg.row_shards > 1) ||
                              (e->logistic_tc && e->model.family == TNUTS_LOGISTIC);
  int done = 0;
  // option "persist" (default on): tcgen05 logistic likelihood on one GPU (the per-leapfrog NCCL
  // all-reduce of row shards cannot run inside the persistent kernel)
  // ... and D <= 128: the CTA-pair kernel of wider models (k_logistic_tc_wide) has no persistent form
  // (dense metric: the gathered views and the persistent kernel do not carry the [D*D][C] arrays)
  const bool dense = e->st.minv_dense != nullptr;
  const bool persist_ok = e->persist && e->logistic_tc && e->model.family == TNUTS_LOGISTIC && !dense &&
                          e->model.D <= 128 && !(e->world > 1 && e->cfg.row_shards > 1);
  e->persist_seconds = 0.0;
  e->persist_pumps = 0;
  if (e->time_kernels) timer_of(e)->begin_run();
  static const bool trace = getenv("TNUTS_PUMP_TRACE") != nullptr;
  long long pumps_done = 0;
  Compactor cp;   // the state the pump works on: the engine's arrays, later narrower copies of them
  cp.begin(e);
  for (long long pump = 0; pump < max_pumps; ++pump) {
    pumps_done = pump;
    if (trace && pump % 4096 == 0)
      fprintf(stderr, "[tnuts pump] %lld done=%d listed=%d width=%d\n", pump, (int)*h_done, n_max, cp.cap);
    if (!lockstep_ranks) {
      done = *h_done;
    } else if (pump % 16 == 0 && pump > 0) {
      // snapshot of the counter as of this pump, consumed 8 pumps later: deterministic, and the
      // stream never drains
      if (cudaMemcpyAsync(ls->h_count + 1, ls->count + 1, sizeof(int), cudaMemcpyDeviceToHost, e->stream) !=
              cudaSuccess ||
          cudaEventRecord(ev_cnt, e->stream) != cudaSuccess) {
        rc = set_error(e, TNUTS_E_CUDA, "pump: counter snapshot failed");
        break;
      }
    } else if (pump % 16 == 8 && pump > 16) {
      if (cudaEventSynchronize(ev_cnt) != cudaSuccess) {
        rc = set_error(e, TNUTS_E_CUDA, "pump: counter snapshot failed");
        break;
      }
      done = ls->h_count[1];
    }
    if (done >= C) break;
    if (done != listed_done && (pump % 8 == 0)) {  // refresh the gradient work list
      listed_done = done;
      n_max = C - done;
      if (e->compact && !dense && n_max <= Compactor::narrow_below(cp.cap)) {
        if ((rc = cp.compact(n_max))) break;       // leaves the identity list of the narrower view
        if (trace) fprintf(stderr, "[tnuts pump] %lld: %d unfinished chains gathered into a view of width %d\n", pump, cp.n_cols, cp.cap);
      } else {
        k_ls_running_flags<<<(cp.cap + 255) / 256, 256, 0, e->stream>>>(cp.ls, cp.cap, cp.ls.ts.sub_active);
        k_ls_compact<<<1, 1024, 0, e->stream>>>(cp.ls.ts.sub_active, cp.cap, cp.ls.list, ls->count);
        e->launches += 2;
      }
      if (persist_ok && n_max <= tc::kPersistMaxTiles * tc::kChainsPerCta) {
        // The listed chains fit one wave of (chain tile, row chunk) CTAs: the rest of the run is ONE
        // persistent cooperative kernel (tc_persist.cuh), same arithmetic and same refresh schedule;
        // it only comes back to have its chains gathered into a narrower view.
        tc::Data *td = tc_data_of(e);
        if (!td->ready) {
          const char *msg = tc::prepare(e->model.X, e->model.y, e->model.N, e->model.D, e->stream, td);
          if (msg) {
            rc = set_error(e, TNUTS_E_UNSUPPORTED, msg);
            break;
          }
          e->launches += 2;
        }
        GradWork *gw = gw_of(e);
        cudaEvent_t pe0 = nullptr, pe1 = nullptr;
        if (e->time_kernels) {
          cudaEventCreate(&pe0);
          cudaEventCreate(&pe1);
        }
        long long p0 = pump;
        for (;;) {
          long long ran = 0;
          int nd = 0, reason = 0, sched_done = 0;
          const PersistView pv{&cp.st, &cp.ls, &cp.out, C, e->compact ? Compactor::narrow_below(cp.cap) : 0};
          if (e->time_kernels) cudaEventRecord(pe0, e->stream);
          rc = tc_persist_run(e, *td, pv, gw->part, gw->stride, ls->count + 1, ls->h_count, p0, max_pumps, listed_done,
                              n_max, &ran, &nd, &reason, &sched_done);
          if (e->time_kernels) {
            float ms = 0.f;
            if (!rc && cudaEventRecord(pe1, e->stream) == cudaSuccess && cudaEventSynchronize(pe1) == cudaSuccess &&
                cudaEventElapsedTime(&ms, pe0, pe1) == cudaSuccess)
              e->persist_seconds += ms * 1e-3;
          }
          e->persist_pumps += ran;
          p0 += ran;
          pumps_done = p0;
          if (trace)
            fprintf(stderr, "[tnuts pump] persistent kernel, width %d: %lld pumps up to pump %lld, done=%d\n", cp.cap, ran,
                    p0, nd);
          if (rc || reason != 3) break;
          // at a list refresh with few enough chains left for a narrower view
          listed_done = sched_done;
          n_max = C - sched_done;
          if ((rc = cp.compact(n_max))) break;
        }
        if (e->time_kernels) {
          cudaEventDestroy(pe0);
          cudaEventDestroy(pe1);
        }
        break;   // the completion check below reports chains the pump budget left unfinished
      }
    }
    if (dense) {
      // whitened pump: the family's kernel sees theta = U'phi (wT, written by k_dense_post of the previous
      // pump) and its gradient (wG) becomes grad_phi = U grad_theta (zg)
      if (pump > 0) {
        if ((rc = grad_launch(e, cp.ls.wT, cp.cap, cp.ls.ts.zlp, cp.ls.wG, cp.ls.list, ls->count, cp.cap, n_max,
                              gw_of(e))))
          break;
        k_dense_grad<<<tiles(cp.cap), kTileThreads, 0, e->stream>>>(cp.st, cp.ls);
      }
      k_ls_pump_dense<<<tiles(cp.cap), kTileThreads, 0, e->stream>>>(cp.st, cp.ls, rp, e->model, cp.out, ls->count + 1,
                                                                     ls->h_count);
      k_dense_post<<<tiles(cp.cap), kTileThreads, 0, e->stream>>>(cp.st, cp.ls);
      e->launches += 3;
    } else {
    if (pump > 0) {  // on the first pump no leaf is pending yet
      if ((rc = grad_launch(e, cp.ls.zt, cp.cap, cp.ls.ts.zlp, cp.ls.zg, cp.ls.list, ls->count, cp.cap, n_max,
                            gw_of(e))))
        break;
    }
    k_ls_pump<<<tiles(cp.cap), kTileThreads, 0, e->stream>>>(cp.st, cp.ls, rp, e->model, cp.out, ls->count + 1,
                                                             ls->h_count);
    e->launches += 1;
    }
    if (!lockstep_ranks && pump % 64 == 63) {
      const int slot = (int)((pump / 64) & 1);
      if (pump >= 127) cudaEventSynchronize(ev[slot]);  // the event recorded 128 pumps ago
      cudaEventRecord(ev[slot], e->stream);
    }
  }
  cudaEventDestroy(ev[0]);
  cudaEventDestroy(ev[1]);
  cudaEventDestroy(ev_cnt);
  if (trace) fprintf(stderr, "[tnuts pump] finished after %lld pumps\n", pumps_done);
  {
    const int rc2 = cp.finish();   // chains back into the engine's arrays
    if (!rc) rc = rc2;
  }
  if (rc) return rc;
  TN_CUDA(e, cudaGetLastError());
  // The loop is bounded (max_pumps): make sure every chain really finished, otherwise the rows of
  // the kept-draw buffer that were never written would be reported as draws.
  TN_CUDA(e, cudaMemcpyAsync(ls->h_count + 1, ls->count + 1, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  TN_CUDA(e, cudaStreamSynchronize(e->stream));
  if (e->time_kernels) timer_of(e)->finish(e);
  if (ls->h_count[1] < C)
    return set_error(e, TNUTS_E_UNSUPPORTED,
                     "tnuts_run: " + std::to_string(C - ls->h_count[1]) + " chains did not finish their transitions within the "
                     "pump budget (HMCDA with a collapsing step size: lambda / eps leapfrogs per transition)");
  return TNUTS_OK;
}

// times `reps` gradient evaluations of ALL chains at their current positions (CUDA events on the
// engine's stream); used by bench.py for the roofline of the gradient kernel alone
int lockstep_bench_gradient(Engine *e, int reps, double *seconds) {
  LockstepState *ls = e->ls;
  ChainState &st = e->st;
  const int C = st.C;
  int rc;
  if ((rc = grad_launch(e, st.theta, C, ls->ts.zlp, ls->zg, nullptr, nullptr, C, C, gw_of(e)))) return rc;
  TN_CUDA(e, cudaEventRecord(e->ev0, e->stream));
  for (int i = 0; i < reps; ++i)
    if ((rc = grad_launch(e, st.theta, C, ls->ts.zlp, ls->zg, nullptr, nullptr, C, C, gw_of(e)))) return rc;
  TN_CUDA(e, cudaEventRecord(e->ev1, e->stream));
  TN_CUDA(e, cudaStreamSynchronize(e->stream));
  float ms = 0;
  TN_CUDA(e, cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  *seconds = ms * 1e-3;
  return TNUTS_OK;
}

// ------------------------------------------------------------------------------------------
// point-wise entry points: n arbitrary points as n "chains"
// ------------------------------------------------------------------------------------------
__global__ void k_transpose_in(const double *rm /* [n][D] */, double *cm /* [D][n] */, long long n, int D) {
  const long long tot = n * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot;
       i += (long long)gridDim.x * blockDim.x) {
    const long long c = i % n;
    const int d = (int)(i / n);
    cm[i] = rm[c * D + d];
  }
}
__global__ void k_transpose_out(const double *cm /* [D][n] */, double *rm /* [n][D] */, long long n, int D) {
  const long long tot = n * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < tot;
       i += (long long)gridDim.x * blockDim.x) {
    const long long c = i % n;
    const int d = (int)(i / n);
    rm[c * D + d] = cm[i];
  }
}

struct Tmp {
  std::vector<void *> p;
  ~Tmp() {
    for (void *q : p) cudaFree(q);
  }
  double *get(size_t n) {
    void *q = nullptr;
    if (cudaMalloc(&q, 8 * (n ? n : 1)) != cudaSuccess) return nullptr;
    p.push_back(q);
    return (double *)q;
  }
};

int lockstep_point_lpgrad(Engine *e, const double *theta_dev, long long n, double *lp_dev,
                          double *grad_dev) {
  const int D = e->D;
  Tmp tmp;
  double *th = tmp.get((size_t)n * D), *g = tmp.get((size_t)n * D);
  if (!th || !g) return set_error(e, TNUTS_E_CUDA, "point lpgrad: cudaMalloc failed");
  GradWork gw;
  int rc = gradwork_alloc(e, &gw, (int)n, &tmp.p);
  if (rc) return rc;
  k_transpose_in<<<148 * 2, 256, 0, e->stream>>>(theta_dev, th, n, D);
  if ((rc = grad_launch(e, th, (int)n, lp_dev, g, nullptr, nullptr, (int)n, (int)n, &gw))) return rc;
  if (grad_dev) k_transpose_out<<<148 * 2, 256, 0, e->stream>>>(g, grad_dev, n, D);
  e->launches += 2;
  TN_CUDA(e, cudaStreamSynchronize(e->stream));
  return TNUTS_OK;
}

__global__ void __launch_bounds__(kTileThreads)
k_ls_point_constrain(ModelDev m, const double *th /* [D][n] */, const double *lp, int n,
                     double *params /* [n][P] */, double *logp /* [n][3] */) {
  __shared__ double red[2 * 64];
  const Tile t;
  const bool in = t.c < n;
  double prior, lik;
  ls_user_logp(m, th, n, t, in, in ? lp[t.c] : 0.0, red, prior, lik);
  if (!in) return;
  for (int d = t.dl; d < m.D; d += kTileD) {
    const double x = th[(size_t)d * n + t.c];
    params[(size_t)t.c * m.D + d] = ls_is_positive(m, d) ? exp(x) : x;
  }
  if (t.dl == 0) {
    logp[(size_t)t.c * 3 + 0] = prior;
    logp[(size_t)t.c * 3 + 1] = lik;
    logp[(size_t)t.c * 3 + 2] = prior + lik;
  }
}

int lockstep_point_constrain(Engine *e, const double *theta_dev, long long n, double *params_dev,
                             double *logp_dev) {
  const int D = e->D;
  Tmp tmp;
  double *th = tmp.get((size_t)n * D), *g = tmp.get((size_t)n * D), *lp = tmp.get((size_t)n);
  if (!th || !g || !lp) return set_error(e, TNUTS_E_CUDA, "point constrain: cudaMalloc failed");
  GradWork gw;
  int rc = gradwork_alloc(e, &gw, (int)n, &tmp.p);
  if (rc) return rc;
  k_transpose_in<<<148 * 2, 256, 0, e->stream>>>(theta_dev, th, n, D);
  if ((rc = grad_launch(e, th, (int)n, lp, g, nullptr, nullptr, (int)n, (int)n, &gw))) return rc;
  k_ls_point_constrain<<<tiles((int)n), kTileThreads, 0, e->stream>>>(e->model, th, lp, (int)n, params_dev, logp_dev);
  e->launches += 2;
  TN_CUDA(e, cudaStreamSynchronize(e->stream));
  return TNUTS_OK;
}

// fixed-momentum trajectories: kick/drift, gradient, kick + energy, one launch each per step
__global__ void __launch_bounds__(kTileThreads)
k_ls_traj_step(const double *minv, int n, int D, double eps, int phase, double *th, double *r,
               const double *g, const double *lp, int l, int L, double *tt, double *tr, double *tH) {
  __shared__ double red[2 * 64];
  const Tile t;
  const bool in = t.c < n;
  const double half = 0.5 * eps;
  double a[2] = {0.0, 0.0};
  if (in) {
    for (int d = t.dl; d < D; d += kTileD) {
      const size_t o = (size_t)d * n + t.c;
      double rr = r[o];
      if (phase == 1) {  // second half kick of step l (l >= 1)
        rr += half * g[o];
        r[o] = rr;
      }
      a[0] += minv[d] * rr * rr;
      a[1] += isfinite(g[o]) ? 0.0 : 1.0;
      // record point l
      tt[((size_t)t.c * (L + 1) + l) * D + d] = th[o];
      tr[((size_t)t.c * (L + 1) + l) * D + d] = rr;
    }
  }
  tile_reduce<2>(a, red);
  if (in && t.dl == 0) {
    const double K = 0.5 * a[0];
    tH[(size_t)t.c * (L + 1) + l] = (isfinite(lp[t.c]) && a[1] == 0.0 && isfinite(K)) ? (-lp[t.c] + K) : CUDART_INF;
  }
  if (in && l < L) {  // first half kick + drift of step l+1
    for (int d = t.dl; d < D; d += kTileD) {
      const size_t o = (size_t)d * n + t.c;
      const double rr = r[o] + half * g[o];
      r[o] = rr;
      th[o] = th[o] + eps * (minv[d] * rr);
    }
  }
}

int lockstep_point_leapfrog(Engine *e, const double *theta_dev, const double *r_dev,
                            const double *minv_dev, double eps, int L, long long n, double *tt_dev,
                            double *tr_dev, double *tH_dev) {
  const int D = e->D;
  Tmp tmp;
  double *th = tmp.get((size_t)n * D), *r = tmp.get((size_t)n * D), *g = tmp.get((size_t)n * D),
         *lp = tmp.get((size_t)n);
  if (!th || !r || !g || !lp) return set_error(e, TNUTS_E_CUDA, "point leapfrog: cudaMalloc failed");
  GradWork gw;
  int rc = gradwork_alloc(e, &gw, (int)n, &tmp.p);
  if (rc) return rc;
  k_transpose_in<<<148 * 2, 256, 0, e->stream>>>(theta_dev, th, n, D);
  k_transpose_in<<<148 * 2, 256, 0, e->stream>>>(r_dev, r, n, D);
  e->launches += 2;
  for (int l = 0; l <= L; ++l) {
    if ((rc = grad_launch(e, th, (int)n, lp, g, nullptr, nullptr, (int)n, (int)n, &gw))) return rc;
    k_ls_traj_step<<<tiles((int)n), kTileThreads, 0, e->stream>>>(minv_dev, (int)n, D, eps, l == 0 ? 0 : 1, th, r, g,
                                                               lp, l, L, tt_dev, tr_dev, tH_dev);
    e->launches += 1;
  }
  TN_CUDA(e, cudaStreamSynchronize(e->stream));
  return TNUTS_OK;
}

}  // namespace tnuts
