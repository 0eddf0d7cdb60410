// tc_persist.cuh -- the tail of a lock-step run as ONE persistent cooperative kernel.
//
// The host pump loop (lockstep.cu) issues three launches per leapfrog of the unfinished chains:
// gradient (k_logistic_tc), reduce + prior (k_logistic_reduce_finish), per-chain NUTS state machine
// (k_ls_pump).  A run of BASELINE.json configs[2] needs ~2 x 10^5 such pumps, and all but the first
// ~25 000 serve fewer than 512 chains: they are bound by launch and kernel-boundary latency
// (85 us per pump for 25 us of tensor work), not by throughput.  Once the work list fits one wave of
// CTAs (chain tiles x row chunks <= SMs, i.e. <= 512 listed chains) this kernel takes over and runs
// every remaining pump without returning to the host:
//
//   per pump   G  gradient of this CTA's (128-chain tile, row chunk): TcCore -- the same TMA /
//                 tcgen05 / epilogue schedule as k_logistic_tc, TMEM and barriers set up once
//              -- grid barrier --
//              R  fixed-order sum of the row-chunk partials + prior, spread over all CTAs
//              -- grid barrier --
//              P  NUTS state machine (ls_pump_body), one warp per listed chain, the chains dealt
//                 round-robin over the CTAs (WarpTeam: no CTA barrier, shuffle reductions)
//              -- grid barrier --
//
// It is bit-identical to the host-pumped path by construction: the same row cut per number of chain
// tiles, the same summation order of the partials, the same per-chain arithmetic, and the same
// schedule of work-list refreshes (the finished-chain counter is sampled at pump % 16 == 0, consumed
// at pump % 16 == 8, exactly as the host loop does with its stream-ordered snapshots) -- so chains do
// not depend on whether, or when, the persistent kernel takes over (tests/test_gpu_tc.py).
//
// Cross-SM visibility: every phase boundary is a grid barrier whose fences release this CTA's writes
// and invalidate its SM's L1 (the acquire side), so what another SM wrote in an earlier phase is read
// from L2; the few loads that bypass L1 explicitly (__ldcg) are the ones polled inside a phase.
#pragma once
#include "lockstep_kernels.cuh"
#include "logistic_tc.cuh"
#include "logistic_finish.cuh"

namespace tnuts {
namespace tc {

struct PersistCut {                          // row cut per number of chain tiles (index 1..4)
  int chunks[kPersistMaxTiles + 1];
  int rpc[kPersistMaxTiles + 1];
};
struct PersistArgs {
  long long pump0;             // global pump index of the first pump of this launch
  long long max_pumps;         // budget: the launch ends before this global pump index
  int listed_done;             // finished-chain count the current work list was built for
  int n_max;                   // C - listed_done: bound of the listed chains
  int bar_mode;                // grid barrier variant (experiments)
  unsigned long long *gbar;    // grid barrier counter, zero at launch
  int *n_done;                 // finished chains (device)
  int *h_done;                 // mapped pinned mirror of n_done for the host's progress trace
  long long *cta_times;        // [grid][3] ns every CTA spent inside G, R, P (barrier waits excluded)
  long long *out_state;        // [4] pumps executed, gradient passes of CTA 0, final n_done, reason;
                               // [4..10] ns CTA 0 spent in G, barrier, R, barrier, P, barrier, refresh
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the (co-resident) grid; `target` is this thread's running arrival total.
// mode (experiments, TNUTS_BAR_MODE): 0 sc fences + acquire polls, 1 acq_rel fences + relaxed polls,
// 2 release arrival + acquire polls, 3 sc fences + volatile polls (cooperative groups' recipe)
__device__ __forceinline__ void grid_barrier(unsigned long long *bar, unsigned long long &target, unsigned nblk,
                                             int mode) {
  target += nblk;
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    if (mode == 1) {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(bar) : "memory");
      unsigned long long v;
      do {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
        if (clock64() - t0 > 8000000000LL) __trap();
      } while (v < target);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    } else if (mode == 2) {
      asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(bar) : "memory");
      while (ld_acquire_u64(bar) < target) {
        if (clock64() - t0 > 8000000000LL) __trap();
      }
    } else if (mode == 3) {
      __threadfence();
      atomicAdd(bar, 1ull);
      while (*(volatile unsigned long long *)bar < target) {
        if (clock64() - t0 > 8000000000LL) __trap();
      }
      __threadfence();
    } else {
      __threadfence();
      atomicAdd(bar, 1ull);
      while (ld_acquire_u64(bar) < target) {
        if (clock64() - t0 > 8000000000LL) __trap();   // a protocol bug traps instead of hanging the GPU
      }
      __threadfence();
    }
  }
  __syncthreads();
}

// ordered compaction of the unfinished chains into the work list, by one CTA (any block size that is
// a multiple of 32 up to 1024); same result as k_ls_running_flags + k_ls_compact
__device__ __forceinline__ void persist_compact(const int *phase, int C, int *list, int *count, int *wsum /*[33]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) wsum[32] = 0;
  __syncthreads();
  for (int c0 = 0; c0 < C; c0 += (int)blockDim.x) {
    const int c = c0 + (int)threadIdx.x;
    const int f = (c < C && __ldcg(phase + c) != 2) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    const int pre = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int off = wsum[32];
    for (int w = 0; w < warp; ++w) off += wsum[w];
    if (f) list[off + pre] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < nw; ++w) tot += wsum[w];
      wsum[32] += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = wsum[32];
}

template <int KP>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_persist(const __grid_constant__ CUtensorMap mapX, const float *__restrict__ yf, long long nrows,
             double *part /* [chunks][D+1][cap] */, int cap, ChainState st, LockstepState ls, RunParams rp,
             ModelDev m, DrawBuffers out, PersistCut cut, PersistArgs pa) {
  using Core = TcCore<KP>;
  using Team = WarpTeam;            // state machine: one epilogue warp per listed chain
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[Core::NBARS];
  __shared__ uint32_t tmem_slot;
  __shared__ double lp_part[kEpiWarps / 4][kChainsPerCta];
  __shared__ int s_maxlev[kEpiWarps];
  __shared__ int wsum[33];

  Core core;
  core.setup(smem_raw, bars, &tmem_slot, &mapX, st.D);
  const int C = st.C, D = st.D;
  const unsigned nblk = gridDim.x;
  const int b = (int)blockIdx.x;
  const int etid = (int)threadIdx.x - 64;             // index among the 512 epilogue threads
  unsigned long long bar_target = 0;
  uint32_t tt = 0, pass = 0;                          // running tile / pass counters of this CTA
  // the host consumed its pending counter snapshot at pump0 (that is when it refreshed the list and
  // handed over), so the kernel starts with snapshot == done == listed_done
  int listed_done = pa.listed_done, n_max = pa.n_max, done = pa.listed_done, snap = pa.listed_done;
  long long p = pa.pump0;
  int reason = 0;
  // phase timing (first epilogue thread of every CTA): where a pump's time goes, for the host's trace
  long long tph[7] = {0, 0, 0, 0, 0, 0, 0};
  auto now = [] {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return (long long)t;
  };
  long long tmark = now();
  auto lap = [&](int i) {
    const long long t = now();
    tph[i] += t - tmark;
    tmark = t;
  };
  const double sd = m.hyper[0], is2 = 1.0 / (sd * sd);

  for (; p < pa.max_pumps; ++p) {
    // ---- the host loop's schedule of counter snapshots and work-list refreshes ----
    if (p % 16 == 0 && p > 0) snap = __ldcg(pa.n_done);          // the count after pump p - 1
    else if (p % 16 == 8 && p > 16) done = snap;
    if (done >= C) {
      reason = 1;
      break;
    }
    if (done != listed_done && (p % 8 == 0)) {
      if (b == 0) persist_compact(ls.ts.phase, C, ls.list, ls.count, wsum);
      listed_done = done;
      n_max = C - done;
      grid_barrier(pa.gbar, bar_target, nblk, pa.bar_mode);
      lap(6);
    }
    const int n = __ldcg(ls.count);                               // listed chains
    const int n_tiles = (n_max + kChainsPerCta - 1) / kChainsPerCta;
    const int chunks = cut.chunks[n_tiles], rpc = cut.rpc[n_tiles];
    const bool grad_pass = p > 0;                                 // on the first pump of a run no leaf is pending

    // ---- G: this CTA's (chain tile, row chunk) ----
    if (grad_pass && b < n_tiles * chunks) {
      const int tile = b % n_tiles, chunk = b / n_tiles;
      const int k0 = tile * kChainsPerCta;
      const long long rbeg = (long long)chunk * rpc;
      if (k0 < n && rbeg < nrows) {
        const long long rend = (rbeg + rpc < nrows) ? rbeg + rpc : nrows;
        const int ntile = (int)((rend - rbeg + Core::T - 1) / Core::T);
        if (core.warp == 0) {
          core.produce(&mapX, rbeg, ntile, tt);
        } else if (core.warp == 1) {
          core.mma(ntile, tt, pass);
        } else {
          const int k = k0 + core.kl();
          const bool live = k < n;
          const int c = live ? __ldcg(ls.list + k) : 0;
          core.template epi_theta<true>(ls.zt, C, c, live);
          const double lp = core.epi_tiles(yf, nrows, rbeg, ntile, tt, nullptr, false);
          core.epi_store(part + (size_t)chunk * (D + 1) * cap, cap, k, live, lp, lp_part, pass);
        }
        tt += (uint32_t)ntile;
        pass += 1;
      }
    }
    lap(0);
    grid_barrier(pa.gbar, bar_target, nblk, pa.bar_mode);
    lap(1);

    // ---- R: sum of the row-chunk partials in chunk order + prior -> zg, zlp (k_logistic_reduce_finish) ----
    if (grad_pass && etid >= 0) {
      const size_t cs = (size_t)(D + 1) * cap;
      const long long total = (long long)(D + 1) * n;
      // 32 consecutive outputs per warp, warps dealt round-robin over the CTAs: every SM pulls its
      // share of the partials (a few CTAs alone are bound by their outstanding-load limit)
      const long long wstep = (long long)nblk * kEpiWarps * 32;
      for (long long o = ((long long)(etid >> 5) * nblk + b) * 32 + (etid & 31); o < total; o += wstep) {
        const int d = (int)(o / n), k = (int)(o - (long long)d * n);
        const int c = __ldcg(ls.list + k);
        const double *q = part + (size_t)d * cap + k;
        double s = 0.0;
        int ch = 0;
        for (; ch + 16 <= chunks; ch += 16) {   // sixteen loads in flight, summed in chunk order
          double a[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) a[i] = __ldcg(q + (size_t)(ch + i) * cs);
#pragma unroll
          for (int i = 0; i < 16; ++i) s += a[i];
        }
        for (; ch < chunks; ++ch) s += __ldcg(q + (size_t)ch * cs);
        if (d < D) {
          ls.zg[(size_t)d * C + c] = logistic_finish_grad(s, __ldcg(ls.zt + (size_t)d * C + c), is2);
        } else {
          double qq = 0.0;
          int j = 0;
          for (; j + 20 <= D; j += 20) {          // twenty loads in flight, squared and summed in order
            double a[20];
#pragma unroll
            for (int i = 0; i < 20; ++i) a[i] = __ldcg(ls.zt + (size_t)(j + i) * C + c);
#pragma unroll
            for (int i = 0; i < 20; ++i) qq = __fma_rn(a[i], a[i], qq);
          }
          for (; j < D; ++j) {
            const double t = __ldcg(ls.zt + (size_t)j * C + c);
            qq = __fma_rn(t, t, qq);
          }
          ls.ts.zlp[c] = logistic_finish_lp(s, qq, is2, D, sd);
        }
      }
    }
    lap(2);
    grid_barrier(pa.gbar, bar_target, nblk, pa.bar_mode);
    lap(3);

    // ---- P: per-chain state machine, one warp per listed chain, chains dealt over the CTAs ----
    if (b == 0 && threadIdx.x == 0 && (p & 63) == 0) *(volatile int *)pa.h_done = __ldcg(pa.n_done);
    if (etid >= 0) {
      const int w = etid >> 5;
      const int slot = w * (int)nblk + b;      // n <= kPersistMaxTiles * 128 <= kEpiWarps * grid
      if (slot < n) {
        const Tile t = Tile::of_warp(etid & 31, __ldcg(ls.list + slot));
        ls_pump_body<Team>(t, st, ls, rp, m, out, pa.n_done, nullptr, &s_maxlev[w]);
      }
    }
    lap(4);
    grid_barrier(pa.gbar, bar_target, nblk, pa.bar_mode);
    lap(5);
    if (__ldcg(pa.n_done) >= C) {          // every chain has finished its transitions
      ++p;
      reason = 2;
      break;
    }
  }
  if (threadIdx.x == 64 && pa.cta_times) {   // the first epilogue thread: it has work in every phase
    pa.cta_times[3 * b + 0] = tph[0];
    pa.cta_times[3 * b + 1] = tph[2];
    pa.cta_times[3 * b + 2] = tph[4];
  }
  if (b == 0 && threadIdx.x == 64) {
    pa.out_state[0] = p - pa.pump0;
    pa.out_state[1] = (long long)pass;
    pa.out_state[2] = __ldcg(pa.n_done);
    pa.out_state[3] = reason;
    for (int i = 0; i < 7; ++i) pa.out_state[4 + i] = tph[i];
    *(volatile int *)pa.h_done = __ldcg(pa.n_done);
  }
  core.teardown();
}

}  // namespace tc
}  // namespace tnuts
