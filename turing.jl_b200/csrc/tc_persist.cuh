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
  int n_max;                   // c_total - listed_done: bound of the listed chains
  int c_total;                 // chains of the engine (st.C is the width of the view the kernel works on)
  int exit_below;              // return at a list refresh with <= this many chains unfinished (0: never)
  int bar_mode;                // grid barrier variant (experiments)
  unsigned long long *gbar;    // grid barrier counter, zero at launch
  int *n_done;                 // finished chains (device)
  int *h_done;                 // mapped pinned mirror of n_done for the host's progress trace
  long long *cta_times;        // [grid][3] ns every CTA spent inside G, R, P (barrier waits excluded)
  long long *out_state;        // [4] pumps executed, gradient passes of CTA 0, final n_done, reason;
                               // [4..10] ns CTA 0 spent in G, barrier, R, barrier, P, barrier, refresh;
                               // [11] the refresh schedule's finished-chain count at the end;
                               // [12..14] ns inside G: positions -> TMEM, row tiles, drain + partial sums
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the (co-resident) grid; `target` is this thread's running arrival total.
// mode (experiments, TNUTS_BAR_MODE): 0 sc fences + acquire polls, 1 acq_rel fences + relaxed polls,
// 2 release arrival + acquire polls, 3 sc fences + volatile polls (cooperative groups' recipe)
// After the barrier completes, thread 0 also reads two device words (the finished-chain counter and
// the length of the work list) and hands them to the CTA through shared memory: 576 threads x 148
// CTAs loading the same sector right after every barrier cost ~30 us per pump.
__device__ __forceinline__ void grid_barrier(unsigned long long *bar, unsigned long long &target, unsigned nblk,
                                             int mode, const int *w0 = nullptr, const int *w1 = nullptr,
                                             int *s_words = nullptr) {
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
    if (s_words) {
      s_words[0] = __ldcg(w0);
      s_words[1] = __ldcg(w1);
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

__device__ __forceinline__ long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
  return (long long)t;
}

// The gradient pass is a separate, NOT inlined function: it gets the CTA's whole register budget (96
// per thread at 18 warps) for itself.  Inlined into the pump loop, the loop's live state (the pointers
// of five parameter structs, counters, timers) cost the tcgen05 epilogue 500 bytes of spills inside
// its per-tile loop and made a row tile take 3.1 us instead of 1.8.  (The reduction and the state
// machine stay inlined: as functions they read the kernel's parameter structs through generic
// pointers, which measured slower than the spills they avoid.)
struct GradPass {
  long long rbeg;
  int ntile, k0, n, C, cap;
  uint32_t tt, pass;
  const float *yf;
  long long nrows;
  const double *zt;
  const int *list;
  double *pc;          // partial sums of this row chunk
  long long *tgr;      // trace: ns inside the three parts (first epilogue thread of CTA 0), or null
};

template <int KP>
__device__ __noinline__ void persist_grad_pass(TcCore<KP> core, const CUtensorMap *mapX, GradPass g,
                                               double (*lp_part)[kChainsPerCta]) {
  if (core.warp == 0) {
    core.produce(mapX, g.rbeg, g.ntile, g.tt);
  } else if (core.warp == 1) {
    core.mma(g.ntile, g.tt, g.pass);
  } else {
    const int k = g.k0 + core.kl();
    const bool live = k < g.n;
    const int c = live ? __ldcg(g.list + k) : 0;
    const long long g0 = g.tgr ? gtime_ns() : 0;
    core.template epi_theta<true>(g.zt, g.C, c, live);
    const long long g1 = g.tgr ? gtime_ns() : 0;
    const double lp = core.epi_tiles(g.yf, g.nrows, g.rbeg, g.ntile, g.tt, nullptr, false);
    const long long g2 = g.tgr ? gtime_ns() : 0;
    core.epi_store(g.pc, g.cap, k, live, lp, lp_part, g.pass);
    if (g.tgr) {
      g.tgr[0] += g1 - g0;
      g.tgr[1] += g2 - g1;
      g.tgr[2] += gtime_ns() - g2;
    }
  }
}

// R: fixed-order sum of the row-chunk partials + prior -> zg, zlp (same order as k_logistic_reduce_finish).
// Latency-bound (a few hundred outputs, each the sum of up to 148 partials that sit in L2): what
// matters is the number of loads in flight.  Plain loads -- the grid barrier before this phase has
// invalidated the SM's L1, and nobody writes the partials or the positions during the phase -- in
// unpredicated batches, so that the compiler issues a whole batch before the first add (with ld.cg, or
// with a predicate per load, it alternates load and add: one L2 round trip per partial, 24 us).
// x, but not before `last` (the value of the batch's last load) has arrived: the compiler otherwise
// pipelines "load, add" four deep to save registers, and a batch of 32 loads becomes 8 L2 round trips.
// The bit pattern compared with is a signalling NaN no computation produces.
__device__ __forceinline__ double after_all(double x, double last) {
  return __double_as_longlong(last) == 0x7ff4dead0000beefLL ? last : x;
}

__device__ __noinline__ void persist_reduce(const double *__restrict__ part, int cap, int chunks, int n, int C,
                                            int D, int b, unsigned nblk, int etid, const int *__restrict__ list,
                                            const double *__restrict__ zt, double *__restrict__ zg,
                                            double *__restrict__ zlp, double is2, double sd) {
  const size_t cs = (size_t)(D + 1) * cap;
  const int total = (D + 1) * n;
  // 32 consecutive outputs per warp, warps dealt round-robin over the CTAs: every SM pulls its
  // share of the partials (a few CTAs alone are bound by their outstanding-load limit)
  const int wstep = (int)nblk * kEpiWarps * 32;
  for (int o = ((etid >> 5) * (int)nblk + b) * 32 + (etid & 31); o < total; o += wstep) {
    const int d = o / n, k = o - d * n;
    const int c = list[k];
    const double *q = part + (size_t)d * cap + k;
    const double th = d < D ? zt[(size_t)d * C + c] : 0.0;
    double s = 0.0;
    int ch = 0;
    for (; ch + 32 <= chunks; ch += 32) {
      double a[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = q[(size_t)(ch + i) * cs];
      s += after_all(a[0], a[31]);
#pragma unroll
      for (int i = 1; i < 32; ++i) s += a[i];
    }
    if (ch + 16 <= chunks) {
      double a[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = q[(size_t)(ch + i) * cs];
      s += after_all(a[0], a[15]);
#pragma unroll
      for (int i = 1; i < 16; ++i) s += a[i];
      ch += 16;
    }
    if (ch < chunks) {      // up to 15 left: clamp the index, then add only the real ones
      double a[15];
#pragma unroll
      for (int i = 0; i < 15; ++i) a[i] = q[(size_t)min(ch + i, chunks - 1) * cs];
      s += after_all(a[0], a[14]);
#pragma unroll
      for (int i = 1; i < 15; ++i)
        if (ch + i < chunks) s += a[i];
    }
    if (d < D) {
      zg[(size_t)d * C + c] = logistic_finish_grad(s, th, is2);
    } else {
      double qq = 0.0;
      int j = 0;
      for (; j + 25 <= D; j += 25) {          // 25 loads in flight, squared and summed in order
        double a[25];
#pragma unroll
        for (int i = 0; i < 25; ++i) a[i] = zt[(size_t)(j + i) * C + c];
        const double a0 = after_all(a[0], a[24]);
        qq = __fma_rn(a0, a0, qq);
#pragma unroll
        for (int i = 1; i < 25; ++i) qq = __fma_rn(a[i], a[i], qq);
      }
      for (; j < D; ++j) {
        const double t = zt[(size_t)j * C + c];
        qq = __fma_rn(t, t, qq);
      }
      zlp[c] = logistic_finish_lp(s, qq, is2, D, sd);
    }
  }
}

// P: the NUTS state machine of one listed chain on one warp
__device__ __forceinline__ void persist_pump_warp(int lane, int chain, const ChainState &st, const LockstepState &ls,
                                               const RunParams &rp, const ModelDev &m, const DrawBuffers &out,
                                               int *n_done, int *s_maxlev, long long *tprof) {
  const Tile t = Tile::of_warp(lane, chain);
  ls_pump_body<WarpTeam>(t, st, ls, rp, m, out, n_done, nullptr, s_maxlev, tprof);
}

template <int KP>
__global__ void __launch_bounds__(kThreads, 1)
k_tc_persist(const __grid_constant__ CUtensorMap mapX, const float *__restrict__ yf, long long nrows,
             double *part /* [chunks][D+1][cap] */, int cap, const __grid_constant__ ChainState st,
             const __grid_constant__ LockstepState ls, const __grid_constant__ RunParams rp,
             const __grid_constant__ ModelDev m, const __grid_constant__ DrawBuffers out, PersistCut cut,
             PersistArgs pa) {
  using Core = TcCore<KP>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[Core::NBARS];
  __shared__ uint32_t tmem_slot;
  __shared__ double lp_part[kEpiWarps / 4][kChainsPerCta];
  __shared__ int s_maxlev[kEpiWarps];
  __shared__ int wsum[33];
  __shared__ int s_pmax;
  __shared__ int s_words[2];      // finished chains, listed chains: read by thread 0 inside the barriers

  Core core;
  core.setup(smem_raw, bars, &tmem_slot, &mapX, st.D);
  const int C = st.C, D = st.D;             // width of the (possibly compacted) state arrays
  const int Ctot = pa.c_total;
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
  // only the first epilogue thread of a CTA keeps time: %globaltimer is one chip-wide register, and
  // 85 000 threads reading it after every barrier queue up for tens of microseconds
  const bool timekeeper = threadIdx.x == 64;
  auto now = [&] { return timekeeper ? gtime_ns() : 0LL; };
  long long tgr[3] = {0, 0, 0};   // inside G: positions -> TMEM, row tiles, drain + partial sums
  long long tpp[7] = {0, 0, 0, 0, 0, 0, 0};   // inside P (first state-machine warp of CTA 0), per section
  long long tmark = now();
  long long tx[2] = {0, 0};
  auto lapx = [&](int i) {   // sub-laps of the G slot (their time is NOT added to tph[0])
    const long long t = now();
    tx[i] += t - tmark;
    tmark = t;
  };
  auto lap = [&](int i) {
    const long long t = now();
    tph[i] += t - tmark;
    tmark = t;
  };
  const double sd = m.hyper[0], is2 = 1.0 / (sd * sd);
  if (threadIdx.x == 0) s_pmax = 0;
  int n = __ldcg(ls.count);                 // listed chains (changes at list refreshes only)
  int n_done_now = __ldcg(pa.n_done);       // finished chains as of the last barrier

  for (; p < pa.max_pumps; ++p) {
    // ---- the host loop's schedule of counter snapshots and work-list refreshes ----
    if (p % 16 == 0 && p > 0) snap = n_done_now;                 // the count after pump p - 1
    else if (p % 16 == 8 && p > 16) done = snap;
    if (done >= Ctot) {
      reason = 1;
      break;
    }
    if (done != listed_done && (p % 8 == 0)) {
      if (Ctot - done <= pa.exit_below) {   // few enough chains for a narrower view: back to the host
        reason = 3;
        break;
      }
      if (b == 0) persist_compact(ls.ts.phase, C, ls.list, ls.count, wsum);
      listed_done = done;
      n_max = Ctot - done;
      grid_barrier(pa.gbar, bar_target, nblk, pa.bar_mode, pa.n_done, ls.count, s_words);
      n = s_words[1];
      lap(6);
    }
    lapx(0);
    lapx(1);
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
        GradPass g;
        g.rbeg = rbeg;
        g.ntile = ntile;
        g.k0 = k0;
        g.n = n;
        g.C = C;
        g.cap = cap;
        g.tt = tt;
        g.pass = pass;
        g.yf = yf;
        g.nrows = nrows;
        g.zt = ls.zt;
        g.list = ls.list;
        g.pc = part + (size_t)chunk * (D + 1) * cap;
        g.tgr = (pa.cta_times && b == 0 && threadIdx.x == 64) ? tgr : nullptr;
        persist_grad_pass<KP>(core, &mapX, g, lp_part);
        tt += (uint32_t)ntile;
        pass += 1;
      }
    }
    lap(0);
    grid_barrier(pa.gbar, bar_target, nblk, pa.bar_mode);
    lap(1);

    // ---- R: sum of the row-chunk partials in chunk order + prior -> zg, zlp (k_logistic_reduce_finish) ----
    if (grad_pass && etid >= 0)
      persist_reduce(part, cap, chunks, n, C, D, b, nblk, etid, ls.list, ls.zt, ls.zg, ls.ts.zlp, is2, sd);
    lap(2);
    grid_barrier(pa.gbar, bar_target, nblk, pa.bar_mode);
    lap(3);

    // ---- P: per-chain state machine, one warp per listed chain, chains dealt over the CTAs ----
    if (b == 0 && threadIdx.x == 0 && (p & 63) == 0) *(volatile int *)pa.h_done = __ldcg(pa.n_done);
    if (etid >= 0) {
      const int w = etid >> 5;
      const int slot = w * (int)nblk + b;      // n <= kPersistMaxTiles * 128 <= kEpiWarps * grid
      if (slot < n) {
        const long long w0 = (pa.cta_times && (etid & 31) == 0) ? gtime_ns() : 0;
        persist_pump_warp(etid & 31, __ldcg(ls.list + slot), st, ls, rp, m, out, pa.n_done, &s_maxlev[w],
                          (pa.cta_times && b == 0 && w == 0) ? tpp : nullptr);
        if (pa.cta_times && (etid & 31) == 0) atomicMax(&s_pmax, (int)(gtime_ns() - w0));
      }
    }
    lap(4);
    grid_barrier(pa.gbar, bar_target, nblk, pa.bar_mode, pa.n_done, ls.count, s_words);
    n_done_now = s_words[0];
    lap(5);
    if (timekeeper) {        // trace: the slowest state-machine warp of this CTA in this pump
      tx[0] += s_pmax;
      s_pmax = 0;
    }
    if (n_done_now >= Ctot) {              // every chain has finished its transitions
      ++p;
      reason = 2;
      break;
    }
  }
  if (threadIdx.x == 64 && pa.cta_times) {   // the first epilogue thread: it has work in every phase
    pa.cta_times[5 * b + 0] = tph[0];
    pa.cta_times[5 * b + 1] = tph[2];
    pa.cta_times[5 * b + 2] = tph[4];
    pa.cta_times[5 * b + 3] = tx[0];
    pa.cta_times[5 * b + 4] = tx[1];
  }
  if (b == 0 && threadIdx.x == 64) {
    pa.out_state[0] = p - pa.pump0;
    pa.out_state[1] = (long long)pass;
    pa.out_state[2] = __ldcg(pa.n_done);
    pa.out_state[3] = reason;
    for (int i = 0; i < 7; ++i) pa.out_state[4 + i] = tph[i];
    pa.out_state[11] = done;
    for (int i = 0; i < 3; ++i) pa.out_state[12 + i] = tgr[i];
    for (int i = 0; i < 7; ++i) pa.out_state[15 + i] = tpp[i];
    *(volatile int *)pa.h_done = __ldcg(pa.n_done);
  }
  core.teardown();
}

}  // namespace tc
}  // namespace tnuts
