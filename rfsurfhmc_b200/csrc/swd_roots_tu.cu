// Translation unit of the phase-velocity root search (K1 thread-mapped, K1t team-mapped, retry).
// Compiled with -fmad=false (see swd_roots_launch.h): every kernel in here evaluates the secular
// function with the same roundings, so all mappings return the same bits.
#include "swd_roots_team.cuh"

namespace rfs {

// ---- K1: one thread per (model, sequence)
// croot : [nmode_out][nsolve][B]   cwork : [nsolve][B] (only touched when nmode > 1)
// ierr  : [nseq][B] int
#ifndef RFS_ROOTS_MINBLOCKS
#define RFS_ROOTS_MINBLOCKS 4
#endif
#ifndef RFS_ROOTS_BLOCK
#define RFS_ROOTS_BLOCK 128
#endif
// perm (optional): the length-sorted job order of swd_sched_sort_kernel, [nseq][Bp] model indices
// (-1 = padding), Bp = B rounded up to a multiple of 32 so that a warp never mixes sequences.
__global__ void __launch_bounds__(RFS_ROOTS_BLOCK, RFS_ROOTS_MINBLOCKS)
    swd_roots_kernel(SwdPlan plan, SwdBlocks blk, long long B, int n,
                     const double *__restrict__ periods, int all_modes,
                     double *__restrict__ croot, double *__restrict__ cwork,
                     int *__restrict__ ierr, unsigned long long *__restrict__ neval_total,
                     const int *__restrict__ perm, long long Bp) {
  __shared__ double wsm_all[RFS_ROOTS_BLOCK / 32][33];
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  bool valid;
  long long b;
  int s;
  if (perm) {
    valid = i < Bp * plan.nseq;
    s = valid ? (int)(i / Bp) : 0;
    const int pb = valid ? __ldg(perm + i) : -1;
    valid = pb >= 0;
    b = valid ? pb : 0;
  } else {
    valid = i < B * plan.nseq;
    b = valid ? i % B : 0;
    s = valid ? (int)(i / B) : 0;
  }
  SwdModel M(blk.root[plan.seq[s].ifunc == 2 ? 0 : 1], B, n);
  unsigned int nev = 0;
  const int e = swd_solve_sequence(M, b, plan.seq[s], periods, plan.nmode, all_modes, croot,
                                   (long long)plan.nsolve * B, cwork, B, nev, valid,
                                   wsm_all[threadIdx.x >> 5], -1);
  if (valid) ierr[(long long)s * B + b] = e;
  if (neval_total) {
    // one aggregated atomic per warp: algorithmic-work counter for the roofline (bench.py)
    unsigned int w = nev;
    for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(neval_total, (unsigned long long)w);
    atomicMax(neval_total + 1, (unsigned long long)nev);          // slowest thread
    if (nev > 2000u) atomicAdd(neval_total + 2, 1ull);            // heavy threads (> 2000 evals)
  }
}

// ---- length-sorted scheduling of the thread-mapped search (large batches).
// A sequence costs (c(T_longest) - cc) / dc scan steps plus a few refinement steps per period
// (the scan of period k starts 1.5 dc below the root of period k-1, surfdisp96.f:271), and models
// differ: on the C1 sampler set 830 +- 125 evaluations per sequence, so a warp of 32 unrelated
// models waits for its slowest lane with 21 % of its lane-time idle.  swd_sched_key_kernel
// estimates c(T_longest) of every job with RFS_SCHED_BISECT bisection steps of the real secular
// function between the start value cc and the fastest layer; swd_sched_sort_kernel orders each
// sequence's jobs by that estimate (longest first), and swd_roots_kernel then runs one warp per
// block over 32 jobs of near-equal length.  The order changes which lane solves which job and
// nothing else: every job is solved by the same code on the same inputs.
#define RFS_SCHED_BISECT 6
#define RFS_SCHED_BINS 1024
__global__ void __launch_bounds__(128)
    swd_sched_key_kernel(SwdPlan plan, SwdBlocks blk, long long B, int n,
                         const double *__restrict__ periods, unsigned int *__restrict__ key) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= B * plan.nseq) return;
  const long long b = i % B;
  const SwdSeq sq = plan.seq[(int)(i / B)];
  SwdModel M(blk.root[sq.ifunc == 2 ? 0 : 1], B, n);
  int llw;
  float betmx, cc1;
  swd_start_values(M, b, llw, betmx, cc1);
  double tmax = 0.0;
  for (int k = 0; k < sq.nper; k++) tmax = fmax(tmax, __ldg(periods + sq.per_off + k));
  unsigned int kv = 0;
  if (sq.nper > 0 && tmax > 0.0 && betmx > cc1) {
    const double omega = 2.0 * RFS_PI64 / (tmax * sq.scale), iomega = 1.0 / omega;
    auto f = [&](double c) {
      const double wv = omega / c;
      return (sq.ifunc == 1) ? dltar1_dev(wv, omega, M, b, llw) : dltar4_dev(wv, omega, iomega, M, b, llw);
    };
    double lo = (double)cc1, hi = (double)betmx;
    const double s0 = sgn1(f(lo));
    for (int it = 0; it < RFS_SCHED_BISECT; it++) {
      const double mid = 0.5 * (lo + hi);
      if (sgn1(f(mid)) != s0) hi = mid; else lo = mid;
    }
    const double steps = (0.5 * (lo + hi) - (double)cc1) / (double)0.005f;
    kv = (unsigned int)fmin(fmax(steps, 0.0), 1.0e6);
  }
  key[i] = kv;
}

// one block per sequence: counting sort of its B keys into RFS_SCHED_BINS bins, largest first
__global__ void __launch_bounds__(1024)
    swd_sched_sort_kernel(const unsigned int *__restrict__ key, long long B, long long Bp,
                          int *__restrict__ perm) {
  __shared__ unsigned int s_lo, s_hi;
  __shared__ int hist[RFS_SCHED_BINS];
  __shared__ int wsum[32];
  const int t = threadIdx.x;
  const unsigned int *k = key + (long long)blockIdx.x * B;
  int *out = perm + (long long)blockIdx.x * Bp;
  if (t == 0) { s_lo = 0xffffffffu; s_hi = 0u; }
  for (int j = t; j < RFS_SCHED_BINS; j += blockDim.x) hist[j] = 0;
  __syncthreads();
  unsigned int lo = 0xffffffffu, hi = 0u;
  for (long long j = t; j < B; j += blockDim.x) {
    const unsigned int v = __ldg(k + j);
    lo = min(lo, v);
    hi = max(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((t & 31) == 0) { atomicMin(&s_lo, lo); atomicMax(&s_hi, hi); }
  __syncthreads();
  lo = s_lo;
  const unsigned long long span = (unsigned long long)(s_hi - lo) + 1ull;
  // bin 0 holds the largest keys
  auto bin_of = [&](unsigned int v) {
    return RFS_SCHED_BINS - 1 - (int)(((unsigned long long)(v - lo) * RFS_SCHED_BINS) / span);
  };
  for (long long j = t; j < B; j += blockDim.x) atomicAdd(&hist[bin_of(__ldg(k + j))], 1);
  __syncthreads();
  // exclusive prefix sum over the bins (blockDim.x == RFS_SCHED_BINS == 1024)
  const int h = hist[t];
  int inc = h;
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, inc, o);
    if ((t & 31) >= o) inc += u;
  }
  if ((t & 31) == 31) wsum[t >> 5] = inc;
  __syncthreads();
  if (t < 32) {
    int w = wsum[t];
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, w, o);
      if (t >= o) w += u;
    }
    wsum[t] = w;
  }
  __syncthreads();
  hist[t] = inc - h + ((t >> 5) ? wsum[(t >> 5) - 1] : 0);
  __syncthreads();
  for (long long j = t; j < B; j += blockDim.x) out[atomicAdd(&hist[bin_of(__ldg(k + j))], 1)] = (int)j;
  for (long long j = B + t; j < Bp; j += blockDim.x) out[j] = -1;
}

// ---- per-period retries of _surfdisp (surfdisp.cpp:93-100): when the fundamental mode failed in
// the main pass, every period whose reported value is zero / NaN is searched again as a fresh
// single-period problem (start value cc, scan upward in dc steps: hundreds of evaluations).
// The reference does these one after the other and stops at the first one that fails again; the
// jobs are independent, so they run here as one thread per (model, sequence, period) — in a warp
// that is otherwise idle, whose 31 spare lanes take over the look-ahead scan — and
// swd_retry_finish_kernel re-imposes the sequential stop rule.
//   rstat [nsolve][B] int: -1 not retried, 0 retried ok, 1 retried and failed again
__global__ void __launch_bounds__(RFS_ROOTS_BLOCK, RFS_ROOTS_MINBLOCKS)
    swd_retry_kernel(SwdPlan plan, SwdBlocks blk, long long B, int n,
                     const double *__restrict__ periods, int all_modes,
                     double *__restrict__ croot, double *__restrict__ cwork,
                     const int *__restrict__ ierr, int *__restrict__ rstat,
                     unsigned long long *__restrict__ neval_total) {
  __shared__ double wsm_all[RFS_ROOTS_BLOCK / 32][33];
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const bool inrange = i < B * plan.nsolve;
  const long long b = inrange ? i % B : 0;
  const int solve = inrange ? (int)(i / B) : 0;
  int s = 0;
  for (int q = 0; q < plan.nseq; q++)
    if (solve >= plan.seq[q].out_off && solve < plan.seq[q].out_off + plan.seq[q].nper) s = q;
  const int k = solve - plan.seq[s].out_off;
  bool valid = false;
  if (inrange && ierr[(long long)s * B + b] != 0) {
    const double *clast = croot + (all_modes ? (long long)(plan.nmode - 1) * plan.nsolve * B : 0);
    const double v = clast[(long long)solve * B + b];
    valid = (v == 0.0 || isnan(v));
  }
  // whole warps without work leave (the warp-cooperative loop needs all 32 lanes of a live warp)
  if (__ballot_sync(0xffffffffu, valid) == 0u) {
    if (inrange) rstat[(long long)solve * B + b] = -1;
    return;
  }
  SwdModel M(blk.root[plan.seq[s].ifunc == 2 ? 0 : 1], B, n);
  unsigned int nev = 0;
  const int e = swd_solve_sequence(M, b, plan.seq[s], periods, plan.nmode, all_modes, croot,
                                   (long long)plan.nsolve * B, cwork, B, nev, valid,
                                   wsm_all[threadIdx.x >> 5], k);
  if (inrange) rstat[(long long)solve * B + b] = valid ? e : -1;
  if (neval_total) {
    unsigned int w = nev;
    for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(neval_total, (unsigned long long)w);
  }
}


static inline unsigned grid_for(long long total, int block) {
  return (unsigned)((total + block - 1) / block);
}

cudaError_t launch_roots_thread(const SwdPlan &P, const SwdBlocks &blk, long long B, int n,
                                const double *periods, int all_modes, double *croot, double *cwork,
                                int *ierr, unsigned long long *counter, const int *perm,
                                cudaStream_t st) {
  if (perm) {
    // sorted order: one warp per block, so that the block scheduler spreads the (longest-first)
    // warps evenly over the SMs
    const long long Bp = (B + 31) / 32 * 32;
    swd_roots_kernel<<<grid_for(Bp * P.nseq, 32), 32, 0, st>>>(P, blk, B, n, periods, all_modes, croot,
                                                               cwork, ierr, counter, perm, Bp);
  } else {
    swd_roots_kernel<<<grid_for(B * P.nseq, RFS_ROOTS_BLOCK), RFS_ROOTS_BLOCK, 0, st>>>(
        P, blk, B, n, periods, all_modes, croot, cwork, ierr, counter, nullptr, 0);
  }
  return cudaGetLastError();
}

cudaError_t launch_sched_keys(const SwdPlan &P, const SwdBlocks &blk, long long B, int n,
                              const double *periods, unsigned int *key, cudaStream_t st) {
  swd_sched_key_kernel<<<grid_for(B * P.nseq, 128), 128, 0, st>>>(P, blk, B, n, periods, key);
  return cudaGetLastError();
}

cudaError_t launch_sched_sort(const SwdPlan &P, long long B, const unsigned int *key, int *perm,
                              cudaStream_t st) {
  swd_sched_sort_kernel<<<P.nseq, RFS_SCHED_BINS, 0, st>>>(key, B, (B + 31) / 32 * 32, perm);
  return cudaGetLastError();
}

cudaError_t launch_roots_retry(const SwdPlan &P, const SwdBlocks &blk, long long B, int n,
                               const double *periods, int all_modes, double *croot, double *cwork,
                               const int *ierr, int *rstat, unsigned long long *counter,
                               cudaStream_t st) {
  swd_retry_kernel<<<grid_for(B * P.nsolve, RFS_ROOTS_BLOCK), RFS_ROOTS_BLOCK, 0, st>>>(
      P, blk, B, n, periods, all_modes, croot, cwork, ierr, rstat, counter);
  return cudaGetLastError();
}

bool team_shape_supported(int T, int S) {
  static const int ok[][2] = {{2, 2}, {4, 1}, {4, 4}, {8, 1}, {8, 2}, {8, 8}, {16, 1}, {16, 2},
                              {32, 1}, {32, 2}, {32, 4}};
  for (auto &p : ok)
    if (p[0] == T && p[1] == S) return true;
  return false;
}

template <int T, int S>
static cudaError_t launch_team(const SwdPlan &P, const SwdBlocks &blk, long long B, int n,
                               const double *periods, int all_modes, double *croot, double *cwork,
                               int *ierr, unsigned long long *counter, cudaStream_t st) {
  // threads per block: as many teams as fit ~64 KB of staged layer parameters
  int threads = 128;
  const size_t per_team = sizeof(double) * RFS_TEAM_NF * (size_t)n;
  while (threads > 32 && threads > T && (threads / T) * per_team > 64 * 1024) threads /= 2;
  const size_t sm = (((threads / T) * per_team + 15) & ~(size_t)15) +
                    (size_t)(threads / 32) * RFS_TEAM_NP * 32 * sizeof(double2);
  if (sm > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(swd_roots_team_kernel<T, S>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return e;
  }
  swd_roots_team_kernel<T, S><<<grid_for(B * P.nseq, threads / T), threads, sm, st>>>(
      P, blk, B, n, periods, all_modes, croot, cwork, ierr, counter);
  return cudaGetLastError();
}

cudaError_t launch_roots_team(int T, int S, const SwdPlan &P, const SwdBlocks &blk, long long B, int n,
                              const double *periods, int all_modes, double *croot, double *cwork,
                              int *ierr, unsigned long long *counter, cudaStream_t st) {
#define RFS_TEAM(TT, SS) \
  if (T == TT && S == SS) return launch_team<TT, SS>(P, blk, B, n, periods, all_modes, croot, cwork, ierr, counter, st);
  RFS_TEAM(2, 2) RFS_TEAM(4, 1) RFS_TEAM(4, 4) RFS_TEAM(8, 1) RFS_TEAM(8, 2) RFS_TEAM(8, 8) RFS_TEAM(16, 1)
  RFS_TEAM(16, 2)
  RFS_TEAM(32, 1) RFS_TEAM(32, 2) RFS_TEAM(32, 4)
#undef RFS_TEAM
  return cudaErrorInvalidValue;
}

}  // namespace rfs
