/* rfsurfhmc.h — C ABI of librfsurf_b200.so (hand-written sm_100a CUDA behind plain C).
 *
 * Drop-in boundary for the hot path of nqdu/RfSurfHmc.  Every entry point names the reference
 * interface it replaces (paths relative to the reference repository root):
 *
 *   libsurf.forward / libsurf.adjoint_kernel      src/SWD/main.cpp:14-82  (pybind11 module :84-94)
 *     -> surfdisp96_, sregn96_, sregnpu_, slegn96_, slegnpu_   src/SWD/surfdisp.hpp:17-93
 *   librf.forward / librf.kernel / librf.kernel_all             src/RF/main.cpp:17-189 (:191-213)
 *     -> cal_rf_freq_, cal_rf_par_freq_, cal_rf_par_freq_all_   src/RF/rf_cal.hpp:10-51
 *   Joint_RF_SWD.misfit_and_grad (+ SurfWD / ReceiverFunc)      model/model_rf_swd_vs_thk.py:66-86,
 *                                                               model/model_surf.py:155-228,
 *                                                               model/model_rf.py:137-197
 *   HamitonianMC.sample / HMCDualAveraging.sample               pyhmc/hmc.py:228-276,
 *                                                               pyhmc/hmcda.py:280-369
 *
 * Conventions
 *   - all entry points return 0 on success, a negative RFS_E_* code otherwise;
 *     rfs_last_error(ctx) returns a human-readable message for the last failure on that context.
 *   - no torch / C++ types cross the boundary: plain pointers, sizes and an opaque context.
 *   - *_dev functions take DEVICE pointers and a cudaStream_t (passed as void*); they never
 *     allocate in steady state (workspace is grown on the first call for a given batch size) and
 *     are stream-ordered.  *_host functions take HOST pointers and perform the H2D/D2H copies.
 *   - arrays are row-major, float64 unless stated; batch index first.
 *   - wave types: 0 "Rc", 1 "Rg", 2 "Lc", 3 "Lg".  rf_type: 1 P, 2 S.  method: 0 time, 1 freq.
 *   - there is NO CPU fallback: if no CUDA device is usable rfs_create fails.
 */
#ifndef RFSURFHMC_H
#define RFSURFHMC_H

#ifdef __cplusplus
extern "C" {
#endif

#define RFS_OK 0
#define RFS_E_CUDA -1      /* CUDA runtime error (message has the detail) */
#define RFS_E_ARG -2       /* invalid argument (bad wave type / rf type / sizes) */
#define RFS_E_UNSUPPORTED -3 /* reference feature not built (currently unused) */
#define RFS_E_CONFIG -4    /* context not configured for this call */

typedef struct rfs_ctx rfs_ctx;

/* ---- lifetime ------------------------------------------------------------------------------ */
int rfs_create(rfs_ctx **out, int device);
void rfs_destroy(rfs_ctx *ctx);
const char *rfs_last_error(rfs_ctx *ctx);
/* version string of the library, e.g. "rfsurf_b200 0.1 sm_100a" */
const char *rfs_version(void);
/* number of kernels launched by this context since creation (bench.py's gpu_launches) */
long long rfs_launch_count(rfs_ctx *ctx);

/* ---- configuration (replaces SurfWD.__init__, ReceiverFunc.__init__, Joint_RF_SWD.__init__ +
 *      set_obsdata: model/model_surf.py:5-29, model/model_rf.py:5-18,
 *      model/model_rf_swd_vs_thk.py:6-25) -------------------------------------------------------- */
/* period lists may be empty (n*=0).  mode: 0 fundamental.  sphere: 1 = earth-flattening transformation.
 * stale_group_kernel=1 reproduces sregnpu/slegnpu's stale first term (sregn96.f90:1841-1844). */
int rfs_config_swd(rfs_ctx *ctx, int nlayer, int ntRc, const double *tRc, int ntRg,
                   const double *tRg, int ntLc, const double *tLc, int ntLg, const double *tLg,
                   int mode, int sphere, int stale_group_kernel);
int rfs_config_rf(rfs_ctx *ctx, int nlayer, double ray_p, int nt, double dt, double gauss,
                  double time_shift, double water, int rf_type, int method);
/* Extensions of the two calls above for objectives with several modes / several ray parameters
 * (BASELINE configs 2 and 3; the reference's classes hold one `mode` and one `ray_p`, so such data
 * sets need one model object per mode / ray parameter there).  modes[] ascending; the SWD data
 * vector becomes [mode][Rc,Rg,Lc,Lg], the RF data vector [ray parameter][nt].  All modes are solved
 * in ONE chained root-search pass (surfdisp96.f:232-368 solves modes 1..mode+1 anyway). */
int rfs_config_swd_modes(rfs_ctx *ctx, int nlayer, int ntRc, const double *tRc, int ntRg,
                         const double *tRg, int ntLc, const double *tLc, int ntLg, const double *tLg,
                         int nmodes, const int *modes, int sphere, int stale_group_kernel);
int rfs_config_rf_rays(rfs_ctx *ctx, int nlayer, int nray, const double *ray_p, int nt, double dt,
                       double gauss, double time_shift, double water, int rf_type, int method);
/* dobs (host pointer): [nt_rf + n_swd] joint, or the matching sub-vector for which=1/2 */
int rfs_config_obs(rfs_ctx *ctx, double sigma1, double sigma2, const double *dobs, int ndobs);

/* ---- the hot path: batched misfit + gradient ------------------------------------------------
 * x [B][2n] = [vs(n), thk(n)];  U [B];  grad [B][2n];  dsyn [B][ndata];  flag [B] (1 ok / 0 fail)
 * which: 0 Joint_RF_SWD.misfit_and_grad, 1 ReceiverFunc.misfit_and_grad, 2 SurfWD.misfit_and_grad */
int rfs_misfit_grad_dev(rfs_ctx *ctx, long long B, const double *x, int which, double *U,
                        double *grad, double *dsyn, unsigned char *flag, void *stream);
int rfs_misfit_grad_host(rfs_ctx *ctx, long long B, const double *x, int which, double *U,
                         double *grad, double *dsyn, unsigned char *flag);

/* ---- libsurf drop-ins, batched over B models (host pointers) --------------------------------
 * thk,vp,vs,rho [B][n] (cast to float32 inside, as src/SWD/main.cpp:9);  period [nT]
 * c [B][nT];  dcda,dcdb,dcdr,dcdh [B][nT][n];  ok [B] */
int rfs_surf_forward(rfs_ctx *ctx, long long B, int n, const double *thk, const double *vp,
                     const double *vs, const double *rho, int nT, const double *period,
                     int wavetype, int mode, int sphere, double *c, unsigned char *ok);
int rfs_surf_adjoint_kernel(rfs_ctx *ctx, long long B, int n, const double *thk, const double *vp,
                            const double *vs, const double *rho, int nT, const double *period,
                            int wavetype, int mode, int sphere, int stale_group_kernel, double *c,
                            double *dcda, double *dcdb, double *dcdr, double *dcdh,
                            unsigned char *ok);
/* all modes 0..mode at once (BASELINE config 2): c [B][mode+1][nT], kernels [B][mode+1][nT][n] */
int rfs_surf_adjoint_kernel_modes(rfs_ctx *ctx, long long B, int n, const double *thk,
                                  const double *vp, const double *vs, const double *rho, int nT,
                                  const double *period, int wavetype, int mode,
                                  int stale_group_kernel, double *c, double *dcda, double *dcdb,
                                  double *dcdr, double *dcdh, unsigned char *ok);

/* ---- librf drop-ins, batched over B models (host pointers) ----------------------------------
 * thk,rho,vp,vs,qa,qb [B][n] (RF argument order, src/RF/main.cpp:17-19);  rf [B][nt]
 * par_type: 1 rho, 2 vp, 3 vs, 4 thickness;  drf [B][n][nt] (kernel) or [B][4][n][nt] (kernel_all) */
int rfs_rf_forward(rfs_ctx *ctx, long long B, int n, const double *thk, const double *rho,
                   const double *vp, const double *vs, const double *qa, const double *qb,
                   double ray_p, int nt, double dt, double gauss, double time_shift, int method,
                   double water, int rf_type, double *rf);
int rfs_rf_kernel(rfs_ctx *ctx, long long B, int n, const double *thk, const double *rho,
                  const double *vp, const double *vs, const double *qa, const double *qb,
                  double ray_p, int nt, double dt, double gauss, double time_shift, int method,
                  double water, int rf_type, int par_type, double *rf, double *drf);
int rfs_rf_kernel_all(rfs_ctx *ctx, long long B, int n, const double *thk, const double *rho,
                      const double *vp, const double *vs, const double *qa, const double *qb,
                      double ray_p, int nt, double dt, double gauss, double time_shift, int method,
                      double water, int rf_type, double *rf, double *drf);

/* ---- device-resident HMC (replaces the Python loops of pyhmc/hmc.py:140-276 and
 *      pyhmc/hmcda.py:170-369; chain i uses NumPy-legacy MT19937 seeded with seed+chain_id[i]) ---
 * which  : objective sampled, as in rfs_misfit_grad_dev (0 Joint_RF_SWD, 1 ReceiverFunc, 2 SurfWD:
 *          the reference samplers take any object with misfit_and_grad, pyhmc/hmc.py:113-119).
 * sampler: 0 HamitonianMC (fixed dt, L ~ randint[Lmin,Lmax]), 1 HMCDualAveraging (L0, target;
 *          Lmax > 0 caps L = max(1,int(lambda/dt)) — an extension, 0 = unlimited as the reference).
 * bounds [2n][2] (low, high), shared by all chains (host pointer).
 * Outputs (host pointers, may be NULL): samples [C][nsamples][2n], misfit [C][nsamples],
 * syn [C][nsamples][ndata], initmodel [C][2n], n_iter [C] (trajectories run), n_acc [C],
 * dt_final [C], accept_seq [C][max_iter_log] (1/0 per trajectory, -1 padding; for RNG parity tests).
 * max_iters bounds the number of trajectories per chain (0 = unlimited). */
int rfs_hmc_run(rfs_ctx *ctx, int sampler, int which, long long C, const long long *chain_id,
                const double *bounds, double dt, int Lmin, int Lmax, int L0, double target_ratio,
                long long seed, int nsamples, int ndraws, long long max_iters, double *samples,
                double *misfit, double *syn, double *initmodel, long long *n_iter,
                long long *n_acc, double *dt_final, signed char *accept_seq,
                long long max_iter_log);
/* number of misfit_and_grad evaluations / of global steps (one batched evaluation each) performed by
 * the last rfs_hmc_run */
long long rfs_hmc_last_evals(rfs_ctx *ctx);
long long rfs_hmc_last_steps(rfs_ctx *ctx);
/* Options of rfs_hmc_run (no reference counterpart).  resident_slots > 0: at most that many chains are
 * resident on the device at a time; a finished chain hands its slot to the next queued chain on the
 * device, so the evaluated batch stays full (results do not depend on it).  max_seconds > 0: stop the
 * run after that wall-clock time; unfinished chains keep what they have (rfs_hmc_run returns 1). */
int rfs_set_hmc_options(rfs_ctx *ctx, long long resident_slots, double max_seconds);

/* ---- measurement helpers (no reference counterpart; used by bench.py) --------------------------
 * rfs_count_evals(ctx,1) zeroes and enables a device counter of secular-function evaluations
 * (the algorithmic work of the root-search kernel); rfs_read_evals synchronises and reads it.
 * rfs_measure_fp64_peak runs a DFMA micro-benchmark: the roofline denominator of the FP64 path. */
int rfs_count_evals(rfs_ctx *ctx, int enable);
/* Root-search mapping (no reference counterpart; results are bit-identical for every mapping):
 * T < 0 automatic by batch size (default), T = 0 one thread per (model, period sequence),
 * T in {2,4,8,16,32}: a team of T lanes per sequence, S speculative scan points per round
 * (csrc/swd_roots_team.cuh).  Environment override at rfs_create: RFS_ROOTS_TEAM="T,S". */
int rfs_set_roots_team(rfs_ctx *ctx, int T, int S);
int rfs_last_roots_team(rfs_ctx *ctx, int *T, int *S);
/* Length-sorted job order of the thread-mapped root search (no reference counterpart): large batches
 * are solved longest-job-first through a permutation built on the device (a bisection estimate of the
 * phase velocity at the longest period predicts the scan length of every (model, sequence) job).  The
 * order changes which lane solves which job and nothing else: results are bit-identical.
 * mode < 0 automatic (on for large, throughput-bound batches; default), 0 off, 1 on.
 * rfs_last_roots_sched: 1 if the last root search ran in sorted order.  Environment override at
 * rfs_create: RFS_ROOTS_SCHED. */
int rfs_set_roots_sched(rfs_ctx *ctx, int mode);
int rfs_last_roots_sched(rfs_ctx *ctx);
long long rfs_read_evals(rfs_ctx *ctx);
int rfs_read_eval_stats(rfs_ctx *ctx, long long *out3); /* total, slowest thread, threads > 2000 */
int rfs_measure_fp64_peak(rfs_ctx *ctx, double *tflops);
/* One rfs_misfit_grad_dev call with CUDA events around every kernel launch (the RF branch is
 * serialised behind the SWD branch for this call, so each kernel is timed alone).  ms[RFS_PROF_NK] =
 * milliseconds per kernel class, launches[RFS_PROF_NK] (may be NULL) = launch counts;
 * rfs_profile_kernel_name(i) names class i.  bench.py derives every roofline figure from this. */
#define RFS_PROF_NK 10
int rfs_profile_eval(rfs_ctx *ctx, long long B, const double *x, int which, double *U, double *grad,
                     double *dsyn, unsigned char *flag, void *stream, double *ms,
                     long long *launches);
const char *rfs_profile_kernel_name(int i);
/* self-test: the constant-bank exp / sincos / rsqrt of the root search against the CUDA math library
 * on n device-generated arguments; mismatches[6] = exp, sin/cos (large args), sin/cos (small), rsqrt */
int rfs_selftest_math(rfs_ctx *ctx, long long n, long long *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* RFSURFHMC_H */
