// Host driver of the device-resident HMC (included at the end of capi.cu).
// Replaces HamitonianMC.sample (pyhmc/hmc.py:228-276) and HMCDualAveraging.sample
// (pyhmc/hmcda.py:280-369) for C chains at once; see hmc_kernels.cuh for the per-chain logic.

namespace {

struct Carver {
  char *base;
  size_t off = 0;
  template <class T>
  T *take(size_t count) {
    off = (off + 255) & ~(size_t)255;
    T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
    off += sizeof(T) * count;
    return p;
  }
};

// R resident slots (trajectory state), C chains (results)
void carve_hmc(Carver &cv, HmcDev &D, long long R, long long C, int n2, int nd) {
  D.xcur = cv.take<double>(R * n2);
  D.xnew = cv.take<double>(R * n2);
  D.pnew = cv.take<double>(R * n2);
  D.gcur = cv.take<double>(R * n2);
  D.dcur = cv.take<double>(R * nd);
  D.xeval = cv.take<double>(R * n2);
  D.Ucur = cv.take<double>(R);
  D.Hcur = cv.take<double>(R);
  D.dt = cv.take<double>(R);
  D.dtbar = cv.take<double>(R);
  D.h0 = cv.take<double>(R);
  D.fdH = cv.take<double>(R);
  D.gauss = cv.take<double>(R);
  D.phase = cv.take<int>(R);
  D.istep = cv.take<int>(R);
  D.L = cv.take<int>(R);
  D.okcur = cv.take<int>(R);
  D.fd_it = cv.take<int>(R);
  D.fd_a = cv.take<int>(R);
  D.mti = cv.take<int>(R);
  D.has_gauss = cv.take<int>(R);
  D.iacc = cv.take<long long>(R);
  D.ncount = cv.take<long long>(R);
  D.mt = cv.take<unsigned int>(R * 624);
  D.n_active = cv.take<int>(4);
  D.queue_head = cv.take<int>(4);
  D.slot = cv.take<int>(R);
  D.idx = cv.take<int>(R);
  D.chain = cv.take<int>(R);
  D.xg = cv.take<double>(R * n2);
  D.status = cv.take<int>(C);
  D.nevals = cv.take<long long>(C);
  D.o_iter = cv.take<long long>(C);
  D.o_acc = cv.take<long long>(C);
  D.o_dt = cv.take<double>(C);
}

}  // namespace

extern "C" {

long long rfs_hmc_last_evals(rfs_ctx *ctx) { return ctx ? ctx->hmc_evals : 0; }
long long rfs_hmc_last_steps(rfs_ctx *ctx) { return ctx ? ctx->hmc_steps : 0; }

int rfs_set_hmc_options(rfs_ctx *ctx, long long resident_slots, double max_seconds) {
  if (!ctx) return RFS_E_ARG;
  if (resident_slots < 0 || max_seconds < 0.0) return fail(ctx, RFS_E_ARG, "bad HMC options");
  ctx->hmc_resident = resident_slots;
  ctx->hmc_max_seconds = max_seconds;
  return RFS_OK;
}

int rfs_hmc_run(rfs_ctx *ctx, int sampler, int which, long long C, const long long *chain_id,
                const double *bounds, double dt, int Lmin, int Lmax, int L0, double target_ratio,
                long long seed, int nsamples, int ndraws, long long max_iters, double *samples,
                double *misfit, double *syn, double *initmodel, long long *n_iter,
                long long *n_acc, double *dt_final, signed char *accept_seq,
                long long max_iter_log) {
  if (!ctx) return RFS_E_ARG;
  if (sampler != 0 && sampler != 1) return fail(ctx, RFS_E_ARG, "sampler must be 0 or 1");
  if (C <= 0 || !chain_id || !bounds || nsamples <= 0 || ndraws < 0)
    return fail(ctx, RFS_E_ARG, "bad HMC arguments");
  if (which < 0 || which > 2) return fail(ctx, RFS_E_ARG, "which must be 0,1,2");
  const bool use_swd = which != 1, use_rf = which != 2;
  if ((use_swd && !ctx->has_swd) || (use_rf && !ctx->has_rf) || !ctx->has_obs)
    return fail(ctx, RFS_E_CONFIG, "objective not configured (rfs_config_swd/rf/obs)");
  if (use_swd && use_rf && ctx->n_swd != ctx->n_rf)
    return fail(ctx, RFS_E_CONFIG, "SWD and RF layer counts differ");
  // NumPy's RandomState.seed raises for seeds outside [0, 2**32-1]; chain i is seeded seed + id_i
  for (long long c = 0; c < C; c++)
    if (chain_id[c] < 0 || seed < 0 || seed + chain_id[c] > 0xffffffffLL)
      return fail(ctx, RFS_E_ARG, "Seed must be between 0 and 2**32 - 1");
  if (sampler == 0 && (Lmin < 1 || Lmax < Lmin)) return fail(ctx, RFS_E_ARG, "bad Lrange");
  CK(cudaSetDevice(ctx->device));
  const int n = use_swd ? ctx->n_swd : ctx->n_rf, n2 = 2 * n;
  const int nd = (use_rf ? ctx->nt * (int)ctx->ray_ps.size() : 0) +
                 (use_swd ? ctx->plan.ndata * (int)ctx->modes.size() : 0);
  if ((int)ctx->dobs.size() != nd) return fail(ctx, RFS_E_CONFIG, "dobs length != ndata");
  cudaStream_t st = ctx->stream;
  HmcCfg cfg;
  cfg.n2 = n2;
  cfg.ndata = nd;
  cfg.sampler = sampler;
  cfg.Lmin = Lmin;
  cfg.Lmax = Lmax;
  cfg.L0 = L0;
  cfg.nsamples = nsamples;
  cfg.ndraws = ndraws;
  cfg.dt0 = dt;
  cfg.target = target_ratio;
  cfg.lambda = L0 * dt;        // hmcda.py:76
  cfg.mu = log(10 * dt);       // hmcda.py:300 (uses the configured dt)
  cfg.max_iters = max_iters;
  cfg.max_log = accept_seq ? max_iter_log : 0;

  // resident slots: all chains at once unless rfs_set_hmc_options bounds them (then finished chains
  // hand their slot to queued ones on the device)
  const long long R = (ctx->hmc_resident > 0 && ctx->hmc_resident < C) ? ctx->hmc_resident : C;
  HmcDev D;
  memset(&D, 0, sizeof(D));
  Carver probe{nullptr};
  carve_hmc(probe, D, R, C, n2, nd);
  int rc;
  if ((rc = ensure(ctx, ctx->h_state, probe.off + 256))) return rc;
  Carver cv{(char *)ctx->h_state.p};
  carve_hmc(cv, D, R, C, n2, nd);
  // evaluation outputs
  if ((rc = ensure(ctx, ctx->io_U, sizeof(double) * R))) return rc;
  if ((rc = ensure(ctx, ctx->io_grad, sizeof(double) * R * n2))) return rc;
  if ((rc = ensure(ctx, ctx->io_dsyn, sizeof(double) * R * nd))) return rc;
  if ((rc = ensure(ctx, ctx->io_flag, R))) return rc;
  D.Ue = (const double *)ctx->io_U.p;
  D.ge = (const double *)ctx->io_grad.p;
  D.de = (const double *)ctx->io_dsyn.p;
  D.fe = (const unsigned char *)ctx->io_flag.p;
  // misc inputs: bounds, chain ids
  if ((rc = ensure(ctx, ctx->h_misc, sizeof(double) * 2 * n2 + sizeof(long long) * C + 512)))
    return rc;
  double *d_bounds = (double *)ctx->h_misc.p;
  long long *d_ids = (long long *)((char *)ctx->h_misc.p + ((sizeof(double) * 2 * n2 + 255) & ~255));
  CK(cudaMemcpyAsync(d_bounds, bounds, sizeof(double) * 2 * n2, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_ids, chain_id, sizeof(long long) * C, cudaMemcpyHostToDevice, st));
  D.bounds = d_bounds;
  D.chain_id = d_ids;
  D.seed = seed;
  D.C_total = C;
  {
    const int qh[4] = {(int)R, 0, 0, 0};
    CK(cudaMemcpyAsync(D.queue_head, qh, sizeof(qh), cudaMemcpyHostToDevice, st));
  }
  // outputs
  size_t ob = 0;
  const size_t o_mis = ob; ob += misfit ? sizeof(double) * C * nsamples : 0; ob = (ob + 255) & ~(size_t)255;
  const size_t o_smp = ob; ob += samples ? sizeof(double) * C * nsamples * n2 : 0; ob = (ob + 255) & ~(size_t)255;
  const size_t o_syn = ob; ob += syn ? sizeof(double) * C * nsamples * nd : 0; ob = (ob + 255) & ~(size_t)255;
  const size_t o_ini = ob; ob += initmodel ? sizeof(double) * C * n2 : 0; ob = (ob + 255) & ~(size_t)255;
  const size_t o_log = ob; ob += accept_seq ? (size_t)C * max_iter_log : 0; ob = (ob + 255) & ~(size_t)255;
  if ((rc = ensure(ctx, ctx->h_out, ob + 256))) return rc;
  char *ob_base = (char *)ctx->h_out.p;
  D.misfit = misfit ? (double *)(ob_base + o_mis) : nullptr;
  D.samples = samples ? (double *)(ob_base + o_smp) : nullptr;
  D.syn = syn ? (double *)(ob_base + o_syn) : nullptr;
  D.initmodel = initmodel ? (double *)(ob_base + o_ini) : nullptr;
  D.alog = accept_seq ? (signed char *)(ob_base + o_log) : nullptr;
  if (D.misfit) CK(cudaMemsetAsync(D.misfit, 0, sizeof(double) * C * nsamples, st));
  if (D.samples) CK(cudaMemsetAsync(D.samples, 0, sizeof(double) * C * nsamples * n2, st));
  if (D.syn) CK(cudaMemsetAsync(D.syn, 0, sizeof(double) * C * nsamples * nd, st));
  if (D.alog) CK(cudaMemsetAsync(D.alog, 0xff, (size_t)C * max_iter_log, st));

  LAUNCH(hmc_init_kernel, gridFor(R, 64), 64, 0, st, D, cfg, R);
  int h_active = 1;
  long long steps = 0;
  long long Ba = R;      // rows of the evaluated batch
  bool packed = false;   // false: row == slot
  const int check_every = 4;  // one stream sync per 4 global steps: at most 3 wasted steps at the end
  const auto t_start = std::chrono::steady_clock::now();
  while (h_active > 0) {
    for (int s = 0; s < check_every; s++) {
      if (packed) LAUNCH(hmc_gather_kernel, gridFor(Ba * n2, 256), 256, 0, st, D, n2, Ba);
      rc = rfs_misfit_grad_dev(ctx, Ba, packed ? D.xg : D.xeval, which, (double *)ctx->io_U.p,
                               (double *)ctx->io_grad.p, (double *)ctx->io_dsyn.p,
                               (unsigned char *)ctx->io_flag.p, st);
      if (rc) return rc;
      if (s == check_every - 1) CK(cudaMemsetAsync(D.n_active, 0, 2 * sizeof(int), st));
      LAUNCH(hmc_advance_kernel, gridFor(R, 64), 64, 0, st, D, cfg, R);
      steps++;
    }
    CK(cudaMemcpyAsync(&h_active, D.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h_active > 0 && ctx->hmc_max_seconds > 0.0 &&
        std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() >
            ctx->hmc_max_seconds) {
      LAUNCH(hmc_abort_kernel, gridFor(R, 128), 128, 0, st, D, R);
      break;
    }
    // re-pack once at least 1/16 of the evaluated rows belong to slots that have run dry
    if (h_active > 0 && (long long)h_active <= Ba - std::max<long long>(1, Ba / 16)) {
      LAUNCH(hmc_compact_kernel, gridFor(R, 256), 256, 0, st, D, R);
      Ba = h_active;
      packed = true;
    }
  }
  // results
  std::vector<long long> nev(C);
  CK(cudaMemcpyAsync(nev.data(), D.nevals, sizeof(long long) * C, cudaMemcpyDeviceToHost, st));
  if (misfit) CK(cudaMemcpyAsync(misfit, D.misfit, sizeof(double) * C * nsamples, cudaMemcpyDeviceToHost, st));
  if (samples)
    CK(cudaMemcpyAsync(samples, D.samples, sizeof(double) * C * nsamples * n2, cudaMemcpyDeviceToHost, st));
  if (syn) CK(cudaMemcpyAsync(syn, D.syn, sizeof(double) * C * nsamples * nd, cudaMemcpyDeviceToHost, st));
  if (initmodel) CK(cudaMemcpyAsync(initmodel, D.initmodel, sizeof(double) * C * n2, cudaMemcpyDeviceToHost, st));
  if (accept_seq) CK(cudaMemcpyAsync(accept_seq, D.alog, (size_t)C * max_iter_log, cudaMemcpyDeviceToHost, st));
  if (n_iter) CK(cudaMemcpyAsync(n_iter, D.o_iter, sizeof(long long) * C, cudaMemcpyDeviceToHost, st));
  if (n_acc) CK(cudaMemcpyAsync(n_acc, D.o_acc, sizeof(long long) * C, cudaMemcpyDeviceToHost, st));
  if (dt_final) CK(cudaMemcpyAsync(dt_final, D.o_dt, sizeof(double) * C, cudaMemcpyDeviceToHost, st));
  std::vector<int> status(C);
  CK(cudaMemcpyAsync(status.data(), D.status, sizeof(int) * C, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  long long tot = 0;
  for (long long c = 0; c < C; c++) tot += nev[c];
  ctx->hmc_evals = tot;
  ctx->hmc_steps = steps;
  long long nbad2 = 0, nbad3 = 0, nbad4 = 0, nbad5 = 0;
  for (long long c = 0; c < C; c++) {
    if (status[c] == 2) nbad2++;
    if (status[c] == 3) nbad3++;
    if (status[c] == 4) nbad4++;
    if (status[c] == 5) nbad5++;
  }
  if (nbad2 || nbad3 || nbad4 || nbad5) {
    char msg[400];
    snprintf(msg, sizeof(msg),
             "%lld chain(s) stopped by max_iters, %lld chain(s) stuck at a state whose forward model "
             "fails (reference would loop forever), %lld chain(s) failed inside _find_initial_dt "
             "(reference: exit(1)), %lld chain(s) stopped by the wall-clock budget",
             nbad2, nbad3, nbad4, nbad5);
    ctx->err = msg;
    return 1;  // positive: completed with per-chain early stops (outputs are valid for the others)
  }
  return RFS_OK;
}

}  // extern "C"
