"""Bounded dual-averaging run (evidence for DESIGN.md): step-size adaptation and acceptance."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfsurfhmc_b200._lib import Context
from rfsurfhmc_b200.fixtures import *
from bench import make_dobs_gpu, workload
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg, x0, _ = workload(1, 0)
ctx = Context(0)
ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"]); ctx.config_rf(7, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"], cfg["rf_type"], cfg["method"]); dobs = make_dobs_gpu(ctx, cfg, x0); ctx.config_obs(dobs)
b = driver_bounds(x0)
t0 = time.time()
out = ctx.hmc_run(1, np.arange(nch), b, dt=0.02, Lrange=(1, 40), L0=10, target_ratio=0.65, seed=991206, nsamples=200, ndraws=100,
                  max_iters=500, want_samples=True, log_accepts=500)
dt = time.time() - t0
seq = out["accept_seq"].astype(float); seq[seq < 0] = np.nan
late = np.nanmean(seq[:, 150:], axis=1)
fin = out["n_acc"] >= 300
print(json.dumps({"sampler": "HMCDualAveraging dt=0.02 L0=10 target=0.65, L capped at 40 (extension)", "chains": nch, "seconds": round(dt, 1),
                  "finished": int(fin.sum()), "evals": out["evals"], "evals_per_s": round(out["evals"] / dt),
                  "accept_ratio_after_warmup(median,p10,p90)": [round(float(v), 3) for v in np.nanpercentile(late, [50, 10, 90])],
                  "dt_final(p1,p50,p99)": [round(float(v), 4) for v in np.percentile(out["dt"], [1, 50, 99])],
                  "misfit_median_first_last": [round(float(np.median(out["misfit"][fin, 0])), 4),
                                               round(float(np.median(out["misfit"][fin, -1])), 4)] if fin.any() else None,
                  "warning": out["warning"]}), flush=True)
