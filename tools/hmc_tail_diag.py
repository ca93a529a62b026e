"""Bounded HMC run to look at the distribution of per-chain trajectory counts (stuck chains)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfsurfhmc_b200._lib import Context
from rfsurfhmc_b200.fixtures import *
from bench import make_dobs_gpu, workload
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
max_iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
cfg, x0, _ = workload(1, 0)
ctx = Context(0)
ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"]); ctx.config_rf(7, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"], cfg["rf_type"], cfg["method"]); dobs = make_dobs_gpu(ctx, cfg, x0); ctx.config_obs(dobs)
b = driver_bounds(x0)
for sampler, kw in ((0, dict(dt=0.1, Lrange=(5, 20))), (1, dict(dt=0.1, L0=10, target_ratio=0.65))):
    t0 = time.time()
    out = ctx.hmc_run(sampler, np.arange(nch), b, seed=991206, nsamples=800, ndraws=200, max_iters=max_iters,
                      want_samples=True, **kw)
    dt = time.time() - t0
    fin = out["n_acc"] >= 1000
    S = out["samples"]; mis = out["misfit"]
    best = np.array([S[c][np.argsort(mis[c])[:10]].mean(0) for c in np.where(fin)[0]])
    print(json.dumps({"sampler": sampler, "chains": nch, "seconds": round(dt, 1), "finished": int(fin.sum()),
                      "n_iter_pcts(50,90,99,100)": [int(v) for v in np.percentile(out["n_iter"], [50, 90, 99, 100])],
                      "n_acc_min": int(out["n_acc"].min()), "evals": out["evals"],
                      "evals_per_s": round(out["evals"] / dt), "accepted_per_s": round(float(out["n_acc"].sum()) / dt),
                      "dt_pcts(1,50,99)": [round(float(v), 4) for v in np.percentile(out["dt"], [1, 50, 99])],
                      "misfit_median_first_last": [round(float(np.median(mis[fin, 0])), 4), round(float(np.median(mis[fin, -1])), 4)] if fin.any() else None,
                      "best10_mean_abs_vs_err": [round(float(v), 3) for v in np.abs(best[:, :7] - x0[:7]).mean(0)] if fin.any() else None,
                      "warning": out["warning"]}), flush=True)
