"""A full-length run of the reference's default sampler settings (pyhmc/hmc.py defaults of
param.yaml: dt=0.1, L in [5,20], ndraws=200, nsamples=800) for many chains at once, bounded by
max_iters.  Prints one JSON line: wall time, trajectories/s, accepted samples/s, evaluations/s,
distribution of per-chain trajectory counts and the posterior summary of the finished chains.

    python tools/hmc_full_run.py [chains=16384] [max_iters=1100]"""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfsurfhmc_b200._lib import Context
from rfsurfhmc_b200.fixtures import driver_bounds
from bench import make_dobs_gpu, workload

nch = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
max_iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1100
cfg, x0, _ = workload(1, 0)
ctx = Context(0)
ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"])
ctx.config_rf(7, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"],
              cfg["rf_type"], cfg["method"])
dobs = make_dobs_gpu(ctx, cfg, x0)
ctx.config_obs(dobs)
b = driver_bounds(x0)
t0 = time.time()
out = ctx.hmc_run(0, np.arange(nch), b, 0.1, Lrange=(5, 20), seed=991206, nsamples=800, ndraws=200,
                  max_iters=max_iters, want_samples=True)
dt = time.time() - t0
fin = out["n_acc"] >= 1000
S, mis = out["samples"], out["misfit"]
idx = np.where(fin)[0]
best = np.array([S[c][np.argsort(mis[c])[:10]].mean(0) for c in idx[:4096]])
post = S[idx[:4096], 400:, :7].reshape(-1, 7)
print(json.dumps({
    "sampler": "HamitonianMC dt=0.1 L in [5,20] ndraws=200 nsamples=800", "chains": nch, "max_iters": max_iters,
    "seconds": round(dt, 1), "finished_chains": int(fin.sum()),
    "trajectories_per_s": round(float(out["n_iter"].sum()) / dt), "accepted_samples_per_s": round(float(out["n_acc"].sum()) / dt),
    "evals": out["evals"], "evals_per_s": round(out["evals"] / dt),
    "n_iter_pcts(50,90,99,100)": [int(v) for v in np.percentile(out["n_iter"], [50, 90, 99, 100])],
    "misfit_median_first_last": [round(float(np.median(mis[fin, 0])), 4), round(float(np.median(mis[fin, -1])), 4)],
    "true_vs": [round(float(v), 2) for v in x0[:7]],
    "posterior_mean_vs(second half of the samples)": [round(float(v), 3) for v in post.mean(0)],
    "posterior_std_vs": [round(float(v), 3) for v in post.std(0)],
    "best10_mean_abs_vs_err": [round(float(v), 3) for v in np.abs(best[:, :7] - x0[:7]).mean(0)],
    "warning": out["warning"]}), flush=True)
