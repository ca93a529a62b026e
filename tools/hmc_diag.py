"""Diagnose evaluation cost on mid-chain states vs initial models (run on GPU)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rfsurfhmc_b200._lib import Context
from rfsurfhmc_b200.fixtures import *
from bench import make_dobs_gpu, workload
cfg, x0, X0 = workload(16384, 5)
ctx = Context(0)
ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"]); ctx.config_rf(7, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"], cfg["rf_type"], cfg["method"]); dobs = make_dobs_gpu(ctx, cfg, x0); ctx.config_obs(dobs)
def timeit(X, label):
    ctx.misfit_grad_host(X)
    ctx.count_evals(True)
    t0 = time.perf_counter(); U, g, d, f = ctx.misfit_grad_host(X); t1 = time.perf_counter()
    ne, mx, heavy = ctx.read_eval_stats(); ctx.count_evals(False)
    print(f"{label}: {1e3*(t1-t0):.1f} ms  evals/model {ne/len(X):.0f}  slowest thread {mx}  threads>2000: {heavy}  fail {np.sum(~f)}  nan-grad {np.sum(np.isnan(g).any(1))}")
    return U, g, d, f
timeit(X0, "initial models")
b = driver_bounds(x0)
for dt, L in ((0.02, 20), (0.1, 12)):
    out = ctx.hmc_run(0, np.arange(16384), b, dt, Lrange=(L, L), seed=991206, nsamples=3, ndraws=0, max_iters=3, want_samples=True)
    acc = out["n_acc"] >= 1
    Xs = out["samples"][acc, 0]
    print("dt", dt, "L", L, "accepted chains", acc.sum(), "acc ratio", out["n_acc"].sum() / out["n_iter"].sum())
    timeit(Xs[:16384], f"after 1 accepted trajectory dt={dt}")
    rng = np.random.default_rng(0)
# random mid-trajectory-like states: initial + random walk of scale 0.3 reflected
Xw = X0 + 0.3 * np.random.default_rng(1).standard_normal(X0.shape)
Xw = np.clip(Xw, b[:, 0], b[:, 1])
timeit(Xw, "random-walk states")

# HMC initial models
out = ctx.hmc_run(0, np.arange(16384), b, 0.02, Lrange=(20, 20), seed=991206, nsamples=1, ndraws=0, max_iters=1, want_samples=True)
U, g, d, f = timeit(out["initmodel"], "hmc initial models")
