"""GPU-vs-oracle error DISTRIBUTION at the C1/C4 sizes (run on a GPU box):

    python tests/gpu_parity_report.py [--n 16384] [--out gpurun_out/parity_report.json]

Two model sets of `n` models each:
  * "sampler":  the distribution chains start from (HamitonianMC.set_initial_model, pyhmc/hmc.py:74-99);
  * "trajectory": states inside leapfrog trajectories — accepted states of the device sampler after a few
    trajectories, moved one leapfrog drift x + dt p (p ~ 0.5 N(0,1), dt = 0.1) and mirrored into the box.
For every set the fused joint objective of the CUDA path is compared with the oracle (checker build: -O2,
no FMA contraction), and — the yard-stick for what "identical" can mean for this algorithm — the oracle's
own -O3/FMA build is compared with the checker build.  Per quantity: fraction of models within the
north-star tolerance and the maximum error.  The committed copy is profiles/r02_parity_report.json."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from rfsurfhmc_b200._lib import Context  # noqa: E402
from rfsurfhmc_b200.fixtures import f1_config, f1_true_model, driver_bounds, sorted_uniform_models  # noqa: E402

TOL_C, TOL_RF, TOL_G = 1e-6, 1e-5, 1e-4


def compare(a, b, nt):
    """a, b = (U, grad, dsyn, flag): error statistics of a against b, per model"""
    Ua, ga, da, fa = a
    Ub, gb, db, fb = b
    out = {"models": int(len(Ua)), "flags_equal": bool(np.array_equal(fa, fb)),
           "flag_mismatches": int((fa != fb).sum()), "failed_models": int((~fb).sum())}
    m = fa & fb
    rf_a, rf_b = da[m, :nt], db[m, :nt]
    e_rf = np.abs(rf_a - rf_b).max(axis=1) / np.abs(rf_b).max(axis=1)
    sw_a, sw_b = da[m, nt:], db[m, nt:]
    e_c = (np.abs(sw_a - sw_b) / np.abs(sw_b))
    e_cph, e_cgr = e_c[:, :36].max(axis=1), e_c[:, 36:].max(axis=1)
    e_u = np.abs(Ua[m] - Ub[m]) / np.maximum(np.abs(Ub[m]), 1e-300)
    fin = np.isfinite(ga[m]).all(axis=1) & np.isfinite(gb[m]).all(axis=1)
    e_g = np.full(m.sum(), np.nan)
    e_g[fin] = np.abs(ga[m][fin] - gb[m][fin]).max(axis=1) / np.abs(gb[m][fin]).max(axis=1)

    def stat(e, tol):
        e = e[np.isfinite(e)]
        return {"tolerance": tol, "fraction_within": float(np.mean(e <= tol)), "max": float(e.max()),
                "median": float(np.median(e)), "p99": float(np.quantile(e, 0.99)),
                "p999": float(np.quantile(e, 0.999)), "count_outside": int((e > tol).sum())}
    out["phase_velocity_rel"] = stat(e_cph, TOL_C)
    out["group_velocity_rel"] = stat(e_cgr, TOL_C)
    out["rf_over_peak"] = stat(e_rf, TOL_RF)
    out["misfit_rel"] = stat(e_u, 1e-4)
    out["gradient_over_max_component"] = stat(e_g, TOL_G)
    out["gradient_nonfinite_either_side"] = int((~fin).sum())
    out["phase_roots_bit_identical_fraction"] = float(np.mean((sw_a[:, :36] == sw_b[:, :36]).all(axis=1)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16384)
    ap.add_argument("--out", default="gpurun_out/parity_report.json")
    a = ap.parse_args()
    nth = os.cpu_count() or 8
    cfg, x0 = f1_config(), f1_true_model()
    bounds = driver_bounds(x0)
    O, Of = Oracle(), Oracle(fast=True)
    nt = cfg["nt"]
    _, _, d0, _ = O.joint_batch(x0[None, :], np.zeros(nt + 72), cfg)
    dobs = d0[0]
    ctx = Context(0)
    ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"])
    ctx.config_rf(7, cfg["ray_p"], nt, cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"], cfg["rf_type"],
                  cfg["method"])
    ctx.config_obs(dobs)
    sets = {"sampler": sorted_uniform_models(bounds, a.n, seed=20261017)}
    # mid-trajectory states
    ho = ctx.hmc_run(0, np.arange(a.n), bounds, 0.1, Lrange=(5, 20), seed=991206, nsamples=1, ndraws=3,
                     max_iters=12, want_samples=True)
    xs = ho["samples"][:, 0, :]
    ok = np.isfinite(xs).all(axis=1) & (ho["n_acc"] >= 4)
    xs = xs[ok]
    rng = np.random.default_rng(7)
    xm = xs + 0.1 * 0.5 * rng.standard_normal(xs.shape)
    lo, hi = bounds[:, 0], bounds[:, 1]
    for _ in range(4):
        xm = np.where(xm > hi, 2 * hi - xm, xm)
        xm = np.where(xm < lo, 2 * lo - xm, xm)
    sets["trajectory"] = xm
    rep = {"workload": "C1/C4 joint objective (n=7, 36 Rc + 36 Rg, RF nt=125), tolerances of BASELINE.json: c,U 1e-6 "
                       "relative, RF 1e-5 of the peak, gradient 1e-4 of its largest component",
           "oracle_checker": "g++ -O2 -ffp-contract=off", "oracle_fma": "g++ -O3 -march=x86-64-v3 (FMA contraction)",
           "sets": {}}
    for name, X in sets.items():
        t0 = time.time()
        ref = O.joint_batch(X, dobs, cfg, nthreads=nth)
        fma = Of.joint_batch(X, dobs, cfg, nthreads=nth)
        t1 = time.time()
        gpu = ctx.misfit_grad_host(X)
        rep["sets"][name] = {"gpu_vs_oracle": compare(gpu, ref, nt),
                             "oracle_fma_vs_oracle": compare(fma, ref, nt),
                             "gpu_vs_oracle_fma": compare(gpu, fma, nt),
                             "oracle_seconds": round(t1 - t0, 1), "root_search_mapping": list(ctx.last_roots_team())}
        print(name, json.dumps(rep["sets"][name]["gpu_vs_oracle"]["gradient_over_max_component"]), flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(rep, open(a.out, "w"), indent=1)
    print("written", a.out)


if __name__ == "__main__":
    main()
