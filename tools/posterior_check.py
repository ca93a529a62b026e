"""Run both samplers with the reference's default param set (F1) on many chains and summarise
the posterior (evidence for DESIGN.md): accept ratio, misfit decay, recovered model vs truth."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, yaml
from rfsurfhmc_b200 import driver
from rfsurfhmc_b200.fixtures import f1_true_model

param = yaml.safe_load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "f1_param.yaml")))
param["hmc"]["OUTPUT_DIR"] = "/tmp/rfs_results/"
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x0 = f1_true_model()
for sampler in ("base", "da"):
    t0 = time.time()
    misfit, n_iter, out = driver.run(param, sampler, nchains=nch, save_chains=False)
    dt = time.time() - t0
    S = out["samples"]                       # [C, nsamples, 14]
    best = np.array([S[c][np.argsort(misfit[c])[:10]].mean(0) for c in range(nch)])
    err_vs = np.abs(best[:, :7] - x0[:7]).mean(0)
    print(json.dumps({"sampler": sampler, "chains": nch, "seconds": round(dt, 1),
                      "accepted_samples_per_s": round(nch * 1000 / dt, 1),
                      "evals": out["evals"], "evals_per_s": round(out["evals"] / dt, 1),
                      "accept_ratio_mean": round(float((1000 / n_iter).mean()), 3),
                      "final_dt_median": round(float(np.median(out["dt"])), 4),
                      "misfit_median_first_last": [float(np.median(misfit[:, 0])), float(np.median(misfit[:, -1]))],
                      "best10_mean_abs_vs_error_per_layer": [round(float(v), 3) for v in err_vs],
                      "status_warning": out["warning"]}))
