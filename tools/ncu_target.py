"""small targets for ncu captures: python tools/ncu_target.py <what>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rfsurfhmc_b200._lib import Context
from rfsurfhmc_b200.fixtures import f1_config, f1_true_model, driver_bounds, sorted_uniform_models
what = sys.argv[1] if len(sys.argv) > 1 else "team64"
cfg, x0 = f1_config(), f1_true_model()
ctx = Context(0)
ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"])
ctx.config_rf(7, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"], cfg["rf_type"], cfg["method"])
ctx.config_obs(np.full(197, 3.0))
B = {"team64": 64, "team2048": 2048, "thread16k": 16384}[what]
X = sorted_uniform_models(driver_bounds(x0), B, seed=5)
for _ in range(2):
    ctx.misfit_grad_host(X)
print("mapping", ctx.last_roots_team())
