"""Generate tests/golden/*.npz from the CPU oracle (oracle/), run in the build container:

    python tests/golden/make_golden.py

The reference itself cannot be executed here (Fortran + FFTW3), so these vectors
pin the ORACLE (and through it the CUDA path) against regressions; the oracle in turn is pinned by
the analytic / finite-difference / identity tests in tests/test_oracle_*.py."""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.oracle import Oracle, brocher  # noqa: E402
from rfsurfhmc_b200.fixtures import f1_config, f1_true_model, driver_bounds, sorted_uniform_models, perturbed_models  # noqa: E402


def main():
    O = Oracle()
    x0 = f1_true_model()
    vs, thk = x0[:7], x0[7:]
    vp, rho = brocher(vs)
    T = np.arange(5., 41.)
    out = {"thk": thk, "vs": vs, "vp": vp, "rho": rho, "T": T}
    for wt in ("Rc", "Rg", "Lc", "Lg"):
        for mode in (0, 1, 2):
            c, ok = O.surf_forward(thk, vp, vs, rho, T, wt, mode=mode)
            out[f"fwd_{wt}_{mode}"] = c
            out[f"fwd_{wt}_{mode}_ok"] = np.array(ok)
        c, da, db, dr, dh, ok = O.surf_adjoint_kernel(thk, vp, vs, rho, T, wt)
        out[f"ker_{wt}_c"], out[f"ker_{wt}_da"], out[f"ker_{wt}_db"] = c, da, db
        out[f"ker_{wt}_dr"], out[f"ker_{wt}_dh"] = dr, dh
    q = thk * 0 + 9999.
    for rft in ("P", "S"):
        rf, kl = O.rf_kernel_all(thk, rho, vp, vs, q, q, 0.045, 125, 0.4, 1.5, 5.0, "freq", 0.001, rft)
        out[f"rf_{rft}"], out[f"rf_{rft}_kl"] = rf, kl
    # F2 "smoke-time" parameter set (reference test_forward.py:13-30), forward only
    vs2 = np.array([3.2, 3.4, 3.46, 3.7, 3.9, 4.5, 4.7])
    vp2, rho2 = brocher(vs2)
    out["f2_vs"] = vs2
    out["f2_rf_time"] = O.rf_forward(thk, rho2, vp2, vs2, q, q, 0.045, 500, 0.1, 1.0, 5.0, "time", 0.001, "P")
    out["f2_rf_freq"] = O.rf_forward(thk, rho2, vp2, vs2, q, q, 0.045, 500, 0.1, 1.0, 5.0, "freq", 0.001, "P")
    Tf = np.linspace(5, 40, 36)
    out["f2_Rc"], _ = O.surf_forward(thk, vp2, vs2, rho2, Tf, "Rc")
    out["f2_Rg"], _ = O.surf_forward(thk, vp2, vs2, rho2, Tf, "Rg")
    np.savez_compressed(os.path.join(HERE, "f1_dropin.npz"), **out)

    # fused joint misfit+gradient on two realistic model sets
    cfg = f1_config()
    _, _, d, _ = O.joint_batch(x0[None, :], np.zeros(197), cfg)
    dobs = d[0]
    Xa = sorted_uniform_models(driver_bounds(x0), 48, seed=11)
    Xb = perturbed_models(x0, 48, seed=12)
    X = np.vstack((Xa, Xb))
    U, g, ds, f = O.joint_batch(X, dobs, cfg, nthreads=8)
    np.savez_compressed(os.path.join(HERE, "f1_joint.npz"), X=X, dobs=dobs, U=U, grad=g, dsyn=ds, flag=f)
    print("golden written:", os.listdir(HERE))


if __name__ == "__main__":
    main()
