"""Device-resident HMC vs the NumPy restatement of the reference samplers (oracle/hmc_ref.py):
identical initial models, identical accept/reject sequences and (within the forward-model
tolerances) identical samples under the shared NumPy-legacy random stream."""
import os
import numpy as np
import pytest
from oracle import hmc_ref
from rfsurfhmc_b200.fixtures import f1_config, f1_true_model, driver_bounds

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _setup(ctx):
    cfg = f1_config()
    dobs = np.load(os.path.join(ROOT, "tests", "golden", "f1_joint.npz"))["dobs"]
    ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"])
    ctx.config_rf(7, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"],
                  cfg["rf_type"], cfg["method"])
    ctx.config_obs(dobs)
    return cfg, dobs, driver_bounds(f1_true_model())


def test_base_sampler_accept_sequence_parity(ctx, oracle):
    cfg, dobs, bounds = _setup(ctx)
    ids = [0, 1, 2, 3]
    niter = 25
    out = ctx.hmc_run(0, ids, bounds, 0.1, Lrange=(5, 20), seed=991206, nsamples=40, ndraws=5,
                      max_iters=niter, want_samples=True, want_syn=True, log_accepts=niter)
    f = hmc_ref.oracle_joint_f(oracle, dobs, cfg)
    for i, cid in enumerate(ids):
        R = hmc_ref.run_base(f, bounds, 0.1, (5, 20), 991206 + cid, nsamples=40, ndraws=5, max_iters=niter)
        assert np.allclose(out["initmodel"][i], R.initmodel, rtol=0, atol=1e-14)
        seq = out["accept_seq"][i][:out["n_iter"][i]]
        assert list(seq) == R.accepts, (cid, list(seq), R.accepts)
        assert out["n_acc"][i] == R.n_acc
        ns = max(0, R.n_acc - 5)
        if ns > 0:
            assert np.allclose(out["samples"][i][:ns], R.samples[:ns], rtol=1e-6, atol=1e-9)
            assert np.allclose(out["misfit"][i][:ns], R.misfit[:ns], rtol=1e-4)
    assert out["evals"] > 0


def test_dual_averaging_sampler_parity(ctx, oracle):
    cfg, dobs, bounds = _setup(ctx)
    ids = [0, 5]
    niter = 12
    out = ctx.hmc_run(1, ids, bounds, 0.02, L0=10, target_ratio=0.65, seed=991206, nsamples=20, ndraws=4,
                      max_iters=niter, want_samples=True, log_accepts=niter)
    f = hmc_ref.oracle_joint_f(oracle, dobs, cfg)
    for i, cid in enumerate(ids):
        R = hmc_ref.run_da(f, bounds, 0.02, 10, 0.65, 991206 + cid, nsamples=20, ndraws=4, max_iters=niter)
        seq = out["accept_seq"][i][:out["n_iter"][i]]
        assert list(seq) == R.accepts, (cid, list(seq), R.accepts)
        assert np.isclose(out["dt"][i], R.dt, rtol=1e-6)


def test_batch_compaction_leaves_every_chain_unchanged(ctx):
    """Chains draw different L and reject differently, so they finish at different global steps;
    the driver re-packs the evaluated batch as they finish.  A chain's samples must not depend on
    which other chains ran beside it: 48 chains together == the same chains in two smaller runs."""
    cfg, dobs, bounds = _setup(ctx)
    ids = np.arange(48)
    kw = dict(seed=991206, nsamples=6, ndraws=1, max_iters=14, want_samples=True, log_accepts=14)
    full = ctx.hmc_run(0, ids, bounds, 0.1, Lrange=(5, 20), **kw)
    assert len(set(full["n_iter"].tolist())) > 1  # chains really do finish at different times
    for part in (ids[:5], ids[29:48]):
        sub = ctx.hmc_run(0, part, bounds, 0.1, Lrange=(5, 20), **kw)
        for k in ("samples", "misfit", "n_iter", "n_acc", "accept_seq", "initmodel"):
            if k in full and full[k] is not None:
                assert np.array_equal(np.asarray(full[k])[part], np.asarray(sub[k]), equal_nan=True), k


def test_slot_refill_and_budget_leave_every_chain_unchanged(ctx):
    """Chain-level refill: with fewer resident slots than chains, a finished chain hands its slot to the
    next queued chain on the device.  Every chain's result must be the one it has when all chains are
    resident; a wall-clock budget stops the run with the unfinished chains flagged."""
    cfg, dobs, bounds = _setup(ctx)
    ids = np.arange(40)
    kw = dict(seed=991206, nsamples=5, ndraws=1, max_iters=12, want_samples=True, log_accepts=12)
    full = ctx.hmc_run(0, ids, bounds, 0.1, Lrange=(5, 20), **kw)
    try:
        for resident in (7, 16):
            ctx.set_hmc_options(resident=resident)
            part = ctx.hmc_run(0, ids, bounds, 0.1, Lrange=(5, 20), **kw)
            assert part["global_steps"] > full["global_steps"]        # the same work through fewer slots
            for k in ("samples", "misfit", "n_iter", "n_acc", "accept_seq", "initmodel", "dt"):
                assert np.array_equal(np.asarray(full[k]), np.asarray(part[k]), equal_nan=True), (resident, k)
            assert part["evals"] == full["evals"]
        da_full = ctx.hmc_run(1, ids[:12], bounds, 0.02, L0=10, target_ratio=0.65, seed=991206, nsamples=6,
                              ndraws=3, max_iters=10, want_samples=True, log_accepts=10)
        ctx.set_hmc_options(resident=5)
        da_part = ctx.hmc_run(1, ids[:12], bounds, 0.02, L0=10, target_ratio=0.65, seed=991206, nsamples=6,
                              ndraws=3, max_iters=10, want_samples=True, log_accepts=10)
        for k in ("samples", "misfit", "n_iter", "n_acc", "accept_seq", "dt"):
            assert np.array_equal(np.asarray(da_full[k]), np.asarray(da_part[k]), equal_nan=True), k
        # wall-clock budget: a run that cannot finish in time returns what it has
        ctx.set_hmc_options(resident=4, max_seconds=1e-3)
        cut = ctx.hmc_run(0, ids, bounds, 0.1, Lrange=(5, 20), seed=991206, nsamples=400, ndraws=100)
        assert "wall-clock budget" in cut["warning"] and (cut["n_acc"] < 500).all()
    finally:
        ctx.set_hmc_options(0, 0.0)


def test_sampler_front_ends_and_result_files(ctx, tmp_path):
    import yaml
    from rfsurfhmc_b200 import driver
    param = yaml.safe_load(open(os.path.join(ROOT, "tests", "golden", "f1_param.yaml")))
    param["hmc"].update(nsamples=6, ndraws=2, OUTPUT_DIR=str(tmp_path) + "/")
    misfit, n_iter, out = driver.run(param, "base", nchains=3)
    assert misfit.shape == (3, 6) and np.all(misfit > 0) and np.all(n_iter >= 8)
    assert os.path.exists(tmp_path / "misfit.npy") and os.path.exists(tmp_path / "real_syn.npy")
    z = np.load(tmp_path / "chain_joint.1.npz")
    assert z["models"].shape == (6, 14) and z["syn"].shape == (6, 197) and z["mean/model"].shape == (14,)
    assert np.allclose(z["obs"], np.load(tmp_path / "real_syn.npy"))


def test_device_samplers_reproduce_the_reference_drivers_own_output(ctx):
    """tests/golden/reference_code.npz holds real_syn.npy and misfit.npy written by the reference's
    main_base.py and main_DA.py, run unmodified (tests/golden/make_reference_golden.py).  The device
    samplers, given the same observations, bounds, seed and settings, must reproduce those misfit
    histories -- no restatement in between."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "reference_code.npz"))
    cfg = f1_config()
    ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"])
    ctx.config_rf(7, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"],
                  cfg["rf_type"], cfg["method"])
    ctx.config_obs(z["drvbase_real_syn"])
    b = z["drvbase_bounds"]
    ns, ndr, dt = z["drvbase_cfg"]
    out = ctx.hmc_run(0, [0], b, float(dt), Lrange=(5, 20), seed=991206, nsamples=int(ns), ndraws=int(ndr))
    assert out["n_acc"][0] == int(ns + ndr)
    assert np.allclose(out["misfit"][0], z["drvbase_misfit"][0], rtol=1e-4)
    ns, ndr, dt = z["drvda_cfg"]
    out = ctx.hmc_run(1, [0], b, float(dt), L0=10, target_ratio=0.65, seed=991206, nsamples=int(ns),
                      ndraws=int(ndr))
    assert out["n_acc"][0] == int(ns + ndr)
    assert np.allclose(out["misfit"][0], z["drvda_misfit"][0], rtol=1e-4)
