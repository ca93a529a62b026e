"""The warp-cooperative (team) root search must return the SAME BITS as the thread-mapped kernel: it
runs the reference's scan / Neville state machine (surfdisp96.f:398-491,568-701) with the same secular
function, only scheduled differently (csrc/swd_roots_team.cuh).  Every mapping is compared with the
thread-mapped kernel through the C ABI on the drop-in and on the fused path."""
import numpy as np
import pytest
from oracle.oracle import brocher
from rfsurfhmc_b200.fixtures import f1_config, f1_true_model, driver_bounds, sorted_uniform_models

pytestmark = pytest.mark.gpu
TEAMS = [(2, 2), (4, 1), (4, 4), (8, 1), (8, 2), (8, 8), (16, 1), (16, 2), (32, 1), (32, 2), (32, 4)]


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.fixture()
def tctx(ctx):
    yield ctx
    ctx.set_roots_team(-1)


def _wild(B, seed):
    """strongly inverted random models: the state machine takes its rare paths here (downward scans,
    clow resets, failed modes, Neville fall-backs)"""
    x0 = f1_true_model()
    return np.random.default_rng(seed).uniform(0.5, 1.5, (B, 14)) * x0 + 0.01


def test_team_roots_bit_identical_on_65k_models(tctx):
    """>= 65 536 models at the C1/C4 sizes: fused joint objective, every output bit for bit, and the
    same number of secular evaluations consumed (speculated grid points that are not consumed do not
    count)."""
    ctx = tctx
    cfg, x0 = f1_config(), f1_true_model()
    B = 65536
    X = sorted_uniform_models(driver_bounds(x0), B, seed=2024)
    X[B // 2:] = _wild(B - B // 2, 5)
    ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"])
    ctx.config_rf(7, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"],
                  cfg["rf_type"], cfg["method"])
    ctx.config_obs(np.full(125 + 72, 3.0))
    ctx.set_roots_team(0)
    ctx.count_evals(True)
    ref = ctx.misfit_grad_host(X)
    nev0 = ctx.read_evals()
    assert ctx.last_roots_team() == (0, 1)
    for T, S in [(8, 1), (32, 4)]:
        ctx.set_roots_team(T, S)
        ctx.count_evals(True)
        got = ctx.misfit_grad_host(X)
        nev = ctx.read_evals()
        assert ctx.last_roots_team() == (T, S)
        for a, b in zip(ref, got):
            assert _same(a, b), (T, S)
        assert nev == nev0, (T, S, nev, nev0)
    ctx.count_evals(False)


@pytest.mark.parametrize("T,S", TEAMS)
def test_every_team_shape_all_wave_types_and_modes(tctx, T, S):
    """Rayleigh + Love, phase + group, modes 0-2 (mode chaining through cwork), realistic and wild
    models, n = 7: roots, group velocities and kernels identical to the thread mapping."""
    ctx = tctx
    x0 = f1_true_model()
    B = 192
    X = sorted_uniform_models(driver_bounds(x0), B, seed=11)
    X[B // 2:] = _wild(B - B // 2, 12)
    vs, thk = X[:, :7], X[:, 7:]
    vp, rho = brocher(vs)
    Tp = np.arange(5., 41., 2.5)
    for wt in ("Rc", "Lc", "Rg", "Lg"):
        for mode in (0, 2):
            ctx.set_roots_team(0)
            ref = ctx.surf_adjoint_kernel(thk, vp, vs, rho, Tp, wt, mode=mode)
            ctx.set_roots_team(T, S)
            got = ctx.surf_adjoint_kernel(thk, vp, vs, rho, Tp, wt, mode=mode)
            assert ctx.last_roots_team() == (T, S)
            for a, b in zip(ref, got):
                assert _same(a, b), (wt, mode)
    # all modes at once (BASELINE config 2 entry point)
    ctx.set_roots_team(0)
    ref = ctx.surf_adjoint_kernel(thk, vp, vs, rho, Tp, "Rc", mode=2, all_modes=True)
    ctx.set_roots_team(T, S)
    got = ctx.surf_adjoint_kernel(thk, vp, vs, rho, Tp, "Rc", mode=2, all_modes=True)
    for a, b in zip(ref, got):
        assert _same(a, b)


@pytest.mark.parametrize("T,S", [(2, 2), (4, 4), (8, 1), (16, 2), (32, 1), (32, 4)])
def test_team_roots_many_layers_water_and_sphere(tctx, T, S):
    """n = 40 (several build rounds per evaluation), n = 200, a water layer on top (llw = 2) and the
    earth-flattening model blocks."""
    ctx = tctx
    rng = np.random.default_rng(3)
    for n, B, nT in ((40, 48, 20), (200, 6, 12)):
        i = np.arange(n - 1)
        thk = np.hstack((20.0 / n + 4.0 / n * i / n, [0.0]))[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (B, n)))
        thk[:, -1] = 0.0
        vs0 = 2.0 + 2.7 * (np.arange(n) / (n - 1.0))**0.7
        vs = np.clip(vs0[None, :] * (1 + 0.04 * rng.standard_normal((B, n))), 1.5, 5.0)
        vp, rho = brocher(vs)
        Tp = np.geomspace(2, 100, nT)
        for wt, mode, sph in (("Rc", 1, False), ("Lg", 0, False), ("Rg", 0, True)):
            ctx.set_roots_team(0)
            ref = ctx.surf_adjoint_kernel(thk, vp, vs, rho, Tp, wt, mode=mode, sphere=sph)
            ctx.set_roots_team(T, S)
            got = ctx.surf_adjoint_kernel(thk, vp, vs, rho, Tp, wt, mode=mode, sphere=sph)
            for a, b in zip(ref, got):
                assert _same(a, b), (n, wt, mode, sph)
    # ocean model: water layer on top
    B = 40
    thk = np.tile([2.5, 4.0, 8.0, 12.0, 0.0], (B, 1)) * (1 + 0.1 * rng.uniform(-1, 1, (B, 5)))
    thk[:, -1] = 0
    vs = np.tile([0.0, 2.4, 3.3, 3.8, 4.5], (B, 1)) * (1 + 0.05 * rng.uniform(-1, 1, (B, 5)))
    vp = np.tile([1.5, 4.4, 5.9, 6.7, 8.0], (B, 1)) * (1 + 0.05 * rng.uniform(-1, 1, (B, 5)))
    rho = np.tile([1.03, 2.4, 2.7, 2.9, 3.3], (B, 1))
    Tp = np.geomspace(4, 60, 14)
    for wt in ("Rc", "Rg", "Lc"):
        ctx.set_roots_team(0)
        ref = ctx.surf_adjoint_kernel(thk, vp, vs, rho, Tp, wt)
        ctx.set_roots_team(T, S)
        got = ctx.surf_adjoint_kernel(thk, vp, vs, rho, Tp, wt)
        for a, b in zip(ref, got):
            assert _same(a, b), wt


def test_mapping_is_chosen_by_batch_size(tctx):
    """Small batches go to the team kernel, full batches to the thread-mapped one; the choice never
    changes a result bit (chunk- and mapping-independence of every model's answer)."""
    ctx = tctx
    cfg, x0 = f1_config(), f1_true_model()
    ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"])
    ctx.config_rf(7, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"],
                  cfg["rf_type"], cfg["method"])
    ctx.config_obs(np.full(125 + 72, 3.0))
    X = sorted_uniform_models(driver_bounds(x0), 16384, seed=9)
    ctx.set_roots_team(-1)
    big = ctx.misfit_grad_host(X)
    assert ctx.last_roots_team()[0] == 0
    small = ctx.misfit_grad_host(X[:64])
    assert ctx.last_roots_team()[0] > 0
    for a, b in zip(big, small):
        assert _same(a[:64], b)


def test_length_sorted_schedule_is_bit_identical(tctx):
    """Large batches run the thread-mapped search in length-sorted job order (swd_sched_*_kernel in
    csrc/swd_roots_tu.cu): which lane solves which (model, sequence) changes, nothing else.  Every output
    bit and the evaluation count must equal the unsorted run — ragged batch sizes, wild models (failed
    modes, retries), Love + Rayleigh in one objective, several modes, spherical earth, n > 8 (model not
    staged in shared memory) included."""
    ctx = tctx
    cfg, x0 = f1_config(), f1_true_model()
    Tl = np.asarray(cfg["tRc"])[:11]
    cases = [(40000 + 13, 0, False, None), (9001, [0, 1], False, Tl), (8192 + 5, 0, True, Tl)]
    try:
        for B, modes, sphere, tL in cases:
            X = sorted_uniform_models(driver_bounds(x0), B, seed=77)
            X[B // 2:] = _wild(B - B // 2, 9)
            ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"], tLc=tL, mode=modes, sphere=sphere)
            ctx.config_obs(np.full(ctx.n_swd_data, 3.0))
            ctx.set_roots_team(0)
            out = []
            for mode in (0, 1):
                ctx.set_roots_sched(mode)
                ctx.count_evals(True)
                res = ctx.misfit_grad_host(X, which=2)
                out.append((res, ctx.read_evals()))
                assert ctx.last_roots_sched() == bool(mode)
            names = ("U", "grad", "dsyn", "flag")
            for nm, a, b in zip(names, out[0][0], out[1][0]):
                assert _same(a, b), (B, modes, sphere, nm, int(np.sum(~np.isclose(a, b, rtol=0, atol=0, equal_nan=True))))
            assert out[0][1] == out[1][1]
        # n = 40 (global-memory model accessor), four wave types
        Tp = np.geomspace(2, 100, 20)
        rng = np.random.default_rng(3)
        B, n = 3000, 40
        i = np.arange(n - 1)
        thk = np.hstack((0.5 + 0.75 * i / n, [0.0]))[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (B, n)))
        thk[:, -1] = 0.0
        vs = np.clip((2.0 + 2.7 * (np.arange(n) / (n - 1.0))**0.7)[None, :] * (1 + 0.04 * rng.standard_normal((B, n))), 1.5, 5.0)
        X = np.hstack((vs, thk))
        ctx.config_swd(n, Tp, Tp, Tp, Tp, mode=0)
        ctx.config_obs(np.full(80, 3.0))
        ctx.set_roots_team(0)
        res = []
        for mode in (0, 1):
            ctx.set_roots_sched(mode)
            res.append(ctx.misfit_grad_host(X, which=2))
        for a, b in zip(*res):
            assert _same(a, b)
        # automatic: on for a large thread-mapped batch, off for a small one
        ctx.set_roots_sched(-1)
        ctx.set_roots_team(-1)
        ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"])
        ctx.config_obs(np.full(72, 3.0))
        ctx.misfit_grad_host(sorted_uniform_models(driver_bounds(x0), 16384, seed=3), which=2)
        assert ctx.last_roots_sched() and ctx.last_roots_team() == (0, 1)
        ctx.misfit_grad_host(sorted_uniform_models(driver_bounds(x0), 512, seed=3), which=2)
        assert not ctx.last_roots_sched()
    finally:
        ctx.count_evals(False)
        ctx.set_roots_sched(-1)
