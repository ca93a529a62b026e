"""Pins of the RF oracle: FFT vs NumPy, analytic half-space answer, finite differences, invariance."""
import numpy as np
import pytest
from independent import surface_response
from oracle.oracle import brocher
from rfsurfhmc_b200.fixtures import f1_true_model

X0 = f1_true_model()
VS, THK = X0[:7], X0[7:]
VP, RHO = brocher(VS)
Q = THK * 0 + 9999.
ARGS = dict(ray_p=0.045, nt=125, dt=0.4, gauss=1.5, time_shift=5., method="freq", water=0.001, rf_type="P")


@pytest.mark.parametrize("n", [2, 8, 128, 1024])
def test_fft_matches_numpy(oracle, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    X = oracle.rfft(x)
    assert np.allclose(X, np.fft.rfft(x), rtol=1e-12, atol=1e-12)
    Y = rng.standard_normal(n // 2 + 1) + 1j * rng.standard_normal(n // 2 + 1)
    # c2r ignores the imaginary parts of the DC and Nyquist bins (FFTW convention)
    assert np.allclose(oracle.irfft(Y, n), np.fft.irfft(Y, n), rtol=1e-12, atol=1e-12)


def test_halfspace_rf_is_single_gaussian_pulse(oracle):
    # no interface -> no conversions: the radial/vertical ratio is constant and the RF is one
    # Gaussian pulse at t = 0 (sample index time_shift/dt)
    n = 3
    vs = np.full(n, 3.6); thk = np.array([10., 10., 0.])
    vp, rho = brocher(vs)
    q = thk * 0 + 9999.
    rf = oracle.rf_forward(thk, rho, vp, vs, q, q, 0.06, 256, 0.1, 2.5, 5.0, "freq", 0.001, "P")
    t = np.arange(256) * 0.1 - 5.0
    k = np.argmax(np.abs(rf))
    assert abs(t[k]) < 0.05
    # time-domain image of exp(-(w/2a)^2) is exp(-a^2 t^2); the complex-frequency trick
    # (response at w - i sigma, trace times exp(sigma t), RFModule.f90:381-407) turns it into
    # exp(-a^2 t^2 + sigma t) with sigma = 4/(nft dt)
    sigma = 4.0 / (256 * 0.1)
    g = np.exp(-(2.5 * t)**2 + sigma * t)
    shape = rf / rf[k]
    assert np.max(np.abs(shape - g / g[k])[np.abs(t) < 1.5]) < 2e-3
    assert np.max(np.abs(rf[t > 2.0])) < 1e-3 * abs(rf[k])


def test_f1_trace_matches_independent_probe(oracle):
    rf = oracle.rf_forward(THK, RHO, VP, VS, Q, Q, **ARGS)
    # SURVEY.md Appendix C (independent NumPy scratch restatement)
    assert abs(rf.sum() - 1.095717149) < 1e-8
    assert np.argmax(rf) == 12 and abs(rf[12] - 0.222843) < 1e-6
    assert np.allclose(rf[10:16], [0.024547, 0.106417, 0.222843, 0.218859, 0.090575, 0.018077], atol=1e-6)


def test_kernels_finite_difference_and_variants(oracle):
    rf, kl = oracle.rf_kernel_all(THK, RHO, VP, VS, Q, Q, **ARGS)
    assert np.allclose(rf, oracle.rf_forward(THK, RHO, VP, VS, Q, Q, **ARGS), rtol=0, atol=1e-13)
    names = ["rho", "vp", "vs", "h"]
    arrs = {"rho": RHO, "vp": VP, "vs": VS, "h": THK}
    for ip, nm in enumerate(names):
        _, k1 = oracle.rf_kernel(THK, RHO, VP, VS, Q, Q, par_type=nm, **ARGS)
        assert np.array_equal(k1, kl[ip])
        for j in range(7):
            base = {k: v.copy() for k, v in arrs.items()}
            h = 1e-5
            base[nm][j] += h
            fp = oracle.rf_forward(base["h"], base["rho"], base["vp"], base["vs"], Q, Q, **ARGS)
            base[nm][j] -= 2 * h
            fm = oracle.rf_forward(base["h"], base["rho"], base["vp"], base["vs"], Q, Q, **ARGS)
            fd = (fp - fm) / (2 * h)
            assert np.max(np.abs(fd - kl[ip, j])) < 2e-8 + 1e-6 * np.max(np.abs(kl[ip, j])), (nm, j)
    # density invariance: sum_j rho_j dRF/drho_j = 0
    assert np.max(np.abs((RHO[:, None] * kl[0]).sum(0))) < 1e-13
    # half-space thickness has no influence
    assert np.all(kl[3, 6] == 0.0)


def test_s_type_and_time_method_run(oracle):
    a = dict(ARGS); a["rf_type"] = "S"
    rf, kl = oracle.rf_kernel_all(THK, RHO, VP, VS, Q, Q, **a)
    assert np.all(np.isfinite(rf)) and np.all(np.isfinite(kl))
    a = dict(ARGS); a["method"] = "time"; a["gauss"] = 1.0
    rft = oracle.rf_forward(THK, RHO, VP, VS, Q, Q, **a)
    a["method"] = "freq"
    rff = oracle.rf_forward(THK, RHO, VP, VS, Q, Q, **a)
    # iterative and water-level deconvolution agree on the main features
    assert abs(int(np.argmax(rft)) - int(np.argmax(rff))) <= 1
    assert np.max(np.abs(rft - rff)) < 0.15 * np.max(np.abs(rff))
    with pytest.raises(ValueError):
        oracle.rf_forward(THK, RHO, VP, VS, Q, Q, 0.045, 125, 0.4, 1.5, 5.0, "freq", 0.001, "X")


def test_layer_over_halfspace_conversion_and_multiple_times(oracle):
    """Independent of any reference code: for one layer over a half-space the P receiver function
    has its Ps conversion at H(eta_b - eta_a), the PpPs multiple at H(eta_b + eta_a) (both positive
    for a velocity increase) and the PpSs+PsPs multiple at 2 H eta_b (negative), with
    eta = sqrt(1/v^2 - p^2).  Pins timing, polarity and the time_shift convention of the trace."""
    H, p, dt, nt, tshift = 35.0, 0.06, 0.05, 1024, 5.0
    vs = np.array([3.5, 4.5]); vp = np.array([6.1, 8.0]); rho = np.array([2.7, 3.3])
    thk = np.array([H, 0.0])
    q = thk * 0 + 9999.
    for method in ("freq", "time"):
        rf = oracle.rf_forward(thk, rho, vp, vs, q, q, p, nt, dt, 3.0, tshift, method, 0.001, "P")
        t = np.arange(nt) * dt - tshift
        eta_a, eta_b = np.sqrt(1 / vp[0]**2 - p**2), np.sqrt(1 / vs[0]**2 - p**2)
        t_ps, t_ppps, t_ppss = H * (eta_b - eta_a), H * (eta_b + eta_a), 2 * H * eta_b

        def extremum(t0, sign):
            w = np.abs(t - t0) < 0.6
            k = np.argmax(sign * rf[w])
            return t[w][k], rf[w][k]
        k0 = np.argmax(rf)
        assert abs(t[k0]) < 1.5 * dt and rf[k0] > 0           # direct P at t = 0
        tp, ap = extremum(t_ps, +1)
        assert abs(tp - t_ps) <= 2 * dt and 0.05 * rf[k0] < ap < rf[k0], method
        tm, am = extremum(t_ppps, +1)
        assert abs(tm - t_ppps) <= 2 * dt and am > 0.02 * rf[k0], method
        tn, an = extremum(t_ppss, -1)
        assert abs(tn - t_ppss) <= 2 * dt and an < -0.02 * rf[k0], method


def test_halfspace_amplitude_is_the_free_surface_response_ratio(oracle):
    """Independent of any reference code: a P wave with horizontal slowness p hitting the free surface of
    a half-space gives radial / vertical displacement = 2 p eta_b / (1/b^2 - 2 p^2),
    eta_b = sqrt(1/b^2 - p^2); the receiver function is that ratio times the Gaussian low-pass
    exp(-w^2/4a^2), whose time-domain peak is a/sqrt(pi).  Pins the absolute amplitude convention."""
    for vsv, p, a, tol in ((3.6, 0.06, 2.5, 1e-6), (4.0, 0.07, 3.0, 1e-6), (3.2, 0.045, 1.5, 3e-3)):
        vs = np.full(3, vsv); thk = np.array([10., 10., 0.])
        vp, rho = brocher(vs)
        q = thk * 0 + 9999.
        for method in ("freq", "time"):
            rf = oracle.rf_forward(thk, rho, vp, vs, q, q, p, 512, 0.05, a, 5.0, method, 0.001, "P")
            b = vs[0]
            ratio = 2 * p * np.sqrt(1 / b**2 - p**2) / (1 / b**2 - 2 * p**2)
            want = ratio * a / np.sqrt(np.pi)
            t2 = 5e-3 if method == "time" else tol   # deconit represents the pulse by discrete spikes
            assert abs(rf.max() - want) <= max(tol, t2) * want, (vsv, p, a, method, rf.max(), want)


def test_multilayer_rf_against_an_independent_plane_wave_solution(oracle):
    """cal_rf_freq for layered models against a from-scratch solution of the same physics: every
    layer carries four plane-wave potentials (P/S, up/down), the half-space the incident P plus two
    radiating waves; stress-free surface and welded interfaces give ONE linear system (NumPy), whose
    surface displacement ratio ux/uz is the receiver-function spectrum.  Conventions shared with the
    reference: Gaussian exp(-w^2/4a^2), complex frequency w - i sigma with sigma = 4/(nft dt) undone
    by exp(sigma (t - t_shift)), velocities v (1 + 1/(8Q^2) + i/(2Q)) with the reference's dummy
    Q = 9999.  Agreement: 3e-7 of the peak on the 7-layer F1 model (float32 pi, water level)."""
    cases = [(np.array([6., 6, 13, 5, 10, 30, 0]), np.array([3.2, 2.8, 3.46, 3.3, 3.9, 4.5, 4.7]), 0.045, 125, 0.4, 1.5),
             (np.array([4., 18., 12., 0.]), np.array([2.6, 3.5, 3.9, 4.6]), 0.07, 250, 0.2, 2.0)]
    for thk, vs, p, nt, dt, a in cases:
        vp, rho = brocher(vs)
        q = thk * 0 + 9999.
        tshift = 5.0
        rf = oracle.rf_forward(thk, rho, vp, vs, q, q, p, nt, dt, a, tshift, "freq", 0.001, "P")
        nft = 1
        while nft < nt:
            nft *= 2
        sigma = 4.0 / (nft * dt)
        qf = 1 + 1 / (8 * 9999.**2) + 1j / (2 * 9999.)
        H = np.zeros(nft // 2 + 1, dtype=complex)
        for kf in range(nft // 2 + 1):
            w = 2 * np.pi * kf / (nft * dt) - 1j * sigma
            ux, uz = surface_response(w, p, thk, vp * qf, vs * qf, rho)
            H[kf] = ux / uz
        wr = 2 * np.pi * np.arange(nft // 2 + 1) / (nft * dt)
        spec = H * np.exp(-wr**2 / (4 * a * a)) * np.exp(-1j * wr * tshift)
        tr = np.fft.irfft(spec, nft)[:nt] / dt * np.exp(sigma * (np.arange(nt) * dt - tshift))
        assert np.max(np.abs(tr - rf)) <= 2e-6 * np.max(np.abs(rf)), (p, np.max(np.abs(tr - rf)) / np.max(np.abs(rf)))
        # S receiver function: incident SV wave, vertical over radial, time shift of opposite sign
        # (src/RF/main.cpp:35); a ray parameter for which P is still propagating in the half-space
        ps = 0.1
        rfs = oracle.rf_forward(thk, rho, vp, vs, q, q, ps, nt, dt, a, tshift, "freq", 0.001, "S")
        for kf in range(nft // 2 + 1):
            w = 2 * np.pi * kf / (nft * dt) - 1j * sigma
            ux, uz = surface_response(w, ps, thk, vp * qf, vs * qf, rho, incident="S")
            H[kf] = uz / ux
        spec = H * np.exp(-wr**2 / (4 * a * a)) * np.exp(+1j * wr * tshift)
        tr = np.fft.irfft(spec, nft)[:nt] / dt * np.exp(sigma * (np.arange(nt) * dt + tshift))
        assert np.max(np.abs(tr - rfs)) <= 2e-6 * np.max(np.abs(rfs)), ("S", np.max(np.abs(tr - rfs)) / np.max(np.abs(rfs)))
