"""Pins of the SWD oracle (no reference test exists): analytic known answers,
finite differences of its own forward, exact scaling identities, mode bookkeeping."""
import numpy as np
import pytest
from oracle.oracle import brocher
from rfsurfhmc_b200.fixtures import f1_true_model

X0 = f1_true_model()
VS, THK = X0[:7], X0[7:]
VP, RHO = brocher(VS)


def rayleigh_halfspace_speed(a, b):
    """root of (2-k^2)^2 = 4 sqrt(1-g^2 k^2) sqrt(1-k^2), k=c/b, g=b/a (same equation gtsolh iterates,
    surfdisp96.f:380-394), by bisection in double precision."""
    g = b / a
    f = lambda k: (2 - k * k)**2 - 4 * np.sqrt(1 - g * g * k * k) * np.sqrt(1 - k * k)
    lo, hi = 0.5, 0.999999
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if f(lo) * f(mid) <= 0:
            hi = mid
        else:
            lo = mid
    return 0.5 * (lo + hi) * b


def test_rayleigh_homogeneous_halfspace(oracle):
    # identical layers == half-space: non-dispersive Rayleigh wave at the analytic speed
    n = 4
    vs = np.full(n, 3.5); vp = np.full(n, 6.0); rho = np.full(n, 2.7); thk = np.array([5., 5., 5., 0.])
    T = np.array([2., 5., 10., 20., 40.])
    c, ok = oracle.surf_forward(thk, vp, vs, rho, T, "Rc")
    cr = rayleigh_halfspace_speed(float(np.float32(6.0)), float(np.float32(3.5)))
    assert ok
    # nevill's bias (<= ~1e-6 relative, always low) + float32 output rounding
    assert np.all(np.abs(c - cr) / cr < 1.5e-6)
    u, ok = oracle.surf_forward(thk, vp, vs, rho, T, "Rg")
    assert ok and np.all(np.abs(u - cr) / cr < 5e-6)  # U == c without dispersion


def rayleigh_det_2layer(cc, Tp, a1, b1, r1, H, a2, b2, r2):
    """Textbook boundary-condition determinant of P-SV motion for ONE layer over a half-space, written
    from scratch with potentials (independent of the Dunkin/Haskell formulation of the reference):
    layer phi = A e^{-ra z} + B e^{ra z}, psi = C e^{-rb z} + D e^{rb z}; half-space decaying only.
    Unknowns (A,B,C,D,E,F); stress-free surface (2 rows) + welded interface (4 rows)."""
    w = 2 * np.pi / Tp
    k = w / cc
    ra1 = np.sqrt(complex(k * k - (w / a1)**2)); rb1 = np.sqrt(complex(k * k - (w / b1)**2))
    ra2 = np.sqrt(complex(k * k - (w / a2)**2)); rb2 = np.sqrt(complex(k * k - (w / b2)**2))
    mu1, mu2 = r1 * b1 * b1, r2 * b2 * b2

    def col_p(r, mu, bvel, z, sgn):   # phi = e^{sgn r z}: (ux, uz, tzz, txz), e^{i(kx-wt)} dropped
        e = np.exp(sgn * r * z)
        return np.array([1j * k, sgn * r, mu * (2 * k * k - (w / bvel)**2), 2j * mu * k * sgn * r]) * e

    def col_s(r, mu, bvel, z, sgn):   # psi = e^{sgn r z}: ux = -dpsi/dz, uz = dpsi/dx
        e = np.exp(sgn * r * z)
        return np.array([-sgn * r, 1j * k, 2j * mu * k * sgn * r, -mu * (2 * k * k - (w / bvel)**2)]) * e
    M = np.zeros((6, 6), dtype=complex)
    top = [col_p(ra1, mu1, b1, 0.0, -1), col_p(ra1, mu1, b1, 0.0, +1),
           col_s(rb1, mu1, b1, 0.0, -1), col_s(rb1, mu1, b1, 0.0, +1)]
    bot = [col_p(ra1, mu1, b1, H, -1), col_p(ra1, mu1, b1, H, +1),
           col_s(rb1, mu1, b1, H, -1), col_s(rb1, mu1, b1, H, +1)]
    half = [col_p(ra2, mu2, b2, 0.0, -1), col_s(rb2, mu2, b2, 0.0, -1)]
    for j in range(4):
        M[0, j], M[1, j] = top[j][2], top[j][3]
        M[2:6, j] = bot[j]
    for j in range(2):
        M[2:6, 4 + j] = -half[j]
    M[:, 1] /= np.exp(ra1.real * H) if ra1.real > 0 else 1.0   # keep the growing columns O(1)
    M[:, 3] /= np.exp(rb1.real * H) if rb1.real > 0 else 1.0
    return np.linalg.det(M)


def rayleigh_root_2layer(c_guess, Tp, *par, width=3e-4):
    """Root of the independent determinant next to c_guess, by bisection to machine precision."""
    lo, hi = c_guess * (1 - width), c_guess * (1 + width)
    d0 = rayleigh_det_2layer(lo, Tp, *par)
    ph = d0 / abs(d0)
    f = lambda x: (rayleigh_det_2layer(x, Tp, *par) / ph).real
    flo, fhi = f(lo), f(hi)
    assert np.sign(flo) != np.sign(fhi), (c_guess, Tp, flo, fhi)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        fm = f(mid)
        if np.sign(fm) == np.sign(flo):
            lo, flo = mid, fm
        else:
            hi = mid
    return 0.5 * (lo + hi)


def test_rayleigh_kernels_and_group_velocity_against_the_independent_determinant(oracle):
    """sregn96/sregnpu outputs for a layer over a half-space -- analytic group velocity (energy
    integrals) and the phase-velocity kernels dc/dvp, dc/dvs, dc/drho, dc/dh -- against derivatives
    of the roots of the from-scratch determinant above, taken in double precision around the
    float32-rounded model the reference works on."""
    H = 10.0
    vs = np.array([3.0, 4.2]); vp = np.array([5.2, 7.4]); rho = np.array([2.5, 3.2]); thk = np.array([H, 0.0])
    f32 = lambda v: float(np.float32(v))
    base = [f32(vp[0]), f32(vs[0]), f32(rho[0]), f32(H), f32(vp[1]), f32(vs[1]), f32(rho[1])]
    T = np.array([3.0, 8.0, 15.0])
    c, da, db, dr, dh, ok = oracle.surf_adjoint_kernel(thk, vp, vs, rho, T, "Rc")
    u, ok2 = oracle.surf_forward(thk, vp, vs, rho, T, "Rg")
    assert ok and ok2
    # index of each parameter in `base` and the oracle array / layer it belongs to
    wrt = [(0, da, 0), (1, db, 0), (2, dr, 0), (3, dh, 0), (4, da, 1), (5, db, 1), (6, dr, 1)]
    for i, Tp in enumerate(T):
        c0 = rayleigh_root_2layer(c[i], Tp, *base)
        assert abs(c0 - c[i]) < 2e-6 * c0                      # float32 output + nevill's 1e-6 bracket
        # group velocity U = c / (1 + (T/c) dc/dT) from the independent roots
        e = 1e-4 * Tp
        dcdT = (rayleigh_root_2layer(c0, Tp + e, *base) - rayleigh_root_2layer(c0, Tp - e, *base)) / (2 * e)
        assert abs(u[i] - c0 / (1 + Tp / c0 * dcdT)) < 2e-5 * u[i], (Tp, u[i])
        for j, arr, layer in wrt:
            h = 1e-5 * base[j]
            pp, pm = list(base), list(base)
            pp[j] += h
            pm[j] -= h
            fd = (rayleigh_root_2layer(c0, Tp, *pp) - rayleigh_root_2layer(c0, Tp, *pm)) / (2 * h)
            # the reference forms 2*b*b and rho*rho in REAL*4 inside dnka/hska: its eigenfunctions carry
            # ~1e-7 relative noise, visible in the small (nearly cancelling) density kernels
            big = max(np.max(np.abs(k_[i])) for k_ in (da, db, dr, dh))
            assert abs(arr[i, layer] - fd) < 2e-4 * abs(fd) + 1e-5 * big, (Tp, j, arr[i, layer], fd)


def _independent_root(c_guess, Tp, par, width=3e-4):
    from independent import rayleigh_secular
    lo, hi = c_guess * (1 - width), c_guess * (1 + width)
    d0 = rayleigh_secular(lo, Tp, *par)
    ph = d0 / abs(d0)
    f = lambda x: (rayleigh_secular(x, Tp, *par) / ph).real
    flo = f(lo)
    assert np.sign(flo) != np.sign(f(hi)), (c_guess, Tp)
    for _ in range(50):
        mid = 0.5 * (lo + hi)
        fm = f(mid)
        if np.sign(fm) == np.sign(flo):
            lo, flo = mid, fm
        else:
            hi = mid
    return 0.5 * (lo + hi)


def test_f1_rayleigh_against_the_independent_multilayer_system(oracle):
    """The reference's default 7-layer model (param.yaml) against tests/independent.py: one linear
    system of plane-wave potentials with all boundary conditions (26 x 26), nothing of the
    Haskell/Dunkin machinery.  Roots of the fundamental and first higher mode (1.5e-6: nevill's
    bracket + float32 output), analytic group velocity, and dc/dvs, dc/dh of EVERY layer at 20 s."""
    f32 = lambda a: np.float32(a).astype(float)
    par = [f32(THK), f32(VP), f32(VS), f32(RHO)]
    T = np.array([5., 8., 12., 20., 30., 40.])
    for mode in (0, 1):
        c, ok = oracle.surf_forward(THK, VP, VS, RHO, T, "Rc", mode)
        assert ok
        for Tp, ck in zip(T, c):
            if ck == 0.0:
                continue   # mode does not exist at this period
            assert abs(_independent_root(ck, Tp, par) - ck) < 1.5e-6 * ck, (mode, Tp)
    Tp = 20.0
    c, da, db, dr, dh, ok = oracle.surf_adjoint_kernel(THK, VP, VS, RHO, np.array([Tp]), "Rc")
    u, _ = oracle.surf_forward(THK, VP, VS, RHO, np.array([Tp]), "Rg")
    c0 = _independent_root(c[0], Tp, par)
    e = 1e-4 * Tp
    dcdT = (_independent_root(c0, Tp + e, par) - _independent_root(c0, Tp - e, par)) / (2 * e)
    assert abs(u[0] - c0 / (1 + Tp / c0 * dcdT)) < 2e-5 * u[0]
    big = max(np.max(np.abs(k_)) for k_ in (da, db, dr, dh))
    for which, arr in ((2, db), (0, dh)):        # index into par: 2 = vs, 0 = thickness
        for m in range(7 if which == 2 else 6):
            h = 1e-5 * max(par[which][m], 1.0)
            pp = [a.copy() for a in par]
            pm = [a.copy() for a in par]
            pp[which][m] += h
            pm[which][m] -= h
            fd = (_independent_root(c0, Tp, pp) - _independent_root(c0, Tp, pm)) / (2 * h)
            assert abs(arr[0, m] - fd) < 2e-4 * abs(fd) + 1e-5 * big, (which, m, arr[0, m], fd)


def test_f1_love_against_the_independent_multilayer_system(oracle):
    """Same model, Love waves: roots of modes 0 and 1, analytic group velocity and dc/dvs, dc/drho,
    dc/dh of every layer at 20 s against the from-scratch SH system of tests/independent.py."""
    from independent import love_secular
    f32 = lambda a: np.float32(a).astype(float)
    par = [f32(THK), f32(VS), f32(RHO)]

    def root(c_guess, Tp, pr, width=3e-4):
        lo, hi = c_guess * (1 - width), c_guess * (1 + width)
        d0 = love_secular(lo, Tp, *pr)
        ph = d0 / abs(d0)
        f = lambda x: (love_secular(x, Tp, *pr) / ph).real
        flo = f(lo)
        assert np.sign(flo) != np.sign(f(hi)), (c_guess, Tp)
        for _ in range(50):
            mid = 0.5 * (lo + hi)
            fm = f(mid)
            if np.sign(fm) == np.sign(flo):
                lo, flo = mid, fm
            else:
                hi = mid
        return 0.5 * (lo + hi)
    T = np.array([5., 8., 12., 20., 30., 40.])
    for mode in (0, 1):
        c, ok = oracle.surf_forward(THK, VP, VS, RHO, T, "Lc", mode)
        assert ok
        for Tp, ck in zip(T, c):
            if ck != 0.0:
                assert abs(root(ck, Tp, par) - ck) < 1.5e-6 * ck, (mode, Tp)
    Tp = 20.0
    c, da, db, dr, dh, ok = oracle.surf_adjoint_kernel(THK, VP, VS, RHO, np.array([Tp]), "Lc")
    u, _ = oracle.surf_forward(THK, VP, VS, RHO, np.array([Tp]), "Lg")
    c0 = root(c[0], Tp, par)
    e = 1e-4 * Tp
    dcdT = (root(c0, Tp + e, par) - root(c0, Tp - e, par)) / (2 * e)
    assert abs(u[0] - c0 / (1 + Tp / c0 * dcdT)) < 2e-5 * u[0]
    big = max(np.max(np.abs(k_)) for k_ in (db, dr, dh))
    for which, arr, nl in ((1, db, 7), (2, dr, 7), (0, dh, 6)):
        for m in range(nl):
            h = 1e-5 * max(par[which][m], 1.0)
            pp = [a.copy() for a in par]
            pm = [a.copy() for a in par]
            pp[which][m] += h
            pm[which][m] -= h
            fd = (root(c0, Tp, pp) - root(c0, Tp, pm)) / (2 * h)
            assert abs(arr[0, m] - fd) < 2e-4 * abs(fd) + 1e-5 * big, (which, m, arr[0, m], fd)


def test_group_velocity_kernels_against_the_independent_determinant(oracle):
    """sregnpu builds dU/dm from three solves at T, 1.05 T and 0.95 T (a +-5 % period difference),
    and the reference takes the first term from the wrong solve (the "stale array" defect,
    sregn96.f90:1841-1844).  Against derivatives of U computed from the from-scratch determinant:
    the corrected form (stale=False) is the +-5 % difference approximation of the true kernel
    (within 5 % of the largest kernel of the period), and the reference's form is measurably
    further away -- which is why it is reproduced by switch, not by accident."""
    H = 10.0
    vs = np.array([3.0, 4.2]); vp = np.array([5.2, 7.4]); rho = np.array([2.5, 3.2]); thk = np.array([H, 0.0])
    f32 = lambda v: float(np.float32(v))
    base = [f32(vp[0]), f32(vs[0]), f32(rho[0]), f32(H), f32(vp[1]), f32(vs[1]), f32(rho[1])]
    T = np.array([3.0, 8.0, 15.0])
    c, _ = oracle.surf_forward(thk, vp, vs, rho, T, "Rc")

    def U_ind(cg, Tp, par):
        c0 = rayleigh_root_2layer(cg, Tp, *par)
        e = 1e-4 * Tp
        d = (rayleigh_root_2layer(c0, Tp + e, *par) - rayleigh_root_2layer(c0, Tp - e, *par)) / (2 * e)
        return c0 / (1 + Tp / c0 * d)
    err = {}
    for stale in (False, True):
        u, da, db, dr, dh, ok = oracle.surf_adjoint_kernel(thk, vp, vs, rho, T, "Rg", 0, False, stale)
        assert ok
        worst = 0.0
        for i, Tp in enumerate(T):
            fds, got = [], []
            for j, (arr, layer) in enumerate([(da, 0), (db, 0), (dr, 0), (dh, 0), (da, 1), (db, 1), (dr, 1)]):
                h = 1e-4 * base[j]
                pp, pm = list(base), list(base)
                pp[j] += h
                pm[j] -= h
                fds.append((U_ind(c[i], Tp, pp) - U_ind(c[i], Tp, pm)) / (2 * h))
                got.append(arr[i, layer])
            fds, got = np.array(fds), np.array(got)
            worst = max(worst, np.max(np.abs(got - fds)) / np.max(np.abs(fds)))
        err[stale] = worst
    assert err[False] < 0.05, err
    assert err[True] > 1.5 * err[False], err


def test_rayleigh_layer_over_halfspace_limits_and_exact_secular_root(oracle):
    """Layer over a half-space, independent of the reference code:
    (i) short periods see only the layer, long periods only the half-space: c -> the analytic
        Rayleigh speeds of the two media; the curve is monotonic in between;
    (ii) at an intermediate period the root must annihilate the textbook 4x4 boundary-condition
         determinant (free surface + welded interface + radiation), written here from scratch with
         potentials in numpy -- a different formulation than the Dunkin compound matrix."""
    H = 10.0
    vs = np.array([3.0, 4.2]); vp = np.array([5.2, 7.4]); rho = np.array([2.5, 3.2]); thk = np.array([H, 0.0])
    T = np.array([0.4, 1.0, 3.0, 8.0, 15.0, 40.0, 150.0, 600.0])
    c, ok = oracle.surf_forward(thk, vp, vs, rho, T, "Rc")
    assert ok and np.all(np.diff(c) > -1e-5)   # flat (non-dispersive) in the short-period limit
    f32 = lambda v: float(np.float32(v))
    assert abs(c[0] - rayleigh_halfspace_speed(f32(vp[0]), f32(vs[0]))) < 2e-4 * c[0]
    assert abs(c[-1] - rayleigh_halfspace_speed(f32(vp[1]), f32(vs[1]))) < 5e-3 * c[-1]  # H/lambda = 0.4 %

    par = [f32(vp[0]), f32(vs[0]), f32(rho[0]), H, f32(vp[1]), f32(vs[1]), f32(rho[1])]
    det = lambda cc, Tp: rayleigh_det_2layer(cc, Tp, *par)

    for Tp, ck in ((3.0, c[2]), (8.0, c[3]), (15.0, c[4])):
        cs = ck * (1 + np.array([-2e-4, -1e-4, 0.0, 1e-4, 2e-4]))
        d = np.array([det(x, Tp) for x in cs])
        # the determinant is (up to a constant phase) real and changes sign at the root:
        ph = d[0] / abs(d[0])
        dr = (d / ph).real
        assert np.sign(dr[0]) != np.sign(dr[-1]), (Tp, dr)
        # linear interpolation of the sign change lands on the reported root within float32 + 1e-6
        root = cs[0] + (cs[-1] - cs[0]) * (0 - dr[0]) / (dr[-1] - dr[0])
        assert abs(root - ck) < 3e-6 * ck, (Tp, root, ck)

    # (iii) higher modes: the first and second higher-mode roots annihilate the same determinant, and
    # it has NO further sign change between consecutive returned roots (no mode is skipped).  Below
    # and above the layer's S velocity the determinant is purely real resp. imaginary, so sign
    # changes are counted per segment.
    def sign_changes(lo, hi, Tp, step=2e-4):
        g = np.arange(lo, hi, step)
        d = np.array([det(x, Tp) for x in g])
        r = d / (d[0] / abs(d[0]))
        assert np.max(np.abs(r.imag)) < 1e-9 * np.max(np.abs(d))
        return int(np.sum(np.diff(np.sign(r.real)) != 0))
    Tm = np.array([0.8, 1.5, 2.5])
    cm = [oracle.surf_forward(thk, vp, vs, rho, Tm, "Rc", mode)[0] for mode in (0, 1, 2)]
    b1 = f32(vs[0])
    for i, Tp in enumerate(Tm):
        x0, x1, x2 = cm[0][i], cm[1][i], cm[2][i]
        assert x0 < b1 < x1 < x2
        for ck in (x1, x2):
            cs = ck * (1 + np.array([-2e-4, 2e-4]))
            d = np.array([det(x, Tp) for x in cs])
            dr = (d / (d[0] / abs(d[0]))).real
            assert np.sign(dr[0]) != np.sign(dr[1])
            root = cs[0] + (cs[1] - cs[0]) * (0 - dr[0]) / (dr[1] - dr[0])
            assert abs(root - ck) < 3e-6 * ck, (Tp, root, ck)
        assert sign_changes(x0 + 3e-4, b1 - 3e-4, Tp) == 0
        assert sign_changes(b1 + 3e-4, x1 - 3e-4, Tp) == 0
        assert sign_changes(x1 + 3e-4, x2 - 3e-4, Tp) == 0


def test_love_layer_over_halfspace(oracle):
    # tan(q h) = mu2 nu2 / (mu1 q): residual of the classical dispersion relation at the oracle roots
    thk = np.array([20., 0.]); vs = np.array([3.0, 4.5]); vp = vs * 1.75; rho = np.array([2.5, 3.2])
    T = np.array([5., 8., 12., 20., 30.])
    for mode in (0, 1):
        c, ok = oracle.surf_forward(thk, vp, vs, rho, T, "Lc", mode=mode)
        assert ok
        for t, cc in zip(T, c):
            if cc == 0.0:
                continue
            w = 2 * np.pi / t
            k = w / cc
            b1, b2 = float(np.float32(3.0)), float(np.float32(4.5))
            r1, r2 = float(np.float32(2.5)), float(np.float32(3.2))
            q = np.sqrt((w / b1)**2 - k * k)
            nu2 = np.sqrt(k * k - (w / b2)**2)
            lhs = np.tan(q * 20.0)
            rhs = (r2 * b2 * b2 * nu2) / (r1 * b1 * b1 * q)
            # the root is only located to ~1e-6 relative; compare through the phase
            ph = np.arctan(rhs) + mode * np.pi
            assert abs(q * 20.0 - ph) < 2e-4, (mode, t, cc, lhs, rhs)


def test_love_kernels_and_group_velocity_against_the_analytic_dispersion_relation(oracle):
    """slegn96/slegnpu outputs for a layer over a half-space against derivatives of the roots of the
    closed-form Love dispersion relation mu1 q sin(qH) = mu2 nu2 cos(qH) (q = sqrt(w^2/b1^2 - k^2),
    nu2 = sqrt(k^2 - w^2/b2^2)), taken in double precision around the float32-rounded model:
    analytic group velocity, dc/dvs, dc/drho of both media and dc/dh, fundamental and first higher mode."""
    f32 = lambda v: float(np.float32(v))
    thk = np.array([20.0, 0.0]); vs = np.array([3.0, 4.5]); vp = np.array([5.2, 7.8]); rho = np.array([2.5, 3.2])
    base = [f32(vs[0]), f32(rho[0]), f32(thk[0]), f32(vs[1]), f32(rho[1])]

    def G(cc, Tp, b1, r1, H, b2, r2):
        w = 2 * np.pi / Tp
        k = w / cc
        q = np.sqrt((w / b1)**2 - k * k)
        nu2 = np.sqrt(k * k - (w / b2)**2)
        return r1 * b1 * b1 * q * np.sin(q * H) - r2 * b2 * b2 * nu2 * np.cos(q * H)

    def root(c_guess, Tp, *par):
        lo, hi = c_guess * (1 - 3e-4), c_guess * (1 + 3e-4)
        flo = G(lo, Tp, *par)
        assert np.sign(flo) != np.sign(G(hi, Tp, *par))
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            fm = G(mid, Tp, *par)
            if np.sign(fm) == np.sign(flo):
                lo, flo = mid, fm
            else:
                hi = mid
        return 0.5 * (lo + hi)
    for mode, T in ((0, np.array([4.0, 10.0, 25.0])), (1, np.array([3.0, 6.0]))):
        c, da, db, dr, dh, ok = oracle.surf_adjoint_kernel(thk, vp, vs, rho, T, "Lc", mode)
        u, ok2 = oracle.surf_forward(thk, vp, vs, rho, T, "Lg", mode)
        assert ok and ok2 and np.all(c > 0)
        wrt = [(0, db, 0), (1, dr, 0), (2, dh, 0), (3, db, 1), (4, dr, 1)]
        for i, Tp in enumerate(T):
            c0 = root(c[i], Tp, *base)
            assert abs(c0 - c[i]) < 2e-6 * c0
            e = 1e-4 * Tp
            dcdT = (root(c0, Tp + e, *base) - root(c0, Tp - e, *base)) / (2 * e)
            assert abs(u[i] - c0 / (1 + Tp / c0 * dcdT)) < 2e-5 * u[i], (mode, Tp)
            big = max(np.max(np.abs(k_[i])) for k_ in (db, dr, dh))
            for j, arr, layer in wrt:
                h = 1e-5 * base[j]
                pp, pm = list(base), list(base)
                pp[j] += h
                pm[j] -= h
                fd = (root(c0, Tp, *pp) - root(c0, Tp, *pm)) / (2 * h)
                assert abs(arr[i, layer] - fd) < 2e-4 * abs(fd) + 1e-5 * big, (mode, Tp, j, arr[i, layer], fd)


def test_f1_values_and_modes(oracle):
    T = np.arange(5., 41.)
    c, ok = oracle.surf_forward(THK, VP, VS, RHO, T, "Rc")
    assert ok
    # independent scratch-probe values of SURVEY.md Appendix C (float32-rounded on output)
    ref = {5: 2.811252546, 6: 2.802712276, 10: 2.876952767, 20: 3.247148090, 40: 3.882736766}
    for t, v in ref.items():
        assert abs(c[int(t) - 5] - v) / v < 1e-7
    # reversed dispersion below 6 s (low-velocity second layer) exercises the downward search
    assert c[1] < c[0]
    # mode 2 does not exist at long periods: zeros with flag True (surfdisp96.f:317,356-362)
    c2, ok2 = oracle.surf_forward(THK, VP, VS, RHO, np.array([2., 11.]), "Rc", mode=2)
    assert ok2 and c2[0] > 3.0 and c2[1] == 0.0
    n_ev = oracle.surfdisp96_evals(THK, VP, VS, RHO, T)
    assert 700 < n_ev < 1100  # ~125 scan steps once + ~21 per later period


@pytest.mark.parametrize("wt", ["Rc", "Lc"])
def test_phase_kernels_finite_difference(oracle, wt):
    T = np.array([5., 12., 25., 40.])
    c, da, db, dr, dh, ok = oracle.surf_adjoint_kernel(THK, VP, VS, RHO, T, wt)
    assert ok

    def roots(thk, vp, vs, rho):
        # tight-root evaluation through the kernel's own U/c is impossible; use many-period forward
        return oracle.surf_forward(thk, vp, vs, rho, T, wt)[0]
    # the forward is float32-quantised (6e-8) and nevill-biased (1e-6), so FD needs big steps:
    # compare with a relative tolerance that reflects that noise (few %) on the dominant entries
    for arr, ker in ((VS, db), (THK, dh)):
        for j in range(6):
            h = 0.02 * max(arr[j], 1.0)
            p = arr.copy(); p[j] += h
            m = arr.copy(); m[j] -= h
            if arr is VS:
                fd = (roots(THK, VP, p, RHO) - roots(THK, VP, m, RHO)) / (2 * h)
            else:
                fd = (roots(p, VP, VS, RHO) - roots(m, VP, VS, RHO)) / (2 * h)
            big = np.abs(ker[:, j]) > 0.02
            if big.any():
                assert np.max(np.abs(fd[big] - ker[big, j]) / np.abs(ker[big, j])) < 0.03


def test_euler_homogeneity_and_density_invariance(oracle):
    T = np.array([5., 8., 12., 20., 30., 40.])
    c, da, db, dr, dh, ok = oracle.surf_adjoint_kernel(THK, VP, VS, RHO, T, "Rc")
    euler = (da * VP).sum(1) + (db * VS).sum(1) + (dh * THK).sum(1)
    assert np.max(np.abs(euler - c) / c) < 5e-6
    assert np.max(np.abs((dr * RHO).sum(1))) < 2e-5
    c, da, db, dr, dh, ok = oracle.surf_adjoint_kernel(THK, VP, VS, RHO, T, "Lc")
    euler = (db * VS).sum(1) + (dh * THK).sum(1)
    assert np.max(np.abs(euler - c) / c) < 5e-6
    assert np.max(np.abs((dr * RHO).sum(1))) < 2e-5
    assert np.all(da == 0.0)


def test_group_velocity_consistent_with_dispersion(oracle):
    # U = c / (1 + (T/c) dc/dT): energy-integral U vs numerical differentiation of the phase curve
    T = np.linspace(8., 36., 15)
    u, ok = oracle.surf_forward(THK, VP, VS, RHO, T, "Rg")
    h = 0.25
    cp, _ = oracle.surf_forward(THK, VP, VS, RHO, T + h, "Rc")
    cm, _ = oracle.surf_forward(THK, VP, VS, RHO, T - h, "Rc")
    c0, _ = oracle.surf_forward(THK, VP, VS, RHO, T, "Rc")
    un = c0 / (1 + (T / c0) * (cp - cm) / (2 * h))
    assert np.max(np.abs(u - un) / u) < 2e-4


def test_group_kernel_identity_and_stale_switch(oracle):
    # dU/dm = (U/c)(2-U/c) dcdm[first] - (U/c)^2 T (dcdm(T2)-dcdm(T1))/(T2-T1)   (sregn96.f90:1839-1844)
    T = np.array([8., 20., 35.])
    t1, t2 = T * (1.0 + 0.05), T * (1.0 - 0.05)
    c0, a0, b0, r0, h0, _ = oracle.surf_adjoint_kernel(THK, VP, VS, RHO, T, "Rc")
    c1, a1, b1, r1, h1, _ = oracle.surf_adjoint_kernel(THK, VP, VS, RHO, t1, "Rc")
    c2, a2, b2, r2, h2, _ = oracle.surf_adjoint_kernel(THK, VP, VS, RHO, t2, "Rc")
    u, _ = oracle.surf_forward(THK, VP, VS, RHO, T, "Rg")
    for stale in (True, False):
        ug, ga, gb, gr, gh, ok = oracle.surf_adjoint_kernel(THK, VP, VS, RHO, T, "Rg", stale=stale)
        assert ok and np.allclose(ug, u, rtol=1e-12)
        uc = (u / c0)[:, None]
        first = b2 if stale else b0
        expect = uc * (2 - uc) * first - uc**2 * T[:, None] * (b2 - b1) / (t2 - t1)[:, None]
        assert np.allclose(gb, expect, rtol=1e-9, atol=1e-13)
    # the two forms differ by percents: the reference's stale form is what parity needs
    ga_s = oracle.surf_adjoint_kernel(THK, VP, VS, RHO, T, "Rg", stale=True)[2]
    ga_f = oracle.surf_adjoint_kernel(THK, VP, VS, RHO, T, "Rg", stale=False)[2]
    assert np.max(np.abs(ga_s - ga_f)) / np.max(np.abs(ga_f)) > 5e-3


def test_invalid_wavetype_raises(oracle):
    with pytest.raises(ValueError):
        oracle.surf_forward(THK, VP, VS, RHO, [5.], "Xx")


def test_spherical_earth_oracle_sanity(oracle):
    """sphere=True: flattening is a small, period-growing correction; the phase kernels of the
    spherical model still satisfy c = sum(vp dc/dvp + vs dc/dvs + h dc/dh) approximately."""
    T = np.array([5., 10., 20., 40., 80.])
    for wt in ("Rc", "Lc"):
        cf, _ = oracle.surf_forward(THK, VP, VS, RHO, T, wt, 0, False)
        cs, ok = oracle.surf_forward(THK, VP, VS, RHO, T, wt, 0, True)
        d = (cs - cf) / cf
        assert ok and np.all(np.abs(d) < 0.03) and abs(d[-1]) > abs(d[0])
        c, da, db, dr, dh, ok = oracle.surf_adjoint_kernel(THK, VP, VS, RHO, T, wt, 0, True)
        assert ok and np.all(np.isfinite(db)) and db.min() > -0.02 and db.max() > 0.1
    ug, ok = oracle.surf_forward(THK, VP, VS, RHO, T, "Rg", 0, True)
    uf, _ = oracle.surf_forward(THK, VP, VS, RHO, T, "Rg", 0, False)
    assert ok and np.all(np.abs(ug - uf) / uf < 0.05)


def test_water_layer_against_the_independent_ocean_system(oracle):
    """Ocean model (fluid top layer) against tests/independent.py::rayleigh_secular_ocean -- two
    compressional waves in the water, pressure-free surface, no shear traction at the sea floor:
    roots at five periods, analytic group velocity, and at 15 s the kernels dc/dvp (water and solids),
    dc/dvs (solids), dc/drho and dc/dh (including the water depth).  Pins the fluid branches of the
    restated surfdisp96 (dltar4 water term) and sregn96 (dnka/hska/intijr/energy/getdcdh)."""
    from independent import rayleigh_secular_ocean
    thk = np.array([3.0, 2.0, 5.0, 12.0, 0.0])
    vs = np.array([0.0, 2.2, 3.3, 3.9, 4.6])
    vp = np.array([1.5, 4.2, 5.9, 6.8, 8.1])
    rho = np.array([1.03, 2.3, 2.7, 2.9, 3.3])
    f32 = lambda a: np.float32(a).astype(float)
    par = [f32(thk), f32(vp), f32(vs), f32(rho)]

    def root(c_guess, Tp, pr, width=3e-4):
        lo, hi = c_guess * (1 - width), c_guess * (1 + width)
        d0 = rayleigh_secular_ocean(lo, Tp, *pr)
        ph = d0 / abs(d0)
        f = lambda x: (rayleigh_secular_ocean(x, Tp, *pr) / ph).real
        flo = f(lo)
        assert np.sign(flo) != np.sign(f(hi)), (c_guess, Tp)
        for _ in range(50):
            mid = 0.5 * (lo + hi)
            fm = f(mid)
            if np.sign(fm) == np.sign(flo):
                lo, flo = mid, fm
            else:
                hi = mid
        return 0.5 * (lo + hi)
    T = np.array([6., 10., 15., 25., 40.])
    c, da, db, dr, dh, ok = oracle.surf_adjoint_kernel(thk, vp, vs, rho, T, "Rc")
    u, ok2 = oracle.surf_forward(thk, vp, vs, rho, T, "Rg")
    assert ok and ok2
    for i, Tp in enumerate(T):
        c0 = root(c[i], Tp, par)
        assert abs(c0 - c[i]) < 1.5e-6 * c0, Tp
        e = 1e-4 * Tp
        dcdT = (root(c0, Tp + e, par) - root(c0, Tp - e, par)) / (2 * e)
        assert abs(u[i] - c0 / (1 + Tp / c0 * dcdT)) < 3e-5 * u[i], Tp
    i, Tp = 2, 15.0
    c0 = root(c[i], Tp, par)
    big = max(np.max(np.abs(k_[i])) for k_ in (da, db, dr, dh))
    for which, arr, layers in ((1, da, range(5)), (2, db, range(1, 5)), (3, dr, range(5)), (0, dh, range(4))):
        for m in layers:
            h = 1e-5 * max(par[which][m], 1.0)
            pp = [a.copy() for a in par]
            pm = [a.copy() for a in par]
            pp[which][m] += h
            pm[which][m] -= h
            fd = (root(c0, Tp, pp) - root(c0, Tp, pm)) / (2 * h)
            assert abs(arr[i, m] - fd) < 3e-4 * abs(fd) + 1e-5 * big, (which, m, arr[i, m], fd)


def test_water_layer_oracle_identities(oracle):
    """The fluid branches of the restated sregn96 satisfy the scaling identities; Love ignores water."""
    thk = np.array([3.0, 2.0, 5.0, 12.0, 0.0])
    vs = np.array([0.0, 2.2, 3.3, 3.9, 4.6])
    vp = np.array([1.5, 4.2, 5.9, 6.8, 8.1])
    rho = np.array([1.03, 2.3, 2.7, 2.9, 3.3])
    T = np.array([6., 10., 15., 25., 40.])
    c, da, db, dr, dh, ok = oracle.surf_adjoint_kernel(thk, vp, vs, rho, T, "Rc")
    assert ok and np.all(c > 1.0)
    v32 = lambda a: a.astype(np.float32).astype(np.float64)
    euler = (da * v32(vp)).sum(1) + (db * v32(vs)).sum(1) + (dh * v32(thk)).sum(1)
    assert np.max(np.abs(euler - c) / c) < 1e-4
    assert np.max(np.abs((dr * v32(rho)).sum(1))) < 3e-4
    # Love waves do not see the water: same as the model without the top layer
    cl, okl = oracle.surf_forward(thk, vp, vs, rho, T, "Lc")
    cs, oks = oracle.surf_forward(thk[1:], vp[1:], vs[1:], rho[1:], T, "Lc")
    assert okl and oks and np.allclose(cl, cs, rtol=2e-6)
