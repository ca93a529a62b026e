"""ORACLE — TEST INFRASTRUCTURE ONLY.  This file IS pinned against the reference: the reference's own
samplers, imported unmodified, reproduce its decisions and states (tests/golden/reference_code.npz,
tests/golden/make_reference_golden.py); the forward model underneath stays PARITY UNPINNED
(oracle/README.md).

NumPy restatement of the reference samplers, one chain at a time, driven by any
`misfit_and_grad(x) -> (U, grad, dsyn, flag)` callable (in the tests: the C++ oracle):

  /root/reference/pyhmc/hmc.py    set_initial_model :74-93, _check_init_is_in_boundary :95-99,
                                  _mirror :121-137, _leapfrog :140-201, sample :228-263
  /root/reference/pyhmc/hmcda.py  _find_initial_dt :170-220, _leapfrog :222-278, sample :280-355

It uses NumPy's own legacy global RandomState (np.random.seed / rand / randn / randint), i.e. the
very stream the reference draws from, so it pins the device MT19937 + Box-Muller + randint
implementation as well as the accept/reject logic.  Unlike the reference it returns the per-iteration
accept flags and stops after `max_iters` trajectories.
"""
import math
import numpy as np


class ChainResult:
    def __init__(self):
        self.accepts = []
        # per-trajectory trace (for the fixtures made from the reference's own Python):
        # L, dt used, alpha (dual averaging), x returned by _leapfrog
        self.trace_L, self.trace_dt, self.trace_alpha, self.trace_x = [], [], [], []
        self.misfit = None
        self.samples = None
        self.syn = None
        self.initmodel = None
        self.dt = None
        self.n_acc = 0
        self.n_iter = 0


def _initial_model(bounds, rs):
    n2 = bounds.shape[0]
    n = n2 // 2
    while True:
        x = np.zeros(n2)
        for i in range(n2):
            x[i] = bounds[i, 0] + (bounds[i, 1] - bounds[i, 0]) * rs.rand()
        idx = np.argsort(x[:n])
        x[:n] = x[:n][idx]
        x[n:] = x[n:][idx]
        ok = all(not (x[i] < bounds[i, 0] or x[i] > bounds[i, 1]) for i in range(n2 - 1))
        if ok:
            return x


def _mirror(x, p, bounds):
    x = x.copy()
    p = p.copy()
    lo, hi = bounds[:, 0], bounds[:, 1]
    i1 = x > hi
    i2 = x < lo
    it = 0
    while np.sum(np.logical_or(i1, i2)) > 0:
        x[i1] = 2 * hi[i1] - x[i1]
        p[i1] = -p[i1]
        x[i2] = 2 * lo[i2] - x[i2]
        p[i2] = -p[i2]
        i1 = x > hi
        i2 = x < lo
        it += 1
        if it >= 64:
            # the reference keeps bouncing (|x|/(hi-lo) passes, forever for +-inf); fold the
            # remainder in closed form exactly as rfsurfhmc_b200/csrc/hmc_kernels.cuh does
            for j in np.nonzero(np.logical_or(i1, i2))[0]:
                w = hi[j] - lo[j]
                u = x[j] - lo[j]
                if u < 0.0:
                    u = -u
                    p[j] = -p[j]
                y = math.fmod(u, 2.0 * w) if np.isfinite(u) and w != 0.0 else np.nan
                if y > w:
                    y = 2.0 * w - y
                    p[j] = -p[j]
                x[j] = lo[j] + y
                if not (lo[j] <= x[j] <= hi[j]):
                    x[j] = np.nan
            break
    return x, p


def _kin(p):
    return np.dot(p, p) * 0.5


def _trajectory(f, xcur, dt, L, bounds, rs):
    """Common body of both _leapfrog variants. Returns (xnew, Unew, dsyn, Hcur, Hnew, failed)."""
    n = len(xcur)
    p = rs.randn(n) * 0.5
    x = xcur * 1.0
    K = _kin(p)
    U, g, d, flag = f(x)
    if flag is False or np.sum(np.isnan(d)) > 0:
        return None
    Hcur = K + U
    p = p - dt * g * 0.5
    for i in range(L):
        x = x + dt * p
        x, p = _mirror(x, p, bounds)
        if np.sum(np.isnan(x)) > 0:
            return None
        U, g, d, flag = f(x)
        if np.sum(np.isnan(g)) > 0:
            return None
        if flag is False or np.sum(np.isnan(d)) > 0:
            return None
        if i < L - 1:
            p = p - dt * g
        else:
            p = p - dt * g * 0.5
    Hnew = _kin(-p) + U
    return x, U, d, Hcur, Hnew


def run_base(f, bounds, dt, Lrange, seed, nsamples, ndraws, max_iters=None):
    """HamitonianMC.sample for one chain whose seed is already `seed + rank`."""
    rs = np.random.RandomState()
    rs.seed(seed)
    R = ChainResult()
    x = _initial_model(bounds, rs)
    R.initmodel = x.copy()
    nd = None
    R.misfit = np.zeros(nsamples)
    R.samples = np.zeros((nsamples, len(x)))
    i = 0
    ncount = 0
    while i < ndraws + nsamples:
        if max_iters is not None and ncount >= max_iters:
            break
        L = rs.randint(Lrange[0], Lrange[1] + 1)
        t = _trajectory(f, x, dt, L, bounds, rs)
        acc = False
        if t is not None:
            xn, Un, dn, Hc, Hn = t
            u = rs.rand()
            if u < np.exp(-(Hn - Hc)):
                acc = True
                x = xn
                if i >= ndraws:
                    if R.syn is None:
                        R.syn = np.zeros((nsamples, len(dn)))
                    R.misfit[i - ndraws] = Un
                    R.samples[i - ndraws] = xn
                    R.syn[i - ndraws] = dn
                i += 1
        R.accepts.append(1 if acc else 0)
        R.trace_L.append(L)
        R.trace_dt.append(dt)
        R.trace_x.append(np.array(x, dtype=float).copy())
        ncount += 1
    R.n_acc = i
    R.n_iter = ncount
    R.dt = dt
    return R


def _find_initial_dt(f, dt0, x, bounds, rs):
    xc = x.copy()
    dt = dt0
    p = rs.randn(len(xc)) * 0.5
    U, g, _, flag = f(xc)
    Hcur = U + _kin(p)
    a = 0.
    p = p - 0.5 * dt * g
    for it in range(20):
        xc = xc + dt * p
        xc, p = _mirror(xc, p, bounds)
        U, g, _, flag = f(xc)
        if not flag:
            raise RuntimeError("error in chain (reference exits here)")
        p = p - 0.5 * dt * g
        Hnew = U + _kin(p)
        ediff = -(Hnew - Hcur)
        if it == 0:
            a = 2 * (ediff > np.log(0.5)) - 1
        if ediff < np.log(0.5):
            break
        p = p - 0.5 * dt * g
        Hcur = Hnew * 1.
        dt = dt * 2**a
    return dt


def run_da(f, bounds, dt_cfg, L0, target, seed, nsamples, ndraws, max_iters=None, max_L=0):
    """HMCDualAveraging.sample for one chain whose seed is already `seed + rank`.
    max_L > 0: the cap on L that rfs_hmc_run offers as an extension (0 = the reference)."""
    rs = np.random.RandomState()
    rs.seed(seed)
    R = ChainResult()
    x = _initial_model(bounds, rs)
    R.initmodel = x.copy()
    R.misfit = np.zeros(nsamples)
    R.samples = np.zeros((nsamples, len(x)))
    lam = L0 * dt_cfg
    gamma, t0, kappa = 0.05, 10., 0.75
    dt = _find_initial_dt(f, dt_cfg, x, bounds, rs)
    dtbar = dt * 1.
    h0 = 0.
    mu = np.log(10 * dt_cfg)
    ncount = 0
    i = 0
    while i < ndraws + nsamples:
        if max_iters is not None and ncount >= max_iters:
            break
        L = max(1, int(lam / dt))
        if max_L > 0:
            L = min(L, max_L)
        t = _trajectory(f, x, dt, L, bounds, rs)
        alpha = 0.
        if t is not None:
            xn, Un, dn, Hc, Hn = t
            alpha = min(1., np.exp(-(Hn - Hc)))
        R.trace_L.append(L)
        R.trace_dt.append(dt)
        R.trace_alpha.append(float(alpha))
        R.trace_x.append(np.array(xn if t is not None else x, dtype=float).copy())
        u = rs.rand()
        acc = False
        if u < alpha:
            acc = True
            x = xn.copy()
            if i >= ndraws:
                if R.syn is None:
                    R.syn = np.zeros((nsamples, len(dn)))
                R.misfit[i - ndraws] = Un
                R.samples[i - ndraws] = xn
                R.syn[i - ndraws] = dn
            i += 1
        if ncount < ndraws:
            m = ncount + 1
            fac = 1. / (m + t0)
            h0 = (1 - fac) * h0 + fac * (target - alpha)
            logdt = mu - np.sqrt(m) / gamma * h0
            dt = np.exp(logdt)
            fac = m**(-kappa)
            logdtbar = fac * logdt + (1 - fac) * np.log(dtbar)
            dtbar = np.exp(logdtbar)
        else:
            dt = dtbar * 1.
        R.accepts.append(1 if acc else 0)
        ncount += 1
    R.n_acc = i
    R.n_iter = ncount
    R.dt = dt
    return R


def oracle_joint_f(O, dobs, cfg, which=0):
    """misfit_and_grad callable backed by the C++ oracle (which: 0 Joint_RF_SWD, 1 ReceiverFunc,
    2 SurfWD semantics)."""
    def f(x):
        U, g, d, fl = O.joint_batch(np.asarray(x)[None, :], dobs, cfg, which=which)
        if not fl[0]:
            return 0.0, np.zeros(len(x)), dobs, False
        return float(U[0]), g[0], d[0], True
    return f
