"""ReceiverFunc — mirror of /root/reference/model/model_rf.py (class ReceiverFunc :4-197).

Same constructor / `init(**kargs)` / `set_obsdata` / `set_thk` / `empirical_relation` / `forward` /
`misfit` / `misfit_and_grad` (3-tuple, no flag, as the reference).  `misfit_and_grad` runs the fused
CUDA path (rfs_misfit_grad_host, which=1): propagator + spectral deconvolution + adjoint gradient."""
import numpy as np
from .lib import librf
from .._lib import Context


class ReceiverFunc:
    def __init__(self, ray_p, nt, dt, gauss, time_shift, water_level=0.001, type_="p", method="time"):
        self.ray_p = ray_p
        self.nt = nt
        self.dt = dt
        self.gauss = gauss
        self.time_shift = time_shift
        self.water_level = water_level
        self.rf_type = type_
        self.t = np.arange(nt) * dt - time_shift
        self.method = method
        self._ctx = None
        self._ctx_n = None

    @classmethod
    def init(self, **kargs):
        return ReceiverFunc(kargs['ray_p'], kargs['nt'], kargs['dt'], kargs['gauss'], kargs['time_shift'],
                            kargs['water_level'], kargs['type'], kargs['method'])

    def set_obsdata(self, dobs):
        self.dobs = dobs
        self._ctx_n = None

    def set_thk(self, thk):
        self.thk = thk * 1.0

    def empirical_relation(self, vs, deriv=False):
        vp = 0.9409 + 2.0947 * vs - 0.8206 * vs**2 + 0.2683 * vs**3 - 0.0251 * vs**4
        rho = 1.6612 * vp - 0.4721 * vp**2 + 0.0671 * vp**3 - 0.0043 * vp**4 + 0.000106 * vp**5
        if deriv:
            drda = 1.6612 - 0.4721 * 2 * vp + 0.0671 * 3 * vp**2 - 0.0043 * 4 * vp**3 + 0.000106 * 5 * vp**4
            dadb = 2.0947 - 0.8206 * 2 * vs + 0.2683 * 3 * vs**2 - 0.0251 * 4 * vs**3
            return vp, rho, drda, dadb
        return vp, rho

    def _context(self, n):
        if self._ctx is None:
            self._ctx = Context(0)
        if self._ctx_n != n:
            self._ctx.config_rf(n, self.ray_p, self.nt, self.dt, self.gauss, self.time_shift,
                                self.water_level, self.rf_type, self.method)
            self._ctx.config_obs(self.dobs)
            self._ctx_n = n
        return self._ctx

    def forward(self, x: np.ndarray):
        layers = int(len(x) / 2)
        vs = x[:layers]
        thk = x[layers:]
        vp, rho = self.empirical_relation(vs, deriv=False)
        qa = thk * 0 + 9999.
        qb = thk * 0 + 9999.
        return librf.forward(thk, rho, vp, vs, qa, qb, self.ray_p, self.nt, self.dt, self.gauss,
                             self.time_shift, self.method, self.water_level, self.rf_type)

    def misfit(self, x):
        d = self.forward(x)
        return 0.5 * np.sum((d - self.dobs)**2)

    def misfit_and_grad(self, x):
        x = np.asarray(x, dtype=np.float64)
        n = int(x.shape[0] / 2)
        U, g, d, _ = self._context(n).misfit_grad_host(x[None, :], which=1)
        return float(U[0]), g[0], d[0]
