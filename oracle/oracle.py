"""ORACLE — TEST INFRASTRUCTURE ONLY (PARITY UNPINNED, see oracle/README.md).

ctypes front end of oracle/_build/liboracle*.so exposing look-alikes of the reference's
pybind11 modules (`libsurf`, `librf`: /root/reference/src/SWD/main.cpp:84-94,
/root/reference/src/RF/main.cpp:191-213) plus the batched joint misfit/gradient used as the
checker of the fused GPU path and as bench.py's CPU baseline.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (rfsurfhmc_b200) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_WT = {"Rc": 0, "Rg": 1, "Lc": 2, "Lg": 3}
_PAR = {"rho": 1, "vp": 2, "alpha": 2, "vs": 3, "beta": 3, "h": 4, "thick": 4}
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    """Compile the restatement with g++ (checker build + fast baseline build)."""
    out = os.path.join(_HERE, "_build", "liboracle.so")
    if force or not os.path.exists(out) or not os.path.exists(out.replace(".so", "_fast.so")):
        subprocess.check_call(["make", "-C", _HERE, "-j2"], stdout=subprocess.DEVNULL)
    return out


def _p(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class Oracle:
    def __init__(self, fast=False):
        build()
        name = "liboracle_fast.so" if fast else "liboracle.so"
        self.lib = C.CDLL(os.path.join(_HERE, "_build", name))
        L = self.lib
        L.orc_surf_forward.restype = C.c_int
        L.orc_surf_forward.argtypes = [_dp] * 4 + [C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
        L.orc_surf_kernel.restype = C.c_int
        L.orc_surf_kernel.argtypes = [_dp] * 4 + [C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                                  C.c_int] + [_dp] * 5
        L.orc_surfdisp96_evals.restype = C.c_long
        L.orc_surfdisp96_evals.argtypes = [_dp] * 4 + [C.c_int, _dp, C.c_int, C.c_int, C.c_int]
        rf_common = [_dp] * 6 + [C.c_int, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                                 C.c_int, C.c_double, C.c_int]
        L.orc_rf_forward.restype = None
        L.orc_rf_forward.argtypes = rf_common + [_dp]
        L.orc_rf_kernel.restype = None
        L.orc_rf_kernel.argtypes = rf_common + [C.c_int, _dp, _dp]
        L.orc_rf_kernel_all.restype = None
        L.orc_rf_kernel_all.argtypes = rf_common + [_dp, _dp]
        L.orc_rfft.restype = None
        L.orc_rfft.argtypes = [_dp, _dp, C.c_int]
        L.orc_irfft.restype = None
        L.orc_irfft.argtypes = [_dp, _dp, C.c_int]
        L.orc_joint_batch.restype = None
        L.orc_joint_batch.argtypes = [C.c_int, C.c_int, _dp, _dp,
                                      C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp,
                                      C.c_int, C.c_int,
                                      C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                                      C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                      C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _ip]

    # ---- libsurf look-alike -------------------------------------------------------------
    def surf_forward(self, thk, vp, vs, rho, period, wavetype, mode=0, sphere=False):
        if wavetype not in _WT:
            raise ValueError("wavetype should be one of [Rc,Rg,Lc,Lg]")
        thk, vp, vs, rho, t = map(_f64, (thk, vp, vs, rho, period))
        cg = np.zeros(t.size)
        ok = self.lib.orc_surf_forward(_p(thk), _p(vp), _p(vs), _p(rho), thk.size, _p(t), t.size,
                                       _WT[wavetype], int(mode), int(bool(sphere)), _p(cg))
        return cg, bool(ok == 1)

    def surf_adjoint_kernel(self, thk, vp, vs, rho, period, wavetype, mode=0, sphere=False,
                            stale=True):
        if wavetype not in _WT:
            raise ValueError("wavetype should be one of [Rc,Rg,Lc,Lg]")
        thk, vp, vs, rho, t = map(_f64, (thk, vp, vs, rho, period))
        n, nt = thk.size, t.size
        c = np.zeros(nt)
        da, db, dr, dh = (np.zeros((nt, n)) for _ in range(4))
        ok = self.lib.orc_surf_kernel(_p(thk), _p(vp), _p(vs), _p(rho), n, _p(t), nt,
                                      _WT[wavetype], int(mode), int(bool(sphere)), int(bool(stale)),
                                      _p(c), _p(da), _p(db), _p(dr), _p(dh))
        return c, da, db, dr, dh, bool(ok == 1)

    def surfdisp96_evals(self, thk, vp, vs, rho, period, iwave=2, mode1=1):
        thk, vp, vs, rho, t = map(_f64, (thk, vp, vs, rho, period))
        return int(self.lib.orc_surfdisp96_evals(_p(thk), _p(vp), _p(vs), _p(rho), thk.size, _p(t),
                                                 t.size, iwave, mode1))

    # ---- librf look-alike ---------------------------------------------------------------
    @staticmethod
    def _rft(rf_type):
        if rf_type in ("P", "p"):
            return 1
        if rf_type in ("S", "s"):
            return 2
        raise ValueError("rf_type should be one of [P,p,S,s]")

    def rf_forward(self, thk, rho, vp, vs, qa, qb, ray_p, nt, dt, gauss, time_shift,
                   method="time", water=0.001, rf_type="P"):
        a = list(map(_f64, (thk, rho, vp, vs, qa, qb)))
        rf = np.zeros(nt)
        self.lib.orc_rf_forward(*map(_p, a), a[0].size, ray_p, nt, dt, gauss, time_shift,
                                0 if method == "time" else 1, water, self._rft(rf_type), _p(rf))
        return rf

    def rf_kernel(self, thk, rho, vp, vs, qa, qb, ray_p, nt, dt, gauss, time_shift,
                  method="time", water=0.001, rf_type="P", par_type="vs"):
        if par_type not in _PAR:
            raise ValueError("par_type should be one of [vp,vs,rho,thick]")
        a = list(map(_f64, (thk, rho, vp, vs, qa, qb)))
        n = a[0].size
        rf = np.zeros(nt)
        drf = np.zeros((n, nt))
        self.lib.orc_rf_kernel(*map(_p, a), n, ray_p, nt, dt, gauss, time_shift,
                               0 if method == "time" else 1, water, self._rft(rf_type),
                               _PAR[par_type], _p(rf), _p(drf))
        return rf, drf

    def rf_kernel_all(self, thk, rho, vp, vs, qa, qb, ray_p, nt, dt, gauss, time_shift,
                      method="time", water=0.001, rf_type="P"):
        a = list(map(_f64, (thk, rho, vp, vs, qa, qb)))
        n = a[0].size
        rf = np.zeros(nt)
        drf = np.zeros((4, n, nt))
        self.lib.orc_rf_kernel_all(*map(_p, a), n, ray_p, nt, dt, gauss, time_shift,
                                   0 if method == "time" else 1, water, self._rft(rf_type),
                                   _p(rf), _p(drf))
        return rf, drf

    def rfft(self, x):
        x = _f64(x)
        out = np.zeros(2 * (x.size // 2 + 1))
        self.lib.orc_rfft(_p(x), _p(out), x.size)
        return out[0::2] + 1j * out[1::2]

    def irfft(self, X, n):
        buf = np.zeros(2 * (n // 2 + 1))
        buf[0::2] = X.real
        buf[1::2] = X.imag
        out = np.zeros(n)
        self.lib.orc_irfft(_p(buf), _p(out), n)
        return out

    # ---- batched joint misfit + gradient ------------------------------------------------
    def joint_batch(self, x, dobs, cfg, which=0, nthreads=1):
        """x [B,2n]; cfg = dict(tRc,tRg,tLc,tLg,mode,sphere,ray_p,nt,dt,gauss,time_shift,water,
        rf_type,method,sigma1,sigma2,stale).  which: 0 joint, 1 RF only, 2 SWD only."""
        x = _f64(x)
        B, n2 = x.shape
        n = n2 // 2
        per = [_f64(cfg.get(k, [])) for k in ("tRc", "tRg", "tLc", "tLg")]
        nsw = sum(p.size for p in per)
        nt = int(cfg["nt"])
        nd = nt + nsw if which == 0 else (nt if which == 1 else nsw)
        dobs = _f64(dobs)
        U = np.zeros(B)
        g = np.zeros((B, 2 * n))
        d = np.zeros((B, nd))
        flag = np.zeros(B, dtype=np.int32)
        self.lib.orc_joint_batch(
            B, n, _p(x), _p(dobs),
            per[0].size, _p(per[0]), per[1].size, _p(per[1]), per[2].size, _p(per[2]),
            per[3].size, _p(per[3]),
            int(cfg.get("mode", 0)), int(bool(cfg.get("sphere", False))),
            float(cfg["ray_p"]), nt, float(cfg["dt"]), float(cfg["gauss"]),
            float(cfg["time_shift"]), float(cfg.get("water", 0.001)),
            self._rft(cfg.get("rf_type", "P")), 0 if cfg.get("method", "freq") == "time" else 1,
            float(cfg.get("sigma1", 1.0)), float(cfg.get("sigma2", 1.0)),
            int(bool(cfg.get("stale", True))), which, nthreads,
            _p(U), _p(g), _p(d), flag.ctypes.data_as(_ip))
        return U, g, d, flag.astype(bool)


def brocher(vs):
    """model/model_surf.py:64-67 (vp, rho from vs)."""
    vs = np.asarray(vs, dtype=np.float64)
    vp = 0.9409 + 2.0947 * vs - 0.8206 * vs**2 + 0.2683 * vs**3 - 0.0251 * vs**4
    rho = 1.6612 * vp - 0.4721 * vp**2 + 0.0671 * vp**3 - 0.0043 * vp**4 + 0.000106 * vp**5
    return vp, rho
