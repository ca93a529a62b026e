"""`libsurf` — same names, argument order, defaults and return tuples as the reference's pybind11
module (/root/reference/src/SWD/main.cpp:84-94), computed by the sm_100a kernels through the
C ABI (include/rfsurfhmc.h: rfs_surf_forward / rfs_surf_adjoint_kernel).

Differences by design: an invalid `wavetype` raises ValueError (the reference prints and calls
exit(0), main.cpp:19-25); a fluid layer is accepted only at the top of the stack;
Love `dcda` is returned as zeros (the reference returns uninitialised memory, main.cpp:68)."""
import numpy as np
from ..._lib import default_context, wavetype_code

__doc__ = "Surface wave dispersion and sensivity kernel\n"


def forward(thk, vp, vs, rho, period, wavetype, mode=0, sphere=False):
    """Surface wave dispersion -> (ndarray[nT] float64, bool)."""
    wavetype_code(wavetype)
    c, ok = default_context().surf_forward(thk, vp, vs, rho, period, wavetype, mode, sphere)
    return c[0], bool(ok[0])


def adjoint_kernel(thk, vp, vs, rho, period, wavetype, mode=0, sphere=False):
    """Surface wave dispersion sensitivity kernel ->
    (c[nT], dcda[nT,n], dcdb[nT,n], dcdr[nT,n], dcdh[nT,n], bool)."""
    wavetype_code(wavetype)
    c, da, db, dr, dh, ok = default_context().surf_adjoint_kernel(thk, vp, vs, rho, period, wavetype,
                                                                  mode, sphere)
    return c[0], da[0], db[0], dr[0], dh[0], bool(ok[0])
