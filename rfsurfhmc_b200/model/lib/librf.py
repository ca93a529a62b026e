"""`librf` — same names, argument order, defaults and returns as the reference's pybind11 module
(/root/reference/src/RF/main.cpp:191-213), computed by the sm_100a kernels through the C ABI
(include/rfsurfhmc.h: rfs_rf_forward / rfs_rf_kernel / rfs_rf_kernel_all).

Differences by design: invalid `rf_type` / `par_type` raise ValueError (reference: exit(-1),
main.cpp:36-41,107-111); `kernel_all`'s accidental default rf_type ("P"+docstring, main.cpp:211-212)
is "P"."""
from ..._lib import default_context, rf_type_code, PARTYPES

__doc__ = "Receiver function and partial derivative\n"


def forward(thk, rho, vp, vs, qa, qb, ray_p, nt, dt, gauss, time_shift, method="time", water=0.001,
            rf_type="P"):
    """receiver function -> ndarray[nt]."""
    rf_type_code(rf_type)
    return default_context().rf_forward(thk, rho, vp, vs, qa, qb, ray_p, nt, dt, gauss, time_shift,
                                        method, water, rf_type)[0]


def kernel(thk, rho, vp, vs, qa, qb, ray_p, nt, dt, gauss, time_shift, method="time", water=0.001,
           rf_type="P", par_type="vs"):
    """receiver function and its kernel for one parameter type -> (rf[nt], drf[n,nt])."""
    rf_type_code(rf_type)
    if par_type not in PARTYPES:
        raise ValueError("par_type should be one of [vp,vs,rho,thick]")
    rf, drf = default_context().rf_kernel(thk, rho, vp, vs, qa, qb, ray_p, nt, dt, gauss, time_shift,
                                          method, water, rf_type, par_type)
    return rf[0], drf[0]


def kernel_all(thk, rho, vp, vs, qa, qb, ray_p, nt, dt, gauss, time_shift, method="time", water=0.001,
               rf_type="P"):
    """receiver function and all kernels -> (rf[nt], drf[4,n,nt]) in the order rho, vp, vs, h."""
    rf_type_code(rf_type)
    rf, drf = default_context().rf_kernel_all(thk, rho, vp, vs, qa, qb, ray_p, nt, dt, gauss,
                                              time_shift, method, water, rf_type)
    return rf[0], drf[0]
