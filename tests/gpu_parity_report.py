"""Print GPU-vs-oracle error statistics (run on a GPU box: `python tests/gpu_parity_report.py`)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import Oracle, brocher
from rfsurfhmc_b200._lib import Context

O = Oracle()
ctx = Context(0)
thk = np.array([6, 6, 13, 5, 10, 30, 0.]); vs = np.array([3.2, 2.8, 3.46, 3.3, 3.9, 4.5, 4.7])
vp, rho = brocher(vs)
T = np.arange(5, 41.)


def rel(a, b):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


def relk(a, b):
    s = np.max(np.abs(b))
    return np.max(np.abs(a - b)) / s

for wt in ["Rc", "Rg", "Lc", "Lg"]:
    for mode in [0, 1]:
        c0, ok0 = O.surf_forward(thk, vp, vs, rho, T, wt, mode=mode)
        c1, ok1 = ctx.surf_forward(thk, vp, vs, rho, T, wt, mode=mode)
        print(f"forward {wt} mode{mode}: ok {ok0} {ok1[0]}  maxrel {rel(c1[0][c0>0], c0[c0>0]):.3e} zeros {np.sum(c0==0)} {np.sum(c1[0]==0)}")
        r0 = O.surf_adjoint_kernel(thk, vp, vs, rho, T, wt, mode=mode)
        r1 = ctx.surf_adjoint_kernel(thk, vp, vs, rho, T, wt, mode=mode)
        m = r0[0] > 0
        msg = f"kernel  {wt} mode{mode}: c {rel(r1[0][0][m], r0[0][m]):.3e}"
        for i, nm in zip(range(1, 5), ["da", "db", "dr", "dh"]):
            if np.max(np.abs(r0[i][m])) > 0:
                msg += f" {nm} {relk(r1[i][0][m], r0[i][m]):.3e}"
        print(msg)

q = thk * 0 + 9999.
args = dict(ray_p=0.045, nt=125, dt=0.4, gauss=1.5, time_shift=5., method="freq", water=0.001, rf_type="P")
rf0, kl0 = O.rf_kernel_all(thk, rho, vp, vs, q, q, **args)
rf1, kl1 = ctx.rf_kernel_all(thk, rho, vp, vs, q, q, **args)
print("rf fwd maxabs/peak", np.max(np.abs(rf1[0] - rf0)) / np.max(np.abs(rf0)))
for i, nm in enumerate(["rho", "vp", "vs", "h"]):
    print("rf kern", nm, np.max(np.abs(kl1[0][i] - kl0[i])) / np.max(np.abs(kl0[i])))
rff = ctx.rf_forward(thk, rho, vp, vs, q, q, **args)
print("rf forward-only", np.max(np.abs(rff[0] - rf0)))
args["rf_type"] = "S"
rf0, kl0 = O.rf_kernel_all(thk, rho, vp, vs, q, q, **args)
rf1, kl1 = ctx.rf_kernel_all(thk, rho, vp, vs, q, q, **args)
print("S rf fwd", np.max(np.abs(rf1[0] - rf0)) / np.max(np.abs(rf0)), "kern", np.max(np.abs(kl1[0] - kl0)) / np.max(np.abs(kl0)))

# ---- fused joint path on random models around F1
rng = np.random.default_rng(1)
B = 512
x0 = np.hstack((vs, thk))
lo = np.hstack((np.maximum(vs * 0.2, 1.5), thk * 0.8)); hi = np.hstack((np.minimum(vs * 1.8, 5.0), thk * 1.2)); hi[-1] = 2.0
X = lo + (hi - lo) * rng.random((B, 14))
X[0] = x0
cfg = dict(tRc=T, tRg=T, tLc=[], tLg=[], mode=0, sphere=False, ray_p=0.045, nt=125, dt=0.4, gauss=1.5,
           time_shift=5., water=0.001, rf_type="P", method="freq", sigma1=1., sigma2=1., stale=True)
_, _, dobs, _ = O.joint_batch(x0[None, :], np.zeros(197), cfg)
dobs = dobs[0]
ctx.config_swd(7, tRc=T, tRg=T)
ctx.config_rf(7, 0.045, 125, 0.4, 1.5, 5.0, 0.001, "P", "freq")
ctx.config_obs(dobs)
t0 = time.time(); U0, g0, d0, f0 = O.joint_batch(X, dobs, cfg, nthreads=8); t1 = time.time()
U1, g1, d1, f1 = ctx.misfit_grad_host(X); t2 = time.time()
U1, g1, d1, f1 = ctx.misfit_grad_host(X); t3 = time.time()
print(f"oracle {B/(t1-t0):.1f} eval/s (8 thr)   gpu first {B/(t2-t1):.1f}  second {B/(t3-t2):.1f} eval/s")
print("flags equal", np.array_equal(f0, f1), "nfail", np.sum(~f0))
m = f0 & f1
m[0] = False  # X[0] is the true model: U = 0 and grad = 0 exactly, relative errors are undefined there
print("U rel", rel(U1[m], U0[m]), " U[0] (true model)", U0[0], U1[0])
print("dsyn rf abs/peak", np.max(np.abs(d1[m, :125] - d0[m, :125])) / np.max(np.abs(d0[m, :125])))
print("dsyn swd rel", rel(d1[m, 125:], d0[m, 125:]))
gs = np.max(np.abs(g0[m]), axis=1, keepdims=True)
e = np.abs(g1[m] - g0[m]) / gs
print("grad err / max|grad| : max", e.max(), "median", np.median(e.max(1)), "frac>1e-4", np.mean(e.max(1) > 1e-4))
bad = np.argsort(-e.max(1))[:3]
for b in bad:
    print("  worst", b, e[b].max(), "dsyn swd rel", rel(d1[m][b, 125:], d0[m][b, 125:]))
for which in (1, 2):
    dd = dobs[:125] if which == 1 else dobs[125:]
    ctx.config_obs(dd)
    Ua, ga, da_, fa = O.joint_batch(X, dd, cfg, which=which, nthreads=8)
    Ub, gb, db_, fb = ctx.misfit_grad_host(X, which=which)
    mm = fa & fb
    mm[0] = False
    print("which", which, "flags", np.array_equal(fa, fb), "U", rel(Ub[mm], Ua[mm]), "g",
          np.max(np.abs(gb[mm] - ga[mm]) / np.max(np.abs(ga[mm]), axis=1, keepdims=True)))
print("launches", ctx.launches)

# ---- wild models (velocity inversions, thin layers): error quantiles instead of maxima
ctx.config_obs(dobs)
Xw = np.random.default_rng(13).uniform(0.5, 1.5, (B, 14)) * x0 + 0.01
U0, g0, d0, f0 = O.joint_batch(Xw, dobs, cfg, nthreads=8)
U1, g1, d1, f1 = ctx.misfit_grad_host(Xw)
m = f0 & f1
ed = np.abs(d1[m, 125:] - d0[m, 125:]) / np.abs(d0[m, 125:])
eg = (np.abs(g1[m] - g0[m]) / np.max(np.abs(g0[m]), axis=1, keepdims=True)).max(1)
print("wild: flags equal", np.array_equal(f0, f1), "ok", int(m.sum()), "of", B)
print("wild: dsyn swd rel  max %.3e  99%% %.3e  median %.3e" % (ed.max(), np.quantile(ed.max(1), 0.99), np.median(ed.max(1))))
print("wild: grad err/max|g| max %.3e  99%% %.3e  median %.3e  frac>1e-4 %.4f" %
      (eg.max(), np.quantile(eg, 0.99), np.median(eg), np.mean(eg > 1e-4)))
