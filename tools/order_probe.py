"""Probe: does the order of the models inside a batch matter?  Lanes of a warp hold consecutive
models; similar models take the same branches (oscillatory/evanescent layers, scan/refine phases)
and need similar numbers of secular evaluations.

    python tools/order_probe.py [--chains 16384]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rfsurfhmc_b200._lib import Context
from bench import make_dobs_gpu, workload, N_LAYERS

ap = argparse.ArgumentParser()
ap.add_argument("--chains", type=int, default=16384)
a = ap.parse_args()
dev = torch.device("cuda", 0)
cfg, x0, X = workload(a.chains, 1000)
ctx = Context(0)
ctx.config_swd(N_LAYERS, tRc=cfg["tRc"], tRg=cfg["tRg"])
ctx.config_rf(N_LAYERS, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"],
              cfg["rf_type"], cfg["method"])
dobs = make_dobs_gpu(ctx, cfg, x0)
ctx.config_obs(dobs)
B = X.shape[0]
nd = dobs.size
U = torch.empty(B, dtype=torch.float64, device=dev)
G = torch.empty(B, 2 * N_LAYERS, dtype=torch.float64, device=dev)
D = torch.empty(B, nd, dtype=torch.float64, device=dev)
F = torch.empty(B, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream


def rate(Xo, label):
    xd = torch.from_numpy(np.ascontiguousarray(Xo)).to(dev)
    for _ in range(2):
        ctx.misfit_grad_dev(B, xd.data_ptr(), 0, U.data_ptr(), G.data_ptr(), D.data_ptr(), F.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ctx.misfit_grad_dev(B, xd.data_ptr(), 0, U.data_ptr(), G.data_ptr(), D.data_ptr(), F.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-44s %7.3f ms/step  %8.0f evals/s" % (label, ms, B / ms * 1e3))


rate(X, "as generated")
rate(X[np.random.default_rng(0).permutation(B)], "random permutation")
n = N_LAYERS
rate(X[np.argsort(X[:, :n].mean(1))], "sorted by mean vs")
rate(X[np.argsort(X[:, 0])], "sorted by vs of the top layer")
w = X[:, n:2 * n - 1]
rate(X[np.argsort((X[:, :n - 1] * w).sum(1) / w.sum(1))], "sorted by thickness-weighted mean vs")
_, _, d, f = ctx.misfit_grad_host(X)
rate(X[np.argsort(d[:, cfg["nt"] + 15])], "sorted by c(T=20 s) (needs a first evaluation)")
rate(X[np.lexsort((d[:, cfg["nt"] + 35], d[:, cfg["nt"]]))], "sorted by c(5 s) then c(40 s)")
