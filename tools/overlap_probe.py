"""Probe: do two half-size batches on two contexts/streams, skewed by half a step, beat one full
batch?  (Would hide the eigen/assemble kernels of one half behind the root search of the other.)

    python tools/overlap_probe.py [--chains 16384] [--steps 20]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rfsurfhmc_b200._lib import Context
from bench import make_dobs_gpu, workload, N_LAYERS


def setup(ctx, cfg, dobs):
    ctx.config_swd(N_LAYERS, tRc=cfg["tRc"], tRg=cfg["tRg"])
    ctx.config_rf(N_LAYERS, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"],
                  cfg["rf_type"], cfg["method"])
    ctx.config_obs(dobs)


class Lane:
    def __init__(self, cfg, dobs, X, dev):
        self.ctx = Context(0)
        setup(self.ctx, cfg, dobs)
        self.B = X.shape[0]
        self.x = torch.from_numpy(X).to(dev)
        nd = dobs.size
        self.U = torch.empty(self.B, dtype=torch.float64, device=dev)
        self.G = torch.empty(self.B, 2 * N_LAYERS, dtype=torch.float64, device=dev)
        self.D = torch.empty(self.B, nd, dtype=torch.float64, device=dev)
        self.F = torch.empty(self.B, dtype=torch.uint8, device=dev)
        self.s = torch.cuda.Stream()

    def step(self):
        self.ctx.misfit_grad_dev(self.B, self.x.data_ptr(), 0, self.U.data_ptr(), self.G.data_ptr(),
                                 self.D.data_ptr(), self.F.data_ptr(), self.s.cuda_stream)


def run(lanes, steps, skew):
    torch.cuda.synchronize()
    for l in lanes:
        l.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if skew and len(lanes) > 1:
        lanes[0].step()
        time.sleep(skew)
    for i in range(steps):
        for l in lanes:
            l.step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = sum(l.B for l in lanes) * steps + (lanes[0].B if skew and len(lanes) > 1 else 0)
    return n / dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg, x0, X = workload(a.chains, 1000)
    c0 = Context(0)
    setup(c0, cfg, np.zeros(cfg["nt"] + 72))
    dobs = make_dobs_gpu(c0, cfg, x0)
    full = [Lane(cfg, dobs, X, dev)]
    print("1 lane  x %6d : %.0f evals/s" % (a.chains, run(full, a.steps, 0)))
    for k in (2, 4):
        parts = np.array_split(X, k)
        lanes = [Lane(cfg, dobs, p, dev) for p in parts]
        print("%d lanes x %6d, no skew   : %.0f evals/s" % (k, parts[0].shape[0], run(lanes, a.steps, 0)))
        print("%d lanes x %6d, skew 3 ms : %.0f evals/s" % (k, parts[0].shape[0], run(lanes, a.steps, 0.003)))


if __name__ == "__main__":
    main()
