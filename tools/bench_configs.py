"""Secondary workloads of BASELINE.json (configs[1], configs[2]) through the fused device path.

    python tools/bench_configs.py [--batch 65536]

C2: SWD-only forward+Frechet objective, n=40 layers, 60 periods, Rc+Rg+Lc+Lg, modes 0,1,2
    (three calls, as the reference needs one call per mode).
C3: RF-only forward+Frechet objective, nt=2048, dt=0.05, Gaussian a=2.5, water-level, 3 ray parameters.
Synthetic models per SURVEY.md §8d.  Prints one JSON line per config (models/s, device-resident)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rfsurfhmc_b200._lib import Context


def c2_models(B, seed=2, n=40):
    rng = np.random.default_rng(seed)
    i = np.arange(n - 1)
    thk = np.hstack((0.5 + 0.1 * i, [0.0]))[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (B, n)))
    thk[:, -1] = 0.0
    vs0 = 2.0 + 2.7 * (np.arange(n) / (n - 1.0))**0.7
    vs = np.clip(vs0[None, :] * (1 + 0.04 * rng.standard_normal((B, n))), 1.5, 5.0)
    return np.hstack((vs, thk))


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--cpu", action="store_true", help="also time the oracle on the host cores (bounded sample)")
    a = ap.parse_args()
    B, n = a.batch, 40
    dev = torch.device("cuda", 0)
    X = c2_models(B)
    xd = torch.from_numpy(X).to(dev)
    st = torch.cuda.current_stream().cuda_stream
    T = np.geomspace(2, 100, 60)
    ctx = Context(0)
    U = torch.empty(B, dtype=torch.float64, device=dev)
    G = torch.empty(B, 2 * n, dtype=torch.float64, device=dev)
    Fl = torch.empty(B, dtype=torch.uint8, device=dev)
    # ---- C2
    nd = 240
    D = torch.empty(B, nd, dtype=torch.float64, device=dev)
    tot = 0.0
    fails = []
    for mode in (0, 1, 2):
        ctx.config_swd(n, T, T, T, T, mode=mode)
        ctx.config_obs(np.full(nd, 3.0))
        t = timed(lambda: ctx.misfit_grad_dev(B, xd.data_ptr(), 2, U.data_ptr(), G.data_ptr(), D.data_ptr(),
                                              Fl.data_ptr(), st), reps=1)
        tot += t
        fails.append(int((Fl == 0).sum().item()))
        zeros = float((D == 0).float().mean().item())
        print(f"# C2 mode {mode}: {t*1e3:.1f} ms, failed models {fails[-1]}, zeroed (missing-mode) data fraction {zeros:.4f}",
              file=sys.stderr)
    print(json.dumps({"config": "C2 SWD-only forward+Frechet: n=40, 60 periods, Rc+Rg+Lc+Lg, modes 0-2 (3 calls)",
                      "batch": B, "models_per_s": B / tot, "seconds": tot, "failed_models_per_mode": fails}))
    # ---- C3
    nt = 2048
    D3 = torch.empty(B, nt, dtype=torch.float64, device=dev)
    tot = 0.0
    for p in (0.04, 0.06, 0.08):
        ctx.config_rf(n, p, nt, 0.05, 2.5, 5.0, 1e-3, "P", "freq")
        ctx.config_obs(np.zeros(nt))
        t = timed(lambda: ctx.misfit_grad_dev(B, xd.data_ptr(), 1, U.data_ptr(), G.data_ptr(), D3.data_ptr(),
                                              Fl.data_ptr(), st), reps=1)
        tot += t
        print(f"# C3 p={p}: {t*1e3:.1f} ms", file=sys.stderr)
    print(json.dumps({"config": "C3 RF-only forward+Frechet (freq): n=40, nt=2048, a=2.5, 3 ray parameters (3 calls)",
                      "batch": B, "models_per_s": B / tot, "seconds": tot}))
    # ---- C3, time-domain method (iterative deconvolution of the forward trace and of all 4n Frechet
    # traces, <= 200 iterations each): a bounded sub-batch, one ray parameter
    Bt = min(B, 256)
    ctx.config_rf(n, 0.06, nt, 0.05, 2.5, 5.0, 1e-3, "P", "time")
    ctx.config_obs(np.zeros(nt))
    t = timed(lambda: ctx.misfit_grad_dev(Bt, xd.data_ptr(), 1, U.data_ptr(), G.data_ptr(), D3.data_ptr(),
                                          Fl.data_ptr(), st), reps=1)
    print(json.dumps({"config": "C3 RF-only forward+Frechet (time-domain deconvolution): n=40, nt=2048, 1 ray parameter",
                      "batch": Bt, "models_per_s": Bt / t, "seconds": t}))


def cpu_leg():
    from oracle.oracle import Oracle
    O = Oracle(fast=True)
    nth = os.cpu_count() or 1
    n = 40
    X = c2_models(4 * nth)
    T = np.geomspace(2, 100, 60)
    base = dict(ray_p=0.06, nt=2048, dt=0.05, gauss=2.5, time_shift=5.0, water=1e-3, rf_type="P", method="freq",
                tRc=T, tRg=T, tLc=T, tLg=T)
    t0 = time.perf_counter()
    for mode in (0, 1, 2):
        O.joint_batch(X, np.full(240, 3.0), dict(base, mode=mode), which=2, nthreads=nth)
    t2 = time.perf_counter() - t0
    print(json.dumps({"config": "C2 (CPU oracle port)", "cores": nth, "models_per_s": len(X) / t2,
                      "sample": "%d models" % len(X)}))
    Xr = X[:max(2, nth // 2)]
    t0 = time.perf_counter()
    for p in (0.04, 0.06, 0.08):
        O.joint_batch(Xr, np.zeros(2048), dict(base, ray_p=p), which=1, nthreads=nth)
    t3 = time.perf_counter() - t0
    print(json.dumps({"config": "C3 (CPU oracle port, O(n^2) reference algorithm)", "cores": nth,
                      "models_per_s": len(Xr) / t3, "sample": "%d models" % len(Xr)}))


def c5_leg(B):
    """C5: fine parameterisation joint evaluation (n=200, 128 Rc + 128 Rg periods, nt=4096)."""
    n = 200
    rng = np.random.default_rng(5)
    thk = 0.4 * (1 + 0.1 * rng.uniform(-1, 1, (B, n)))
    thk[:, -1] = 0.0
    vs0 = 2.0 + 2.7 * (np.arange(n) / (n - 1.0))**0.7
    vs = np.clip(vs0[None, :] * (1 + 0.04 * rng.standard_normal((B, n))), 1.5, 5.0)
    X = np.hstack((vs, thk))
    dev = torch.device("cuda", 0)
    xd = torch.from_numpy(X).to(dev)
    T = np.geomspace(1, 150, 128)
    nt = 4096
    ctx = Context(0)
    ctx.config_swd(n, T, T)
    ctx.config_rf(n, 0.06, nt, 0.025, 2.5, 5.0, 1e-3, "P", "freq")
    nd = nt + 256
    ctx.config_obs(np.hstack((np.zeros(nt), np.full(256, 3.3))))
    U = torch.empty(B, dtype=torch.float64, device=dev)
    G = torch.empty(B, 2 * n, dtype=torch.float64, device=dev)
    D = torch.empty(B, nd, dtype=torch.float64, device=dev)
    Fl = torch.empty(B, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    t = timed(lambda: ctx.misfit_grad_dev(B, xd.data_ptr(), 0, U.data_ptr(), G.data_ptr(), D.data_ptr(),
                                          Fl.data_ptr(), st), reps=1)
    print(json.dumps({"config": "C5 joint evaluation: n=200, 128 Rc + 128 Rg periods, nt=4096", "batch": B,
                      "evals_per_s": B / t, "seconds": t, "failed": int((Fl == 0).sum().item()),
                      "finite_grad_fraction": float(torch.isfinite(G).all(dim=1).float().mean().item())}))


if __name__ == "__main__":
    if "--c5" in sys.argv:
        c5_leg(int(sys.argv[sys.argv.index("--c5") + 1]))
        sys.exit(0)
    if "--cpu" in sys.argv:
        cpu_leg()
    main()
