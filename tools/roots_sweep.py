"""Root-search mapping sweep on the GPU: thread-mapped kernel vs every team shape, by batch size.

    python tools/roots_sweep.py [--out gpurun_out/roots_sweep.json] [--quick]

For each workload size and each mapping: milliseconds of the root-search kernel (CUDA events around the
launch, rfs_profile_eval) and of the whole evaluation.  The thresholds of pick_team() in
csrc/capi.cu come from this table (kept under profiles/)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rfsurfhmc_b200._lib import Context
from rfsurfhmc_b200.fixtures import f1_config, f1_true_model, driver_bounds, sorted_uniform_models

# (T, S); T = 0: thread-mapped, S = 1 with the length-sorted job order forced on, S = 0 forced off
TEAMS = [(0, 0), (0, 1), (2, 2), (4, 1), (4, 4), (8, 1), (8, 2), (8, 8), (16, 1), (16, 2), (32, 1), (32, 2), (32, 4)]


def layered(B, n, seed):
    rng = np.random.default_rng(seed)
    i = np.arange(n - 1)
    thk = np.hstack((20.0 / n + 30.0 / n * i / n, [0.0]))[None, :] * (1 + 0.1 * rng.uniform(-1, 1, (B, n)))
    thk[:, -1] = 0.0
    vs0 = 2.0 + 2.7 * (np.arange(n) / (n - 1.0))**0.7
    vs = np.clip(vs0[None, :] * (1 + 0.04 * rng.standard_normal((B, n))), 1.5, 5.0)
    return np.hstack((vs, thk))


def run(ctx, X, which, nd, reps):
    dev = torch.device("cuda", 0)
    B, n2 = X.shape
    xd = torch.from_numpy(X).to(dev)
    U = torch.empty(B, dtype=torch.float64, device=dev)
    G = torch.empty(B, n2, dtype=torch.float64, device=dev)
    D = torch.empty(B, nd, dtype=torch.float64, device=dev)
    F = torch.empty(B, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    ref = None
    for T, S in TEAMS:
        ctx.set_roots_team(T, max(S, 1))
        ctx.set_roots_sched(S if T == 0 else -1)
        try:
            ctx.profile_eval(B, xd.data_ptr(), which, U.data_ptr(), G.data_ptr(), D.data_ptr(), F.data_ptr(), st)
        except Exception as e:  # shape not available for this layer count (shared memory)
            rows.append({"T": T, "S": S, "error": str(e)})
            continue
        ms_r, ms_t = [], []
        for _ in range(reps):
            p = ctx.profile_eval(B, xd.data_ptr(), which, U.data_ptr(), G.data_ptr(), D.data_ptr(), F.data_ptr(), st)
            ms_r.append(p["swd_roots"][0])
            ms_t.append(sum(v[0] for v in p.values()))
        torch.cuda.synchronize()
        out = (U.cpu().numpy().copy(), G.cpu().numpy().copy(), D.cpu().numpy().copy())
        same = True
        if ref is None:
            ref = out
        else:
            same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(ref, out))
        rows.append({"T": T, "S": S, "roots_ms": float(np.median(ms_r)), "eval_ms": float(np.median(ms_t)),
                     "bit_identical_to_thread": bool(same)})
    ctx.set_roots_team(-1)
    ctx.set_roots_sched(-1)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/roots_sweep.json")
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    res = []
    ctx = Context(0)
    # ---- C1/C4 sizes
    cfg, x0 = f1_config(), f1_true_model()
    ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"])
    ctx.config_rf(7, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"],
                  cfg["rf_type"], cfg["method"])
    ctx.config_obs(np.full(197, 3.0))
    sizes = [64, 2048, 16384] if a.quick else [64, 256, 1024, 2048, 4096, 8192, 16384, 32768]
    for B in sizes:
        X = sorted_uniform_models(driver_bounds(x0), B, seed=100 + B)
        rows = run(ctx, X, 0, 197, 3)
        res.append({"workload": "C1 joint n=7, 36 Rc + 36 Rg (3 sequences/model)", "B": B, "rows": rows})
        print(json.dumps(res[-1]), flush=True)
    # ---- C2-like: n=40, 60 periods, four wave types, fundamental mode (6 sequences/model)
    Tp = np.geomspace(2, 100, 60)
    ctx2 = Context(0)
    ctx2.config_swd(40, Tp, Tp, Tp, Tp, mode=0)
    ctx2.config_obs(np.full(240, 3.0))
    for B in ([256] if a.quick else [64, 512, 4096]):
        rows = run(ctx2, layered(B, 40, 7), 2, 240, 2)
        res.append({"workload": "C2-like SWD n=40, 60 periods x Rc,Rg,Lc,Lg (6 sequences/model)", "B": B, "rows": rows})
        print(json.dumps(res[-1]), flush=True)
    # ---- C5-like: n=200, 128 Rc + 128 Rg
    Tp = np.geomspace(1, 150, 128)
    ctx3 = Context(0)
    ctx3.config_swd(200, Tp, Tp)
    ctx3.config_obs(np.full(256, 3.3))
    for B in ([128] if a.quick else [64, 1024]):
        rows = run(ctx3, layered(B, 200, 8), 2, 256, 1)
        res.append({"workload": "C5-like SWD n=200, 128 Rc + 128 Rg (3 sequences/model)", "B": B, "rows": rows})
        print(json.dumps(res[-1]), flush=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
