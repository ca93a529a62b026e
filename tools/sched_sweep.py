"""Length-sorted scheduling of the thread-mapped root search: on vs off, by batch size.

    python tools/sched_sweep.py [--out gpurun_out/sched_sweep.json]

Per workload and batch size: milliseconds of the root-search class (CUDA events around the launches,
rfs_profile_eval: key + sort + search kernels when sorted) and of the whole evaluation as the caller
sees it (RF branch overlapped), with the schedule off and on.  RFS_SCHED_MIN_JOBS in csrc/capi.cu comes
from this table (kept under profiles/)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rfsurfhmc_b200._lib import Context
from rfsurfhmc_b200.fixtures import f1_config, f1_true_model, driver_bounds, sorted_uniform_models
from tools.roots_sweep import layered


def run(ctx, X, which, nd, reps):
    dev = torch.device("cuda", 0)
    B, n2 = X.shape
    xd = torch.from_numpy(X).to(dev)
    U = torch.empty(B, dtype=torch.float64, device=dev)
    G = torch.empty(B, n2, dtype=torch.float64, device=dev)
    D = torch.empty(B, nd, dtype=torch.float64, device=dev)
    F = torch.empty(B, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()
    args = (B, xd.data_ptr(), which, U.data_ptr(), G.data_ptr(), D.data_ptr(), F.data_ptr(), st.cuda_stream)
    row, ref = {}, None
    ctx.set_roots_team(0)
    for mode in (0, 1):
        ctx.set_roots_sched(mode)
        ctx.profile_eval(*args)
        ms_r = [ctx.profile_eval(*args)["swd_roots"][0] for _ in range(reps)]
        ms_e = []
        for _ in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            ctx.misfit_grad_dev(*args)
            e1.record(st)
            torch.cuda.synchronize()
            ms_e.append(e0.elapsed_time(e1))
        out = (U.cpu().numpy().copy(), G.cpu().numpy().copy(), D.cpu().numpy().copy())
        if ref is None:
            ref = out
        row["on" if mode else "off"] = {"roots_ms": float(np.median(ms_r)), "eval_ms": float(np.median(ms_e[1:]))}
        if mode:
            row["bit_identical"] = bool(all(np.array_equal(a, b, equal_nan=True) for a, b in zip(ref, out)))
    ctx.set_roots_sched(-1)
    ctx.set_roots_team(-1)
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sched_sweep.json")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    res = []
    cfg, x0 = f1_config(), f1_true_model()
    ctx = Context(0)
    ctx.config_swd(7, tRc=cfg["tRc"], tRg=cfg["tRg"])
    ctx.config_rf(7, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"], cfg["water"],
                  cfg["rf_type"], cfg["method"])
    ctx.config_obs(np.full(197, 3.0))
    for B in (2048, 4096, 8192, 12288, 16384, 24576, 32768, 65536):
        r = run(ctx, sorted_uniform_models(driver_bounds(x0), B, seed=100 + B), 0, 197, 3)
        res.append(dict(workload="C1 joint n=7 (3 sequences/model)", B=B, **r))
        print(json.dumps(res[-1]), flush=True)
    Tp = np.geomspace(2, 100, 60)
    ctx2 = Context(0)
    ctx2.config_swd(40, Tp, Tp, Tp, Tp, mode=0)
    ctx2.config_obs(np.full(240, 3.0))
    for B in (2048, 8192, 32768):
        r = run(ctx2, layered(B, 40, 7), 2, 240, 2)
        res.append(dict(workload="C2-like SWD n=40, 60 periods x Rc,Rg,Lc,Lg (6 sequences/model)", B=B, **r))
        print(json.dumps(res[-1]), flush=True)
    Tp = np.geomspace(1, 150, 128)
    ctx3 = Context(0)
    ctx3.config_swd(200, Tp, Tp)
    ctx3.config_obs(np.full(256, 3.3))
    for B in (8192, 32768):
        r = run(ctx3, layered(B, 200, 8), 2, 256, 1)
        res.append(dict(workload="C5-like SWD n=200, 128 Rc + 128 Rg (3 sequences/model)", B=B, **r))
        print(json.dumps(res[-1]), flush=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
