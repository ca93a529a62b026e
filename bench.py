#!/usr/bin/env python
"""bench.py — joint RF+SWD forward+gradient evaluations/s on N B200s (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            (our arm: sm_100a CUDA through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  (CPU arm: oracle restatement of the
                                                              reference, all host threads)
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, chains sharded)

One "step" = one batched Joint_RF_SWD.misfit_and_grad over `--chains` chain states per GPU at the
C1/C4 sizes of SURVEY.md §8 (n=7 layers, 36 Rc + 36 Rg periods, RF nt=125 -> nft=128; reference
param.yaml), i.e. the evaluation every leapfrog step of every chain performs.  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# frozen algorithmic-work constants (hand counts of SURVEY.md §8d)
F_R = 375.0      # flop per Rayleigh secular-function layer step (dltar4 body)
N_LAYERS = 7


def workload(nchains, seed):
    from rfsurfhmc_b200.fixtures import f1_config, f1_true_model, driver_bounds, sorted_uniform_models
    cfg = f1_config()
    x0 = f1_true_model()
    X = sorted_uniform_models(driver_bounds(x0), nchains, seed)
    return cfg, x0, X


def make_dobs(cfg, x0):
    """CPU arm only: observations = synthetics of the true model (main_base.py:49-56) from the
    oracle, which is the implementation that arm measures."""
    from oracle.oracle import Oracle
    O = Oracle()
    nd = cfg["nt"] + len(cfg["tRc"]) + len(cfg["tRg"])
    _, _, d, _ = O.joint_batch(x0[None, :], np.zeros(nd), cfg)
    return d[0]


def make_dobs_gpu(ctx, cfg, x0):
    """Our arm: observations = synthetics of the true model computed by the CUDA path itself, as
    the reference driver does with its own forward model (main_base.py:49-56)."""
    nd = cfg["nt"] + len(cfg["tRc"]) + len(cfg["tRg"])
    ctx.config_obs(np.zeros(nd))
    _, _, d, f = ctx.misfit_grad_host(x0[None, :])
    assert f[0]
    return d[0].copy()


class ClockSampler(threading.Thread):
    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev = dev
        self.rows = []
        self.stop = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.rows.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons}


def cpu_arm(cfg, dobs, X, seconds, nthreads):
    """Oracle restatement (-O3 build) timed on the host cores: evaluations/s."""
    from oracle.oracle import Oracle
    O = Oracle(fast=True)
    chunk = max(64, 16 * nthreads)
    done = 0
    t0 = time.perf_counter()
    i = 0
    while True:
        xb = X[(i * chunk) % len(X):][:chunk]
        if len(xb) < chunk:
            xb = X[:chunk]
        O.joint_batch(xb, dobs, cfg, nthreads=nthreads)
        done += len(xb)
        i += 1
        if time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=16384, help="chain states per GPU (C4: 16384)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-hmc", action="store_true", help="skip the short device-resident HMC leg")
    ap.add_argument("--hmc-traj", type=int, default=10, help="trajectories per chain in the HMC leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1
    config = {"workload": "C4/C1 joint RF+SWD misfit_and_grad: n=7 layers, 36 Rc + 36 Rg periods (5-40 s), "
                          "RF nt=125 (nft=128) freq-domain P, %d chain states per GPU" % args.chains,
              "chains_per_gpu": args.chains, "parallelism": "chains sharded, no data-path collective",
              "l2": "inputs rotated over 4 batches + 256 MiB L2 flush between steps (inside the timed region)"}

    # ------------------------------------------------------------------ reference (CPU) arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        cfg, x0, X = workload(4096, 1234)
        dobs = make_dobs(cfg, x0)
        per_step = max(1.0, min(10.0, 120.0 / max(1, args.steps + args.warmup)))
        for _ in range(min(args.warmup, 1)):
            cpu_arm(cfg, dobs, X, 0.5, ncores)
        tot, tt = 0, 0.0
        for _ in range(args.steps):
            v, done, dt = cpu_arm(cfg, dobs, X, per_step, ncores)
            tot += done
            tt += dt
        val = tot / tt
        print(json.dumps({
            "impl": "reference", "metric": "joint RF+SWD forward+gradient evaluations/s",
            "value": val, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": val, "unit": "evals/s", "cores": ncores, "kind": "port",
                             "sample": "%d evaluations in %.1f s (oracle restatement of the reference CPU "
                                       "path, g++ -O3, %d threads; reference itself not buildable: no "
                                       "gfortran/FFTW3)" % (tot, tt, ncores)},
            "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    from rfsurfhmc_b200._lib import Context
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (rfsurfhmc_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.chains
    cfg, x0, _ = workload(1, 0)
    ctx = Context(local_rank)
    ctx.config_swd(N_LAYERS, tRc=cfg["tRc"], tRg=cfg["tRg"])
    ctx.config_rf(N_LAYERS, cfg["ray_p"], cfg["nt"], cfg["dt"], cfg["gauss"], cfg["time_shift"],
                  cfg["water"], cfg["rf_type"], cfg["method"])
    dobs = make_dobs_gpu(ctx, cfg, x0)
    nd = dobs.size
    ctx.config_obs(dobs)
    nrot = 4
    Xs = [workload(B, 1000 + 97 * rank + i)[2] for i in range(nrot)]
    xd = [torch.from_numpy(x).to(dev) for x in Xs]
    U = torch.empty(B, dtype=torch.float64, device=dev)
    G = torch.empty(B, 2 * N_LAYERS, dtype=torch.float64, device=dev)
    D = torch.empty(B, nd, dtype=torch.float64, device=dev)
    Fl = torch.empty(B, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step(i):
        flush.zero_()
        ctx.misfit_grad_dev(B, xd[i % nrot].data_ptr(), 0, U.data_ptr(), G.data_ptr(), D.data_ptr(),
                            Fl.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    fp64_peak = ctx.measure_fp64_peak()
    ctx.count_evals(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - l0
    nev = ctx.read_evals()
    ctx.count_evals(False)
    sampler.stop = True
    n_fail = int((Fl == 0).sum().item())

    # kernel-level timing of the dominant kernel (swd_roots) with CUDA events on the same stream:
    # re-run the SWD-only path where the roots kernel is >70 % of the time is not exact, so time the
    # full step with the flush excluded instead and attribute by the ncu share (profiles/).
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record(stream)
    for i in range(5):
        ctx.misfit_grad_dev(B, xd[i % nrot].data_ptr(), 0, U.data_ptr(), G.data_ptr(), D.data_ptr(),
                            Fl.data_ptr(), stream.cuda_stream)
    ev[1].record(stream)
    torch.cuda.synchronize()
    ms_noflush = ev[0].elapsed_time(ev[1]) / 5

    # ---- e2e: the public host-buffer API (rfsurfhmc_b200.batched.HostPipeline): every step copies
    # its inputs from pinned host memory and reads U, grad, dsyn, flag back to the host; the two
    # slots overlap the copies of one batch with the kernels of the next
    from rfsurfhmc_b200.batched import HostPipeline
    xh = [torch.from_numpy(x).pin_memory() for x in Xs]
    pipe = HostPipeline(cfg, dobs, N_LAYERS, B, device=local_rank)
    for i in range(3):
        pipe.submit(xh[i % nrot])
    pipe.drain()
    barrier()
    l_e2e0 = pipe.launches
    t0 = time.perf_counter()
    chk = 0.0
    for i in range(args.steps):
        done = pipe.submit(xh[i % nrot])
        if done is not None:
            chk += float(done[0][0])        # the step's result is read on the host
    last = pipe.drain()
    chk += float(last[0][0])
    barrier()
    t_e2e = time.perf_counter() - t0
    launches_e2e = pipe.launches - l_e2e0

    # ---- secondary metric: device-resident HMC (C4: L=20 leapfrog steps per trajectory)
    hmc = None
    if not args.no_hmc:
        from rfsurfhmc_b200.fixtures import driver_bounds
        from rfsurfhmc_b200.distributed import shard_chains
        ids = shard_chains(B * world, rank, world)
        ntraj = args.hmc_traj
        barrier()
        t0 = time.perf_counter()
        ho = ctx.hmc_run(0, ids, driver_bounds(x0), 0.02, Lrange=(20, 20), seed=991206, nsamples=ntraj,
                         ndraws=0, max_iters=ntraj, want_samples=False, want_syn=False)
        barrier()
        th = time.perf_counter() - t0
        hmc = [th, float(ho["n_iter"].sum()), float(ho["n_acc"].sum()), float(ho["evals"])]
        # C4 names the dual-averaging sampler (main_DA.py): lambda = L0*dt = 20*0.02, step size adapted
        # during the first half of the trajectories; L = int(lambda/dt) capped at 40 (extension, see
        # the reference's gamma = 0.05 lets L explode after one rejected warm-up trajectory)
        barrier()
        t0 = time.perf_counter()
        hd = ctx.hmc_run(1, ids, driver_bounds(x0), 0.02, Lrange=(1, 40), L0=20, target_ratio=0.65,
                         seed=991206, nsamples=ntraj - ntraj // 2, ndraws=ntraj // 2, max_iters=ntraj,
                         want_samples=False, want_syn=False)
        barrier()
        td = time.perf_counter() - t0
        hmc += [td, float(hd["n_iter"].sum()), float(hd["n_acc"].sum()), float(hd["evals"])]

    # max over ranks
    tt = torch.tensor([ms, t_e2e * 1e3, ms_noflush], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_noflush = [float(v) for v in tt.tolist()]
    if hmc is not None:
        hv = torch.tensor(hmc, dtype=torch.float64, device=dev)
        hmax = hv.clone()
        if world > 1:
            dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(hv, op=dist.ReduceOp.SUM)
        th = float(hmax[0])
        hmc = {"sampler": "HamitonianMC, L=20, dt=0.02, %d chains/GPU, %d trajectories each (initial models, "
                          "includes chain initialisation and the first evaluation)" % (B, args.hmc_traj),
               "trajectories_per_s": float(hv[1]) / th, "accepted_samples_per_s": float(hv[2]) / th,
               "evals_per_s": float(hv[3]) / th, "seconds": th}
        td = float(hmax[4])
        hmc["dual_averaging"] = {
            "sampler": "HMCDualAveraging, L0=20, dt0=0.02, target 0.65, L capped at 40, %d chains/GPU, %d "
                       "trajectories each (first half adapts the step size; includes _find_initial_dt)"
                       % (B, args.hmc_traj),
            "trajectories_per_s": float(hv[5]) / td, "accepted_samples_per_s": float(hv[6]) / td,
            "evals_per_s": float(hv[7]) / td, "seconds": td}
    value = world * B * args.steps / (ms * 1e-3)
    e2e_val = world * B * args.steps / (ms_e2e * 1e-3)

    out = None
    if rank == 0:
        # ---- roofline of the dominant kernel (swd_roots_kernel, FP64-pipe bound)
        share = None
        prof = os.path.join(ROOT, "profiles", "r01_launch_shares.json")
        if os.path.exists(prof):
            try:
                share = json.load(open(prof)).get("swd_roots_kernel")
            except Exception:
                share = None
        evals_per_launch = nev / max(1, args.steps)
        flop_per_launch = evals_per_launch * (N_LAYERS - 1) * F_R
        t_kernel = (ms_noflush * 1e-3) * (share if share else 1.0)
        achieved = flop_per_launch / t_kernel / 1e12
        roofline = {"bound": "fp64", "kernel": "swd_roots_kernel", "achieved": achieved, "peak": fp64_peak,
                    "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None, "traffic": None,
                    "note": "peak = DFMA micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 "
                            "figure); achieved = secular evaluations counted on device x (n-1) layer steps x "
                            "F_R=375 flop / (step time without flush x ncu share %s)" %
                            ("%.3f" % share if share else "unknown -> 1.0, lower bound"),
                    "secular_evals_per_launch": evals_per_launch, "step_ms_no_flush": ms_noflush}
        tr = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tr):
            try:
                roofline["traffic"] = json.load(open(tr)).get("swd_roots_kernel_dram_bytes_per_launch")
            except Exception:
                pass
        cpu = None
        if not args.no_cpu_baseline:
            cfg2, _, Xc = workload(4096, 4321)
            v, done, dtc = cpu_arm(cfg2, dobs, Xc, args.cpu_seconds, ncores)
            cpu = {"value": v, "unit": "evals/s", "cores": ncores, "kind": "port",
                   "sample": "%d evaluations of the same workload in %.1f s (oracle restatement, g++ -O3, "
                             "%d threads)" % (done, dtc, ncores)}
        out = {"metric": "joint RF+SWD forward+gradient evaluations/s", "value": value, "unit": "evals/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": config, "clocks": sampler.summary(),
               "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": pipe.h2d_bytes,
                       "d2h_bytes_per_step": pipe.d2h_bytes, "gpu_launches": launches_e2e,
                       "api": "rfsurfhmc_b200.batched.HostPipeline (2 slots: copies overlap the next batch)"},
               "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "hmc": hmc,
               "failed_models_last_step": n_fail}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
