"""Turn the ncu outputs of a gpurun call into the tracked summaries under profiles/.

    python tools/summarize_ncu.py <round-tag> gpurun_out/launches.csv gpurun_out/prof.ncu-rep

writes profiles/<tag>_launches.md (per-kernel launch list + shares), profiles/<tag>_launch_shares.json
(read by bench.py for the roofline), profiles/<tag>_kernels.md (key `ncu --set full` metrics per
kernel) and profiles/<tag>_traffic.json (DRAM bytes per launch of the dominant kernel)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0].replace('void ', '').replace('rfs::', '')
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        v *= {'us': 1e-3, 'usecond': 1e-3, 'ns': 1e-6, 'nsecond': 1e-6, 's': 1e3, 'second': 1e3}.get(u, 1.0)
        agg.setdefault(name, []).append(v)
    return agg


def main():
    tag, lcsv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
    pd = os.path.join(ROOT, "profiles")
    os.makedirs(pd, exist_ok=True)
    agg = launches(lcsv)
    ours = {k: v for k, v in agg.items() if not k.startswith('at::') and 'dfma_peak' not in k}
    tot = sum(sum(v) for v in ours.values())
    shares = {k.split('<')[0]: sum(v) / tot for k, v in ours.items()}
    with open(os.path.join(pd, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: launch list of `bench.py --steps 2 --warmup 3` under "
                "`ncu --metrics gpu__time_duration.sum --clock-control none`\n\n"
                "Per-launch times are cold-cache and serialised under the profiler: compare SHARES.\n"
                "Shares are over this repo's kernels of one step (torch fill/copy and the DFMA peak probe excluded).\n\n"
                "| kernel | launches | mean ms | share of step |\n|---|---|---|---|\n")
        for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.3f} | {sum(v)/tot:.3f} |\n")
        f.write("\nOther launches seen (not ours): " +
                ", ".join(f"`{k[:60]}` x{len(v)}" for k, v in agg.items() if k not in ours) + "\n")
    json.dump(shares, open(os.path.join(pd, f"{tag}_launch_shares.json"), "w"), indent=1)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
            'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
            'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
            'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
            'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
    seen = set()
    traffic = {}
    with open(os.path.join(pd, f"{tag}_kernels.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` of one launch per kernel (bench.py workload)\n\n")
        for r in rows[2:]:
            name = r[idx['Kernel Name']].split('(')[0].replace('void ', '')
            if name in seen:
                continue
            seen.add(name)
            f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for w in want:
                if w in idx:
                    f.write(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |\n")
            f.write("\n")
            try:
                def tobytes(key):
                    v = float(r[idx[key]].replace(',', ''))
                    return v * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}.get(units[idx[key]], 1.0)
                traffic[name.split('<')[0] + "_dram_bytes_per_launch"] = tobytes('dram__bytes_read.sum') + tobytes('dram__bytes_write.sum')
            except Exception:
                pass
    json.dump(traffic, open(os.path.join(pd, f"{tag}_traffic.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(pd)))


if __name__ == "__main__":
    main()
