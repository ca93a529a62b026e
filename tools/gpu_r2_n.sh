#!/bin/bash
# round-2 GPU session N: rf_propagate block size 128 vs 64 (co-residency with the root search) on C1 / C3 / C5
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/n_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/n_build.log; exit 1; }
for blk in 128 64 32; do
  RFS_RF_BLOCK=$blk timeout 900 python bench.py --no-hmc --no-cpu-baseline > gpurun_out/n_bench_$blk.json 2> gpurun_out/n_bench_$blk.err
  python - <<PY
import json
L=[l for l in open("gpurun_out/n_bench_$blk.json") if l.strip().startswith("{")]
d=json.loads(L[-1])
print("block $blk: C1 %.3f M/s %.2f ms/step e2e %.3f M/s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6), {k: round(v["value"]) for k, v in d["configs"].items()}, "rf_propagate ms", d["kernels"]["rf_propagate"]["ms"])
PY
done
