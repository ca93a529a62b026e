#!/bin/bash
# round-2 GPU session X: where the root search of the first chunk runs (caller's stream / front stream, priority, RF delay)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/x_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/x_build.log; exit 1; }
for cfg in "0 1" "1 1" "2 1" "1 0" "3 1" "0 1"; do
  set -- $cfg
  RFS_FRONT_MODE=$1 RFS_FRONT_PRIO=$2 timeout 300 python bench.py --no-configs --no-cpu-baseline --hmc-traj 20 --da-budget 3 > gpurun_out/x_bench_$1_$2.json 2> gpurun_out/x_bench_$1_$2.err
  python - <<PY
import json
L=[l for l in open("gpurun_out/x_bench_$1_$2.json") if l.strip().startswith("{")]
d=json.loads(L[-1])
print("front_mode $1 prio $2: value %.3f M (%.2f ms) e2e %.3f M hmc c4 %.3f M da %.3f M" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["hmc"]["c4_strong"]["evals_per_s"]/1e6, d["hmc"]["da_capped"]["evals_per_s"]/1e6))
PY
done
