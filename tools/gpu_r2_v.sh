#!/bin/bash
# round-2 GPU session V: power-of-two rescaling every 2nd (v8) / 4th (v8b) layer, 9 bisection steps in the key (v8c)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/v_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/v_build.log; exit 1; }
for v in v8 v8b; do
  timeout 300 python tools/compare_libs.py rfsurfhmc_b200/lib/librfsurf_b200.so build/$v.so --batch 32768 > gpurun_out/v_cmp_$v.log 2>&1
  echo "compare $v rc=$?"; tail -n 2 gpurun_out/v_cmp_$v.log
done
export CHAINS="16384 65536 2048"
timeout 900 bash tools/quick_bench.sh default build/v8.so build/v8b.so build/v8c.so default 2>&1 | tee gpurun_out/v_quick.log
