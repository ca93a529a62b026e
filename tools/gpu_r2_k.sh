#!/bin/bash
# round-2 GPU session K: full suite, smoke, full bench after the eigen (all layer counts) and mapping-threshold changes
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/k_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/k_build.log; exit 1; }
timeout 1800 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/k_all_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/k_all_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/k_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/k_smoke.log
timeout 1200 python bench.py > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
tail -n 3 gpurun_out/k_all_tests.log gpurun_out/k_smoke.log
