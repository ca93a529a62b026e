#!/bin/bash
# round-2 GPU session G: full suite + smoke + full bench after the eigen-kernel change
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/g_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/g_build.log; exit 1; }
timeout 1800 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/g_all_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/g_all_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/g_smoke.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"swd_eigen" -c 1 -o gpurun_out/g_eigen python tools/ncu_target.py thread16k > gpurun_out/g_ncu.log 2>&1
timeout 1200 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
tail -n 3 gpurun_out/g_all_tests.log gpurun_out/g_smoke.log
