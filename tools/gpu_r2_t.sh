#!/bin/bash
# round-2 GPU session T: length-sorted job order in the tree — full suite, smoke, full bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/t_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/t_build.log; exit 1; }
timeout 1800 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/t_all_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/t_all_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/t_smoke.log
timeout 1200 python bench.py > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
tail -n 3 gpurun_out/t_all_tests.log gpurun_out/t_smoke.log
