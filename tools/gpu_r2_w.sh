#!/bin/bash
# round-2 GPU session W: rescaling every 4th layer + high-priority search streams — suite, smoke, 3 quick benches (run-to-run spread), full bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/w_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/w_build.log; exit 1; }
timeout 1800 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/w_all_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/w_all_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/w_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/w_smoke.log
CHAINS="16384" timeout 600 bash tools/quick_bench.sh default default default 2>&1 | tee gpurun_out/w_quick.log
timeout 1200 python bench.py > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err
tail -n 3 gpurun_out/w_all_tests.log gpurun_out/w_smoke.log
