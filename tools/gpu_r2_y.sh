#!/bin/bash
# round-2 GPU session Y: final state (front stream + hold kernel default) — suite, smoke, full bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/y_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/y_build.log; exit 1; }
timeout 1800 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/y_all_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/y_all_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/y_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/y_smoke.log
timeout 1200 python bench.py > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err
tail -n 3 gpurun_out/y_all_tests.log gpurun_out/y_smoke.log
