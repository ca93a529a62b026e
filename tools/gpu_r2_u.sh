#!/bin/bash
# round-2 GPU session U: sorted order restricted to staged layer counts — sorted-order test, full bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/u_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/u_build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_roots_team.py -q -m gpu > gpurun_out/u_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/u_tests.log
timeout 1200 python bench.py > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err
tail -n 3 gpurun_out/u_tests.log; tail -c 600 gpurun_out/u_bench.json
