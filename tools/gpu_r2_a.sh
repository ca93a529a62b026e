#!/bin/bash
# round-2 GPU session A: team root search (bit-identity tests + mapping sweep), regression of the suite, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/a_build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_roots_team.py -x -q -m gpu > gpurun_out/a_team_tests.log 2>&1
echo "team tests rc=$?" >> gpurun_out/a_team_tests.log
timeout 900 python tools/roots_sweep.py --out gpurun_out/roots_sweep.json > gpurun_out/roots_sweep.log 2>&1
echo "sweep rc=$?" >> gpurun_out/roots_sweep.log
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_roots_team.py --durations=15 > gpurun_out/a_all_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/a_all_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -3 gpurun_out/a_team_tests.log gpurun_out/a_all_tests.log
