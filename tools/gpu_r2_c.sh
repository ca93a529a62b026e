#!/bin/bash
# round-2 GPU session C: power-of-two renormalisation + parked-layer team kernel: identity, parity, profile, bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/c_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/c_build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_roots_team.py -q -m gpu > gpurun_out/c_team_tests.log 2>&1
echo "team tests rc=$?" >> gpurun_out/c_team_tests.log
timeout 300 python tools/water_diag.py > gpurun_out/c_water_diag.log 2>&1
timeout 900 python tools/roots_sweep.py --out gpurun_out/roots_sweep_c.json > gpurun_out/roots_sweep_c.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_roots_team.py --durations=10 > gpurun_out/c_all_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/c_all_tests.log
timeout 600 python tests/gpu_parity_report.py --out gpurun_out/parity_report.json > gpurun_out/c_parity.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:swd_roots_team -c 1 -o gpurun_out/c_team_small python tools/ncu_target.py team64 > gpurun_out/c_ncu1.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
tail -n 3 gpurun_out/c_team_tests.log gpurun_out/c_all_tests.log
