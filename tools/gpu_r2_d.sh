#!/bin/bash
# round-2 GPU session D: leaner team chain; tests, sweep, parity report, ncu captures, full bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/d_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/d_build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_roots_team.py -q -m gpu > gpurun_out/d_team_tests.log 2>&1
echo "team tests rc=$?" >> gpurun_out/d_team_tests.log
timeout 900 python tools/roots_sweep.py --out gpurun_out/roots_sweep_d.json > gpurun_out/roots_sweep_d.log 2>&1
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_roots_team.py --durations=10 > gpurun_out/d_all_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/d_all_tests.log
timeout 600 python tests/gpu_parity_report.py --out gpurun_out/parity_report_d.json > gpurun_out/d_parity.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:swd_roots_team -c 1 -o gpurun_out/d_team_small python tools/ncu_target.py team64 > gpurun_out/d_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"swd_roots_kernel|swd_eigen|rf_propagate|rf_decon|joint_assemble|rf_layer|prep_models" -c 14 -o gpurun_out/d_kernels python tools/ncu_target.py thread16k > gpurun_out/d_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/d_launches.csv python bench.py --steps 2 --warmup 3 --no-hmc --no-configs --no-cpu-baseline > gpurun_out/d_bench_under_ncu.json 2> gpurun_out/d_ncu3.err
timeout 1200 python bench.py > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
tail -n 3 gpurun_out/d_team_tests.log gpurun_out/d_all_tests.log
