#!/bin/bash
# round-2 GPU session H: final state — full suite, smoke, ncu launch list + per-kernel captures, full bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/h_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/h_build.log; exit 1; }
timeout 1800 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/h_all_tests.log 2>&1
echo "suite rc=$?" >> gpurun_out/h_all_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/h_smoke.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"swd_roots_kernel|swd_eigen|rf_propagate|rf_decon|joint_assemble|rf_layer|prep_models" -c 7 -o gpurun_out/h_kernels python tools/ncu_target.py thread16k > gpurun_out/h_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 2 --warmup 3 --no-hmc --no-configs --no-cpu-baseline > gpurun_out/h_bench_under_ncu.json 2> gpurun_out/h_ncu3.err
timeout 600 python tests/gpu_parity_report.py --out gpurun_out/parity_report_h.json > gpurun_out/h_parity.log 2>&1
timeout 1200 python bench.py > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
tail -n 3 gpurun_out/h_all_tests.log gpurun_out/h_smoke.log
