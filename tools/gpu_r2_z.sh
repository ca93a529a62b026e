#!/bin/bash
# round-2 GPU session Z: ncu evidence of the final state — per-kernel --set full captures and the launch list of bench.py
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/z_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/z_build.log; exit 1; }
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"swd_roots_kernel|swd_sched|swd_eigen|rf_propagate|rf_decon|joint_assemble|rf_layer|prep_models" -c 9 -o gpurun_out/z_kernels python tools/ncu_target.py thread16k > gpurun_out/z_ncu2.log 2>&1
echo "ncu full rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 3 --no-hmc --no-configs --no-cpu-baseline > gpurun_out/z_bench_under_ncu.json 2> gpurun_out/z_ncu3.err
echo "ncu launches rc=$?"; tail -n 2 gpurun_out/z_ncu2.log; wc -l gpurun_out/z_launches.csv
