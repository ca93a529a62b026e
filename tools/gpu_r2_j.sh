#!/bin/bash
# round-2 GPU session J: thread-mapped kernel with two layer matrices per trip (shape 1x1): identity, sweep, headline
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/j_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/j_build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_roots_team.py -q -m gpu > gpurun_out/j_team_tests.log 2>&1
echo "team tests rc=$?" >> gpurun_out/j_team_tests.log
timeout 900 python tools/roots_sweep.py --out gpurun_out/roots_sweep_j.json > gpurun_out/roots_sweep_j.log 2>&1
RFS_ROOTS_TEAM=1,1 timeout 600 python bench.py --steps 10 --no-hmc --no-configs --no-cpu-baseline > gpurun_out/j_bench_paired.json 2> gpurun_out/j_bench_paired.err
timeout 600 python bench.py --steps 10 --no-hmc --no-configs --no-cpu-baseline > gpurun_out/j_bench_auto.json 2> gpurun_out/j_bench_auto.err
RFS_ROOTS_TEAM=1,1 timeout 600 python bench.py --steps 5 --chains 65536 --no-hmc --no-configs --no-cpu-baseline > gpurun_out/j_bench_paired_64k.json 2> gpurun_out/j_bench_paired_64k.err
RFS_ROOTS_TEAM=0 timeout 600 python bench.py --steps 5 --chains 65536 --no-hmc --no-configs --no-cpu-baseline > gpurun_out/j_bench_thread_64k.json 2> gpurun_out/j_bench_thread_64k.err
tail -n 3 gpurun_out/j_team_tests.log
