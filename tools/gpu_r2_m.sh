#!/bin/bash
# round-2 GPU session M: length-sorted scheduling of the thread-mapped root search — identity test, on/off sweep, bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/m_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/m_build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_roots_team.py -q -m gpu -x --durations=5 > gpurun_out/m_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/m_tests.log
timeout 900 python tools/sched_sweep.py --out gpurun_out/m_sched_sweep.json > gpurun_out/m_sweep.log 2>&1
echo "sweep rc=$?" >> gpurun_out/m_sweep.log
timeout 600 python bench.py --no-hmc --no-configs --no-cpu-baseline > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err
tail -n 3 gpurun_out/m_tests.log; cat gpurun_out/m_sweep.log | cut -c1-400; tail -c 1500 gpurun_out/m_bench.json
