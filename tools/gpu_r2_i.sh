#!/bin/bash
# round-2 GPU session I: paired-layer GL=1 team kernels + new shapes: identity tests and the mapping sweep
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/i_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/i_build.log; exit 1; }
timeout 900 python -m pytest tests/test_gpu_roots_team.py -q -m gpu > gpurun_out/i_team_tests.log 2>&1
echo "team tests rc=$?" >> gpurun_out/i_team_tests.log
timeout 900 python tools/roots_sweep.py --out gpurun_out/roots_sweep_i.json > gpurun_out/roots_sweep_i.log 2>&1
tail -n 3 gpurun_out/i_team_tests.log
