#!/bin/bash
# round-2 GPU session: root-search mapping sweep with the final kernels (thread unsorted / sorted / every team shape)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/sw_build.log 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 170 python tools/roots_sweep.py --out gpurun_out/roots_sweep_final.json > gpurun_out/sw.log 2>&1
echo "sweep rc=$?"; tail -c 300 gpurun_out/sw.log
