#!/bin/bash
# round-2 GPU session: parity-distribution report with the final kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/pr_build.log 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 150 python tests/gpu_parity_report.py --out gpurun_out/parity_report_final.json > gpurun_out/pr.log 2>&1
echo "report rc=$?"; tail -n 3 gpurun_out/pr.log | cut -c1-300
