#!/bin/bash
# round-2 GPU session Q: sorted order + RF branch released at the root-search launch (v6b: 6 bisections, v6c: 3)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/q_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/q_build.log; exit 1; }
RFS_ROOTS_SCHED=1 timeout 300 python tools/compare_libs.py rfsurfhmc_b200/lib/librfsurf_b200.so build/v6b.so --batch 32768 > gpurun_out/q_cmp.log 2>&1
echo "compare v6b rc=$?"; tail -n 2 gpurun_out/q_cmp.log
export CHAINS="16384 65536 8192 32768"
timeout 600 bash tools/quick_bench.sh default 2>&1 | tee gpurun_out/q_quick.log
echo "--- v6b sched off"; RFS_ROOTS_SCHED=0 CHAINS=16384 timeout 600 bash tools/quick_bench.sh build/v6b.so 2>&1 | tee -a gpurun_out/q_quick.log
echo "--- v6b sched on"; RFS_ROOTS_SCHED=1 timeout 600 bash tools/quick_bench.sh build/v6b.so 2>&1 | tee -a gpurun_out/q_quick.log
echo "--- v6c sched on"; RFS_ROOTS_SCHED=1 timeout 600 bash tools/quick_bench.sh build/v6c.so 2>&1 | tee -a gpurun_out/q_quick.log
