#!/bin/bash
# round-2 GPU session P: sorted job order with balanced 128-thread blocks (build/v6.so, RFS_ROOTS_SCHED=1) vs in-tree
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/p_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/p_build.log; exit 1; }
RFS_ROOTS_SCHED=1 timeout 300 python tools/compare_libs.py rfsurfhmc_b200/lib/librfsurf_b200.so build/v6.so --batch 32768 > gpurun_out/p_cmp_v6.log 2>&1
echo "compare v6 rc=$?"; tail -n 3 gpurun_out/p_cmp_v6.log
export CHAINS="16384 65536 8192"
timeout 600 bash tools/quick_bench.sh default 2>&1 | tee gpurun_out/p_quick.log
echo "--- v6 sched off"; RFS_ROOTS_SCHED=0 timeout 600 bash tools/quick_bench.sh build/v6.so 2>&1 | tee -a gpurun_out/p_quick.log
echo "--- v6 sched on"; RFS_ROOTS_SCHED=1 timeout 600 bash tools/quick_bench.sh build/v6.so 2>&1 | tee -a gpurun_out/p_quick.log
