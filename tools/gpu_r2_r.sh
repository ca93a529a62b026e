#!/bin/bash
# round-2 GPU session R: block composition of the sorted order (pairs vs rounds) and RF block size beside it
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/r_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/r_build.log; exit 1; }
RFS_ROOTS_SCHED=2 timeout 300 python tools/compare_libs.py rfsurfhmc_b200/lib/librfsurf_b200.so build/v6e.so --batch 32768 > gpurun_out/r_cmp.log 2>&1
echo "compare v6e mode 2 rc=$?"; tail -n 2 gpurun_out/r_cmp.log
export CHAINS="16384"
timeout 600 bash tools/quick_bench.sh default 2>&1 | tee gpurun_out/r_quick.log
for m in 1 2; do for rb in 128 64; do
  echo "--- v6e sched $m rf_block $rb"; RFS_ROOTS_SCHED=$m RFS_RF_BLOCK=$rb timeout 600 bash tools/quick_bench.sh build/v6e.so 2>&1 | tee -a gpurun_out/r_quick.log
done; done
echo "--- v6e sched 2 rf 128 at 65536/32768"; CHAINS="65536 32768" RFS_ROOTS_SCHED=2 timeout 600 bash tools/quick_bench.sh build/v6e.so 2>&1 | tee -a gpurun_out/r_quick.log
timeout 600 bash tools/quick_bench.sh default 2>&1 | tee -a gpurun_out/r_quick.log
