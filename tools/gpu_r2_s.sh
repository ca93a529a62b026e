#!/bin/bash
# round-2 GPU session S: sorted order, block-to-SM compositions 2 (rounds), 3 (uneven SMs), 4 (plain longest-first)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/s_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/s_build.log; exit 1; }
RFS_ROOTS_SCHED=3 timeout 300 python tools/compare_libs.py rfsurfhmc_b200/lib/librfsurf_b200.so build/v6f.so --batch 32768 > gpurun_out/s_cmp.log 2>&1
echo "compare v6f mode 3 rc=$?"; tail -n 2 gpurun_out/s_cmp.log
for m in 2 3 4; do
  echo "--- v6f sched $m"; CHAINS="16384 12288 20480 65536" RFS_ROOTS_SCHED=$m timeout 600 bash tools/quick_bench.sh build/v6f.so 2>&1 | tee -a gpurun_out/s_quick.log
done
echo "--- default"; CHAINS="12288 20480" timeout 600 bash tools/quick_bench.sh default 2>&1 | tee -a gpurun_out/s_quick.log
