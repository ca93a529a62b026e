#!/bin/bash
# round-2 GPU session O: A/B of root-search variants built into build/*.so (bit identity + quick bench)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/o_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/o_build.log; exit 1; }
for v in $VARIANTS; do
  timeout 300 python tools/compare_libs.py rfsurfhmc_b200/lib/librfsurf_b200.so build/$v.so > gpurun_out/o_cmp_$v.log 2>&1
  echo "compare $v rc=$?"; tail -n 3 gpurun_out/o_cmp_$v.log
done
LIBS="default"; for v in $VARIANTS; do LIBS="$LIBS build/$v.so"; done
CHAINS="16384 65536 2048" timeout 900 bash tools/quick_bench.sh $LIBS default 2>&1 | tee gpurun_out/o_quick.log
