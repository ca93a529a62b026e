#!/bin/bash
# usage: [CHAINS="16384 65536"] tools/quick_bench.sh lib1.so lib2.so ...   ("default" = in-tree library)
# prints: evals/s (device-resident), ms/step, roofline fraction, evals/s through host buffers
for lib in "$@"; do
  for ch in ${CHAINS:-16384}; do
    L=""; [ "$lib" != "default" ] && L=$PWD/$lib
    echo "== lib=$lib chains=$ch"
    RFS_LIB=$L python bench.py --steps 10 --warmup 3 --chains $ch --no-cpu-baseline --no-hmc 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
  done
done
