#!/bin/bash
# usage: tools/quick_bench.sh  (on GPU box)
for ov in 0 1; do
  for ch in 16384 65536; do
    echo "== RFS_NO_OVERLAP=$ov chains=$ch"
    RFS_NO_OVERLAP=$ov python bench.py --steps 5 --warmup 3 --chains $ch --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
  done
done
