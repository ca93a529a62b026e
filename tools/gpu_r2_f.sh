#!/bin/bash
# round-2 GPU session F (N GPUs): torchrun bench at N = $1
N=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/f_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/f_build.log; exit 1; }
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/f_bench$N.json 2> gpurun_out/f_bench$N.err
echo "bench rc=$?" >> gpurun_out/f_bench$N.err
tail -n 3 gpurun_out/f_bench$N.err; wc -c gpurun_out/f_bench$N.json
