#!/bin/bash
# round-2 GPU session E (2 GPUs): reworked time-domain kernel (tests + C3 legs), then the 2-GPU bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/e_build.log 2>&1 || { echo BUILD FAILED; tail -5 gpurun_out/e_build.log; exit 1; }
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -q -m gpu -k "time_domain or c3_rf or finite_q or dropins or golden or test_script" > gpurun_out/e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/e_tests.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 5 --no-hmc --no-cpu-baseline > gpurun_out/e_bench1.json 2> gpurun_out/e_bench1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/e_bench2.json 2> gpurun_out/e_bench2.err
echo "bench2 rc=$?" >> gpurun_out/e_bench2.err
tail -n 3 gpurun_out/e_tests.log; tail -n 2 gpurun_out/e_bench2.err
