#!/bin/bash
# round 2, visit E (2 GPUs): full GPU suite incl. multi-device + multi-process tests; bench at 1 and 2 GPUs
mkdir -p gpurun_out; L=gpurun_out/r2e.log; rm -f $L
nvidia-smi -L >> $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 >> $L
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2e_bench1.json 2> gpurun_out/r2e_bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench2.json 2> gpurun_out/r2e_bench2.err
tail -5 gpurun_out/r2e_bench1.err gpurun_out/r2e_bench2.err >> $L
cat $L; cat gpurun_out/r2e_bench1.json gpurun_out/r2e_bench2.json
