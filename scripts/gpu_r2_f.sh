#!/bin/bash
# round 2, visit F (2 GPUs): the multi-device / multi-process / session tests, then the whole suite, bench at 2
mkdir -p gpurun_out; L=gpurun_out/r2f.log; rm -f $L
timeout 420 python -m pytest tests -m gpu -x -q -k "multidev or peers or sessions" 2>&1 | tail -25 >> $L
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_multidev.py --deselect tests/test_gpu_peers.py 2>&1 | tail -15 >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench2.json 2> gpurun_out/r2f_bench2.err
tail -5 gpurun_out/r2f_bench2.err >> $L
cat $L; cat gpurun_out/r2f_bench2.json
