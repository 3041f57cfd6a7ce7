#!/bin/bash
# round 2: strong-scaling bench on 8 GPUs (one process per GPU, IPC communicator) + config 5 over 8 GPUs
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_scale_8.json 2> gpurun_out/r2_scale_8.err
tail -3 gpurun_out/r2_scale_8.err; cat gpurun_out/r2_scale_8.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_cfg5_8.json 2> gpurun_out/r2_cfg5_8.err
tail -3 gpurun_out/r2_cfg5_8.err; cat gpurun_out/r2_cfg5_8.json
