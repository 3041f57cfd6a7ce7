#!/bin/bash
# round 2: strong-scaling bench on N GPUs (one process per GPU, IPC communicator) + config 5 over the same GPUs
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_scale_$N.json 2> gpurun_out/r2f_scale_$N.err
tail -3 gpurun_out/r2f_scale_$N.err; cat gpurun_out/r2f_scale_$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_cfg5_$N.json 2> gpurun_out/r2f_cfg5_$N.err
tail -3 gpurun_out/r2f_cfg5_$N.err; cat gpurun_out/r2f_cfg5_$N.json
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multidev.py tests/test_gpu_peers.py -m gpu -x -q 2>&1 | tail -4; fi
