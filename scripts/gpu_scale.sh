#!/bin/bash
# strong-scaling bench on N GPUs of one node (argument: N)
N=$1
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_$N.json 2> gpurun_out/scale_$N.err
fi
tail -3 gpurun_out/scale_$N.err; cat gpurun_out/scale_$N.json
