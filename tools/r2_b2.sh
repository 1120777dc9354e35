#!/bin/bash
N=${1:-2}; OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29545 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r02_scale_$N.json 2> $OUT/r02_scale_$N.err; echo rc=$?
cut -c1-200 $OUT/r02_scale_$N.json; grep -o '"halo_exchange".*' $OUT/r02_scale_$N.json | cut -c1-600; grep -o '"e2e".*"gpu_launches": [0-9]*' $OUT/r02_scale_$N.json | cut -c1-500; grep -v "^$\|OMP_NUM\|\*\*\*\|NCCL version" $OUT/r02_scale_$N.err | tail -3
