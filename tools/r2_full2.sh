#!/bin/bash
# full test + driver-shaped bench at 1 and N GPUs on one box: tools/r2_full2.sh N
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/r2_pytest_gpu_$N.log 2>&1; echo rc=$?; tail -5 $OUT/r2_pytest_gpu_$N.log
echo "== reference arm N=1"; timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/r2_ref_1.json 2>$OUT/r2_ref_1.err; echo rc=$?; cut -c1-300 $OUT/r2_ref_1.json
echo "== bench N=1 (driver command)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2_bench_full_1.json 2> $OUT/r2_bench_full_1.err; echo rc=$?; tail -c 2500 $OUT/r2_bench_full_1.json; tail -3 $OUT/r2_bench_full_1.err
echo "== reference arm N=$N (torchrun)"; timeout 600 $TR --master-port 29548 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $OUT/r2_ref_$N.json 2>$OUT/r2_ref_$N.err; echo rc=$?; grep -o '"cores": [0-9]*' $OUT/r2_ref_$N.json | head -1; cut -c1-120 $OUT/r2_ref_$N.json
echo "== bench N=$N (driver command)"; timeout 900 $TR --master-port 29549 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r2_bench_full_$N.json 2> $OUT/r2_bench_full_$N.err; echo rc=$?; tail -c 2500 $OUT/r2_bench_full_$N.json; grep -v "^$\|OMP_NUM\|\*\*\*\|NCCL version" $OUT/r2_bench_full_$N.err | tail -5
