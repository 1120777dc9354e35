#!/bin/bash
# tools/r2_quick.sh N : fused-path parity + the driver's bench command at N GPUs (and 1 GPU on the same box)
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multirank_check fused x$N"; timeout 600 $TR --master-port 29541 tests/multirank_check.py > $OUT/r2_check_fused_$N.log 2>&1; echo rc=$?; grep -E "peer-memory|parity|poisson|toggle|MULTIRANK|Error|error" $OUT/r2_check_fused_$N.log | head -20
echo "== bench 1 GPU"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-solve --no-cpu-baseline --no-e2e --no-parity > $OUT/r2_bench_1_of_$N.json 2> $OUT/r2_bench_1_of_$N.err; cut -c1-200 $OUT/r2_bench_1_of_$N.json
echo "== bench x$N fused (driver command)"
timeout 600 $TR --master-port 29545 bench.py --gpus $N --steps 20 --warmup 5 --no-solve --no-cpu-baseline --trace-steps > $OUT/r2_bench_fused_$N.json 2> $OUT/r2_bench_fused_$N.err; echo rc=$?
cut -c1-200 $OUT/r2_bench_fused_$N.json; grep "trace rank" $OUT/r2_bench_fused_$N.err | cut -c1-120; grep -o '"e2e".*"gpu_launches": [0-9]*' $OUT/r2_bench_fused_$N.json | cut -c1-400
timeout 600 $TR --master-port 29546 bench.py --gpus $N --steps 200 --warmup 20 --no-solve --no-cpu-baseline --no-e2e --no-parity > $OUT/r2_bench_fused_${N}_long.json 2>/dev/null; cut -c1-200 $OUT/r2_bench_fused_${N}_long.json
