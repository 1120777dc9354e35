#!/bin/bash
# Round-2 multi-GPU check on one box: tools/r2_check.sh N [quick]
#   parity (peer-memory fused kernel, round-1 split path, NCCL path), then the driver's exact bench command at 1 and N GPUs
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -$N
echo "== multirank_check fused x$N"; timeout 600 $TR --master-port 29541 tests/multirank_check.py > $OUT/r2_check_fused_$N.log 2>&1; echo rc=$?; grep -E "peer-memory|parity|poisson|toggle|MULTIRANK|Error|error" $OUT/r2_check_fused_$N.log | head -20
if [ "$2" != "quick" ]; then
echo "== multirank_check split x$N"; FSP_P2P_MODE=split timeout 600 $TR --master-port 29542 tests/multirank_check.py > $OUT/r2_check_split_$N.log 2>&1; echo rc=$?; grep -E "MULTIRANK|rror" $OUT/r2_check_split_$N.log | head
echo "== multirank_check nccl x$N"; FSP_P2P=0 timeout 600 $TR --master-port 29543 tests/multirank_check.py > $OUT/r2_check_nccl_$N.log 2>&1; echo rc=$?; grep -E "MULTIRANK|rror" $OUT/r2_check_nccl_$N.log | head
fi
echo "== bench 1 GPU (driver command)"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-solve --no-cpu-baseline > $OUT/r2_bench_1_of_$N.json 2> $OUT/r2_bench_1_of_$N.err; echo rc=$?; cut -c1-260 $OUT/r2_bench_1_of_$N.json
for mode in fused split; do
  echo "== bench x$N $mode (driver command) + step trace"
  FSP_P2P_MODE=$mode timeout 600 $TR --master-port 29545 bench.py --gpus $N --steps 20 --warmup 5 --no-solve --no-cpu-baseline --trace-steps > $OUT/r2_bench_${mode}_$N.json 2> $OUT/r2_bench_${mode}_$N.err; echo rc=$?
  cut -c1-260 $OUT/r2_bench_${mode}_$N.json; grep "trace rank" $OUT/r2_bench_${mode}_$N.err | cut -c1-200
  FSP_P2P_MODE=$mode timeout 600 $TR --master-port 29546 bench.py --gpus $N --steps 200 --warmup 20 --no-solve --no-cpu-baseline --no-e2e --no-parity > $OUT/r2_bench_${mode}_${N}_long.json 2>/dev/null; cut -c1-200 $OUT/r2_bench_${mode}_${N}_long.json
done
echo "== bench x$N nccl"; FSP_P2P=0 timeout 600 $TR --master-port 29547 bench.py --gpus $N --steps 200 --warmup 20 --no-solve --no-cpu-baseline --no-e2e --no-parity > $OUT/r2_bench_nccl_$N.json 2>/dev/null; cut -c1-200 $OUT/r2_bench_nccl_$N.json
