#!/bin/bash
# 8-GPU validation + strong-scaling measurement of the peer-memory path (one box): tools/scale8.sh [N]
N=${1:-8}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multirank_check p2p x$N"; timeout 400 $TR --master-port 29541 tests/multirank_check.py > $OUT/p2p_check_$N.log 2>&1; echo rc=$?; grep -E "peer-memory|parity|poisson|toggle|MULTIRANK" $OUT/p2p_check_$N.log
echo "== bench 1 GPU"; timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_1_of_$N.json | cut -c1-200
echo "== bench p2p x$N"; timeout 300 $TR --master-port 29543 bench.py --gpus $N --steps 400 --warmup 40 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_p2p_$N.json | cut -c1-200
echo "== bench nccl x$N"; FSP_P2P=0 timeout 300 $TR --master-port 29544 bench.py --gpus $N --steps 400 --warmup 40 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_nccl_$N.json | cut -c1-200
for s in krylov cvode; do
  echo "== lattice_solve 465 $s p2p x$N"; MASTER_PORT=29552 timeout 300 tools/launch_ranks.sh $N build/examples/lattice_solve --edge 465 --solver $s --repeat 2 2>&1 | tail -1 | tee -a $OUT/lattice_solve_$N.log
done
echo "== lattice_solve 465 krylov nccl x$N"; FSP_P2P=0 MASTER_PORT=29553 timeout 300 tools/launch_ranks.sh $N build/examples/lattice_solve --edge 465 --solver krylov --repeat 2 2>&1 | tail -1 | tee -a $OUT/lattice_solve_$N.log
