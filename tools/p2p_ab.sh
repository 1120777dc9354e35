#!/bin/bash
# A/B of the peer-memory fast path against the NCCL path on N GPUs of one box: tools/p2p_ab.sh N [edge]
N=${1:-2}; EDGE=${2:-465}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multirank_check p2p"; timeout 600 $TR --master-port 29541 tests/multirank_check.py > $OUT/p2p_check_$N.log 2>&1; echo rc=$?; grep -E "peer-memory|parity|poisson|toggle|MULTIRANK" $OUT/p2p_check_$N.log
echo "== multirank_check nccl"; FSP_P2P=0 timeout 600 $TR --master-port 29542 tests/multirank_check.py > $OUT/nccl_check_$N.log 2>&1; echo rc=$?; grep -E "peer-memory|MULTIRANK" $OUT/nccl_check_$N.log
for t in test_mat test_fss test_ode test_fsp_solver; do
  echo "== $t x$N"; MASTER_PORT=29551 timeout 600 tools/launch_ranks.sh $N build/tests/$t 2>&1 | tail -2
done
echo "== bench p2p"; timeout 600 $TR --master-port 29543 bench.py --gpus $N --lattice $EDGE --steps 200 --warmup 20 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_p2p_$N.json | cut -c1-400
echo "== bench nccl"; FSP_P2P=0 timeout 600 $TR --master-port 29544 bench.py --gpus $N --lattice $EDGE --steps 200 --warmup 20 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | tee $OUT/bench_nccl_$N.json | cut -c1-400
for s in krylov cvode; do
  echo "== lattice_solve $s p2p"; MASTER_PORT=29552 timeout 600 tools/launch_ranks.sh $N build/examples/lattice_solve --edge 215 --solver $s 2>&1 | tail -1 | tee -a $OUT/lattice_solve_$N.log
  echo "== lattice_solve $s nccl"; FSP_P2P=0 MASTER_PORT=29553 timeout 600 tools/launch_ranks.sh $N build/examples/lattice_solve --edge 215 --solver $s 2>&1 | tail -1 | tee -a $OUT/lattice_solve_$N.log
done
