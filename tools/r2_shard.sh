#!/bin/bash
# sharded state set: 2-rank check + the C++ programs with FSP_SHARDED_SET=1
OUT=gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/multirank_sharded_check.py > $OUT/r02_sharded_$N.log 2>&1
echo "sharded check rc=$?"; grep -v "^\[W\|^W1\|^\*\*\*" $OUT/r02_sharded_$N.log | tail -45
for prog in test_fss test_mat test_fsp_solver; do
  FSP_SHARDED_SET=1 timeout 300 tools/launch_ranks.sh 2 build/tests/$prog > $OUT/r02_sharded_cpp_$prog.log 2>&1
  echo "$prog (sharded sets, 2 ranks) rc=$?"; grep -E "FAILED|tests ran|Failure|error" $OUT/r02_sharded_cpp_$prog.log | head -12
done
# BASELINE config 3 (transcr_reg_6d, adaptive solve) and hog1p on N ranks: replicated vs sharded state set
for ex in transcr_reg_6d hog1p; do
  for mode in 0 1; do
    FSP_SHARDED_SET=$mode MASTER_PORT=2958$mode timeout 400 tools/launch_ranks.sh $N build/examples/$ex --solver krylov > $OUT/r02_sharded_${ex}_${N}_$mode.log 2>&1
    echo "$ex krylov on $N ranks, FSP_SHARDED_SET=$mode rc=$?: $(tail -1 $OUT/r02_sharded_${ex}_${N}_$mode.log | cut -c1-420)"
  done
done
