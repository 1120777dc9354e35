#!/bin/bash
# where the sharded construction spends its time: set trace + component timing of config 3 on N ranks
OUT=gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/multirank_sharded_check.py > $OUT/r02_sharded_$N.log 2>&1
echo "sharded check rc=$?"; grep -E "Expand\(\)|SHARDED|solve " $OUT/r02_sharded_$N.log
for mode in 0 1; do
  FSP_SET_TRACE=$mode FSP_SHARDED_SET=$mode MASTER_PORT=2958$mode timeout 400 tools/launch_ranks.sh $N build/examples/transcr_reg_6d --solver krylov --log > $OUT/r02_sharded_transcr_reg_6d_${N}_$mode.log 2>&1
  echo "transcr_reg_6d krylov on $N ranks, FSP_SHARDED_SET=$mode rc=$?"; grep -E "wall_s|timing_s" $OUT/r02_sharded_transcr_reg_6d_${N}_$mode.log | cut -c1-330
done
grep "\[set\]" $OUT/r02_sharded_transcr_reg_6d_${N}_1.log | awk 'NR<=3 || NR%6==0' | cut -c1-250
