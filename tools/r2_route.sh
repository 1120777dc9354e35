#!/bin/bash
OUT=gpurun_out
timeout 300 python -m pytest tests/test_gpu_vec.py -q -m gpu -k "route or scatter" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_multirank.py -q -m gpu > $OUT/r02_pytest_multirank_2b.log 2>&1; echo "multirank pytest rc=$?"; tail -4 $OUT/r02_pytest_multirank_2b.log
for mode in 0 1; do
  FSP_SHARDED_SET=$mode MASTER_PORT=2958$mode timeout 400 tools/launch_ranks.sh 2 build/examples/transcr_reg_6d --solver krylov --log > $OUT/r02_route_transcr_2_$mode.log 2>&1
  echo "transcr_reg_6d krylov 2 ranks FSP_SHARDED_SET=$mode rc=$?"; grep -E "wall_s|timing_s" $OUT/r02_route_transcr_2_$mode.log | cut -c1-330
done
