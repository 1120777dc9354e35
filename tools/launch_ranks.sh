#!/bin/bash
# Launch a stand-alone C++ program as N ranks on one node, one rank per GPU (the counterpart of `mpirun -np N`):
#   tools/launch_ranks.sh N program [args...]
# Each rank gets RANK, LOCAL_RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT; PACMENSLInit() uses them to pick its GPU
# and to join the NCCL world communicator.
N=$1; shift
export WORLD_SIZE=$N MASTER_ADDR=127.0.0.1 MASTER_PORT=${MASTER_PORT:-29533}
pids=()
for ((r=0; r<N; r++)); do
  if [ $r -eq 0 ]; then RANK=$r LOCAL_RANK=$r "$@" & else RANK=$r LOCAL_RANK=$r "$@" > /dev/null 2>&1 & fi
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=$?; done
exit $rc
