#!/bin/bash
OUT=gpurun_out
N=${1:-2}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multirank_check.py > $OUT/r02_check_a2a_$N.log 2>&1; echo "multirank_check rc=$?"; grep -E "alltoallv|MULTIRANK|poisson|toggle_custom|action parity" $OUT/r02_check_a2a_$N.log | cut -c1-160
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/multirank_sharded_check.py > $OUT/r02_sharded_$N.log 2>&1; echo "sharded check rc=$?"; grep -E "SHARDED|solve " $OUT/r02_sharded_$N.log | cut -c1-180
FSP_GEN_TRACE=1 MASTER_PORT=29560 timeout 300 tools/launch_ranks.sh $N build/examples/transcr_reg_6d --solver krylov --log > $OUT/r02_gen_trace_$N.log 2>&1
tail -2 $OUT/r02_gen_trace_$N.log | cut -c1-300
python - <<PY
import re,collections
t=collections.defaultdict(float); c=collections.Counter()
for l in open("$OUT/r02_gen_trace_$N.log"):
    m=re.match(r"\[gen n=\d+\]\s+(.+?)\s+([\d.]+) ms",l)
    if m: t[m.group(1).strip()]+=float(m.group(2)); c[m.group(1).strip()]+=1
for k,v in t.items(): print("%-40s %8.1f ms over %d" % (k,v,c[k]))
PY
MASTER_PORT=29562 timeout 300 tools/launch_ranks.sh $N build/examples/transcr_reg_6d --solver krylov --log 2>&1 | grep -E "wall_s|timing_s" | cut -c1-300
