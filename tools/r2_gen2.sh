#!/bin/bash
OUT=gpurun_out
FSP_GEN_TRACE=1 MASTER_PORT=29560 timeout 300 tools/launch_ranks.sh 2 build/examples/transcr_reg_6d --solver krylov --log > $OUT/r02_gen_trace_2.log 2>&1
tail -2 $OUT/r02_gen_trace_2.log | cut -c1-300
python - <<PY
import re,collections
t=collections.defaultdict(float); c=collections.Counter()
for l in open("$OUT/r02_gen_trace_2.log"):
    m=re.match(r"\[gen n=\d+\]\s+(.+?)\s+([\d.]+) ms",l)
    if m: t[m.group(1).strip()]+=float(m.group(2)); c[m.group(1).strip()]+=1
for k,v in t.items(): print("%-40s %8.1f ms over %d" % (k,v,c[k]))
PY
grep "\[gen" $OUT/r02_gen_trace_2.log | tail -12
