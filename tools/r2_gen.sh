#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for i in 1 2; do
for mode in "" "--host-propensities"; do
  FSP_GEN_TRACE=1 timeout 300 build/examples/transcr_reg_6d --solver cvode --log $mode > $OUT/gen_$i$mode.log 2>&1
  tail -1 $OUT/gen_$i$mode.log | cut -c1-200
  python - <<PY
import re,collections
t=collections.defaultdict(float); mx=collections.defaultdict(float)
for l in open("$OUT/gen_$i$mode.log"):
    m=re.match(r"\[gen n=(\d+)\] (.+?)\s+([\d.]+) ms",l)
    if m:
        t[m.group(2).strip()]+=float(m.group(3)); mx[m.group(2).strip()]=max(mx[m.group(2).strip()],float(m.group(3)))
print("mode '$mode':", {k:(round(v,1),round(mx[k],1)) for k,v in t.items()})
PY
done; done
