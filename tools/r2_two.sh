#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== generation trace, transcr_reg_6d on 2 ranks"
FSP_GEN_TRACE=1 MASTER_PORT=29560 timeout 300 tools/launch_ranks.sh 2 build/examples/transcr_reg_6d --solver krylov --log > $OUT/r02_gen_trace_2.log 2>&1; tail -2 $OUT/r02_gen_trace_2.log | cut -c1-300
grep "\[gen" $OUT/r02_gen_trace_2.log | tail -12
python - <<PY
import re,collections
t=collections.defaultdict(float)
for l in open("$OUT/r02_gen_trace_2.log"):
    m=re.match(r"\[gen n=\d+\] (.+?)\s+([\d.]+) ms",l)
    if m: t[m.group(1).strip()]+=float(m.group(2))
print({k:round(v,1) for k,v in t.items()})
PY
echo "== 1 rank for comparison"; FSP_GEN_TRACE=1 timeout 300 build/examples/transcr_reg_6d --solver krylov --log > $OUT/r02_gen_trace_1.log 2>&1; tail -2 $OUT/r02_gen_trace_1.log | cut -c1-300
python - <<PY
import re,collections
t=collections.defaultdict(float)
for l in open("$OUT/r02_gen_trace_1.log"):
    m=re.match(r"\[gen n=\d+\] (.+?)\s+([\d.]+) ms",l)
    if m: t[m.group(1).strip()]+=float(m.group(2))
print({k:round(v,1) for k,v in t.items()})
PY
FSP_GEN_TRACE=1 timeout 300 build/examples/hog1p --solver cvode --log > $OUT/r02_gen_trace_hog1p.log 2>&1; tail -2 $OUT/r02_gen_trace_hog1p.log | cut -c1-300
python - <<PY
import re,collections
t=collections.defaultdict(float)
for l in open("$OUT/r02_gen_trace_hog1p.log"):
    m=re.match(r"\[gen n=\d+\] (.+?)\s+([\d.]+) ms",l)
    if m: t[m.group(1).strip()]+=float(m.group(2))
print({k:round(v,1) for k,v in t.items()})
PY
echo "== pytest -m gpu (2 GPUs: multi-rank tests included)"; timeout 1700 python -m pytest tests -q -m gpu -x > $OUT/r02_pytest_gpu_2.log 2>&1; echo rc=$?; tail -4 $OUT/r02_pytest_gpu_2.log
echo "== lattice 215^3 krylov: 1 vs 2 GPUs"
build/examples/lattice_solve --edge 215 --solver krylov --repeat 3 | tail -1 | cut -c1-260
MASTER_PORT=29570 tools/launch_ranks.sh 2 build/examples/lattice_solve --edge 215 --solver krylov --repeat 3 | tail -1 | cut -c1-260
