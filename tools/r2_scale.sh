#!/bin/bash
# tools/r2_scale.sh N : the N-GPU validation + scaling evidence of round 2 on one box (N = 2, 4 or 8)
N=${1:-8}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multirank_check (fused peer-memory path) x$N"; timeout 600 $TR --master-port 29541 tests/multirank_check.py > $OUT/r02_check_fused_$N.log 2>&1; echo rc=$?; grep -E "peer-memory|parity|poisson|toggle|MULTIRANK|rror" $OUT/r02_check_fused_$N.log | head -20
echo "== bench 1 GPU (same box)"; timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-solve --no-cpu-baseline --no-extra --no-parity > $OUT/r02_scale_1_of_$N.json 2> /dev/null; cut -c1-200 $OUT/r02_scale_1_of_$N.json
echo "== bench x$N (driver command; e2e + solve_to_tf included)"; timeout 900 $TR --master-port 29545 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --trace-steps > $OUT/r02_scale_$N.json 2> $OUT/r02_scale_$N.err; echo rc=$?
cut -c1-200 $OUT/r02_scale_$N.json; grep "trace rank" $OUT/r02_scale_$N.err | cut -c1-110; grep -o '"e2e".*"gpu_launches": [0-9]*' $OUT/r02_scale_$N.json | cut -c1-330; grep -o '"solve_to_tf".*' $OUT/r02_scale_$N.json | cut -c1-700
echo "== NVLink counters around 2000 Actions (GPU 0)"
nvidia-smi nvlink -gt d -i 0 > $OUT/r02_nvlink_before_$N.txt 2>&1
timeout 600 $TR --master-port 29546 bench.py --gpus $N --steps 2000 --warmup 20 --no-solve --no-cpu-baseline --no-e2e --no-parity > $OUT/r02_scale_${N}_long.json 2>/dev/null; cut -c1-200 $OUT/r02_scale_${N}_long.json
nvidia-smi nvlink -gt d -i 0 > $OUT/r02_nvlink_after_$N.txt 2>&1
python - <<PY
import re
def tot(p):
    tx=rx=0
    for l in open(p):
        m=re.search(r"Data (Tx|Rx): (\d+) KiB", l)
        if m:
            if m.group(1)=="Tx": tx+=int(m.group(2))
            else: rx+=int(m.group(2))
    return tx,rx
try:
    a=tot("$OUT/r02_nvlink_before_$N.txt"); b=tot("$OUT/r02_nvlink_after_$N.txt")
    print("GPU0 NVLink data over the run: Tx %.1f MiB, Rx %.1f MiB  -> per Action (2020 Actions): Tx %.1f KiB, Rx %.1f KiB" % ((b[0]-a[0])/1024,(b[1]-a[1])/1024,(b[0]-a[0])/2020,(b[1]-a[1])/2020))
except Exception as e: print("nvlink counters unavailable", e)
PY
echo "== PCIe ceiling (N GPUs copying both ways at once)"; timeout 300 $TR --master-port 29547 tools/pcie_ceiling.py > $OUT/r02_pcie_$N.log 2>&1; grep -v "^$\|\*\*\*\|OMP_NUM" $OUT/r02_pcie_$N.log | head -40
timeout 300 $TR --master-port 29548 tools/pcie_ceiling.py --bind > $OUT/r02_pcie_bind_$N.log 2>&1; grep "rank\|aggregate" $OUT/r02_pcie_bind_$N.log | head -12
echo "== config 3: transcr_reg_6d partitioned across $N GPUs"
for s in cvode krylov; do MASTER_PORT=29560 timeout 300 tools/launch_ranks.sh $N build/examples/transcr_reg_6d --solver $s --log 2>&1 | tail -2 | cut -c1-330; done | tee $OUT/r02_transcr_$N.log
echo "== lattice 215^3 solves x$N"
for s in krylov cvode; do MASTER_PORT=29570 timeout 300 tools/launch_ranks.sh $N build/examples/lattice_solve --edge 215 --solver $s --repeat 2 2>&1 | tail -1 | cut -c1-330; done | tee $OUT/r02_lattice215_$N.log
