#!/bin/bash
# sharded construction at N ranks (ranks without states, uneven blocks) + the driver's bench command
OUT=gpurun_out
N=${1:-8}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 tests/multirank_sharded_check.py > $OUT/r02_sharded_$N.log 2>&1
echo "sharded check rc=$?"; grep -E "sharded set|Expand\(\)|SHARDED|solve |action parity|remembered" $OUT/r02_sharded_$N.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r02_bench_sharded_default_$N.json 2> $OUT/r02_bench_sharded_default_$N.err; echo "bench rc=$?"
python - <<PY
import json
try:
    j=json.loads(open("$OUT/r02_bench_sharded_default_$N.json").read().strip().splitlines()[-1])
    print({k:j[k] for k in ("value","ms_per_step","n_gpus")}, "parity", j.get("parity",{}).get("ok"), j.get("parity",{}).get("max_rel_err"), "e2e", j["e2e"]["value"], "solve", {k:(v.get("wall_s") if isinstance(v,dict) else v) for k,v in j.get("solve_to_tf",{}).items()})
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/r02_bench_sharded_default_$N.err").read()[-2000:])
PY
