#!/bin/bash
# full GPU suite + driver bench at N GPUs (sharded state sets are the multi-GPU default)
OUT=gpurun_out
N=${1:-2}
timeout 2400 python -m pytest tests -q -m gpu -x > $OUT/r02_pytest_sharded_default_$N.log 2>&1; echo "pytest -m gpu rc=$?"; tail -5 $OUT/r02_pytest_sharded_default_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r02_bench_sharded_default_$N.json 2> $OUT/r02_bench_sharded_default_$N.err; echo "bench rc=$?"
python - <<PY
import json
try:
    j=json.loads(open("$OUT/r02_bench_sharded_default_$N.json").read().strip().splitlines()[-1])
    print({k:j[k] for k in ("value","ms_per_step","n_gpus")}, "parity", j.get("parity",{}).get("ok"), j.get("parity",{}).get("max_rel_err"), "e2e", j["e2e"]["value"], "solve", {k:(v.get("wall_s") if isinstance(v,dict) else v) for k,v in j.get("solve_to_tf",{}).items()})
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/r02_bench_sharded_default_$N.err").read()[-2000:])
PY
