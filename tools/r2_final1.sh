#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1700 python -m pytest tests -x -q -m gpu > $OUT/r02_pytest_final_1.log 2>&1; echo rc=$?; tail -3 $OUT/r02_pytest_final_1.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench default flags"; ( time timeout 900 python bench.py > $OUT/r02_bench_default.json 2> $OUT/r02_bench_default.err ) 2>&1 | grep real; echo rc=$?; python - <<PY
import json
j=json.loads(open("$OUT/r02_bench_default.json").read().strip().splitlines()[-1])
print({k:j[k] for k in ("value","ms_per_step","steps","warmup","gpu_launches")}, j["roofline"]["frac"], j["roofline"]["traffic"])
print("e2e",{k:j["e2e"][k] for k in ("value","ms_per_step","copies_only_ms_per_step","frac_of_link_floor","bit_identical_to_device_action")})
print("cpu",j["cpu_baseline"]); print("parity ok",j["parity"]["ok"],j["parity"]["max_rel_err"])
print("other", {k:(round(v.get("frac",0),3) if "frac" in v else v) for k,v in j["roofline_other_operators"].items()})
print("solve", {k:(v.get("wall_s"),v.get("action_calls"),v.get("l1_err_vs_poisson")) for k,v in j["solve_to_tf"].items() if isinstance(v,dict)})
PY
echo "== bench driver flags + reference arm"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r02_bench_driver_1.json 2>/dev/null; cut -c1-160 $OUT/r02_bench_driver_1.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/r02_ref_driver_1.json 2>/dev/null; cut -c1-200 $OUT/r02_ref_driver_1.json
