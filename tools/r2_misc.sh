#!/bin/bash
OUT=gpurun_out
timeout 900 build/tests/test_sensfsp_solver 2>&1 | grep -E "OK|FAILED|tests ran|fixed-set|FIM|Failure|failed" | tail -12
for c in 16 32 64 128; do
  FSP_HOST_CHUNKS=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-solve --no-cpu-baseline --no-extra --no-parity 2>/dev/null | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=j['e2e']
print('chunks $c: e2e %.1f GB/s, %.2f ms/step, copies only %.2f ms, frac %.3f, identical %s' % (e['value'], e['ms_per_step'], e['copies_only_ms_per_step'], e['frac_of_link_floor'], e['bit_identical_to_device_action']))"
done
