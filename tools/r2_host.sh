#!/bin/bash
for c in ${CHUNKS:-32}; do
FSP_HOST_TRACE=1 FSP_HOST_CHUNKS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-solve --no-cpu-baseline --no-extra --no-parity 2>/dev/null > gpurun_out/host_trace_$c.log
grep "host pipeline" gpurun_out/host_trace_$c.log | head -40
tail -1 gpurun_out/host_trace_$c.log | python -c "
import sys,json
j=json.loads(sys.stdin.read()); e=j['e2e']
print('chunks $c: e2e %.1f GB/s, %.2f ms/step, copies only %.2f ms, frac %.3f' % (e['value'], e['ms_per_step'], e['copies_only_ms_per_step'], e['frac_of_link_floor']))"
done
