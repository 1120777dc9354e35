#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== test_examples_small"; timeout 1200 python -m pytest tests/test_examples_small.py -x -q -m gpu -s > $OUT/r02_examples_small.log 2>&1; echo rc=$?; grep -E "to t=|\|\|p|passed|failed|Error|assert" $OUT/r02_examples_small.log | cut -c1-250
echo "== traces"; timeout 600 python tools/trace_solves.py $OUT > $OUT/r02_trace_solves.log 2>&1; echo rc=$?; grep "==== done" $OUT/r02_trace_solves.log
echo "== lattice krylov 128 verbose"; build/examples/lattice_solve --edge 128 --solver krylov --verbose > $OUT/r02_krylov128_trace.log 2>&1; tail -1 $OUT/r02_krylov128_trace.log | cut -c1-300
echo "== examples again"; for i in 1 2; do build/examples/repressilator --solver cvode --log | tail -2 | cut -c1-300; done
echo "== ncu launch list of the bench step (no parity leg)"
python bench.py --steps 2 --warmup 3 --no-solve --no-cpu-baseline --no-extra --no-parity > $OUT/r02_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-solve --no-cpu-baseline --no-extra --no-parity > $OUT/r02_ncu_bench.log 2>&1; echo rc=$?
