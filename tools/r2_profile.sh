#!/bin/bash
# 1-GPU evidence run: event-timed rooflines of every operator, then ncu (full set) on the kernels that ship.
OUT=gpurun_out; mkdir -p $OUT
echo "== event-timed rooflines (no profiler)"
python tools/profile_actions.py --steps 20 --warmup 3 > $OUT/r02_rooflines.json 2> $OUT/r02_rooflines.err; echo rc=$?; cat $OUT/r02_rooflines.json | head -c 3000; echo
echo "== examples (1 GPU, --log)"
for ex in repressilator hog1p transcr_reg_6d; do for s in cvode krylov; do timeout 300 build/examples/$ex --solver $s --log 2>&1 | tail -2; done; done > $OUT/r02_examples_1.log 2>&1; cat $OUT/r02_examples_1.log | cut -c1-330
echo "== ncu launch list of the bench step"
python bench.py --steps 2 --warmup 3 --no-solve --no-cpu-baseline --no-extra > $OUT/r02_bench_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-solve --no-cpu-baseline --no-extra > $OUT/r02_ncu_bench.log 2>&1; echo rc=$?
echo "== ncu full: Action kernels on all operators"
python tools/profile_actions.py --steps 1 --warmup 1 > $OUT/r02_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fsp_action_lean -c 16 -o $OUT/r02_action_full -f python tools/profile_actions.py --steps 1 --warmup 1 > $OUT/r02_ncu_actions.log 2>&1; echo rc=$?; tail -3 $OUT/r02_ncu_actions.log
echo "== ncu full: fused-epilogue kernel inside a CVODE solve (215^3)"
build/examples/lattice_solve --edge 215 --solver cvode > $OUT/r02_cvode_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fsp_action_epi -s 30 -c 2 -o $OUT/r02_epi_full -f build/examples/lattice_solve --edge 215 --solver cvode > $OUT/r02_ncu_epi.log 2>&1; echo rc=$?
echo "== ncu launch list of a Krylov solve (215^3)"
build/examples/lattice_solve --edge 215 --solver krylov > $OUT/r02_krylov_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file $OUT/r02_krylov_launches.csv build/examples/lattice_solve --edge 215 --solver krylov > $OUT/r02_ncu_krylov.log 2>&1; echo rc=$?
ls -la $OUT | grep r02_
