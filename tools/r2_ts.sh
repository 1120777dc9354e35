#!/bin/bash
OUT=gpurun_out
timeout 600 build/tests/test_ode 2>&1 | grep -E "RUN|OK|FAILED|tests ran" | tail -12
timeout 900 build/tests/test_fsp_solver 2>&1 | grep -E "OK|FAILED|tests ran|Poisson|marginal" | tail -14
timeout 900 build/tests/test_mat 2>&1 | grep -E "FAILED|tests ran" | tail -3
timeout 1200 python -m pytest tests/test_examples_small.py -q -m gpu -s -k "tight" 2>&1 | grep -E "p_tight|passed|failed|Error" | cut -c1-230
echo "== krylov orth kernel: parity + timing"
timeout 900 python -m pytest tests/test_oracle_krylov.py tests/test_examples_small.py -q -m gpu -s -k "krylov" 2>&1 | grep -E "p_gpu|passed|failed|Error|to t=" | cut -c1-200
for ex in hog1p repressilator transcr_reg_6d; do
  timeout 300 build/examples/$ex --solver krylov --log 2>&1 | tail -2 | cut -c1-300
  FSP_KRYLOV_ORTH=0 timeout 300 build/examples/$ex --solver krylov --log 2>&1 | tail -2 | cut -c1-300
done
build/examples/lattice_solve --edge 100 --solver krylov --repeat 3 | tail -1 | cut -c1-250
FSP_KRYLOV_ORTH=0 build/examples/lattice_solve --edge 100 --solver krylov --repeat 3 | tail -1 | cut -c1-250
