#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1700 python -m pytest tests -q -m gpu -s > $OUT/r02_pytest_gpu_1.log 2>&1; echo rc=$?; grep -E "passed|failed|FAILED|Error|to t=|\|\|p_|tight|warm|marginal|FIM" $OUT/r02_pytest_gpu_1.log | cut -c1-260 | head -60
echo "== examples: defaults (warm restart, device propensities) vs reference-like (--cold --host-propensities)"
for ex in repressilator hog1p transcr_reg_6d; do
  timeout 300 build/examples/$ex --solver cvode --log 2>&1 | tail -2 | cut -c1-330
  timeout 300 build/examples/$ex --solver cvode --log --cold --host-propensities 2>&1 | tail -2 | cut -c1-330
  timeout 300 build/examples/$ex --solver krylov --log 2>&1 | tail -2 | cut -c1-330
done > $OUT/r02_examples_warm.log 2>&1; cat $OUT/r02_examples_warm.log
