#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_examples_small.py -q -m gpu -s -k warm > $OUT/r02_warm_test.log 2>&1; echo rc=$?; grep -E "cold|passed|failed|L1 vs" $OUT/r02_warm_test.log | grep -v print | cut -c1-250
for ex in repressilator hog1p transcr_reg_6d; do
  for mode in "" "--warm"; do timeout 300 build/examples/$ex --solver cvode --log $mode 2>&1 | tail -2 | cut -c1-330; done
  FSP_WARM_RESTART=carry timeout 300 build/examples/$ex --solver cvode --warm 2>&1 | tail -1 | cut -c1-330
done > $OUT/r02_examples_taylor.log 2>&1; cat $OUT/r02_examples_taylor.log
