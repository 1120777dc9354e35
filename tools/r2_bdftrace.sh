#!/bin/bash
OUT=gpurun_out
FSP_BDF_TRACE=1 build/examples/repressilator --solver cvode > $OUT/bdf_cold.log 2>&1
FSP_BDF_TRACE=1 build/examples/repressilator --solver cvode --warm > $OUT/bdf_warm.log 2>&1
tail -1 $OUT/bdf_cold.log | cut -c1-250; tail -1 $OUT/bdf_warm.log | cut -c1-250
echo COLD; grep "\[bdf\]" $OUT/bdf_cold.log | tail -25 | cut -c1-200
echo WARM; grep "\[bdf\]" $OUT/bdf_warm.log | tail -25 | cut -c1-200
