#!/bin/bash
OUT=gpurun_out
timeout 600 build/tests/test_ode 2>&1 | grep -E "RUN|OK|FAILED|tests ran" | tail -12
timeout 900 build/tests/test_fsp_solver 2>&1 | grep -E "OK|FAILED|tests ran|Poisson|marginal" | tail -14
timeout 900 build/tests/test_mat 2>&1 | grep -E "FAILED|tests ran" | tail -3
timeout 1200 python -m pytest tests/test_examples_small.py -q -m gpu -s -k "tight" 2>&1 | grep -E "p_tight|passed|failed|Error" | cut -c1-230
