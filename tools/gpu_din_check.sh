#!/bin/bash
# K4 with the key gather fused in: kernel + model tests, cfg4 step time and kernel table.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "din or DIN" > gpurun_out/din_tests.log 2>&1
echo "rc=$?" >> gpurun_out/din_tests.log
grep -n "passed\|failed\|FAILED\|Error\|rc=" gpurun_out/din_tests.log | head -20
timeout 120 python tools/bench_models.py din 2>&1 | tail -n 1
PTREC_DIN_FUSED_GATHER=0 timeout 120 python tools/bench_models.py din 2>&1 | tail -n 1
timeout 120 python tools/profile_step.py din r2_din_ids > gpurun_out/r2_step_kernels_din_ids.txt 2>&1
head -n 16 gpurun_out/r2_step_kernels_din_ids.txt | cut -c1-150
