#!/bin/bash
# K5 on the CTA-pair GEMM: DCN kernel / model tests in both kernel modes, kernel timings, cfg3 step time and kernel table.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "dcn or cross or row_dot" > gpurun_out/dcn_tests.log 2>&1
echo "rc=$?" >> gpurun_out/dcn_tests.log
timeout 300 python -m pytest tests -m gpu -x -q -k "dcn or cross or DCN or row_dot or tower" >> gpurun_out/dcn_tests.log 2>&1
echo "rc=$?" >> gpurun_out/dcn_tests.log
grep -n "passed\|failed\|FAILED\|Error\|rc=" gpurun_out/dcn_tests.log | head -20
timeout 120 python tools/bench_dcn_kernels.py > gpurun_out/dcn_kernels.txt 2>&1
cat gpurun_out/dcn_kernels.txt | tail -n 12
timeout 120 python tools/bench_models.py dcn 2>&1 | tail -n 2
timeout 120 python tools/profile_step.py dcn r2_dcn_pair > gpurun_out/r2_step_kernels_dcn_pair.txt 2>&1
head -n 14 gpurun_out/r2_step_kernels_dcn_pair.txt
