#!/bin/bash
# Last short call of the round: the double-buffered fp16x2 GEMM (PTREC_TC_BN=128).
mkdir -p gpurun_out
export PTREC_TC_BN=128
timeout 60 python -m pytest tests/test_gpu_tc_h2.py -m gpu -x -q > gpurun_out/db_tests.log 2>&1
echo "rc=$?" >> gpurun_out/db_tests.log
timeout 60 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/db_bench.json 2> gpurun_out/db_bench.err
echo "rc=$?" >> gpurun_out/db_bench.err
timeout 40 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "ctr_models or golden_ctr or cuda_graph or dcn or din" > gpurun_out/db_models.log 2>&1
echo "rc=$?" >> gpurun_out/db_models.log
timeout 30 python tools/bench_tc_linear.py > gpurun_out/db_bench_tc.log 2>&1
tail -n 3 gpurun_out/db_tests.log gpurun_out/db_models.log 2>/dev/null
