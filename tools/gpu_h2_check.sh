#!/bin/bash
# One short GPU call: validate the fp16x2 operand format of K6 (opt-in), the HBM-resident reader, and time both formats.
# Usage (from the repo root, through gpurun):  bash tools/gpu_h2_check.sh
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/h2_gpu.txt 2>&1
PTREC_TC_MODE=fp16x2 timeout 200 python -m pytest tests/test_gpu_tc_h2.py tests/test_reader.py -m gpu -x -q \
    > gpurun_out/h2_tests.log 2>&1
echo "rc=$?" >> gpurun_out/h2_tests.log
timeout 60 python tools/bench_tc_linear.py > gpurun_out/h2_bench_tc.log 2>&1
echo "rc=$?" >> gpurun_out/h2_bench_tc.log
BK=64 timeout 60 python tools/bench_tc_linear.py > gpurun_out/h2_bench_tc_bk64.log 2>&1
PTREC_TC_MODE=fp16x2 timeout 150 python -m pytest tests/test_gpu_models.py -m gpu -x -q \
    -k "ctr_models or golden_ctr or cuda_graph or dense_layer or prefetched or staged or reference_models" \
    > gpurun_out/h2_models.log 2>&1
echo "rc=$?" >> gpurun_out/h2_models.log
PTREC_TC_MODE=fp16x2 timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline \
    > gpurun_out/h2_bench.json 2> gpurun_out/h2_bench.err
echo "rc=$?" >> gpurun_out/h2_bench.err
timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/h2_bench_bf16x3.json 2> gpurun_out/h2_bench_bf16x3.err
tail -n 3 gpurun_out/h2_tests.log; tail -n 3 gpurun_out/h2_models.log; head -n 60 gpurun_out/h2_bench_tc.log
