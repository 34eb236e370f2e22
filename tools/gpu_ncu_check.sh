#!/bin/bash
# ncu launch list of one eager cfg2 step + one full capture of the K6 kernels in the default operand format.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline"
timeout 120 $CMD > gpurun_out/ncu_plain_bench.json 2> gpurun_out/ncu_plain_bench.err || exit 1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/r1_launches_fp16x2.csv $CMD > gpurun_out/ncu_bench.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_bench.log
timeout 60 python tools/run_tc_once.py > gpurun_out/run_tc_once.log 2>&1 || exit 1
timeout 200 ncu --set full --clock-control none --import-source on \
    -k regex:'gemm_split3_2sm_kernel|split_kernel|absmax_kernel' --launch-skip 8 -c 3 \
    -o gpurun_out/r1_ncu_full_k6_fp16x2 -f python tools/run_tc_once.py > gpurun_out/ncu_full.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_full.log
ncu -i gpurun_out/r1_ncu_full_k6_fp16x2.ncu-rep --page raw --csv > gpurun_out/r1_ncu_full_k6_fp16x2_raw.csv 2>> gpurun_out/ncu_full.log
ls -la gpurun_out | tail -n 12
