#!/bin/bash
# Source-level (per-instruction stall samples) capture of the K6 fused GEMMs.
mkdir -p gpurun_out
timeout 100 python tools/run_tc_fused_once.py || { echo "plain run failed"; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_split3_2sm_kernel' --launch-skip 3 -c 3 \
    -o gpurun_out/k6_fused -f python tools/run_tc_fused_once.py > gpurun_out/ncu_k6_fused.log 2>&1
echo "capture rc=$?"
ncu -i gpurun_out/k6_fused.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_k6_fused_raw.csv 2>> gpurun_out/ncu_k6_fused.log
ncu -i gpurun_out/k6_fused.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_ncu_k6_fused_source.csv 2>> gpurun_out/ncu_k6_fused.log
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -5
