#!/bin/bash
# Source-level (per-line stall samples) capture of the K4 backward at the cfg4 shape.
mkdir -p gpurun_out
timeout 100 python tools/run_din_once.py || { echo "plain run failed"; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'din_bwd_tc_kernel' --launch-skip 2 -c 1 \
    -o gpurun_out/din_bwd -f python tools/run_din_once.py > gpurun_out/ncu_din_bwd.log 2>&1
echo "capture rc=$?"
ncu -i gpurun_out/din_bwd.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_ncu_din_bwd_source_sass.csv 2>> gpurun_out/ncu_din_bwd.log
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -3
