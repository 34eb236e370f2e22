#!/bin/bash
# Round-2 ncu evidence (B200_PROFILING.md recipe): (1) the launch list of the bench command — each program first exits 0
# WITHOUT ncu; (2) one `--set full` capture per dominant kernel family.  Outputs under gpurun_out/ (copied to profiles/).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-cfg5 --no-models"
timeout 200 $CMD > gpurun_out/ncu_plain_bench.json 2> gpurun_out/ncu_plain_bench.err || { echo "plain bench failed"; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_bench_summary.txt 2>&1
head -30 gpurun_out/r2_launches_bench_summary.txt
# embedding kernels (K1, K2a one-sweep, K2b) at the cfg2 shape
timeout 100 python tools/run_update_once.py > /dev/null 2>&1 || { echo "run_update_once failed"; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:'gather_onehot_kernel|fused_update_kernel|os_pass_kernel|os_hist_kernel|os_dedup_kernel' --launch-skip 10 -c 8 \
    -o gpurun_out/r2_ncu_full_embedding -f python tools/run_update_once.py > gpurun_out/ncu_full_emb.log 2>&1
echo "embedding capture rc=$?"
ncu -i gpurun_out/r2_ncu_full_embedding.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_embedding_raw.csv 2>> gpurun_out/ncu_full_emb.log
# K6 fused-tower GEMMs (forward: planes + bit mask out; input gradient: bit mask in, planes + column sums out; plain fp32)
# at the cfg2 hidden-layer shape
timeout 100 python tools/run_tc_fused_once.py > /dev/null 2>&1 || { echo "run_tc_fused_once failed"; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_split3_2sm_kernel' --launch-skip 3 -c 3 \
    -o gpurun_out/r2_ncu_full_k6_fused -f python tools/run_tc_fused_once.py > gpurun_out/ncu_full_k6.log 2>&1
echo "k6 capture rc=$?"
ncu -i gpurun_out/r2_ncu_full_k6_fused.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_k6_fused_raw.csv 2>> gpurun_out/ncu_full_k6.log
# K4 on tensor cores at the cfg4 shape
timeout 100 python tools/run_din_once.py > /dev/null 2>&1 || { echo "run_din_once failed"; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on \
    -k regex:'din_fwd_tc_kernel|din_bwd_tc_kernel' --launch-skip 2 -c 2 \
    -o gpurun_out/r2_ncu_full_din_tc -f python tools/run_din_once.py > gpurun_out/ncu_full_din.log 2>&1
echo "din capture rc=$?"
ncu -i gpurun_out/r2_ncu_full_din_tc.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_din_tc_raw.csv 2>> gpurun_out/ncu_full_din.log
rm -f gpurun_out/*.ncu-rep   # the raw pages are what is kept (the reports exceed the transfer limit)
ls -la gpurun_out | tail -n 14
