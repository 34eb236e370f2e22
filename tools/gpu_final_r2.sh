#!/bin/bash
# End-of-round evidence: full GPU suite, smoke(), the bench line, the ncu launch list of the bench command and a full
# capture of K5 on the CTA-pair kernel (each program first exits 0 without ncu).
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/full_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/full_gpu.log
tail -n 3 gpurun_out/full_gpu.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "rc=$?" >> gpurun_out/smoke.log
tail -n 2 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-cfg5 --no-models"
timeout 200 $CMD > gpurun_out/ncu_plain_bench.json 2> gpurun_out/ncu_plain_bench.err || { echo "plain bench failed"; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r2_launches_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_bench_summary.txt 2>&1
head -12 gpurun_out/r2_launches_bench_summary.txt
timeout 100 python tools/run_dcn_once.py > /dev/null 2>&1 || { echo "run_dcn_once failed"; exit 1; }
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_split3_2sm_kernel' --launch-skip 3 -c 3 \
    -o gpurun_out/r2_ncu_full_k5_pair -f python tools/run_dcn_once.py > gpurun_out/ncu_full_k5.log 2>&1
echo "k5 capture rc=$?"
ncu -i gpurun_out/r2_ncu_full_k5_pair.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_k5_pair_raw.csv 2>> gpurun_out/ncu_full_k5.log
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -n 8
