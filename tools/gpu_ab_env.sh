#!/bin/bash
# A/B of an environment switch on the bench line: tools/gpu_ab_env.sh VAR  (full GPU suite first, then VAR=1 / 0 alternating)
VAR=$1
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests.log 2>&1
echo "rc=$?" >> gpurun_out/ab_tests.log
tail -n 3 gpurun_out/ab_tests.log
for p in 1 0 1 0; do
  env $VAR=$p timeout 200 python bench.py --no-cpu-baseline --no-cfg5 > gpurun_out/bench_ab$p.json 2> gpurun_out/bench_ab$p.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_ab$p.json')); print('$VAR=$p', round(d['ms_per_step'],4), round(d['value']/1e6,2), 'e2e', round(d['e2e']['ms_per_step'],4), 'cfg3', round(d['models']['cfg3_dcn']['ms_per_step'],3), 'cfg4', round(d['models']['cfg4_din']['ms_per_step'],3))"
done
