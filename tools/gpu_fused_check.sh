#!/bin/bash
# K6 fused tower: its own tests, the per-layer K6 tests (the staged epilogue serves both), then the bench line.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_tc_fused.py -x -q > gpurun_out/fused_tests.log 2>&1
echo "rc=$?" >> gpurun_out/fused_tests.log
tail -n 25 gpurun_out/fused_tests.log
timeout 240 python -m pytest tests/test_gpu_tc_h2.py tests/test_gpu_kernels.py -x -q -k "tc or gemm or split or dense or mlp" > gpurun_out/k6_tests.log 2>&1
echo "rc=$?" >> gpurun_out/k6_tests.log
tail -n 6 gpurun_out/k6_tests.log
timeout 200 python bench.py --no-cpu-baseline --no-cfg5 --no-models > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_fused.json').read().strip().splitlines()[-1])
    print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
    print('roofline', d['roofline']['seconds_per_launch'], d['roofline']['frac'])
except Exception as e:
    print('no bench line', e)
PY
tail -n 5 gpurun_out/bench_fused.err
