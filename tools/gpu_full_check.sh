#!/bin/bash
# Full GPU suite in the default configuration, smoke(), and one bench line (through gpurun).
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -x -q > gpurun_out/full_gpu.log 2>&1
echo "rc=$?" >> gpurun_out/full_gpu.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "rc=$?" >> gpurun_out/smoke.log
timeout 150 python bench.py --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "rc=$?" >> gpurun_out/bench_default.err
tail -n 4 gpurun_out/full_gpu.log; tail -n 2 gpurun_out/smoke.log
