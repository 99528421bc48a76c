#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python tools/batch_phases.py 148 > gpurun_out/phases148.log 2>&1
python tools/batch_phases.py 512 > gpurun_out/phases512.log 2>&1
( timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu ) > gpurun_out/bench_quick.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/phases148.log; head -1 gpurun_out/phases512.log; tail -1 gpurun_out/bench_quick.log | cut -c1-2500
