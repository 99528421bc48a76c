#!/bin/bash
# One short gpurun call: parity tests, smoke, default bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
( time timeout 1200 python -m pytest tests -x -q -m gpu --durations=15 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench.log 2>&1
tail -25 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -5 gpurun_out/bench.log | cut -c1-3000
