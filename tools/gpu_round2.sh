#!/bin/bash
# full GPU parity suite + smoke + bench (product, reference) + batch phase clocks
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
( timeout 300 python tools/batch_phases.py 512 ) > gpurun_out/batch_phases.txt 2>&1
( time timeout 1500 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/batch_phases.txt; tail -c 6000 gpurun_out/bench.log; tail -c 1500 gpurun_out/bench_ref.log
