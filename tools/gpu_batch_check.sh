#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_batch.py -x -q ) > gpurun_out/pytest_batch.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_batch.log
( timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_batch.py -x -q -k "small and persistent" ) > gpurun_out/sanitizer_batch.log 2>&1
( time timeout 600 python bench.py --steps 3 --warmup 3 --no-dense ) > gpurun_out/bench_batchp.log 2>&1
( QPALM_B200_BATCH_ENGINE=lockstep timeout 600 python bench.py --steps 2 --warmup 3 --no-dense --no-cpu ) > gpurun_out/bench_batch_lockstep.log 2>&1
tail -5 gpurun_out/pytest_batch.log; tail -5 gpurun_out/sanitizer_batch.log; tail -2 gpurun_out/bench_batchp.log; tail -1 gpurun_out/bench_batch_lockstep.log
