#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sparse.py -x -q ) > gpurun_out/pytest_sparse.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sparse.log
tail -25 gpurun_out/pytest_sparse.log
( timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sparse.py -x -q -k "grid31 and (factor or 5-3)" ) > gpurun_out/sanitizer_sparse.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/sanitizer_sparse.log
tail -8 gpurun_out/sanitizer_sparse.log
( SPARSE_TIME_PROF=1 timeout 900 python tools/sparse_time.py 100 300 ) > gpurun_out/sparse_time.log 2>&1
cat gpurun_out/sparse_time.log
