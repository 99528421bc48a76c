#!/bin/bash
# TMA (bulk copy + mbarrier producer warp) GEMM vs the cp.async version: parity tests, SYRK / Cholesky throughput, C3 time
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_updown.py tests/test_gpu_batch.py -x -q -m gpu ) > gpurun_out/gemm_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/gemm_tests.log
{
for mode in tma ldgsts; do
  if [ $mode = ldgsts ]; then export QPALM_B200_GEMM_LDGSTS=1; else unset QPALM_B200_GEMM_LDGSTS; fi
  echo "== $mode"
  timeout 120 python tools/prof_dense.py 8000 syrk 1408
  timeout 120 python tools/prof_dense.py 8000 syrk 4000
  timeout 120 python tools/prof_dense.py 8000 potrf
  timeout 120 python tools/prof_dense.py 4096 potrf
  timeout 300 python tools/c3_trace.py 2>/dev/null | grep rep
done
} > gpurun_out/gemm_ab.txt 2>&1
tail -3 gpurun_out/gemm_tests.log; cat gpurun_out/gemm_ab.txt
