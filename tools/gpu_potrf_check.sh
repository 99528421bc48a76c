#!/bin/bash
# quick check of the dense factorization path: potrf timings + the parity tests that exercise it
mkdir -p gpurun_out
python tools/prof_dense.py 8000 potrf > gpurun_out/potrf.txt 2>&1
python tools/prof_dense.py 2048 potrf >> gpurun_out/potrf.txt 2>&1
python tools/prof_dense.py 1024 potrf >> gpurun_out/potrf.txt 2>&1
python tools/prof_dense.py 2048 potrf_prof >> gpurun_out/potrf.txt 2>&1
python tools/prof_dense.py 8000 potrf_prof >> gpurun_out/potrf.txt 2>&1
cat gpurun_out/potrf.txt
if [ -z "$NOTEST" ]; then ( timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_solve.py tests/test_gpu_batch.py -x -q -m gpu ) > gpurun_out/pytest_sub.log 2>&1
tail -5 gpurun_out/pytest_sub.log; fi
