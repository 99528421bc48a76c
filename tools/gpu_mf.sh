#!/bin/bash
# sparse multifrontal path: parity tests, then the per-kernel breakdown of the n = 90 000 grid problem (cluster vs per-block)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_kkt.py -x -q ) > gpurun_out/pytest_sparse.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sparse.log
tail -15 gpurun_out/pytest_sparse.log
( timeout 600 python tools/prof_config.py grid300 ) > gpurun_out/prof_grid300_cluster.txt 2>&1
head -16 gpurun_out/prof_grid300_cluster.txt
( QPALM_B200_MF_PER_BLOCK=1 timeout 600 python tools/prof_config.py grid300 ) > gpurun_out/prof_grid300_block.txt 2>&1
head -4 gpurun_out/prof_grid300_block.txt
