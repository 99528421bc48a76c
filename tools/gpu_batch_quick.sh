#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_batch.py -x -q -m gpu ) > gpurun_out/batch_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/batch_tests.log
( timeout 300 python tools/batch_phases.py 512 ) > gpurun_out/batch_phases.txt 2>&1
( QPALM_B200_BATCH_UPDOWN=0 timeout 300 python tools/batch_phases.py 512 ) > gpurun_out/batch_phases_noupdown.txt 2>&1
tail -3 gpurun_out/batch_tests.log; head -20 gpurun_out/batch_phases.txt; head -3 gpurun_out/batch_phases_noupdown.txt
