#!/bin/bash
# batch engine A/B: parity tests, then phase clocks with the default settings and with the listed env variants
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_batch.py -x -q -m gpu ) > gpurun_out/batch_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/batch_tests.log
( timeout 300 python tools/batch_phases.py 512 ) > gpurun_out/batch_phases.txt 2>&1
for v in "$@"; do
  ( env $v timeout 300 python tools/batch_phases.py 512 ) > "gpurun_out/batch_phases_$v.txt" 2>&1
done
tail -3 gpurun_out/batch_tests.log; head -24 gpurun_out/batch_phases.txt
for v in "$@"; do echo "== $v"; head -24 "gpurun_out/batch_phases_$v.txt"; done
