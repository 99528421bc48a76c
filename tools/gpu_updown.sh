#!/bin/bash
# rank-k dataflow sweep: parity tests, micro-benchmark, chain stage clocks, C3 trace
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_updown.py -x -q -m gpu ) > gpurun_out/updown_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/updown_tests.log
( timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "rank_k" ) >> gpurun_out/updown_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/updown_tests.log
( timeout 300 python tools/prof_dense.py 8000 updown ) > gpurun_out/updown_bench.txt 2>&1
( timeout 120 python tools/prof_dense.py 2048 updown ) >> gpurun_out/updown_bench.txt 2>&1
( timeout 120 python tools/prof_dense.py 8000 updown_clocks ) >> gpurun_out/updown_bench.txt 2>&1
( timeout 600 python tools/c3_trace.py ) > gpurun_out/c3_trace.txt 2>&1
tail -15 gpurun_out/updown_tests.log; cat gpurun_out/updown_bench.txt; grep -v trace gpurun_out/c3_trace.txt | tail -5
