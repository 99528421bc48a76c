#!/bin/bash
# generator-form rank-k update (updown_gen.cu): parity tests, micro-benchmark against the dataflow sweep, C3 / C1 solves
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_updown.py -x -q -m gpu ) > gpurun_out/updown_gen_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/updown_gen_tests.log
( timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "rank_k" ) >> gpurun_out/updown_gen_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/updown_gen_tests.log
tail -12 gpurun_out/updown_gen_tests.log
for n in 8000 1024; do
  echo "== generator form n=$n"; ( timeout 300 python tools/prof_dense.py $n updown ) 2>&1 | tee -a gpurun_out/updown_gen_bench.txt
  echo "== dataflow sweep n=$n"; ( QPALM_B200_UPDOWN_GEN=0 timeout 300 python tools/prof_dense.py $n updown ) 2>&1 | tee -a gpurun_out/updown_gen_bench.txt
done
( timeout 600 python tools/c3_trace.py ) > gpurun_out/c3_trace_gen.txt 2>&1
grep -v trace gpurun_out/c3_trace_gen.txt | tail -4
( timeout 300 python tools/prof_config.py c1 ) > gpurun_out/prof_c1_gen.txt 2>&1
head -12 gpurun_out/prof_c1_gen.txt
