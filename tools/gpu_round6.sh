#!/bin/bash
# Closing pass of round 2: parity suite, smoke, default bench line, sparse / C1 / C3 kernel breakdowns after the pipelined sparse solves.
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -q -m gpu -rs ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/smoke.log 2>&1
( time timeout 1200 python bench.py --steps 5 --warmup 3 ) > gpurun_out/bench.log 2>&1
( timeout 600 python tools/prof_config.py grid300; QPALM_B200_MF_PER_BLOCK=1 QPALM_B200_MF_FWD_GENERIC=1 timeout 600 python tools/prof_config.py grid300 ) > gpurun_out/prof_grid300_final.txt 2>&1
( timeout 300 python tools/prof_config.py c3; timeout 300 python tools/prof_config.py c1 ) > gpurun_out/prof_c3_c1_final.txt 2>&1
tail -6 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log | head -1; grep "^grid300\|^c3\|^c1" gpurun_out/prof_grid300_final.txt gpurun_out/prof_c3_c1_final.txt; tail -1 gpurun_out/bench.log | cut -c1-400
