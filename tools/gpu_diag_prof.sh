#!/bin/bash
# ncu source-level capture of the 128x128 diagonal-block kernel + event timing of potrf and latency microbenchmarks.
mkdir -p gpurun_out
python tools/microbench.py > gpurun_out/microbench.txt 2>&1
python tools/prof_dense.py 8000 potrf > gpurun_out/potrf8000.txt 2>&1
python tools/prof_dense.py 2048 potrf >> gpurun_out/potrf8000.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_diag_block -s 3 -c 1 -f -o gpurun_out/full_diag \
   python tools/prof_dense.py 2048 potrf > gpurun_out/ncu_full_diag.log 2>&1
cat gpurun_out/microbench.txt gpurun_out/potrf8000.txt
