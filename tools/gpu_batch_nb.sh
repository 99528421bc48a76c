#!/bin/bash
# phase clocks of the persistent batch engine at several batch sizes (CTAs per SM): isolates intrinsic chain latency from contention
mkdir -p gpurun_out
for nb in "$@"; do
  ( timeout 300 python tools/batch_phases.py $nb ) > gpurun_out/batch_phases_nb$nb.txt 2>&1
  echo "== nb=$nb"; head -2 gpurun_out/batch_phases_nb$nb.txt; grep -E "sweep|potrf: diag16|solve  |gemv|sort" gpurun_out/batch_phases_nb$nb.txt
done
