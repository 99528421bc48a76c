#!/bin/bash
# phase clocks of the 512-instance sweep under each of the given env settings (A/B of engine knobs)
mkdir -p gpurun_out
for v in "$@"; do
  ( env $v timeout 300 python tools/batch_phases.py 512 ) > "gpurun_out/batch_phases_$v.txt" 2>&1
  echo "== $v"; head -2 "gpurun_out/batch_phases_$v.txt"; grep -E "H syrk|potrf  |rank-k" "gpurun_out/batch_phases_$v.txt"
done
