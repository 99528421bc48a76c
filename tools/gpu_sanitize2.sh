#!/bin/bash
# compute-sanitizer memcheck + racecheck on the kernels added in the second half of round 2: udgen::k_fwd_multi (tagged packets,
# cooperative launch), k_gen_* / k_gen_apply, mfc::k_mf_front (cluster barriers)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for which in dense front; do
  for tool in memcheck racecheck; do
    log=gpurun_out/sanitizer2_${tool}_${which}.txt
    ( timeout 600 $CS --tool $tool --print-limit 20 python tools/sanitize_case.py $which ) > $log 2>&1
    echo "exit $?" >> $log
    echo "== $tool $which"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit |^batch|^dense|^sparse|^front" $log | tail -4
  done
done
