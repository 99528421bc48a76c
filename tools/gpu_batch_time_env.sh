#!/bin/bash
# device time of the 512-instance sweep under each env setting ("-" = defaults)
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "-" ]; then echo "== defaults"; timeout 300 python tools/batch_time.py 512 8
  else echo "== $v"; env $v timeout 300 python tools/batch_time.py 512 8; fi
done 2>&1 | tee gpurun_out/batch_time_env.txt
