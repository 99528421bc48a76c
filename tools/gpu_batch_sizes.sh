#!/bin/bash
# device time of the resident batch solve at several batch sizes
for nb in "$@"; do timeout 300 python tools/batch_time.py $nb 5; done 2>&1 | tee gpurun_out/batch_sizes.txt
