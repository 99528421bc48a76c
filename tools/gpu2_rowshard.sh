#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
( time timeout 900 python -m pytest tests/test_gpu_rowshard.py -x -q ) > gpurun_out/pytest_rowshard.log 2>&1
tail -15 gpurun_out/pytest_rowshard.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload dense --steps 2 --warmup 3 --no-cpu ) > gpurun_out/bench_dense_2gpu.log 2>&1
tail -3 gpurun_out/bench_dense_2gpu.log | cut -c1-1800
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu ) > gpurun_out/bench_batch_2gpu.log 2>&1
tail -3 gpurun_out/bench_batch_2gpu.log | cut -c1-1500
