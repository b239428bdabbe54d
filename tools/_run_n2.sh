#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_e2e.py -q -x -p no:cacheprovider -k "two_rank or nccl" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_c2_2gpu.json 2> gpurun_out/r2f_bench_c2_2gpu.err
cut -c1-600 gpurun_out/r2f_bench_c2_2gpu.json; tail -3 gpurun_out/r2f_bench_c2_2gpu.err
