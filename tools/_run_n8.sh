#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_c2_8gpu.json 2> gpurun_out/r2f_bench_c2_8gpu.err
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,ENV timeout 300 $TR --master-port 29532 bench.py --gpus 8 --config c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_c5_8gpu.json 2> gpurun_out/r2f_bench_c5_8gpu.err
grep -i -m 12 "nvls\|algo\|channels\|via P2P" gpurun_out/r2f_bench_c5_8gpu.err | cut -c1-200 > gpurun_out/r2f_nccl_info_8gpu.txt
for f in gpurun_out/r2f_bench_c2_8gpu gpurun_out/r2f_bench_c5_8gpu; do cut -c1-300 $f.json; done; cat gpurun_out/r2f_nccl_info_8gpu.txt; tail -3 gpurun_out/r2f_bench_c5_8gpu.err | cut -c1-300
