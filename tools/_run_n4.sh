#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_c2_4gpu.json 2> gpurun_out/r2f_bench_c2_4gpu.err
timeout 400 $TR --master-port 29522 bench.py --gpus 4 --config c4 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_c4_4gpu.json 2> gpurun_out/r2f_bench_c4_4gpu.err
for f in gpurun_out/r2f_bench_c2_4gpu gpurun_out/r2f_bench_c4_4gpu; do cut -c1-300 $f.json; tail -2 $f.err; done
