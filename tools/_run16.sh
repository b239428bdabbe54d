mkdir -p gpurun_out
nvidia-smi -L | head -3
(timeout 600 python -m pytest tests/test_gpu_e2e.py -x -q -m gpu -p no:cacheprovider --timeout 300 -k "two_rank" > gpurun_out/t16_tests.log 2>&1; echo "rc=$?" >> gpurun_out/t16_tests.log); tail -5 gpurun_out/t16_tests.log
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_c2_2gpu.json 2> gpurun_out/r2_bench_c2_2gpu.err; tail -3 gpurun_out/r2_bench_c2_2gpu.err; cut -c1-2500 gpurun_out/r2_bench_c2_2gpu.json
SBR_NO_NCCL_REGISTER=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_c2_2gpu_noreg.json 2> gpurun_out/r2_bench_c2_2gpu_noreg.err; python -c "
import json
for f in ('r2_bench_c2_2gpu','r2_bench_c2_2gpu_noreg'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms_all'], d.get('multi_rank_cost_check'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-600
