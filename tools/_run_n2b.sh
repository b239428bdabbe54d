#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f_bench_c2_2gpu.json 2> gpurun_out/r2f_bench_c2_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_c2_2gpu.json'))
print(d['value'], d['ms_per_step'], d.get('per_rank'), d.get('multi_rank_cost_check'))
PY
tail -3 gpurun_out/r2f_bench_c2_2gpu.err
