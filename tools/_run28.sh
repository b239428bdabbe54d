#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "tma_staged or rating or cce_gradients" 2>&1 | tail -3
timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/g28_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/g28_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms_all'])
PY
