#!/bin/bash
timeout 900 python -m pytest tests -q -x -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 200 python tools/bench_tc.py gemm c2 2>&1 | grep "^{" | cut -c1-200
for e in "" "SBR_GATHER_NO_TMA=1"; do
env $e timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/g27_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/g27_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms_all'])
PY
done
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/g27_bench_c3.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/g27_bench_c3.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms_all'])
PY
