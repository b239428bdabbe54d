#!/bin/bash
P='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["cell"], d["H"], d["B"], d["env"], d["fwd_cycles_per_step"], d["bwd_cycles_per_step"], d["stage_ms"])
'
timeout 900 python -m pytest tests -q -x -m gpu -p no:cacheprovider 2>&1 | tail -3
timeout 120 python tools/tl_c2.py LSTM 200 8 200 2>&1 | python -c "$P"
timeout 120 python tools/tl_c2.py LSTM 200 120 200 2>&1 | python -c "$P"
timeout 200 python tools/bench_tc.py gemm c2 2>&1 | grep "^{" | cut -c1-200
timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/g26_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/g26_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms_all'])
PY
