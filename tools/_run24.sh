#!/bin/bash
P='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["cell"], d["H"], d["B"], d["env"], d["fwd_cycles_per_step"], d["bwd_cycles_per_step"])
'
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "cce_gradients or stacked or mixed or 8_row or bidirectional or trajectory" 2>&1 | tail -3
for b in 8 120; do
  for x in 0 4096 8192 12288; do
    SBR_TC_EXPERIMENT=$x timeout 120 python tools/tl_c2.py LSTM 200 $b 200 2>&1 | python -c "$P"
  done
done
timeout 120 python tools/tl_c2.py GRU 200 120 200 2>&1 | python -c "$P"
timeout 120 python tools/tl_c2.py GRU 96 120 200 2>&1 | python -c "$P"
timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/g24_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/g24_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms_all'])
PY
