#!/bin/bash
for x in 0 1024 4096 5120 8192 9216 12288 13312 16 48 112; do
  SBR_TC_EXPERIMENT=$x timeout 120 python tools/tl_c2.py LSTM 200 8 200 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['env'], d['fwd_cycles_per_step'], d['bwd_cycles_per_step'])
"
done
