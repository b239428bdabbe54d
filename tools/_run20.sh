#!/bin/bash
mkdir -p gpurun_out
L=$PWD/sequence-based-recommendations_b200/libsbr_b200_timeline.so
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "cce_gradients or stacked or mixed or 8_row or bidirectional or trajectory" 2>&1 | tail -5
for x in 0 256; do
  SBR_TC_EXPERIMENT=$x timeout 120 python tools/tl_c2.py LSTM 200 120 200
  SBR_TC_EXPERIMENT=$x timeout 120 python tools/tl_c2.py LSTM 200 8 200
done
SBR_B200_LIB=$L SBR_TC_TIMELINE=1 timeout 120 python tools/tl_c2.py LSTM 200 120 200 2>&1 | grep -v "^$"
SBR_B200_LIB=$L SBR_TC_TIMELINE=1 timeout 120 python tools/tl_c2.py LSTM 200 8 200 2>&1 | grep -v "^$"
SBR_TG_TIMELINE=1 timeout 200 python tools/bench_tc.py gemm c2 2>&1 | cut -c1-700
timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/g20_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/g20_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms_all'])
PY
