#!/bin/bash
mkdir -p gpurun_out
L=$PWD/sequence-based-recommendations_b200/libsbr_b200_timeline.so
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_tc_gemm.py -q -x -p no:cacheprovider 2>&1 | tail -5
for x in 0 2048 256; do
  SBR_TC_EXPERIMENT=$x timeout 120 python tools/tl_c2.py LSTM 200 120 200
  SBR_TC_EXPERIMENT=$x timeout 120 python tools/tl_c2.py LSTM 200 8 200
done
SBR_TC_EXPERIMENT=0 timeout 120 python tools/tl_c2.py GRU 200 120 200
SBR_B200_LIB=$L SBR_TC_TIMELINE=1 timeout 120 python tools/tl_c2.py LSTM 200 120 200 2>&1 | grep -v "^$"
SBR_B200_LIB=$L SBR_TC_TIMELINE=1 timeout 120 python tools/tl_c2.py LSTM 200 8 200 2>&1 | grep -v "^$"
timeout 300 python tools/bench_tc.py scan c3,c5 - SBR_SCAN_ACQ_SPIN=1 2>&1 | cut -c1-600
timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/g21_bench.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/g21_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['stage_ms_all'])
PY
