#!/bin/bash
mkdir -p gpurun_out
L=$PWD/sequence-based-recommendations_b200/libsbr_b200_timeline.so
python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "cce_gradients or stacked or mixed or 8_row or bidirectional" 2>&1 | tail -5
for x in 0 256; do
  SBR_TC_EXPERIMENT=$x python tools/tl_c2.py LSTM 200 120 200
  SBR_TC_EXPERIMENT=$x python tools/tl_c2.py GRU 200 120 200
done
SBR_B200_LIB=$L SBR_TC_TIMELINE=1 python tools/tl_c2.py LSTM 200 120 200 2>&1 | grep -v "^$"
SBR_B200_LIB=$L SBR_TC_TIMELINE=1 SBR_TC_EXPERIMENT=256 python tools/tl_c2.py LSTM 200 120 200 2>&1 | grep -v "^$"
python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-400
