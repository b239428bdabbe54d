#!/bin/bash
P='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(d["B"], d["env"], d["fwd_cycles_per_step"], d["bwd_cycles_per_step"])
'
O=$PWD/sequence-based-recommendations_b200/libsbr_b200_old.so
for b in 8 120; do
  SBR_B200_LIB=$O timeout 120 python tools/tl_c2.py LSTM 200 $b 200 2>&1 | python -c "$P"
  SBR_B200_LIB=$O SBR_TC_EXPERIMENT=16 timeout 120 python tools/tl_c2.py LSTM 200 $b 200 2>&1 | python -c "$P"
  SBR_B200_LIB=$O SBR_TC_EXPERIMENT=112 timeout 120 python tools/tl_c2.py LSTM 200 $b 200 2>&1 | python -c "$P"
done
SBR_TC_EXPERIMENT=12288 timeout 120 python tools/tl_c2.py LSTM 200 120 200 2>&1 | python -c "$P"
SBR_TC_EXPERIMENT=12400 timeout 120 python tools/tl_c2.py LSTM 200 8 200 2>&1 | python -c "$P"
