"""Per-step cycle timeline of the H <= 224 cluster scans (run on the GPU box with the --timeline build):
    SBR_B200_LIB=$PWD/sequence-based-recommendations_b200/libsbr_b200_timeline.so SBR_TC_TIMELINE=1 python tools/tl_c2.py [cell] [H] [B] [T]
Without the timeline build it still prints the stage times of a step (fwd / bwd scan, ms and cycles per step at the
given clock)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sbr_b200 import _capi  # noqa: E402

cell = sys.argv[1] if len(sys.argv) > 1 else "LSTM"
H = int(sys.argv[2]) if len(sys.argv) > 2 else 200
B = int(sys.argv[3]) if len(sys.argv) > 3 else 120
T = int(sys.argv[4]) if len(sys.argv) > 4 else 200
N = 3706
rng = np.random.RandomState(1)
e = _capi.Engine(n_items=N, cell=cell, layers=(H,), max_length=T, batch_size=B)
e.set_all_param_values([rng.normal(0, 0.05, size=s).astype(np.float32) for _, s in e.param_infos()])
X = rng.randint(0, N, size=(B, T, 1)).astype(np.int32)
mask = np.ones((B, T), np.float32)
Y = rng.randint(0, N, size=B).astype(np.int32)
pop = np.ones(B, np.float32)
for _ in range(10):
    c = e.train_step_cce(X, mask, Y, pop)
e.set_profiling(True)
acc = {}
for _ in range(5):
    c = e.train_step_cce(X, mask, Y, pop)
    for k, v in e.stage_times().items():
        acc[k] = acc.get(k, 0.0) + v / 5
mhz = float(os.environ.get("SM_MHZ", "1965"))
print(json.dumps({"cell": cell, "H": H, "B": B, "T": T, "env": os.environ.get("SBR_TC_EXPERIMENT", ""), "cost": float(c),
                  "stage_ms": {k: round(v, 4) for k, v in acc.items()},
                  "fwd_cycles_per_step": round(acc["rnn_fwd"] * 1e3 * mhz / T), "bwd_cycles_per_step": round(acc["rnn_bwd"] * 1e3 * mhz / T)}), flush=True)
e.close()
