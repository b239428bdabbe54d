import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SBR_TG_TIMELINE"] = "1"
from sbr_b200 import _capi
e = _capi.Engine(n_items=16, cell="GRU", layers=(8,), max_length=4, batch_size=2)
rng = np.random.RandomState(0)
for (ta, tb, M, N, K) in [(False, True, 4096, 4096, 4096), (True, True, 4096, 4096, 4096), (False, False, 4096, 4096, 4096), (True, False, 4096, 4096, 4096)]:
    A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
    B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
    for _ in range(2):
        _, ms = e.debug_gemm(A, B, ta=ta, tb=tb, engine=1, reps=1)
    print("ms", ms, flush=True)
e.close()
