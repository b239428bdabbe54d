"""Micro-benchmarks of the tcgen05 GEMM kernel and of the per-step scans (run on the GPU box):
    python tools/bench_tc.py gemm        # device time of representative GEMM shapes, tensor-core vs FFMA
    python tools/bench_tc.py scan        # stage times of one training step at C3 / C4-shard / C5-shard layer shapes
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sbr_b200 import _capi  # noqa: E402


def gemm():
    e = _capi.Engine(n_items=16, cell="GRU", layers=(8,), max_length=4, batch_size=2)
    rng = np.random.RandomState(0)
    shapes = [  # name, ta, tb, M, N, K
        ("c2 logits  h W^T", False, True, 128, 3706, 200),
        ("c2 dh      d W", False, False, 128, 200, 3706),
        ("c2 dW_out  d^T h", True, False, 3706, 200, 128),
        ("c3 in-gemm hs W_in", False, False, 65536, 1024, 256),
        ("c3 dgrad   dXg W_in^T", False, True, 65536, 256, 1024),
        ("c3 wgrad   hs^T dXg", True, False, 256, 1024, 65536),
        ("c5 wgrad   hs^T dXg", True, False, 512, 1536, 65536),
        ("c5 logits  (N=100k slice)", False, True, 256, 100000, 512),
        ("square 4096", False, True, 4096, 4096, 4096),
    ]
    out = []
    for name, ta, tb, M, N, K in shapes:
        if len(sys.argv) > 2 and not any(name.startswith(c) for c in sys.argv[2].split(",")):
            continue
        A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
        B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
        reps = 5
        _, ms1 = e.debug_gemm(A, B, ta=ta, tb=tb, engine=1, reps=reps)
        _, ms1 = e.debug_gemm(A, B, ta=ta, tb=tb, engine=1, reps=reps)
        _, ms0 = e.debug_gemm(A, B, ta=ta, tb=tb, engine=0, reps=reps)
        fl = 2.0 * M * N * K
        out.append({"shape": name, "M": M, "N": N, "K": K, "tc_ms": ms1 / reps, "ffma_ms": ms0 / reps,
                    "tc_tflops_fp32_equiv": fl / (ms1 / reps) / 1e9, "ffma_tflops": fl / (ms0 / reps) / 1e9})
        print(json.dumps(out[-1]), flush=True)
    e.close()


def scan():
    rng = np.random.RandomState(1)
    for name, cell, layers, B, T, N in [("c2 LSTM 1x200 B128 T200", "LSTM", (200,), 128, 200, 3706),
                                        ("c2g GRU 1x200 B128 T200", "GRU", (200,), 128, 200, 3706),
                                        ("c3 LSTM 2x256 B512 T200", "LSTM", (256, 256), 512, 200, 2000),
                                        ("c4-shard LSTM 1x512 B256 T200", "LSTM", (512,), 256, 200, 2000),
                                        ("c5-shard GRU 2x512 B256 T500", "GRU", (512, 512), 256, 500, 2000)]:
        envs = [{}, {"SBR_DISABLE_STEP_SCAN": "1"}]
        if len(sys.argv) > 2 and not any(name.startswith(c) for c in sys.argv[2].split(",")):
            continue
        if len(sys.argv) > 3:
            envs = [dict(kv.split("=") for kv in e.split(",")) if e != "-" else {} for e in sys.argv[3:]]
        for env in envs:
            for k in ("SBR_DISABLE_STEP_SCAN", "SBR_SCAN_FENCE", "SBR_DISABLE_PERSISTENT_SCAN", "SBR_DISABLE_SPLITK_SCAN",
                      "SBR_TC_EXPERIMENT"):
                os.environ.pop(k, None)
            os.environ.update(env)
            e = _capi.Engine(n_items=N, cell=cell, layers=layers, max_length=T, batch_size=B)
            vals = [rng.normal(0, 0.05, size=s).astype(np.float32) for _, s in e.param_infos()]
            e.set_all_param_values(vals)
            lens = np.sort(rng.randint(2, T + 1, size=B))
            lens[-1] = T
            X = rng.randint(0, N, size=(B, T, 1)).astype(np.int32)
            mask = (np.arange(T)[None, :] < lens[:, None]).astype(np.float32)
            Y = rng.randint(0, N, size=B).astype(np.int32)
            pop = np.ones(B, np.float32)
            e.train_step_cce(X, mask, Y, pop)
            e.set_profiling(True)
            t0 = time.time()
            c = e.train_step_cce(X, mask, Y, pop)
            wall = time.time() - t0
            st = e.stage_times()
            print(json.dumps({"config": name, "env": env, "cost": float(c), "wall_ms": wall * 1e3, "valid_steps": int(lens.sum()),
                              "stage_ms": {k: round(v, 3) for k, v in st.items()}}), flush=True)
            e.close()
    os.environ.pop("SBR_DISABLE_STEP_SCAN", None)


if __name__ == "__main__":
    {"gemm": gemm, "scan": scan}[sys.argv[1]]()
