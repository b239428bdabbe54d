"""Run BASELINE.json configs 3-5 at their full per-GPU sizes for a few steps (sanity + timing)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sbr_b200 import _capi

def batch(rng, B, T, N):
    base = rng.randint(0, N, T)
    lens = np.sort(rng.randint(2, T + 1, B))
    X = np.zeros((B, T, 1), np.int32); mask = np.zeros((B, T), np.float32)
    for b in range(B):
        X[b, :lens[b], 0] = base[:lens[b]]; mask[b, :lens[b]] = 1
    return X, mask, lens

def run(name, steps=3, **kw):
    rng = np.random.RandomState(0)
    B, T, N = kw["batch_size"], kw["max_length"], kw["n_items"]
    t0 = time.time()
    eng = _capi.Engine(**kw)
    # small random init
    vals = eng.get_all_param_values()
    for v in vals:
        if v.ndim == 2:
            v[...] = rng.normal(0, 0.05, size=v.shape).astype(np.float32)
    eng.set_all_param_values(vals)
    print("%s: %.2f M params, create+init %.1f s" % (name, eng.total_params() / 1e6, time.time() - t0), flush=True)
    X, mask, lens = batch(rng, B, T, N)
    Y = rng.randint(0, N, B).astype(np.int32); pop = np.ones(B, np.float32)
    eng.set_profiling(True)
    for s in range(steps):
        t0 = time.time()
        if kw["loss"] == "CCE":
            c = eng.train_step_cce(X, mask, Y, pop)
        elif kw["loss"] in ("BPR", "TOP1", "Blackout"):
            c = eng.train_step_sampled(X, mask, Y, rng.randint(0, N, kw["n_samples"]).astype(np.int32), pop)
        else:
            off = np.arange(B + 1, dtype=np.int32); w = np.full(B, 1.0 / N, np.float32)
            c = eng.train_step_margin(X, mask, off, Y, w, None, True)
        dt = time.time() - t0
        st = eng.stage_times()
        print("  step %d cost %.5f  %.1f ms  (%.0f seq/s)  stages %s" % (s, c, dt * 1e3, B / dt, {k: round(v, 2) for k, v in st.items()}), flush=True)
        assert np.isfinite(c)
    eng.close()

if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c4", "c5"]
    if "c3" in which:
        run("C3 LSTM-2x256 N=50k T=200 B=512 BPR S=32", cell="LSTM", layers=(256, 256), n_items=50000, max_length=200, batch_size=512, loss="BPR", n_samples=32)
    if "c4" in which:
        run("C4/4 LSTM-1x512 N=200k T=200 B=256 hinge", cell="LSTM", layers=(512,), n_items=200000, max_length=200, batch_size=256, loss="hinge")
    if "c5" in which:
        run("C5/8 GRU-2x512 N=500k T=500 B=256 CCE", cell="GRU", layers=(512, 512), n_items=500000, max_length=500, batch_size=256, loss="CCE")
