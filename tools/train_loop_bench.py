#!/usr/bin/env python
"""Throughput of the *caller's* loop on C2: `batch = next(generator); cost = train_function(*batch)` exactly as
RNNBase.train runs it (rnn_base.py:289-290), with the Python batch assembly inside the timed region -- plain, and
with the background prefetch (`threaded_generator`, --prefetch).  One JSON line per mode."""
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    cfg = bench.CONFIGS["c2"]
    ds = bench.make_dataset(cfg)
    pred = bench.make_predictor(cfg, ds)
    pred._compile_train_function()
    from sbr_b200.neural_networks.rnn_base import threaded_generator
    devnull = open(os.devnull, "w")
    for mode in ("plain", "prefetch"):
        random.seed(7)
        np.random.seed(7)
        stdout, sys.stdout = sys.stdout, devnull          # "Opening file" chatter of the generator
        try:
            gen = pred._gen_mini_batch(ds.training_set())
            if mode == "prefetch":
                gen = threaded_generator(gen, num_cached=32)
            for _ in range(20):
                pred.train_function(*next(gen))
            pred.engine.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                cost = pred.train_function(*next(gen))
            pred.engine.synchronize()
            dt = time.perf_counter() - t0
            if mode == "prefetch":
                gen.close()
        finally:
            sys.stdout = stdout
        print(json.dumps({"loop": mode, "steps": steps, "ms_per_step": dt / steps * 1e3,
                          "sequences_per_s": cfg["B"] * steps / dt, "last_cost": float(cost)}), flush=True)


if __name__ == "__main__":
    main()
