#!/usr/bin/env python
"""bench.py -- training throughput of the RNN hot path in user-sequences/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c2]

A "step" is one call of the reference's `train_function(*batch)` (neural_networks/rnn_base.py:290):
gather -> LSTM scan -> full softmax + CCE -> BPTT -> scatter -> (all-reduce) -> Adam, on one
synthetic mini-batch.  Workload (BASELINE.json configs[1], the configuration `metric` is quoted
on): RNNOneHot, LSTM 1x200, MovieLens-1M-shaped synthetic data (6040 users, 3706 items, mean
sequence length ~165), max_length 200, 128 rows per GPU, Adam, fp32.  Batches come from the
host mirror of `_gen_mini_batch` (nested prefixes of one user's sequence).

One JSON line on stdout (rank 0):
  value      whole-job sequences/sec with the batches already resident in HBM (device slots),
             timed on the device with a cudaEvent pair on the library's stream, max over ranks;
  e2e        the same metric through the public API (`RNNOneHot.train_function(X, mask, Y, pop)`)
             with HOST numpy buffers: pinned staging + H2D of the inputs and the D2H read of the
             cost are inside the timed region of every step;
  roofline   the dominant kernel (largest stage of the step), algorithmic FLOPs / its measured
             launch time, against MEASURED_PEAKS.json;
  cpu_baseline  the numpy restatement of the reference's Theano CPU path (oracle/, "port") timed on
             this box's host cores on a bounded sample of the same batches (rank 0, N=1 only).

`--impl reference` times that CPU restatement alone (the reference itself -- Python 2 + Theano +
Lasagne -- cannot be installed here) with all host threads numpy's BLAS will use.
"""
import argparse
import json
import os
import random
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: model + data shape
    "c1": dict(label="C1 RNNOneHot GRU-1x100, 500 items, 200 users, seq-len<=20, batch 16",
               cell="GRU", layers=(100,), n_items=500, n_users=200, T=20, B=16, uniform_len=(5, 40)),
    "c2": dict(label="C2 RNNOneHot LSTM-1x200, ML-1M shape (3706 items, 6040 users, mean len 165), max_length 200, "
                     "batch 128/GPU, full softmax + CCE, Adam",
               cell="LSTM", layers=(200,), n_items=3706, n_users=6040, T=200, B=128, uniform_len=None),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------
def make_dataset(cfg):
    from sbr_b200.helpers import synthetic
    from sbr_b200.helpers.data_handling import DataHandler
    tag = "sbr_bench_%s_%d_%d" % (cfg["cell"], cfg["n_users"], cfg["n_items"])
    d = os.path.join(tempfile.gettempdir(), tag)
    if not os.path.exists(os.path.join(d, "data", "stats")):
        tmp = d + ".%d.tmp" % os.getpid()
        kw = dict(uniform_len=cfg["uniform_len"]) if cfg["uniform_len"] else {}
        synthetic.write_dataset(tmp, cfg["n_users"], cfg["n_items"], seed=1234, **kw)
        try:
            os.rename(tmp, d)
        except OSError:
            pass  # another rank won the race
    return DataHandler(d + "/")


def make_predictor(cfg, dataset, n_ranks=1, rank=0, nccl_id=None, device=0, n_slots=1, create_engine=True):
    from sbr_b200.neural_networks.recurrent_layers import RecurrentLayers
    from sbr_b200.neural_networks.rnn_one_hot import RNNOneHot
    from sbr_b200.neural_networks.update_manager import Adam
    p = RNNOneHot(recurrent_layer=RecurrentLayers(layer_type=cfg["cell"], layers=list(cfg["layers"])),
                  updater=Adam(), max_length=cfg["T"], batch_size=cfg["B"] * n_ranks,
                  use_ratings_features=False, use_movies_features=False, use_users_features=False,
                  device=device, n_ranks=n_ranks, rank=rank, nccl_id=nccl_id)
    p._engine_extra = dict(n_slots=n_slots)
    base = p._engine_extra_kwargs
    p._engine_extra_kwargs = lambda: dict(base(), n_slots=n_slots)
    if create_engine:
        p.prepare_model(dataset)
    else:
        p.n_items = dataset.n_items
    p.set_dataset(dataset)
    return p


def make_batches(predictor, dataset, n):
    """n global mini-batches from the host mirror of _gen_mini_batch (same on every rank)."""
    random.seed(1234)
    np.random.seed(1234)
    devnull = open(os.devnull, "w")
    stdout, sys.stdout = sys.stdout, devnull      # "Opening file (n)" chatter of the generator
    try:
        gen = predictor._gen_mini_batch(dataset.training_set())
        return [next(gen) for _ in range(n)]
    finally:
        sys.stdout = stdout
        devnull.close()


def step_flops(cfg, batch):
    """Algorithmic FLOPs of one step on this batch: 3 x forward, forward =
    2*V*H*G*H per layer (+ input GEMM for layers >= 1) + 2*B*H*N, V = valid (b, t) pairs."""
    mask = batch[1]
    V = float(mask.sum())
    B = mask.shape[0]
    G = 4 if cfg["cell"] == "LSTM" else 3
    fwd = 0.0
    prev = None
    for H in cfg["layers"]:
        fwd += 2.0 * V * H * G * H
        if prev is not None:
            fwd += 2.0 * V * prev * G * H
        prev = H
    fwd += 2.0 * B * cfg["layers"][-1] * cfg["n_items"]
    return 3.0 * fwd, V


# ----------------------------------------------------------------------------------------------
# CPU arm: numpy restatement of the reference graph (oracle/), float32 like a tuned Theano run
# ----------------------------------------------------------------------------------------------
def cpu_reference_rate(cfg, dataset, batches, steps, warmup, budget_s=25.0):
    from oracle import sbr_oracle as O
    spec = O.Spec(n_items=cfg["n_items"], cell=cfg["cell"], layers=tuple(cfg["layers"]), loss="CCE")
    vals = O.init_params(spec, np.random.RandomState(1), np.float32)
    upd = O.Updater("adam", lr=1e-3)
    times = []
    t_begin = time.perf_counter()
    n_done = 0
    for i in range(warmup + steps):
        X, mask, Y, pop, _ = batches[i % len(batches)]
        t0 = time.perf_counter()
        cost = O.train_step(spec, vals, upd, X, mask, Y=Y, pop=pop)
        dt = time.perf_counter() - t0
        if not np.isfinite(cost):
            raise RuntimeError("oracle cost is not finite")
        if i >= warmup:
            times.append(dt)
            n_done += 1
        if time.perf_counter() - t_begin > budget_s and len(times) >= 1:
            break
    B = batches[0][0].shape[0]
    mean = float(np.mean(times))
    return B / mean, mean, n_done


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    W = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    K = args.steps

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log("warning: WORLD_SIZE=%d but --gpus %d; using WORLD_SIZE" % (world, args.gpus))
    n_gpus = world

    base = {"metric": "user-sequences/sec (training step, device-timed)", "unit": "sequences/s", "n_gpus": n_gpus,
            "steps": K, "warmup": W, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (ML-1M-shaped, seed 1234), random-init weights",
            "config": {"workload": cfg["label"], "global_batch": cfg["B"] * n_gpus, "seq_len": cfg["T"],
                       "parallelism": "dp%d" % n_gpus,
                       "l2": "no explicit flush: a step streams ~330 MB of activations (> 126 MB L2) and every "
                             "step uses a different batch"}}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        dataset = make_dataset(cfg)
        pred = make_predictor(cfg, dataset, create_engine=False)
        n_b = min(K + W, 8)
        batches = make_batches(pred, dataset, n_b)
        steps = max(1, min(K, 12))
        rate, sec, done = cpu_reference_rate(cfg, dataset, batches, steps, min(W, 1), budget_s=150.0)
        cores = blas_threads()
        out = dict(base)
        out.update({"impl": "reference", "n_gpus": n_gpus, "value": rate, "ms_per_step": sec * 1e3, "steps": done,
                    "gpu_launches": 0,
                    "cpu_baseline": {"value": rate, "unit": "sequences/s", "cores": cores, "kind": "port",
                                     "sample": "%d full %s steps (batch %d) of the numpy float32 restatement of the "
                                               "Theano CPU graph, BLAS threads=%d of %d host cores"
                                               % (done, args.config, cfg["B"], cores, os.cpu_count())},
                    "e2e": {"value": rate, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        out["config"] = dict(out["config"], global_batch=cfg["B"],
                             note="Theano/Lasagne (Python 2) cannot be installed here; this is the oracle port")
        print(json.dumps(out))
        return 0

    # ------------------------------------------------------------------ B200 arm
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ["NCCL_DEBUG"] = "NONE"       # stdout carries exactly one JSON line (NCCL prints its banner there)
    from sbr_b200 import _capi
    dist = None
    nccl_id = None
    if n_gpus > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group(backend="gloo")     # control plane only; the gradients use the library's NCCL
        obj = [_capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        nccl_id = obj[0]

    dataset = make_dataset(cfg)
    n_batches = K + W
    pred = make_predictor(cfg, dataset, n_ranks=n_gpus, rank=rank, nccl_id=nccl_id, device=local_rank,
                          n_slots=n_batches)
    pred._compile_train_function()
    eng = pred.engine
    batches = make_batches(pred, dataset, n_batches)
    B_global = cfg["B"] * n_gpus

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- leg 1: batches resident in HBM ------------------------------------------------------
    for i, b in enumerate(batches):
        X, mask, Y, pop, _ = b
        sl = pred._split_rows
        eng.stage_cce(i, sl(X), sl(mask), sl(Y), sl(pop))
    for i in range(W):
        eng.train_step_staged(i, want_cost=False)
    eng.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = eng.kernel_launches()
    eng.timer_start()
    t_enq = time.perf_counter()
    for i in range(W, W + K):
        eng.train_step_staged(i, want_cost=False)
    host_enqueue_ms = (time.perf_counter() - t_enq) * 1e3 / K     # host time to enqueue a step (no sync inside)
    ms = eng.timer_stop()
    barrier()
    launches = eng.kernel_launches() - launches0
    last_cost = eng.synchronize(want_cost=True)
    ms = max_over_ranks(ms)
    value = B_global * K / (ms * 1e-3)

    # ---- leg 2: end to end through the public API, host buffers ------------------------------
    for i in range(W):
        pred.train_function(*batches[i])
    eng.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(W, W + K):
        cost = pred.train_function(*batches[i])
    eng.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None     # sampled across both timed legs
    barrier()
    e2e_value = B_global * K / e2e_s
    Bl, T = cfg["B"], cfg["T"]
    h2d = Bl * T * 4 + Bl * 4 * 3
    d2h = 4

    # ---- leg 3: per-stage device times (separate pass; profiling syncs every step) -------------
    eng.set_profiling(True)
    acc = {}
    n_prof = K                    # the same batches as the timed leg, so the stage times add up to ms_per_step
    for i in range(W, W + n_prof):
        eng.train_step_staged(i, want_cost=False)
        for k, v in eng.stage_times().items():
            acc[k] = acc.get(k, 0.0) + v / n_prof
    eng.set_profiling(False)
    flops = [step_flops(cfg, b) for b in batches[W:W + n_prof]]
    mean_V = float(np.mean([v for _, v in flops])) / n_gpus
    G = 4 if cfg["cell"] == "LSTM" else 3
    H = cfg["layers"][0]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "fallback 1.4 PFLOP/s sustained"
    # FLOPs attributable to each recurrent kernel: forward scan 2*V*H*G*H, backward scan (dh only) the same
    stage_flops = {"rnn_fwd": 2.0 * mean_V * H * G * H * len(cfg["layers"]),
                   "rnn_bwd": 2.0 * mean_V * H * G * H * len(cfg["layers"])}
    dom = max(("rnn_fwd", "rnn_bwd"), key=lambda k: acc.get(k, 0.0))
    dom_ms = acc.get(dom, 0.0)
    achieved = stage_flops[dom] / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
    except Exception:
        pass
    kern = {"rnn_fwd": "rnn_fwd_tc_kernel<4,8>", "rnn_bwd": "rnn_bwd_tc_kernel<4,2,8>"}[dom] if cfg["cell"] == "LSTM" else dom
    roofline = {"kernel": kern, "stage": dom, "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (achieved / peak_tf) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "math": "3xTF32 on tcgen05 (fp32-accurate): 3 MMA passes at the TF32 rate = 1/6 of the bf16 peak per "
                        "algorithmic FLOP",
                "frac_of_3xtf32_peak": (achieved / (peak_tf / 6.0)) if achieved else None,
                "note": "cluster-persistent scan: up to %d strictly sequential steps per launch, 8-row tiles on 8-CTA "
                        "clusters (B=%d -> 16 tiles on the 15 co-resident cluster slots = 120 of 148 SMs); bound by the "
                        "per-step latency chain (DSMEM exchange ~20 B/clk/SM, tcgen05 issue at the tf32 rate with N=16 "
                        "half used, gate math), not by FLOPs or HBM" % (cfg["T"], cfg["B"]),
                "algorithmic_flops_per_launch": stage_flops[dom], "valid_steps_per_launch": mean_V,
                "stage_ms": {k: round(v, 4) for k, v in acc.items()}}

    out = None
    if rank == 0:
        out = dict(base)
        out.update({"impl": "b200", "value": value, "ms_per_step": ms / K, "host_enqueue_ms_per_step": host_enqueue_ms,
                    "clocks": clocks,
                    "e2e": {"value": e2e_value, "unit": "sequences/s", "h2d_bytes_per_step": h2d * n_gpus,
                            "d2h_bytes_per_step": d2h * n_gpus, "ms_per_step": e2e_s / K * 1e3},
                    "gpu_launches": int(launches) * n_gpus, "roofline": roofline,
                    "last_cost": float(last_cost), "e2e_last_cost": float(cost)})
        if n_gpus == 1 and not args.no_cpu_baseline:
            cb_batches = [batches[i] for i in range(min(4, len(batches)))]
            rate, sec, done = cpu_reference_rate(cfg, dataset, cb_batches, 6, 1, budget_s=25.0)
            cores = blas_threads()
            out["cpu_baseline"] = {"value": rate, "unit": "sequences/s", "cores": cores, "kind": "port",
                                   "sample": "%d full steps (batch %d) of the numpy float32 restatement of the "
                                             "reference graph, %.2f s/step, BLAS threads=%d of %d host cores"
                                             % (done, cfg["B"], sec, cores, os.cpu_count())}
        print(json.dumps(out))
    eng.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
