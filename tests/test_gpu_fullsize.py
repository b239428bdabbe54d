"""GPU parity at the BASELINE shapes, against the float64 oracle (not only size-independent properties):

  * C2 at full size -- RNNOneHot LSTM 1x200, 3706 items, B = 128, max_length 200, a real nested-prefix batch from
    the host mirror of `_gen_mini_batch`: step-0 cost and every gradient tensor, then 10 Adam steps of costs.  200
    strictly sequential 3xTF32 steps with SFU gate math are exactly what the short (T <= 20) parity cases do not see.
  * C3-shaped -- RNNSampling BPR, LSTM 2x256 (per-step tensor-core scans + tensor-core input GEMMs), T = 200.
  * C5-shaped -- RNNOneHot GRU 2x512, T = 120, reduced catalog.

Tolerances as everywhere: cost 1e-4 absolute, gradients 2e-4 * max|g| per tensor."""
import os
import sys

import numpy as np
import pytest

from oracle import sbr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _nested_prefix_batch(rng, B, T, N, n_users=3, min_len=2):
    """Rows = nested prefixes of a few users' sequences, ascending length inside a user (rnn_base.py:396-415)."""
    X = np.zeros((B, T, 1), np.int32)
    mask = np.zeros((B, T), np.float32)
    per = B // n_users
    b = 0
    for u in range(n_users):
        n = per if u < n_users - 1 else B - b
        seq = rng.permutation(N)[:T + 40]
        ls = np.sort(rng.choice(np.arange(min_len, T + 40), size=n, replace=False))
        for l in ls:
            w = seq[max(0, l - T):l]
            X[b, :len(w), 0] = w
            mask[b, :len(w)] = 1
            b += 1
    return X, mask


def _compare_step(eng, spec, vals64, X, mask, kw_gpu, kw_ref, step):
    eng.set_skip_update(True)
    cost = step(eng, X, mask, kw_gpu)
    grads = eng.get_all_grads()
    c0, g0 = O.loss_and_grads(spec, vals64, X, mask, **kw_ref)
    assert abs(float(cost) - float(c0)) <= 1e-4, (float(cost), float(c0))
    for (name, _), a, b in zip(O.param_names_shapes(spec), g0, grads):
        tol = 2e-4 * np.abs(a).max() + 1e-7
        err = np.abs(a - b).max()
        assert err <= tol, "%s: max err %.3e > tol %.3e (scale %.3e)" % (name, err, tol, np.abs(a).max())
    eng.set_skip_update(False)
    return float(c0)


def test_c2_full_size_step_and_ten_adam_steps_match_the_oracle(tmp_path):
    sys.path.insert(0, ROOT)
    import bench
    cfg = dict(bench.CONFIGS["c2"])
    dataset = bench.make_dataset(cfg)
    pred = bench.make_predictor(cfg, dataset)
    pred._compile_train_function()
    eng = pred.engine
    try:
        batches = bench.make_batches(pred, dataset, 10)
        spec = O.Spec(n_items=cfg["n_items"], cell="LSTM", layers=(200,), loss="CCE")
        vals = [v.astype(np.float64) for v in eng.get_all_param_values()]
        X, mask, Y, pop, _ = batches[0]
        assert X.shape == (128, 200, 1) and mask.sum(axis=1).max() >= 190      # a real full-length nested-prefix batch
        _compare_step(eng, spec, vals, X, mask, dict(Y=Y, pop=pop), dict(Y=Y, pop=pop.astype(np.float64)),
                      lambda e, X, m, kw: e.train_step_cce(X, m, kw["Y"], kw["pop"]))
        upd = O.Updater("adam", lr=1e-3)
        for i, (X, mask, Y, pop, _) in enumerate(batches):
            c_gpu = pred.train_function(X, mask, Y, pop)
            c_ref = O.train_step(spec, vals, upd, X, mask, Y=Y, pop=pop.astype(np.float64))
            assert abs(float(c_gpu) - float(c_ref)) <= 1e-4, (i, float(c_gpu), float(c_ref))
        for a, b in zip(vals, eng.get_all_param_values()):
            assert np.abs(a - b).max() <= 2e-4
    finally:
        eng.close()


def test_c3_shaped_bpr_two_layer_256_step_matches_the_oracle():
    from sbr_b200 import _capi
    rng = np.random.RandomState(31)
    B, T, N, S = 192, 200, 4000, 32
    spec = O.Spec(n_items=N, cell="LSTM", layers=(256, 256), loss="BPR")
    vals = O.init_params(spec, rng, np.float64)
    X, mask = _nested_prefix_batch(rng, B, T, N)
    Y = rng.randint(0, N, B).astype(np.int32)
    samples = rng.randint(0, N, S).astype(np.int32)
    pop = rng.uniform(0.5, 2.0, B)
    eng = _capi.Engine(n_items=N, cell="LSTM", layers=(256, 256), loss="BPR", max_length=T, batch_size=B, n_samples=S)
    try:
        eng.set_all_param_values(vals)
        _compare_step(eng, spec, vals, X, mask, dict(Y=Y, samples=samples, pop=pop), dict(Y=Y, samples=samples, pop=pop),
                      lambda e, X, m, kw: e.train_step_sampled(X, m, kw["Y"], kw["samples"], kw["pop"]))
        upd = O.Updater("adam", lr=1e-3)
        for i in range(2):
            c_gpu = eng.train_step_sampled(X, mask, Y, samples, pop)
            c_ref = O.train_step(spec, vals, upd, X, mask, Y=Y, samples=samples, pop=pop)
            assert abs(float(c_gpu) - float(c_ref)) <= 1e-4, (i, float(c_gpu), float(c_ref))
    finally:
        eng.close()


def test_c5_shaped_gru_two_layer_512_step_matches_the_oracle():
    from sbr_b200 import _capi
    rng = np.random.RandomState(32)
    B, T, N = 160, 120, 3000
    spec = O.Spec(n_items=N, cell="GRU", layers=(512, 512), loss="CCE")
    vals = O.init_params(spec, rng, np.float64)
    X, mask = _nested_prefix_batch(rng, B, T, N, n_users=2)
    Y = rng.randint(0, N, B).astype(np.int32)
    pop = rng.uniform(0.5, 2.0, B)
    eng = _capi.Engine(n_items=N, cell="GRU", layers=(512, 512), loss="CCE", max_length=T, batch_size=B)
    try:
        eng.set_all_param_values(vals)
        _compare_step(eng, spec, vals, X, mask, dict(Y=Y, pop=pop), dict(Y=Y, pop=pop),
                      lambda e, X, m, kw: e.train_step_cce(X, m, kw["Y"], kw["pop"]))
        upd = O.Updater("adam", lr=1e-3)
        for i in range(2):
            c_gpu = eng.train_step_cce(X, mask, Y, pop)
            c_ref = O.train_step(spec, vals, upd, X, mask, Y=Y, pop=pop)
            assert abs(float(c_gpu) - float(c_ref)) <= 1e-4, (i, float(c_gpu), float(c_ref))
    finally:
        eng.close()
