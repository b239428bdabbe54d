"""GPU parity: the CUDA path (through the C ABI, via ctypes) against the float64 oracle on the
same seeded inputs.  Tolerances (fp32 device arithmetic vs float64 oracle):
  cost            |d| <= 1e-4 absolute        (north_star: per-step loss within 1e-4)
  gradients       |d| <= 2e-4 * max|g| + 1e-7 per parameter tensor
  K-step training per-step cost within 1e-4; parameters within 2e-4 absolute after K Adam steps
"""
import numpy as np
import pytest

from oracle import sbr_oracle as O
from tests.test_oracle import make_batch

pytestmark = pytest.mark.gpu


def _engine(spec, B, T, **kw):
    from sbr_b200 import _capi
    args = dict(n_items=spec.n_items, cell=spec.cell, layers=spec.layers, loss=spec.loss, max_length=T,
                batch_size=B, embedding=spec.embedding, n_extra_ids=spec.n_extra_ids,
                ids_per_step=spec.ids_per_step, grad_clip=spec.grad_clip, regularization=spec.regularization,
                last_layer_tanh=spec.last_layer_tanh, bidirectional=spec.bidirectional)
    args.update(kw)
    return _capi.Engine(**args)


def _init(spec, seed):
    rng = np.random.RandomState(seed)
    vals = O.init_params(spec, rng, np.float64)
    for v in vals:
        if not v.any():
            v[...] = rng.normal(0, 0.05, size=v.shape)
    return rng, vals


def _batch_kwargs(spec, rng, B, N, X, lens, S=7):
    if spec.loss == "CCE":
        return dict(Y=rng.randint(0, N, size=B), pop=rng.uniform(0.5, 2.0, size=B))
    if spec.loss in O.SAMPLING_LOSSES:
        return dict(Y=rng.randint(0, N, size=B), samples=rng.randint(0, N, size=S), pop=rng.uniform(0.5, 2.0, size=B))
    in_seqs = [list(X[b, :lens[b], 0]) for b in range(B)]
    targets = [list(rng.choice(N, size=1 + b % 3, replace=False)) for b in range(B)]
    Y, W = O.margin_targets(N, in_seqs, targets)
    return dict(Ymat=Y, Wmat=W, _targets=targets, _in_seqs=in_seqs)


def _gpu_step(eng, spec, X, mask, kw, ragged=False):
    if spec.loss == "CCE":
        return eng.train_step_cce(X, mask, kw["Y"], kw["pop"])
    if spec.loss in O.SAMPLING_LOSSES:
        return eng.train_step_sampled(X, mask, kw["Y"], kw["samples"], kw["pop"])
    if not ragged:
        return eng.train_step_margin_dense(X, mask, kw["Ymat"], kw["Wmat"])
    tg, ins = kw["_targets"], kw["_in_seqs"]
    off = np.zeros(len(tg) + 1, np.int32)
    off[1:] = np.cumsum([len(t) for t in tg])
    ids = np.concatenate([np.asarray(t, np.int32) for t in tg])
    N = spec.n_items
    w = np.array([1.0 * len(t) / (N - len(t) - len(s)) for t, s in zip(tg, ins)], np.float32)
    return eng.train_step_margin(X, mask, off, ids, w, None, True)


def _okw(kw):
    return {k: v for k, v in kw.items() if not k.startswith("_")}


def check_grads(spec, B, T, seed=0, K=1, ragged=False, **ekw):
    rng, vals = _init(spec, seed)
    N = spec.n_items
    X, mask, lens = make_batch(rng, B, T, N, K, spec.n_extra_ids)
    kw = _batch_kwargs(spec, rng, B, N, X, lens)
    eng = _engine(spec, B, T, ids_per_step=K, n_samples=7, **ekw)
    try:
        eng.set_all_param_values(vals)
        back = eng.get_all_param_values()
        for a, b in zip(vals, back):
            np.testing.assert_array_equal(a.astype(np.float32), b)
        eng.set_skip_update(True)
        cost = _gpu_step(eng, spec, X, mask, kw, ragged)
        grads = eng.get_all_grads()
        c0, g0 = O.loss_and_grads(spec, vals, X, mask, **_okw(kw))
        assert abs(float(cost) - float(c0)) <= 1e-4, (cost, c0)
        for (name, _), a, b in zip(O.param_names_shapes(spec), g0, grads):
            tol = 2e-4 * np.abs(a).max() + 1e-7
            err = np.abs(a - b).max()
            assert err <= tol, "%s: max err %.3e > tol %.3e (scale %.3e)" % (name, err, tol, np.abs(a).max())
        # second call on the same handle gives the same answer (no state leaks between steps)
        cost2 = _gpu_step(eng, spec, X, mask, kw, ragged)
        assert abs(float(cost2) - float(cost)) <= 1e-6
    finally:
        eng.close()


@pytest.mark.parametrize("cell,H", [("GRU", 100), ("LSTM", 200), ("Vanilla", 48), ("GRU", 20), ("LSTM", 50),
                                    ("GRU", 37), ("LSTM", 264)])
def test_cce_gradients(cell, H):
    spec = O.Spec(n_items=211, cell=cell, layers=(H,), loss="CCE", regularization=0.01)
    check_grads(spec, B=11, T=9)


@pytest.mark.parametrize("rows", ["8", "16"])
@pytest.mark.parametrize("cell,H,B", [("LSTM", 200, 32), ("GRU", 100, 16), ("Vanilla", 48, 48), ("LSTM", 64, 128)])
def test_cce_gradients_tensor_core_wgrad(cell, H, B, rows, monkeypatch):
    """B % 16 == 0: the tcgen05 scans also emit the K-major hi/lo copies and dW_hid comes from wgrad_tc_kernel.
    Both cluster-tile heights (8 and 16 batch rows, SBR_TC_BT) must give the same gradients."""
    monkeypatch.setenv("SBR_TC_BT", rows)
    spec = O.Spec(n_items=211, cell=cell, layers=(H,), loss="CCE")
    check_grads(spec, B=B, T=11, seed=12)


@pytest.mark.parametrize("cell,H,B", [("LSTM", 200, 128), ("GRU", 100, 64), ("Vanilla", 48, 32), ("LSTM", 64, 128)])
def test_cce_gradients_mixed_tiling(cell, H, B, monkeypatch):
    """Mixed tiling: the shortest 16-row group runs as one 16-row tile on a second stream next to the 8-row tiles
    (what a batch of 128 rows does by default on the 15 co-resident cluster slots of a B200)."""
    monkeypatch.setenv("SBR_TC_FORCE_MIXED", "1")
    spec = O.Spec(n_items=211, cell=cell, layers=(H,), loss="CCE")
    check_grads(spec, B=B, T=11, seed=5)


def test_stacked_layers_8_row_tiles():
    """Upper-layer gradient (dhs) streamed by TMA into 8-row tiles, and the mixed tiling on a stacked model."""
    spec = O.Spec(n_items=150, cell="LSTM", layers=(32, 40), loss="CCE", regularization=-0.02)
    check_grads(spec, B=16, T=8, seed=2)
    check_grads(spec, B=128, T=6, seed=3)


@pytest.mark.parametrize("cell,H", [("GRU", 100), ("LSTM", 200), ("Vanilla", 48), ("LSTM", 52)])
def test_cce_gradients_8_row_tiles_ragged_batch(cell, H, monkeypatch):
    """8-row cluster tiles with a batch that is not a multiple of 8 (last tile partly empty)."""
    monkeypatch.setenv("SBR_TC_BT", "8")
    spec = O.Spec(n_items=211, cell=cell, layers=(H,), loss="CCE", regularization=0.01)
    check_grads(spec, B=11, T=9)


@pytest.mark.parametrize("cell", ["GRU", "LSTM"])
def test_cce_gradients_wide_layer_global_weights(cell):
    """H=512: the W_hid slice does not fit in shared memory -> global-memory weight path."""
    spec = O.Spec(n_items=97, cell=cell, layers=(512,), loss="CCE")
    check_grads(spec, B=9, T=5)


@pytest.mark.parametrize("cell,layers", [("GRU", (40, 24)), ("LSTM", (32, 32)), ("LSTM", (24, 40, 16))])
def test_stacked_layers(cell, layers):
    spec = O.Spec(n_items=150, cell=cell, layers=layers, loss="CCE", regularization=-0.02)
    check_grads(spec, B=13, T=8, seed=1)


@pytest.mark.parametrize("cell,layers,B", [("GRU", (40,), 11), ("LSTM", (32,), 11), ("Vanilla", (48,), 16),
                                           ("GRU", (40, 24), 11), ("LSTM", (24, 40, 16), 16), ("LSTM", (200,), 32),
                                           ("GRU", (256,), 64)])
def test_bidirectional_stacks(cell, layers, B):
    """--r_bi (recurrent_layers.py:72-78): every depth is a forward and a backwards layer over the same input, outputs
    concatenated; the final state is [forward state after the last item | backward state after the first item]."""
    spec = O.Spec(n_items=211, cell=cell, layers=layers, loss="CCE", bidirectional=True)
    check_grads(spec, B=B, T=9)


@pytest.mark.parametrize("layers,emb,bi,B", [((48, 32), 0, False, 11), ((40,), 12, False, 16), ((32, 24, 16), 0, True, 16),
                                            ((256, 64), 0, False, 32), ((200, 100), 0, False, 128)])
def test_dense_vanilla_layers(layers, emb, bi, B):
    """Vanilla layers fed by a dense input are Lasagne RecurrentLayers (recurrent_layers.py:98-99): rectifier, parameters
    listed hid_init, W_in_to_hid, b, W_hid_to_hid; layer 0 without an embedding stays the in-tree tanh cell."""
    spec = O.Spec(n_items=211, cell="Vanilla", layers=layers, embedding=emb, loss="CCE", bidirectional=bi)
    check_grads(spec, B=B, T=9)


def test_bidirectional_embedding_and_sampling_loss():
    spec = O.Spec(n_items=150, cell="GRU", layers=(32, 16), loss="CCE", embedding=12, bidirectional=True)
    check_grads(spec, B=8, T=7)
    spec = O.Spec(n_items=150, cell="LSTM", layers=(32,), loss="TOP1", bidirectional=True)
    check_grads(spec, B=8, T=7)
    spec = O.Spec(n_items=101, cell="GRU", layers=(24,), loss="CCE", n_extra_ids=10, ids_per_step=2, bidirectional=True)
    check_grads(spec, B=9, T=6, K=2)


@pytest.mark.parametrize("cell", ["GRU", "LSTM"])
def test_embedding_path(cell):
    spec = O.Spec(n_items=120, cell=cell, layers=(32,), embedding=12, loss="CCE")
    check_grads(spec, B=10, T=7, seed=2)


@pytest.mark.parametrize("cell,H,K", [("GRU", 100, 1), ("LSTM", 48, 2)])
def test_tma_staged_gather(cell, H, K, monkeypatch):
    """SBR_GATHER_TMA=1: stage 1 through cp.async.bulk row copies into a shared-memory ring (one table row per id)."""
    monkeypatch.setenv("SBR_GATHER_TMA", "1")
    spec = O.Spec(n_items=211, cell=cell, layers=(H,), loss="CCE", n_extra_ids=10 if K == 2 else 0, ids_per_step=K)
    check_grads(spec, B=11, T=9, K=K)


def test_rating_feature_ids():
    spec = O.Spec(n_items=90, cell="GRU", layers=(32,), n_extra_ids=10, ids_per_step=2, loss="CCE")
    check_grads(spec, B=10, T=7, seed=3, K=2)


def test_grad_clip_bites():
    spec = O.Spec(n_items=90, cell="GRU", layers=(32,), loss="CCE", grad_clip=1e-3)
    check_grads(spec, B=10, T=7, seed=4)
    spec = O.Spec(n_items=90, cell="LSTM", layers=(32,), loss="CCE", grad_clip=1e-3)
    check_grads(spec, B=10, T=7, seed=4)


@pytest.mark.parametrize("loss", ["BPR", "BPRI", "TOP1", "Blackout"])
def test_sampling_losses(loss):
    spec = O.Spec(n_items=300, cell="LSTM", layers=(48,), loss=loss)
    check_grads(spec, B=12, T=8, seed=5)


def test_sampling_tanh_and_collisions():
    spec = O.Spec(n_items=12, cell="GRU", layers=(16,), loss="TOP1", last_layer_tanh=True)
    check_grads(spec, B=12, T=6, seed=6)   # 12 items, 12 targets + 7 samples: duplicate cells guaranteed


@pytest.mark.parametrize("loss", ["hinge", "logit", "logsig"])
@pytest.mark.parametrize("ragged", [False, True])
def test_margin_losses(loss, ragged):
    spec = O.Spec(n_items=180, cell="LSTM", layers=(40,), loss=loss)
    check_grads(spec, B=9, T=8, seed=7, ragged=ragged)


def test_large_batch_tiles_and_short_rows():
    """B > 148/8 tiles: exercises the 16/32-row tiles; rows of length 1; one empty tile tail."""
    spec = O.Spec(n_items=400, cell="LSTM", layers=(64,), loss="CCE")
    check_grads(spec, B=300, T=12, seed=8)
    spec = O.Spec(n_items=400, cell="GRU", layers=(64,), loss="CCE")
    check_grads(spec, B=700, T=6, seed=9)


@pytest.mark.parametrize("updater", ["adam", "adagrad", "rmsprop", "adadelta", "nesterov"])
def test_training_trajectory(updater):
    """K steps from identical init on identical batches: per-step cost and final parameters."""
    spec = O.Spec(n_items=500, cell="GRU", layers=(100,), loss="CCE")
    B, T, K = 16, 20, 12
    rng, vals = _init(spec, 10)
    lr = {"adam": 1e-3, "adagrad": 0.01, "rmsprop": 1e-3, "adadelta": 1.0, "nesterov": 0.05}[updater]
    eng = _engine(spec, B, T, updater=updater, lr=lr)
    upd = O.Updater(updater, lr=lr)
    try:
        eng.set_all_param_values(vals)
        for step in range(K):
            X, mask, lens = make_batch(rng, B, T, 500)
            Y = rng.randint(0, 500, B)
            pop = rng.uniform(0.5, 2, B)
            c_gpu = eng.train_step_cce(X, mask, Y, pop)
            c_ref = O.train_step(spec, vals, upd, X, mask, Y=Y, pop=pop)
            assert abs(float(c_gpu) - float(c_ref)) <= 1e-4, (step, c_gpu, c_ref)
        # adagrad divides by sqrt(acc + 1e-6): a gradient that is tiny on both sides (|g| << 1e-3) is
        # amplified by lr / 1e-3 = 50 per step, so its fp32-vs-float64 noise shows up 50x larger
        ptol = 1e-3 if updater == "adagrad" else 2e-4
        for (name, _), a, b in zip(O.param_names_shapes(spec), vals, eng.get_all_param_values()):
            assert np.abs(a - b).max() <= ptol, name
    finally:
        eng.close()


def test_scores_and_topk():
    spec = O.Spec(n_items=321, cell="LSTM", layers=(48,), loss="CCE")
    rng, vals = _init(spec, 11)
    B, T = 6, 10
    X, mask, lens = make_batch(rng, B, T, 321)
    eng = _engine(spec, B, T)
    try:
        eng.set_all_param_values(vals)
        s = eng.scores(X, mask)
        ref = O.scores(spec, vals, X, mask)
        np.testing.assert_allclose(s, ref, rtol=2e-4, atol=1e-7)
        excl = [list(np.unique(X[b, :lens[b], 0])) for b in range(B)]
        ids = eng.topk(X, mask, k=10, exclude=excl)
        ex = np.zeros((B, 321))
        for b in range(B):
            ex[b, excl[b]] = 1
        ref_ids = O.top_k(O.test_scores(spec, vals, X, mask, exclude=ex), 10)
        np.testing.assert_array_equal(ids, ref_ids)
        ids_inf = eng.topk(X, mask, k=5, exclude=excl, neg_inf=True)
        for b in range(B):
            assert not set(ids_inf[b]) & set(excl[b])
    finally:
        eng.close()


def test_error_behaviour():
    from sbr_b200 import _capi
    spec = O.Spec(n_items=50, cell="GRU", layers=(16,), loss="CCE")
    eng = _engine(spec, 4, 5)
    try:
        X = np.zeros((4, 5, 1), np.int32)
        mask = np.ones((4, 5), np.float32)
        mask[1, 2] = 0  # hole
        with pytest.raises(_capi.SbrError) as e:
            eng.train_step_cce(X, mask, np.zeros(4, np.int32), np.ones(4, np.float32))
        assert e.value.code == -4
        mask[:] = 1
        X[2, 3, 0] = 50
        with pytest.raises(_capi.SbrError) as e:
            eng.train_step_cce(X, mask, np.zeros(4, np.int32), np.ones(4, np.float32))
        assert e.value.code == -5
        X[2, 3, 0] = 0
        # the handle is still usable after argument errors
        c = eng.train_step_cce(X, mask, np.zeros(4, np.int32), np.ones(4, np.float32))
        assert np.isfinite(c)
        # NaN parameters give a NaN cost, which the python train loop turns into ValueError (rnn_base.py:291)
        vals = eng.get_all_param_values()
        vals[-1][:] = np.nan
        eng.set_all_param_values(vals)
        assert np.isnan(eng.train_step_cce(X, mask, np.zeros(4, np.int32), np.ones(4, np.float32)))
    finally:
        eng.close()
